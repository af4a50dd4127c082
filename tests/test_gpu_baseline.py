"""GPU parity tests on the BASELINE.json configurations themselves (run with -m gpu on a B200).

What VERDICT round 1 asked for: converged SCF energies on the named molecules (1e-8 Eh against the oracle's SCF from
the same starting guess), J/K blocks with PHYSICAL densities (the core-Hamiltonian projector of SURVEY 8d, density 1)
on the three large configurations, O(1) block-diagonal densities on (H2O)64 (fixed-point resolution), the reference's
own screening counts for threshold > 0, and a pure-P (type -1) shell.
Tolerances are north_star's: every J/K element within 1e-10 absolute, SCF energies within 1e-8 Eh.
"""
import os

import numpy as np
import pytest

import scf_harness as H
from chinium_b200.inputs import load_fixture_molecule

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def Int4C2E():
    from chinium_b200 import Int4C2E as cls
    return cls


def _engine(Int4C2E, fb, exx=1.0, thr=-1.0, **kw):
    e = Int4C2E(fb, exx, thr, **kw)
    e.getRepulsionDiag(0); e.getRepulsionLength(0); e.getRepulsionIndices(0); e.getThreadPointers(1, 0); e.CalculateIntegrals(0, 0)
    return e


def _shell_pairs(fb, n_far=10, seed=11):
    """>= 16 shell pairs: the first s|s diagonal block of every distinct element, same-atom pairs of every l, and seeded far pairs."""
    l = np.abs(np.asarray(fb.type))
    s2a = np.asarray(fb.shell2atom)
    pairs = []
    first = {}
    for s in range(fb.nshell):
        first.setdefault((int(l[s]), 0 if s2a[s] == s2a[0] else 1), s)
    for (ll, _), s in sorted(first.items()):
        pairs.append((s, s))                                    # diagonal blocks incl. the tightest s|s
    on0 = [s for s in range(fb.nshell) if s2a[s] == s2a[0]]
    for a in on0[1:4]:
        pairs.append((a, on0[0]))                               # same atom, different shells
    rng = np.random.default_rng(seed)
    while len(pairs) < 16 + n_far - 10 or len(pairs) < 16:
        a, b = (int(x) for x in rng.integers(0, fb.nshell, 2))
        if a != b and (a, b) not in pairs:
            pairs.append((a, b))
    return pairs


def _check_blocks(oracle, fb, pairs, J, Dtot, Ks):
    """Ks: list of (K matrix, its density).  Exact oracle blocks (unscreened) for every pair."""
    worst_j = worst_k = 0.0
    for sa, sb in pairs:
        ia, ib = fb.shell2bf[sa], fb.shell2bf[sb]
        na, nb = fb.nfun[sa], fb.nfun[sb]
        for idx, (K, D) in enumerate(Ks):
            Jb, Kb = oracle.jk_block(fb, Dtot, D, sa, sb)
            if idx == 0:
                worst_j = max(worst_j, float(np.abs(Jb - J[ia:ia + na, ib:ib + nb]).max()))
            worst_k = max(worst_k, float(np.abs(Kb - K[ia:ia + na, ib:ib + nb]).max()))
    return worst_j, worst_k


# ------------------------------------------------------------------------------------------------------------------
# SCF energies on the BASELINE molecules (Restricted/SP.cpp:38-73, Unrestricted/SP.cpp:42-94)
# ------------------------------------------------------------------------------------------------------------------
def test_scf_energy_h2o_rhf(Int4C2E, oracle):
    """examples/h2o.inp as shipped (RHF / cc-pVDZ): engine inside the SCF loop vs the reference's stored-integral path
    restated by the oracle, same core-Hamiltonian guess, same CDIIS."""
    mol, fb = load_fixture_molecule("h2o")
    S, T, V = oracle.one_electron(fb, mol.Z, mol.xyz_bohr)
    enuc = H.nuclear_repulsion(mol.Z, mol.xyz_bohr)
    eng = _engine(Int4C2E, fb)
    E_gpu, D_gpu, _, it_gpu = H.rhf(S, T + V, 5, lambda d, a, b: eng.ContractInts(d, a, b, 1, 0), enuc)
    h = oracle.store_build(fb)
    E_cpu, D_cpu, _, it_cpu = H.rhf(S, T + V, 5, lambda d, a, b: oracle.store_contract(h, fb.nbf, d, a, b), enuc)
    oracle.store_free(h)
    eng.close()
    assert abs(E_gpu - E_cpu) < 1e-8, (E_gpu, E_cpu)
    assert it_gpu == it_cpu
    assert -76.03 < E_gpu < -76.02          # RHF/cc-pVDZ water


def test_scf_energy_bo3h3_hf(Int4C2E, oracle):
    """examples/bo3h3.inp geometry and basis (6-31G**), Hartree-Fock part of the path (EXX = 1): RHF energy vs the oracle."""
    mol, fb = load_fixture_molecule("bo3h3")
    S, T, V = oracle.one_electron(fb, mol.Z, mol.xyz_bohr)
    enuc = H.nuclear_repulsion(mol.Z, mol.xyz_bohr)
    nocc = mol.nelec // 2
    eng = _engine(Int4C2E, fb)
    E_gpu, *_ = H.rhf(S, T + V, nocc, lambda d, a, b: eng.ContractInts(d, a, b, 1, 0), enuc)
    h = oracle.store_build(fb)
    E_cpu, *_ = H.rhf(S, T + V, nocc, lambda d, a, b: oracle.store_contract(h, fb.nbf, d, a, b), enuc)
    oracle.store_free(h)
    eng.close()
    assert abs(E_gpu - E_cpu) < 1e-8, (E_gpu, E_cpu)


def test_scf_energy_fe4s4_uhf(Int4C2E, oracle):
    """examples/fe4s4.inp (UHF, charge +2, 2S+1 = 19, 6-31G*; J + Ka + Kb with d and f shells) from the core-Hamiltonian
    guess with the reference's CDIIS settings (20 vectors, max|commutator| < 1e-6, 300 iterations:
    Unrestricted/SP.cpp:99-100).  This open-shell cluster needs ~180 iterations.  The engine runs the whole SCF; the
    oracle's stored-integral path (the reference's algorithm, 3.5 GiB, ~1 s per iteration on the host) runs
      (1) the first 8 iterations from the same guess: energies must agree iteration by iteration, and
      (2) the SCF restarted at the engine's converged density until ITS convergence test passes: the two converged
          energies must agree to 1e-8 Eh (a fixed point of one is a fixed point of the other)."""
    mol, fb = load_fixture_molecule("fe4s4")
    S, T, V = oracle.one_electron(fb, mol.Z, mol.xyz_bohr)
    enuc = H.nuclear_repulsion(mol.Z, mol.xyz_bohr)
    na, nb = mol.nalpha_nbeta
    assert (na, nb) == (92, 74)
    eng = _engine(Int4C2E, fb)
    kw = dict(tol=1e-6, diis_space=20, max_iter=300)
    tr_gpu = []
    E_gpu, (Da, Db), _, it_gpu = H.uhf(S, T + V, na, nb, lambda d, a, b: eng.ContractInts(d, a, b, 1, 0), enuc, trace=tr_gpu, **kw)
    eng.close()
    h = oracle.store_build(fb)
    jk_cpu = lambda d, a, b: oracle.store_contract(h, fb.nbf, d, a, b)
    tr_cpu = []
    H.uhf(S, T + V, na, nb, jk_cpu, enuc, trace=tr_cpu, raise_on_fail=False, **dict(kw, max_iter=8))
    E_cpu, _, _, it_cpu = H.uhf(S, T + V, na, nb, jk_cpu, enuc, D0=(Da, Db), **kw)
    oracle.store_free(h)
    print("fe4s4 UHF: E_gpu %.10f after %d iterations, oracle restart %.10f after %d" % (E_gpu, it_gpu, E_cpu, it_cpu))
    assert max(abs(a - b) for a, b in zip(tr_gpu, tr_cpu)) < 1e-8, [a - b for a, b in zip(tr_gpu, tr_cpu)]
    assert it_cpu <= 3
    assert abs(E_gpu - E_cpu) < 1e-8, (E_gpu, E_cpu)
    assert abs(E_gpu - (-6638.57444219)) < 1e-6      # value the oracle's own SCF from the core guess reaches (recorded while writing this test)


def test_rohf_three_density_energy(Int4C2E, oracle):
    """The three-density call of the restricted open-shell / two-determinant drivers (Universal.cpp:33-47):
    (J, Kd, Ka, Kb) = ContractInts(Dd, Da, Db), F_0 = H + J - Kd - Ka/2 - Kb/2, F_1 = H + J - Kd - Ka + c Kb,
    F_2 = H + J - Kd - Kb + c Ka, E = sum_t occ_t/2 D_t o (H + F_t), occ = {2, 1, 1} (Universal.h:27).  Triplet CH2 with
    core-guess orbitals split into 3 doubly occupied, one alpha-only and one beta-only orbital; a few steepest-descent
    style re-diagonalisations of F_0 move the densities so that more than one point is compared."""
    mol, fb = load_fixture_molecule("ch2")
    S, T, V = oracle.one_electron(fb, mol.Z, mol.xyz_bohr)
    Hc = T + V
    enuc = H.nuclear_repulsion(mol.Z, mol.xyz_bohr)
    eng = _engine(Int4C2E, fb)
    coupling = 1.0

    def energy_and_fock(jk, Dd, Da, Db):
        J, Kd, Ka, Kb = jk(Dd, Da, Db)
        F = [Hc + J - Kd - 0.5 * Ka - 0.5 * Kb, Hc + J - Kd - Ka + coupling * Kb, Hc + J - Kd - Kb + coupling * Ka]
        E = sum(0.5 * occ * np.sum(D * (Hc + Ft)) for occ, D, Ft in zip((2, 1, 1), (Dd, Da, Db), F))
        return E + enuc, F

    from scipy.linalg import eigh
    Fcur = Hc
    for step in range(4):
        e, C_ = eigh(Fcur, S)
        Dd = C_[:, :3] @ C_[:, :3].T
        Da = C_[:, 3:4] @ C_[:, 3:4].T
        Db = C_[:, 4:5] @ C_[:, 4:5].T
        E_gpu, F_gpu = energy_and_fock(lambda d, a, b: eng.ContractInts(d, a, b, 1, 0), Dd, Da, Db)
        E_cpu, F_cpu = energy_and_fock(lambda d, a, b: oracle.direct_jk(fb, d, a, b)[:4], Dd, Da, Db)
        assert abs(E_gpu - E_cpu) < 1e-8, (step, E_gpu, E_cpu)
        for Fg, Fc in zip(F_gpu, F_cpu):
            assert np.abs(Fg - Fc).max() < 4 * TOL
        Fcur = F_cpu[0]
    eng.close()


# ------------------------------------------------------------------------------------------------------------------
# physical densities on the large configurations: exact oracle blocks (SURVEY 8d density 1 = core-Hamiltonian projector)
# ------------------------------------------------------------------------------------------------------------------
def test_c18_core_density_blocks(Int4C2E, oracle):
    mol, fb = load_fixture_molecule("c18")
    S, T, V = oracle.one_electron(fb, mol.Z, mol.xyz_bohr)
    D = H.core_density(S, T + V, mol.nelec // 2)
    assert 0.5 < np.abs(D).max() < 3.0                       # O(1) entries, unlike the 1/nbf stress density
    eng = _engine(Int4C2E, fb)
    J, K, _, _ = eng.ContractInts(D, None, None, 1, 0)
    st = eng.stats
    pairs = _shell_pairs(fb)
    assert len(pairs) >= 16
    ej, ek = _check_blocks(oracle, fb, pairs, J, 2 * D, [(K, D)])
    eng.close()
    assert ej < TOL and ek < TOL, (ej, ek, st["j_two_limb_last"], st["j_rounding_estimate_last"])


def test_fe4s4_core_density_blocks(Int4C2E, oracle):
    mol, fb = load_fixture_molecule("fe4s4")
    S, T, V = oracle.one_electron(fb, mol.Z, mol.xyz_bohr)
    na, nb = mol.nalpha_nbeta
    Da, Db = H.core_density(S, T + V, na), H.core_density(S, T + V, nb)
    eng = _engine(Int4C2E, fb)
    J, _, Ka, Kb = eng.ContractInts(None, Da, Db, 1, 0)
    pairs = _shell_pairs(fb)
    assert len(pairs) >= 16
    ej, ek = _check_blocks(oracle, fb, pairs, J, Da + Db, [(Ka, Da), (Kb, Db)])
    eng.close()
    assert ej < TOL and ek < TOL, (ej, ek)


def _h2o64_fragment_density(oracle, mol, fb):
    """O(1) block-diagonal density of (H2O)64: the core-Hamiltonian projector (5 occupied orbitals) of every water
    molecule on its own, placed on the molecule's diagonal block (a superposition of fragment densities)."""
    import copy
    n = fb.nbf
    D = np.zeros((n, n))
    s2a = np.asarray(fb.shell2atom)
    per = fb.nshell // 64
    for k in range(64):
        sh = np.arange(k * per, (k + 1) * per)
        atoms = np.unique(s2a[sh])
        sub = copy.copy(fb)
        sub = type(fb)(type=fb.type[sh], nprim=fb.nprim[sh], prim_offset=fb.prim_offset[sh], exps=fb.exps, coefs_raw=fb.coefs_raw,
                       coefs_normalized=fb.coefs_normalized, center_xyz=np.ascontiguousarray(np.asarray(fb.center_xyz).reshape(-1, 3)[sh]),
                       shell2atom=(s2a[sh] - atoms[0]).astype(np.int32))
        S, T, V = oracle.one_electron(sub, mol.Z[atoms], mol.xyz_bohr[atoms])
        Dk = H.core_density(S, T + V, 5)
        o = fb.shell2bf[sh[0]]
        D[o:o + sub.nbf, o:o + sub.nbf] = Dk
    return D


def test_h2o64_block_diagonal_density_resolution(Int4C2E, oracle):
    """(H2O)64 / def2-TZVP with an O(1) block-diagonal density (fragment core-Hamiltonian projectors): sampled J/K blocks,
    including the tight s|s diagonal blocks where ~10^6 fixed-point adds pile up, within 1e-10 of the UNSCREENED oracle.
    Round 1 measured 2.2e-10 here (one global scale bounding the sum of absolute values); now the scale bounds the final
    value only (sums are exact modulo 2^64) and J gets a low limb when the rounding estimate asks for it."""
    mol, fb = load_fixture_molecule("h2o64")
    D = _h2o64_fragment_density(oracle, mol, fb)
    assert 0.5 < np.abs(D).max() < 3.0
    eng = _engine(Int4C2E, fb, thr=1e-13)
    J, K, _, _ = eng.ContractInts(D, None, None, 1, 0)
    st = eng.stats
    per = fb.nshell // 64
    pairs = [(0, 0), (1, 1), (2, 0), (per * 20, per * 20), (per * 20 + 3, per * 20 + 3), (per * 63 + 8, per * 63 + 8), (5, 5), (9, 9),
             (4, 7 * per + 2), (per * 20 + 9, per * 20 + 3), (per * 63 + 8, 5), (per * 33 + 12, per * 32 + 1), (per * 10, per * 11),
             (per * 40 + 6, per * 41 + 6), (per * 5 + 16, per * 5 + 2), (per * 50 + 13, per * 12 + 13)]
    ej, ek = _check_blocks(oracle, fb, pairs, J, 2 * D, [(K, D)])
    print("h2o64 block-diagonal O(1) density: max|dJ| %.2e max|dK| %.2e, two-limb J %d (estimate %.1e), |J|max %.1f"
          % (ej, ek, st["j_two_limb_last"], st["j_rounding_estimate_last"], np.abs(J).max()))
    # forced single limb: the tightened scale alone
    eng1 = _engine(Int4C2E, fb, thr=1e-13, j_two_limb=-1)
    J1, K1, _, _ = eng1.ContractInts(D, None, None, 1, 0)
    ej1, ek1 = _check_blocks(oracle, fb, pairs[:6], J1, 2 * D, [(K1, D)])
    print("   single limb: max|dJ| %.2e max|dK| %.2e" % (ej1, ek1))
    eng.close(); eng1.close()
    assert ej < TOL and ek < TOL, (ej, ek)
    assert ej1 < 2 * TOL


def test_two_limb_j_is_bit_stable_and_matches(Int4C2E, oracle):
    """j_two_limb = 1 (forced) on a small molecule: same J within rounding as the single-limb build, bit-identical run
    to run and across a 2-way partition (both limbs are integer sums)."""
    import torch
    mol, fb = load_fixture_molecule("bo3h3")
    n = fb.nbf
    D = H.random_symmetric_density(n, 0) * n
    e1 = _engine(Int4C2E, fb, j_two_limb=-1)
    e2 = _engine(Int4C2E, fb, j_two_limb=1)
    J1, K1, _, _ = e1.ContractInts(D, None, None, 1, 0)
    J2, K2, _, _ = e2.ContractInts(D, None, None, 1, 0)
    J2b, _, _, _ = e2.ContractInts(D, None, None, 1, 0)
    assert e2.stats["j_two_limb_last"] == 1 and e1.stats["j_two_limb_last"] == 0
    Jo, Ko, _, _, _ = oracle.direct_jk(fb, D)
    assert np.abs(J1 - Jo).max() < TOL and np.abs(J2 - Jo).max() < TOL and (K1 == K2).all()
    assert (J2 == J2b).all()
    e1.close(); e2.close()
    Dt = torch.from_numpy(D).cuda()
    acc_sum = None
    engs = []
    for r in range(2):
        e = Int4C2E(fb, 1.0, -1.0, rank=r, world_size=2, j_two_limb=1)
        acc = torch.zeros(e.acc_len(1), dtype=torch.int64, device="cuda")
        e.accumulate_device(Dt.data_ptr(), None, None, acc.data_ptr(), None)
        torch.cuda.synchronize()
        if acc_sum is None:
            acc_sum = acc
        else:
            acc_sum[: e.acc_reduce_len(1)] += acc[: e.acc_reduce_len(1)]
        engs.append(e)
    J = torch.empty((n, n), dtype=torch.float64, device="cuda"); K = torch.empty_like(J)
    engs[0].finalize_device(acc_sum.data_ptr(), (1, 0, 0), J.data_ptr(), K.data_ptr(), None, None, None)
    torch.cuda.synchronize()
    assert (J.cpu().numpy() == J2).all()
    for e in engs:
        e.close()


# ------------------------------------------------------------------------------------------------------------------
# the reference's own counts and shell types
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,thr", [("h2o", -1.0), ("h2o", 1e-6), ("h2o", 1e-3), ("hf_tz", 1e-5), ("bo3h3", -1.0), ("bo3h3", 1e-8),
                                      ("bo3h3", 1e-4)])
def test_reference_counts_match_restatement(Int4C2E, oracle, name, thr):
    """RepulsionLength / ShellQuartetLength as the reference computes them (getRepulsionLength, Int4C2E.cpp:79-128: its
    loop nest s4 <= max(s2,s3), function-level uniqueness predicate and Schwarz test on Diag1212) against the oracle's
    literal restatement -- for the reference's default threshold -1 and for threshold > 0."""
    mol, fb = load_fixture_molecule(name)
    eng = _engine(Int4C2E, fb, thr=thr)
    D = H.random_symmetric_density(fb.nbf, 0)
    *_, counts = oracle.reference_jk(fb, D, threshold=thr)
    assert (eng.RepulsionLength, eng.ShellQuartetLength) == tuple(counts), (eng.RepulsionLength, eng.ShellQuartetLength, counts)
    eng.close()


def test_pure_p_shell(Int4C2E, oracle):
    """Type -1 (pure P, order y, z, x: src/Grid/AO/PureP.hpp:1-3, Int2C1E.cpp:256-259) next to the default Cartesian +1."""
    import copy
    mol, fb = load_fixture_molecule("h2o")
    fbp = copy.deepcopy(fb)
    fbp.type = np.where(np.asarray(fb.type) == 1, -1, fb.type).astype(np.int32)
    fbp.__post_init__()
    assert fbp.nbf == fb.nbf and (np.asarray(fbp.type) == -1).sum() == 4
    D = H.random_symmetric_density(fb.nbf, 3)
    eng = _engine(Int4C2E, fbp)
    J, K, _, _ = eng.ContractInts(D, None, None, 1, 0)
    Jo, Ko, _, _, _ = oracle.direct_jk(fbp, D)
    assert np.abs(J - Jo).max() < TOL and np.abs(K - Ko).max() < TOL
    # and it really is a permutation of the Cartesian result: p functions (x,y,z) -> (y,z,x)
    perm = np.arange(fb.nbf)
    for s in range(fb.nshell):
        if fb.type[s] == 1:
            o = fb.shell2bf[s]
            perm[o:o + 3] = [o + 1, o + 2, o]
    engc = _engine(Int4C2E, fb)
    Dc = np.zeros_like(D); Dc[np.ix_(perm, perm)] = D
    Jc, Kc, _, _ = engc.ContractInts(Dc, None, None, 1, 0)
    assert np.abs(Jc[np.ix_(perm, perm)] - J).max() < TOL and np.abs(Kc[np.ix_(perm, perm)] - K).max() < TOL
    eng.close(); engc.close()


def test_multi_density_rejects_bad_shapes(Int4C2E):
    """ContractInts([D...]) validates every matrix like the single-density call."""
    from chinium_b200 import FockEngineError
    mol, fb = load_fixture_molecule("h2o")
    n = fb.nbf
    eng = _engine(Int4C2E, fb)
    with pytest.raises(FockEngineError):
        eng.ContractInts([np.eye(n), np.eye(n - 1)], 1, 0)
    with pytest.raises(FockEngineError):
        eng.ContractInts([np.eye(n + 2)], 1, 0)
    with pytest.raises(FockEngineError):
        eng.ContractInts([np.eye(n), None], 1, 0)
    assert eng.ContractInts([], 1, 0) == []
    Gs = eng.ContractInts([np.eye(n)], nthreads=4, output=0)
    assert len(Gs) == 1 and Gs[0].shape == (n, n)
    eng.close()


# ------------------------------------------------------------------------------------------------------------------
# one-electron integrals on the device and a device-resident RHF iteration
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["h2o", "hf_tz", "bo3h3", "fe4s4"])
def test_one_electron_integrals(Int4C2E, oracle, name):
    """S, T, V from the device kernels (oneint.cuh; Int2C1E.cpp:18-67, :313-333) against the oracle's McMurchie-Davidson
    one-electron integrals: s..f shells, contracted and uncontracted, 2 to 8 nuclei."""
    from chinium_b200 import Int2C1E
    mol, fb = load_fixture_molecule(name)
    eng = Int4C2E(fb, 1.0, -1.0)
    i1 = Int2C1E(fb, mol.Z, mol.xyz_bohr, engine=eng)
    i1.CalculateIntegrals(0, 0)
    S, T, V = oracle.one_electron(fb, mol.Z, mol.xyz_bohr)
    assert np.abs(i1.Overlap - S).max() < 1e-12
    assert np.abs(i1.Kinetic - T).max() < 1e-10 * max(1.0, np.abs(T).max() / 100)
    assert np.abs(i1.Nuclear - V).max() < 1e-10 * max(1.0, np.abs(V).max() / 100)
    for M in (i1.Overlap, i1.Kinetic, i1.Nuclear):
        assert np.abs(M - M.T).max() == 0.0
    assert np.abs(np.diag(i1.Overlap) - 1.0).max() < 1e-12
    eng.close()


def test_device_resident_scf(Int4C2E, oracle):
    """RHF with D, J, K, F, S, H resident on the device (tests/scf_device.py): the energies of examples/h2o.inp and of the
    CH3ClF- golden case equal the oracle's / Chinium's to 1e-8 Eh."""
    import scf_device
    for name, golden in (("h2o", None), ("sn2", -598.514802895)):
        mol, fb = load_fixture_molecule(name)
        enuc = H.nuclear_repulsion(mol.Z, mol.xyz_bohr)
        nocc = mol.nelec // 2
        eng = Int4C2E(fb, 1.0, -1.0)
        E_dev, D_dev, it = scf_device.rhf_device(eng, mol.Z, mol.xyz_bohr, nocc, enuc)
        eng.close()
        S, T, V = oracle.one_electron(fb, mol.Z, mol.xyz_bohr)
        h = oracle.store_build(fb)
        E_cpu, D_cpu, _, it_cpu = H.rhf(S, T + V, nocc, lambda d, a, b: oracle.store_contract(h, fb.nbf, d, a, b), enuc)
        oracle.store_free(h)
        assert abs(E_dev - E_cpu) < 1e-8, (name, E_dev, E_cpu)
        assert np.abs(D_dev.cpu().numpy() - D_cpu).max() < 1e-6
        if golden is not None:
            assert abs(E_dev - golden) < 1e-7


@pytest.mark.parametrize("name,exx", [("h2o", 1.0), ("hf_tz", 0.7), ("bo3h3", 0.0)])
def test_contract_grads_matrix_form(Int4C2E, oracle, name, exx):
    """Int4C2E::ContractGrads(D, output) (Int4C2E.cpp:766-790): the 3*natoms matrices G^(atom,xyz)[D] against the oracle's
    literal getRepulsion1 restatement (:312-408), s..f shells; consistency with the fused vector form; GradCache."""
    mol, fb = load_fixture_molecule(name)
    n = fb.nbf
    D, D1 = H.random_symmetric_density(n, 31) * n, H.random_symmetric_density(n, 32) * n
    eng = _engine(Int4C2E, fb, exx=exx)
    Gs = eng.ContractGrads(D, 0)
    ref = oracle.grad_matrices(fb, D, exx)
    assert len(Gs) == len(ref) == 3 * (int(np.max(fb.shell2atom)) + 1)
    scale = max(1.0, max(np.abs(r).max() for r in ref))
    for G, R in zip(Gs, ref):
        assert np.abs(G - R).max() < 1e-9 * scale
        assert np.abs(G - G.T).max() == 0.0
    v = eng.ContractGrads(D1, D, 0)                          # fused vector form == D1 o matrices
    assert np.abs(v - np.array([np.sum(D1 * G) for G in Gs])).max() < 1e-9 * max(1.0, np.abs(v).max())
    vs = eng.ContractGrads([D1, D], D, 0)
    assert np.abs(vs[0] - v).max() < 1e-9 * max(1.0, np.abs(v).max())
    assert eng.ContractGrads(D, 0) is Gs                     # served from GradCache
    eng.close()
