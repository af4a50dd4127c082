"""GPU parity tests (run with -m gpu on a B200): the CUDA engine, called through the C ABI, against the CPU
oracle on the same seeded inputs.  Tolerances are north_star's: every J/K element within 1e-10 absolute,
converged SCF energies within 1e-8 Eh."""
import os

import numpy as np
import pytest

import scf_harness as H
from chinium_b200.inputs import load_fixture_molecule

pytestmark = pytest.mark.gpu
TOL = 1e-10
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def Int4C2E():
    from chinium_b200 import Int4C2E as cls
    from chinium_b200.fock import device_info
    info = device_info()
    assert info["cc"][0] == 10, info
    return cls


def _engine(Int4C2E, fb, exx=1.0, thr=-1.0, **kw):
    e = Int4C2E(fb, exx, thr, **kw)
    # the reference's setup sequence (SelfConsistentField.cpp:49-53)
    e.getRepulsionDiag(0); e.getRepulsionLength(0); e.getRepulsionIndices(0); e.getThreadPointers(1, 0); e.CalculateIntegrals(0, 0)
    return e


@pytest.mark.parametrize("name", ["h2o", "hf_tz", "bo3h3"])
def test_jk_parity_rhf_uhf_rohf(Int4C2E, oracle, name):
    mol, fb = load_fixture_molecule(name)
    n = fb.nbf
    Dd, Da, Db = (H.random_symmetric_density(n, s) for s in (0, 1, 2))
    eng = _engine(Int4C2E, fb)
    for dens in ((Dd, None, None), (None, Da, Db), (Dd, Da, Db)):      # RHF, UHF, ROHF-type calls
        got = eng.ContractInts(*dens, 1, 0)
        ref = oracle.direct_jk(fb, *dens)[:4]
        for g, r in zip(got, ref):
            if r is None:
                assert np.abs(g).max() == 0.0          # absent density -> zero matrix (Int4C2E.cpp:616-619)
            else:
                assert np.abs(g - r).max() < TOL
        assert np.abs(got[0] - got[0].T).max() == 0.0  # exactly symmetric outputs
    eng.close()


def test_golden_fixture_and_reference_stored_path(Int4C2E, oracle):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "jk_golden.npz"))
    for name in ("h2o", "hf_tz"):
        mol, fb = load_fixture_molecule(name)
        eng = _engine(Int4C2E, fb)
        D, Da, Db = (H.random_symmetric_density(fb.nbf, s) for s in (0, 1, 2))
        J, K, _, _ = eng.ContractInts(D, None, None, 1, 0)
        assert np.abs(J - g[name + "_J"]).max() < TOL and np.abs(K - g[name + "_K"]).max() < TOL
        J2, _, Ka, Kb = eng.ContractInts(None, Da, Db, 1, 0)
        assert np.abs(J2 - g[name + "_J_ab"]).max() < TOL
        assert np.abs(Ka - g[name + "_Ka"]).max() < TOL and np.abs(Kb - g[name + "_Kb"]).max() < TOL
        assert eng.RepulsionLength == int(g[name + "_counts"][0])    # the reference's RepulsionLength
        n = fb.nbf                                                    # gradient fixture (tests/golden/make_jk_golden.py)
        eng.EXX = 0.7
        gr = eng.ContractGrads(H.random_symmetric_density(n, 21) * n, H.random_symmetric_density(n, 22) * n, 0)
        ref = g[name + "_grad"]
        assert gr.shape == ref.shape and np.abs(gr - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())
        eng.close()


def test_repulsion_diag(Int4C2E, oracle):
    mol, fb = load_fixture_molecule("hf_tz")
    eng = _engine(Int4C2E, fb)
    assert np.abs(eng.RepulsionDiags[0] - oracle.repulsion_diag(fb)).max() < TOL
    eng.close()


def test_exx_semantics(Int4C2E, oracle):
    mol, fb = load_fixture_molecule("bo3h3")
    D = H.random_symmetric_density(fb.nbf, 0)
    eng = _engine(Int4C2E, fb, exx=1.0)
    eng.EXX = 0.2                                   # overwritten after construction, read at contract time
    J, K, _, _ = eng.ContractInts(D, None, None, 1, 0)
    Jo, Ko, _, _, _ = oracle.direct_jk(fb, D, exx=0.2)
    assert np.abs(J - Jo).max() < TOL and np.abs(K - Ko).max() < TOL
    eng.EXX = 0.0                                   # pure DFT: J only, K zeros (Int4C2E.cpp:638)
    J0, K0, _, _ = eng.ContractInts(D, None, None, 1, 0)
    assert np.abs(J0 - Jo).max() < TOL and np.abs(K0).max() == 0.0
    eng.close()


def test_edge_cases(Int4C2E):
    from chinium_b200 import FockEngineError
    mol, fb = load_fixture_molecule("h2o")
    eng = _engine(Int4C2E, fb)
    n = fb.nbf
    # 0x0 matrices mean "absent" (Int4C2E.cpp:608-610)
    J, Kd, Ka, Kb = eng.ContractInts(np.eye(n), np.zeros((0, 0)), np.zeros((0, 0)), 1, 0)
    assert np.abs(Ka).max() == 0 and np.abs(Kb).max() == 0 and np.abs(J).max() > 0
    # zero density -> zero J/K
    J, Kd, _, _ = eng.ContractInts(np.zeros((n, n)), None, None, 1, 0)
    assert np.abs(J).max() == 0 and np.abs(Kd).max() == 0
    with pytest.raises(FockEngineError):
        eng.ContractInts(np.eye(n + 1), None, None, 1, 0)
    with pytest.raises(FockEngineError):
        eng.ContractInts(None, None, None, 1, 0)
    # non-symmetric input is symmetrised defensively (the reference assumes symmetric D)
    A = np.random.default_rng(5).uniform(-1, 1, (n, n))
    J1, K1, _, _ = eng.ContractInts(A, None, None, 1, 0)
    J2, K2, _, _ = eng.ContractInts(0.5 * (A + A.T), None, None, 1, 0)
    assert np.abs(J1 - J2).max() < 1e-12 and np.abs(K1 - K2).max() < 1e-12
    # huge density: the fixed-point scale follows the density, so relative accuracy is kept
    Jh, Kh, _, _ = eng.ContractInts(1e6 * np.eye(n), None, None, 1, 0)
    J1, K1, _, _ = eng.ContractInts(np.eye(n), None, None, 1, 0)
    assert np.abs(Jh - 1e6 * J1).max() < 1e-6 and np.abs(Kh - 1e6 * K1).max() < 1e-6
    with pytest.raises(FockEngineError):
        eng.ContractInts(np.full((n, n), np.nan), None, None, 1, 0)
    eng.close()


def test_size_independent_properties_c18(Int4C2E):
    """Full-size BASELINE config: linearity, symmetry, bit-stability, partition independence."""
    mol, fb = load_fixture_molecule("c18")
    n = fb.nbf
    D1, D2 = H.random_symmetric_density(n, 0), H.random_symmetric_density(n, 7)
    full = _engine(Int4C2E, fb, pair_cutoff=1e-300)     # nothing dropped: the reference's unscreened counts
    st = full.stats
    npair_bf = n * (n + 1) // 2
    # unscreened RepulsionLength = N(N+1)/2 with N = nbf(nbf+1)/2 (SURVEY 8d prints 10 668 085 485: a transcription slip)
    assert st["canonical_quartets"] == 132690195 and st["unique_integrals"] == npair_bf * (npair_bf + 1) // 2 == 10668295485
    assert abs(st["flops_alg_jk"][1] / 2.365e12 - 1) < 5e-3
    full.close()
    eng = _engine(Int4C2E, fb)
    assert eng.stats["canonical_quartets"] <= 132690195
    J1, K1, _, _ = eng.ContractInts(D1, None, None, 1, 0)
    J1b, K1b, _, _ = eng.ContractInts(D1, None, None, 1, 0)
    assert (J1 == J1b).all() and (K1 == K1b).all()          # bit-stable run to run
    J2, K2, _, _ = eng.ContractInts(D2, None, None, 1, 0)
    J3, K3, _, _ = eng.ContractInts(0.5 * D1 - 2.0 * D2, None, None, 1, 0)
    assert np.abs(J3 - (0.5 * J1 - 2.0 * J2)).max() < TOL   # linearity
    assert np.abs(K3 - (0.5 * K1 - 2.0 * K2)).max() < TOL
    assert np.abs(J1 - J1.T).max() == 0 and np.abs(K1 - K1.T).max() == 0
    # energy-like checksum: sum D1 o J[D2] == sum D2 o J[D1]
    assert abs(np.sum(D1 * J2) - np.sum(D2 * J1)) < 1e-9
    assert abs(np.sum(D1 * K2) - np.sum(D2 * K1)) < 1e-9
    eng.close()


def test_sampled_blocks_c18(Int4C2E, oracle):
    """c18 / cc-pVTZ at full size: exact oracle J and K blocks for a sample of shell pairs (s..f)."""
    mol, fb = load_fixture_molecule("c18")
    n = fb.nbf
    D = H.random_symmetric_density(n, 0)
    eng = _engine(Int4C2E, fb)
    J, K, _, _ = eng.ContractInts(D, None, None, 1, 0)
    # shells of atom 0: s s s s p p p d d f ; pick pairs covering the classes, on near and far atoms
    per_atom = fb.nshell // 18
    for sa, sb in ((0, 0), (per_atom * 9 + 4, 2), (8, per_atom * 5 + 7), (9, per_atom * 17 + 9), (per_atom * 3 + 9, 5)):
        Jb, Kb = oracle.jk_block(fb, 2 * D, D, sa, sb)
        ia, ib = fb.shell2bf[sa], fb.shell2bf[sb]
        assert np.abs(Jb - J[ia:ia + fb.nfun[sa], ib:ib + fb.nfun[sb]]).max() < TOL, (sa, sb)
        assert np.abs(Kb - K[ia:ia + fb.nfun[sa], ib:ib + fb.nfun[sb]]).max() < TOL, (sa, sb)
    eng.close()


def test_partition_independence_bitwise(Int4C2E):
    """Two-rank partition summed in integers == single-rank result, bit for bit (multi-GPU path)."""
    import torch
    mol, fb = load_fixture_molecule("bo3h3")
    n = fb.nbf
    D = torch.from_numpy(H.random_symmetric_density(n, 0)).cuda()
    outs = []
    for world in (1, 2, 3):
        acc_sum = None
        engs = []
        for r in range(world):
            e = Int4C2E(fb, 1.0, -1.0, rank=r, world_size=world)
            acc = torch.zeros(e.acc_len(1), dtype=torch.int64, device="cuda")
            e.accumulate_device(D.data_ptr(), None, None, acc.data_ptr(), None)
            torch.cuda.synchronize()
            if acc_sum is None:
                acc_sum = acc                          # keeps this build's scale tail (identical on every rank)
            else:
                acc_sum[: e.acc_reduce_len(1)] += acc[: e.acc_reduce_len(1)]   # the all-reduce: integer sum of the payload
            engs.append(e)
        J = torch.empty((n, n), dtype=torch.float64, device="cuda"); K = torch.empty_like(J)
        engs[0].finalize_device(acc_sum.data_ptr(), (1, 0, 0), J.data_ptr(), K.data_ptr(), None, None, None)
        torch.cuda.synchronize()
        outs.append((J.cpu().numpy().copy(), K.cpu().numpy().copy()))
        for e in engs:
            e.close()
    for J, K in outs[1:]:
        assert (J == outs[0][0]).all() and (K == outs[0][1]).all()


def test_scf_energy_sn2_golden(Int4C2E, oracle):
    """RHF/cc-pVDZ CH3ClF-: the engine inside the SCF loop reproduces Chinium's recorded energy
    (tools/sn2/sn2.cnm.log:204) and the oracle's SCF energy from the same guess to 1e-8 Eh."""
    mol, fb = load_fixture_molecule("sn2")
    S, T, V = oracle.one_electron(fb, mol.Z, mol.xyz_bohr)
    enuc = H.nuclear_repulsion(mol.Z, mol.xyz_bohr)
    eng = _engine(Int4C2E, fb)
    E_gpu, *_ = H.rhf(S, T + V, 18, lambda d, a, b: eng.ContractInts(d, a, b, 1, 0), enuc)
    h = oracle.store_build(fb)
    E_cpu, *_ = H.rhf(S, T + V, 18, lambda d, a, b: oracle.store_contract(h, fb.nbf, d, a, b), enuc)
    oracle.store_free(h)
    eng.close()
    assert abs(E_gpu - E_cpu) < 1e-8
    assert abs(E_gpu - (-598.514802895)) < 1e-7


def test_scf_energy_h2o_uhf_triplet(Int4C2E, oracle):
    """Open-shell path (J, Ka, Kb): triplet CH2 / cc-pVDZ (examples/ch2.inp) UHF energy vs the oracle's."""
    mol, fb = load_fixture_molecule("ch2")
    S, T, V = oracle.one_electron(fb, mol.Z, mol.xyz_bohr)
    enuc = H.nuclear_repulsion(mol.Z, mol.xyz_bohr)
    na, nb = mol.nalpha_nbeta
    eng = _engine(Int4C2E, fb)
    E_gpu, *_ = H.uhf(S, T + V, na, nb, lambda d, a, b: eng.ContractInts(d, a, b, 1, 0), enuc)
    E_cpu, *_ = H.uhf(S, T + V, na, nb, lambda d, a, b: oracle.direct_jk(fb, d, a, b)[:4], enuc)
    eng.close()
    assert abs(E_gpu - E_cpu) < 1e-8


@pytest.mark.parametrize("name,nmat,exx", [("h2o", 3, 0.5), ("bo3h3", 4, 0.2), ("hf_tz", 2, 1.0), ("h2o", 5, 0.0)])
def test_multi_density(Int4C2E, oracle, name, nmat, exx):
    """ContractInts(std::vector<EigenMatrix>&) (Int4C2E.cpp:685-745): G_k = J[2 D_k] - EXX K[D_k].  Batches of three
    densities share one pass over the integrals (QuartetTask::nj); 4 and 5 matrices exercise the 3+1 / 3+2 split,
    EXX = 0 the Coulomb-only form."""
    mol, fb = load_fixture_molecule(name)
    Ds = [H.random_symmetric_density(fb.nbf, 30 + s) * (1.0 + s) for s in range(nmat)]
    eng = _engine(Int4C2E, fb, exx=exx)
    Gs = eng.ContractInts(Ds, 1, 0)
    assert len(Gs) == nmat
    for D, G in zip(Ds, Gs):
        J, K, _, _, _ = oracle.direct_jk(fb, D, exx=exx)
        assert np.abs(G - (J - K)).max() < TOL
        assert np.abs(G - G.T).max() == 0.0
    st = eng.stats
    assert st["quartets_evaluated_last"] == st["canonical_quartets"]      # one pass per batch, every quartet once
    eng.close()


def test_fe4s4_uhf_blocks(Int4C2E, oracle):
    """fe4s4 / 6-31G* (d, f shells, deep contractions, UHF-type call): sampled exact blocks."""
    mol, fb = load_fixture_molecule("fe4s4")
    n = fb.nbf
    Da, Db = H.random_symmetric_density(n, 1), H.random_symmetric_density(n, 2)
    eng = _engine(Int4C2E, fb)
    J, _, Ka, Kb = eng.ContractInts(None, Da, Db, 1, 0)
    nfe = 12  # shells per Fe
    for sa, sb in ((0, 0), (11, 5), (nfe * 4 + 3, 9), (nfe + 10, nfe * 2 + 11)):
        Jb, Kab = oracle.jk_block(fb, Da + Db, Da, sa, sb)
        ia, ib = fb.shell2bf[sa], fb.shell2bf[sb]
        assert np.abs(Jb - J[ia:ia + fb.nfun[sa], ib:ib + fb.nfun[sb]]).max() < TOL, (sa, sb)
        assert np.abs(Kab - Ka[ia:ia + fb.nfun[sa], ib:ib + fb.nfun[sb]]).max() < TOL, (sa, sb)
    eng.close()


def test_cpp_adaptor_matches_oracle(Int4C2E, oracle, tmp_path):
    """The C++ adaptor class (chinium_b200/cpp/Int4C2E_b200.hpp), driven like SelfConsistentField.cpp:47-53 +
    Restricted/SP.cpp:47 by tests/cpp/adaptor_test.cpp, against the oracle."""
    import subprocess
    import cpp_adaptor
    exe = cpp_adaptor.build()
    mol, fb = load_fixture_molecule("h2o")
    n = fb.nbf
    D = H.random_symmetric_density(n, 0)
    inp, outp = tmp_path / "in.txt", tmp_path / "out.txt"
    cpp_adaptor.write_input(str(inp), fb, D)
    r = subprocess.run([exe, str(inp), str(outp)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Done in" in r.stdout and "After screening" in r.stdout
    lines = open(outp).read().split()
    assert int(lines[0]) == 45150 and int(lines[1]) == 3214        # the reference's own counts for h2o (its loop keeps 3214 of 4368 quartets)
    ih = lines.index("HESS")
    v = np.array([float(x) for x in lines[2:ih]])
    J = v[:n * n].reshape(n, n, order="F"); K = v[n * n:2 * n * n].reshape(n, n, order="F")
    G = v[2 * n * n:3 * n * n].reshape(n, n, order="F")
    Jo, Ko, _, _, _ = oracle.direct_jk(fb, D, exx=0.5)
    assert np.abs(J - Jo).max() < TOL and np.abs(K - Ko).max() < TOL
    assert np.abs(G - (Jo - Ko)).max() < TOL
    assert v[3 * n * n] == 0.0      # absent Ka -> zeros
    ng = int(v[3 * n * n + 1])
    g = v[3 * n * n + 2:3 * n * n + 2 + ng]
    go = oracle.contract_grads(fb, D, D, 0.5)
    assert ng == len(go) and np.abs(g - go).max() < 1e-9 * max(1.0, np.abs(go).max())     # Restricted/Grad.cpp:66 call
    nh = int(lines[ih + 1])                                                                  # Restricted/Hess.cpp:67 call
    Hc = np.array([float(x) for x in lines[ih + 2:ih + 2 + nh * nh]]).reshape(nh, nh)
    Ho = oracle.contract_hess(fb, D, 0.5)
    assert nh == Ho.shape[0] and np.abs(Hc - Ho).max() < 1e-9 * max(1.0, np.abs(Ho).max())


def test_h2o64_schwarz_screened_blocks(Int4C2E, oracle):
    """(H2O)64 / def2-TZVP (nbf 2752; BASELINE config 5): the unscreened job is 2.7e11 quartets, so the engine runs it
    with a Cauchy-Schwarz threshold (Int4C2E's `threshold`, Int4C2E.cpp:108-113) of 1e-13.  Sampled J/K blocks must still
    match the UNSCREENED oracle to 1e-10 (the reference itself screens nothing)."""
    mol, fb = load_fixture_molecule("h2o64")
    n = fb.nbf
    D = H.random_symmetric_density(n, 0)
    eng = _engine(Int4C2E, fb, thr=1e-13)
    st = eng.stats
    assert 1.2e10 < st["canonical_quartets"] < 1.5e10          # 1.338e10 of 2.7375e11 survive
    J, K, _, _ = eng.ContractInts(D, None, None, 1, 0)
    assert np.abs(J - J.T).max() == 0 and np.abs(K - K.T).max() == 0
    per = fb.nshell // 64   # shells per water molecule
    for sa, sb in ((0, 0), (4, 7 * per + 2), (per * 20 + 9, per * 20 + 3), (per * 63 + 8, 5)):
        Jb, Kb = oracle.jk_block(fb, 2 * D, D, sa, sb)
        ia, ib = fb.shell2bf[sa], fb.shell2bf[sb]
        assert np.abs(Jb - J[ia:ia + fb.nfun[sa], ib:ib + fb.nfun[sb]]).max() < TOL, (sa, sb)
        assert np.abs(Kb - K[ia:ia + fb.nfun[sa], ib:ib + fb.nfun[sb]]).max() < TOL, (sa, sb)
    eng.close()


def test_schwarz_threshold_matches_reference_count(Int4C2E, oracle):
    """threshold > 0: the surviving shell quartets / unique integrals follow the reference's predicate
    sqrt(|Diag(bf1,bf2) Diag(bf3,bf4)|) > Threshold on the shell-pair maxima, and J/K stay within the bound."""
    mol, fb = load_fixture_molecule("bo3h3")
    D = H.random_symmetric_density(fb.nbf, 0)
    full = _engine(Int4C2E, fb)
    Jf, Kf, _, _ = full.ContractInts(D, None, None, 1, 0)
    nfull = full.stats["canonical_quartets"]
    full.close()
    for thr in (1e-6, 1e-9):
        eng = _engine(Int4C2E, fb, thr=thr)
        st = eng.stats
        assert 0 < st["canonical_quartets"] < nfull
        J, K, _, _ = eng.ContractInts(D, None, None, 1, 0)
        # every dropped quartet has |(ab|cd)| <= thr; at most nbf^2 of them touch one element, |D| <= 1/nbf
        assert np.abs(J - Jf).max() < thr * fb.nbf * 4 and np.abs(K - Kf).max() < thr * fb.nbf * 4
        eng.close()


def test_density_weighted_screening(Int4C2E, oracle):
    """cf_set_density_threshold: with a small density (a late-SCF difference density) the effective
    Schwarz threshold rises to dthr / max|D|, fewer quartets are evaluated, and J/K still match the unscreened oracle
    far below the 1e-10 bar (every neglected contribution is < dthr)."""
    mol, fb = load_fixture_molecule("bo3h3")
    n = fb.nbf
    dD = 1e-5 * H.random_symmetric_density(n, 7) * n        # entries ~1e-5: a typical late-iteration difference
    eng = _engine(Int4C2E, fb)
    J0, K0, _, _ = eng.ContractInts(dD, None, None, 1, 0)
    st0 = eng.stats
    assert st0["quartets_evaluated_last"] == st0["canonical_quartets"]
    assert st0["threshold_effective_last"] == 0.0
    eng.setDensityThreshold(1e-13)
    J1, K1, _, _ = eng.ContractInts(dD, None, None, 1, 0)
    st1 = eng.stats
    assert 0 < st1["quartets_evaluated_last"] < st0["canonical_quartets"]
    dmax = 2 * np.abs(dD).max()
    assert 0.2e-13 / dmax < st1["threshold_effective_last"] < 5e-13 / dmax     # max|D| is taken in the Cartesian basis
    Jo, Ko, _, _, _ = oracle.direct_jk(fb, dD)
    assert np.abs(J0 - Jo).max() < TOL and np.abs(K0 - Ko).max() < TOL
    assert np.abs(J1 - Jo).max() < TOL and np.abs(K1 - Ko).max() < TOL
    eng.setDensityThreshold(0.0)                             # off again: bit-identical to the first build
    J2, K2, _, _ = eng.ContractInts(dD, None, None, 1, 0)
    assert (J2 == J0).all() and (K2 == K0).all()
    eng.close()


def test_incremental_scf_energy(Int4C2E, oracle):
    """Incremental Fock builds G_n = G_(n-1) + G[D_n - D_(n-1)] with density-weighted screening reproduce the
    energy of the plain SCF (full build every iteration) to 1e-8 Eh and skip quartets in the late iterations."""
    mol, fb = load_fixture_molecule("sn2")
    S, T, V = oracle.one_electron(fb, mol.Z, mol.xyz_bohr)
    enuc = H.nuclear_repulsion(mol.Z, mol.xyz_bohr)
    eng = _engine(Int4C2E, fb)
    E_full, *_ = H.rhf(S, T + V, 18, lambda d, a, b: eng.ContractInts(d, a, b, 1, 0), enuc)
    evaluated, log = [], []

    def jk(D, full):
        eng.setDensityThreshold(0.0 if full else 1e-12)
        J, K, _, _ = eng.ContractInts(D, None, None, 1, 0)
        evaluated.append(eng.stats["quartets_evaluated_last"])
        return J, K

    E_inc, *_ = H.rhf_incremental(S, T + V, 18, jk, enuc, log=log)
    total = eng.stats["canonical_quartets"]
    eng.close()
    assert abs(E_inc - E_full) < 1e-8
    assert abs(E_inc - (-598.514802895)) < 1e-7
    inc = [e for e, f in zip(evaluated, log) if not f]
    assert inc and min(inc) < 0.8 * total        # late difference densities are small: quartets are skipped


@pytest.mark.parametrize("name,exx", [("h2o", 1.0), ("hf_tz", 0.7), ("bo3h3", 0.2)])
def test_contract_grads_parity(Int4C2E, oracle, name, exx):
    """ContractGrads(D1, D2) (Int4C2E.cpp:747-763 over getRepulsion1 :312-408): the fused derivative-ERI kernels vs the
    oracle's literal restatement (12 derivative buffers per quartet, per-atom scatter, D1 o G), s..f shells."""
    mol, fb = load_fixture_molecule(name)
    n = fb.nbf
    D1, D2 = H.random_symmetric_density(n, 21) * n, H.random_symmetric_density(n, 22) * n
    eng = _engine(Int4C2E, fb, exx=exx)
    g = eng.ContractGrads(D1, D2, 0)
    go = oracle.contract_grads(fb, D1, D2, exx)
    assert g.shape == go.shape
    assert np.abs(g - go).max() < 1e-9 * max(1.0, np.abs(go).max()), (np.abs(g - go).max(), np.abs(go).max())
    g2 = eng.ContractGrads(D1, D2, 0)
    assert (g2 == g).all()                                       # fixed-order row sums: repeatable bit for bit
    natom = len(go) // 3
    assert np.abs(g.reshape(natom, 3).sum(axis=0)).max() < 1e-9 * max(1.0, np.abs(go).max())
    # J-only (EXX <= 0) and the usual D1 == D2 call of Restricted/Grad.cpp:66
    eng.EXX = 0.0
    assert np.abs(eng.ContractGrads(D2, D2, 0) - oracle.contract_grads(fb, D2, D2, 0.0)).max() < 1e-9 * max(1.0, np.abs(go).max())
    eng.close()


@pytest.mark.parametrize("minq", ["0", "1000000000000"])
def test_contract_grads_both_kernel_families(oracle, minq, tmp_path):
    """Every class that has a thread-per-quartet gradient kernel also has the CTA-per-quartet one; which one runs
    depends on the task's quartet count.  Force each family in turn (fresh process: the switch is read once) on a
    molecule with s..f shells."""
    import subprocess, sys, textwrap
    code = textwrap.dedent("""
        import sys, os
        sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
        import numpy as np
        from chinium_b200 import Int4C2E
        from chinium_b200.inputs import load_fixture_molecule
        from oracle_lib import Oracle
        import scf_harness as H
        worst = 0.0
        for name, exx in (("hf_tz", 0.7), ("bo3h3", 0.2)):
            mol, fb = load_fixture_molecule(name)
            n = fb.nbf
            D1, D2 = H.random_symmetric_density(n, 41) * n, H.random_symmetric_density(n, 42) * n
            eng = Int4C2E(fb, exx, -1.0)
            g = eng.ContractGrads(D1, D2, 0)
            go = Oracle().contract_grads(fb, D1, D2, exx)
            worst = max(worst, float(np.abs(g - go).max() / max(1.0, np.abs(go).max())))
            eng.close()
        print("WORST", worst)
    """ % (ROOT, ROOT))
    env = dict(os.environ, CF_GRAD_TPQ_MINQ=minq)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    worst = float(r.stdout.strip().split("WORST")[-1])
    assert worst < 1e-9, worst


def test_contract_grads_partition_sum(Int4C2E, oracle):
    """world_size = 2: the two partitions' gradient shares add up to the whole (the caller's all-reduce)."""
    mol, fb = load_fixture_molecule("h2o")
    n = fb.nbf
    D = H.random_symmetric_density(n, 23) * n
    whole = _engine(Int4C2E, fb)
    g = whole.ContractGrads(D, D, 0)
    whole.close()
    parts = []
    for r in range(2):
        e = _engine(Int4C2E, fb, rank=r, world_size=2)
        parts.append(e.ContractGrads(D, D, 0))
        e.close()
    assert np.abs(parts[0] + parts[1] - g).max() < 1e-12 * max(1.0, np.abs(g).max())


def test_sn2_recorded_forces_gpu(Int4C2E, oracle):
    """The forces Chinium recorded for CH3ClF- (tools/sn2/sn2.cnm.log:211-216) with the engine's ContractGrads in the
    reference's gradient assembly (Restricted/Grad.cpp:60-70); see tests/test_oracle.py::test_sn2_recorded_forces for
    why the bound is 1e-5 Eh/bohr."""
    mol, fb = load_fixture_molecule("sn2")
    S, T, V = oracle.one_electron(fb, mol.Z, mol.xyz_bohr)
    enuc = H.nuclear_repulsion(mol.Z, mol.xyz_bohr)
    eng = _engine(Int4C2E, fb)
    E, D, F, _ = H.rhf(S, T + V, 18, lambda d, a, b: eng.ContractInts(d, a, b, 1, 0), enuc, tol=1e-9)
    g2 = eng.ContractGrads(D, D, 0)
    eng.close()
    forces = -H.rhf_total_gradient(oracle, fb, mol, D, F, S, 18, g2)
    assert abs(E - (-598.514802895)) < 1e-7
    assert np.abs(forces - H.SN2_FORCES_LOG).max() < 1e-5, forces
    assert np.abs(forces.sum(axis=0)).max() < 1e-7


@pytest.mark.parametrize("name", ["h2o", "hf_tz", "bo3h3"])
def test_contract_hess_golden(Int4C2E, name):
    """ContractHesss (Int4C2E.cpp:792-811 over getRepulsion2 :410-492): the fused second-derivative kernels (eri_hess.cuh)
    vs the committed outputs of the oracle's literal restatement (tests/golden/make_hess_golden.py: 78 buffers per quartet,
    per-atom scatter, raw + raw^T - diag), s..f shells, with and without exchange."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "hess_golden.npz"))
    mol, fb = load_fixture_molecule(name)
    n = fb.nbf
    D = H.random_symmetric_density(n, 31) * n
    eng = _engine(Int4C2E, fb, exx=0.6)
    Hm = eng.ContractHesss(D, D, 0)
    ref = g[name + "_hess"]
    assert Hm.shape == ref.shape
    assert np.abs(Hm - Hm.T).max() == 0.0
    assert np.abs(Hm - ref).max() < 1e-9 * max(1.0, np.abs(ref).max()), (np.abs(Hm - ref).max(), np.abs(ref).max())
    natom = ref.shape[0] // 3
    assert np.abs(Hm.reshape(3 * natom, natom, 3).sum(axis=1)).max() < 1e-8 * max(1.0, np.abs(ref).max())   # translation
    eng.EXX = 0.0                                                    # Coulomb part only (reference: kscale > 0 guard, :475)
    Hj = eng.ContractHesss(None, D, 0)                               # D1 is ignored like in the reference (:793)
    refj = g[name + "_hess_j"]
    assert np.abs(Hj - refj).max() < 1e-9 * max(1.0, np.abs(refj).max())
    eng.close()


def test_contract_hess_oracle_and_partition(Int4C2E, oracle):
    """Against the live oracle with a different density and EXX, and world_size = 2: the partitions' shares add up."""
    mol, fb = load_fixture_molecule("h2o")
    n = fb.nbf
    D = H.random_symmetric_density(n, 77) * n
    eng = _engine(Int4C2E, fb, exx=1.0)
    Hm = eng.ContractHesss(D, D, 0)
    eng.close()
    Ho = oracle.contract_hess(fb, D, 1.0)
    assert np.abs(Hm - Ho).max() < 1e-9 * max(1.0, np.abs(Ho).max())
    parts = []
    for r in range(2):
        e = _engine(Int4C2E, fb, exx=1.0, rank=r, world_size=2)
        parts.append(e.ContractHesss(D, D, 0))
        e.close()
    assert np.abs(parts[0] + parts[1] - Hm).max() < 1e-11 * max(1.0, np.abs(Hm).max())


def test_primitive_cutoff_margin(tmp_path):
    """The J/K kernels skip primitive quartets with |c_ab c_cd wgt| < 1e-18 (engine.cu, CF_PRIM_CUT_DEFAULT; scan in
    profiles/r2l_primcut.txt).  Against a build with the cutoff switched off (developer knob CF_PRIM_CUT=0, fresh process: the
    knob is read once) an O(1) all-positive density -- skipped positive integrals add up coherently -- must agree to 1e-11,
    a tenth of the parity bar, on a molecule with s..f shells and on bo3h3."""
    import subprocess, sys, textwrap
    code = textwrap.dedent("""
        import sys, os
        sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
        import numpy as np
        from chinium_b200 import Int4C2E
        from chinium_b200.inputs import load_fixture_molecule
        import scf_harness as H
        for name in ("hf_tz", "bo3h3"):
            mol, fb = load_fixture_molecule(name)
            D = np.abs(H.random_symmetric_density(fb.nbf, 5) * fb.nbf)
            eng = Int4C2E(fb, 1.0, -1.0)
            J, K, _, _ = eng.ContractInts(D, None, None, 1, 0)
            np.save(os.path.join(%r, name + "_" + os.environ.get("TAGX", "x") + ".npy"), np.stack([J, K]))
            eng.close()
    """ % (ROOT, ROOT, str(tmp_path)))
    for tag, extra in (("default", {}), ("nocut", {"CF_PRIM_CUT": "0"})):
        env = {k: v for k, v in os.environ.items() if k != "CF_PRIM_CUT"}
        env.update(TAGX=tag, **extra)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
    for name in ("hf_tz", "bo3h3"):
        a, b = np.load(tmp_path / (name + "_default.npy")), np.load(tmp_path / (name + "_nocut.npy"))
        assert np.abs(b).max() > 1.0                                   # O(1) and larger elements
        assert np.abs(a - b).max() < 1e-11, (name, np.abs(a - b).max())

