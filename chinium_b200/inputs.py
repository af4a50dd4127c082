"""Input marshalling for the J/K path: molecule + Gaussian94 basis -> flat `cf_basis` arrays.

This is the host-side replacement of what the reference does before it can call its
ERI layer:

* `.gbs` semantics follow the reference's reader (legacy/MwfnIO/Mwfn.cpp:606-682):
  `SP` lines split into an S shell followed by a P shell on the same exponents,
  P shells are Cartesian (type +1), D/F/G/H/I are pure (type -2..-6).
* shell order = centres in input order x shells in file order
  (src/Integral/Macro.h:1-25, `__Make_Basis_Set__`).
* coefficient renormalisation restates what `Normalize` copies back from the
  integral library (src/Integral/Normalization.cpp:11-18): every primitive is
  multiplied by the norm of its axis-aligned primitive, then the contraction is
  scaled so that the axis-aligned contracted function has unit self-overlap.
* coordinates: Angstrom * 1.8897259886 (src/Gateway.cpp:44-49, src/Macro/Unit.h:1).

Nothing here touches a GPU; the arrays produced are what `cf_create` consumes.
"""
from __future__ import annotations

import json
import math
import os
import re
from dataclasses import dataclass, field

import numpy as np

ANGSTROM_TO_BOHR = 1.8897259886  # the reference's constant (src/Macro/Unit.h:1)

_SYMBOLS = (
    "H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn "
    "Ga Ge As Se Br Kr Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe"
).split()
SYMBOL_TO_Z = {s.upper(): i + 1 for i, s in enumerate(_SYMBOLS)}

_L_OF = {"S": 0, "P": 1, "D": 2, "F": 3, "G": 4, "H": 5, "I": 6}


def shell_type_of(letter: str) -> int:
    """Reference convention: S=0, P=+1 (Cartesian), D..I = -2..-6 (pure)."""
    l = _L_OF[letter]
    return l if l <= 1 else -l


def nfun_of_type(t: int) -> int:
    l = abs(t)
    return 2 * l + 1 if t < 0 else (l + 1) * (l + 2) // 2


def parse_gbs(path: str, elements=None) -> dict:
    """Parse a Gaussian94 basis file -> {SYMBOL: [(type, exps, raw_coefs), ...]}."""
    out: dict = {}
    cur = None
    wanted = None if elements is None else {e.upper() for e in elements}
    with open(path) as f:
        lines = f.read().splitlines()
    i = 0
    num = lambda s: float(re.sub("[Dd]", "E", s))
    while i < len(lines):
        toks = lines[i].split()
        i += 1
        if not toks or toks[0].startswith("!"):
            continue
        w = toks[0]
        if w == "****":
            cur = None
        elif w[0] == "-" and not re.match(r"^-?\d", w):
            cur = w[1:].upper()
            out[cur] = []
        elif w.upper() in ("S", "SP", "P", "D", "F", "G", "H", "I") and cur is not None:
            n = int(toks[1])
            rows = [lines[i + k].split() for k in range(n)]
            i += n
            exps = [num(r[0]) for r in rows]
            if w.upper() == "SP":
                out[cur].append((0, exps, [num(r[1]) for r in rows]))
                out[cur].append((1, list(exps), [num(r[2]) for r in rows]))
            else:
                out[cur].append((shell_type_of(w.upper()), exps, [num(r[1]) for r in rows]))
    if wanted is not None:
        out = {k: v for k, v in out.items() if k in wanted}
    return out


def _dfact(n: int) -> float:  # (n)!! with (-1)!! = 1
    r = 1.0
    while n > 1:
        r *= n
        n -= 2
    return r


def normalize_shell(l: int, exps, coefs) -> np.ndarray:
    """Normalised contraction coefficients (what the reference stores in
    MwfnShell.NormalizedCoefficients, src/Integral/Normalization.cpp:15)."""
    a = np.asarray(exps, dtype=np.float64)
    c = np.asarray(coefs, dtype=np.float64).copy()
    pi32 = math.pi ** 1.5
    df = _dfact(2 * l - 1)
    c *= np.sqrt((2.0 ** l) * (2.0 * a) ** (l + 1.5) / (pi32 * df))
    s = a[:, None] + a[None, :]
    norm = np.sum(c[:, None] * c[None, :] * df * pi32 / ((2.0 ** l) * s ** (l + 1.5)))
    return c / math.sqrt(norm)


@dataclass
class Molecule:
    symbols: list
    xyz_angstrom: np.ndarray
    charge: int = 0
    multiplicity: int = 1
    basis: str = ""
    name: str = ""

    @property
    def Z(self):
        return np.array([SYMBOL_TO_Z[s.upper()] for s in self.symbols], dtype=np.int32)

    @property
    def xyz_bohr(self):
        return np.asarray(self.xyz_angstrom, dtype=np.float64) * ANGSTROM_TO_BOHR

    @property
    def nelec(self):
        return int(self.Z.sum()) - self.charge

    @property
    def nalpha_nbeta(self):
        nun = self.multiplicity - 1
        ne = self.nelec
        return (ne + nun) // 2, (ne - nun) // 2


@dataclass
class FlatBasis:
    """Host arrays of `struct cf_basis` (include/chinium_fock.h)."""
    type: np.ndarray           # int32 [nshell]
    nprim: np.ndarray          # int32 [nshell]
    prim_offset: np.ndarray    # int32 [nshell]
    exps: np.ndarray           # f64 [sum nprim]
    coefs_raw: np.ndarray      # f64 [sum nprim]
    coefs_normalized: np.ndarray
    center_xyz: np.ndarray     # f64 [nshell,3] bohr
    shell2atom: np.ndarray     # int32 [nshell]
    shell2bf: np.ndarray = field(default=None)

    def __post_init__(self):
        nf = np.array([nfun_of_type(int(t)) for t in self.type], dtype=np.int32)
        self.nfun = nf
        self.shell2bf = np.concatenate([[0], np.cumsum(nf)[:-1]]).astype(np.int32)
        self.nbf = int(nf.sum())

    @property
    def nshell(self):
        return len(self.type)


def build_basis(mol: Molecule, library: dict) -> FlatBasis:
    """library: {SYMBOL: [(type, exps, coefs), ...]} for this molecule's basis set."""
    types, nprim, exps, craw, cnorm, xyz, s2a = [], [], [], [], [], [], []
    R = mol.xyz_bohr
    for ia, sym in enumerate(mol.symbols):
        for (t, e, c) in library[sym.upper()]:
            types.append(int(t))
            nprim.append(len(e))
            exps += list(e)
            craw += list(c)
            cnorm += list(normalize_shell(abs(int(t)), e, c))
            xyz.append(R[ia])
            s2a.append(ia)
    nprim = np.array(nprim, dtype=np.int32)
    off = np.concatenate([[0], np.cumsum(nprim)[:-1]]).astype(np.int32)
    return FlatBasis(
        type=np.array(types, dtype=np.int32), nprim=nprim, prim_offset=off,
        exps=np.array(exps, dtype=np.float64), coefs_raw=np.array(craw, dtype=np.float64),
        coefs_normalized=np.array(cnorm, dtype=np.float64),
        center_xyz=np.ascontiguousarray(np.array(xyz, dtype=np.float64)),
        shell2atom=np.array(s2a, dtype=np.int32))


def read_inp(path: str) -> Molecule:
    """Keyword / next-line `.inp` reader (src/Gateway.cpp); only the keys the path needs."""
    with open(path) as f:
        lines = [ln.rstrip("\n") for ln in f]
    kv = {}
    i = 0
    syms, xyz = [], []
    while i < len(lines):
        key = lines[i].strip().lower()
        i += 1
        if not key:
            continue
        if key == "xyz":
            n = int(lines[i].split()[0]); i += 1
            for _ in range(n):
                t = lines[i].split(); i += 1
                syms.append(t[0]); xyz.append([float(x) for x in t[1:4]])
        else:
            if i < len(lines):
                kv[key] = lines[i].strip(); i += 1
    return Molecule(symbols=syms, xyz_angstrom=np.array(xyz), charge=int(kv.get("charge", 0)),
                    multiplicity=int(kv.get("spin", 1)), basis=kv.get("basis", "").split()[0].lower(),
                    name=os.path.splitext(os.path.basename(path))[0])


# ----------------------------------------------------------------------------------------------
# committed fixtures (tests/golden): molecules + the per-element shells of the basis sets they use
# ----------------------------------------------------------------------------------------------
_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def load_fixture_molecule(name: str, golden_dir: str = _GOLDEN):
    """-> (Molecule, FlatBasis) for one of the BASELINE configs, from committed fixtures."""
    with open(os.path.join(golden_dir, "molecules.json")) as f:
        mols = json.load(f)
    with open(os.path.join(golden_dir, "basis_library.json")) as f:
        lib = json.load(f)
    m = mols[name]
    mol = Molecule(symbols=[a[0] for a in m["atoms"]],
                   xyz_angstrom=np.array([a[1:4] for a in m["atoms"]], dtype=np.float64),
                   charge=m["charge"], multiplicity=m["multiplicity"], basis=m["basis"], name=name)
    if "xyz_bohr" in m:  # geometry given directly in bohr (sn2 pipeline, see make_fixtures.py)
        mol.xyz_angstrom = np.array(m["xyz_bohr"], dtype=np.float64) / ANGSTROM_TO_BOHR
    blib = {k: [(s[0], s[1], s[2]) for s in v] for k, v in lib[m["basis"]].items()}
    return mol, build_basis(mol, blib)
