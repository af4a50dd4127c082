#!/usr/bin/env python3
"""Regenerate tests/golden/{molecules,basis_library}.json from the reference's DATA files.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_fixtures.py
Only input data is taken from the reference (geometries of examples/*.inp and tools/sn2, and the
per-element shells of the Gaussian94 basis files the configs name); no reference source code.
"""
import json, os, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from chinium_b200.inputs import read_inp, parse_gbs  # noqa: E402

REF = "/root/reference"


def mol_entry(m, basis=None):
    return {"atoms": [[s] + [float(x) for x in r] for s, r in zip(m.symbols, m.xyz_angstrom)],
            "charge": m.charge, "multiplicity": m.multiplicity, "basis": basis or m.basis}


def h2o64():
    """SURVEY 8d generator: 4x4x4 simple-cubic lattice (3.104 A) of randomly rotated waters."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(64)
    rots = Rotation.random(64, random_state=rng).as_matrix()
    mono = np.array([[0.0, 0.0, 0.0], [0.0, -0.757, 0.587], [0.0, 0.757, 0.587]])  # examples/h2o.inp:6-8
    atoms = []
    n = 0
    for iz in range(4):
        for iy in range(4):
            for ix in range(4):
                o = 3.104 * np.array([ix, iy, iz], dtype=float)
                for s, r in zip("OHH", mono):
                    atoms.append([s] + [float(x) for x in np.round(o + rots[n] @ r, 10)])
                n += 1
    return {"atoms": atoms, "charge": 0, "multiplicity": 1, "basis": "def2-tzvp"}


def sn2():
    """tools/sn2/sn2.cnm.gjf:8-14 through the pipeline Chinium actually saw:
    Gaussian converts A->bohr with 0.52917721092 and prints 12 decimals, tools/gau_cnm.sh:57-59
    divides by 1.8897259886 and prints %.8f A, Chinium multiplies by 1.8897259886."""
    syms, xyz = [], []
    with open(os.path.join(REF, "tools/sn2/sn2.cnm.gjf")) as f:
        lines = f.read().splitlines()
    for ln in lines[8:14]:
        t = ln.split()
        syms.append(t[0]); xyz.append([float(x) for x in t[1:4]])
    xyz = np.array(xyz)
    bohr_g = np.round(xyz / 0.52917721092, 12)
    ang_c = np.round(bohr_g / 1.8897259886, 8)
    return {"atoms": [[s] + [float(x) for x in r] for s, r in zip(syms, ang_c)],
            "charge": -1, "multiplicity": 1, "basis": "cc-pvdz",
            "golden_energy_hartree": -598.514802895, "golden_source": "tools/sn2/sn2.cnm.log:204"}


def main():
    mols = {}
    for name in ("h2o", "bo3h3", "c18", "fe4s4", "ch2"):
        mols[name] = mol_entry(read_inp(os.path.join(REF, "examples", name + ".inp")))
    mols["h2o64"] = h2o64()
    mols["sn2"] = sn2()
    # textbook anchors: geometries given in bohr
    mols["h2_sto3g"] = {"atoms": [["H", 0, 0, 0], ["H", 0, 0, 0]], "xyz_bohr": [[0, 0, 0], [0, 0, 1.4]],
                        "charge": 0, "multiplicity": 1, "basis": "sto-3g"}
    mols["h2o_sto3g"] = {"atoms": [["O", 0, 0, 0], ["H", 0, 0, 0], ["H", 0, 0, 0]],
                         "xyz_bohr": [[0.0, -0.143225816552, 0.0], [1.638036840407, 1.136548822547, 0.0],
                                      [-1.638036840407, 1.136548822547, 0.0]],
                         "charge": 0, "multiplicity": 1, "basis": "sto-3g"}
    # a small f-shell case for kernel parity: HF molecule with cc-pVTZ (H: up to d, F: up to f)
    mols["hf_tz"] = {"atoms": [["F", 0.0, 0.0, 0.0], ["H", 0.0, 0.3, 0.85]], "charge": 0, "multiplicity": 1,
                     "basis": "cc-pvtz"}
    need = {}
    for m in mols.values():
        need.setdefault(m["basis"], set()).update(a[0].upper() for a in m["atoms"])
    lib = {}
    for b, els in sorted(need.items()):
        parsed = parse_gbs(os.path.join(REF, "BasisSets", b + ".gbs"), els)
        assert set(parsed) == els, (b, els, set(parsed))
        lib[b] = {el: [[t, e, c] for (t, e, c) in parsed[el]] for el in sorted(parsed)}
    with open(os.path.join(HERE, "molecules.json"), "w") as f:
        json.dump(mols, f, indent=0)
    with open(os.path.join(HERE, "basis_library.json"), "w") as f:
        json.dump(lib, f, indent=0)
    print({k: len(v["atoms"]) for k, v in mols.items()})


if __name__ == "__main__":
    main()
