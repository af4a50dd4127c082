"""CPU tests of the host side: inputs, the C-ABI library (load + exported symbols, loud failure without a
GPU), and the multi-rank plumbing with gloo (world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import scf_harness as H
from chinium_b200.inputs import load_fixture_molecule

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fixture_sizes_match_survey():
    exp = {"h2o": (12, 24), "bo3h3": (33, 71), "c18": (180, 540), "fe4s4": (80, 208), "h2o64": (1216, 2752), "sn2": (29, 61)}
    for name, (nsh, nbf) in exp.items():
        mol, fb = load_fixture_molecule(name)
        assert (fb.nshell, fb.nbf) == (nsh, nbf), name
    mol, _ = load_fixture_molecule("fe4s4")
    assert mol.nalpha_nbeta == (92, 74)


def test_abi_exports_every_declared_symbol():
    from chinium_b200.fock import LIB_PATH
    assert os.path.exists(LIB_PATH), "build the library first (__graft_entry__.build())"
    hdr = open(os.path.join(ROOT, "include", "chinium_fock.h")).read()
    names = set(re.findall(r"\b(cf_[a-z0-9_]+)\s*\(", hdr))
    lib = ctypes.CDLL(LIB_PATH)
    for n in sorted(names):
        assert hasattr(lib, n), n
    assert len(names) >= 14


def test_fails_loudly_without_gpu(have_gpu):
    if have_gpu:
        pytest.skip("GPU present")
    from chinium_b200 import Int4C2E, FockEngineError
    mol, fb = load_fixture_molecule("h2o")
    eng = Int4C2E(fb)
    with pytest.raises(FockEngineError, match="no CPU fallback"):
        eng.ContractInts(np.eye(fb.nbf), None, None, 1, 0)


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or any CPU fallback); developer scripts that use the oracle as a
    checker live under tests/ (dev_*.py, dist_grad_check.py), never under tools/ or the package."""
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith((".py", ".sh")):
            src = open(os.path.join(ROOT, "tools", f)).read()
            assert "liboracle" not in src and "oracle_lib" not in src, f
    for dirpath, _, files in os.walk(os.path.join(ROOT, "chinium_b200")):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")) and f != "rys_tables_data.h":
                src = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in src and "oracle_lib" not in src and "oracle/" not in src.replace("the oracle/", ""), f


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
from chinium_b200.inputs import load_fixture_molecule
from oracle_lib import Oracle
import scf_harness as H
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
mol, fb = load_fixture_molecule("h2o")
D = H.random_symmetric_density(fb.nbf, 0)
o = Oracle()
# each rank digests its static share of the bra pairs (stand-in for the device partition), converts to the
# engine's 64-bit fixed point, and the partial accumulators are summed with an INTEGER all-reduce
J, K, _, _, _ = o.direct_jk(fb, D, nthreads=1, stride=world, offset=rank)
scale = 2.0 ** 40
acc = torch.from_numpy(np.rint(np.stack([J, K]) * scale).astype(np.int64))
dist.all_reduce(acc, op=dist.ReduceOp.SUM)
if rank == 0:
    Jf, Kf, _, _, _ = o.direct_jk(fb, D, nthreads=1)
    got = acc.numpy().astype(np.float64) / scale
    err = max(np.abs(got[0] - Jf).max(), np.abs(got[1] - Kf).max())
    print("GLOO_ERR", err)
dist.destroy_process_group()
"""


def test_two_rank_integer_allreduce_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    m = re.search(r"GLOO_ERR ([0-9.e+-]+)", out.stdout)
    assert m and float(m.group(1)) < 1e-10, out.stdout


def test_bench_reference_arm_runs_on_cpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--workload", "h2o"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and "cpu_baseline" in line


def test_bench_gradient_reference_arm_runs_on_cpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--path", "grad", "--impl", "reference", "--workload", "h2o"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"


def test_cpp_adaptor_compiles_and_fails_loudly_without_gpu(tmp_path, have_gpu):
    """The C++ adaptor (reference's Int4C2E method names) builds against the C ABI; with no GPU it throws
    instead of falling back."""
    import cpp_adaptor
    exe = cpp_adaptor.build()
    if have_gpu:
        pytest.skip("GPU present (covered by the gpu test)")
    mol, fb = load_fixture_molecule("h2o")
    inp = tmp_path / "in.txt"
    cpp_adaptor.write_input(str(inp), fb, H.random_symmetric_density(fb.nbf, 0))
    out = subprocess.run([exe, str(inp), str(tmp_path / "out.txt")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 3, (out.returncode, out.stderr)
    assert "no CPU fallback" in out.stderr or "no CUDA device" in out.stderr


def test_wg_strides_table_matches_bank_model():
    """chinium_b200/csrc/wg_strides_data.h is generated by tools/wg_bank_model.py --emit: every warp-group configuration of
    eri_wg.cuh (wg_cfg_base) must be known to the tool with the same (MK, swap, HS, MINB), and the header must hold the
    strides the model picks for it, within the kernel's size / alignment constraints."""
    import re, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    import importlib
    wbm = importlib.import_module("wg_bank_model")
    src = open(os.path.join(root, "chinium_b200", "csrc", "eri_wg.cuh")).read()
    body = src[src.index("constexpr int wg_cfg_base(int key)"):src.index("constexpr int wg_cfg_alt1(int key)")]
    WGC = lambda mk, sw, hs, minb: mk | (sw << 8) | (hs << 12) | (minb << 16)
    cfg = {}
    for m in re.finditer(r"case (\d+): return ([^;]+);", body):
        v = eval(m.group(2), {"WGC": WGC})
        cfg[int(m.group(1))] = (v & 255, (v >> 8) & 1, ((v >> 12) & 15) or 1, ((v >> 16) & 15) or 2)
    assert len(cfg) >= 30
    for key, c in cfg.items():
        assert wbm.CFG[key] == c, (key, c, wbm.CFG[key])
    hdr = {}
    for line in open(os.path.join(root, "chinium_b200", "csrc", "wg_strides_data.h")):
        m = re.match(r"\s+case (\d+): return (\d+)LL \| \((\d+)LL << 8\) \| \((\d+)LL << 20\);", line)
        if m:
            hdr[int(m.group(1))] = tuple(int(x) for x in m.groups()[1:])
    for key in cfg:
        shp, cur, cb, best = wbm.search(key)
        la, lb, lc, ld, mk, hs, minb = shp
        k = ((((la * 4 + lb) * 4 + lc) * 4 + ld) * 8 + mk) * 16 + hs
        assert k in hdr, (key, shp)
        labp, tsz, scr = hdr[k]
        assert (labp, tsz, scr) == (best[2], best[3], best[1]), (key, hdr[k], best)
        assert wbm.constraints(la, lb, lc, ld, mk, hs, labp, tsz, scr)

