"""Multi-GPU inside the library (cf_create_multi): ONE process, worker thread + stream per device, NCCL int64 all-reduce
of the fixed-point accumulators.  Needs >= 2 GPUs (skipped otherwise); run with `gpurun --gpus 2 -- pytest tests/test_gpu_multi.py -m gpu`."""
import subprocess

import numpy as np
import pytest

import scf_harness as H
from chinium_b200.inputs import load_fixture_molecule

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def ndev():
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    return min(n, 8)


def test_cpp_adaptor_multi_device(ndev, oracle, tmp_path):
    """tests/cpp/adaptor_test.cpp with a device count: the C++ adaptor's MultiDevice handle (no Python, no torch in that
    process) returns J, K, G bit-identical to the single-GPU handle and the same gradient."""
    import cpp_adaptor
    exe = cpp_adaptor.build()
    mol, fb = load_fixture_molecule("bo3h3")
    n = fb.nbf
    D = H.random_symmetric_density(n, 0)
    inp, outp = tmp_path / "in.txt", tmp_path / "out.txt"
    cpp_adaptor.write_input(str(inp), fb, D)
    r = subprocess.run([exe, str(inp), str(outp), str(ndev)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    last = open(outp).read().strip().splitlines()[-1].split()
    assert last[0] == "MULTI" and int(last[1]) == ndev and int(last[2]) == 1, last
    assert float(last[3]) < 1e-11
    v = np.array([float(x) for x in open(outp).read().split()[2:2 + n * n]])
    Jo, Ko, _, _, _ = oracle.direct_jk(fb, D, exx=0.5)
    assert np.abs(v.reshape(n, n, order="F") - Jo).max() < 1e-10


@pytest.mark.parametrize("name", ["bo3h3", "c18"])
def test_multi_handle_bit_identical(ndev, name):
    from chinium_b200 import Int4C2E
    import time
    mol, fb = load_fixture_molecule(name)
    n = fb.nbf
    Dd, Da, Db = (H.random_symmetric_density(n, s) for s in (0, 1, 2))
    one = Int4C2E(fb, 1.0, -1.0)
    multi = Int4C2E(fb, 1.0, -1.0, ndevices=ndev)
    ref = one.ContractInts(Dd, Da, Db, 1, 0)
    t = time.perf_counter()
    got = multi.ContractInts(Dd, Da, Db, 1, 0)
    dt = time.perf_counter() - t
    for a, b in zip(ref, got):
        assert (a == b).all()
    st1, stn = one.stats, multi.stats
    assert stn["canonical_quartets"] == st1["canonical_quartets"] and stn["quartets_evaluated_last"] == st1["quartets_evaluated_last"]
    assert stn["primitive_quartets_executed_last"] == st1["primitive_quartets_executed_last"]
    print("%s on %d GPUs in one process: host call %.2f ms (device %.2f ms; one GPU %.2f ms)" % (name, ndev, dt * 1e3, stn["ms_device_last"], st1["ms_device_last"]))
    Gs1 = one.ContractInts([Dd, Da], 1, 0)
    Gsn = multi.ContractInts([Dd, Da], 1, 0)
    for a, b in zip(Gs1, Gsn):
        assert (a == b).all()
    if name == "bo3h3":
        g1 = one.ContractGrads(Dd * n, Dd * n, 0)
        gn = multi.ContractGrads(Dd * n, Dd * n, 0)
        assert np.abs(g1 - gn).max() < 1e-12 * max(1.0, np.abs(g1).max())
    one.close(); multi.close()
