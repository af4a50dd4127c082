#!/usr/bin/env python3
"""In-library multi-GPU (cf_create_multi, ONE process): host-call time of cf_build_jk for 1..N GPUs.
usage: bench_multi.py [workload ...]   (run under gpurun --gpus N)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from chinium_b200 import Int4C2E
from chinium_b200.inputs import load_fixture_molecule
import scf_harness as H

thr = {"h2o64": 1e-13}
for w in (sys.argv[1:] or ["c18"]):
    mol, fb = load_fixture_molecule(w)
    n = fb.nbf
    D = H.random_symmetric_density(n, 0)
    pin_in = torch.empty((n, n), dtype=torch.float64).pin_memory(); pin_in.numpy()[...] = D
    outs = [torch.empty((n, n), dtype=torch.float64).pin_memory().numpy().T for _ in range(2)]
    base = None
    nd = 1
    while nd <= torch.cuda.device_count():
        t0 = time.perf_counter()
        eng = Int4C2E(fb, 1.0, thr.get(w, -1.0), ndevices=0 if nd == 1 else nd)
        eng._ensure()
        setup = time.perf_counter() - t0
        for _ in range(3):
            eng._contract(pin_in.numpy().T, None, None, out=[outs[0], outs[1], None, None])
        ts = []
        for _ in range(5):
            t = time.perf_counter(); eng._contract(pin_in.numpy().T, None, None, out=[outs[0], outs[1], None, None]); ts.append(time.perf_counter() - t)
        st = eng.stats
        ms = float(np.median(ts)) * 1e3
        base = base or ms
        print(json.dumps({"workload": w, "n_gpus": nd, "one_process": True, "host_call_ms": ms, "device_ms": st["ms_device_last"], "eri_ms_slowest_device": st["ms_eri_last"],
                          "speedup": base / ms, "efficiency": base / ms / nd, "setup_s": setup, "quartets_per_s": st["canonical_quartets"] / (ms * 1e-3),
                          "checksum": float(np.abs(outs[0]).sum())}), flush=True)
        eng.close()
        nd *= 2
