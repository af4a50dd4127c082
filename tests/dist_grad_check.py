#!/usr/bin/env python3
"""torchrun check of the multi-GPU paths: DistributedInt4C2E.ContractInts and .ContractGrads on N ranks (NCCL) against the
CPU oracle (rank 0).  usage: torchrun --nproc-per-node N tests/dist_grad_check.py [molecule]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
from chinium_b200.inputs import load_fixture_molecule
from chinium_b200.distributed import DistributedInt4C2E
import scf_harness as H

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
name = sys.argv[1] if len(sys.argv) > 1 else "bo3h3"
mol, fb = load_fixture_molecule(name)
n = fb.nbf
D = H.random_symmetric_density(n, 5) * n
eng = DistributedInt4C2E(fb, 0.5, -1.0)
J, K, _, _ = eng.ContractInts(D, None, None, 1, 0)
g = eng.ContractGrads(D, D, 0)
if rank == 0:
    from oracle_lib import Oracle
    o = Oracle()
    Jo, Ko, _, _, _ = o.direct_jk(fb, D, exx=0.5)
    go = o.contract_grads(fb, D, D, 0.5)
    eJ, eK = np.abs(J - Jo).max(), np.abs(K - Ko).max()
    eg = np.abs(g - go).max() / max(1.0, np.abs(go).max())
    print("DIST_CHECK world %d %s: max|dJ| %.2e max|dK| %.2e rel|dg| %.2e" % (world, name, eJ, eK, eg))
    assert eJ < 1e-10 and eK < 1e-10 and eg < 1e-9
    print("DIST_CHECK OK")
eng.close()
dist.destroy_process_group()
