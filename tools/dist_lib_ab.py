#!/usr/bin/env python
"""One leg of an A/B of libpmg.so builds on several GPUs (the library is chosen per process: AB_LIB=path):
V-cycle time per cycle at N = 16385 and 32769, 20 cycles per solve, max over ranks.
    AB_LIB=ab/libpmg_x.so torchrun --nproc-per-node R tools/dist_lib_ab.py [N ...]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pmg_b200 as pmg  # noqa: E402

lib = os.environ.get("AB_LIB")
if lib:
    sys.modules["_pmg_b200_pkg"].LIB_PATH = os.path.join(ROOT, lib)  # before the first call loads the library
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world, dev = pmg.init_distributed_from_torch(local)
for n in [int(a) for a in sys.argv[1:]] or [16385, 32769]:
    s = pmg.Solver(n, omega=2.0 / 3.0, device=dev, rank=rank, n_ranks=world)
    s.set_rhs_sine()
    ms = []
    for it in range(5):
        s.zero_guess()
        torch.cuda.synchronize()
        dist.barrier()
        k, hist = s.solve(pmg.V, 0.0, 20)
        ms.append(s.last_ms)
    t = torch.tensor(ms[1:], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("lib=%s ranks=%d N=%d: %d cycles, ms %s -> %.1f us/cycle, last norm %r"
              % (lib, world, n, k, [round(float(v), 3) for v in t], 1e3 * float(t.min()) / k, float(hist[-1])), flush=True)
    s.close()
pmg.comm_finalize()
dist.destroy_process_group()
