#!/usr/bin/env python
"""Multi-GPU tuning under torchrun: solve time at N for several agglomeration thresholds, plus (with
PMG_DIST_TRACE=1) the per-phase device times of the distributed cycle on rank 0 and the last rank.
  torchrun ... tools/dist_sweep.py N thr1,thr2,..."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pmg_b200 as pmg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16385
thrs = [int(t) for t in (sys.argv[2] if len(sys.argv) > 2 else "2049").split(",")]
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world, dev = pmg.init_distributed_from_torch(local)
pmg.lib().pmg_dist_trace_dump.argtypes = [__import__("ctypes").c_int]
pmg.lib().pmg_dist_trace_dump.restype = None
for thr in thrs:
    s = pmg.Solver(n, omega=2.0 / 3.0, device=dev, rank=rank, n_ranks=world, agglomerate_below=thr)
    s.set_rhs_sine()
    ms = []
    for it in range(4):
        s.zero_guess()
        torch.cuda.synchronize()
        dist.barrier()
        k, hist = s.solve(pmg.V, 1e-8, 100)
        ms.append(s.last_ms)
        if it == 0:
            pmg.lib().pmg_dist_trace_dump(-1)  # drop the warm-up marks
    t = torch.tensor(ms[1:], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("ranks=%d N=%d agglomerate_below=%d: cycles=%d solve ms %s -> %.3f ms/cycle" %
              (world, n, thr, k, [round(float(v), 2) for v in t], float(t.min()) / k), flush=True)
    pmg.lib().pmg_dist_trace_dump(0)
    dist.barrier()
    pmg.lib().pmg_dist_trace_dump(world - 1)
    dist.barrier()
    s.close()
pmg.comm_finalize()
dist.destroy_process_group()
