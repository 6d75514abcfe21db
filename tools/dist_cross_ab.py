#!/usr/bin/env python
"""A/B of the cross-cycle pass on row slabs (same box, same process): V-cycle time per cycle with PMG_CROSS off / on at
N = 16385 and N = 32769, 20 cycles per solve.  torchrun --nproc-per-node R tools/dist_cross_ab.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pmg_b200 as pmg  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world, dev = pmg.init_distributed_from_torch(local)
sizes = [int(a) for a in sys.argv[1:]] or [16385, 32769]
for n in sizes:
    ref = None
    for cross in (0, 1, 0, 1):
        pmg.set_cross_cycle(cross)
        s = pmg.Solver(n, omega=2.0 / 3.0, device=dev, rank=rank, n_ranks=world)
        s.set_rhs_sine()
        ms = []
        for it in range(4):
            s.zero_guess()
            torch.cuda.synchronize()
            dist.barrier()
            k, hist = s.solve(pmg.V, 0.0, 20)
            ms.append(s.last_ms)
        t = torch.tensor(ms[1:], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if ref is None:
            ref = hist.copy()
        dev_rel = float(np.max(np.abs(hist - ref) / ref))
        if rank == 0:
            print("ranks=%d N=%d cross=%d: %d cycles, ms %s -> %.1f us/cycle, history max rel dev vs first %.2e"
                  % (world, n, cross, k, [round(float(v), 3) for v in t], 1e3 * float(t.min()) / k, dev_rel), flush=True)
        s.close()
pmg.set_cross_cycle(-1)
pmg.comm_finalize()
dist.destroy_process_group()
