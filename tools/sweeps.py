#!/usr/bin/env python
"""BASELINE config 5: standalone Jacobi-smoother bandwidth sweep 1K^2 .. 32K^2 (one GPU) and, under torchrun, a
weak-scaling point (pass N; DOF per GPU = N^2 / world).

  python tools/sweeps.py jacobi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweeps.py weak 32769
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pmg_b200 as pmg  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "jacobi"
if mode == "jacobi":
    for k in range(10, 16):
        n = 2 ** k + 1
        s = pmg.Solver(n, omega=2.0 / 3.0)
        s.set_rhs_sine()
        s.zero_guess()
        s.smooth(4, 1)
        sweeps = 100 if n <= 4097 else 20
        s.smooth(sweeps, 1)
        t1 = s.last_ms
        s.smooth(sweeps, 4)
        t4 = s.last_ms
        print(json.dumps({"n": n, "sweeps": sweeps, "ms_per_sweep": round(t1 / sweeps, 5),
                          "gbs_24B_per_point": round(24.0 * n * n * sweeps / t1 / 1e6, 1),
                          "blocked4_effective_gbs": round(24.0 * n * n * sweeps / t4 / 1e6, 1)}), flush=True)
        s.close()
else:
    import torch
    import torch.distributed as dist
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 32769
    # fp64 floor of the UN-SCALED residual norm: ~4 eps/h^2 per entry times sqrt(N^2) = 4.8e-8 ||r0|| at N = 32769
    # (1.2e-8 at 16385, which is why the reference's last cycles there slow down) -- 1e-8 is out of reach here
    tol = float(sys.argv[3]) if len(sys.argv) > 3 else (1e-8 if n <= 16385 else 1e-6)
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        rank, world, dev = pmg.init_distributed_from_torch(local)
        s = pmg.Solver(n, omega=2.0 / 3.0, device=dev, rank=rank, n_ranks=world)
    else:
        rank, dev = 0, local
        s = pmg.Solver(n, omega=2.0 / 3.0, device=dev)
    s.set_rhs_sine()
    best = None
    for _ in range(3):
        s.zero_guess()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
        k, hist = s.solve(pmg.V, tol, 100)
        best = s.last_ms if best is None else min(best, s.last_ms)
    if rank == 0:
        print(json.dumps({"n": n, "gpus": world, "dof_per_gpu_M": round(n * n / world / 1e6, 1), "cycles": k,
                          "solve_ms": round(best, 3), "ms_per_cycle": round(best / k, 4),
                          "gdof_cycle_per_s_per_gpu": round(n * n * k / best / 1e6 / world, 2),
                          "rel_tol": tol, "converged": bool(hist[-1] < tol * hist[0]),
                          "final_rel": float(hist[-1] / hist[0])}), flush=True)
    s.close()
    if world > 1:
        pmg.comm_finalize()
        dist.destroy_process_group()
