"""bench_dist.py -- the N > 1 leg of bench.py: one process per GPU (torchrun), strong scaling of the same
N = 16385 solve over row slabs.  torch.distributed (NCCL) is plumbing only: rendezvous, the barrier around
the timed region, the MAX over ranks of the elapsed time.  All solver traffic (halo rows, the agglomerated
coarse level, the norm all-gather) goes through libpmg's own NCCL communicator inside pmg_solve.
"""
import json
import os
import time

import numpy as np

import bench


def pin_to_gpu_numa_node(local):
    """Run this rank on the CPU cores next to ITS GPU before anything is allocated, so that the pinned host buffers
    of the e2e leg are first touched -- hence placed -- on the GPU's own NUMA node.  (Round 1: 8 ranks with default
    placement reached 16.8 GB/s per rank over PCIe.)"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "rank pinned to the %d cores NVML lists for GPU %d (%d-%d)" % (len(cpus), local, cpus[0], cpus[-1])
        return "NVML affinity empty; default placement"
    except Exception as e:  # noqa: BLE001
        return "default placement (%s)" % repr(e)[:80]


def run(args):
    import torch
    import torch.distributed as dist
    import pmg_b200 as pmg

    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa_note = pin_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world, dev = pmg.init_distributed_from_torch(local)
    n = args.n
    peak, peak_src = bench.hbm_peak()
    prolong = pmg.PROLONG_FULL if args.prolong == "full" else pmg.PROLONG_REFERENCE
    s = pmg.Solver(n, omega=bench.OMEGA, prolong_mode=prolong, device=dev, rank=rank, n_ranks=world,
                   agglomerate_below=args.agglomerate_below)
    y0, y1 = s.local_rows
    ny = y1 - y0
    cross_cycle = s.cross_cycle
    s.set_rhs_sine()
    max_cycles = 100

    def step():
        s.zero_guess()
        return s.solve(pmg.V, rel_tol=bench.REL_TOL, max_cycles=max_cycles)

    def fence():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        k, hist = step()
    fence()
    clocks = bench.ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = pmg.kernel_launches()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        k, hist = step()
        dev_ms += s.last_ms
    fence()
    wall = time.perf_counter() - t0
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([wall, dev_ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall, dev_ms = float(t[0]), float(t[1])
    lt = torch.tensor([pmg.kernel_launches() - launches0], dtype=torch.float64, device="cuda")
    dist.all_reduce(lt, op=dist.ReduceOp.SUM)

    # dominant kernel on this rank's slab, timed alone (no communication inside)
    t_down = s.bench_pass(0, 0, 5)
    t_upn = s.bench_pass(1, 0, 5)
    tt = torch.tensor([t_down, t_upn], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_down, t_upn = float(tt[0]), float(tt[1])

    # e2e: every rank pushes ITS slab of f from pinned host memory and pulls its slab of phi back
    def fill_f(out):
        h = 1.0 / (n - 1)
        sx = np.sin(np.pi * (np.arange(n) * h))
        np.multiply((2.0 * np.pi * np.pi * sx)[None, :], sx[y0:y1, None], out=out)

    def reduce_max_early(v):
        tv = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        return float(tv[0])

    e2e = bench.e2e_measure(pmg, s, n, ny, args, fill_f, max_cycles, fence, reduce_max_early)
    e2e["d2h_bytes_per_step"] = n * n * 8 + (e2e["cycles"] + 1) * 8 * world
    e2e["numa"] = numa_note
    s.close()

    # the other BASELINE configs on the same ranks (W and F at N = 16385, N = 4097, the N = 32769 weak-scaling point)
    def reduce_max(v):
        tv = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        return float(tv[0])

    legs = None
    if not args.no_legs:
        legs = bench.config_legs(lambda nn, **cfg: pmg.Solver(nn, omega=bench.OMEGA, prolong_mode=prolong, device=dev,
                                                              rank=rank, n_ranks=world,
                                                              agglomerate_below=args.agglomerate_below, **cfg),
                                 peak, world, reduce_max)
    pmg.comm_finalize()

    if rank == 0:
        alg_bytes = bench.BYTES_PER_POINT_PASS * n * ny
        dom_ms, dom_name = (t_upn, "k_up<nu2=2,prolong,norm>") if t_upn >= t_down else (t_down, "k_down<nu1=2,resid>")
        achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
        cycle_gbs = bench.BYTES_PER_DOF_CYCLE * n * n * k / (dev_ms / args.steps * 1e-3) / 1e9
        gold = bench.goldens()
        parity = bench.history_check(hist[1:], gold.get(("V_n16385_full" if args.prolong == "full" else "V_n16385")
                                                        if n == 16385 else "V_n%d" % n))
        line = {"metric": bench.METRIC, "value": n * n / (wall / args.steps) / 1e9, "unit": bench.UNIT,
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": bench.workload_config(args, world),
                "cycles_to_converge": k, "converged": bool(hist[-1] < bench.REL_TOL * hist[0]),
                "final_rel_residual": float(hist[-1] / hist[0]),
                "cycles_match": parity["cycles_match"], "history_max_rel_dev": parity["history_max_rel_dev"],
                "history_parity": parity, "cross_cycle_pass": cross_cycle, "legs": legs,
                "gdof_cycle_per_s": n * n * k / (dev_ms / args.steps * 1e-3) / 1e9,
                "device_ms_per_step": dev_ms / args.steps,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": None, "kernel": dom_name + " on one rank's slab",
                             "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": dom_ms, "peak_source": peak_src,
                             "pass_down_ms": t_down, "pass_up_norm_ms": t_upn,
                             "vcycle_effective_gbs_at_69.3B_per_dof": cycle_gbs,
                             "vcycle_frac_of_aggregate_peak": cycle_gbs / (peak * world)},
                "cpu_baseline": None,
                "e2e": e2e,
                "gpu_launches": int(lt[0]), "clocks": clk}
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0
