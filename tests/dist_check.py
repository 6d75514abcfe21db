#!/usr/bin/env python
"""Multi-GPU equivalence check, run under torchrun on a box with >= 2 B200s:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_check.py [N ...]

Every rank solves its row slab; rank 0 also solves the same problem on its own GPU alone.  Checks: same
cycle count, per-cycle norms equal to <= 1e-12 relative (only the summation order differs), and the
gathered iterate BIT-IDENTICAL to the single-GPU iterate (V, W and F cycles, sine and random RHS).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pmg_b200 as pmg  # noqa: E402


def timing(sizes, rank, world, dev):
    """solve time per cycle for the latency options (small-level kernel generation, halo prologue, middle graph) and,
    for the last combination, the phase trace of the distributed cycle on rank 0"""
    import ctypes
    L = pmg.lib()
    L.pmg_dist_trace_dump.argtypes = [ctypes.c_int]
    L.pmg_dist_trace_dump.restype = None
    L.pmg_dist_trace_enable.argtypes = [ctypes.c_int]
    L.pmg_dist_trace_enable.restype = None
    ref_hist = {}
    for n in sizes:
        # (small-level kernel generation, halo prologue, middle graph [+ no interior/boundary split], phase trace)
        combos = [(2, 0, 0, 0), (3, 0, 0, 0), (3, 1, 0, 0), (3, 1, 1, 0), (2, 1, 2, 0), (3, 1, 2, 0), (3, 1, 2, 1)]
        for small, prologue, graph, trace in combos:
            pmg.set_small_vcycle_version(small)
            pmg.set_halo_prologue(prologue)
            os.environ["PMG_MID_GRAPH"] = "1" if graph else "0"  # read by pmg_create
            if graph == 2:  # with the prologue the interior warps never wait: no need to split tall slabs
                os.environ["PMG_SPLIT_MIN_ROWS"] = "1000000000"
            else:
                os.environ.pop("PMG_SPLIT_MIN_ROWS", None)
            L.pmg_dist_trace_enable(trace)
            s = pmg.Solver(n, omega=2.0 / 3.0, device=dev, rank=rank, n_ranks=world)
            s.set_rhs_sine()
            ms = []
            for it in range(4):
                s.zero_guess()
                torch.cuda.synchronize()
                dist.barrier()
                k, hist = s.solve(pmg.V, 1e-8, 100)
                ms.append(s.last_ms)
                if it == 0 and trace:
                    L.pmg_dist_trace_dump(-1)  # drop the warm-up marks
            t = torch.tensor(ms[1:], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            # every option must reproduce the first combination's residual history bit for bit
            if ref_hist.get(n) is None:
                ref_hist[n] = hist.copy()
            hist_ok = bool(len(hist) == len(ref_hist[n]) and np.array_equal(hist, ref_hist[n]))
            if rank == 0:
                print("timing ranks=%d N=%d small_kernel=%d halo_prologue=%d mid_graph=%d trace=%d: cycles=%d hist_ok=%s ms %s "
                      "-> %.1f us/cycle" % (world, n, small, prologue, graph, trace, k, hist_ok,
                                            [round(float(v), 3) for v in t], 1e3 * float(t.min()) / k), flush=True)
            if trace:
                L.pmg_dist_trace_dump(0)
                dist.barrier()
            s.close()
        L.pmg_dist_trace_enable(0)
    pmg.set_small_vcycle_version(0)
    pmg.set_halo_prologue(True)  # the default since round 2
    os.environ.pop("PMG_MID_GRAPH", None)
    os.environ.pop("PMG_SPLIT_MIN_ROWS", None)


def main():
    args = sys.argv[1:]
    timing_n = []
    if "--timing" in args:  # --timing N1,N2: per-cycle times of the latency options + one phase trace per size
        i = args.index("--timing")
        timing_n = [int(v) for v in args[i + 1].split(",")]
        args = args[:i] + args[i + 2:]
    if "--prologue" in args:  # (the halo prologue of Pass A is the default since round 2; flag kept for old commands)
        args.remove("--prologue")
    if "--no-prologue" in args:  # run the equivalence checks with in-place streaming of the halo rows instead
        args.remove("--no-prologue")
        pmg.set_halo_prologue(False)
        print("halo prologue OFF", flush=True) if int(os.environ.get("RANK", "0")) == 0 else None
    sizes = [int(a) for a in args] or [1025, 4097]
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world, dev = pmg.init_distributed_from_torch(local)
    ok = True
    for n in sizes:
        for kind, gamma, rhs in ((pmg.V, 1, "sine"), (pmg.W, 2, "random"), (pmg.V, 1, "random")):
            y0, y1 = pmg.partition_rows(n, world, rank)
            rng = np.random.default_rng(7)
            if rhs == "random":
                f = np.zeros((n, n))
                f[1:-1, 1:-1] = rng.uniform(-1, 1, (n - 2, n - 2))
            else:
                x = np.sin(np.pi * np.arange(n) / (n - 1))
                f = 2 * np.pi ** 2 * np.outer(x, x)
            s = pmg.Solver(n, omega=2.0 / 3.0, gamma=gamma, device=dev, rank=rank, n_ranks=world,
                           agglomerate_below=min(257, n // 4))
            s.set_rhs(np.ascontiguousarray(f[y0:y1]))
            s.zero_guess()
            k, hist = s.solve(kind, rel_tol=1e-8, max_cycles=60)
            mine = torch.from_numpy(s.get_solution()).cuda()
            ms = s.last_ms
            s.close()
            parts = [torch.empty((pmg.partition_rows(n, world, r)[1] - pmg.partition_rows(n, world, r)[0], n),
                                 dtype=torch.float64, device="cuda") for r in range(world)]
            dist.all_gather(parts, mine)
            if rank == 0:
                full = torch.cat(parts).cpu().numpy()
                one = pmg.Solver(n, omega=2.0 / 3.0, gamma=gamma, device=dev)
                one.set_rhs(f)
                one.zero_guess()
                k1, h1 = one.solve(kind, rel_tol=1e-8, max_cycles=60)
                ref = one.get_solution()
                ms1 = one.last_ms
                one.close()
                same = np.array_equal(full, ref)
                rel = float(np.max(np.abs(hist - h1) / h1)) if k == k1 else float("inf")
                good = same and k == k1 and rel <= 1e-12
                ok &= good
                print("N=%d kind=%s rhs=%s ranks=%d: cycles %d/%d iterate_bit_identical=%s max_rel_norm_dev=%.2e "
                      "ms dist=%.3f single=%.3f %s" % (n, "VWF"[kind], rhs, world, k, k1, same, rel, ms, ms1,
                                                       "OK" if good else "FAIL"), flush=True)
    # F-cycle (the runner's full-multigrid wrapper, MultiGridTestRunner.hpp:192-205): two passes from a random start
    # whose ring is non-zero; iterate bit-identical to one GPU, the runner's residual norm to <= 1e-12
    for n in sizes:
        for prolong in (pmg.PROLONG_REFERENCE, pmg.PROLONG_FULL):
            y0, y1 = pmg.partition_rows(n, world, rank)
            x = np.sin(np.pi * np.arange(n) / (n - 1))
            f = 2 * np.pi ** 2 * np.outer(x, x)
            phi0 = np.random.default_rng(3).standard_normal((n, n))
            s = pmg.Solver(n, omega=2.0 / 3.0, device=dev, rank=rank, n_ranks=world, prolong_mode=prolong,
                           agglomerate_below=min(257, n // 4))
            s.set_rhs(np.ascontiguousarray(f[y0:y1]))
            s.set_guess(np.ascontiguousarray(phi0[y0:y1]))
            norms, ms = [], 0.0
            for _ in range(2):
                norms.append(s.cycle(pmg.F))
                ms = s.last_ms
            mine = torch.from_numpy(s.get_solution()).cuda()
            s.close()
            parts = [torch.empty((pmg.partition_rows(n, world, r)[1] - pmg.partition_rows(n, world, r)[0], n),
                                 dtype=torch.float64, device="cuda") for r in range(world)]
            dist.all_gather(parts, mine)
            if rank == 0:
                full = torch.cat(parts).cpu().numpy()
                one = pmg.Solver(n, omega=2.0 / 3.0, device=dev, prolong_mode=prolong)
                one.set_rhs(f)
                one.set_guess(phi0)
                n1 = [one.cycle(pmg.F) for _ in range(2)]
                ms1 = one.last_ms
                ref = one.get_solution()
                one.close()
                same = np.array_equal(full, ref)
                rel = max(abs(a - b) / b for a, b in zip(norms, n1))
                good = same and rel <= 1e-12
                ok &= good
                print("N=%d kind=F prolong=%d ranks=%d: iterate_bit_identical=%s max_rel_norm_dev=%.2e ms dist=%.3f "
                      "single=%.3f %s" % (n, prolong, world, same, rel, ms, ms1, "OK" if good else "FAIL"), flush=True)
    if timing_n:
        timing(timing_n, rank, world, dev)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    pmg.comm_finalize()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
