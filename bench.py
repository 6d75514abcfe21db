#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

metric   V-cycle GDOF/s to a 1e-8 relative residual: N^2 / time(phi0 = 0 -> ||r|| < 1e-8 ||r0||)
workload 2-D Poisson, N = 16385 (268M DOF), V(2,2) weighted Jacobi omega = 2/3, fp64, RHS A
         (f = 2 pi^2 sin(pi x) sin(pi y), synthetic), reference-parity prolongation => 39 cycles.
step     one complete solve (reset phi to 0, cycle until converged, residual norm every cycle).

  python bench.py [--gpus N] [--steps K] [--warmup W]          our CUDA path
  python bench.py --impl reference ...                         the reference's CPU path (oracle/_ref)

`value` is measured with f resident in HBM; `e2e` goes through the C ABI with pinned HOST buffers
(H2D of f and phi0, the solve, D2H of phi inside the timed region).  `roofline` is the dominant kernel
(the fused level-0 pass) timed alone with CUDA events on the solver stream against the measured HBM
copy peak.  Only the `cpu_baseline` / `--impl reference` legs touch oracle/ (as the CPU baseline).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

METRIC = "V-cycle GDOF/s to 1e-8 residual"
UNIT = "GDOF/s"
OMEGA = 2.0 / 3.0
REL_TOL = 1e-8
BYTES_PER_DOF_CYCLE = 69.3   # 52 B per level-point * 1.3334 (fused two-pass minimum)
BYTES_PER_POINT_PASS = 26.0  # one fused pass: read x, read f, write x', + coarse array traffic (2 B)


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_prefix, n):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture (profiles/r1_ncu_full_fused_passes_n16385.json, level-0 passes at N = 16385)."""
    if n != 16385:
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "r1_ncu_full_fused_passes_n16385.json")) as fh:
            recs = json.load(fh)
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        tot = []
        for r in recs:
            if r["kernel"].startswith(kernel_prefix):
                v = 0.0
                for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    num, u = r[key].split()
                    v += float(num) * unit[u]
                tot.append(v)
        return sum(tot) / len(tot) if tot else None
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (profiling recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 9 and r[5 + k].lower() == "active" for r in self.rows)]
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows), "power_w_max": max(pw) if pw else None}


def pinned(pmg, shape):
    """numpy view over pinned host memory obtained through the C ABI."""
    nbytes = int(np.prod(shape)) * 8
    p = ctypes.c_void_p()
    pmg.check(pmg.lib().pmg_host_alloc_pinned(ctypes.byref(p), nbytes))
    buf = (ctypes.c_double * (nbytes // 8)).from_address(p.value)
    return np.frombuffer(buf, dtype=np.float64).reshape(shape), p


def sine_rhs(n, out):
    """DynamicGridUtils::compute_rhs with a = p = q = 1, via its separable form (same rounding order:
    (factor * sin(pi x)) * sin(pi y))."""
    h = 1.0 / (n - 1)
    sx = np.sin(1.0 * np.pi * (np.arange(n) * h) / 1.0)
    factor = (np.pi * np.pi / 1.0) * 2.0
    np.multiply((factor * sx)[None, :], sx[:, None], out=out)


# ------------------------------------------------------------------------------------------------------
def cpu_reference_sample(n_sample, cycles, n_target, cycles_target):
    """Time `cycles` V(2,2) cycles of the reference CPU path at n_sample and scale by DOF*cycles to the
    target workload.  Returns (GDOF/s to solution at n_target, seconds measured, kind)."""
    import cpu_checkers as cc
    lib = cc.load("ref")
    kind = "reference"
    if lib is None:
        lib, kind = cc.load("orc"), "port"
    f = lib.rhs(n_sample)
    phi = np.zeros((n_sample, n_sample))
    t0 = time.perf_counter()
    for _ in range(cycles):
        lib.cycle(phi, f, kind=cc.V, omega=OMEGA, eps=0.0, alpha=1, v1=1, v2=1)
    dt = time.perf_counter() - t0
    dof_cycles_per_s = n_sample * n_sample * cycles / dt
    t_target = n_target * n_target * cycles_target / dof_cycles_per_s
    return n_target * n_target / t_target / 1e9, dt, kind


def reference_cuda_build(sizes=((4097, 3), (16385, 2))):
    """The reference's own CUDA multigrid (3_part_parallel, unmodified, compiled for sm_100a into
    oracle/_ref/ref_gpu_exec) timed on this GPU, ParallelTestRunner::run_v_cycle protocol, data prefetched.
    Time per V-cycle only: it has no convergence test, no omega, a racy in-place smoother, and leaks ~16 N^2
    bytes of managed memory per cycle (hence the small cycle counts)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_gpu_exec")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/ref_gpu_exec not built"}
    out = {}
    for n, cycles in sizes:
        try:
            p = subprocess.run([exe, str(n), str(cycles), "1"], capture_output=True, text=True, timeout=180)
            rec = json.loads(p.stdout.strip().splitlines()[-1])
            out[str(n)] = {"cycle_ms": rec["cycle_ms"], "best_cycle_ms": min(rec["cycle_ms"]),
                           "gdof_cycle_per_s": n * n / (min(rec["cycle_ms"]) * 1e-3) / 1e9}
        except Exception as e:  # noqa: BLE001
            out[str(n)] = {"error": repr(e)[:200]}
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores.  Each step
    is a bounded sample (one V(2,2) cycle at N = 4097, ~2 s) scaled by DOF*cycles to the 39-cycle solve at
    N = 16385.  The reference is single-threaded (no OpenMP/threads anywhere): cores = 1."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n_s, cyc_s = 4097, 1
    for _ in range(args.warmup):
        cpu_reference_sample(n_s, cyc_s, args.n, args.cycles_expected)
    vals, secs, kind = [], 0.0, "reference"
    for _ in range(args.steps):
        v, dt, kind = cpu_reference_sample(n_s, cyc_s, args.n, args.cycles_expected)
        vals.append(v)
        secs += dt
    value = len(vals) / sum(1.0 / v for v in vals)  # total work / total time
    ms_per_step = 1e3 * args.n * args.n / (value * 1e9)
    sample = ("%d V(2,2) cycle(s) at N=%d per step, scaled by DOF*cycles to %d cycles at N=%d"
              % (cyc_s, n_s, args.cycles_expected, args.n))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                             "host_cores_available": os.cpu_count(), "measured_seconds": secs},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def workload_config(args, n_gpus):
    return {"workload": "2D Poisson N=%d (%.1fM DOF) V(2,2) weighted-Jacobi omega=2/3 fp64 to rel residual 1e-8, "
                        "RHS 2pi^2 sin(pi x)sin(pi y), phi0=0" % (args.n, args.n * args.n / 1e6),
            "n": args.n, "cycle": "V(2,2)", "omega": OMEGA, "rel_tol": REL_TOL,
            "prolongation": args.prolong, "engine": "fused", "l2": "inputs_exceed_l2 (2.1 GB per array)",
            "parallelism": "1 GPU" if n_gpus == 1 else "row slabs x%d" % n_gpus}


# ------------------------------------------------------------------------------------------------------
def run_single(args):
    import pmg_b200 as pmg
    if pmg.device_count() < 1:
        raise SystemExit("bench.py needs a B200: libpmg.so has no CPU fallback")
    n = args.n
    peak, peak_src = hbm_peak()
    prolong = pmg.PROLONG_FULL if args.prolong == "full" else pmg.PROLONG_REFERENCE
    s = pmg.Solver(n, omega=OMEGA, prolong_mode=prolong, device=0)
    s.set_rhs_sine()
    max_cycles = 100

    def step():
        s.zero_guess()
        return s.solve(pmg.V, rel_tol=REL_TOL, max_cycles=max_cycles)

    for _ in range(args.warmup):
        k, hist = step()
    pmg.check(pmg.lib().pmg_device_synchronize())
    clocks = ClockSampler(0)
    clocks.start()
    launches0 = pmg.kernel_launches()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        k, hist = step()
        dev_ms += s.last_ms
    pmg.check(pmg.lib().pmg_device_synchronize())
    wall = time.perf_counter() - t0
    launches = pmg.kernel_launches() - launches0
    clk = clocks.stop()
    ms_per_step = 1e3 * wall / args.steps
    value = n * n / (wall / args.steps) / 1e9
    converged = bool(hist[-1] < REL_TOL * hist[0])

    # ---- roofline of the dominant kernel (level-0 fused passes), timed alone with CUDA events ----
    t_down = s.bench_pass(0, 0, 5)
    t_upn = s.bench_pass(1, 0, 5)
    alg_bytes = BYTES_PER_POINT_PASS * n * n
    dom_ms, dom_name = (t_upn, "k_up<nu2=2,prolong,norm>") if t_upn >= t_down else (t_down, "k_down<nu1=2,resid>")
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    traffic = ncu_traffic("k_up" if t_upn >= t_down else "k_down", n)
    cycle_gbs = BYTES_PER_DOF_CYCLE * n * n * k / (dev_ms / args.steps * 1e-3) / 1e9
    # Jacobi sweep sub-metric: one HBM pass per sweep, 24 B/point
    s.smooth(3, 1)
    s.smooth(10, 1)
    jac_gbs = 24.0 * n * n * 10 / (s.last_ms * 1e-3) / 1e9
    s.smooth(12, 4)
    jac_blocked_gbs = 24.0 * n * n * 12 / (s.last_ms * 1e-3) / 1e9

    # ---- e2e through the C ABI with pinned HOST buffers ----
    f_host, pf = pinned(pmg, (n, n))
    x_host, px = pinned(pmg, (n, n))
    out_host, po = pinned(pmg, (n, n))
    sine_rhs(n, f_host)
    x_host[:] = 0.0

    def e2e_step():
        s.set_rhs(f_host)
        s.set_guess(x_host)
        kk, hh = s.solve(pmg.V, rel_tol=REL_TOL, max_cycles=max_cycles)
        s.get_solution(out_host)
        return kk

    e2e_step()
    e_steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(e_steps):
        ke = e2e_step()
    e_wall = (time.perf_counter() - t0) / e_steps
    e2e = {"value": n * n / e_wall / 1e9, "unit": UNIT, "h2d_bytes_per_step": 2 * n * n * 8,
           "d2h_bytes_per_step": n * n * 8 + (ke + 1) * 8, "ms_per_step": 1e3 * e_wall, "cycles": ke}
    for p in (pf, px, po):
        pmg.lib().pmg_host_free_pinned(p)
    s.close()

    # ---- CPU baseline (bounded sample of the same workload on the host cores) ----
    cpu = None
    if not args.no_cpu_baseline:
        v, secs, kind = cpu_reference_sample(4097, 4, n, k)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": kind,
               "sample": "4 V(2,2) cycles at N=4097 (%.1f s), scaled by DOF*cycles to %d cycles at N=%d; the "
                         "reference is single-threaded" % (secs, k, n),
               "host_cores_available": os.cpu_count()}

    ref_cuda = None if args.no_ref_cuda else reference_cuda_build()
    if ref_cuda and str(n) in ref_cuda and "best_cycle_ms" in ref_cuda[str(n)]:
        ref_cuda["ours_cycle_ms_same_n"] = dev_ms / args.steps / k
        ref_cuda["speedup_per_cycle_same_n"] = ref_cuda[str(n)]["best_cycle_ms"] / (dev_ms / args.steps / k)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, 1),
            "cycles_to_converge": k, "converged": converged, "final_rel_residual": float(hist[-1] / hist[0]),
            "gdof_cycle_per_s": n * n * k / (dev_ms / args.steps * 1e-3) / 1e9,
            "device_ms_per_step": dev_ms / args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "kernel": dom_name,
                         "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": dom_ms, "peak_source": peak_src,
                         "pass_down_ms": t_down, "pass_up_norm_ms": t_upn,
                         "vcycle_effective_gbs_at_69.3B_per_dof": cycle_gbs, "vcycle_frac": cycle_gbs / peak},
            "jacobi_sweep": {"gbs_24B_per_point": jac_gbs, "frac": jac_gbs / peak,
                             "blocked4_effective_gbs": jac_blocked_gbs},
            "cpu_baseline": cpu, "reference_cuda_build": ref_cuda, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clk}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=16385)
    ap.add_argument("--prolong", default="reference", choices=["reference", "full"])
    ap.add_argument("--cycles-expected", type=int, default=39, dest="cycles_expected")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip timing the reference's own CUDA build")
    ap.add_argument("--agglomerate-below", type=int, default=513, dest="agglomerate_below",
                    help="multi-GPU: levels with n <= this run on rank 0 only")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 or args.gpus > 1:
        import bench_dist
        return bench_dist.run(args)
    return run_single(args)


if __name__ == "__main__":
    sys.exit(main())
