#!/usr/bin/env python
"""GPU tuning sweep of the fused streaming kernels: every variant x {Pass A, Pass A zero-x, Pass B+norm,
Pass B, single Jacobi sweep} at a few sizes; prints ms and algorithmic GB/s (26 / 18 / 26 / 26 / 24 B per
point).  Usage (on a B200): python tools/tune_fused.py [--rows R ...] [N ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pmg_b200 as pmg  # noqa: E402

args = sys.argv[1:]
rows_list = [8]
if args and args[0] == "--rows":
    rows_list = [int(a) for a in args[1].split(",")]
    args = args[2:]
sizes = [int(a) for a in args] or [4097, 16385]
names = {0: ("passA", 26.0), 3: ("passA_zero", 18.0), 1: ("passB_norm", 26.0), 2: ("passB", 26.0)}
for n in sizes:
    for rows in rows_list:
        pmg.lib().pmg_fused_set_min_chunk_rows(rows)
        for v in range(pmg.num_fused_variants()):
            pmg.set_fused_variant(v)
            s = pmg.Solver(n, omega=2.0 / 3.0)
            s.set_rhs_sine()
            s.zero_guess()
            rec = {"n": n, "variant": v, "min_rows": rows}
            for which, (nm, bpp) in names.items():
                ms = s.bench_pass(which, 0, 5)
                rec[nm + "_ms"] = round(ms, 4)
                rec[nm + "_gbs"] = round(bpp * n * n / ms / 1e6, 1)
            s.smooth(2, 1)
            s.smooth(10, 1)
            rec["jacobi1_gbs"] = round(24.0 * n * n * 10 / s.last_ms / 1e6, 1)
            s.zero_guess()
            k, hist = s.solve(pmg.V, 1e-8, 100)
            k, hist = s.solve(pmg.V, 1e-30, 20)
            rec["ms_per_cycle"] = round(s.last_ms / k, 4)
            rec["vcycle_gbs_69.3"] = round(69.3 * n * n * k / s.last_ms / 1e6, 1)
            s.close()
            print(json.dumps(rec), flush=True)
pmg.set_fused_variant(-1)
