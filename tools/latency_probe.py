#!/usr/bin/env python
"""Latency-bound levels, one GPU: solve time per cycle for the combinations of
  * the single-CTA small-level kernel generation (1 / 2), and
  * the deep-prefetch variant of the fused passes for levels n <= threshold,
at sizes where the whole cycle is latency-bound (N = 257 ... 4097), V and W(gamma = 2), plus the headline sizes.
(the second tuple entry is now the top level of the 16-CTA cluster kernel: 0 = off)
Every combination must give the same bits (checked against the first one); prints one JSON line per run.

  python tools/latency_probe.py [quick]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pmg_b200 as pmg  # noqa: E402

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
combos = [(2, 0), (3, 0), (3, 129), (3, 257)]  # (small-kernel generation, cluster-kernel top level)
cases = [(257, pmg.V, 1), (1025, pmg.V, 1), (4097, pmg.V, 1), (1025, pmg.W, 2), (4097, pmg.W, 2)]
if not quick:
    cases += [(16385, pmg.V, 1), (16385, pmg.W, 2)]
ok = True
for n, kind, gamma in cases:
    ref = None
    for small, deep in combos:
        if n <= deep:
            continue
        pmg.set_small_vcycle_version(small)
        pmg.set_cluster_top(deep)
        s = pmg.Solver(n, omega=2.0 / 3.0, gamma=gamma)
        s.set_rhs_sine()
        best = None
        reps = 2 if (n == 16385 and kind == pmg.W) else 4
        for _ in range(reps):
            s.zero_guess()
            k, hist = s.solve(kind, 1e-8, 100)
            best = s.last_ms if best is None else min(best, s.last_ms)
        phi = s.get_solution() if n <= 4097 else None
        s.close()
        same = True
        if ref is None:
            ref = (k, hist, phi)
        else:
            same = (k == ref[0]) and np.array_equal(hist, ref[1]) and (phi is None or np.array_equal(phi, ref[2]))
        ok &= same
        print(json.dumps({"n": n, "cycle": "V" if kind == pmg.V else "W2", "small_kernel": small,
                          "cluster_top": deep, "cycles": k, "solve_ms": round(best, 4),
                          "us_per_cycle": round(1e3 * best / k, 2), "bit_identical_to_first": bool(same)}), flush=True)
pmg.set_small_vcycle_version(0)
pmg.set_cluster_top(-1)
print("latency_probe:", "OK" if ok else "MISMATCH", flush=True)
sys.exit(0 if ok else 1)
