#!/usr/bin/env python
"""Programmatic dependent launch on / off: solve time per cycle on one GPU at latency-bound and bandwidth-bound sizes;
both settings must give the same bits.  python tools/pdl_probe.py [sizes...]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pmg_b200 as pmg  # noqa: E402

cases = [(257, pmg.V, 1), (1025, pmg.V, 1), (4097, pmg.V, 1), (1025, pmg.W, 2), (4097, pmg.W, 2), (16385, pmg.V, 1)]
ok = True
for n, kind, gamma in cases:
    ref = None
    for pdl in (0, 1, 0, 1):
        pmg.set_pdl(pdl)
        s = pmg.Solver(n, omega=2.0 / 3.0, gamma=gamma)
        s.set_rhs_sine()
        best = None
        for _ in range(4):
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
        print(json.dumps({"n": n, "cycle": "V" if kind == pmg.V else "W2", "pdl": pdl, "cycles": k,
                          "solve_ms": round(best, 4), "us_per_cycle": round(1e3 * best / k, 2),
                          "bit_identical_to_first": bool(same)}), flush=True)
pmg.set_pdl(1)
print("pdl_probe:", "OK" if ok else "MISMATCH", flush=True)
sys.exit(0 if ok else 1)
