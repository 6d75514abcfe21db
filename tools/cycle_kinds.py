#!/usr/bin/env python
"""Time-to-1e-8 of the V-, W- (gamma = 2) and F+V strategies at N (BASELINE config 4), one GPU."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pmg_b200 as pmg  # noqa: E402

for n in [int(a) for a in sys.argv[1:]] or [4097, 16385]:
    for name, kind, gamma, prolong in (("V", pmg.V, 1, pmg.PROLONG_REFERENCE), ("W2", pmg.W, 2, pmg.PROLONG_REFERENCE),
                                       ("V_full", pmg.V, 1, pmg.PROLONG_FULL), ("W2_full", pmg.W, 2, pmg.PROLONG_FULL)):
        s = pmg.Solver(n, omega=2.0 / 3.0, gamma=gamma, prolong_mode=prolong)
        s.set_rhs_sine()
        best = None
        for _ in range(3):
            s.zero_guess()
            k, hist = s.solve(kind, 1e-8, 100)
            best = s.last_ms if best is None else min(best, s.last_ms)
        print(json.dumps({"n": n, "cycle": name, "cycles": k, "solve_ms": round(best, 3),
                          "ms_per_cycle": round(best / k, 4), "gdof_per_s": round(n * n / best / 1e6, 3)}), flush=True)
        s.close()
    # one FMG pass (reference F-cycle semantics), then V-cycles
    s = pmg.Solver(n, omega=2.0 / 3.0)
    s.set_rhs_sine()
    s.zero_guess()
    r0 = s.residual_norm()
    rn = s.cycle(pmg.F)
    t_f = s.last_ms
    k, hist = s.solve(pmg.V, 1e-8 * r0 / rn, 100)
    print(json.dumps({"n": n, "cycle": "F+V", "fmg_ms": round(t_f, 3), "v_cycles_after": k,
                      "total_ms": round(t_f + s.last_ms, 3), "rel_after_fmg": rn / r0}), flush=True)
    s.close()
