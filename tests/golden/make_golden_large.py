#!/usr/bin/env python
"""Golden residual histories at the BASELINE sizes that take the CPU minutes: tests/golden/golden_large.json.

Run in the build container (where /root/reference exists), after `make -C oracle`:
    python tests/golden/make_golden_large.py [--with-16385]
Every number is produced by oracle/_ref/libpmg_ref.so, i.e. the unmodified reference headers behind
oracle/ref_driver.cpp (MultigridSolver::{v,w,f}_cycle + the omega-weighted Smoother subclass, epsilon = 0,
v1 = v2 = 1, reference prolongation, RHS A, phi0 = 0).  bench.py compares the GPU histories of its N = 4097 leg
(BASELINE config 2) and of its W / F legs against these; SURVEY.md 8c only recorded the first / last values.
  V_n4097            36 values + r0    (56 s)
  W_alpha2_n4097     W(gamma = 2) history to 1e-8
  F_n4097            ||f - A phi|| after 1 and 2 passes of the runner's F-cycle wrapper, phi0 = 0
  F_n16385           the same after 1 pass at N = 16385 (--with-16385: ~1 min, 15 GB)
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpu_checkers as cc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
W23 = 2.0 / 3.0


def main():
    ref = cc.load("ref")
    if ref is None:
        raise SystemExit("needs oracle/_ref/libpmg_ref.so (build container)")
    out_path = os.path.join(HERE, "golden_large.json")
    out = json.load(open(out_path)) if os.path.exists(out_path) else {}
    out["generator"] = "tests/golden/make_golden_large.py over oracle/_ref/libpmg_ref.so (unmodified reference headers)"

    def timed(name, fn):
        if name in out:
            return
        t0 = time.time()
        out[name] = fn()
        out[name]["seconds"] = round(time.time() - t0, 1)
        print(name, out[name]["seconds"], "s", flush=True)
        json.dump(out, open(out_path, "w"), indent=1)

    def hist(n, kind, alpha):
        def run():
            f = ref.rhs(n)
            phi = np.zeros((n, n))
            k, h = ref.solve(phi, f, kind=kind, omega=W23, eps=0.0, alpha=alpha, rel_tol=1e-8, max_cycles=100)
            return {"n": n, "cycles": int(k), "history": [float(v) for v in h]}
        return run

    def fpass(n, passes):
        def run():
            f = ref.rhs(n)
            phi = np.zeros((n, n))
            norms = []
            for _ in range(passes):
                ref.cycle(phi, f, kind=cc.F, omega=W23, eps=0.0, alpha=1)
                norms.append(float(ref.norm(ref.residual(phi, f, 1.0 / (n - 1)))))
            u = ref.exact(n)
            return {"n": n, "norms": norms, "rel_l2_error": float(ref.norm(phi - u) / ref.norm(u))}
        return run

    timed("V_n4097", hist(4097, cc.V, 1))
    timed("W_alpha2_n4097", hist(4097, cc.W, 2))
    timed("F_n4097", fpass(4097, 2))
    if "--with-16385" in sys.argv:
        timed("F_n16385", fpass(16385, 1))


if __name__ == "__main__":
    main()
