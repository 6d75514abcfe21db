#!/usr/bin/env python
"""A/B of two builds of libpmg.so on the SAME box, alternating, one process per measurement:

    python tools/ab_lib.py ab/libpmg_head.so parallel-geometric-multigrid-for-poisson-problem_b200/libpmg.so

Per library and size: best and median solve time of V(2,2) to 1e-8 (sine RHS), the cycle count and the final residual
norm -- which must agree between the builds (the iterates are bit-identical by construction; the norm is the check).
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import json, sys, statistics
sys.path.insert(0, %(root)r)
import pmg_b200 as pmg
sys.modules["_pmg_b200_pkg"].LIB_PATH = %(lib)r  # before the first call loads the library
out = {}
for n, reps, kind in ((16385, 8, pmg.V), (4097, 30, pmg.V), (1025, 50, pmg.V), (4097, 3, pmg.W)):
    s = pmg.Solver(n, omega=2.0 / 3.0)
    s.set_rhs_sine()
    ts = []
    for _ in range(reps):
        s.zero_guess()
        k, hist = s.solve(kind, 1e-8, 100)
        ts.append(s.last_ms)
    out["%%s%%d" %% ("V" if kind == pmg.V else "W", n)] = {"cycles": int(k), "best_ms": round(min(ts), 4),
        "median_ms": round(statistics.median(ts[1:]), 4), "final_norm": repr(float(s.residual_norm()))}
    s.close()
print(json.dumps(out))
"""


def main():
    libs = [os.path.abspath(p) for p in sys.argv[1:]]
    rounds = int(os.environ.get("AB_ROUNDS", "2"))
    res = {}
    for r in range(rounds):
        for lib in libs:
            p = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "lib": lib}], capture_output=True, text=True)
            if p.returncode != 0:
                print(p.stderr[-2000:])
                sys.exit(1)
            d = json.loads(p.stdout.strip().splitlines()[-1])
            print(json.dumps({"round": r, "lib": os.path.relpath(lib, ROOT), **d}), flush=True)
            res.setdefault(lib, []).append(d)
    a, b = libs[0], libs[-1]
    for key in res[a][0]:
        ta = min(x[key]["median_ms"] for x in res[a])
        tb = min(x[key]["median_ms"] for x in res[b])
        same = res[a][0][key]["final_norm"] == res[b][0][key]["final_norm"] and res[a][0][key]["cycles"] == res[b][0][key]["cycles"]
        print("%-7s %9.4f -> %9.4f ms  (%+.1f %%)  same result: %s" % (key, ta, tb, 100.0 * (tb / ta - 1.0), same), flush=True)


if __name__ == "__main__":
    main()
