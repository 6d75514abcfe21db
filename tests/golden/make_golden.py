#!/usr/bin/env python
"""Generate tests/golden/golden.json + golden_fields.npz from the REAL reference.

Run in the build container (where /root/reference exists):
    make -C oracle && python tests/golden/make_golden.py
Every number below is produced by oracle/_ref/libpmg_ref.so, i.e. the unmodified reference headers
(Smoother.hpp, DynamicGridUtils.hpp, 2_part_MG/MultiGrid.hpp) behind oracle/ref_driver.cpp, except
the entries tagged "source": "oracle" (prolong_mode FULL does not exist in the reference) and the
block "survey" (values recorded in SURVEY.md section 8c from the same reference at sizes that take
the CPU ~20 min and 13 GB; they are copied, not regenerated).
Python's json writes doubles with repr(), which round-trips bit-exactly.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpu_checkers as cc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
TWO_THIRDS = 2.0 / 3.0


def operator_chain(ref, n):
    """SURVEY.md 8c 'per-operator known answers': smooth(num_iter=1) -> residual -> restrict ->
    prolongation of that coarse field into a zero fine field."""
    h = 1.0 / (n - 1)
    m = n // 2
    f = ref.rhs(n)
    x = np.zeros((n, n))
    sm = ref.jacobi(x, f, h, omega=1.0, num_iter=1)
    r = ref.residual(x, f, h)
    rc = ref.restrict_fw(r)
    mc = rc.shape[0] // 2
    p = ref.prolong_add(np.zeros((n, n)), rc)
    xw = np.zeros((n, n))
    smw = ref.jacobi(xw, f, h, omega=TWO_THIRDS, num_iter=1)
    return {
        "n": n,
        "f_norm": ref.norm(f), "f_mid": f[m, m],
        "x_norm": ref.norm(x), "x_mid": x[m, m], "x_11": x[1, 1],
        "smoother_residuals": list(sm),
        "r_mid": r[m, m], "r_11": r[1, 1],
        "rc_norm": ref.norm(rc), "rc_mid": rc[mc, mc], "rc_11": rc[1, 1],
        "p_norm": ref.norm(p), "p_11": p[1, 1], "p_12": p[1, 2], "p_22": p[2, 2], "p_23": p[2, 3],
        "p_33": p[3, 3], "p_last": p[n - 2, n - 2],
        "xw_norm": ref.norm(xw), "xw_mid": xw[m, m], "xw_11": xw[1, 1],
        "smoother_residuals_weighted": list(smw),
    }


def history(lib, source, n, kind, omega, eps, alpha, v1=1, v2=1, prolong=0, rhs="sine",
            max_cycles=100, rel_tol=1e-8, keep_field=None, fields=None):
    f = lib.rhs(n) if rhs == "sine" else cc.random_rhs(n)
    phi = np.zeros((n, n))
    k, hist = lib.solve(phi, f, kind=kind, omega=omega, eps=eps, alpha=alpha, v1=v1, v2=v2,
                        prolong=prolong, rel_tol=rel_tol, max_cycles=max_cycles)
    if keep_field is not None:
        fields[keep_field] = phi
    return {"source": source, "n": n, "kind": "VWF"[kind], "omega": omega, "eps": eps,
            "alpha": alpha, "v1": v1, "v2": v2, "prolong": ["reference", "full"][prolong],
            "rhs": rhs, "rel_tol": rel_tol, "max_cycles": max_cycles, "cycles": k,
            "hist": list(map(float, hist)), "field": keep_field}


def main():
    ref = cc.load("ref")
    orc = cc.load("orc")
    assert ref is not None, "needs oracle/_ref/libpmg_ref.so (build container only)"
    fields = {}
    g = {"generator": "tests/golden/make_golden.py over oracle/_ref/libpmg_ref.so",
         "operators": [operator_chain(ref, n) for n in (9, 33, 257)], "histories": []}
    H = g["histories"]
    V, W, F = cc.V, cc.W, cc.F
    # BASELINE config 1 family: V(2,2) [v1=v2=1], omega=2/3, eps=0, reference prolongation
    for n in (9, 17, 33, 65, 129, 257, 513, 1025):
        H.append(history(ref, "reference", n, V, TWO_THIRDS, 0.0, 2,
                         keep_field=("v_n33" if n == 33 else None), fields=fields))
    # as shipped: omega = 1, eps = 1e-7 (stalls; fixed cycle counts).  Run on the injected smoother
    # (identical arithmetic at omega = 1) because the shipped JacobiSmoother object forms its
    # per-sweep norm over an UNINITIALISED scratch ring (Smoother.hpp:75-77): with eps > 0 its early
    # exit then depends on what malloc recycled, i.e. the reference itself is not deterministic
    # there.  The zero-ring idealisation reproduces the SURVEY.md 8c values (fresh-process run).
    H.append(history(ref, "reference", 33, V, 1.0, 1e-7, 3, max_cycles=8))
    H.append(history(ref, "reference", 257, V, 1.0, 1e-7, 3, max_cycles=12))
    H.append(history(ref, "reference", 257, W, 1.0, 1e-7, 3, max_cycles=3))
    H.append(history(ref, "reference", 257, F, 1.0, 1e-7, 3, max_cycles=3))
    # W-cycles
    for n, alpha in ((33, 2), (257, 2), (257, 3), (1025, 2)):
        H.append(history(ref, "reference", n, W, TWO_THIRDS, 0.0, alpha))
    # F-cycle (one FMG pass per "cycle"; fixed point from pass 2 on)
    for n in (33, 129, 257):
        H.append(history(ref, "reference", n, F, TWO_THIRDS, 0.0, 2, max_cycles=3,
                         keep_field=("f_n33" if n == 33 else None), fields=fields))
    H.append(history(ref, "reference", 65, F, 1.0, 0.0, 2, max_cycles=2))
    # other smoothing counts (v1/v2 are the reference's num_iter: sweeps = v+1)
    H.append(history(ref, "reference", 129, V, TWO_THIRDS, 0.0, 2, v1=0, v2=0))
    H.append(history(ref, "reference", 129, V, TWO_THIRDS, 0.0, 2, v1=2, v2=1))
    H.append(history(ref, "reference", 129, V, 0.8, 0.0, 2, v1=1, v2=2))
    H.append(history(ref, "reference", 65, W, TWO_THIRDS, 0.0, 2, v1=0, v2=2))
    # random RHS B (all modes excited)
    H.append(history(ref, "reference", 129, V, TWO_THIRDS, 0.0, 2, rhs="random",
                     keep_field="v_n129_random", fields=fields))
    H.append(history(ref, "reference", 257, W, TWO_THIRDS, 0.0, 2, rhs="random"))
    # non-reference FULL prolongation: oracle only
    for n in (33, 257, 1025):
        H.append(history(orc, "oracle", n, V, TWO_THIRDS, 0.0, 2, prolong=1))
    H.append(history(orc, "oracle", 257, W, TWO_THIRDS, 0.0, 2, prolong=1))
    H.append(history(orc, "oracle", 129, V, TWO_THIRDS, 0.0, 2, prolong=1, rhs="random"))
    H.append(history(orc, "oracle", 129, F, TWO_THIRDS, 0.0, 2, prolong=1, max_cycles=2))

    # mg_cpu_exec stdout goldens (defaults: 1 cycle, alpha = 3, eps = 1e-7, omega = 1): rel. L2 error
    errs = []
    for n in (129, 257):
        u = ref.exact(n)
        row = {"n": n}
        for kind in (V, W, F):
            phi = np.zeros((n, n))
            ref.cycle(phi, ref.rhs(n), kind=kind, omega=1.0, eps=1e-7, alpha=3)
            row["VWF"[kind]] = ref.norm(phi - u) / ref.norm(u)
        errs.append(row)
    g["mg_cpu_exec_rel_l2_error"] = errs

    g["survey"] = {
        "note": "copied from SURVEY.md section 8c (reference CPU path, g++ -O2, sizes too large to "
                "regenerate in the test suite); RHS sine, phi0 = 0, v1=v2=1, omega=2/3, eps=0",
        "cycles_to_1e-8_V_reference": {"33": 23, "257": 29, "513": 31, "1025": 32, "2049": 34,
                                       "4097": 36, "8193": 37, "16385": 39},
        "r0_formula": "pi^2*(N-1)",
        "V_n4097_first3": [39182.0452741017, 77552.685219529783, 130798.97356449421],
        "V_n4097_last2": [0.00047268849590986588, 0.00021352456601704658],
        "V_n8193_first3": [78782.238010839792, 161684.30004397556, 294004.6097136052],
        "V_n8193_last2": [0.0015101061835096044, 0.00069339724665989921],
        "V_n16385": [158221.19135343342, 333879.48448476801, 647359.75515405566, 1000311.2651939917,
                     1246149.9398287286, 1301299.309749264, 1181337.3834970668, 959652.47266434645,
                     713169.03427657916, 493089.36124714522, 321324.54773962044, 199359.47729465924,
                     118706.22008903936, 68270.126750649521, 38120.476811930865, 20753.841917651975,
                     11055.393297277511, 5779.0331324319277, 2971.7221058022774, 1506.3830535142306,
                     754.05754620858909, 373.31209100798714, 183.02169393244483, 88.957662401183541,
                     42.90785987821203, 20.555617851129302, 9.7878015100906932, 4.6353386051236267,
                     2.1845691709396462, 1.025072753004413, 0.47911396485758606, 0.22314495599473488,
                     0.10359841414151541, 0.047962048767302798, 0.022155792583444578,
                     0.010231782551625945, 0.004762003844638419, 0.0023109223427605166,
                     0.0013007435948222939],
        "V_n16385_full_prolong": [34893.079651279775, 7422.5489501253105, 1610.2318444729481,
                                  350.18764377948997, 76.458311564607811, 16.81040557972462,
                                  3.7325944729478397, 0.83830546939087958, 0.19043426425278068,
                                  0.043709771292642315, 0.010152434656297307, 0.0025178806180221874,
                                  0.0010356132485140504],
        "W_alpha2_n16385": [4331297.8564307159, 1818075.2126792707, 570508.60323002306,
                            165509.24560334237, 46292.670416438879, 12690.33266364474,
                            3437.2637317742083, 923.9788767642774, 247.15120029175557,
                            65.890914328388334, 17.527043300908471, 4.6549554705881606,
                            1.234953944298721, 0.32738328523808441, 0.086743599708588082,
                            0.022988415722381098, 0.0061429722229677649, 0.0018220561263716014,
                            0.00095377204445674154],
    }
    with open(os.path.join(HERE, "golden.json"), "w") as fh:
        json.dump(g, fh, indent=1)
    np.savez_compressed(os.path.join(HERE, "golden_fields.npz"), **fields)
    print("wrote golden.json (%d histories) and golden_fields.npz (%s)" % (len(H), list(fields)))


if __name__ == "__main__":
    main()
