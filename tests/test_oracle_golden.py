"""The oracle (plain-C restatement) against the committed goldens generated from the real reference
(tests/golden/make_golden.py) -- this is what PINS the oracle.  CPU only."""
import numpy as np
import pytest

import cpu_checkers as cc

KIND = {"V": cc.V, "W": cc.W, "F": cc.F}
PROLONG = {"reference": cc.PROLONG_REFERENCE, "full": cc.PROLONG_FULL}


def _run(lib, h):
    n = h["n"]
    f = lib.rhs(n) if h["rhs"] == "sine" else cc.random_rhs(n)
    phi = np.zeros((n, n))
    k, hist = lib.solve(phi, f, kind=KIND[h["kind"]], omega=h["omega"], eps=h["eps"], alpha=h["alpha"],
                        v1=h["v1"], v2=h["v2"], prolong=PROLONG[h["prolong"]], rel_tol=h["rel_tol"],
                        max_cycles=h["max_cycles"])
    return k, hist, phi


def test_operator_known_answers(orc, golden):
    """SURVEY.md 8c per-operator table: bit-exact (both sides built with -ffp-contract=off)."""
    for g in golden["operators"]:
        n = g["n"]
        h, m = 1.0 / (n - 1), n // 2
        f = orc.rhs(n)
        assert orc.norm(f) == g["f_norm"] and f[m, m] == g["f_mid"]
        x = np.zeros((n, n))
        sm = orc.jacobi(x, f, h, omega=1.0, num_iter=1)
        assert list(sm) == g["smoother_residuals"]
        assert (orc.norm(x), x[m, m], x[1, 1]) == (g["x_norm"], g["x_mid"], g["x_11"])
        r = orc.residual(x, f, h)
        assert (r[m, m], r[1, 1]) == (g["r_mid"], g["r_11"])
        rc = orc.restrict_fw(r)
        mc = rc.shape[0] // 2
        assert (orc.norm(rc), rc[mc, mc], rc[1, 1]) == (g["rc_norm"], g["rc_mid"], g["rc_11"])
        p = orc.prolong_add(np.zeros((n, n)), rc)
        assert p[1, 1] == 0.0 and p[1, 2] == 0.0, "reference prolongation skips fine row/col 1"
        assert (orc.norm(p), p[2, 2], p[2, 3], p[3, 3], p[n - 2, n - 2]) == (
            g["p_norm"], g["p_22"], g["p_23"], g["p_33"], g["p_last"])
        xw = np.zeros((n, n))
        smw = orc.jacobi(xw, f, h, omega=2.0 / 3.0, num_iter=1)
        assert list(smw) == g["smoother_residuals_weighted"]
        assert (orc.norm(xw), xw[m, m], xw[1, 1]) == (g["xw_norm"], g["xw_mid"], g["xw_11"])


def test_survey_recorded_values(orc, golden):
    """Values quoted in SURVEY.md 8c / BASELINE.md (independent of make_golden.py)."""
    assert orc.norm(orc.rhs(257)) == 2526.6187266788884
    assert orc.norm(orc.rhs(33)) == 315.82734083485929
    phi = np.zeros((257, 257))
    k, hist = orc.solve(phi, orc.rhs(257), kind=cc.V, omega=2.0 / 3.0, eps=0.0)
    assert k == 29
    assert list(hist[1:7]) == [2328.5378873946133, 3442.3370409628619, 3692.9077962260526,
                               3036.7198115267674, 2090.0880160348338, 1275.1619796586242]
    assert list(hist[27:30]) == [8.8570783621105275e-05, 3.8243083254876706e-05, 1.6507447860347508e-05]


@pytest.mark.parametrize("idx", range(32))
def test_history_bit_exact(orc, golden, idx):
    h = golden["histories"][idx]
    if h["n"] > 600:
        pytest.skip("large case covered by test_history_large")
    k, hist, phi = _run(orc, h)
    assert k == h["cycles"]
    assert list(hist) == h["hist"], "oracle history differs from the reference golden"
    if h["field"]:
        assert np.array_equal(phi, golden["fields"][h["field"]])


def test_history_large(orc, golden):
    for h in golden["histories"]:
        if h["n"] > 600 and h["kind"] == "V" and h["prolong"] == "reference":
            k, hist, _ = _run(orc, h)
            assert k == h["cycles"] and list(hist) == h["hist"]
            break


def test_mg_cpu_exec_stdout_errors(orc, golden):
    """`Final Relative L2 Error` of mg_cpu_exec (defaults: 1 cycle, alpha=3, eps=1e-7, omega=1)."""
    for row in golden["mg_cpu_exec_rel_l2_error"]:
        n = row["n"]
        u = orc.exact(n)
        for kind in "VWF":
            phi = np.zeros((n, n))
            orc.cycle(phi, orc.rhs(n), kind=KIND[kind], omega=1.0, eps=1e-7, alpha=3)
            assert orc.norm(phi - u) / orc.norm(u) == row[kind]
    r257 = [r for r in golden["mg_cpu_exec_rel_l2_error"] if r["n"] == 257][0]
    assert "%.6g" % r257["V"] == "0.171973" and "%.6g" % r257["W"] == "0.000650242"
    assert "%.6g" % r257["F"] == "0.000344973"
