"""include/pmg.hpp: the C++ class shims that mirror the reference surface.  tests/cpp/test_shims.cpp is a
reference-style driver (MultigridTestRunner / ParallelTestRunner protocols) written against the header only;
it is compiled with plain g++ here and run through libpmg.so."""
import json
import os
import subprocess

import numpy as np
import pytest

import pmg_b200 as pmg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.dirname(pmg.LIB_PATH)
EXE = os.path.join(ROOT, "tests", "cpp", "test_shims")


def _build():
    src = os.path.join(ROOT, "tests", "cpp", "test_shims.cpp")
    if os.path.exists(EXE) and os.path.getmtime(EXE) > max(os.path.getmtime(src), os.path.getmtime(pmg.LIB_PATH)):
        return
    pmg.lib()
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"), src, "-o", EXE,
                    "-L" + PKG, "-lpmg", "-Wl,-rpath," + PKG], check=True)


def _run():
    _build()
    p = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    return p.returncode, [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]


def test_shims_compile_and_fail_loudly_without_gpu():
    """No torch / CUDA types in the header: plain g++ builds it.  Without a GPU the driver must stop with
    PMG_ERR_NO_DEVICE -- there is no CPU fallback behind the reference-shaped classes."""
    _build()
    if pmg.device_count() > 0:
        pytest.skip("a GPU is visible")
    rc, out = _run()
    assert rc == 3 and out[-1]["test"] == "error" and out[-1]["status"] == 3


@pytest.mark.gpu
def test_reference_style_driver_matches_goldens(golden):
    rc, out = _run()
    assert rc == 0, out
    by = {}
    for o in out:
        by.setdefault(o["test"], []).append(o)
    # mg_cpu_exec stdout ("Final Relative L2 Error", 1 cycle, alpha=3, eps=1e-7)
    want = {r["n"]: r for r in golden["mg_cpu_exec_rel_l2_error"]}
    assert len(by["mg_cpu_exec"]) == 8
    for o in by["mg_cpu_exec"]:
        # "F_runner_shape": compute_coarsest_grid + f_cycle(phi_coarse, f_coarse, n_coarse, h_coarse) called exactly as
        # MultiGridTestRunner.hpp:192-205 does (3.44973e-4 at N = 257)
        w = want[o["n"]]["F" if o["cycle"] == "F_runner_shape" else o["cycle"]]
        assert abs(o["rel_l2_error"] - w) <= 1e-12 * w
    assert all(o["prolong_mode_change_seen"] for o in by["knobs"]) and len(by["knobs"]) == 2
    # residual history through pmg::Solver::solve (BASELINE config 1)
    h = by["history"][0]
    g = [x for x in golden["histories"] if x["n"] == 257 and x["kind"] == "V" and x["eps"] == 0.0
         and x["prolong"] == "reference" and x["rhs"] == "sine" and x["v1"] == 1 and abs(x["omega"] - 2 / 3) < 1e-12][0]
    assert h["cycles"] == g["cycles"] == 29
    assert np.max(np.abs(np.array(h["hist"]) - np.array(g["hist"])) / np.array(g["hist"])) <= 1e-10
    # class Parallel operators: SURVEY.md 8c known answers, bit exact
    ops = {o["n"]: o for o in golden["operators"]}
    for o in by["parallel_ops"]:
        g = ops[o["n"]]
        for k in ("x_mid", "x_11", "r_mid", "r_11", "rc_mid", "rc_11", "p_11", "p_22", "p_23", "p_33"):
            assert o[k] == g[k], (o["n"], k)
    # Smoother::smooth(num_iter = 1) -> 2 sweeps, per-sweep residuals
    sm = by["smoother"][0]
    g = ops[33]
    assert sm["sweeps"] == 2 and sm["x_mid"] == g["x_mid"]
    assert abs(sm["res0"] - g["smoother_residuals"][0]) <= 1e-12 * sm["res0"]
    assert abs(sm["res1"] - g["smoother_residuals"][1]) <= 1e-12 * sm["res1"]
    assert len(by["gpu_exec_v3"]) == 3
    # the other Smoother.hpp classes behind the same interface, against the CPU checkers (pinned to the reference classes
    # in tests/test_oracle_smoothers.py)
    import cpu_checkers as cc
    orc = cc.load("orc")
    n = 65
    f = orc.rhs(n)
    kinds = {"gs_lex": (cc.SMOOTHER_GS_LEX, 1, 10), "rbgs": (cc.SMOOTHER_RBGS, 1, 10), "chebyshev": (cc.SMOOTHER_CHEBYSHEV, 2, 11)}
    for o in by["injected"]:
        sm, nu, coarse = kinds[o["smoother"]]
        want = np.zeros((n, n))
        orc.cycle_s(want, f, kind=cc.V, smoother=sm, alpha=2, nu1=nu, nu2=nu, coarse_sweeps=coarse)
        assert o["phi_mid"] == want[n // 2, n // 2], o["smoother"]
    x = np.zeros((n, n))
    res = orc.gs(x, f, 1.0 / (n - 1), 3)
    g = by["gs_smooth"][0]
    assert g["sweeps"] == 3 and g["x_mid"] == x[n // 2, n // 2] and abs(g["res_last"] - res[-1]) <= 1e-12 * res[-1]
    cres = orc.cg(np.zeros((n, n)), f, 1.0 / (n - 1), 10)
    c = by["cg_smooth"][0]
    # (the manufactured right-hand side is an eigenvector of A: CG is done after one step and the later entries are
    # rounding noise on both sides -- tests/test_gpu_smoothers.py compares whole CG histories on random right-hand sides)
    assert c["entries"] == 11 and abs(c["res0"] - cres[0]) <= 1e-12 * cres[0] and c["res_last"] < 1e-9 * c["res0"]


@pytest.mark.gpu
def test_pmg_runner_writes_reference_output_formats(golden, tmp_path):
    """tools/pmg_runner.cpp: run_all_cycles_err_h / time_h protocols, OUTPUT_RESULT text formats
    (2_part_MG/save_to_file.hpp:59-151) with the reference's own numbers in them."""
    exe = os.path.join(PKG, "pmg_runner")
    if not os.path.exists(exe):
        subprocess.run(["make", "-s", "-C", PKG, "pmg_runner"], check=True)
    out = str(tmp_path / "OUTPUT_RESULT")
    p = subprocess.run([exe, "err_h", "--n", "129,257", "--iters", "1", "--out", out], capture_output=True, text=True,
                       timeout=300)
    assert p.returncode == 0, p.stderr
    want = {r["n"]: r for r in golden["mg_cpu_exec_rel_l2_error"]}
    for tag, key in (("v", "V"), ("w", "W"), ("f", "F")):
        rows = [l.split() for l in open(os.path.join(out, "h_errors_%s_cycle1.txt" % tag)).read().splitlines()]
        assert [int(r[0]) for r in rows] == [129, 257]
        for r in rows:
            assert abs(float(r[1]) - want[int(r[0])][key]) <= 2e-6 * want[int(r[0])][key]  # 6 printed digits
    assert "Final Relative L2 Error: 0.171973" in p.stdout
    p = subprocess.run([exe, "ops", "--n", "257,1025", "--out", out], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    for name in ("residual", "jacobi", "restriction", "prolungator"):
        rows = [l.split() for l in open(os.path.join(out, "timings_%s_gpu.txt" % name)).read().splitlines()]
        assert [r[:2] for r in rows] == [["32", "257"], ["32", "1025"]] and all(float(r[2]) > 0 for r in rows)


@pytest.mark.gpu
def test_pmg_runner_smoother_study_and_error_field_dumps(orc, tmp_path):
    """the Smoother::test hook (Smoother.hpp:50-57,100-115) and the `errors` output through the shims:
    per-sweep residual and relative-error norms against the CPU checker, error FIELDS in the reference's
    save_vector_err_file.hpp format (length on the first line, one component per line, 6 significant digits)."""
    exe = os.path.join(PKG, "pmg_runner")
    n, sweeps = 33, 21
    p = subprocess.run([exe, "smoother", "--smoother", "jacobi", "--n", str(n), "--iters", str(sweeps), "--out", "out"],
                       capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
    assert p.returncode == 0, p.stderr
    f, u = orc.rhs(n), orc.exact(n)
    x = np.zeros((n, n))
    res = orc.jacobi(x, f, 1.0 / (n - 1), omega=1.0, num_iter=sweeps - 1)
    rows = [l.split() for l in open(tmp_path / "out" / ("smoother_jacobi_N%d.txt" % n)).read().splitlines()]
    assert len(rows) == sweeps
    got = np.array([float(r[1]) for r in rows])
    assert np.max(np.abs(got - res) / res) <= 1e-12
    # relative error after the LAST sweep (x now holds it)
    assert abs(float(rows[-1][2]) - orc.norm(x - u) / orc.norm(u)) <= 1e-12
    # error fields: before the first sweep, after sweeps 0, 10, 20
    vec = tmp_path / "OUTPUT_RESULT" / "ERR_VECTOR"
    for i in (0, 10, 20, 30):
        lines = open(vec / ("iteration_%d.txt" % i)).read().split()
        assert int(lines[0]) == n * n and len(lines) == n * n + 1
    first = np.array([float(v) for v in open(vec / "iteration_0.txt").read().split()[1:]]).reshape(n, n)
    assert np.allclose(first, -u, rtol=1e-5, atol=1e-12)  # x = 0: the error is -u, printed with 6 digits
    last = np.array([float(v) for v in open(vec / "iteration_30.txt").read().split()[1:]]).reshape(n, n)
    assert np.allclose(last, x - u, rtol=1e-5, atol=1e-12)
    # the multigrid runner's error-field dump (flag_err_vector_iteration)
    p = subprocess.run([exe, "err_vector", "--n", "65", "--iters", "2", "--omega", "0.6666666666666666", "--eps", "0"],
                       capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
    assert p.returncode == 0, p.stderr
    lines = open(vec / "iteration_last_gpu.txt").read().split()
    assert int(lines[0]) == 65 * 65 and len(lines) == 65 * 65 + 1
    # the other smoothers through the same study
    for name in ("gs", "rbgs", "chebyshev", "cg"):
        p = subprocess.run([exe, "smoother", "--smoother", name, "--n", "33", "--iters", "6", "--out", "out"],
                           capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
        assert p.returncode == 0, (name, p.stderr)
        rows = open(tmp_path / "out" / ("smoother_%s_N33.txt" % name)).read().splitlines()
        assert len(rows) == (7 if name == "cg" else 6)
