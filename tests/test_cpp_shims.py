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
