"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/pmg.h
declares, validates arguments, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import pmg_b200 as pmg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pmg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pmg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    L = pmg.lib()
    declared = _declared_symbols()
    assert len(declared) >= 30
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, "declared in include/pmg.h but not exported by libpmg.so: %s" % missing
    assert sorted(pmg.ABI_SYMBOLS) == declared, "python binding list out of sync with include/pmg.h"


def test_version_and_status_strings():
    L = pmg.lib()
    assert b"sm_100a" in L.pmg_version()
    assert L.pmg_status_string(0) == b"ok"
    assert b"no CPU fallback" in L.pmg_status_string(3)


def test_config_defaults_match_reference():
    cfg = pmg.default_config(257)
    # MultiGrid.hpp:15-16 v1=v2=1 -> 2 sweeps; :19 N_coarse=5; :61 10+1 sweeps; :153 3+1; main.cpp:15 alpha=3
    assert (cfg.n, cfg.nu1, cfg.nu2, cfg.n_coarse, cfg.coarse_sweeps, cfg.fmg_sweeps, cfg.gamma) == \
        (257, 2, 2, 5, 11, 4, 3)
    assert cfg.omega == 1.0 and cfg.prolong_mode == pmg.PROLONG_REFERENCE and cfg.smoother_eps == 0.0


@pytest.mark.parametrize("n", [0, 2, 4, 6, 100, 256, 258, -5, 70000])
def test_create_rejects_bad_sizes(n):
    with pytest.raises(pmg.PmgError) as e:
        pmg.Solver(n)
    assert e.value.status == 1  # PMG_ERR_INVALID, before any device is touched


def test_create_rejects_bad_parameters():
    for kw in ({"nu1": -1}, {"gamma": 0}, {"omega": 0.0}, {"n_coarse": 6}, {"coarse_sweeps": -2}):
        with pytest.raises(pmg.PmgError) as e:
            pmg.Solver(33, **kw)
        assert e.value.status == 1


def test_no_device_means_error_not_fallback():
    if pmg.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(pmg.PmgError) as e:
        pmg.Solver(33)
    assert e.value.status == 3  # PMG_ERR_NO_DEVICE
    p = ctypes.c_void_p()
    assert pmg.lib().pmg_device_alloc(ctypes.byref(p), 64) == 3
    one = ctypes.c_double()
    assert pmg.lib().pmg_norm2(ctypes.c_void_p(8), 1, ctypes.byref(one), None) == 3


@pytest.mark.parametrize("n", [9, 17, 257, 4097, 16385, 32769])
@pytest.mark.parametrize("ranks", [1, 2, 3, 4, 8])
def test_partition_rows(n, ranks):
    """Row slabs tile [0, n), start on even rows (coarse row jc lives with fine row 2jc) and the coarse
    partition induced by halving is exactly the partition of the coarse level's own rows."""
    prev = 0
    for r in range(ranks):
        y0, y1 = pmg.partition_rows(n, ranks, r)
        assert y0 == prev and y1 >= y0 and y0 % 2 == 0
        prev = y1
    assert prev == n
    if (n - 1) // 2 + 1 >= 3:
        nc = (n - 1) // 2 + 1
        for r in range(ranks):
            y0, y1 = pmg.partition_rows(n, ranks, r)
            c0, c1 = y0 // 2, (nc if r == ranks - 1 else y1 // 2)
            assert 2 * c0 == y0 and (r == ranks - 1 or 2 * c1 == y1)


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` runs on the host cores (no GPU needed) and prints one JSON line with the
    contract's keys; it is the only place outside tests/ that executes oracle/ (as the CPU baseline)."""
    import json
    import subprocess
    import sys
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    line = json.loads(p.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["dtype"] == "f64" and line["unit"] == "GDOF/s"
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0
    assert "workload" in line["config"] and "model" not in line["config"]


def test_product_never_touches_the_oracle():
    """The product path (package, csrc, include, bench_dist) must not import, link or name oracle/."""
    pkg = os.path.dirname(pmg.LIB_PATH)
    offenders = []
    for base, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".cu", ".h", ".hpp", ".py", ".cpp")) or fn == "Makefile":
                text = open(os.path.join(base, fn), errors="ignore").read()
                if "oracle" in text.replace("the oracle", "").lower() and ("liboracle" in text or "cpu_checkers" in text
                                                                          or "oracle/" in text):
                    offenders.append(os.path.join(base, fn))
    for fn in ("include/pmg.h", "include/pmg.hpp", "pmg_b200.py"):
        text = open(os.path.join(ROOT, fn)).read()
        if "liboracle" in text or "cpu_checkers" in text or "oracle/" in text:
            offenders.append(fn)
    assert not offenders, offenders
    out = subprocess_check_ldd(pmg.LIB_PATH)
    assert "oracle" not in out and "pmg_ref" not in out


def subprocess_check_ldd(path):
    import subprocess
    return subprocess.run(["ldd", path], capture_output=True, text=True).stdout
