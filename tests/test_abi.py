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
