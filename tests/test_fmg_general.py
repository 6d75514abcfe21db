"""General-RHS full multigrid (PMG_CYCLE_FMG) -- NOT a reference function: the reference's F-cycle
regenerates the analytic right-hand side on every level and zeroes the ring.  oracle/pmg_oracle.c::orc_fmg_general is
its specification, assembled from the reference's operators.

CPU part: the specification behaves like full multigrid (ring kept, discretisation-level accuracy after ONE pass on a
smooth problem, V-cycles afterwards converge at the V-cycle rate).
The GPU half (libpmg reproduces it bit for bit) is tests/test_zz_gpu_fmg_general.py.
"""
import numpy as np
import pytest

import cpu_checkers as cc


def test_specification_is_a_full_multigrid_pass(orc):
    n = 257
    h = 1.0 / (n - 1)
    f = orc.rhs(n)
    exact = orc.exact(n)
    phi = np.zeros((n, n))
    orc.fmg_general(phi, f, prolong=cc.PROLONG_FULL)
    err_fmg = np.linalg.norm(phi - exact) / np.linalg.norm(exact)
    v = np.zeros((n, n))
    orc.cycle(v, f, kind=cc.V, prolong=cc.PROLONG_FULL)
    err_v = np.linalg.norm(v - exact) / np.linalg.norm(exact)
    assert err_fmg < 2e-3 and err_fmg < 0.05 * err_v  # one FMG pass ~ discretisation error; one V-cycle is far off
    # non-homogeneous Dirichlet data: u = x + y is discretely harmonic, f = 0
    x = np.arange(n) * h
    u = x[None, :] + x[:, None]
    phi = u.copy()
    phi[1:-1, 1:-1] = 7.0  # garbage in the interior: FMG restarts it
    zero = np.zeros((n, n))
    orc.fmg_general(phi, zero, prolong=cc.PROLONG_FULL)
    assert np.array_equal(phi[0], u[0]) and np.array_equal(phi[:, -1], u[:, -1])
    r0 = np.zeros((n, n))
    r0[[0, -1], :] = u[[0, -1], :]
    r0[:, [0, -1]] = u[:, [0, -1]]
    e0 = np.abs(r0 - u).max()
    # the ring enters through r0 = f - A x_b (lifting), a boundary layer the full-weighting hierarchy only partly sees:
    # one pass removes ~90 % of the error here, the V-cycles after it the rest
    assert np.abs(phi - u).max() < 0.15 * e0
    for _ in range(8):
        orc.cycle(phi, zero, kind=cc.V, prolong=cc.PROLONG_FULL)
    assert np.abs(phi - u).max() < 1e-6
