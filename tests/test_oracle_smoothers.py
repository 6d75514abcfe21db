"""The CPU checkers for the smoothers beyond weighted Jacobi and for the Krylov wrapper (SURVEY.md 8f-3,
oracle/pmg_oracle_smoothers.c): pinned against the REAL reference classes where the reference has them
(GaussSeidelSmoother, ConjugateGradientSmoother, MultigridSolver with a GaussSeidelSmoother injected -- all through
oracle/ref_driver.cpp over the unmodified headers), and against hand-computed / structural facts where it has not
(red-black ordering, Chebyshev weights, preconditioned CG)."""
import numpy as np
import pytest

import cpu_checkers as cc


def _rand(shape, seed):
    return np.random.default_rng(seed).standard_normal(shape)


@pytest.mark.parametrize("n,iters", [(5, 1), (9, 3), (33, 2), (65, 1)])
def test_gs_equals_reference_gauss_seidel_smoother(orc, ref, n, iters):
    f, x0 = cc.random_rhs(n, seed=3), _rand((n, n), 4)
    a, b = x0.copy(), x0.copy()
    ra = orc.gs(a, f, 1.0 / (n - 1), iters)
    rb = ref.gs(b, f, 1.0 / (n - 1), iters)
    assert np.array_equal(a, b) and np.array_equal(ra, rb) and len(ra) == iters


def test_gs_eps_exit_matches_reference(orc, ref):
    n = 9
    f, x0 = cc.random_rhs(n, seed=5), np.zeros((n, n))
    a, b = x0.copy(), x0.copy()
    ra = orc.gs(a, f, 1.0 / (n - 1), 200, eps=1e-3)
    rb = ref.gs(b, f, 1.0 / (n - 1), 200, eps=1e-3)
    assert np.array_equal(a, b) and np.array_equal(ra, rb) and 1 < len(ra) < 200 and ra[-1] < 1e-3


@pytest.mark.parametrize("n,kind,alpha,nu", [(17, cc.V, 1, (1, 1)), (33, cc.V, 1, (2, 1)), (33, cc.W, 2, (1, 1)), (65, cc.W, 3, (1, 2))])
def test_gs_multigrid_cycle_equals_reference_with_injected_gauss_seidel(orc, ref, n, kind, alpha, nu):
    """MultigridSolver(&GaussSeidelSmoother, alpha, N): v1 / v2 ARE the sweep counts for this smoother (its loop is `<`) and
    the coarsest solve is 10 sweeps (MultiGrid.hpp:61)."""
    f, phi0 = cc.random_rhs(n, seed=6), _rand((n, n), 7)
    a, b = phi0.copy(), phi0.copy()
    for _ in range(2):
        orc.cycle_s(a, f, kind=kind, smoother=cc.SMOOTHER_GS_LEX, alpha=alpha, nu1=nu[0], nu2=nu[1], coarse_sweeps=10)
        ref.cycle_s(b, f, kind=kind, smoother=cc.SMOOTHER_GS_LEX, alpha=alpha, nu1=nu[0], nu2=nu[1], coarse_sweeps=10)
    assert np.array_equal(a, b)


def test_red_black_hand_computed_case(orc):
    """5 x 5 grid, h = 1/4, x = 0, f = 16 everywhere on the interior (h^2 f = 1).  Red = (col + row) even.
    Red half sweep: every red point becomes 0.25 * (0 + 0 + 0 + 0 + 1) = 0.25.
    Black half sweep: a black point with k red INTERIOR neighbours becomes 0.25 * (0.25 k + 1):
    (1,2),(2,1),(2,3),(3,2) have 3 red interior neighbours -> 0.4375."""
    n = 5
    x = np.zeros((n, n))
    f = np.zeros((n, n))
    f[1:-1, 1:-1] = 16.0
    orc.rbgs(x, f, 0.25, 1)
    want = np.zeros((n, n))
    for r in range(1, 4):
        for c in range(1, 4):
            want[r, c] = 0.25 if (r + c) % 2 == 0 else 0.4375
    assert np.array_equal(x, want)


def test_red_half_sweep_is_jacobi_on_the_red_points_and_one_row_grids_are_gauss_seidel(orc):
    n = 17
    f, x0 = cc.random_rhs(n, seed=8), _rand((n, n), 9)
    h = 1.0 / (n - 1)
    # after one red-black sweep the RED points hold what the Gauss-Seidel expression gives from the OLD field
    x = x0.copy()
    orc.rbgs(x, f, h, 1)
    for r in range(1, n - 1):
        for c in range(1, n - 1):
            if (r + c) % 2 == 0:
                want = 0.25 * (x0[r, c - 1] + x0[r, c + 1] + x0[r - 1, c] + x0[r + 1, c] + h * h * f[r, c])
                assert x[r, c] == want
    # and the black points use the NEW red values
    r, c = 3, 4
    want = 0.25 * (x[r, c - 1] + x[r, c + 1] + x[r - 1, c] + x[r + 1, c] + h * h * f[r, c])
    assert x[r, c] == want
    # a grid with ONE interior point: every ordering is the same
    f3, x3 = cc.random_rhs(3, seed=10), _rand((3, 3), 11)
    a, b = x3.copy(), x3.copy()
    orc.rbgs(a, f3, 0.5, 3)
    orc.gs(b, f3, 0.5, 3)
    assert np.array_equal(a, b)


def test_chebyshev_is_a_sequence_of_reference_pinned_weighted_sweeps(orc, ref):
    n = 33
    f, x0 = cc.random_rhs(n, seed=12), _rand((n, n), 13)
    h = 1.0 / (n - 1)
    for nu in (1, 2, 3):
        w = orc.chebyshev_weights(nu)
        d, c = 1.25, 0.75
        assert np.allclose(w, [1.0 / (d - c * np.cos(np.pi * (2 * k + 1) / (2 * nu))) for k in range(nu)], rtol=1e-15)
        a, b = x0.copy(), x0.copy()
        orc.jacobi_weights(a, f, h, w)
        for wk in w:
            ref.jacobi(b, f, h, omega=float(wk), num_iter=0)
        assert np.array_equal(a, b)
        # and the cycle-level dispatcher uses exactly these weights
        p, q = x0.copy(), x0.copy()
        orc.cycle_s(p, f, smoother=cc.SMOOTHER_CHEBYSHEV, nu1=nu, nu2=nu, coarse_sweeps=nu)
    # min-max property on [1/2, 2]: at the LOW edge of the smoothing range (eigenvalue of D^-1 A ~ 0.53) two Chebyshev
    # sweeps damp by 0.19 where two sweeps of omega = 2/3 only reach 0.42 (at the top edge they give 0.22 vs 0.11:
    # the bound 1 / T_2(5/3) = 0.22 holds over the whole range, 2/3 reaches 0.44 somewhere in it)
    e = np.zeros((n, n))
    kx, ky = 11, 11
    e[1:-1, 1:-1] = np.outer(np.sin(np.pi * ky * np.arange(1, n - 1) / (n - 1)), np.sin(np.pi * kx * np.arange(1, n - 1) / (n - 1)))
    z = np.zeros((n, n))
    a, b = e.copy(), e.copy()
    orc.jacobi_weights(a, z, h, orc.chebyshev_weights(2))
    orc.jacobi(b, z, h, omega=2.0 / 3.0, num_iter=1)
    assert np.abs(a).max() < np.abs(b).max()


@pytest.mark.parametrize("n,iters", [(9, 5), (33, 12)])
def test_cg_equals_reference_conjugate_gradient_smoother(orc, ref, n, iters):
    f = cc.random_rhs(n, seed=14)
    a, b = _rand((n, n), 15), _rand((n, n), 15)  # the reference zeroes x first (Smoother.hpp:186)
    ra = orc.cg(a, f, 1.0 / (n - 1), iters)
    rb = ref.cg(b, f, 1.0 / (n - 1), iters)
    assert np.array_equal(a, b) and np.array_equal(ra, rb) and len(ra) == iters + 1


def test_pcg_without_preconditioner_follows_the_pinned_cg_and_multigrid_preconditioning_pays(orc):
    n = 33
    f = orc.rhs(n)
    h = 1.0 / (n - 1)
    x = np.zeros((n, n))
    k, hist = orc.pcg(x, f, precond=0, rel_tol=0.0, max_iter=15)
    ref_hist = orc.cg(np.zeros((n, n)), f, h, 15)
    assert k == 15 and np.allclose(hist, ref_hist[:16], rtol=1e-9)
    # one V(2,2) cycle (full-interior prolongation, omega = 2/3) as preconditioner: an order of magnitude fewer steps
    n = 65
    f = cc.random_rhs(n, seed=16)
    x0, x1 = np.zeros((n, n)), np.zeros((n, n))
    k0, h0 = orc.pcg(x0, f, precond=0, rel_tol=1e-8, max_iter=500)
    k1, h1 = orc.pcg(x1, f, precond=1, rel_tol=1e-8, max_iter=500)
    assert h1[-1] < 1e-8 * h1[0] and h0[-1] < 1e-8 * h0[0]
    assert k1 <= 12 and k0 >= 8 * k1
    assert np.allclose(x0, x1, atol=1e-7 * np.abs(x0).max())
