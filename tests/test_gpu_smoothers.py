"""GPU parity of the smoothers beyond weighted Jacobi and of the Krylov wrapper, through the C ABI,
against the CPU checkers of oracle/pmg_oracle_smoothers.c -- which tests/test_oracle_smoothers.py pins against the
reference's own GaussSeidelSmoother / ConjugateGradientSmoother / MultigridSolver classes where the reference has them.
Bit-exact (np.array_equal) for every field; 1e-9 relative for the CG scalars (tree sums on the device, left-to-right sums
on the CPU)."""
import numpy as np
import pytest

import cpu_checkers as cc
import pmg_b200 as pmg

pytestmark = pytest.mark.gpu


def _rand(shape, seed):
    return np.random.default_rng(seed).standard_normal(shape)


@pytest.mark.parametrize("n,sweeps", [(3, 2), (5, 1), (9, 3), (33, 2), (129, 1), (65, 4)])
def test_gauss_seidel_orderings_bit_exact(orc, n, sweeps):
    f, x0 = cc.random_rhs(n, seed=21), _rand((n, n), 22)
    h = 1.0 / (n - 1)
    for ordering, fn in ((0, lambda x: orc.gs(x, f, h, sweeps)), (1, lambda x: orc.rbgs(x, f, h, sweeps))):
        want = x0.copy()
        fn(want)
        dx, df = pmg.DeviceArray.from_numpy(x0), pmg.DeviceArray.from_numpy(f)
        pmg.gauss_seidel(dx, df, h, sweeps, ordering)
        assert np.array_equal(dx.numpy(), want), (n, ordering)


def test_lexicographic_gauss_seidel_equals_the_reference_class(ref):
    n = 33
    f, x0 = cc.random_rhs(n, seed=23), _rand((n, n), 24)
    want = x0.copy()
    ref.gs(want, f, 1.0 / (n - 1), 3)
    dx, df = pmg.DeviceArray.from_numpy(x0), pmg.DeviceArray.from_numpy(f)
    pmg.gauss_seidel(dx, df, 1.0 / (n - 1), 3, 0)
    assert np.array_equal(dx.numpy(), want)


SMOOTHERS = {"rbgs": (pmg.SMOOTHER_RBGS, cc.SMOOTHER_RBGS), "gs_lex": (pmg.SMOOTHER_GS_LEX, cc.SMOOTHER_GS_LEX),
             "chebyshev": (pmg.SMOOTHER_CHEBYSHEV, cc.SMOOTHER_CHEBYSHEV)}


@pytest.mark.parametrize("name", sorted(SMOOTHERS))
@pytest.mark.parametrize("n,kind,gamma,nu,prolong", [
    (17, "V", 1, (1, 1), pmg.PROLONG_REFERENCE),
    (65, "W", 2, (2, 1), pmg.PROLONG_FULL),
    (257, "V", 1, (2, 2), pmg.PROLONG_FULL),
])
def test_cycles_with_injected_smoothers_bit_exact(orc, name, n, kind, gamma, nu, prolong):
    sm_gpu, sm_cpu = SMOOTHERS[name]
    f, phi0 = cc.random_rhs(n, seed=25), _rand((n, n), 26)
    want = phi0.copy()
    for _ in range(2):
        orc.cycle_s(want, f, kind=cc.W if kind == "W" else cc.V, smoother=sm_cpu, alpha=gamma, nu1=nu[0], nu2=nu[1],
                    coarse_sweeps=10, prolong=prolong)
    with pmg.Solver(n, smoother=sm_gpu, gamma=gamma, nu1=nu[0], nu2=nu[1], coarse_sweeps=10, prolong_mode=prolong) as s:
        s.set_rhs(f)
        s.set_guess(phi0)
        for _ in range(2):
            norm = s.cycle(pmg.W if kind == "W" else pmg.V)
        got = s.get_solution()
    assert np.array_equal(got, want)
    r = orc.residual(want, f, 1.0 / (n - 1))
    assert abs(norm - orc.norm(r)) <= 1e-10 * orc.norm(r)


def test_gauss_seidel_multigrid_equals_the_reference_with_injected_smoother(ref):
    """MultigridSolver(&GaussSeidelSmoother, alpha, N) of the reference itself (oracle/_ref), V(1,1) and W."""
    n = 65
    f, phi0 = cc.random_rhs(n, seed=27), _rand((n, n), 28)
    for kind, ckind, gamma in ((pmg.V, cc.V, 1), (pmg.W, cc.W, 3)):
        want = phi0.copy()
        ref.cycle_s(want, f, kind=ckind, smoother=cc.SMOOTHER_GS_LEX, alpha=gamma, nu1=1, nu2=1, coarse_sweeps=10)
        with pmg.Solver(n, smoother=pmg.SMOOTHER_GS_LEX, gamma=gamma, nu1=1, nu2=1, coarse_sweeps=10) as s:
            s.set_rhs(f)
            s.set_guess(phi0)
            s.cycle(kind, want_norm=False)
            assert np.array_equal(s.get_solution(), want)


def test_red_black_multigrid_converges_faster_than_jacobi():
    """What the smoother is for: V(1,1) with red-black Gauss-Seidel and full-interior prolongation beats V(2,2) weighted
    Jacobi per cycle."""
    n = 1025
    f = cc.random_rhs(n, seed=29)
    res = {}
    for name, cfg in (("jacobi", dict(omega=2.0 / 3.0)), ("rbgs", dict(smoother=pmg.SMOOTHER_RBGS, nu1=1, nu2=1))):
        with pmg.Solver(n, prolong_mode=pmg.PROLONG_FULL, **cfg) as s:
            s.set_rhs(f)
            s.zero_guess()
            res[name] = s.solve(pmg.V, rel_tol=1e-8, max_cycles=60)
    assert res["rbgs"][1][-1] < 1e-8 * res["rbgs"][1][0]
    assert res["rbgs"][0] <= res["jacobi"][0]


@pytest.mark.parametrize("n,precond,engine", [(33, 0, pmg.ENGINE_FUSED), (65, 1, pmg.ENGINE_FUSED), (65, 1, pmg.ENGINE_OPERATOR),
                                              (257, 1, pmg.ENGINE_FUSED)])
def test_pcg_matches_the_cpu_specification(orc, n, precond, engine):
    f = cc.random_rhs(n, seed=31)
    phi0 = np.zeros((n, n))
    phi0[0, :] = 1.0  # a Dirichlet ring that is not zero
    want = phi0.copy()
    kw, hw = orc.pcg(want, f, precond=precond, rel_tol=1e-9, max_iter=60 if precond else 400)
    with pmg.Solver(n, omega=2.0 / 3.0, prolong_mode=pmg.PROLONG_FULL, engine=engine) as s:
        s.set_rhs(f)
        s.set_guess(phi0)
        k, hist = s.pcg(precond=precond, rel_tol=1e-9, max_iter=60 if precond else 400)
        got = s.get_solution()
        # the solver is still usable for cycles afterwards (its level-0 arrays were only borrowed)
        s.cycle(pmg.V)
    assert abs(k - kw) <= (0 if precond else 2)
    m = min(len(hist), len(hw), 12)
    assert np.max(np.abs(hist[:m] - hw[:m]) / hw[:m]) <= 1e-9
    assert hist[-1] < 1e-9 * hist[0]
    assert np.max(np.abs(got - want)) <= 1e-9 * np.max(np.abs(want))
    if precond:
        assert k <= 14


def test_pcg_with_vcycle_preconditioner_at_scale():
    """N = 4097: one V(2,2) cycle per step as preconditioner; 1e-8 in about ten steps."""
    n = 4097
    with pmg.Solver(n, omega=2.0 / 3.0, prolong_mode=pmg.PROLONG_FULL) as s:
        s.set_rhs_sine()
        s.zero_guess()
        k, hist = s.pcg(precond=1, rel_tol=1e-8, max_iter=40)
    assert hist[-1] < 1e-8 * hist[0] and k <= 12


def test_fp32_smoother_experiment_shows_the_precision_floor():
    """SURVEY.md 8f-4, mixed precision as an EXPERIMENT (pmg_config.smoother_fp32): Jacobi sweeps evaluated in fp32 on the
    fp64 fields, residual and grid transfers in fp64.  The early cycles follow the fp64 history closely; the residual then
    stalls at the fp32 rounding floor instead of reaching 1e-8 -- which is why the product path smooths in fp64."""
    n = 1025
    hist = {}
    for fp32 in (0, 1):
        with pmg.Solver(n, omega=2.0 / 3.0, prolong_mode=pmg.PROLONG_FULL, engine=pmg.ENGINE_OPERATOR, smoother_fp32=fp32) as s:
            s.set_rhs_sine()
            s.zero_guess()
            hist[fp32] = s.solve(pmg.V, rel_tol=1e-10, max_cycles=25)[1]
    h64, h32 = hist[0], hist[1]
    assert h64[-1] < 1e-10 * h64[0]
    print("fp64 history", h64[:6], "... fp32-smoother history", h32[:6], "... final", h32[-1] / h32[0])
    assert abs(h32[1] - h64[1]) / h64[1] < 1e-2                        # the first cycle contracts alike
    assert len(h32) == 26 and h32[-1] / h32[0] > 1e-9                  # ... then the fp32 floor: 1e-10 is never reached
    assert h32[-1] > 1e3 * h64[-1]
