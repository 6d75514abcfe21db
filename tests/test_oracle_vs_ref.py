"""oracle (restatement) == reference (unmodified headers), bit for bit, on seeded random inputs and
edge sizes.  Skipped where oracle/_ref/libpmg_ref.so is absent.  CPU only."""
import numpy as np
import pytest

import cpu_checkers as cc


def _rand(n, seed, ring=False):
    a = np.random.default_rng(seed).standard_normal((n, n))
    if not ring:
        a[0, :] = a[-1, :] = 0.0
        a[:, 0] = a[:, -1] = 0.0
    return a


@pytest.mark.parametrize("n", [3, 5, 9, 17, 33, 65, 129])
@pytest.mark.parametrize("omega", [1.0, 2.0 / 3.0, 0.8])
def test_jacobi(orc, ref, n, omega):
    h = 1.0 / (n - 1)
    f = _rand(n, 1, ring=True)
    for num_iter in (0, 1, 3, 10):
        xo, xr = _rand(n, 2, ring=True), _rand(n, 2, ring=True)
        ro = orc.jacobi(xo, f, h, omega=omega, num_iter=num_iter)
        rr = ref.jacobi(xr, f, h, omega=omega, num_iter=num_iter)
        assert len(ro) == num_iter + 1 and np.array_equal(ro, rr)
        assert np.array_equal(xo, xr)
    if omega == 1.0:
        # the shipped JacobiSmoother object itself: only x is comparable (its per-sweep norm reads
        # the uninitialised ring of a fresh new[] -- Smoother.hpp:75-77)
        ref.lib.ref_use_shipped_jacobi(1)
        xo, xr = _rand(n, 2, ring=True), _rand(n, 2, ring=True)
        orc.jacobi(xo, f, h, omega=1.0, num_iter=3)
        ref.jacobi(xr, f, h, omega=1.0, num_iter=3)
        ref.lib.ref_use_shipped_jacobi(0)
        assert np.array_equal(xo, xr)


def test_jacobi_early_exit(orc, ref):
    """Smoother.hpp:84-88: absolute-eps break; same sweep count on both sides."""
    n, h = 9, 1.0 / 8
    f = np.zeros((n, n))
    xo, xr = _rand(n, 3) * 1e-9, _rand(n, 3) * 1e-9
    ro = orc.jacobi(xo, f, h, omega=1.0, num_iter=50, eps=1e-7)
    rr = ref.jacobi(xr, f, h, omega=1.0, num_iter=50, eps=1e-7)
    assert 0 < len(ro) < 51 and np.array_equal(ro, rr) and np.array_equal(xo, xr)


@pytest.mark.parametrize("n", [3, 5, 9, 33, 129, 257])
def test_residual_restrict_prolong_norm(orc, ref, n):
    h = 1.0 / (n - 1)
    x, f = _rand(n, 4, ring=True), _rand(n, 5, ring=True)
    r_o, r_r = orc.residual(x, f, h), ref.residual(x, f, h)
    assert np.array_equal(r_o, r_r)
    assert orc.norm(r_o) == ref.norm(r_r)
    if n >= 5:
        assert np.array_equal(orc.restrict_fw(x), ref.restrict_fw(x))
        nc = (n - 1) // 2 + 1
        c = _rand(nc, 6, ring=True)  # ring of the coarse field IS read (coarse[idx_c+1] at ic=nc-2)
        fo, fr = _rand(n, 7, ring=True), _rand(n, 7, ring=True)
        orc.prolong_add(fo, c)
        ref.prolong_add(fr, c)
        assert np.array_equal(fo, fr)


def test_full_prolongation_is_reference_plus_row_col_1(orc):
    n, nc = 17, 9
    c = _rand(nc, 8)
    a, b = np.zeros((n, n)), np.zeros((n, n))
    orc.prolong_add(a, c, cc.PROLONG_REFERENCE)
    orc.prolong_add(b, c, cc.PROLONG_FULL)
    assert np.array_equal(a[2:-1, 2:-1], b[2:-1, 2:-1])
    assert not a[1, :].any() and not a[:, 1].any()
    assert b[1, 2:-1].any() and b[2:-1, 1].any()
    assert not b[0, :].any() and not b[-1, :].any() and not b[:, 0].any() and not b[:, -1].any()
    # standard bilinear weights on the extra row: odd/odd point = mean of the 4 surrounding coarse points
    assert b[1, 1] == 0.25 * (c[0, 0] + c[0, 1] + c[1, 0] + c[1, 1])
    assert b[1, 2] == 0.5 * (c[0, 1] + c[1, 1])


@pytest.mark.parametrize("kind,alpha", [(cc.V, 1), (cc.W, 2), (cc.W, 3), (cc.F, 1)])
@pytest.mark.parametrize("omega,eps", [(2.0 / 3.0, 0.0), (1.0, 1e-7)])
def test_cycles_random_state(orc, ref, kind, alpha, omega, eps):
    n = 65
    f = cc.random_rhs(n, seed=11) if kind != cc.F else orc.rhs(n)
    po, pr = _rand(n, 12), _rand(n, 12)
    for _ in range(2):
        orc.cycle(po, f, kind=kind, omega=omega, eps=eps, alpha=alpha, v1=1, v2=1)
        ref.cycle(pr, f, kind=kind, omega=omega, eps=eps, alpha=alpha, v1=1, v2=1)
        assert np.array_equal(po, pr)


def test_shipped_jacobi_object_equals_weighted_at_omega_1(ref):
    """The injected weighted smoother at omega = 1 IS JacobiSmoother.  eps = 0: the
    shipped object's early exit reads an uninitialised ring (Smoother.hpp:75-77), so with eps > 0
    the reference itself is not deterministic."""
    n = 65
    f = ref.rhs(n)
    out = []
    for shipped in (0, 1):
        ref.lib.ref_use_shipped_jacobi(shipped)
        phi = np.zeros((n, n))
        out.append((ref.solve(phi, f, kind=cc.V, omega=1.0, eps=0.0, alpha=3, max_cycles=15), phi))
    ref.lib.ref_use_shipped_jacobi(0)
    assert out[0][0][0] == out[1][0][0] and np.array_equal(out[0][0][1], out[1][0][1])
    assert np.array_equal(out[0][1], out[1][1])


def test_rhs_and_exact(orc, ref):
    for n in (5, 33, 257):
        assert np.array_equal(orc.rhs(n), ref.rhs(n))
        assert np.array_equal(orc.exact(n), ref.exact(n))
