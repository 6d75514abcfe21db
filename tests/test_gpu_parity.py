"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the committed goldens.

Tolerances.  Every operator output and every iterate is compared BIT-EXACTLY (np.array_equal): the
kernels evaluate the reference's expressions in its order without FMA contraction.  Residual NORMS:
  * norm_mode = SEQUENTIAL (the reference's left-to-right summation order reproduced on the device):
    compared with `==` -- every per-cycle norm is bit-identical to the CPU path, at every size up to
    the full N = 16385 history.
  * norm_mode = TREE (default, fused parallel sum): <= 1e-10 relative per cycle (north-star tolerance)
    up to N = 4097.  At N = 16385 the REFERENCE's running sum of 2.7e8 squares has itself drifted by
    up to 8.5e-10 from the exactly rounded sum (smooth fields give correlated rounding), so the tree sum
    is held to 2e-9 there; the SEQUENTIAL run shows the iterates are identical.
"""
import numpy as np
import pytest

import cpu_checkers as cc
import pmg_b200 as pmg

pytestmark = pytest.mark.gpu

NORM_RTOL = 1e-10
KIND = {"V": pmg.V, "W": pmg.W, "F": pmg.F}
PROLONG = {"reference": pmg.PROLONG_REFERENCE, "full": pmg.PROLONG_FULL}
ENGINES = [pmg.ENGINE_FUSED, pmg.ENGINE_OPERATOR]


def _rand(shape, seed):
    return np.random.default_rng(seed).standard_normal(shape)


def _hist_close(a, b, rtol=NORM_RTOL):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    rel = np.abs(a - b) / np.abs(b)
    assert rel.max() <= rtol, "max relative deviation %.3e" % rel.max()


# ---------------------------------------------------------------------------------------------------
# operator level (class Parallel replacement), dense reference layout
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(3, 3), (5, 5), (9, 9), (33, 33), (257, 257), (17, 40), (130, 67)])
@pytest.mark.parametrize("omega", [1.0, 2.0 / 3.0])
def test_jacobi_bit_exact(orc, shape, omega):
    h = 1.0 / (shape[0] - 1)
    f = _rand(shape, 1)
    for sweeps in (1, 2, 3, 11):
        x = _rand(shape, 2)
        want = x.copy()
        orc.jacobi(want, f, h, omega=omega, num_iter=sweeps - 1)
        dx, df = pmg.DeviceArray.from_numpy(x), pmg.DeviceArray.from_numpy(f)
        pmg.jacobi(dx, df, h, omega=omega, sweeps=sweeps)
        assert np.array_equal(dx.numpy(), want)


@pytest.mark.parametrize("shape", [(3, 3), (5, 5), (33, 33), (257, 257), (40, 17)])
def test_residual_and_norm(orc, shape):
    h = 1.0 / (shape[1] - 1)
    x, f = _rand(shape, 3), _rand(shape, 4)
    want = orc.residual(x, f, h)
    dx, df = pmg.DeviceArray.from_numpy(x), pmg.DeviceArray.from_numpy(f)
    dr = pmg.DeviceArray.from_numpy(np.zeros(shape))
    n2 = pmg.residual(dr, dx, df, h, want_norm2=True)
    assert np.array_equal(dr.numpy(), want)
    assert abs(np.sqrt(n2) - orc.norm(want)) <= 1e-13 * orc.norm(want)
    assert abs(np.sqrt(pmg.norm2(dr)) - orc.norm(want)) <= 1e-13 * orc.norm(want)


@pytest.mark.parametrize("nf", [5, 9, 33, 129, 257, 1025])
def test_restrict_prolong_bit_exact(orc, nf):
    nc = (nf - 1) // 2 + 1
    fine = _rand((nf, nf), 5)
    dc = pmg.DeviceArray.from_numpy(np.zeros((nc, nc)))
    pmg.restrict_fw(pmg.DeviceArray.from_numpy(fine), dc)
    assert np.array_equal(dc.numpy(), orc.restrict_fw(fine))
    coarse = _rand((nc, nc), 6)  # ring included: the reference reads coarse[idx_c + 1] at ic = nc-2
    for mode in (pmg.PROLONG_REFERENCE, pmg.PROLONG_FULL):
        base = _rand((nf, nf), 7)
        want = orc.prolong_add(base.copy(), coarse, mode)
        dfine = pmg.DeviceArray.from_numpy(base)
        pmg.prolong_add(pmg.DeviceArray.from_numpy(coarse), dfine, mode)
        got = dfine.numpy()
        assert np.array_equal(got, want)
        if mode == pmg.PROLONG_REFERENCE:
            assert np.array_equal(got[1, :], base[1, :]) and np.array_equal(got[:, 1], base[:, 1])


def test_operator_known_answers(golden):
    """SURVEY.md 8c per-operator table (generated from the reference) through the CUDA operators."""
    for g in golden["operators"]:
        n = g["n"]
        h, m = 1.0 / (n - 1), n // 2
        orc = cc.load("orc")
        f = orc.rhs(n)
        dx, df = pmg.DeviceArray.from_numpy(np.zeros((n, n))), pmg.DeviceArray.from_numpy(f)
        pmg.jacobi(dx, df, h, omega=1.0, sweeps=2)
        x = dx.numpy()
        assert (x[m, m], x[1, 1]) == (g["x_mid"], g["x_11"])
        dr = pmg.DeviceArray.from_numpy(np.zeros((n, n)))
        pmg.residual(dr, dx, df, h)
        r = dr.numpy()
        assert (r[m, m], r[1, 1]) == (g["r_mid"], g["r_11"])
        nc = (n - 1) // 2 + 1
        dc = pmg.DeviceArray.from_numpy(np.zeros((nc, nc)))
        pmg.restrict_fw(dr, dc)
        rc = dc.numpy()
        assert (rc[nc // 2, nc // 2], rc[1, 1]) == (g["rc_mid"], g["rc_11"])
        dp = pmg.DeviceArray.from_numpy(np.zeros((n, n)))
        pmg.prolong_add(dc, dp)
        p = dp.numpy()
        assert (p[1, 1], p[1, 2], p[2, 2], p[2, 3], p[3, 3], p[n - 2, n - 2]) == (
            g["p_11"], g["p_12"], g["p_22"], g["p_23"], g["p_33"], g["p_last"])


def test_rhs_sine_bit_exact(orc):
    """pmg_set_rhs_sine == DynamicGridUtils::compute_rhs: check through r = f - A*0."""
    for n in (5, 33, 257):
        with pmg.Solver(n) as s:
            s.set_rhs_sine()
            s.zero_guess()
            f = orc.rhs(n)
            assert abs(s.residual_norm() - orc.norm(orc.residual(np.zeros((n, n)), f, 1.0 / (n - 1)))) \
                <= 1e-13 * orc.norm(f)
            # one damped sweep from zero moves f into x: x = w*0.25*h^2*f exactly representable chain
            s2 = pmg.Solver(n, nu1=1, nu2=1, omega=1.0, engine=pmg.ENGINE_OPERATOR)
            s2.set_rhs_sine()
            s2.zero_guess()
            s2.smooth(1)
            want = np.zeros((n, n))
            orc.jacobi(want, f, 1.0 / (n - 1), omega=1.0, num_iter=0)
            assert np.array_equal(s2.get_solution(), want)
            s2.close()


# ---------------------------------------------------------------------------------------------------
# cycle level: residual histories and iterates against the goldens generated from the reference
# ---------------------------------------------------------------------------------------------------
TREE_RTOL_16385 = 2e-9


def _gpu_history(h, engine, **extra):
    n = h["n"]
    f = cc.load("orc").rhs(n) if h["rhs"] == "sine" else cc.random_rhs(n)
    s = pmg.Solver(n, nu1=h["v1"] + 1, nu2=h["v2"] + 1, omega=h["omega"], gamma=h["alpha"],
                   prolong_mode=PROLONG[h["prolong"]], engine=engine, smoother_eps=h["eps"], **extra)
    s.set_rhs(f)
    s.zero_guess()
    k, hist = s.solve(KIND[h["kind"]], rel_tol=h["rel_tol"], max_cycles=h["max_cycles"])
    phi = s.get_solution()
    s.close()
    return k, hist, phi


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("idx", range(32))
def test_history_matches_reference_golden(golden, engine, idx):
    h = golden["histories"][idx]
    k, hist, phi = _gpu_history(h, engine)
    assert k == h["cycles"], "cycle count differs from the reference"
    _hist_close(hist, h["hist"])
    if h["field"]:
        assert np.array_equal(phi, golden["fields"][h["field"]]), "iterate is not bit-identical"
    # the same run with the reference's summation order: the whole history is bit-identical
    k, hist, _ = _gpu_history(h, engine, norm_mode=pmg.NORM_SEQUENTIAL)
    assert k == h["cycles"] and list(hist) == h["hist"]


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("kind,gamma", [(pmg.V, 1), (pmg.W, 2), (pmg.W, 3), (pmg.F, 1)])
def test_iterates_bit_exact_vs_oracle(orc, engine, kind, gamma):
    """Random start (non-zero Dirichlet ring kept, as the reference keeps it), random RHS."""
    n = 129
    f = cc.random_rhs(n, seed=21) if kind != pmg.F else orc.rhs(n)
    phi0 = _rand((n, n), 22)
    want = phi0.copy()
    s = pmg.Solver(n, omega=2.0 / 3.0, gamma=gamma, engine=engine)
    s.set_rhs(f)
    s.set_guess(phi0)
    for _ in range(3):
        orc.cycle(want, f, kind=kind, omega=2.0 / 3.0, eps=0.0, alpha=gamma)
        rn = s.cycle(kind)
        assert np.array_equal(s.get_solution(), want)
        r = orc.residual(want, f, 1.0 / (n - 1))
        assert abs(rn - orc.norm(r)) <= NORM_RTOL * orc.norm(r)
    s.close()


@pytest.mark.parametrize("nu1,nu2", [(1, 1), (1, 2), (3, 2), (2, 3), (4, 4), (3, 1)])
@pytest.mark.parametrize("prolong", [pmg.PROLONG_REFERENCE, pmg.PROLONG_FULL])
def test_sweep_counts_and_prolong_modes(orc, nu1, nu2, prolong):
    n = 65
    f = cc.random_rhs(n, seed=31)
    want = np.zeros((n, n))
    for _ in range(2):
        orc.cycle(want, f, kind=cc.V, omega=0.8, eps=0.0, alpha=1, v1=nu1 - 1, v2=nu2 - 1, prolong=prolong)
    for engine in ENGINES:
        s = pmg.Solver(n, nu1=nu1, nu2=nu2, omega=0.8, prolong_mode=prolong, engine=engine)
        s.set_rhs(f)
        s.zero_guess()
        s.cycle(pmg.V)
        s.cycle(pmg.V)
        assert np.array_equal(s.get_solution(), want)
        s.close()


@pytest.mark.parametrize("n", [3, 5, 9, 17])
def test_tiny_grids(orc, n):
    f = orc.rhs(n)
    for kind in (pmg.V, pmg.W, pmg.F):
        want = np.zeros((n, n))
        wk, whist = orc.solve(want, f, kind=kind, omega=2.0 / 3.0, eps=0.0, alpha=2, max_cycles=5)
        for engine in ENGINES:
            s = pmg.Solver(n, omega=2.0 / 3.0, gamma=2, engine=engine)
            s.set_rhs(f)
            s.zero_guess()
            k, hist = s.solve(kind, max_cycles=5)
            assert k == wk
            if n > 3:  # n == 3: residual is exactly representable noise around 0
                _hist_close(hist, whist)
            assert np.array_equal(s.get_solution(), want)
            s.close()


def test_as_shipped_configuration_with_smoother_eps(golden):
    """omega = 1, eps = 1e-7 (2_part_MG/main.cpp:12): the per-sweep absolute-norm early exit."""
    for h in golden["histories"]:
        if h["eps"] > 0 and h["n"] <= 257:
            k, hist, _ = _gpu_history(h, pmg.ENGINE_OPERATOR)
            assert k == h["cycles"]
            _hist_close(hist, h["hist"])


def test_mg_cpu_exec_stdout_errors(orc, golden):
    """The only numbers the reference prints: Final Relative L2 Error after 1 cycle (alpha=3, omega=1)."""
    for row in golden["mg_cpu_exec_rel_l2_error"]:
        n = row["n"]
        u = orc.exact(n)
        for kind in "VWF":
            s = pmg.Solver(n, omega=1.0, gamma=3, smoother_eps=1e-7)
            s.set_rhs(orc.rhs(n))
            s.zero_guess()
            s.cycle(KIND[kind], want_norm=False)
            err = orc.norm(s.get_solution() - u) / orc.norm(u)
            assert abs(err - row[kind]) <= 1e-12 * row[kind]
            s.close()


# ---------------------------------------------------------------------------------------------------
# full-size checks (BASELINE configs 2 and 3) through size-independent properties + survey goldens
# ---------------------------------------------------------------------------------------------------
def test_n4097_history_against_survey(golden):
    sv = golden["survey"]
    s = pmg.Solver(4097, omega=2.0 / 3.0)
    s.set_rhs_sine()
    s.zero_guess()
    k, hist = s.solve(pmg.V, rel_tol=1e-8, max_cycles=60)
    s.close()
    assert k == sv["cycles_to_1e-8_V_reference"]["4097"] == 36
    assert abs(hist[0] - np.pi ** 2 * 4096) <= 1e-9 * hist[0]
    _hist_close(hist[1:4], sv["V_n4097_first3"])
    _hist_close(hist[35:37], sv["V_n4097_last2"])


def test_n16385_history_against_survey(golden):
    """BASELINE config 3: all 39 per-cycle norms of the reference CPU path, <= 1e-10 relative."""
    sv = golden["survey"]
    s = pmg.Solver(16385, omega=2.0 / 3.0)
    s.set_rhs_sine()
    s.zero_guess()
    k, hist = s.solve(pmg.V, rel_tol=1e-8, max_cycles=60)
    assert k == 39
    _hist_close(hist[1:], sv["V_n16385"], TREE_RTOL_16385)
    s.close()
    s = pmg.Solver(16385, omega=2.0 / 3.0, prolong_mode=pmg.PROLONG_FULL)
    s.set_rhs_sine()
    s.zero_guess()
    k, hist = s.solve(pmg.V, rel_tol=1e-8, max_cycles=60)
    s.close()
    assert k == 13
    _hist_close(hist[1:], sv["V_n16385_full_prolong"], TREE_RTOL_16385)


def test_n16385_history_bit_identical_with_reference_summation_order(golden):
    """BASELINE config 3 with norm_mode = SEQUENTIAL: all 39 residual norms of the reference CPU run
    (20 min, 13 GB on the host) reproduced BIT FOR BIT -- i.e. every iterate is identical."""
    sv = golden["survey"]
    s = pmg.Solver(16385, omega=2.0 / 3.0, norm_mode=pmg.NORM_SEQUENTIAL)
    s.set_rhs_sine()
    s.zero_guess()
    k, hist = s.solve(pmg.V, rel_tol=1e-8, max_cycles=60)
    s.close()
    assert k == 39
    assert list(hist[1:]) == sv["V_n16385"]


def test_n16385_w_cycle_against_survey(golden):
    sv = golden["survey"]
    s = pmg.Solver(16385, omega=2.0 / 3.0, gamma=2)
    s.set_rhs_sine()
    s.zero_guess()
    k, hist = s.solve(pmg.W, rel_tol=1e-8, max_cycles=40)
    s.close()
    assert k == 19
    _hist_close(hist[1:], sv["W_alpha2_n16385"], TREE_RTOL_16385)


@pytest.mark.parametrize("n", [2049, 4097])
def test_engines_and_variants_agree_bitwise(n):
    """Two independent implementations (fused streaming vs one kernel per operator) and all tuning
    variants of the fused kernels give the same bits on a random RHS."""
    f = cc.random_rhs(n, seed=41)
    out = []
    runs = [(pmg.ENGINE_OPERATOR, 0, 0)] + [(pmg.ENGINE_FUSED, v, v % 2) for v in range(pmg.num_fused_variants())]
    for engine, variant, graph in runs:
        pmg.set_fused_variant(variant)
        s = pmg.Solver(n, omega=2.0 / 3.0, gamma=2, engine=engine, use_graph=graph)
        s.set_rhs(f)
        s.zero_guess()
        norms = [s.cycle(pmg.V), s.cycle(pmg.V), s.cycle(pmg.W)]
        out.append((norms, s.get_solution()))
        s.close()
    pmg.set_fused_variant(-1)
    for norms, phi in out[1:]:
        assert np.array_equal(phi, out[0][1])
        _hist_close(norms, out[0][0])


@pytest.mark.parametrize("n,kind,gamma,omega,prolong", [
    (129, pmg.V, 1, 2.0 / 3.0, pmg.PROLONG_REFERENCE),
    (129, pmg.W, 2, 1.0, pmg.PROLONG_FULL),
    (65, pmg.W, 3, 2.0 / 3.0, pmg.PROLONG_REFERENCE),
    (33, pmg.V, 1, 0.8, pmg.PROLONG_REFERENCE),
    (17, pmg.W, 2, 2.0 / 3.0, pmg.PROLONG_FULL),
    (1025, pmg.W, 2, 2.0 / 3.0, pmg.PROLONG_REFERENCE),
])
def test_small_level_kernel_generations_agree_bitwise(orc, n, kind, gamma, omega, prolong):
    """All three generations of the single-CTA kernel for the levels <= 65 (k_vcycle_small: whole-CTA barriers;
    k_vcycle_small2: per-level thread groups on named barriers; k_coarse_local: level sizes as template parameters,
    one-warp deep levels -- the default) against each other and against the oracle."""
    f = cc.random_rhs(n, seed=61)
    got = {}
    for version in (1, 2, 3):
        pmg.set_small_vcycle_version(version)
        assert pmg.small_vcycle_version() == version
        with pmg.Solver(n, omega=omega, gamma=gamma, prolong_mode=prolong) as s:
            s.set_rhs(f)
            s.zero_guess()
            norms = [s.cycle(kind) for _ in range(3)]
            got[version] = (norms, s.get_solution())
    pmg.set_small_vcycle_version(0)
    assert np.array_equal(got[1][1], got[2][1]) and np.array_equal(got[1][1], got[3][1])
    assert got[1][0] == got[2][0] == got[3][0]
    if n <= 129:
        want = np.zeros((n, n))
        for _ in range(3):
            orc.cycle(want, f, kind=cc.W if kind == pmg.W else cc.V, omega=omega, eps=0.0, alpha=gamma, prolong=prolong)
        assert np.array_equal(got[2][1], want)


@pytest.mark.gpu
@pytest.mark.parametrize("n,kind,gamma,omega,prolong,nu", [
    (257, pmg.V, 1, 2.0 / 3.0, pmg.PROLONG_REFERENCE, (2, 2)),
    (513, pmg.W, 2, 2.0 / 3.0, pmg.PROLONG_REFERENCE, (2, 2)),
    (513, pmg.W, 3, 1.0, pmg.PROLONG_FULL, (1, 2)),
    (1025, pmg.V, 1, 0.8, pmg.PROLONG_FULL, (3, 1)),
    (1025, pmg.F, 1, 2.0 / 3.0, pmg.PROLONG_REFERENCE, (2, 2)),
])
def test_cluster_kernel_matches_streaming_path_and_oracle(orc, n, kind, gamma, omega, prolong, nu):
    """The 16-CTA cluster kernel (kernels_coarse.cu: levels 129 or 257 and below in ONE launch, distributed over the
    CTAs' shared memories) against the path it replaces (streaming passes + the single-CTA kernel) and, at sizes the
    oracle finishes quickly, against the oracle: iterates bit-identical, norms equal."""
    f = cc.random_rhs(n, seed=67)
    phi0 = np.random.default_rng(68).standard_normal((n, n))
    got = {}
    for top in (0, 129, 257):
        if top >= n:
            continue
        pmg.set_cluster_top(top)
        with pmg.Solver(n, omega=omega, gamma=gamma, prolong_mode=prolong, nu1=nu[0], nu2=nu[1]) as s:
            assert s.cluster_top == top, "a B200 can co-schedule a 16-CTA cluster"
            s.set_rhs(f)
            s.set_guess(phi0)
            norms = [s.cycle(kind) for _ in range(2)]
            got[top] = (norms, s.get_solution())
    pmg.set_cluster_top(-1)
    for top in got:
        assert np.array_equal(got[top][1], got[0][1]), top
        assert got[top][0] == got[0][0], top
    if n <= 513:
        want = phi0.copy()
        for _ in range(2):
            orc.cycle(want, f, kind={pmg.V: cc.V, pmg.W: cc.W, pmg.F: cc.F}[kind], omega=omega, eps=0.0, alpha=gamma,
                      v1=nu[0] - 1, v2=nu[1] - 1, prolong=prolong)
        assert np.array_equal(got[129][1], want)


@pytest.mark.gpu
@pytest.mark.parametrize("n,prolong,omega", [(1025, pmg.PROLONG_REFERENCE, 2.0 / 3.0), (2049, pmg.PROLONG_FULL, 1.0)])
def test_f_cycle_folded_passes_match_oracle(orc, n, prolong, omega):
    """Nested iteration on the streaming levels (n >= 513) folds "zero the fine grid, add P phi_coarse" into Pass A of the
    level's V-cycle (k_down's prolong-in form) and -- when the right-hand side is the analytic one (pmg_set_rhs_sine) --
    the residual norm into the last Pass B.  Both forms against the oracle's F-cycle (MultiGridTestRunner.hpp:192-205):
    iterates bit-identical; the folded norm equals the separately computed one to the last bits (another summation tree);
    a right-hand side set with pmg_set_rhs (even the same values) takes the unfolded norm."""
    phi0 = np.random.default_rng(5).standard_normal((n, n))
    phi0[0, :] = phi0[-1, :] = phi0[:, 0] = phi0[:, -1] = 0.0
    f = orc.rhs(n)
    want = phi0.copy()
    for _ in range(2):
        orc.cycle(want, f, kind=cc.F, omega=omega, eps=0.0, alpha=1, prolong=prolong)
    want_norm = orc.norm(orc.residual(want, f, 1.0 / (n - 1)))
    got = {}
    for analytic in (True, False):
        with pmg.Solver(n, omega=omega, prolong_mode=prolong) as s:
            if analytic:
                s.set_rhs_sine()
            else:
                s.set_rhs(f)
            s.set_guess(phi0)
            norms = [s.cycle(pmg.F) for _ in range(2)]
            got[analytic] = (norms, s.get_solution(), s.residual_norm())
    for analytic in (True, False):
        assert np.array_equal(got[analytic][1], want), analytic
        assert abs(got[analytic][0][1] - want_norm) <= 1e-10 * want_norm, analytic  # the oracle sums left to right
        assert abs(got[analytic][2] - want_norm) <= 1e-10 * want_norm, analytic
    assert abs(got[True][0][1] - got[False][0][1]) <= 1e-13 * want_norm, "folded against unfolded norm"
    assert got[False][0][1] == got[False][2], "the unfolded norm IS pmg_residual_norm"


def _pinned(shape):
    import ctypes
    nbytes = int(np.prod(shape)) * 8
    p = ctypes.c_void_p()
    pmg.check(pmg.lib().pmg_host_alloc_pinned(ctypes.byref(p), nbytes))
    buf = (ctypes.c_double * (nbytes // 8)).from_address(p.value)
    return np.frombuffer(buf, dtype=np.float64).reshape(shape), p


def test_overlapped_transfers_match_plain_path():
    """pmg_stage_rhs / pmg_commit_rhs / pmg_fetch_solution_begin / _wait (both PCIe directions on copy streams of their
    own beside the solver) over a stream of three different problems: every solution bit-identical to the plain
    set_rhs / solve / get_solution path, and pmg_set_guess(NULL) is the zero start."""
    n = 1025
    fs = [cc.random_rhs(n, seed=80 + i) for i in range(3)]
    plain = []
    with pmg.Solver(n, omega=2.0 / 3.0) as s:
        for f in fs:
            s.set_rhs(f)
            s.zero_guess()
            k, hist = s.solve(pmg.V, rel_tol=1e-8, max_cycles=60)
            plain.append((k, hist, s.get_solution()))
    f_pin = [_pinned((n, n)) for _ in range(2)]
    out_pin = [_pinned((n, n)) for _ in range(2)]
    got = []
    with pmg.Solver(n, omega=2.0 / 3.0) as s:
        s.set_guess(np.ones((n, n)))  # must be replaced by the zero start below
        f_pin[0][0][:] = fs[0]
        s.stage_rhs(f_pin[0][0])
        for k in range(3):
            s.commit_rhs()
            if k + 1 < 3:
                f_pin[(k + 1) % 2][0][:] = fs[k + 1]
                s.stage_rhs(f_pin[(k + 1) % 2][0])
            s.set_guess(None)
            kk, hist = s.solve(pmg.V, rel_tol=1e-8, max_cycles=60)
            s.fetch_solution_wait()
            if k > 0:
                got[-1] = got[-1] + (out_pin[(k - 1) % 2][0].copy(),)
            s.fetch_solution_begin(out_pin[k % 2][0])
            got.append((kk, hist))
        s.fetch_solution_wait()
        got[-1] = got[-1] + (out_pin[2 % 2][0].copy(),)
        with pytest.raises(pmg.PmgError):
            s.commit_rhs()  # nothing staged
    for (k0, h0, x0), (k1, h1, x1) in zip(plain, got):
        assert k0 == k1 and np.array_equal(h0, h1) and np.array_equal(x0, x1)
    for _, p in f_pin + out_pin:
        pmg.lib().pmg_host_free_pinned(p)


def test_scaling_by_two_is_exact():
    """Linearity in a form floating point honours exactly: f -> 2f doubles every iterate bit for bit."""
    n = 4097
    f = cc.random_rhs(n, seed=51)
    res = []
    for scale in (1.0, 2.0):
        s = pmg.Solver(n, omega=2.0 / 3.0)
        s.set_rhs(f * scale)
        s.zero_guess()
        k, hist = s.solve(pmg.V, rel_tol=1e-6, max_cycles=8)
        res.append((k, hist, s.get_solution()))
        s.close()
    assert res[0][0] == res[1][0]
    assert np.array_equal(res[0][2] * 2.0, res[1][2])
    _hist_close(res[1][1], 2.0 * res[0][1])


def test_manufactured_solution_order_h2(orc):
    """The reference's own validation (DynamicGridUtils.hpp:92-124): discretisation error ~ O(h^2)."""
    errs = []
    for n in (65, 129, 257):
        s = pmg.Solver(n, omega=2.0 / 3.0, prolong_mode=pmg.PROLONG_FULL)
        s.set_rhs_sine()
        s.zero_guess()
        s.solve(pmg.V, rel_tol=1e-11, max_cycles=60)
        u = orc.exact(n)
        errs.append(orc.norm(s.get_solution() - u) / orc.norm(u))
        s.close()
    assert abs(errs[2] - 1.254995e-05) < 1e-9  # SURVEY.md section 6, converged error at N = 257
    assert 3.9 < errs[0] / errs[1] < 4.1 and 3.9 < errs[1] / errs[2] < 4.1


def test_jacobi_pass_blocking_is_bit_exact(orc):
    """pmg_smooth with 1, 2 or 4 sweeps per streaming pass == the same number of plain sweeps."""
    n = 1025
    f, x0 = cc.random_rhs(n, seed=61), _rand((n, n), 62)
    want = x0.copy()
    orc.jacobi(want, f, 1.0 / (n - 1), omega=2.0 / 3.0, num_iter=7)
    for block in (1, 2, 3, 4):
        s = pmg.Solver(n, omega=2.0 / 3.0)
        s.set_rhs(f)
        s.set_guess(x0)
        s.smooth(8, block)
        assert np.array_equal(s.get_solution(), want)
        s.close()


@pytest.mark.parametrize("n,prolong,omega", [(257, pmg.PROLONG_REFERENCE, 2.0 / 3.0), (1025, pmg.PROLONG_FULL, 1.0),
                                             (4097, pmg.PROLONG_REFERENCE, 2.0 / 3.0)])
def test_cross_cycle_solve_matches_two_pass_solve(orc, n, prolong, omega):
    """pmg_solve's cross-cycle path (level 0: Pass B of cycle k and Pass A of cycle k+1 in ONE sweep, k_cross; the
    default for V(2,2) solves) against the classic two passes per cycle: same cycle count, iterate bit-identical, norms
    equal to the last bits (the tree sum runs over another strip geometry); max_cycles cut-offs and a second solve on the
    same handle included; at N = 257 also against the oracle."""
    f = cc.random_rhs(n, seed=91)
    phi0 = np.random.default_rng(92).standard_normal((n, n))
    got = {}
    for cross, shape in ((0, 0), (1, 0), (1, 4)):  # shape 0 = default (strips of 128 columns), 4 = strips of 64
        pmg.set_cross_cycle(cross, shape)
        with pmg.Solver(n, omega=omega, prolong_mode=prolong) as s:
            s.set_rhs(f)
            s.set_guess(phi0)
            k, hist = s.solve(pmg.V, rel_tol=1e-8, max_cycles=60)
            x_full = s.get_solution()
            s.set_guess(phi0)
            k3, hist3 = s.solve(pmg.V, rel_tol=0.0, max_cycles=3)    # even / odd numbers of cross passes
            x3 = s.get_solution()
            k4, hist4 = s.solve(pmg.V, rel_tol=0.0, max_cycles=4)    # continues from x3
            x7 = s.get_solution()
            norm_after = s.cycle(pmg.V)                              # the classic cycle still works on the same handle
            got[(cross, shape)] = (k, hist, x_full, k3, hist3, x3, k4, hist4, x7, norm_after)
    pmg.set_cross_cycle(-1, 0)
    a = got[(0, 0)]
    for key in ((1, 0), (1, 4)):
        b = got[key]
        assert a[0] == b[0] and a[3] == b[3] == 3 and a[6] == b[6] == 4, key
        for i in (2, 5, 8):
            assert np.array_equal(a[i], b[i]), (key, i)
        for i in (1, 4, 7):
            assert np.max(np.abs(a[i] - b[i]) / a[i]) <= 1e-13, (key, i)
        assert abs(a[9] - b[9]) <= 1e-13 * a[9], key
    if n == 257:
        want = phi0.copy()
        for _ in range(3):
            orc.cycle(want, f, kind=cc.V, omega=omega, eps=0.0, alpha=1, prolong=prolong)
        assert np.array_equal(b[5], want)
