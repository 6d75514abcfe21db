"""ctypes bindings for the two CPU checkers under oracle/ (test infrastructure only).

`load("orc")` -> oracle/liboracle.so (plain-C restatement), `load("ref")` ->
oracle/_ref/libpmg_ref.so (the unmodified reference headers behind oracle/ref_driver.cpp).
Both export the ABI of oracle/oracle.h with the prefix `orc_` / `ref_`.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORC_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libpmg_ref.so")

V, W, F = 0, 1, 2
PROLONG_REFERENCE, PROLONG_FULL = 0, 1
SMOOTHER_JACOBI, SMOOTHER_RBGS, SMOOTHER_GS_LEX, SMOOTHER_CHEBYSHEV = 0, 1, 2, 3

_dp = ctypes.POINTER(ctypes.c_double)
_i, _d, _l = ctypes.c_int, ctypes.c_double, ctypes.c_long


def _p(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


class Checker:
    """Thin numpy front-end over one of the two libraries."""

    def __init__(self, lib, prefix):
        self.lib, self.prefix = lib, prefix
        g = lambda n: getattr(lib, prefix + n)
        g("jacobi").restype = _i
        g("jacobi").argtypes = [_dp, _dp, _i, _i, _d, _d, _i, _d, _dp]
        g("residual").restype = None
        g("residual").argtypes = [_dp, _dp, _dp, _i, _i, _d]
        g("norm").restype = _d
        g("norm").argtypes = [_dp, _l]
        g("restrict_fw").restype = None
        g("restrict_fw").argtypes = [_dp, _dp, _i, _i]
        g("prolong_add").restype = _i
        g("prolong_add").argtypes = [_dp, _dp, _i, _i, _i]
        g("rhs").restype = None
        g("rhs").argtypes = [_dp, _i, _i, _d]
        g("exact").restype = None
        g("exact").argtypes = [_dp, _d, _i, _i]
        g("cycle").restype = _i
        g("cycle").argtypes = [_dp, _dp, _i, _d, _i, _d, _d, _i, _i, _i, _i]
        g("solve").restype = _i
        g("solve").argtypes = [_dp, _dp, _i, _i, _d, _d, _i, _i, _i, _i, _d, _i, _dp]
        g("fmg_general").restype = _i
        g("fmg_general").argtypes = [_dp, _dp, _i, _d, _d, _i, _i, _i]
        # smoothers beyond weighted Jacobi + the Krylov wrapper (oracle/pmg_oracle_smoothers.c, ref_driver.cpp)
        g("gs").restype = _i
        g("gs").argtypes = [_dp, _dp, _i, _i, _d, _i, _d, _dp]
        g("cycle_s").restype = _i
        g("cycle_s").argtypes = [_dp, _dp, _i, _d, _i, _i, _d, _i, _i, _i, _i, _i]
        g("cg").restype = _i
        g("cg").argtypes = [_dp, _dp, _i, _i, _d, _i, _d, _dp]
        if prefix == "orc_":
            lib.orc_rbgs.restype = _i
            lib.orc_rbgs.argtypes = [_dp, _dp, _i, _i, _d, _i]
            lib.orc_jacobi_weights.restype = _i
            lib.orc_jacobi_weights.argtypes = [_dp, _dp, _i, _i, _d, _dp, _i]
            lib.orc_chebyshev_weights.restype = None
            lib.orc_chebyshev_weights.argtypes = [_d, _d, _i, _dp]
            lib.orc_pcg.restype = _i
            lib.orc_pcg.argtypes = [_dp, _dp, _i, _d, _i, _i, _d, _i, _i, _i, _i, _d, _i, _dp]
        self._g = g

    def jacobi(self, x, f, h, omega=1.0, num_iter=1, eps=0.0):
        """num_iter+1 sweeps in place on x; returns the smoother's own per-sweep ||r|| list."""
        res = np.zeros(num_iter + 1)
        n = self._g("jacobi")(_p(x), _p(f), x.shape[1], x.shape[0], h, omega, num_iter, eps, _p(res))
        return res[:n]

    def residual(self, x, f, h):
        r = np.zeros_like(x)
        self._g("residual")(_p(r), _p(x), _p(f), x.shape[1], x.shape[0], h)
        return r

    def norm(self, v):
        return self._g("norm")(_p(v), v.size)

    def restrict_fw(self, fine):
        nf = fine.shape[0]
        nc = (nf - 1) // 2 + 1
        coarse = np.zeros((nc, nc))
        self._g("restrict_fw")(_p(fine), _p(coarse), nf, nc)
        return coarse

    def prolong_add(self, fine, coarse, mode=PROLONG_REFERENCE):
        rc = self._g("prolong_add")(_p(fine), _p(coarse), fine.shape[0], coarse.shape[0], mode)
        if rc != 0:
            raise NotImplementedError("prolong mode %d not available in %s" % (mode, self.prefix))
        return fine

    def rhs(self, n):
        f = np.zeros((n, n))
        self._g("rhs")(_p(f), n, n, 1.0 / (n - 1))
        return f

    def exact(self, n):
        u = np.zeros((n, n))
        self._g("exact")(_p(u), 1.0 / (n - 1), n, n)
        return u

    def cycle(self, phi, f, kind=V, omega=2.0 / 3.0, eps=0.0, alpha=2, v1=1, v2=1,
              prolong=PROLONG_REFERENCE):
        n = phi.shape[0]
        rc = self._g("cycle")(_p(phi), _p(f), n, 1.0 / (n - 1), kind, omega, eps, alpha, v1, v2, prolong)
        if rc != 0:
            raise NotImplementedError("cycle configuration not available in %s" % self.prefix)
        return phi

    def fmg_general(self, phi, f, omega=2.0 / 3.0, v1=1, v2=1, prolong=PROLONG_REFERENCE):
        """One general-RHS full-multigrid pass in place on phi (ring kept, interior restarted from 0) -- not a
        reference function: the specification of PMG_CYCLE_FMG (oracle.h)."""
        n = phi.shape[0]
        rc = self._g("fmg_general")(_p(phi), _p(f), n, 1.0 / (n - 1), omega, v1, v2, prolong)
        if rc != 0:
            raise NotImplementedError("fmg_general not available in %s" % self.prefix)
        return phi

    # ---- smoothers beyond weighted Jacobi ----
    def gs(self, x, f, h, num_iter, eps=0.0):
        """GaussSeidelSmoother::smooth: num_iter lexicographic sweeps in place; returns the per-sweep ||r|| list."""
        res = np.zeros(max(num_iter, 1))
        n = self._g("gs")(_p(x), _p(f), x.shape[1], x.shape[0], h, num_iter, eps, _p(res))
        return res[:n]

    def rbgs(self, x, f, h, sweeps):
        self.lib.orc_rbgs(_p(x), _p(f), x.shape[1], x.shape[0], h, sweeps)
        return x

    def chebyshev_weights(self, n, lo=0.5, hi=2.0):
        w = np.zeros(n)
        self.lib.orc_chebyshev_weights(lo, hi, n, _p(w))
        return w

    def jacobi_weights(self, x, f, h, w):
        w = np.ascontiguousarray(w, dtype=np.float64)
        self.lib.orc_jacobi_weights(_p(x), _p(f), x.shape[1], x.shape[0], h, _p(w), len(w))
        return x

    def cycle_s(self, phi, f, kind=V, smoother=SMOOTHER_JACOBI, omega=2.0 / 3.0, alpha=2, nu1=2, nu2=2, coarse_sweeps=11,
                prolong=PROLONG_REFERENCE):
        n = phi.shape[0]
        rc = self._g("cycle_s")(_p(phi), _p(f), n, 1.0 / (n - 1), kind, smoother, omega, alpha, nu1, nu2, coarse_sweeps,
                                prolong)
        if rc != 0:
            raise NotImplementedError("cycle_s configuration not available in %s" % self.prefix)
        return phi

    def cg(self, x, f, h, num_iter, eps=0.0):
        """ConjugateGradientSmoother::smooth: x zeroed, num_iter steps; returns the ||r|| list (before + after each step)."""
        res = np.zeros(num_iter + 2)
        n = self._g("cg")(_p(x), _p(f), x.shape[1], x.shape[0], h, num_iter, eps, _p(res))
        return res[:n]

    def pcg(self, x, f, precond=1, smoother=SMOOTHER_JACOBI, omega=2.0 / 3.0, nu1=2, nu2=2, coarse_sweeps=11,
            prolong=PROLONG_FULL, rel_tol=1e-8, max_iter=100):
        n = x.shape[0]
        hist = np.zeros(max_iter + 1)
        k = self.lib.orc_pcg(_p(x), _p(f), n, 1.0 / (n - 1), precond, smoother, omega, nu1, nu2, coarse_sweeps, prolong,
                             rel_tol, max_iter, _p(hist))
        return k, hist[: k + 1].copy()

    def solve(self, phi, f, kind=V, omega=2.0 / 3.0, eps=0.0, alpha=2, v1=1, v2=1,
              prolong=PROLONG_REFERENCE, rel_tol=1e-8, max_cycles=100):
        """Returns (n_cycles, history) with history[0] = ||r0|| and history[k] after cycle k."""
        hist = np.zeros(max_cycles + 1)
        k = self._g("solve")(_p(phi), _p(f), phi.shape[0], kind, omega, eps, alpha, v1, v2, prolong,
                             rel_tol, max_cycles, _p(hist))
        if k < 0:
            raise NotImplementedError("solve configuration not available in %s" % self.prefix)
        return k, hist[: k + 1].copy()


def build():
    """(Re)build the checkers with oracle/Makefile (also builds _ref when /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True, stdout=subprocess.DEVNULL)


def load(which="orc"):
    if which == "orc":
        if not os.path.exists(ORC_SO):
            build()
        return Checker(ctypes.CDLL(ORC_SO), "orc_")
    if which == "ref":
        if not os.path.exists(REF_SO):
            if os.path.isdir("/root/reference"):
                build()
            if not os.path.exists(REF_SO):
                return None
        lib = ctypes.CDLL(REF_SO)
        lib.ref_use_shipped_jacobi.argtypes = [_i]
        return Checker(lib, "ref_")
    raise ValueError(which)


def random_rhs(n, seed=12345):
    """RHS B of SURVEY.md 8d: U(-1,1) on the interior, 0 on the ring (numpy PCG64, seeded)."""
    f = np.zeros((n, n))
    f[1:-1, 1:-1] = np.random.default_rng(seed).uniform(-1.0, 1.0, (n - 2, n - 2))
    return f
