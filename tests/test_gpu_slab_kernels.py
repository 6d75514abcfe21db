"""The multi-GPU flavours of the fused passes on ONE GPU.

Row slabs of 2-4 "ranks" live side by side in the same device memory; each rank's Pass A reads its neighbours' halo
rows through HaloPeers exactly as it does over NVLink (flags pre-set, epoch publication checked) -- streamed in place
(default) or copied by the halo prologue (opt-in flavour) -- with the interior / boundary split of `cycle_dist`, and
Pass B finishes the rows beyond the slab.  Every output is compared BIT FOR BIT with the oracle's operator sequence on
the whole level; halo rows that must come from the neighbour are poisoned with NaN locally.  This is the GPU twin of
tests/cpp/test_fused_kernel_emu.cpp (which runs the same kernel source on the CPU) and gives the slab code paths
hardware coverage on a single-GPU box, where tests/test_gpu_dist.py has to skip.
"""
import ctypes

import numpy as np
import pytest

import cpu_checkers as cc
import pmg_b200 as pmg

pytestmark = pytest.mark.gpu

PADX, PADY = 16, 8
OMEGA = 2.0 / 3.0
EPOCH = 7


class TestSlab(ctypes.Structure):
    __test__ = False
    _fields_ = [("x", ctypes.c_void_p), ("xb", ctypes.c_void_p), ("f", ctypes.c_void_p), ("n", ctypes.c_int),
                ("pitch", ctypes.c_int), ("h", ctypes.c_double), ("ny", ctypes.c_int), ("yoff", ctypes.c_int),
                ("ext_lo", ctypes.c_int), ("ext_hi", ctypes.c_int), ("span_lo", ctypes.c_int), ("span_hi", ctypes.c_int),
                ("x_up", ctypes.c_void_p), ("x_dn", ctypes.c_void_p), ("f_up", ctypes.c_void_p), ("f_dn", ctypes.c_void_p),
                ("f_keep", ctypes.c_void_p), ("x_keep", ctypes.c_void_p), ("flag_up", ctypes.c_void_p),
                ("flag_dn", ctypes.c_void_p), ("pub_up", ctypes.c_void_p), ("pub_dn", ctypes.c_void_p),
                ("epoch", ctypes.c_int), ("err", ctypes.c_void_p)]


def _lib():
    L = pmg.lib()
    L.pmg_test_layout.restype = None
    L.pmg_test_layout.argtypes = [ctypes.c_int] + [ctypes.POINTER(ctypes.c_int)] * 3
    L.pmg_test_fused_down.argtypes = [ctypes.POINTER(TestSlab), ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_double, ctypes.c_int, ctypes.c_int]
    L.pmg_test_fused_up.argtypes = [ctypes.POINTER(TestSlab), ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_double, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
    L.pmg_test_fused_max_partials.argtypes = [ctypes.c_int]
    return L


def level_pitch(n):
    p, px, py = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib().pmg_test_layout(n, ctypes.byref(p), ctypes.byref(px), ctypes.byref(py))
    assert (px.value, py.value) == (PADX, PADY)
    return p.value


class Padded:
    """`rows` rows of an n-column level in the solver's padded layout, host copy + device copy"""

    def __init__(self, n, rows):
        self.n, self.rows, self.pitch = n, rows, level_pitch(n)
        self.h = np.zeros((rows + 2 * PADY, self.pitch))
        self.d = None

    def view(self, a, b):  # logical columns of local rows [a, b)
        return self.h[PADY + a:PADY + b, PADX:PADX + self.n]

    def upload(self):
        self.d = pmg.DeviceArray.from_numpy(self.h)
        return self

    def download(self):
        self.h = self.d.numpy()
        return self

    def ptr(self, row=0):  # device address of logical (row, 0)
        return self.d.ptr + 8 * ((PADY + row) * self.pitch + PADX)


class DevInts:
    def __init__(self, values):
        self.a = np.array(values, dtype=np.int32)
        p = ctypes.c_void_p()
        pmg.check(pmg.lib().pmg_device_alloc(ctypes.byref(p), self.a.nbytes))
        self.ptr = p.value
        pmg.check(pmg.lib().pmg_memcpy(self.ptr, self.a.ctypes.data, self.a.nbytes, 1, 0))

    def get(self):
        pmg.check(pmg.lib().pmg_memcpy(self.a.ctypes.data, self.ptr, self.a.nbytes, 0, 1))
        return self.a.copy()

    def at(self, i):
        return self.ptr + 4 * i

    def __del__(self):
        import sys
        if not sys.is_finalizing() and getattr(self, "ptr", None):
            pmg.lib().pmg_device_free(self.ptr)
            self.ptr = None


def partition(n, ranks, r):
    pairs = (n - 1) // 2
    a, b = pairs * r // ranks, pairs * (r + 1) // ranks
    return 2 * a, (n if r == ranks - 1 else 2 * b)


def expected_visit(orc, x, f, e, h, nu1, nu2, prolong, first_visit):
    xb = np.zeros_like(x) if first_visit else x.copy()
    orc.jacobi(xb, f, h, omega=OMEGA, num_iter=nu1 - 1)
    cf = orc.restrict_fw(orc.residual(xb, f, h))
    xn = xb.copy()
    orc.prolong_add(xn, e, prolong)
    orc.jacobi(xn, f, h, omega=OMEGA, num_iter=nu2 - 1)
    r = orc.residual(xn, f, h)
    return xb, cf, xn, orc.norm(r) ** 2


@pytest.mark.parametrize("n,ranks,first_visit,split,prolong,level0,prologue", [
    (257, 2, False, False, pmg.PROLONG_REFERENCE, True, 0),   # finest level: the iterate is exchanged
    (257, 2, True, False, pmg.PROLONG_REFERENCE, False, 0),   # coarse level, first visit: f is exchanged
    (257, 3, False, False, pmg.PROLONG_FULL, False, 0),       # W re-visit of a coarse level, a middle rank
    (513, 4, False, True, pmg.PROLONG_REFERENCE, True, 0),    # interior / boundary split, two middle ranks
    (257, 2, False, False, pmg.PROLONG_REFERENCE, True, 1),   # the same four with the halo prologue
    (257, 2, True, False, pmg.PROLONG_REFERENCE, False, 1),
    (257, 3, False, False, pmg.PROLONG_FULL, False, 1),
    (513, 4, False, True, pmg.PROLONG_REFERENCE, True, 1),
])
def test_slab_passes_match_oracle_on_one_gpu(orc, n, ranks, first_visit, split, prolong, level0, prologue):
    L = _lib()
    nc, nu1, nu2 = (n - 1) // 2 + 1, 2, 2
    h = 1.0 / (n - 1)
    rng = np.random.default_rng(100 + n + ranks)
    x = rng.uniform(-1, 1, (n, n))
    if not level0:
        x[0, :] = x[-1, :] = x[:, 0] = x[:, -1] = 0.0
    f = np.zeros((n, n))
    f[1:-1, 1:-1] = rng.uniform(-1, 1, (n - 2, n - 2))
    e = np.zeros((nc, nc))
    e[1:-1, 1:-1] = rng.uniform(-1, 1, (nc - 2, nc - 2))
    want_xb, want_cf, want_xn, want_norm2 = expected_visit(orc, x, f, e, h, nu1, nu2, prolong, first_visit)

    R = []
    for r in range(ranks):
        y0, y1 = partition(n, ranks, r)
        c0, c1 = y0 // 2, (nc if y1 == n else y1 // 2)  # nested partition, as pmg_create enforces
        ny, nyc = y1 - y0, c1 - c0
        up, dn = r > 0, r < ranks - 1
        K = dict(y0=y0, y1=y1, ny=ny, c0=c0, nyc=nyc, up=up, dn=dn)
        px, pxb, pf, pcf, pe, pxn = Padded(n, ny), Padded(n, ny), Padded(n, ny), Padded(nc, nyc), Padded(nc, nyc), Padded(n, ny)
        px.view(0, ny)[:] = np.nan if first_visit else x[y0:y1]
        pf.view(0, ny)[:] = f[y0:y1]
        if up:
            px.view(-PADY, 0)[:] = np.nan
        if dn:
            px.view(ny, ny + PADY)[:] = np.nan
        if level0:  # the finest level's f halo is local
            a = max(0, y0 - PADY)
            pf.view(a - y0, 0)[:] = f[a:y0]
            b = min(n, y1 + PADY)
            pf.view(ny, b - y0)[:] = f[y1:b]
        else:
            if up:
                pf.view(-PADY, 0)[:] = np.nan
            if dn:
                pf.view(ny, ny + PADY)[:] = np.nan
        pe.view(-PADY, nyc + PADY)[:] = np.nan
        a, b = max(0, c0 - 4), min(nc, c0 + nyc + 4)
        pe.view(a - c0, b - c0)[:] = e[a:b]
        if c0 - 4 < 0:
            pe.view(-PADY, 0)[:] = 0.0
        if c0 + nyc + 4 > nc:
            pe.view(nyc, nyc + PADY)[:] = 0.0
        for p in (px, pxb, pf, pcf, pe, pxn):
            p.upload()
        K.update(x=px, xb=pxb, f=pf, cf=pcf, e=pe, xn=pxn, inbox=DevInts([EPOCH, EPOCH]), outbox=DevInts([0, 0]),
                 err=DevInts([0]))
        R.append(K)

    # ---- Pass A ----
    for r, K in enumerate(R):
        up, dn, ny = K["up"], K["dn"], K["ny"]
        t = TestSlab()
        t.x, t.xb, t.f = K["x"].ptr(), K["xb"].ptr(), K["f"].ptr()
        t.n, t.pitch, t.h, t.ny, t.yoff = n, K["x"].pitch, h, ny, K["y0"]
        peers = dict(epoch=EPOCH, err=K["err"].ptr)
        if not first_visit:
            peers.update(x_up=R[r - 1]["x"].ptr(R[r - 1]["ny"]) if up else None, x_dn=R[r + 1]["x"].ptr() if dn else None,
                         x_keep=K["x"].ptr())
        if not level0:
            peers.update(f_up=R[r - 1]["f"].ptr(R[r - 1]["ny"]) if up else None, f_dn=R[r + 1]["f"].ptr() if dn else None,
                         f_keep=K["f"].ptr())
        peers.update(flag_up=K["inbox"].at(0) if up else None, flag_dn=K["inbox"].at(1) if dn else None,
                     pub_up=K["outbox"].at(0) if up else None, pub_dn=K["outbox"].at(1) if dn else None)

        def launch(with_peers, lo, hi, publish=True):
            for k in ("x_up", "x_dn", "f_up", "f_dn", "f_keep", "x_keep", "flag_up", "flag_dn", "pub_up", "pub_dn", "err"):
                setattr(t, k, None)
            t.epoch = 0
            if with_peers:
                for k, v in peers.items():
                    if k.startswith("pub_") and not publish:
                        continue
                    setattr(t, k, v)
            t.span_lo, t.span_hi = lo, hi
            pmg.check(L.pmg_test_fused_down(ctypes.byref(t), K["cf"].ptr(), K["cf"].pitch, nu1, OMEGA,
                                            1 if first_visit else 0, prologue))

        if split:  # cycle_dist: boundary strips [-6, 8), [ny - 8, ny + 6) with the peers, interior without
            first = True
            if up:
                launch(True, -6, PADY, publish=first)
                first = False
            if dn:
                launch(True, ny - PADY, ny + 6, publish=first)
            launch(False, PADY if up else 0, ny - PADY if dn else ny)
        else:
            launch(True, -6 if up else 0, ny + 6 if dn else ny)
        assert K["err"].get()[0] == 0
        out = K["outbox"].get()
        assert (not up or out[0] == EPOCH) and (not dn or out[1] == EPOCH), "epoch not published"
        a, b = (-6 if up else 0), (ny + 6 if dn else ny)
        got = K["xb"].download()
        assert np.array_equal(got.view(a, b), want_xb[K["y0"] + a:K["y0"] + b]), "rank %d: xb" % r
        outside = got.h.copy()
        outside[PADY + a:PADY + b, PADX:PADX + n] = 0.0
        assert not outside.any(), "rank %d: Pass A wrote xb outside [-6, ny + 6)" % r
        gcf = K["cf"].download()
        assert np.array_equal(gcf.view(0, K["nyc"]), want_cf[K["c0"]:K["c0"] + K["nyc"]]), "rank %d: coarse f" % r
        outside = gcf.h.copy()
        outside[PADY:PADY + K["nyc"], PADX:PADX + nc] = 0.0
        assert not outside.any(), "rank %d: Pass A wrote the coarse array outside its owned rows" % r

    # ---- Pass B ----
    total = 0.0
    maxp = L.pmg_test_fused_max_partials(n)
    for r, K in enumerate(R):
        up, dn, ny = K["up"], K["dn"], K["ny"]
        t = TestSlab()
        t.x, t.xb, t.f = K["xn"].ptr(), K["xb"].ptr(), K["f"].ptr()
        t.n, t.pitch, t.h, t.ny, t.yoff = n, K["x"].pitch, h, ny, K["y0"]
        ext = 0 if level0 else 4
        t.ext_lo, t.ext_hi = (ext if up else 0), (ext if dn else 0)
        partials = pmg.DeviceArray.from_numpy(np.zeros(maxp))
        np_out = ctypes.c_int()
        pmg.check(L.pmg_test_fused_up(ctypes.byref(t), K["e"].ptr(), K["e"].pitch, nu2, OMEGA, prolong,
                                      partials.ptr if level0 else None, ctypes.byref(np_out)))
        a, b = -t.ext_lo, ny + t.ext_hi
        got = K["xn"].download()
        assert np.array_equal(got.view(a, b), want_xn[K["y0"] + a:K["y0"] + b]), "rank %d: x after Pass B" % r
        outside = got.h.copy()
        outside[PADY + a:PADY + b, PADX:PADX + n] = 0.0
        assert not outside.any(), "rank %d: Pass B wrote x outside its rows" % r
        total += float(partials.numpy()[:np_out.value].sum())
    if level0 and not first_visit:
        assert abs(total - want_norm2) <= 1e-12 * want_norm2


def _lib_cross():
    L = _lib()
    L.pmg_test_fused_cross.argtypes = [ctypes.POINTER(TestSlab), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.POINTER(ctypes.c_int)]
    L.pmg_set_p2p_timeout_ms.restype = None
    L.pmg_set_p2p_timeout_ms.argtypes = [ctypes.c_double]
    return L


def _cross_ranks(orc, n, ranks, prolong, seed):
    """inputs of one cross-cycle pass on `ranks` row slabs + what the oracle's operator sequence gives on the whole level:
    x_k = S^2 (xb + P e), ||f - A x_k||^2, xb' = S^2 x_k, coarse f = R (f - A xb')"""
    nc, h = (n - 1) // 2 + 1, 1.0 / (n - 1)
    rng = np.random.default_rng(seed)
    xb = rng.uniform(-1, 1, (n, n))
    f = np.zeros((n, n))
    f[1:-1, 1:-1] = rng.uniform(-1, 1, (n - 2, n - 2))
    e = np.zeros((nc, nc))
    e[1:-1, 1:-1] = rng.uniform(-1, 1, (nc - 2, nc - 2))
    xk = xb.copy()
    orc.prolong_add(xk, e, prolong)
    orc.jacobi(xk, f, h, omega=OMEGA, num_iter=1)
    norm2 = orc.norm(orc.residual(xk, f, h)) ** 2
    xb2 = xk.copy()
    orc.jacobi(xb2, f, h, omega=OMEGA, num_iter=1)
    cf = orc.restrict_fw(orc.residual(xb2, f, h))
    R = []
    for r in range(ranks):
        y0, y1 = partition(n, ranks, r)
        c0, c1 = y0 // 2, (nc if y1 == n else y1 // 2)
        ny, nyc = y1 - y0, c1 - c0
        up, dn = r > 0, r < ranks - 1
        K = dict(y0=y0, y1=y1, ny=ny, c0=c0, nyc=nyc, up=up, dn=dn)
        pxb, pxk, pout, pf, pcf, pe = (Padded(n, ny), Padded(n, ny), Padded(n, ny), Padded(n, ny), Padded(nc, nyc),
                                       Padded(nc, nyc))
        pxb.view(0, ny)[:] = xb[y0:y1]  # the input array: owned rows; the halo rows must come from the neighbours
        if up:
            pxb.view(-PADY, 0)[:] = np.nan
        if dn:
            pxb.view(ny, ny + PADY)[:] = np.nan
        a, b = max(0, y0 - PADY), min(n, y1 + PADY)  # level 0: the f halo is local
        pf.view(a - y0, b - y0)[:] = f[a:b]
        pe.view(-PADY, nyc + PADY)[:] = np.nan
        a, b = max(0, c0 - 4), min(nc, c0 + nyc + 4)
        pe.view(a - c0, b - c0)[:] = e[a:b]
        if c0 - 4 < 0:
            pe.view(-PADY, 0)[:] = 0.0
        if c0 + nyc + 4 > nc:
            pe.view(nyc, nyc + PADY)[:] = 0.0
        for p in (pxb, pxk, pout, pf, pcf, pe):
            p.upload()
        K.update(xb=pxb, xk=pxk, out=pout, f=pf, cf=pcf, e=pe, inbox=DevInts([EPOCH, EPOCH]), outbox=DevInts([0, 0]),
                 err=DevInts([0]))
        R.append(K)
    return R, dict(xk=xk, xb2=xb2, cf=cf, norm2=norm2, nc=nc, h=h)


def _cross_slab(R, r, n, h):
    K = R[r]
    t = TestSlab()
    t.x, t.xb, t.f = K["xk"].ptr(), K["xb"].ptr(), K["f"].ptr()
    t.n, t.pitch, t.h, t.ny, t.yoff = n, K["xb"].pitch, h, K["ny"], K["y0"]
    t.span_lo, t.span_hi = 0, K["ny"]
    t.x_up = R[r - 1]["xb"].ptr(R[r - 1]["ny"]) if K["up"] else None
    t.x_dn = R[r + 1]["xb"].ptr() if K["dn"] else None
    t.x_keep = K["xb"].ptr()
    t.flag_up = K["inbox"].at(0) if K["up"] else None
    t.flag_dn = K["inbox"].at(1) if K["dn"] else None
    t.pub_up = K["outbox"].at(0) if K["up"] else None
    t.pub_dn = K["outbox"].at(1) if K["dn"] else None
    t.epoch, t.err = EPOCH, K["err"].ptr
    return t


@pytest.mark.parametrize("n,ranks,prolong", [
    (257, 2, pmg.PROLONG_REFERENCE),
    (257, 3, pmg.PROLONG_FULL),
    (513, 4, pmg.PROLONG_REFERENCE),
    (1025, 8, pmg.PROLONG_REFERENCE),
])
def test_cross_cycle_pass_on_slabs_matches_oracle_on_one_gpu(orc, n, ranks, prolong):
    """GPU twin of slab_cross in tests/cpp/test_fused_kernel_emu.cpp: Pass B of cycle k + Pass A of cycle k+1 in one
    sweep over a row slab, the input's halo rows pulled from the neighbours by the halo prologue"""
    L = _lib_cross()
    R, W = _cross_ranks(orc, n, ranks, prolong, 300 + n + ranks)
    maxp = L.pmg_test_fused_max_partials(n)
    total = 0.0
    for r, K in enumerate(R):
        t = _cross_slab(R, r, n, W["h"])
        partials = pmg.DeviceArray.from_numpy(np.zeros(maxp))
        np_out = ctypes.c_int()
        pmg.check(L.pmg_test_fused_cross(ctypes.byref(t), K["out"].ptr(), K["e"].ptr(), K["cf"].ptr(), K["cf"].pitch, OMEGA,
                                         prolong, partials.ptr, ctypes.byref(np_out)))
        assert K["err"].get()[0] == 0
        out = K["outbox"].get()
        assert (not K["up"] or out[0] == EPOCH) and (not K["dn"] or out[1] == EPOCH), "epoch not published"
        ny, y0 = K["ny"], K["y0"]
        for name, want in (("out", W["xb2"]), ("xk", W["xk"])):
            got = K[name].download()
            assert np.array_equal(got.view(0, ny), want[y0:y0 + ny]), "rank %d: %s" % (r, name)
            outside = got.h.copy()
            outside[PADY:PADY + ny, PADX:PADX + n] = 0.0
            assert not outside.any(), "rank %d: the cross pass wrote %s outside the owned rows" % (r, name)
        gcf = K["cf"].download()
        assert np.array_equal(gcf.view(0, K["nyc"]), W["cf"][K["c0"]:K["c0"] + K["nyc"]]), "rank %d: coarse f" % r
        outside = gcf.h.copy()
        outside[PADY:PADY + K["nyc"], PADX:PADX + W["nc"]] = 0.0
        assert not outside.any(), "rank %d: the cross pass wrote the coarse array outside its owned rows" % r
        total += float(partials.numpy()[:np_out.value].sum())
    assert abs(total - W["norm2"]) <= 1e-12 * W["norm2"]


def test_peer_flag_wait_times_out_and_reports(orc):
    """A neighbour that never publishes its epoch: the wait gives up after the wall-time limit (PMG_P2P_TIMEOUT_S; 50 ms
    here), raises the error word and the kernel still terminates -- the path pmg_solve turns into PMG_ERR_COMM"""
    import time
    L = _lib_cross()
    n, ranks = 257, 2
    R, W = _cross_ranks(orc, n, ranks, pmg.PROLONG_REFERENCE, 77)
    K = R[1]
    K["inbox"] = DevInts([EPOCH - 1, EPOCH - 1])  # the upper neighbour is one epoch behind
    t = _cross_slab(R, 1, n, W["h"])
    partials = pmg.DeviceArray.from_numpy(np.zeros(L.pmg_test_fused_max_partials(n)))
    np_out = ctypes.c_int()
    L.pmg_set_p2p_timeout_ms(50.0)
    try:
        t0 = time.perf_counter()
        pmg.check(L.pmg_test_fused_cross(ctypes.byref(t), K["out"].ptr(), K["e"].ptr(), K["cf"].ptr(), K["cf"].pitch, OMEGA,
                                         pmg.PROLONG_REFERENCE, partials.ptr, ctypes.byref(np_out)))
        dt = time.perf_counter() - t0
        assert K["err"].get()[0] != 0, "the time-out did not raise the error word"
        assert 0.04 <= dt < 5.0, "waited %.3f s for a 50 ms limit" % dt
        # the same wait in Pass A (halo prologue flavour)
        K["err"] = DevInts([0])
        t.err = K["err"].ptr
        t.span_lo, t.span_hi = -6, K["ny"]
        pmg.check(L.pmg_test_fused_down(ctypes.byref(t), K["cf"].ptr(), K["cf"].pitch, 2, OMEGA, 0, 1))
        assert K["err"].get()[0] != 0
    finally:
        L.pmg_set_p2p_timeout_ms(0.0)  # back to the default
