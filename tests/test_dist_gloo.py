"""World-size-2 (and 3) CPU tests of the multi-GPU DESIGN over gloo: the row-slab decomposition, the halo
depths (8 rows exchanged on the way down, 6 / 4 rows recomputed instead of exchanged on the way up), the
gather / scatter of the agglomerated level and the rank-ordered norm -- exactly the schedule of
`cycle_dist` in csrc/solver.cu, executed with the CPU oracle's operators on each rank's window.

The check is the one the GPU script tests/dist_check.py makes on real hardware: the gathered iterate after
several cycles is BIT-IDENTICAL to the single-process oracle.  Rows a pass does not write are poisoned with
NaN, so a halo that is one row too shallow shows up as a NaN in an owned row.
"""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

PADY = 8  # halo rows kept / exchanged per slab (csrc/pmg_internal.h)
EXT_A, EXT_B = 6, 4  # rows Pass A / Pass B also finish next to each neighbour (csrc/solver.cu: cycle_dist)


# ---- window operators (numpy, same evaluation order as the oracle; checked against it below) ----------
def restrict_rows(r, g0, n, jc_lo, jc_hi):
    """full weighting (MultiGrid.hpp:187-205) of the window r (global fine rows g0..) for coarse rows [jc_lo, jc_hi)"""
    nc = (n - 1) // 2 + 1
    out = np.zeros((jc_hi - jc_lo, nc))
    for jc in range(max(jc_lo, 1), min(jc_hi, nc - 1)):
        j = 2 * jc - g0
        c, s, nn = r[j], r[j - 1], r[j + 1]
        ic = np.arange(1, nc - 1)
        i = 2 * ic
        edge = ((c[i + 1] + c[i - 1]) + nn[i]) + s[i]
        corner = ((s[i - 1] + s[i + 1]) + nn[i - 1]) + nn[i + 1]
        out[jc - jc_lo, 1:nc - 1] = (0.25 * c[i] + 0.125 * edge) + 0.0625 * corner
    return out


def prolong_rows(x, g0, e, c0, n, lo):
    """x (window, global fine rows g0..) += P e (window, global coarse rows c0..) -- MultiGrid.hpp:208-226;
    lo = 2 (reference: fine row / col 1 skipped) or 1"""
    rows = x.shape[0]
    cols = np.arange(lo, n - 1)
    ic = cols // 2
    even = (cols % 2) == 0
    for j in range(rows):
        g = g0 + j
        if g < lo or g > n - 2:
            continue
        jc = g // 2 - c0
        if jc < 0 or jc + 1 >= e.shape[0]:
            x[j, :] = np.nan  # not computable from this window
            continue
        a, b = e[jc], e[jc + 1]
        if g % 2 == 0:
            corr = np.where(even, a[ic], 0.5 * (a[ic] + a[ic + 1]))
        else:
            corr = np.where(even, 0.5 * (a[ic] + b[ic]), 0.25 * (((a[ic] + a[ic + 1]) + b[ic]) + b[ic + 1]))
        x[j, cols] = x[j, cols] + corr


def partition(n, ranks, r):
    import pmg_b200 as pmg
    return pmg.partition_rows(n, ranks, r)


class Slab:
    """one rank's window of one level: global rows [w0, w1) around the owned rows [y0, y1)"""

    def __init__(self, n, rank, world):
        self.n = n
        self.y0, self.y1 = partition(n, world, rank)
        self.w0, self.w1 = max(0, self.y0 - PADY), min(n, self.y1 + PADY)
        self.x = np.zeros((self.w1 - self.w0, n))
        self.xb = np.zeros_like(self.x)
        self.f = np.zeros_like(self.x)

    def rows(self, arr, a, b):  # view of global rows [a, b)
        return arr[a - self.w0:b - self.w0]

    def poison_outside(self, arr, ext, rank, world):
        lo = self.y0 - (ext if rank > 0 else 0)
        hi = self.y1 + (ext if rank < world - 1 else 0)
        arr[:max(lo, self.w0) - self.w0] = np.nan if rank > 0 else arr[:max(lo, self.w0) - self.w0]
        if rank < world - 1:
            arr[hi - self.w0:] = np.nan


def halo_exchange(slab, arr, rank, world):
    """PADY owned rows to each neighbour's halo (comm_halo_exchange in csrc/comm.cu)"""
    reqs = []
    up_recv = torch.zeros((PADY, slab.n), dtype=torch.float64)
    dn_recv = torch.zeros((PADY, slab.n), dtype=torch.float64)
    if rank > 0:
        reqs.append(dist.isend(torch.from_numpy(slab.rows(arr, slab.y0, slab.y0 + PADY).copy()), rank - 1))
        reqs.append(dist.irecv(up_recv, rank - 1))
    if rank < world - 1:
        reqs.append(dist.isend(torch.from_numpy(slab.rows(arr, slab.y1 - PADY, slab.y1).copy()), rank + 1))
        reqs.append(dist.irecv(dn_recv, rank + 1))
    for q in reqs:
        q.wait()
    if rank > 0:
        slab.rows(arr, slab.y0 - PADY, slab.y0)[:] = up_recv.numpy()
    if rank < world - 1:
        slab.rows(arr, slab.y1, slab.y1 + PADY)[:] = dn_recv.numpy()


def cycle_dist(orc, slabs, l, la, rank, world, omega, gamma, w_form, x_is_zero, lo):
    import cpu_checkers as cc
    L, K = slabs[l], slabs[l + 1]
    n, h = L.n, 1.0 / (L.n - 1)
    if not x_is_zero:
        halo_exchange(L, L.x, rank, world)
    if l > 0 and x_is_zero:
        halo_exchange(L, L.f, rank, world)
    # Pass A: xb = S^2 x on the window, valid EXT_A rows beyond the slab; coarse RHS for owned coarse rows
    xb = np.zeros_like(L.x) if x_is_zero else L.x.copy()
    orc.jacobi(xb, L.f, h, omega=omega, num_iter=1)
    r = orc.residual(xb, L.f, h)
    K.f[:] = 0.0
    K.rows(K.f, K.y0, K.y1)[:] = restrict_rows(r, L.w0, n, K.y0, K.y1)
    L.xb[:] = xb
    L.poison_outside(L.xb, EXT_A, rank, world)
    reps = gamma if w_form else 1
    if l + 1 < la:
        for k in range(reps):
            cycle_dist(orc, slabs, l + 1, la, rank, world, omega, gamma, w_form, k == 0, lo)
    else:
        # gather owned rows of the first agglomerated level to rank 0, solve there, scatter with 4 halo rows
        nc = K.n
        parts = [None] * world  # slabs are uneven (the last rank owns the odd final row)
        dist.all_gather_object(parts, K.rows(K.f, K.y0, K.y1).copy())
        full_x = torch.zeros((nc, nc), dtype=torch.float64)
        if rank == 0:
            f_full = np.concatenate(parts)
            e = np.zeros((nc, nc))
            for k in range(reps):
                orc.cycle(e, f_full, kind=cc.W if w_form else cc.V, omega=omega, eps=0.0, alpha=gamma,
                          prolong=cc.PROLONG_FULL if lo == 1 else cc.PROLONG_REFERENCE)
            full_x = torch.from_numpy(e)
        dist.broadcast(full_x, src=0)
        K.x[:] = np.nan
        a, b = max(0, K.y0 - 4), min(nc, K.y1 + 4)
        K.rows(K.x, a, b)[:] = full_x.numpy()[a:b]
    # Pass B: x = S^2 (xb + P e), valid EXT_B rows beyond the slab
    x = L.xb.copy()
    prolong_rows(x, L.w0, K.x, K.w0, n, lo)
    xs = np.nan_to_num(x, nan=1e300)  # the oracle's norm would choke on NaN; poison survives as 1e300
    orc.jacobi(xs, L.f, h, omega=omega, num_iter=1)
    xs[np.abs(xs) > 1e200] = np.nan
    L.x[:] = xs
    L.poison_outside(L.x, 0 if l == 0 else EXT_B, rank, world)


def _worker(rank, world, port, n, agg_below, w_form, gamma, lo, q):
    import cpu_checkers as cc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = cc.load("orc")
    omega = 2.0 / 3.0
    sizes = [n]
    while sizes[-1] > 5:
        sizes.append((sizes[-1] - 1) // 2 + 1)
    la = 0
    while la < len(sizes) - 1 and sizes[la] > agg_below and ((sizes[la] - 1) // 2) // world * 2 >= 4 * PADY:
        la += 1
    assert la >= 1
    slabs = [Slab(m, rank, world) for m in sizes[:la + 1]]
    f = cc.random_rhs(n, seed=5)
    top = slabs[0]
    top.f[:] = f[top.w0:top.w1]  # set_rhs + its halo exchange
    want = np.zeros((n, n))
    norms = []
    for cyc in range(3):
        cycle_dist(orc, slabs, 0, la, rank, world, omega, gamma, w_form, False, lo)
        orc.cycle(want, f, kind=cc.W if w_form else cc.V, omega=omega, eps=0.0, alpha=gamma,
                  prolong=cc.PROLONG_FULL if lo == 1 else cc.PROLONG_REFERENCE)
        # distributed norm: owned interior rows, combined in rank order (dist_residual_norm2)
        halo_exchange(top, top.x, rank, world)
        r = orc.residual(np.nan_to_num(top.x), top.f, 1.0 / (n - 1))
        ga, gb = max(top.y0, 1), min(top.y1, n - 1)
        mine = torch.tensor([float(np.sum(top.rows(r, ga, gb) ** 2))], dtype=torch.float64)
        allp = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(allp, mine)
        norms.append(float(np.sqrt(sum(float(p) for p in allp))))
    owned = top.rows(top.x, top.y0, top.y1)
    ok_bits = bool(np.array_equal(owned, want[top.y0:top.y1]))
    rref = orc.residual(want, f, 1.0 / (n - 1))
    ok_norm = abs(norms[-1] - orc.norm(rref)) <= 1e-12 * orc.norm(rref)
    q.put((rank, ok_bits, bool(ok_norm), int(np.isnan(owned).sum()), la))
    dist.barrier()
    dist.destroy_process_group()


def _run(world, n, agg_below, w_form=False, gamma=1, lo=2, port=29631):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, agg_below, w_form, gamma, lo, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    return sorted(res)


def test_window_operators_match_oracle():
    import cpu_checkers as cc
    orc = cc.load("orc")
    n, nc = 33, 17
    rng = np.random.default_rng(3)
    r = rng.standard_normal((n, n))
    assert np.array_equal(restrict_rows(r, 0, n, 0, nc), orc.restrict_fw(r))
    e = rng.standard_normal((nc, nc))
    for lo, mode in ((2, cc.PROLONG_REFERENCE), (1, cc.PROLONG_FULL)):
        x = rng.standard_normal((n, n))
        want = orc.prolong_add(x.copy(), e, mode)
        prolong_rows(x, 0, e, 0, n, lo)
        assert np.array_equal(x, want)


@pytest.mark.parametrize("world,n,agg_below,w_form,gamma,lo", [
    (2, 257, 33, False, 1, 2),   # three partitioned levels (257, 129, 65), reference prolongation
    (2, 129, 33, True, 2, 2),    # W-cycle: repeated visits re-exchange the iterate
    (2, 257, 65, False, 1, 1),   # full-interior prolongation
    (3, 513, 129, False, 1, 2),  # three ranks: a middle rank with two neighbours
])
def test_slab_schedule_is_bit_identical_to_single_process(world, n, agg_below, w_form, gamma, lo):
    res = _run(world, n, agg_below, w_form, gamma, lo, port=29631 + world + n % 97)
    assert len(res) == world
    for rank, ok_bits, ok_norm, nans, la in res:
        assert nans == 0, "rank %d: a halo is too shallow (poison reached an owned row)" % rank
        assert ok_bits, "rank %d: iterate differs from the single-process oracle" % rank
        assert ok_norm
        assert la >= 2
