"""World-size-2 / 4 CPU test over gloo of the multi-GPU F-CYCLE schedule (`cycle_f_dist` in csrc/solver.cu):

  (1) phi restricted level by level on row slabs (halo exchange, slab restriction), all-gathered at the first
      agglomerated level, then restricted redundantly on every rank down to 5x5;
  (2) nested iteration upwards: whole levels redundantly; slab levels with the FMG smoothing as passes of <= 2
      sweeps after an 8-row exchange each (2 halo rows kept valid), the analytic right-hand side, the slab
      prolongation reading one coarse halo row, and the V-cycle of `cycle_dist` started at that level.

Executed with the CPU oracle's operators on each rank's window, rows a step does not compute poisoned with NaN.
The gathered result must be BIT-IDENTICAL to the single-process oracle F-cycle (MultiGridTestRunner.hpp:192-205 +
MultiGrid.hpp:138-183), which is the check tests/dist_check.py makes on real GPUs.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from test_dist_gloo import PADY, Slab, cycle_dist, halo_exchange, prolong_rows, restrict_rows  # noqa: E402

FMG_EXT = 2  # rows beyond the slab the FMG smoothing keeps valid (csrc/solver.cu: cycle_f_dist)


def smooth_window(orc, arr, f, h, omega, sweeps):
    """`sweeps` Jacobi sweeps on a window; NaN poison travels as 1e300 through the oracle and back"""
    xs = np.nan_to_num(arr, nan=1e300)
    orc.jacobi(xs, f, h, omega=omega, num_iter=sweeps - 1)
    xs[np.abs(xs) > 1e200] = np.nan
    return xs


def fcycle_dist(orc, slabs, wholes, la, rank, world, omega, lo, fmg_sweeps=4):
    """slabs[l] for l < la (+ slabs[la] = the slab-shaped window of level la), wholes[l] = (x, xb, f) for l >= la"""
    import cpu_checkers as cc
    sizes = [s.n for s in slabs[:la]] + [w[0].shape[0] for w in wholes]
    lc = len(sizes) - 1
    mode = cc.PROLONG_FULL if lo == 1 else cc.PROLONG_REFERENCE
    # (1) restrict phi down
    for l in range(lc):
        if l < la:
            L, K = slabs[l], slabs[l + 1]
            fine = L.x if l == 0 else L.xb
            halo_exchange(L, fine, rank, world)
            out = restrict_rows(fine, L.w0, L.n, K.y0, K.y1)
            if l + 1 == la:
                parts = [None] * world
                rows0 = (K.n - 1) // world  # equally sized slabs; the last rank's extra row is the zero ring
                dist.all_gather_object(parts, out[:rows0].copy())
                full = np.zeros((K.n, K.n))
                full[: K.n - 1] = np.concatenate(parts)
                wholes[0][1][:] = full
            else:
                K.xb[:] = np.nan
                K.rows(K.xb, K.y0, K.y1)[:] = out
        else:
            wholes[l + 1 - la][1][:] = orc.restrict_fw(wholes[l - la][1])
    wholes[-1][0][:] = wholes[-1][1]
    # (2) nested iteration
    wholes[-1][2][:] = orc.rhs(sizes[lc])
    for l in range(lc - 1, -1, -1):
        nK, hK = sizes[l + 1], 1.0 / (sizes[l + 1] - 1)
        if l + 1 >= la:
            xK, _, fK = wholes[l + 1 - la]
            orc.jacobi(xK, fK, hK, omega=omega, num_iter=fmg_sweeps - 1)
        else:
            K = slabs[l + 1]
            left = fmg_sweeps
            cur = K.x
            while True:
                halo_exchange(K, cur, rank, world)
                b = min(2, left)
                if b > 0:
                    cur = smooth_window(orc, cur, K.f, hK, omega, b)
                    K.poison_outside(cur, FMG_EXT, rank, world)
                left -= b
                if left <= 0:
                    break
            K.x[:] = cur
        if l < la:
            L = slabs[l]
            n = L.n
            L.f[:] = orc.rhs(n)[L.w0:L.w1]
            L.x[:] = 0.0
            if l + 1 == la:
                W = slabs[la]
                W.x[:] = np.nan
                a, b = max(0, W.y0 - 4), min(W.n, W.y1 + 4)
                W.rows(W.x, a, b)[:] = wholes[0][0][a:b]
                e, c0 = W.x, W.w0
            else:
                e, c0 = slabs[l + 1].x, slabs[l + 1].w0
            own = L.rows(L.x, L.y0, L.y1)
            prolong_rows(own, L.y0, e, c0, n, lo)
            cycle_dist(orc, slabs, l, la, rank, world, omega, 1, False, False, lo)
        else:
            xL, _, fL = wholes[l - la]
            fL[:] = orc.rhs(sizes[l])
            xL[:] = 0.0
            orc.prolong_add(xL, wholes[l + 1 - la][0], mode)
            orc.cycle(xL, fL, kind=cc.V, omega=omega, eps=0.0, prolong=mode)


def _worker(rank, world, port, n, agg_below, lo, fmg_sweeps, q):
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
    wholes = [(np.zeros((m, m)), np.zeros((m, m)), np.zeros((m, m))) for m in sizes[la:]]
    rng = np.random.default_rng(11)
    phi = rng.standard_normal((n, n))  # a non-trivial start (ring included): exercises the restriction of phi
    f = orc.rhs(n)
    top = slabs[0]
    want = phi.copy()
    top.x[:] = phi[top.w0:top.w1]
    top.f[:] = f[top.w0:top.w1]
    mode = cc.PROLONG_FULL if lo == 1 else cc.PROLONG_REFERENCE
    ok_bits, nans = True, 0
    for _ in range(2):  # the second pass starts from the first one's result
        fcycle_dist(orc, slabs, wholes, la, rank, world, omega, lo, fmg_sweeps)
        if fmg_sweeps == 4:
            orc.cycle(want, f, kind=cc.F, omega=omega, eps=0.0, prolong=mode)
        else:
            want = _fcycle_single(orc, want, omega, mode, fmg_sweeps)
        owned = top.rows(top.x, top.y0, top.y1)
        ok_bits = ok_bits and bool(np.array_equal(owned, want[top.y0:top.y1]))
        nans += int(np.isnan(owned).sum())
    q.put((rank, ok_bits, nans, la))
    dist.barrier()
    dist.destroy_process_group()


def _fcycle_single(orc, phi, omega, mode, fmg_sweeps):
    """the oracle's F-cycle wrapper with a different FMG sweep count, from the oracle's operators"""
    import cpu_checkers as cc
    cur = phi
    while cur.shape[0] > 5:
        cur = orc.restrict_fw(cur)
    m = 5
    x, f = cur.copy(), orc.rhs(5)
    while m < phi.shape[0]:
        if fmg_sweeps > 0:
            orc.jacobi(x, f, 1.0 / (m - 1), omega=omega, num_iter=fmg_sweeps - 1)
        m = 2 * m - 1
        xf, f = np.zeros((m, m)), orc.rhs(m)
        orc.prolong_add(xf, x, mode)
        orc.cycle(xf, f, kind=cc.V, omega=omega, eps=0.0, prolong=mode)
        x = xf
    return x


def _run(world, n, agg_below, lo, fmg_sweeps, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, agg_below, lo, fmg_sweeps, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    return sorted(res)


@pytest.mark.parametrize("world,n,agg_below,lo,fmg_sweeps", [
    (2, 257, 33, 2, 4),   # three slab levels (257, 129, 65), reference prolongation, the reference's 4 FMG sweeps
    (2, 129, 33, 1, 3),   # full-interior prolongation, odd sweep count (copy-back of the last pass)
    (4, 513, 129, 2, 4),  # four ranks: middle ranks with two neighbours
])
def test_fcycle_slab_schedule_is_bit_identical(world, n, agg_below, lo, fmg_sweeps):
    res = _run(world, n, agg_below, lo, fmg_sweeps, port=29731 + world + n % 89)
    assert len(res) == world
    for rank, ok_bits, nans, la in res:
        assert nans == 0, "rank %d: a halo is too shallow (poison reached an owned row)" % rank
        assert ok_bits, "rank %d: F-cycle result differs from the single-process oracle" % rank
        assert la >= 2
