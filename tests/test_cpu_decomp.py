"""Host logic of the multi-GPU path on CPU: tile arithmetic and, with world_size 2 over gloo, the message
logic of the halo exchange and the rho all-gather (loki_b200.decomp).  On the box the pack/unpack
callables are the CUDA kernels lk_halo_pack / lk_halo_unpack; here numpy slicing stands in for them so
the neighbour / ordering logic runs on host arrays (tests/test_gpu_multi.py covers the CUDA side)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from loki_b200.decomp import HaloExchanger, TileLayout, _GroupDist, grid_for, split_extent


def test_split_extent_matches_parallel_array_rule():
    # ParallelArray.C:642-660: remainder cells go to the lowest-index tiles
    assert split_extent(10, 3) == [(0, 4), (4, 3), (7, 3)]
    assert split_extent(512, 4) == [(0, 128), (128, 128), (256, 128), (384, 128)]
    assert split_extent(7, 7) == [(k, 1) for k in range(7)]
    with pytest.raises(ValueError):
        split_extent(3, 4)


def test_layout_neighbours_are_periodic_and_cover_the_domain():
    lay = TileLayout((37, 22), 2, 4, min_tile=5)
    seen = np.zeros((22, 37), dtype=int)
    for r in range(lay.world):
        lx, ly, nx, ny = lay.tile(r)
        seen[ly:ly + ny, lx:lx + nx] += 1
        lo, hi = lay.neighbours(r, 0)
        assert lay.neighbours(hi, 0)[0] == r and lay.neighbours(lo, 0)[1] == r
        lo, hi = lay.neighbours(r, 1)
        assert lay.neighbours(hi, 1)[0] == r and lay.neighbours(lo, 1)[1] == r
    assert (seen == 1).all()
    assert lay.neighbours(0, 1) == (6, 2) and lay.neighbours(7, 0) == (6, 6)
    with pytest.raises(ValueError):
        TileLayout((8, 8), 1, 2, min_tile=5)  # below stencil_width (KineticSpecies.C:495-504)
    assert grid_for(8) == (1, 8) and grid_for(4) == (1, 4) and grid_for(1) == (1, 1)


def _global_array(nx, ny, nv3, nv4):
    i4, i3, i2, i1 = np.meshgrid(np.arange(nv4), np.arange(nv3), np.arange(ny), np.arange(nx), indexing="ij")
    return (i1 + 100.0 * i2 + 1e4 * i3 + 1e6 * i4).astype(np.float64)


def _worker(rank, world, port, px, py, nglobal, ng, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lay = TileLayout(nglobal, px, py, min_tile=ng)
        nv3, nv4 = 3, 2
        G = _global_array(nglobal[0], nglobal[1], nv3, nv4)
        lx, ly, nx, ny = lay.tile(rank)
        f = np.full((nv4, nv3, ny + 2 * ng, nx + 2 * ng), np.nan)
        f[:, :, ng:ng + ny, ng:ng + nx] = G[:, :, ly:ly + ny, lx:lx + nx]
        n1d, n2d = nx + 2 * ng, ny + 2 * ng

        # the same slabs as k_halo_x / k_halo_y: x messages carry interior rows only, y messages the full
        # x extent including the x ghosts filled just before
        def src(side, d):
            if d == 0:
                return f[:, :, ng:ng + ny, (nx if side else ng):(nx if side else ng) + ng]
            return f[:, :, (ny if side else ng):(ny if side else ng) + ng, :]

        def dst(side, d):
            if d == 0:
                return f[:, :, ng:ng + ny, (ng + nx if side else 0):(ng + nx if side else 0) + ng]
            return f[:, :, (ng + ny if side else 0):(ng + ny if side else 0) + ng, :]

        def pack(buf, side, d):
            buf.copy_(torch.from_numpy(np.ascontiguousarray(src(side, d)).ravel()))

        def unpack(buf, side, d):
            dst(side, d)[...] = buf.numpy().reshape(dst(side, d).shape)

        def local_fill(d):
            if d == 0:
                f[:, :, :, :ng] = f[:, :, :, nx:nx + ng]
                f[:, :, :, ng + nx:] = f[:, :, :, ng:2 * ng]
            else:
                f[:, :, :ng, :] = f[:, :, ny:ny + ng, :]
                f[:, :, ng + ny:, :] = f[:, :, ng:2 * ng, :]

        # neighbours in a cut direction may have different extents in the OTHER direction only when that
        # one is cut too with a remainder; messages are sized by the receiver's own tile, equal along a
        # process row/column by construction
        bufs = {0: [torch.empty(nv4 * nv3 * ny * ng, dtype=torch.float64) for _ in range(4)],
                1: [torch.empty(nv4 * nv3 * ng * n1d, dtype=torch.float64) for _ in range(4)]}
        # the product driver binds the halo messages to their own process group (own NCCL communicator and
        # stream on the GPU box, so the rho all-gather never queues behind them): same wrapper here
        halo_group = dist.new_group(ranks=list(range(world)))
        HaloExchanger(lay, rank, _GroupDist(dist, halo_group)).exchange(bufs, pack, unpack, local_fill)
        # expected: the global array with periodic wrap, this rank's window
        ix = (np.arange(lx - ng, lx + nx + ng)) % nglobal[0]
        iy = (np.arange(ly - ng, ly + ny + ng)) % nglobal[1]
        want = G[:, :, iy][:, :, :, ix]
        ok_halo = bool(np.array_equal(f, want))

        # rho tiles -> all ranks hold every tile in rank order (uniform tiles: one all_gather_into_tensor)
        rho = torch.from_numpy(np.ascontiguousarray(G[0, 0, ly:ly + ny, lx:lx + nx]).ravel().copy())
        cells = [lay.tile(r)[2] * lay.tile(r)[3] for r in range(world)]
        ok_rho = True
        if lay.uniform():
            gathered = torch.empty(sum(cells), dtype=torch.float64)
            dist.all_gather_into_tensor(gathered, rho)
            R = np.zeros((nglobal[1], nglobal[0]))
            off = 0
            for r in range(world):
                tx, ty, tnx, tny = lay.tile(r)
                R[ty:ty + tny, tx:tx + tnx] = gathered[off:off + tnx * tny].numpy().reshape(tny, tnx)
                off += tnx * tny
            ok_rho = bool(np.array_equal(R, G[0, 0]))
        out.put((rank, ok_halo, ok_rho))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("px,py,nglobal,ng", [(1, 2, (8, 12), 2), (2, 1, (12, 6), 3), (2, 1, (9, 6), 2)])
def test_halo_exchange_and_rho_gather_world2_gloo(px, py, nglobal, ng):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, px, py, nglobal, ng, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_halo, ok_rho in res:
        assert ok_halo, "rank %d: ghost cells differ from the periodic global array" % rank
        assert ok_rho, "rank %d: gathered rho tiles do not assemble to the global field" % rank


@pytest.mark.parametrize("px,py,nglobal,ng", [(2, 2, (9, 11), 2), (1, 4, (8, 21), 3)])
def test_halo_exchange_world4_gloo_uneven_tiles(px, py, nglobal, ng):
    """four ranks: a 2 x 2 grid with remainders in both directions (5+4, 6+5: the y messages span the x ghosts just
    received, which is what fills the corner cells) and the y-only grid bench.py uses (1 x 4, tiles 6+5+5+5)"""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 4, port, px, py, nglobal, ng, out)) for r in range(4)]
    for p in procs:
        p.start()
    res = [out.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1, 2, 3]
    for rank, ok_halo, ok_rho in res:
        assert ok_halo, "rank %d: ghost cells differ from the periodic global array" % rank


def _dt_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from loki_b200.decomp import allreduce_lambda_max
        # |a_x| peaks in rank 0's tile, |a_y| in rank 1's (two species)
        local = [[3.0, 0.2, 0.5, 0.1], [0.3, 2.0, 0.4, 0.7]][rank]
        out.put((rank, allreduce_lambda_max(torch, dist, local)))
    finally:
        dist.destroy_process_group()


def test_stable_dt_uses_componentwise_global_maxima_world2_gloo():
    """KineticSpecies::computeDt all-reduces MAX on every component of m_lambda_max before it forms the time step
    (KineticSpecies.C:650-656): sum_d max_r lambda_d, not max_r sum_d lambda_d.  With |a_x| largest in one tile
    and |a_y| in the other the minimum over per-rank time steps would be too large."""
    from loki_b200.decomp import compute_dt
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dt_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == res[1] == [3.0, 2.0, 0.5, 0.7]
    dx = (0.5, 0.5, 0.1, 0.1)
    lam_xy = (7.05, 7.05)
    dt_global = compute_dt(lam_xy + (3.0, 2.0), dx, 4)        # what the reference (and one rank) computes
    dt_rank0 = compute_dt(lam_xy + (3.0, 0.2), dx, 4)
    dt_rank1 = compute_dt(lam_xy + (0.3, 2.0), dx, 4)
    assert dt_global < min(dt_rank0, dt_rank1)
    # closed form: beta / (pi * sum_d lambda_d / dx_d)
    assert abs(dt_global - 2.6 / (np.pi * (2 * 7.05 / 0.5 + 3.0 / 0.1 + 2.0 / 0.1))) < 1e-15
