"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): configuration space cut over two
ranks (NCCL halo exchange of f, all-gather of the rho tiles) must reproduce the single-GPU step.
The stencil arithmetic is position-independent, so the only admissible difference is the rounding of
the velocity-space sums (their partial-sum partition depends on the tile shape): <= 1e-12 per cell with
checkTests' relative difference."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

from loki_b200 import decks

pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _mk_deck(order, rk):
    return decks.plane_iaw(n=(16, 20), nv=(16, 12), order=order, rk=rk, A=0.05, ky1=1.0 / 234)


def _run_rank(rank, world, port, px, py, order, rk, nsteps, dt, out):
    import torch
    import torch.distributed as dist
    from loki_b200 import decomp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        deck = _mk_deck(order, rk)
        lay = decomp.TileLayout(deck.n, px, py, min_tile=order + 1)
        vp = decomp.DistributedVP(deck, lay, rank, dev, torch.cuda.current_stream().cuda_stream, dist if world > 1 else None)
        res = []
        for s, sp in enumerate(deck.species):
            f, fx, fv, fnorm = deck.initial_state(sp, vp.tile_lo, vp.tile_n)
            assert vp.H.lk_vp_set_state(vp.sys, s, f.ctypes.data) == 0
            assert vp.H.lk_vp_set_inflow(vp.sys, s, fx.ctypes.data, fv.ctypes.data, fnorm, sp.frac) == 0
        assert vp.H.lk_vp_set_time(vp.sys, 0.3) == 0
        for _ in range(nsteps):
            vp.advance(dt)
        dts = vp.stable_dt()
        vp.synchronize()
        for s in range(vp.nsp):
            o = np.empty(tuple(reversed(vp.geoms[s].nd)))
            assert vp.H.lk_vp_get_state(vp.sys, s, o.ctypes.data) == 0
            res.append(o)
        out.put((rank, vp.tile_lo, vp.tile_n, res, dts))
        vp.close()
    finally:
        if world > 1:
            dist.destroy_process_group()


def _launch(world, px, py, order, rk, nsteps, dt):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_run_rank, args=(r, world, port, px, py, order, rk, nsteps, dt, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res, key=lambda r: r[0])


@pytest.mark.skipif(_ngpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("px,py", [(1, 2), (2, 1)])
@pytest.mark.parametrize("order,rk", [(4, 4), (6, 6)])
def test_two_rank_step_matches_single_rank(lk, px, py, order, rk):
    nsteps, dt = 2, 0.02
    single = _launch(1, 1, 1, order, rk, nsteps, dt)[0]
    multi = _launch(2, px, py, order, rk, nsteps, dt)
    ng = 2 if order == 4 else 3
    for s in range(len(single[3])):
        ref = single[3][s]
        for rank, lo, n, res, dts in multi:
            got = res[s][ng:-ng, ng:-ng, ng:ng + n[1], ng:ng + n[0]]
            want = ref[ng:-ng, ng:-ng, ng + lo[1]:ng + lo[1] + n[1], ng + lo[0]:ng + lo[0] + n[0]]
            assert np.any(want != 0.0)
            den = np.where(want != 0.0, np.abs(want), 1.0)
            err = np.max(np.abs(got - want) / den)
            assert err <= 1e-12, "species %d rank %d: %g" % (s, rank, err)
    # the stable time step is a global minimum: every rank reports the same value, equal to the single-rank
    # one up to the rounding of the net charge density (electron and ion densities cancel to ~1e-5 of
    # their size, and the partial-sum partition of the velocity integrals depends on the tile shape, so
    # the field -- and max|a| with it -- moves at the 1e-11 level; the reference's MPI_Reduce has the same
    # sensitivity to the rank count)
    for rank, lo, n, res, dts in multi:
        assert abs(dts - single[4]) <= 1e-9 * abs(single[4])
        assert dts == multi[0][4]
