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


def _mk_deck(order, rk, kind="iaw"):
    if kind == "iaw":
        return decks.plane_iaw(n=(16, 20), nv=(16, 12), order=order, rk=rk, A=0.05, ky1=1.0 / 234)
    if kind == "iaw_uneven":      # 18 = 9 + 9, 23 = 6 + 6 + 6 + 5: split_extent remainders (ParallelArray.C:642-660)
        return decks.plane_iaw(n=(18, 23), nv=(16, 12), order=order, rk=rk, A=0.05, ky1=1.0 / 234)
    if kind == "iaw_tiles":       # tiles of 64 x 16 per rank on a 2 x 2 grid: the pipelined kernel with its two-part launches
        return decks.plane_iaw(n=(128, 32), nv=(16, 8), order=order, rk=rk, A=0.05, ky1=1.0 / 234)
    if kind == "streams":         # three species, order 6 / RK6 (InterpenetratingStreams)
        return decks.interpenetrating_streams(n=(24, 30), nv=(12, 10))
    raise ValueError(kind)


def _run_rank(rank, world, port, px, py, order, rk, nsteps, dt, out, kind="iaw"):
    import torch
    import torch.distributed as dist
    from loki_b200 import decomp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if kind == "iaw_tiles":
        os.environ["LOKI_SPLIT_STAGES"] = "1"     # two-part stages are the default for a single species only
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        deck = _mk_deck(order, rk, kind)
        order = deck.order
        lay = decomp.TileLayout(deck.n, px, py, min_tile=order + 1)
        vp = decomp.DistributedVP(deck, lay, rank, dev, torch.cuda.current_stream().cuda_stream, dist if world > 1 else None)
        res = []
        for s, sp in enumerate(deck.species):
            f, fx, fv, fnorm = deck.initial_state(sp, vp.tile_lo, vp.tile_n)
            assert vp.H.lk_vp_set_state(vp.sys, s, f.ctypes.data) == 0
            assert deck.set_inflow(vp.H, vp.sys, s, vp.tile_lo, vp.tile_n) == 0
        assert vp.H.lk_vp_set_time(vp.sys, 0.3) == 0
        for _ in range(nsteps):
            vp.advance(dt)
        dts = vp.stable_dt()
        vp.synchronize()
        for s in range(vp.nsp):
            o = np.empty(tuple(reversed(vp.geoms[s].nd)))
            assert vp.H.lk_vp_get_state(vp.sys, s, o.ctypes.data) == 0
            res.append(o)
        out.put((rank, vp.tile_lo, vp.tile_n, res, dts, lk_pipe_count()))
        vp.close()
    finally:
        if world > 1:
            dist.destroy_process_group()


def lk_pipe_count():
    from loki_b200 import capi
    return capi.load().lk_pipe_launch_count()


def _launch(world, px, py, order, rk, nsteps, dt, kind="iaw"):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_run_rank, args=(r, world, port, px, py, order, rk, nsteps, dt, out, kind)) for r in range(world)]
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
        for rank, lo, n, res, dts, _ in multi:
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
    for rank, lo, n, res, dts, _ in multi:
        assert abs(dts - single[4]) <= 1e-9 * abs(single[4])
        assert dts == multi[0][4]


def _compare(single, multi, ng, tol):
    worst = 0.0
    for s in range(len(single[3])):
        ref = single[3][s]
        for rank, lo, n, res, dts, _ in multi:
            got = res[s][ng:-ng, ng:-ng, ng:ng + n[1], ng:ng + n[0]]
            want = ref[ng:-ng, ng:-ng, ng + lo[1]:ng + lo[1] + n[1], ng + lo[0]:ng + lo[0] + n[0]]
            assert np.any(want != 0.0)
            # relative to the cell's stencil neighbourhood along the two velocity directions (Maxwellian tails: a cell
            # many orders below its neighbours has no per-cell bound in either run)
            scale = np.abs(want).copy()
            for ax in (0, 1):
                for sh in (-1, 1):
                    scale = np.maximum(scale, np.abs(np.roll(want, sh, axis=ax)))
            err = float(np.max(np.abs(got - want) / np.where(scale > 0, scale, 1.0)))
            worst = max(worst, err)
            assert err <= tol, "species %d rank %d: %g" % (s, rank, err)
    for rank, lo, n, res, dts, _ in multi:
        assert abs(dts - single[4]) <= 1e-9 * abs(single[4]) and dts == multi[0][4]
    return worst


def _record(**kw):
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    try:
        os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
        with open(os.path.join(root, "gpurun_out", "multi_gpu_parity.jsonl"), "a") as fh:
            fh.write(json.dumps(kw) + "\n")
    except OSError:
        pass


GRID_CASES = [
    # (ranks, px, py, deck kind, order, rk)
    (4, 2, 2, "iaw_uneven", 4, 4),
    (4, 2, 2, "iaw_uneven", 6, 6),
    (4, 2, 2, "streams", 6, 6),
    (4, 2, 2, "iaw_tiles", 4, 4),
    (4, 4, 1, "iaw_tiles", 4, 4),
    (8, 2, 4, "iaw_uneven", 4, 4),
    (8, 2, 4, "streams", 6, 6),
    (8, 4, 2, "iaw_tiles", 4, 4),
]


@pytest.mark.parametrize("world,px,py,kind,order,rk", GRID_CASES)
def test_process_grids_match_single_rank(lk, world, px, py, kind, order, rk):
    """2 x 2 and 2 x 4 process grids (the layouts of grid_for) with uneven tiles (split_extent remainders), order 6 /
    RK6, three species, and tiles the pipelined kernel takes with its two-part launches: two steps equal the
    single-rank run (itself held to the oracle by the system tests) to the rounding of the velocity-space sums"""
    if _ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    nsteps, dt = 2, 0.02
    single = _launch(1, 1, 1, order, rk, nsteps, dt, kind)[0]
    multi = _launch(world, px, py, order, rk, nsteps, dt, kind)
    deck = _mk_deck(order, rk, kind)
    worst = _compare(single, multi, deck.ng, 1e-12)
    pipe = [m[5] for m in multi]
    if kind == "iaw_tiles":
        # every stage kernel of every species in two launches on every rank
        assert all(p == 2 * 2 * 4 * nsteps for p in pipe), pipe
    _record(world=world, grid=[px, py], deck=kind, order=deck.order, rk=deck.rk, species=len(deck.species), steps=nsteps,
            worst_rel_diff=worst, pipelined_launches_per_rank=pipe)
