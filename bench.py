#!/usr/bin/env python
"""bench.py -- headline measurement of the B200-native Vlasov RHS path (BASELINE.json metric:
4D cell-updates/s per RK stage; % of HBM roofline).

  python bench.py --gpus N --steps K --warmup W            # own arm (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path on the host cores

Workload at N=1 (BASELINE.json configs[1]): the planeIAW deck's physics (two species, electron driver,
4th-order WENO, RK4) on 256^2 x 128^2 phase-space cells per species, synthetic (analytic) initial data.
At N>1 configuration space grows with N (weak scaling): every GPU keeps a 256 x 256 (x,y) tile and the
whole velocity space; per stage the ranks all-gather their charge-density tiles and exchange x/y face
halos over NCCL.  A "step" is one RK4 time step = 4 fused stage passes over both species.

The timed region keeps the state resident in HBM (`value`); `e2e` times the same step through the
C ABI with HOST buffers: upload of the state from pinned memory, the step, download of the new state.
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TILE = (256, 256)
NV = (128, 128)
ORDER, RK = 4, 4
# Process grids in (x,y).  The weak-scaling workload cuts y only: x is the contiguous axis, so y faces are long runs that
# pack at full bandwidth, a rank has two neighbours instead of four, the x wrap stays inside the stage kernel, and the
# halo volume is half that of a 2D grid (measured on one 4-GPU box: 1x4 359 G against 2x2 341 G, profiles/r2_multi_gpu.md).
# configs[4] (streams) is a fixed 512 x 512 global box: 2x4 tiles of 256 x 128.  --grid overrides.
GRIDS = {1: (1, 1), 2: (1, 2), 4: (1, 4), 8: (1, 8)}
STREAMS_GRID = {8: (2, 4)}
METRIC = "4D cell-updates/s per RK stage"
UNIT = "cell-updates/s"
# algorithmic bytes per cell-update of the fused RK4 stage kernel (SURVEY 8d / DESIGN.md):
# read f_eval, f_old (+ delta) ; write (delta +) pred -> 32 B in stage 1, 40 B in stages 2,3, 32 B in stage 4
STAGE_BYTES = (32, 40, 40, 32)
# fp64 instructions per cell-update of the stage kernel, from the ncu instruction counts under profiles/
# (order 4: 3.70 G warp-instructions x 32 lanes / 1.074 G cells; order 6: 7.4 G x 32 / 1.074 G)
FP64_INSTR_PER_CELL = {4: 110.3, 6: 221.0}


def grid_of(n_gpus, workload):
    if workload == "streams" and n_gpus in STREAMS_GRID and not GRID_OVERRIDE:
        return STREAMS_GRID[n_gpus]
    return GRIDS[n_gpus]


GRID_OVERRIDE = []


def config_dict(n_gpus, workload="iaw"):
    px, py = grid_of(n_gpus, workload)
    if workload == "iaw6":
        d = config_dict(n_gpus, "iaw")
        d["workload"] = d["workload"].replace("order 4 WENO", "order 6 WENO")
        d["l2"] = "inputs larger than L2 (9.9 GB per array)"
        return d
    if workload == "streams":
        return {"workload": "InterpenetratingStreams electrons, 1 species x 256x128x256x256 cells per GPU (global 512x512 "
                            "x 256x256 on 8 GPUs), order 6 WENO, RK4 (SURVEY 8d S5 variant)",
                "cells_per_species_per_gpu": 256 * 128 * 256 * 256, "species": 1, "stages_per_step": 4,
                "decomposition": "%dx%d tiles in (x,y)" % (px, py), "arithmetic": "production (strict available)",
                "l2": "inputs larger than L2 (19.3 GB per array)"}
    return {"workload": "planeIAW physics, 2 species x %dx%dx%dx%d cells per GPU (global %dx%d), order 4 WENO, RK4"
                        % (TILE[0], TILE[1], NV[0], NV[1], TILE[0] * px, TILE[1] * py),
            "cells_per_species_per_gpu": TILE[0] * TILE[1] * NV[0] * NV[1], "species": 2, "stages_per_step": 4,
            "decomposition": "%dx%d tiles in (x,y)" % (px, py), "arithmetic": "production (strict available)",
            "l2": "inputs larger than L2 (9.4 GB per array)"}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's algorithm on the host cores
# ------------------------------------------------------------------------------------------------
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libloki_ref.so")


def _cpu_worker(args):
    n, order, reps, use_ref = args
    os.environ["OMP_NUM_THREADS"] = "1"       # one process per core; the oracle's own OpenMP stays off
    if use_ref:
        L = C.CDLL(REF_SO)
        L.loki_ref_time_rk4_stage.restype = C.c_double
        return L.loki_ref_time_rk4_stage((C.c_int * 4)(*n), order, (C.c_double * 4)(0.07, 0.07, 0.1, 0.1), reps)
    sys.path.insert(0, os.path.join(ROOT, "tests"))  # the oracle binding is test infrastructure
    import oracle_binding
    ok = oracle_binding.load()
    g = oracle_binding.OkGeom.make(n, order, (0.07, 0.07, 0.1, 0.1))
    ok.ok_time_rk4_stage_reference_style.restype = C.c_double
    return ok.ok_time_rk4_stage_reference_style(C.byref(g), 1, reps)


def cpu_leg(reps=4, box=(32, 32, 64, 64), cores=None):
    """every core owns an independent periodic sub-box (the way the reference's MPI ranks each own a
    ParallelArray block) and runs `reps` RK4 stages done the reference's way (separate sweeps,
    materialised vel3/vel4, unfused zero/copy/axpy, separate velocity reduction).  With oracle/_ref present
    (built where /root/reference is mounted; it travels with the repository snapshot) the sweeps are the
    reference's own Fortran kernels, transliterated to C: kind "reference"; otherwise the hand-written C
    restatement: kind "port"."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    cells = box[0] * box[1] * box[2] * box[3]
    use_ref = os.path.exists(REF_SO)
    ctx = mp.get_context("fork")
    t0 = time.time()
    with ctx.Pool(cores) as pool:
        secs = pool.map(_cpu_worker, [(box, ORDER, reps, use_ref)] * cores)
    wall = time.time() - t0
    per_stage = max(secs)                      # slowest rank, like an MPI step
    value = cells * cores / per_stage
    what = ("the reference's own Fortran kernels (KineticSpeciesF.f transliterated to C by oracle/f77toc.py, gcc -O2 "
            "-ffp-contract=off) in the reference's unfused sequence" if use_ref else
            "oracle C restatement of the reference's unfused passes (gcc -O2 -ffp-contract=off)")
    return {"value": value, "unit": UNIT, "cores": cores, "kind": "reference" if use_ref else "port",
            "sample": "%d cores x independent %dx%dx%dx%d periodic sub-box x %d RK4 stages, %s; %.1f s wall"
                      % (cores, *box, reps, what, wall),
            "_per_stage_s": per_stage}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    t0 = time.time()
    vals = []
    for _ in range(args.warmup):
        cpu_leg(reps=1)
    per = []
    for _ in range(args.steps):
        r = cpu_leg(reps=8)
        vals.append(r["value"])
        per.append(r["_per_stage_s"])
    r.pop("_per_stage_s")
    value = sum(vals) / len(vals)
    r["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(per) / len(per),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(config_dict(args.gpus), reference_sample="the CPU arm times the same per-cell work on "
                           "independent 32x32x64x64 sub-boxes, one per host core (the MPI build's layout); see cpu_baseline.sample"),
            "cpu_baseline": r,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.time() - t0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm.sort()
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def run_own(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import loki_b200
    from loki_b200 import capi, decks, decomp, host

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    if args.gpus not in GRIDS:
        raise SystemExit("--gpus must be one of %s" % sorted(GRIDS))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = loki_b200.load()
    H = host.lib()
    if L.lk_device_count() < 1:
        raise SystemExit("no CUDA device: the hot path has no CPU fallback")
    L.lk_set_strict(0)
    def measure(workload, steps, warmup):
        """one workload: build the system, W warm-up steps, K timed steps (barrier + synchronize on both sides, CUDA
        events, MAX over ranks), live roofline of the stage kernel; returns the pieces of the JSON line"""
        px, py = grid_of(args.gpus, workload)
        tile = TILE if not args.small else (32, 32)
        nv = NV if not args.small else (32, 32)
        if workload == "streams":
            # BASELINE.json configs[4]: InterpenetratingStreams scaled to 512^2 x 256^2, electrons only, order 6 in
            # space with the fused RK4 (the variant SURVEY 8d recommends: 4 arrays of 19.3 GB per GPU)
            if args.gpus != 8 and not args.small:
                raise SystemExit("--workload streams is the 8-GPU configuration (one array is 137 GB globally)")
            tile = (256, 128) if not args.small else (32, 16)
            nv = (256, 256) if not args.small else (32, 32)
            deck = decks.interpenetrating_streams(n=(tile[0] * px, tile[1] * py), nv=nv, order=6, rk=4)
            deck.species = deck.species[:1]
        elif workload == "iaw6":
            # order-6 variant of the headline workload (BASELINE.json configs[2] physics at the headline size): planeIAW,
            # two species, 6th-order WENO, the fused RK4 update (RK6 would need 11 arrays of 9.9 GB per species)
            deck = decks.plane_iaw(n=(tile[0] * px, tile[1] * py), nv=nv, order=6, rk=4)
        else:
            deck = decks.plane_iaw(n=(tile[0] * px, tile[1] * py), nv=nv, order=ORDER, rk=RK)
        layout = decomp.TileLayout(deck.n, px, py, min_tile=deck.order + 1)
        stream = torch.cuda.current_stream().cuda_stream
        vp = decomp.DistributedVP(deck, layout, rank, dev, stream, dist if world > 1 else None)
        sys_ = vp.sys
        tile_lo = vp.tile_lo
        nsp = vp.nsp
        geoms = vp.geoms
        vols = [int(np.prod(g.nd)) for g in geoms]
        cells_rank = sum(int(np.prod([g.n[k] for k in range(4)])) for g in geoms)

        # ---- synthetic initial data: analytic Perturbed-Maxwellian tables, expanded on the device ----
        def device_state(s, amp):
            sp = deck.species[s]
            if sp.stream is not None:
                fx, fx2, fv, fv2 = deck.stream_tables(sp, tile_lo, tile)
                t = lambda a: torch.from_numpy(a).to(dev)
                f = (t(fv)[:, :, None, None] * t(fx)[None, None, :, :])
                f.addcmul_(t(fv2)[:, :, None, None], t(fx2)[None, None, :, :])
                assert deck.set_inflow(H, sys_, s, tile_lo, tile) == 0
                return f.contiguous()
            fx, fv, fnorm = deck.ic_tables(sp, tile_lo, tile)
            dfx = torch.from_numpy(fx).to(dev)
            dfv = torch.from_numpy(fv).to(dev)
            # a smooth spatial perturbation so that the WENO weights are exercised away from 1/2
            x = torch.arange(fx.shape[1], device=dev, dtype=torch.float64)
            y = torch.arange(fx.shape[0], device=dev, dtype=torch.float64)
            pert = 1.0 + amp * torch.cos(0.11 * x)[None, :] * torch.cos(0.07 * y)[:, None]
            f = ((fnorm * dfv)[:, :, None, None] * (dfx * pert)[None, None, :, :]).contiguous()
            H.lk_vp_set_inflow(sys_, s, fx.ctypes.data, fv.ctypes.data, fnorm, 1.0)
            return f

        for s in range(nsp):
            f = device_state(s, 0.05)
            assert f.numel() == vols[s]
            torch.cuda.synchronize()
            # device-to-device copy into the library-owned state array
            t_dst = _wrap(vp.state_ptr(s), vols[s], dev)
            t_dst.copy_(f.view(-1))
            del f
        torch.cuda.synchronize()

        step = vp.advance

        def barrier():
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        dt = 0.02
        # the clock sampler starts with the warm-up steps (same workload, same load): nvidia-smi needs about a
        # second to come up on an 8-GPU box, longer than a short timed region
        clocks = Clocks(local)
        clocks.start()
        for _ in range(warmup):
            step(dt)
        barrier()
        launches0 = L.lk_launch_count()
        pipe0 = L.lk_pipe_launch_count()
        L.lk_profile_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            step(dt)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        clk = clocks.stop()
        n_l, tot_ms = C.c_int64(), C.c_double()
        L.lk_profile_summary(C.byref(n_l), C.byref(tot_ms))
        L.lk_profile_enable(0)
        launches = L.lk_launch_count() - launches0
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
        stages = H.lk_vp_nstages(sys_)
        total_cells = cells_rank * world
        value = total_cells * stages * steps / (ms_max * 1e-3)

        # ---- roofline of the dominant kernel (fused stencil + RK update), live CUDA-event durations ----
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        cells_launch = cells_rank / nsp
        avg_bytes = cells_launch * (sum(STAGE_BYTES) / len(STAGE_BYTES))
        avg_ms = tot_ms.value / max(1, n_l.value)
        share = tot_ms.value / ms if ms > 0 else None
        split_note = None
        n_eval = steps * stages * nsp
        if n_l.value > n_eval:
            # a cut rank evaluates a stage in two launches that run CONCURRENTLY on two streams (the tiles on the cut
            # faces first, so that their halo exchange overlaps the rest): their event durations overlap and do not
            # add up to a stage.  The figure that cannot flatter: everything the device did in a stage evaluation.
            avg_ms = ms / n_eval
            share = None
            split_note = ("%d launches for %d stage evaluations (two-part stages, concurrent): avg_launch_ms is the whole device "
                          "time per stage evaluation, halo exchange and field solve included" % (n_l.value, n_eval))
        achieved = avg_bytes / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
        # dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel, averaged over the launches of
        # one step, from the committed ncu pass of this same command (tools/ncu_traffic.py writes the file)
        traffic, traffic_src = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
            if abs(tr["cells_per_launch"] - cells_launch) < 0.5:
                traffic, traffic_src = tr["dram_bytes_per_launch"], tr.get("source")
        except Exception:
            pass
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic,
                    "kernel": ("k_stage_pipe" if L.lk_pipe_launch_count() > pipe0 else "k_stage_march") +
                              " (fused WENO RHS + RK4 stage update + velocity moments)",
                    "algorithmic_bytes_per_launch": avg_bytes, "avg_launch_ms": avg_ms, "timed_launches": n_l.value,
                    "kernel_share_of_step": share, "peak_source": peak_src,
                    "traffic_source": traffic_src,
                    "note": "co-limited by fp64 issue: see roofline.fp64 (measured ceiling, tools/fp64_peak.cu) and DESIGN.md section 3"}
        if split_note:
            roofline["launches_note"] = split_note
        # the second ceiling: fp64 instructions per cell-update (ncu, profiles/r2_*) against the MEASURED fp64 issue
        # rate of this part under sustained load (tools/fp64_peak.cu -> profiles/r2_fp64_peak.json)
        try:
            fp = json.load(open(os.path.join(ROOT, "profiles", "r2_fp64_peak.json")))["fp64_peak"]
            ipc = FP64_INSTR_PER_CELL[deck.order]
            ach = ipc * cells_launch / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0
            roofline["fp64"] = {"peak": fp["sustained"], "unit": "T lane-instr/s", "achieved": ach, "frac": ach / fp["sustained"],
                                "instr_per_cell_update": ipc, "peak_source": "measured: " + fp["how"]}
        except Exception:
            pass
        return dict(vp=vp, deck=deck, value=value, ms_per_step=ms_max / steps, clocks=clk, launches=int(launches), roofline=roofline,
                    vols=vols, nsp=nsp, total_cells=total_cells, stages=stages, step=step, barrier=barrier, dt=dt, sys_=sys_)

    main = measure(args.workload, args.steps, args.warmup)
    vp, deck, value, clk, launches, roofline = main["vp"], main["deck"], main["value"], main["clocks"], main["launches"], main["roofline"]
    vols, nsp, total_cells, stages, step, barrier, dt, sys_ = (main[k] for k in ("vols", "nsp", "total_cells", "stages", "step", "barrier", "dt", "sys_"))
    ms_max = main["ms_per_step"] * args.steps

    # ---- e2e: the same step with HOST buffers (pinned), H2D of the state + step + D2H of the result ----
    e2e = None
    if not args.no_e2e:
        import psutil
        need = sum(vols) * 8
        pinned = psutil.virtual_memory().available > 3 * need * max(1, world)
        try:
            bufs = [torch.empty(v, dtype=torch.float64, pin_memory=pinned) for v in vols]
        except RuntimeError:
            # page-locking that much host memory can fail on a loaded box: fall back to pageable buffers
            pinned = False
            bufs = [torch.empty(v, dtype=torch.float64) for v in vols]
        for s in range(nsp):
            H.lk_vp_get_state(sys_, s, bufs[s].data_ptr())

        # Every step takes its input from the host buffers and returns its result there.  The download of step k and
        # the upload of step k+1 are in flight together (PCIe is full duplex) through the library's streaming calls:
        # the three rotating arrays of a species are the double buffer.  Separate in / out buffers when they fit.
        outs = bufs
        # (every rank of the box pins its own pair at the same moment: leave a wide margin before asking for the second set)
        if pinned and psutil.virtual_memory().available > (2 if world == 1 else 4) * need * max(1, world):
            try:
                outs = [torch.empty(v, dtype=torch.float64, pin_memory=True) for v in vols]
            except RuntimeError:
                outs = bufs
        streamed = pinned and outs is not bufs
        s_up, s_down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        c_up, c_down = C.c_void_p(s_up.cuda_stream), C.c_void_p(s_down.cuda_stream)

        def e2e_blocking_step():
            for s in range(nsp):
                H.lk_vp_set_state(sys_, s, bufs[s].data_ptr())
            vp.invalidate_halos()
            step(dt)
            vp.synchronize()
            for s in range(nsp):
                H.lk_vp_get_state(sys_, s, bufs[s].data_ptr())

        def upload():
            for s in range(nsp):
                capi.check(H.lk_vp_upload_next(sys_, s, bufs[s].data_ptr(), c_up), "lk_vp_upload_next")
            capi.check(H.lk_vp_adopt_next(sys_), "lk_vp_adopt_next")
            vp.invalidate_halos()

        def e2e_streamed(nrep):
            upload()
            for k in range(nrep):
                step(dt)
                vp.synchronize()
                for s in range(nsp):
                    capi.check(H.lk_vp_download_state(sys_, s, outs[s].data_ptr(), c_down), "lk_vp_download_state")
                if k + 1 < nrep:
                    upload()
            torch.cuda.synchronize(dev)

        nrep = 10 if streamed else 2
        if streamed:
            e2e_streamed(2)
        else:
            e2e_blocking_step()
        barrier()
        t0 = time.perf_counter()
        if streamed:
            e2e_streamed(nrep)
        else:
            for _ in range(nrep):
                e2e_blocking_step()
        barrier()
        wall = time.perf_counter() - t0
        tt = torch.tensor([wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        wall = float(tt.item())
        e2e = {"value": total_cells * stages * nrep / wall, "unit": UNIT, "h2d_bytes_per_step": need * world,
               "d2h_bytes_per_step": need * world, "steps": nrep, "pinned": bool(pinned),
               "api": ("lk_vp_upload_next + lk_vp_adopt_next + lk_vp_advance + lk_vp_download_state: every step uploads its "
                       "input and downloads its result; the download of step k and the upload of step k+1 overlap "
                       "(include/loki_b200_host.h)") if streamed else
                      "lk_vp_set_state + lk_vp_advance + lk_vp_get_state (include/loki_b200_host.h)"}
        del bufs, outs

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_leg(reps=32)
        cpu.pop("_per_stage_s", None)

    # ---- secondary workload on the same box, attached to the line (not the headline): the order-6 kernel at N = 1,
    # BASELINE.json configs[4] (InterpenetratingStreams 512^2 x 256^2) at N = 8 ----
    secondary = None
    sec_name = {1: "iaw6", 8: "streams"}.get(args.gpus)
    if sec_name and not args.no_secondary and not args.small and args.workload == "iaw":
        vp.close()
        vp = None
        torch.cuda.empty_cache()
        sec = measure(sec_name, max(2, min(args.steps, 3)), 1)
        secondary = {"config": config_dict(args.gpus, sec_name), "value": sec["value"], "unit": UNIT, "ms_per_step": sec["ms_per_step"],
                     "steps": max(2, min(args.steps, 3)), "warmup": 1, "clocks": sec["clocks"], "gpu_launches": sec["launches"],
                     "roofline": sec["roofline"]}
        sec["vp"].close()

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(args.gpus, args.workload),
                "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline}
        if args.small:
            line["config"]["workload"] += " [--small: 32x32x32x32 tile, NOT the headline size]"
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if secondary is not None:
            line["config"]["secondary"] = secondary
        print(json.dumps(line))
    if vp is not None:
        vp.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def _wrap(ptr, count, dev):
    """view `count` doubles of library-owned device memory as a torch tensor (plumbing only)"""
    import torch

    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (int(ptr), False), "version": 3, "strides": None}
    return torch.as_tensor(h, device=dev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--small", action="store_true", help="tiny tile for plumbing checks (not a valid bench number)")
    ap.add_argument("--workload", default="iaw", choices=["iaw", "iaw6", "streams"],
                    help="iaw: the headline configuration (BASELINE.json configs[1]); iaw6: the same at order 6; "
                         "streams: configs[4], 8 GPUs only")
    ap.add_argument("--grid", default=None, help="process grid PXxPY in (x,y) instead of the default for this GPU count")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the secondary workload attached as config.secondary (N=1: iaw6; N=8: streams)")
    args = ap.parse_args()
    if args.grid:
        px, py = (int(v) for v in args.grid.lower().split("x"))
        if px * py != args.gpus:
            ap.error("--grid %s does not have %d tiles" % (args.grid, args.gpus))
        GRIDS[args.gpus] = (px, py)
        GRID_OVERRIDE.append(args.grid)
    if args.impl == "reference":
        return run_reference(args)
    return run_own(args)


if __name__ == "__main__":
    sys.exit(main())
