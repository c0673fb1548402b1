"""Configuration-space decomposition over the GPUs of one box and the per-stage exchanges.

The reference block-decomposes all four dimensions over MPI ranks (ParallelArray.C:545-569) and
exchanges ghost slabs with packed MPI messages (ParallelArray.C:925-1113); the charge density is an
MPI_Reduce over the ranks sharing an (x,y) tile (ReductionSchedule.C:69-113) and every Poisson rank
solves the same global 2D problem (EMSolverBase.C:301-345).  Here only (x,y) is cut, the
velocity space stays whole on each GPU, so per stage a rank

  1. all-gathers its rho tile (2D, a few hundred KB),
  2. swaps `ng` layers of f with its x neighbours, then with its y neighbours (periodic wrap is an
     ordinary neighbour; axis-aligned stencils need no corner messages),

one process per GPU over torch.distributed (NCCL on the box, gloo in the CPU tests).  The exchange of a
species' new predictor is issued on a second CUDA stream and its own NCCL communicator as soon as that
species' stage kernel has been queued, so it travels over NVLink while the next species' stage kernel (and
the small moment / field work of the next stage) runs; a stage kernel only waits for its own species'
halos.  This module is the host logic only: tile arithmetic, neighbour ranks, message ordering.  Packing, unpacking and everything
else that touches 4D data are CUDA kernels behind the C ABI (lk_halo_pack / lk_halo_unpack); the
exchanger takes them as callables so the CPU tests can drive the same message logic on host arrays.
"""
import ctypes as C


def split_extent(n, parts):
    """Sub-box rule of ParallelArray.C:642-660: n // parts cells each, the remainder one extra cell to the
    lowest-index tiles.  Returns [(lo, count)]."""
    if parts < 1 or n < parts:
        raise ValueError("cannot cut %d cells into %d tiles" % (n, parts))
    base, extra = divmod(n, parts)
    out, lo = [], 0
    for k in range(parts):
        cnt = base + (1 if k < extra else 0)
        out.append((lo, cnt))
        lo += cnt
    return out


class TileLayout:
    """px x py process grid over a global Nx x Ny configuration space; rank = ry * px + rx."""

    def __init__(self, nglobal, px, py, min_tile=1):
        self.nglobal, self.px, self.py = (int(nglobal[0]), int(nglobal[1])), int(px), int(py)
        self.xs = split_extent(self.nglobal[0], self.px)
        self.ys = split_extent(self.nglobal[1], self.py)
        for lo, cnt in self.xs + self.ys:
            if cnt < min_tile:
                # stencil_width = order + 1 is the minimum interior extent per tile (KineticSpecies.C:495-504)
                raise ValueError("tile extent %d is below the minimum %d" % (cnt, min_tile))

    @property
    def world(self):
        return self.px * self.py

    def coords(self, rank):
        return rank % self.px, rank // self.px

    def rank_of(self, rx, ry):
        return (ry % self.py) * self.px + (rx % self.px)

    def tile(self, rank):
        """(lo_x, lo_y, n_x, n_y) of the rank's interior"""
        rx, ry = self.coords(rank)
        return self.xs[rx][0], self.ys[ry][0], self.xs[rx][1], self.ys[ry][1]

    def reference_rank(self, rank):
        """the rank ParallelArray gives this tile: dimension 0 varies slowest there (getIndexRank, ParallelArray.C:609-626),
        fastest here; a restart dump names a tile's dataset by it (RestartWriter.C:556-559)"""
        rx, ry = self.coords(rank)
        return rx * self.py + ry

    def restart_tiles(self, tiles_by_rank, n_ghosts, nv):
        """({reference rank: tile}, distribInfo) for outputs.RestartWriter.write_parallel_array / outputs.put_species from
        every rank's dataBox (lk_vp_get_state of each rank, gathered by the caller): the split is ParallelArray's
        (split_extent), only the numbering of the ranks differs"""
        from . import outputs
        info = outputs.distrib_info(0, self.world - 1, n_ghosts, [self.nglobal[0], self.nglobal[1], nv[0], nv[1]],
                                    [self.px, self.py, 1, 1])
        return {self.reference_rank(r): t for r, t in tiles_by_rank.items()}, info

    def tiles_flat(self):
        out = []
        for r in range(self.world):
            out += list(self.tile(r))
        return out

    def neighbours(self, rank, direction):
        """(low, high) neighbour ranks along x (0) or y (1), periodic"""
        rx, ry = self.coords(rank)
        if direction == 0:
            return self.rank_of(rx - 1, ry), self.rank_of(rx + 1, ry)
        return self.rank_of(rx, ry - 1), self.rank_of(rx, ry + 1)

    def cut(self, direction):
        return (self.px if direction == 0 else self.py) > 1

    def uniform(self):
        return len({c for _, c in self.xs}) == 1 and len({c for _, c in self.ys}) == 1


def grid_for(world):
    """Process grids used on one box for a free-form (weak-scaling) domain: cut y only.  x is the contiguous axis, so y
    faces are long runs that pack at full bandwidth, a rank has two neighbours instead of four, the x wrap stays inside
    the stage kernel and the halo volume is half that of a 2D grid (measured: profiles/r2_multi_gpu.md).  Any px x py
    TileLayout works (tests/test_gpu_multi.py runs 2x2, 4x1, 2x4, 4x2)."""
    return {1: (1, 1), 2: (1, 2), 4: (1, 4), 8: (1, 8)}[world]


class HaloExchanger:
    """x-then-y face exchange of one 4D array.

    pack(buf, side, direction)   -- fill `buf` with the `ng` interior layers next to face `side` (0 low, 1 high)
    unpack(buf, side, direction) -- write `buf` into the ghost layers of face `side`
    local_fill(direction)        -- periodic wrap inside the array (that direction is not cut)
    bufs[direction] = (send_lo, send_hi, recv_lo, recv_hi) tensors on the communication device.
    The y exchange runs after the x exchange has been unpacked and its messages span the x ghosts, which
    is how the edge/corner cells get their values (ParallelArray.H:580-606 orders its periodic copies the
    same way).
    """

    def __init__(self, layout, rank, dist):
        self.layout, self.rank, self.dist = layout, rank, dist

    def exchange(self, bufs, pack, unpack, local_fill):
        dist = self.dist
        for d in (0, 1):
            if not self.layout.cut(d):
                local_fill(d)
                continue
            slo, shi, rlo, rhi = bufs[d]
            lo_n, hi_n = self.layout.neighbours(self.rank, d)
            pack(slo, 0, d)
            pack(shi, 1, d)
            ops = [dist.P2POp(dist.isend, slo, lo_n, tag=2 * d), dist.P2POp(dist.isend, shi, hi_n, tag=2 * d + 1),
                   # my high ghosts are my high neighbour's low interior layers (its send_lo) and vice versa
                   dist.P2POp(dist.irecv, rhi, hi_n, tag=2 * d), dist.P2POp(dist.irecv, rlo, lo_n, tag=2 * d + 1)]
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            unpack(rlo, 0, d)
            unpack(rhi, 1, d)


class _GroupDist:
    """torch.distributed with every point-to-point op bound to one process group"""

    def __init__(self, dist, group):
        self._d, self._g = dist, group
        self.isend, self.irecv = dist.isend, dist.irecv

    def P2POp(self, op, tensor, peer, tag=0):
        if self._g is None:
            return self._d.P2POp(op, tensor, peer, tag=tag)
        return self._d.P2POp(op, tensor, peer, group=self._g, tag=tag)

    def batch_isend_irecv(self, ops):
        return self._d.batch_isend_irecv(ops)


class DistributedVP:
    """The stage loop of RK4Integrator / RK6Integrator (RK4Integrator.H:66-171) around the C++ host
    mirror when configuration space is cut over several ranks: lk_vp_* stage pieces with the two
    exchanges between them (include/loki_b200_host.h documents the protocol)."""

    def __init__(self, deck, layout, rank, device, stream, dist=None):
        import torch
        from . import capi, host
        self.torch, self.dist = torch, dist
        self.L, self.H = capi.load(), host.lib()
        self.deck, self.layout, self.rank, self.device = deck, layout, rank, device
        self.stream = C.c_void_p(stream)
        lo_x, lo_y, n_x, n_y = layout.tile(rank)
        self.tile_lo, self.tile_n = (lo_x, lo_y), (n_x, n_y)
        self.desc = deck.product_desc(tile_lo=self.tile_lo, tile_n=self.tile_n, ntiles=layout.world)
        self.sys = C.c_void_p()
        capi.check(self.H.lk_vp_create(C.byref(self.sys), C.byref(self.desc), self.stream), "lk_vp_create")
        self.nsp = len(deck.species)
        self.geoms = []
        for s in range(self.nsp):
            g = capi.Geom()
            self.H.lk_vp_species_geom(self.sys, s, C.byref(g))
            self.geoms.append(g)
        self.world = layout.world
        # Two-part stages (lk_vp_stage_finish_species_part) hide a species' exchange behind its OWN kernel.  With
        # several species the exchange already travels under the next species' kernel and the split only costs (two
        # launches, two tails: measured 96-108 ms against 91 ms per step at 4 GPUs, profiles/r2_multi_gpu.md), so it is
        # the default for a single species only; LOKI_SPLIT_STAGES=0/1 overrides.
        import os
        env = os.environ.get("LOKI_SPLIT_STAGES", "")
        self.split_stages = (self.nsp == 1) if env not in ("0", "1") else (env == "1")
        if self.world > 1:
            if dist is None:
                raise ValueError("a cut layout needs torch.distributed")
            flat = layout.tiles_flat()
            self.tiles_arr = (C.c_int * len(flat))(*flat)
            cells = [flat[4 * r + 2] * flat[4 * r + 3] for r in range(self.world)]
            self.tile_cells, self.max_cells = cells, max(cells)
            f64 = torch.float64
            self.rho_tile = torch.zeros(self.max_cells, dtype=f64, device=device)
            self.rho_gather = torch.zeros(sum(cells), dtype=f64, device=device)
            self.rho_padded = None if layout.uniform() else torch.zeros(self.world * self.max_cells, dtype=f64, device=device)
            capi.check(self.H.lk_vp_set_comm_buffers(self.sys, self.rho_tile.data_ptr(), self.rho_gather.data_ptr()),
                       "lk_vp_set_comm_buffers")
            # halo traffic gets its own communicator: c10d gives every process group one NCCL stream, and the
            # (tiny) rho all-gather of the next stage must not queue behind a species' face messages
            # (high-priority streams: the stage kernel fills every SM, so the small pack / NCCL / unpack kernels
            # of the exchange must win the CTA slots that free up instead of queueing behind its 4096 CTAs)
            self.halo_group = None
            if hasattr(dist, "new_group"):
                opts = None
                try:
                    if dist.get_backend() == "nccl":
                        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
                except Exception:
                    opts = None
                self.halo_group = dist.new_group(ranks=list(range(self.world)), pg_options=opts) if opts is not None \
                    else dist.new_group(ranks=list(range(self.world)))
            self.exchanger = HaloExchanger(layout, rank, _GroupDist(dist, self.halo_group))
            on_gpu = device is not None and str(device).startswith("cuda")
            self.main_stream = torch.cuda.current_stream(device) if on_gpu else None
            self.comm_stream = torch.cuda.Stream(device=device, priority=-1) if on_gpu else None
            self.ev_stage = [torch.cuda.Event() for _ in range(self.nsp)] if on_gpu else None
            self.ev_halo = [torch.cuda.Event() for _ in range(self.nsp)] if on_gpu else None
            self.halo_ready = [False] * self.nsp   # the evaluated state of species s has its x/y ghosts
            self.halo = []
            for s in range(self.nsp):
                bufs = {}
                for d in (0, 1):
                    cnt = self.L.lk_halo_count(C.byref(self.geoms[s]), d)
                    bufs[d] = [torch.empty(cnt, dtype=f64, device=device) for _ in range(4)]
                self.halo.append(bufs)

    def close(self):
        if self.sys:
            self.H.lk_vp_destroy(self.sys)
            self.sys = C.c_void_p()

    @property
    def nstages(self):
        return self.H.lk_vp_nstages(self.sys)

    def state_ptr(self, s):
        return self.H.lk_vp_state_ptr(self.sys, s)

    def _exchange_halos(self, s, stream):
        """x-then-y face exchange of species s's evaluated state; every kernel and message on `stream`"""
        L, H = self.L, self.H
        st = C.c_void_p(stream.cuda_stream) if stream is not None else self.stream
        f = self.H.lk_vp_eval_ptr(self.sys, s)
        g = C.byref(self.geoms[s])

        def local_fill(d):
            # a direction that is not cut wraps inside the rank; the fused stage kernel has normally written
            # those ghost cells already and the call is a no-op.  It runs on the system's own stream, so the
            # two streams are ordered around it (x before y, ParallelArray.H:580-606).
            if not H.lk_vp_local_fill_needed(self.sys, s, d):
                return
            if stream is not None:
                e = self.torch.cuda.Event()
                e.record(stream)
                self.main_stream.wait_event(e)
            H.lk_vp_local_fill(self.sys, s, d)
            if stream is not None:
                e = self.torch.cuda.Event()
                e.record(self.main_stream)
                stream.wait_event(e)

        self.exchanger.exchange(
            self.halo[s],
            lambda buf, side, d: L.lk_halo_pack(buf.data_ptr(), f, g, d, side, st),
            lambda buf, side, d: L.lk_halo_unpack(f, buf.data_ptr(), g, d, side, st),
            local_fill)

    def _start_exchange(self, s):
        """queue the exchange of species s's evaluated state behind everything the main stream holds now"""
        torch = self.torch
        if self.comm_stream is None:
            self._exchange_halos(s, None)
        else:
            # behind the face tiles of a two-part stage (their own stream), else behind the main stream
            from . import capi
            capi.check(self.H.lk_vp_wait_faces(self.sys, s, C.c_void_p(self.comm_stream.cuda_stream)), "lk_vp_wait_faces")
            with torch.cuda.stream(self.comm_stream):
                self._exchange_halos(s, self.comm_stream)
                self.ev_halo[s].record(self.comm_stream)
        self.halo_ready[s] = True

    def invalidate_halos(self):
        """call after writing a state from outside (lk_vp_set_state): its ghosts must be exchanged again"""
        if self.world > 1:
            self.halo_ready = [False] * self.nsp

    def _gather_rho(self):
        dist = self.dist
        if self.rho_padded is None:
            dist.all_gather_into_tensor(self.rho_gather, self.rho_tile)
        else:
            dist.all_gather_into_tensor(self.rho_padded, self.rho_tile)
            off = 0
            for r, cnt in enumerate(self.tile_cells):
                self.rho_gather[off:off + cnt].copy_(self.rho_padded[r * self.max_cells:r * self.max_cells + cnt])
                off += cnt

    def advance(self, dt):
        from . import capi
        H = self.H
        if self.world == 1:
            capi.check(H.lk_vp_advance(self.sys, dt), "lk_vp_advance")
            return
        capi.check(H.lk_vp_begin_step(self.sys, dt), "lk_vp_begin_step")
        for stage in range(self.nstages):
            capi.check(H.lk_vp_stage_moments(self.sys, stage), "lk_vp_stage_moments")
            self._gather_rho()
            capi.check(H.lk_vp_stage_field(self.sys, stage, self.tiles_arr), "lk_vp_stage_field")
            for s in range(self.nsp):
                if not self.halo_ready[s]:
                    self._start_exchange(s)       # first stage after a state upload
                if self.comm_stream is not None:
                    self.main_stream.wait_event(self.ev_halo[s])
                if self.split_stages:
                    # the stage kernel in two launches: the tiles on the cut faces first, then the new predictor's
                    # faces leave (second stream) under the launch of the remaining tiles
                    capi.check(H.lk_vp_stage_finish_species_part(self.sys, stage, s, 1), "lk_vp_stage_finish_species_part")
                    self._start_exchange(s)
                    capi.check(H.lk_vp_stage_finish_species_part(self.sys, stage, s, 2), "lk_vp_stage_finish_species_part")
                else:
                    capi.check(H.lk_vp_stage_finish_species(self.sys, stage, s), "lk_vp_stage_finish_species")
                    # the new predictor's faces leave now, under the next species' stage kernel
                    self._start_exchange(s)
        capi.check(H.lk_vp_end_step(self.sys), "lk_vp_end_step")

    def synchronize(self):
        if self.world > 1 and self.comm_stream is not None:
            self.comm_stream.synchronize()

    def stable_dt(self):
        """KineticSpecies::computeDt with configuration space cut over the ranks.  The reference all-reduces
        MAX over each component of m_lambda_max on the species communicator BEFORE it forms
        imLam = sum_d pi lambda_d / dx_d (KineticSpecies.C:650-656), i.e. dt comes from sum_d max_r lambda_d, which
        is >= max_r sum_d lambda_d: when |a_x| peaks in one tile and |a_y| in another, a MIN over per-rank time
        steps would be too large.  So: local (axmax, aymax) of every species -> MAX over ranks -> back into the
        host mirror -> lk_vp_stable_dt (identical on every rank, equal to the single-rank value)."""
        from . import capi
        if self.world > 1:
            pair = (C.c_double * 2)()
            vals = []
            for s in range(self.nsp):
                capi.check(self.H.lk_vp_lambda_max(self.sys, s, C.byref(pair)), "lk_vp_lambda_max")
                vals += [pair[0], pair[1]]
            glob = allreduce_lambda_max(self.torch, self.dist, vals, self.device)
            for s in range(self.nsp):
                pair[0], pair[1] = glob[2 * s], glob[2 * s + 1]
                capi.check(self.H.lk_vp_set_lambda_max(self.sys, s, C.byref(pair)), "lk_vp_set_lambda_max")
        dt = C.c_double()
        capi.check(self.H.lk_vp_stable_dt(self.sys, C.byref(dt)), "lk_vp_stable_dt")
        return dt.value


def allreduce_lambda_max(torch, dist, local_values, device=None):
    """component-wise MAX over the ranks (KineticSpecies.C:651-656: MPI_Allreduce(MPI_MAX) on m_lambda_max)"""
    t = torch.tensor(list(local_values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


def compute_dt(lambda_max, dx, rk_order):
    """KineticSpecies::computeDt without collision operators (KineticSpecies.C:647-694); host-side restatement
    of the formula the C++ mirror evaluates (loki_b200/csrc/lk_host.cu KineticSpecies::computeDt)"""
    import math
    pi = 4.0 * math.atan(1.0)
    im = 0.0
    for d in range(4):
        im += pi * lambda_max[d] / dx[d]
    beta = 2.6 if rk_order == 4 else 3.168
    return math.sqrt(1.0 / (im * im / (beta * beta)))
