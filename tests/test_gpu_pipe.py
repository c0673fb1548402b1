"""The pipelined instantiation of the fused stage kernel (loki_b200/csrc/lk_pipe.cuh) against (a) the generic
marching kernel, bit for bit (same per-cell arithmetic by construction; this pins the synchronisation: a
missed hand-over shows up as a stale or torn value), and (b) the oracle's unfused RHS + RK4 stage update
(KineticSpeciesF.f:1949-2245, 10-38; RK4Integrator.H:149-171) within the north-star tolerance."""
import ctypes as C

import numpy as np
import pytest

from util import Setup, Dev, star_rel_err

pytestmark = pytest.mark.gpu


def chk(lk, status, what):
    assert status == 0, "%s: %s" % (what, lk.lk_last_error().decode())


def _stage(lk, d, s, stage, nmom, wrap, variant, seed=5, bcs=None, f=None, preset=False, tile_sets=None, cut=0):
    """one RK4-shaped fused stage; returns (pred, delta, moment partials, pipelined launches used)"""
    import torch
    import loki_b200 as lkm
    rng = np.random.default_rng(seed)
    f_old = d.t(s.f * (1.0 + 0.01 * rng.uniform(-1, 1, size=s.f.shape)))
    delta = d.t(0.001 * s.f * rng.uniform(-1, 1, size=s.f.shape))
    pred = torch.full_like(d.f, 3.0)
    u = lkm.RkUpdate()
    u.f_old, u.pred = f_old.data_ptr(), pred.data_ptr()
    u.delta_in = None if stage == 1 else delta.data_ptr()
    u.delta_out = None if stage == 4 else delta.data_ptr()
    u.w_delta, u.c_pred, u.use_delta = 0.0123, 0.05, int(stage == 4)
    u.wrap = wrap
    if bcs is not None:
        u.accel_bcs, u.inflow_preset = C.addressof(bcs), int(preset)
    f = d.f if f is None else f
    m, part = None, None
    if nmom:
        parts = lk.lk_stage_moment_parts(C.byref(d.g))
        part = torch.full((nmom * parts * s.n[0] * s.n[1],), float("nan"), dtype=torch.float64, device="cuda")
        m = lkm.StageMoments()
        m.nmom, m.partial, m.capacity = nmom, part.data_ptr(), part.numel()
    old = lk.lk_set_rhs_variant(variant)
    before = lk.lk_pipe_launch_count()
    for ts in (tile_sets or (0,)):
        u.tile_set, u.cut_dirs = ts, cut
        chk(lk, lk.lk_vlasov_stage(None, f.data_ptr(), C.byref(d.g), d.velocities.data_ptr(), C.byref(d.accel), C.byref(u),
                                   C.byref(m) if m else None, None), "stage")
    torch.cuda.synchronize()
    used = lk.lk_pipe_launch_count() - before
    lk.lk_set_rhs_variant(old)
    return pred, delta, part, used


SHAPES = [
    ((32, 8, 8, 5), 4),       # one tile, fewer planes than a chunk
    ((64, 16, 16, 20), 4),    # 2 x 2 x 2 tiles, several chunks
    ((32, 24, 40, 9), 4),     # supertile remainders (3 y-tiles, 5 vx-tiles)
    ((96, 8, 8, 33), 4),
    ((32, 8, 8, 7), 6),
    ((64, 16, 24, 18), 6),
]


@pytest.mark.parametrize("wrap", [0, 3])
@pytest.mark.parametrize("nmom", [0, 1, 3])
@pytest.mark.parametrize("stage", [1, 2, 4])
@pytest.mark.parametrize("n,order", SHAPES)
def test_pipe_equals_generic_kernel(lk, ok, fast, n, order, stage, nmom, wrap):
    import torch
    s = Setup(ok, n, order, rough=0.3)
    d = Dev(lk, s)
    chk(lk, lk.lk_periodic_fill_4d(d.f.data_ptr(), C.byref(d.g), 1, 1, None), "per")
    pa, da, ma, used_a = _stage(lk, d, s, stage, nmom, wrap, 0)
    pb, db, mb, used_b = _stage(lk, d, s, stage, nmom, wrap, 2)
    assert used_a == 1 and used_b == 0, "the aligned RK4-shaped stage must take the pipelined kernel"
    assert torch.equal(pa, pb)
    assert torch.equal(da, db)
    if nmom:
        assert torch.equal(ma, mb)


@pytest.mark.parametrize("n,order", [((64, 16, 16, 12), 4), ((32, 16, 16, 10), 6)])
def test_pipe_repeated_launches_are_deterministic(lk, ok, fast, n, order):
    """scheduling-independent: the hand-overs are data-race free, so 20 launches give one answer"""
    import torch
    s = Setup(ok, n, order, rough=0.3)
    d = Dev(lk, s)
    chk(lk, lk.lk_periodic_fill_4d(d.f.data_ptr(), C.byref(d.g), 1, 1, None), "per")
    ref = _stage(lk, d, s, 2, 3, 3, 2)
    for _ in range(20):
        got = _stage(lk, d, s, 2, 3, 3, 0)
        assert got[3] == 1
        assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1]) and torch.equal(got[2], ref[2])


@pytest.mark.parametrize("stage", [1, 2, 4])
@pytest.mark.parametrize("n,order", [((32, 16, 16, 12), 4), ((32, 8, 16, 9), 6)])
def test_pipe_against_oracle(lk, ok, fast, n, order, stage):
    """the pipelined stage against the oracle's unfused passes on the same inputs"""
    s = Setup(ok, n, order, rough=0.3)
    d = Dev(lk, s)
    # periodic ghosts on both sides
    ok.ok_periodic_fill_4d(s.f.ravel(), C.byref(s.g), 1, 1)
    d.f.copy_(d.t(s.f))
    pred, delta, _, used = _stage(lk, d, s, stage, 0, 0, 0)
    assert used == 1
    rng = np.random.default_rng(5)
    f_old = s.f * (1.0 + 0.01 * rng.uniform(-1, 1, size=s.f.shape))
    delta0 = 0.001 * s.f * rng.uniform(-1, 1, size=s.f.shape)
    rhs = np.zeros_like(s.f)
    vel3, vel4, _, _ = s.vel34(ok)
    ok.ok_advection_derivatives_4d(rhs.ravel(), s.f.ravel(), C.byref(s.g), s.vel1, s.vel2)
    ok.ok_acceleration_derivatives_4d(rhs.ravel(), s.f.ravel(), C.byref(s.g), vel3, vel4)
    w, c = 0.0123, 0.05
    ng = s.ng
    I = (slice(ng, -ng),) * 4
    dl = np.zeros_like(s.f)
    dl[I] = (w * rhs[I]) if stage == 1 else (delta0[I] + w * rhs[I])
    pr = f_old.copy()
    pr[I] = f_old[I] + c * (dl[I] if stage == 4 else rhs[I])
    got = pred.cpu().numpy()
    scale = np.maximum(np.abs(s.f), np.abs(f_old))
    assert star_rel_err(got, pr, scale, ng) <= 1e-12
    if stage != 4:
        gd = delta.cpu().numpy()
        assert star_rel_err(gd, dl, w * np.abs(rhs) + np.abs(delta0) + 1e-3 * scale, ng) <= 1e-12


def _inflow(d, s, kind):
    """the three table-driven inflow descriptions of lk_inflow (PerturbedMaxwellianIC / InterpenetratingStreamIC)"""
    import loki_b200 as lkm
    ic = lkm.Inflow()
    rng = np.random.default_rng(77)
    keep = [d.t(s.fx), d.t(s.fv), d.t(s.fx * rng.uniform(0.5, 1.5, size=s.fx.shape)), d.t(s.fv * rng.uniform(0.5, 1.5, size=s.fv.shape))]
    ic.kind, ic.fx, ic.fv, ic.fnorm, ic.frac = kind, keep[0].data_ptr(), keep[1].data_ptr(), 0.7, 0.9
    if kind in (2, 4):
        ic.fx2 = keep[2].data_ptr()
    if kind == 2:
        ic.fv2 = keep[3].data_ptr()
    return ic, keep


@pytest.mark.parametrize("kind", [1, 2, 4])
@pytest.mark.parametrize("stage", [1, 4])
@pytest.mark.parametrize("n,order", [((32, 8, 8, 5), 4), ((64, 16, 16, 20), 4), ((32, 24, 40, 9), 4), ((32, 8, 8, 7), 6),
                                     ((64, 16, 24, 18), 6), ((32, 8, 16, 17), 6)])
def test_pipe_folds_the_velocity_boundary_fill(lk, ok, fast, n, order, stage, kind):
    """lk_rk_update.accel_bcs + inflow_preset: the pipelined kernel does setaccelerationbcs4d_ (KineticSpeciesF.f:1036-1162)
    inside its boundary tiles -- same bits as `lk_set_acceleration_bcs_4d, then the generic kernel` -- on an array whose
    velocity ghosts hold the inflow sample, and leaves those ghosts alone.  The random acceleration has both signs, so
    inflow and extrapolation are both taken on each of the four boundaries; the grids put the last chunk of the march
    at 1 plane (9 = 8 + 1, 17 = 2 * 8 + 1: top ghosts already in the first window) and at a whole chunk."""
    import torch
    s = Setup(ok, n, order, rough=0.3)
    d = Dev(lk, s)
    ng = s.ng
    chk(lk, lk.lk_periodic_fill_4d(d.f.data_ptr(), C.byref(d.g), 1, 1, None), "per")
    ic, keep = _inflow(d, s, kind)
    # reference: explicit fill, then the generic kernel
    f_ref = d.f.clone()
    at = (C.c_int * 4)(1, 1, 1, 1)
    chk(lk, lk.lk_set_acceleration_bcs_4d(f_ref.data_ptr(), C.byref(d.g), C.byref(d.accel), C.byref(ic), C.byref(at), None), "bcs")
    assert not torch.equal(f_ref, d.f)
    pb, db, mb, used_b = _stage(lk, d, s, stage, 3, 3, 2, f=f_ref)
    # folded: ghosts preset to the inflow sample
    f_pre = d.f.clone()
    chk(lk, lk.lk_preset_inflow_ghosts_4d(f_pre.data_ptr(), C.byref(d.g), C.byref(ic), None), "preset")
    I = (slice(ng, -ng),) * 2
    assert torch.equal(f_pre[I], d.f[I]) and not torch.equal(f_pre, d.f)
    # the preset agrees with the explicit fill wherever that took the inflow branch: some ghost cells, not all
    same = (f_pre == f_ref)
    ghost = torch.ones_like(same)
    ghost[I] = False
    frac = same[ghost].double().mean().item()
    assert 0.2 < frac < 0.9, frac
    before = f_pre.clone()
    pa, da, ma, used_a = _stage(lk, d, s, stage, 3, 3, 0, bcs=ic, f=f_pre, preset=True)
    assert used_a == 1 and used_b == 0
    assert torch.equal(pa, pb) and torch.equal(da, db) and torch.equal(ma, mb)
    assert torch.equal(f_pre, before)
    # without the preset promise, or on the generic kernel (variant 2), the stage runs the fill first: f's ghosts are written
    for variant, preset in ((0, False), (2, True)):
        f2 = d.f.clone()
        pc, dc, mc, used_c = _stage(lk, d, s, stage, 3, 3, variant, bcs=ic, f=f2, preset=preset)
        assert used_c == (1 if variant == 0 else 0) and torch.equal(pc, pb) and torch.equal(mc, mb) and torch.equal(f2, f_ref)


@pytest.mark.parametrize("cut", [1, 2, 3])
@pytest.mark.parametrize("stage", [1, 2, 4])
@pytest.mark.parametrize("n,order", [((96, 24, 16, 12), 4), ((64, 16, 24, 18), 6), ((32, 8, 8, 9), 4)])
def test_pipe_tile_subsets_add_up_to_one_launch(lk, ok, fast, n, order, stage, cut):
    """lk_rk_update.tile_set: the tiles on the faces of the cut directions, then the others (or the other way round)
    leave the same predictor, increment (stages 2-3 update it in place: a tile done twice would show) and moment
    partials as one launch over the whole box; the first launch alone leaves the rest untouched"""
    import torch
    s = Setup(ok, n, order, rough=0.3)
    d = Dev(lk, s)
    chk(lk, lk.lk_periodic_fill_4d(d.f.data_ptr(), C.byref(d.g), 1, 1, None), "per")
    ic, keep = _inflow(d, s, 1)
    chk(lk, lk.lk_preset_inflow_ghosts_4d(d.f.data_ptr(), C.byref(d.g), C.byref(ic), None), "preset")
    whole = _stage(lk, d, s, stage, 3, 3, 0, bcs=ic, preset=True)
    for order_ in ((1, 2), (2, 1)):
        parts = _stage(lk, d, s, stage, 3, 3, 0, bcs=ic, preset=True, tile_sets=order_, cut=cut)
        assert parts[3] == 2 and whole[3] == 1
        assert torch.equal(parts[0], whole[0]) and torch.equal(parts[1], whole[1]) and torch.equal(parts[2], whole[2])
    first = _stage(lk, d, s, stage, 0, 0, 0, bcs=ic, preset=True, tile_sets=(1,), cut=cut)
    ng = s.ng
    inner = first[0][ng:-ng, ng:-ng, ng:-ng, ng:-ng]
    done = inner != 3.0                       # _stage pre-fills the predictor with 3.0
    want = torch.zeros_like(done)
    if cut & 1:
        want[..., :32] = True
        want[..., -32:] = True
    if cut & 2:
        want[:, :, :8, :] = True
        want[:, :, -8:, :] = True
    assert torch.equal(done, want)
    # the generic kernel has no tile subsets: the request fails instead of computing everything twice
    old = lk.lk_set_rhs_variant(2)
    try:
        import loki_b200 as lkm
        u = lkm.RkUpdate()
        pred = torch.zeros_like(d.f)
        u.f_old, u.pred, u.w_delta, u.c_pred, u.tile_set, u.cut_dirs = d.f.data_ptr(), pred.data_ptr(), 0.1, 0.1, 1, cut
        u.delta_out = pred.data_ptr()
        assert lk.lk_vlasov_stage_can_split(None, C.byref(d.g), C.byref(d.accel), C.byref(u)) == 0
        assert lk.lk_vlasov_stage(None, d.f.data_ptr(), C.byref(d.g), d.velocities.data_ptr(), C.byref(d.accel), C.byref(u), None, None) != 0
    finally:
        lk.lk_set_rhs_variant(old)


@pytest.mark.parametrize("mode", [1, 0])
@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("stage", [1, 2, 4])
@pytest.mark.parametrize("n,order", [((32, 8, 8, 9), 4), ((14, 9, 10, 7), 4), ((12, 10, 8, 9), 6)])
def test_krook_term_in_the_fused_stage(lk, ok, n, order, stage, variant, mode):
    """lk_rk_update.krook_*: the fused stage with completeRHS's Krook layer (KineticSpecies.C:1049-1062) against the
    three passes rhs -> lk_append_krook (pinned to appendkrook_) -> lk_rk_stage_update.  Strict arithmetic: the same
    bits; production: the fused kernel contracts the term into an FMA, 1e-13 of the cell's neighbourhood.  The
    aligned grid would take the pipelined kernel without the layer: with it the generic kernel runs."""
    import torch
    import loki_b200 as lkm
    s = Setup(ok, n, order, rough=0.3)
    d = Dev(lk, s)
    chk(lk, lk.lk_periodic_fill_4d(d.f.data_ptr(), C.byref(d.g), 1, 1, None), "per")
    ic, keep = _inflow(d, s, 1)
    n1d, n2d = s.nd[0], s.nd[1]
    nu = np.zeros((n2d, n1d))
    nu[:, : n1d // 3] = np.random.default_rng(4).uniform(0.1, 1.0, size=(n2d, n1d // 3))
    dnu = d.t(nu)
    rng = np.random.default_rng(5)
    f_old = d.t(s.f * (1.0 + 0.01 * rng.uniform(-1, 1, size=s.f.shape)))
    delta0 = d.t(0.001 * s.f * rng.uniform(-1, 1, size=s.f.shape))
    old_mode, old_var = lk.lk_set_strict(mode), lk.lk_set_rhs_variant(variant)
    try:
        def update(pred, delta):
            u = lkm.RkUpdate()
            u.f_old, u.pred = f_old.data_ptr(), pred.data_ptr()
            u.delta_in = None if stage == 1 else delta.data_ptr()
            u.delta_out = None if stage == 4 else delta.data_ptr()
            u.w_delta, u.c_pred, u.use_delta = 0.0123, 0.05, int(stage == 4)
            return u
        # three passes
        rhs = torch.zeros_like(d.f)
        chk(lk, lk.lk_vlasov_rhs(rhs.data_ptr(), d.f.data_ptr(), C.byref(d.g), d.velocities.data_ptr(), C.byref(d.accel), None, None), "rhs")
        plain = rhs.clone()
        chk(lk, lk.lk_append_krook(rhs.data_ptr(), d.f.data_ptr(), C.byref(d.g), dnu.data_ptr(), 0.037, C.byref(ic), None), "krook")
        assert not torch.equal(plain, rhs)
        p3, d3 = torch.full_like(d.f, 3.0), delta0.clone()
        u3 = update(p3, d3)
        chk(lk, lk.lk_rk_stage_update(rhs.data_ptr(), C.byref(d.g), C.byref(u3), None), "update")
        # fused
        p1, d1 = torch.full_like(d.f, 3.0), delta0.clone()
        u1 = update(p1, d1)
        u1.krook_nu, u1.krook_dt, u1.krook_ic = dnu.data_ptr(), 0.037, C.addressof(ic)
        before = lk.lk_pipe_launch_count()
        chk(lk, lk.lk_vlasov_stage(None, d.f.data_ptr(), C.byref(d.g), d.velocities.data_ptr(), C.byref(d.accel), C.byref(u1), None, None), "stage")
        torch.cuda.synchronize()
        assert lk.lk_pipe_launch_count() == before
        if mode == 1:
            assert torch.equal(p1, p3) and torch.equal(d1, d3)
        else:
            scale = np.maximum(np.abs(s.f), 1e-3 * np.abs(s.f).max())
            assert star_rel_err(p1.cpu().numpy(), p3.cpu().numpy(), scale, s.ng) <= 1e-13
            if stage != 4:
                assert star_rel_err(d1.cpu().numpy(), d3.cpu().numpy(), 1e-2 * scale, s.ng) <= 1e-12
        # a layer without tables is refused
        u1.krook_ic = None
        assert lk.lk_vlasov_stage(None, d.f.data_ptr(), C.byref(d.g), d.velocities.data_ptr(), C.byref(d.accel), C.byref(u1), None, None) != 0
    finally:
        lk.lk_set_strict(old_mode)
        lk.lk_set_rhs_variant(old_var)
