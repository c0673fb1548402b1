"""Parity properties at BASELINE.json's full single-GPU size (256 x 256 x 128 x 128 cells per species, 9.4 GB
per array), where the oracle cannot run: size-independent properties of the fused stage kernel.

 * translation: every operation is position-independent and explicitly rounded, so evaluating a periodically
   shifted state (and field) must give the shifted result BIT FOR BIT -- this exercises every tile, halo,
   TMA coordinate and 64-bit offset of the full-size launch;
 * two implementations: the marching TMA kernel against the one-thread-per-cell kernel (variant 1), which
   shares only the face-fit arithmetic -- identical bits expected in production arithmetic;
 * telescoping: the x and y advection terms sum to zero along a periodic line (KineticSpeciesF.f:1990-2025
   reuses uLeft = uRight), up to rounding;
 * RK identities: c_pred = 0 returns f_old unchanged, w_delta = 0 returns delta_in unchanged."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = (256, 256, 128, 128)


def _setup(lkm, torch, n, order=4, seed=3):
    ng = 2 if order == 4 else 3
    nd = [k + 2 * ng for k in n]
    dev = torch.device("cuda:0")
    g = lkm.Geom.make(n, order, (0.07, 0.09, 0.11, 0.11))
    v3 = (torch.arange(nd[2], device=dev, dtype=torch.float64) - ng + 0.5) * 0.11 - 0.055 * n[2]
    v4 = (torch.arange(nd[3], device=dev, dtype=torch.float64) - ng + 0.5) * 0.11 - 0.055 * n[3]
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    # periodic interior fields: smooth modes plus 2 % roughness so that the WENO weights leave 1/2
    xi = torch.arange(n[0], device=dev, dtype=torch.float64)
    yi = torch.arange(n[1], device=dev, dtype=torch.float64)
    fx = 1.0 + 0.1 * torch.cos(2 * np.pi * 3 * xi / n[0])[None, :] * torch.cos(2 * np.pi * 2 * yi / n[1])[:, None]
    fv = torch.exp(-0.5 * (v4[:, None] ** 2 + v3[None, :] ** 2)) / (2 * np.pi)
    f = torch.zeros(nd[3], nd[2], nd[1], nd[0], device=dev, dtype=torch.float64)
    for lo in range(0, nd[3], 8):            # in slabs of vy planes: bounded temporaries at the 2^31-cell size
        sl = slice(lo, min(lo + 8, nd[3]))
        blk = fv[sl, :, None, None] * fx[None, None, :, :]
        blk.mul_(1.0 + 0.02 * (torch.rand(blk.shape, generator=gen, device=dev, dtype=torch.float64) - 0.5))
        f[sl, :, ng:-ng, ng:-ng] = blk
    del blk
    accel = torch.zeros(2, nd[1], nd[0], device=dev, dtype=torch.float64)
    accel[:, ng:-ng, ng:-ng] = 0.05 * (torch.rand(2, n[1], n[0], generator=gen, device=dev, dtype=torch.float64) - 0.5)
    vel = torch.stack([v3[None, :].expand(nd[3], nd[2]), v4[:, None].expand(nd[3], nd[2])]).contiguous()
    vxf = torch.zeros(2, nd[3], nd[2] + 1, device=dev, dtype=torch.float64)
    vxf[1] = v4[:, None]
    vyf = torch.zeros(2, nd[3] + 1, nd[2], device=dev, dtype=torch.float64)
    vyf[0] = v3[None, :]
    return dict(g=g, ng=ng, nd=nd, f=f, accel=accel, vel=vel, vxf=vxf, vyf=vyf, dev=dev)


def _wrap2d(a, ng):
    """periodic ghost fill of a (comp, n2d, n1d) array from its interior"""
    a[:, :, :ng] = a[:, :, -2 * ng:-ng]
    a[:, :, -ng:] = a[:, :, ng:2 * ng]
    a[:, :ng, :] = a[:, -2 * ng:-ng, :]
    a[:, -ng:, :] = a[:, ng:2 * ng, :]


def _accel_desc(lkm, S):
    A = lkm.Accel()
    A.kind, A.field, A.vz = 0, S["accel"].data_ptr(), None
    A.vxface_velocities, A.vyface_velocities = S["vxf"].data_ptr(), S["vyf"].data_ptr()
    A.normalization, A.bz_const = -1.0, 0.0
    return A


def _prepare(lk, lkm, S):
    """x/y periodic wrap + velocity-boundary fill (zero inflow) of S['f'], as a stage does before the kernel"""
    _wrap2d(S["accel"], S["ng"])
    g = S["g"]
    assert lk.lk_periodic_fill_4d(S["f"].data_ptr(), C.byref(g), 1, 1, None) == 0
    A = _accel_desc(lkm, S)
    ic = lkm.Inflow()
    ic.kind = 0
    at = (C.c_int * 4)(1, 1, 1, 1)
    assert lk.lk_set_acceleration_bcs_4d(S["f"].data_ptr(), C.byref(g), C.byref(A), C.byref(ic), C.byref(at), None) == 0
    return A


def _rhs(lk, lkm, S, out):
    A = _accel_desc(lkm, S)
    assert lk.lk_vlasov_rhs(out.data_ptr(), S["f"].data_ptr(), C.byref(S["g"]), S["vel"].data_ptr(), C.byref(A), None, None) == 0, \
        lk.lk_last_error()


# the headline single-GPU size (order 4) and the per-GPU tile of the 8-GPU InterpenetratingStreams
# configuration (order 6): 2^31 interior cells, 2.4e9 elements per array -- every offset needs 64 bits
@pytest.mark.parametrize("n,order,need_gb", [(N, 4, 60), ((256, 128, 256, 256), 6, 130)])
def test_fullsize_translation_and_two_kernels(lk, fast, n, order, need_gb):
    import torch
    import loki_b200 as lkm
    if torch.cuda.get_device_properties(0).total_memory < need_gb * 1e9:
        pytest.skip("needs %d GB of device memory" % need_gb)
    N = n
    S = _setup(lkm, torch, N, order=order)
    ng = S["ng"]
    _prepare(lk, lkm, S)
    rhs = torch.zeros_like(S["f"])
    _rhs(lk, lkm, S, rhs)
    I = (slice(ng, -ng),) * 4
    assert bool(torch.isfinite(rhs[I]).all()) and float(rhs[I].abs().max()) > 0.0
    # ---- telescoping of the advection terms along periodic lines ----
    adv = torch.zeros_like(S["f"])
    assert lk.lk_advection_derivatives_4d(adv.data_ptr(), S["f"].data_ptr(), C.byref(S["g"]), S["vel"].data_ptr(), None) == 0
    scale = float(adv[I].abs().sum())
    # sum over x and y of (x term + y term) vanishes for every (vx, vy)
    line = adv[I].sum(dim=(2, 3))
    assert float(line.abs().max()) <= 1e-9 * scale / (N[2] * N[3])
    del adv, line
    torch.cuda.empty_cache()
    # ---- the one-thread-per-cell kernel computes the same bits ----
    rhs1 = torch.zeros_like(S["f"])
    old = lk.lk_set_rhs_variant(1)
    try:
        _rhs(lk, lkm, S, rhs1)
    finally:
        lk.lk_set_rhs_variant(old)
    # the two kernels fold 1/12 at different places (face vs flux coefficient): agreement to rounding of the
    # neighbourhood, not bit for bit
    d = (rhs1[I] - rhs[I]).abs_()
    den = rhs[I].abs().amax(dim=(2, 3), keepdim=True).clamp_min(1e-300)
    d.div_(den)
    assert float(d.max()) <= 1e-12
    del rhs1, d, den
    torch.cuda.empty_cache()
    # ---- translation by (37, 5) cells in (x, y): bit-identical ----
    sx, sy = 37, 5
    f2 = torch.zeros_like(S["f"])
    f2[:, :, ng:-ng, ng:-ng] = torch.roll(S["f"][:, :, ng:-ng, ng:-ng], shifts=(sy, sx), dims=(2, 3))
    a2 = torch.zeros_like(S["accel"])
    a2[:, ng:-ng, ng:-ng] = torch.roll(S["accel"][:, ng:-ng, ng:-ng], shifts=(sy, sx), dims=(1, 2))
    want = torch.roll(rhs[I], shifts=(sy, sx), dims=(2, 3))
    del rhs
    S2 = dict(S, f=f2, accel=a2)
    _prepare(lk, lkm, S2)
    rhs2 = torch.zeros_like(f2)
    _rhs(lk, lkm, S2, rhs2)
    assert bool(torch.equal(rhs2[I], want))


def test_fullsize_rk_identities(lk, fast):
    import torch
    import loki_b200 as lkm
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs 60 GB of device memory")
    S = _setup(lkm, torch, N, seed=5)
    ng = S["ng"]
    A = _prepare(lk, lkm, S)
    I = (slice(ng, -ng),) * 4
    f_old = S["f"].clone()
    f_old.mul_(1.25)
    delta = torch.full_like(S["f"], 0.5)
    pred = torch.zeros_like(S["f"])
    dout = torch.zeros_like(S["f"])
    U = lkm.RkUpdate()
    U.f_old, U.delta_in, U.delta_out, U.pred = f_old.data_ptr(), delta.data_ptr(), dout.data_ptr(), pred.data_ptr()
    U.w_delta, U.c_pred, U.use_delta = 0.0, 0.0, 0
    assert lk.lk_vlasov_rhs(None, S["f"].data_ptr(), C.byref(S["g"]), S["vel"].data_ptr(), C.byref(A), C.byref(U), None) == 0, \
        lk.lk_last_error()
    assert bool(torch.equal(pred[I], f_old[I]))           # f_old + 0 * rhs
    assert bool(torch.equal(dout[I], delta[I]))           # delta_in + 0 * rhs
    # linearity of the update in c_pred: pred(c) - f_old is c * rhs to rounding
    rhs = torch.zeros_like(S["f"])
    _rhs(lk, lkm, S, rhs)
    U.w_delta, U.c_pred = 0.25, 0.125
    assert lk.lk_vlasov_rhs(None, S["f"].data_ptr(), C.byref(S["g"]), S["vel"].data_ptr(), C.byref(A), C.byref(U), None) == 0
    want = torch.addcmul(f_old[I], rhs[I], torch.tensor(0.125, dtype=torch.float64, device=rhs.device))
    err = (pred[I] - want).abs().max() / f_old[I].abs().max()
    assert float(err) <= 1e-15
    wantd = delta[I] + 0.25 * rhs[I]
    assert float((dout[I] - wantd).abs().max()) <= 1e-15 * float(wantd.abs().max())
