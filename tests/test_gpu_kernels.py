"""GPU parity tests proper: every kernel of the hot path, through the C ABI, against the CPU oracle on
the same seeded inputs.  Bars: strict arithmetic == oracle bit for bit; production arithmetic within
the north-star tolerance (1e-12 relative), written next to each assert."""
import ctypes as C

import numpy as np
import pytest

from util import Setup, Dev, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-12  # BASELINE.json north_star: 1e-12 relative on the distribution


def _np(t):
    return t.cpu().numpy()


def chk(lk, status, what):
    assert status == 0, "%s: %s" % (what, lk.lk_last_error().decode())


# ---------------------------------------------------------------- a1/a2 WENO fits
def _stencils(order, count, seed):
    rng = np.random.default_rng(seed)
    w = order
    u = np.empty((count, w))
    k = count // 4
    u[:k] = rng.uniform(-1, 1, size=(k, w))                                   # rough
    xs = rng.uniform(0, 6, size=(k, 1)) + 0.05 * np.arange(w)[None, :]
    u[k:2 * k] = np.exp(-xs ** 2) * 0.16                                        # smooth Maxwellian-like incl. tiny tails
    u[2 * k:3 * k] = np.where(np.arange(w)[None, :] < rng.integers(0, w + 1, size=(k, 1)), 1.0, 0.0)  # steps
    u[3 * k:] = rng.uniform(0.2, 0.3, size=(count - 3 * k, 1))                  # constant rows -> bl == br ties
    vel = rng.uniform(-1, 1, size=count)
    vel[::7] = 0.0                                                              # vel == 0 takes the else branch
    return np.ascontiguousarray(u), vel


@pytest.mark.parametrize("order", [4, 6])
def test_weno_fit_strict_bit_exact(lk, ok, strict, order):
    import torch
    u, vel = _stencils(order, 40000, 7 + order)
    ref = np.empty(len(vel))
    (ok.ok_weno43_fit_v if order == 4 else ok.ok_weno65_fit_v)(u.ravel(), vel, ref, len(vel))
    du, dv = torch.from_numpy(u).cuda(), torch.from_numpy(vel).cuda()
    out = torch.empty_like(dv)
    chk(lk, lk.lk_weno_fit(order, du.data_ptr(), dv.data_ptr(), out.data_ptr(), len(vel), None), "weno")
    assert np.array_equal(_np(out), ref)


@pytest.mark.parametrize("order", [4, 6])
def test_weno_fit_production_tolerance(lk, ok, fast, order):
    import torch
    u, vel = _stencils(order, 40000, 11 + order)
    ref = np.empty(len(vel))
    (ok.ok_weno43_fit_v if order == 4 else ok.ok_weno65_fit_v)(u.ravel(), vel, ref, len(vel))
    du, dv = torch.from_numpy(u).cuda(), torch.from_numpy(vel).cuda()
    out = torch.empty_like(dv)
    chk(lk, lk.lk_weno_fit(order, du.data_ptr(), dv.data_ptr(), out.data_ptr(), len(vel), None), "weno")
    scale = np.max(np.abs(u), axis=1)
    err = np.abs(_np(out) - ref) / np.maximum(scale, 1e-300)
    assert err.max() < TOL, err.max()


# ---------------------------------------------------------------- geometry cases
CASES = [
    ((16, 8, 16, 8), 4),      # tile aligned
    ((10, 10, 16, 10), 6),    # planeIAW_6 deck grid, ragged tiles
    ((13, 5, 9, 7), 4),       # odd everything (emDamping has Ny = 5)
    ((9, 7, 11, 6), 6),
]


@pytest.mark.parametrize("n,order", CASES)
def test_xpby4d(lk, ok, strict, n, order):
    s = Setup(ok, n, order)
    d = Dev(lk, s)
    y = np.random.default_rng(3).uniform(-1, 1, size=s.f.shape)
    x_ref = s.f.copy()
    ok.ok_xpby4d(x_ref.ravel(), y.ravel(), 0.37, C.byref(s.g))
    dy = d.t(y)
    chk(lk, lk.lk_xpby4d(d.f.data_ptr(), dy.data_ptr(), 0.37, C.byref(d.g), None), "xpby4d")
    assert np.array_equal(_np(d.f), x_ref)       # interior updated, ghosts untouched


@pytest.mark.parametrize("maxwell", [False, True])
@pytest.mark.parametrize("n,order", CASES[:3])
def test_phase_space_vel(lk, ok, strict, n, order, maxwell):
    import torch
    s = Setup(ok, n, order, bz=0.3)
    d = Dev(lk, s, maxwell=maxwell)
    vel3, vel4, ax, ay = s.vel34(ok, maxwell)
    d3 = torch.zeros(vel3.size, dtype=torch.float64, device="cuda")
    d4 = torch.zeros(vel4.size, dtype=torch.float64, device="cuda")
    mx = torch.zeros(2, dtype=torch.float64, device="cuda")
    chk(lk, lk.lk_set_phase_space_vel_4d(d3.data_ptr(), d4.data_ptr(), C.byref(d.g), C.byref(d.accel), mx.data_ptr(), None), "vel")
    assert np.array_equal(_np(d3), vel3) and np.array_equal(_np(d4), vel4)
    assert tuple(_np(mx)) == (ax, ay)
    mx2 = torch.zeros(2, dtype=torch.float64, device="cuda")
    chk(lk, lk.lk_max_accel(C.byref(d.g), C.byref(d.accel), mx2.data_ptr(), None), "max_accel")
    assert tuple(_np(mx2)) == (ax, ay)


@pytest.mark.parametrize("n,order", CASES)
def test_periodic_fill(lk, ok, n, order):
    s = Setup(ok, n, order)
    d = Dev(lk, s)
    ref = s.f.copy()
    ok.ok_periodic_fill_4d(ref.ravel(), C.byref(s.g), 1, 1)
    chk(lk, lk.lk_periodic_fill_4d(d.f.data_ptr(), C.byref(d.g), 1, 1, None), "periodic")
    assert np.array_equal(_np(d.f), ref)


@pytest.mark.parametrize("n,order", CASES)
def test_acceleration_bcs(lk, ok, strict, n, order):
    import loki_b200 as lkm
    s = Setup(ok, n, order)
    d = Dev(lk, s)
    vel3, vel4, _, _ = s.vel34(ok)
    cb = s.ic_callback(0.7, 0.9)
    ref = s.f.copy()
    ok.ok_set_acceleration_bcs_4d(ref.ravel(), C.byref(s.g), vel3, vel4, 1, 1, 1, 1, cb, None)
    ic = lkm.Inflow()
    dfx, dfv = d.t(s.fx), d.t(s.fv)
    ic.kind, ic.fx, ic.fv, ic.fnorm, ic.frac = 1, dfx.data_ptr(), dfv.data_ptr(), 0.7, 0.9
    at = (C.c_int * 4)(1, 1, 1, 1)
    chk(lk, lk.lk_set_acceleration_bcs_4d(d.f.data_ptr(), C.byref(d.g), C.byref(d.accel), C.byref(ic), C.byref(at), None), "bcs")
    out = _np(d.f)
    assert np.array_equal(out, ref)
    # both branches (extrapolation and inflow) must have been taken somewhere
    ng = s.ng
    assert np.any(out[:, :ng] != s.f[:, :ng])


@pytest.mark.parametrize("xper,yper", [(0, 0), (1, 0), (0, 1)])
@pytest.mark.parametrize("n,order", CASES)
def test_advection_bcs_nonperiodic(lk, ok, n, order, xper, yper):
    """setAdvectionBCs4D (non-periodic x / y physical boundaries) against the oracle, which
    is pinned to the reference Fortran: bit for bit in both arithmetic modes (the kernel has one build)"""
    import loki_b200 as lkm
    s = Setup(ok, n, order)
    cb = s.ic_callback(0.7, 0.9)
    ref = s.f.copy()
    ok.ok_set_advection_bcs_4d(ref.ravel(), C.byref(s.g), s.vel1, s.vel2, 1, 1, 1, 1, xper, yper, cb, None)
    for strict_mode in (0, 1):
        old = lk.lk_set_strict(strict_mode)
        try:
            d = Dev(lk, s)
            ic = lkm.Inflow()
            dfx, dfv = d.t(s.fx), d.t(s.fv)
            ic.kind, ic.fx, ic.fv, ic.fnorm, ic.frac = 1, dfx.data_ptr(), dfv.data_ptr(), 0.7, 0.9
            at = (C.c_int * 4)(1, 1, 1, 1)
            chk(lk, lk.lk_set_advection_bcs_4d(d.f.data_ptr(), C.byref(d.g), d.velocities.data_ptr(), C.byref(ic), C.byref(at),
                                               xper, yper, None), "advbcs")
            out = _np(d.f)
        finally:
            lk.lk_set_strict(old)
        assert np.array_equal(out, ref)
    ng = s.ng
    if not xper:
        assert np.any(ref[:, :, :, :ng] != s.f[:, :, :, :ng])
    if not yper:
        assert np.any(ref[:, :, :ng, :] != s.f[:, :, :ng, :])


@pytest.mark.parametrize("maxwell", [False, True])
@pytest.mark.parametrize("n,order", CASES)
def test_jb_boundary_conditions(lk, ok, n, order, maxwell):
    """the "JB" boundary conditions (use_new_bcs; SURVEY 8a row a6'): setAccelerationBCs4DJB and
    setAdvectionBCs4DJB against the oracle (pinned to the reference Fortran), bit for bit"""
    import loki_b200 as lkm
    s = Setup(ok, n, order, bz=0.2)
    cb = s.ic_callback(0.7, 0.9)
    vel3, vel4, _, _ = s.vel34(ok, maxwell)
    ref = s.f.copy()
    ok.ok_set_acceleration_bcs_4d_jb(ref.ravel(), C.byref(s.g), vel3, vel4, 1, 1, 1, 1, cb, None)
    d = Dev(lk, s, maxwell=maxwell)
    ic = lkm.Inflow()
    dfx, dfv = d.t(s.fx), d.t(s.fv)
    ic.kind, ic.fx, ic.fv, ic.fnorm, ic.frac = 1, dfx.data_ptr(), dfv.data_ptr(), 0.7, 0.9
    at = (C.c_int * 4)(1, 1, 1, 1)
    chk(lk, lk.lk_set_acceleration_bcs_4d_jb(d.f.data_ptr(), C.byref(d.g), C.byref(d.accel), C.byref(ic), C.byref(at), None), "jb")
    out = _np(d.f)
    assert np.array_equal(out, ref) and np.any(out != s.f)
    if not maxwell:
        for xper, yper in ((0, 0), (1, 0), (0, 1)):
            ref = s.f.copy()
            ok.ok_set_advection_bcs_4d_jb(ref.ravel(), C.byref(s.g), s.vel1, s.vel2, 1, 1, 1, 1, xper, yper, cb, None)
            d = Dev(lk, s)
            chk(lk, lk.lk_set_advection_bcs_4d_jb(d.f.data_ptr(), C.byref(d.g), d.velocities.data_ptr(), C.byref(ic), C.byref(at),
                                                  xper, yper, None), "jbadv")
            out = _np(d.f)
            assert np.array_equal(out, ref) and np.any(out != s.f)


@pytest.mark.parametrize("n,order", CASES)
def test_append_krook(lk, ok, n, order):
    """appendkrook (Krook-layer damping towards the initial condition) against the oracle, pinned to the
    reference Fortran: bit for bit"""
    import loki_b200 as lkm
    s = Setup(ok, n, order)
    cb = s.ic_callback(0.7, 0.9)
    n1d, n2d = s.nd[0], s.nd[1]
    nu = np.zeros((n2d, n1d))
    nu[:, : n1d // 3] = np.random.default_rng(4).uniform(0.1, 1.0, size=(n2d, n1d // 3))
    rhs0 = np.random.default_rng(5).uniform(-1, 1, size=s.f.shape)
    ref = rhs0.copy()
    ok.ok_append_krook(ref.ravel(), s.f.ravel(), C.byref(s.g), nu.ravel(), 0.037, cb, None)
    d = Dev(lk, s)
    ic = lkm.Inflow()
    dfx, dfv = d.t(s.fx), d.t(s.fv)
    ic.kind, ic.fx, ic.fv, ic.fnorm, ic.frac = 1, dfx.data_ptr(), dfv.data_ptr(), 0.7, 0.9
    drhs, dnu = d.t(rhs0), d.t(nu)
    chk(lk, lk.lk_append_krook(drhs.data_ptr(), d.f.data_ptr(), C.byref(d.g), dnu.data_ptr(), 0.037, C.byref(ic), None), "krook")
    out = _np(drhs)
    assert np.array_equal(out, ref) and np.any(out != rhs0)


def _oracle_rhs(ok, s, maxwell=False):
    vel3, vel4, _, _ = s.vel34(ok, maxwell)
    adv = np.zeros_like(s.f)
    ok.ok_advection_derivatives_4d(adv.ravel(), s.f.ravel(), C.byref(s.g), s.vel1, s.vel2)
    full = adv.copy()
    ok.ok_acceleration_derivatives_4d(full.ravel(), s.f.ravel(), C.byref(s.g), vel3, vel4)
    return adv, full


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("n,order", CASES)
def test_derivatives_strict_bit_exact(lk, ok, strict, n, order, variant):
    s = Setup(ok, n, order, bz=0.2)
    d = Dev(lk, s)
    adv, full = _oracle_rhs(ok, s)
    old = lk.lk_set_rhs_variant(variant)
    try:
        rhs = d.zeros_like_f()
        chk(lk, lk.lk_advection_derivatives_4d(rhs.data_ptr(), d.f.data_ptr(), C.byref(d.g), d.velocities.data_ptr(), None), "adv")
        assert np.array_equal(_np(rhs), adv)
        chk(lk, lk.lk_acceleration_derivatives_4d(rhs.data_ptr(), d.f.data_ptr(), C.byref(d.g), C.byref(d.accel), None), "acc")
        assert np.array_equal(_np(rhs), full)
        fused = d.zeros_like_f()
        chk(lk, lk.lk_vlasov_rhs(fused.data_ptr(), d.f.data_ptr(), C.byref(d.g), d.velocities.data_ptr(), C.byref(d.accel), None, None), "rhs")
        assert np.array_equal(_np(fused), full)
    finally:
        lk.lk_set_rhs_variant(old)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("maxwell", [False, True])
@pytest.mark.parametrize("n,order", CASES)
def test_rhs_production_tolerance(lk, ok, fast, n, order, maxwell, variant):
    s = Setup(ok, n, order, bz=0.2)
    d = Dev(lk, s, maxwell=maxwell)
    _, full = _oracle_rhs(ok, s, maxwell)
    old = lk.lk_set_rhs_variant(variant)
    try:
        fused = d.zeros_like_f()
        chk(lk, lk.lk_vlasov_rhs(fused.data_ptr(), d.f.data_ptr(), C.byref(d.g), d.velocities.data_ptr(), C.byref(d.accel), None, None), "rhs")
        assert rel_err(_np(fused), full) < TOL
    finally:
        lk.lk_set_rhs_variant(old)


def test_rhs_nonseparable_tables_strict(lk, ok, strict):
    """relativistic-style tables: the acceleration depends on the face index along its own sweep"""
    s = Setup(ok, (9, 6, 12, 9), 4, bz=0.4, relativistic_like=True)
    d = Dev(lk, s)
    _, full = _oracle_rhs(ok, s)
    for variant in (0, 1):
        old = lk.lk_set_rhs_variant(variant)
        fused = d.zeros_like_f()
        chk(lk, lk.lk_vlasov_rhs(fused.data_ptr(), d.f.data_ptr(), C.byref(d.g), d.velocities.data_ptr(), C.byref(d.accel), None, None), "rhs")
        lk.lk_set_rhs_variant(old)
        assert np.array_equal(_np(fused), full)


@pytest.mark.parametrize("stage", [1, 2, 4])
@pytest.mark.parametrize("n,order", CASES[:2])
def test_fused_rk_stage_strict(lk, ok, strict, n, order, stage):
    """lk_vlasov_rhs + lk_rk_update == evalRHS, addSolnData(delta), copySolnData, addSolnData(pred)
    (RK4Integrator.H:149-171)"""
    import loki_b200 as lkm
    s = Setup(ok, n, order)
    d = Dev(lk, s)
    _, rhs = _oracle_rhs(ok, s)
    rng = np.random.default_rng(5)
    f_old = s.f * (1.0 + 0.01 * rng.uniform(-1, 1, size=s.f.shape))
    delta = np.zeros_like(s.f) if stage == 1 else 0.001 * rng.uniform(-1, 1, size=s.f.shape)
    w, c = 0.0123, (1.0 if stage == 4 else 0.05)
    dref = delta.copy()
    ok.ok_xpby4d(dref.ravel(), rhs.ravel(), w, C.byref(s.g))
    pref = f_old.copy()
    ok.ok_xpby4d(pref.ravel(), (dref if stage == 4 else rhs).ravel(), c, C.byref(s.g))
    dfold, ddelta = d.t(f_old), d.t(delta)
    dout, pred = d.zeros_like_f(), d.t(f_old)   # pred ghosts = f_old ghosts, as copySolnData leaves them
    u = lkm.RkUpdate()
    u.f_old, u.delta_in = dfold.data_ptr(), (None if stage == 1 else ddelta.data_ptr())
    u.delta_out, u.pred, u.w_delta, u.c_pred, u.use_delta = dout.data_ptr(), pred.data_ptr(), w, c, int(stage == 4)
    chk(lk, lk.lk_vlasov_rhs(None, d.f.data_ptr(), C.byref(d.g), d.velocities.data_ptr(), C.byref(d.accel), C.byref(u), None), "stage")
    ng = s.ng
    I = (slice(ng, -ng),) * 4
    assert np.array_equal(_np(dout)[I], dref[I])
    assert np.array_equal(_np(pred), pref)


# ---------------------------------------------------------------- moments
@pytest.mark.parametrize("n,order", CASES)
def test_charge_density(lk, ok, n, order):
    import torch
    s = Setup(ok, n, order)
    d = Dev(lk, s)
    n1d, n2d = s.nd[0], s.nd[1]
    ref = np.zeros(n1d * n2d)
    dv = s.dx[2] * s.dx[3]
    ok.ok_reduce_4d_to_2d(ref, s.f.ravel(), C.byref(s.g), dv, s.charge)
    out = torch.full((n1d * n2d,), 7.0, dtype=torch.float64, device="cuda")
    old = lk.lk_set_strict(1)
    chk(lk, lk.lk_reduce_4d_to_2d(out.data_ptr(), d.f.data_ptr(), C.byref(d.g), dv, s.charge, None), "reduce")
    assert np.array_equal(_np(out), ref)            # strict: the reference's summation order
    lk.lk_set_strict(0)
    chk(lk, lk.lk_reduce_4d_to_2d(out.data_ptr(), d.f.data_ptr(), C.byref(d.g), dv, s.charge, None), "reduce")
    lk.lk_set_strict(old)
    assert rel_err(_np(out), ref) < 1e-14           # chunked sums: 1e-16*sqrt(N) level


@pytest.mark.parametrize("n,order", CASES[:3])
def test_current_density(lk, ok, n, order):
    import torch
    s = Setup(ok, n, order)
    d = Dev(lk, s)
    n1d, n2d = s.nd[0], s.nd[1]
    J4 = [np.zeros_like(s.f) for _ in range(3)]
    ok.ok_compute_currents(C.byref(s.g), s.velocities, s.f.ravel(), s.vz.ravel(), J4[0].ravel(), J4[1].ravel(), J4[2].ravel())
    dv = s.dx[2] * s.dx[3]
    refs = []
    for J in J4:
        r = np.zeros(n1d * n2d)
        ok.ok_reduce_4d_to_2d(r, J.ravel(), C.byref(s.g), dv, s.charge)
        refs.append(r)
    outs = [torch.zeros(n1d * n2d, dtype=torch.float64, device="cuda") for _ in range(3)]
    old = lk.lk_set_strict(1)
    chk(lk, lk.lk_current_density(outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), d.f.data_ptr(), C.byref(d.g),
                                  d.velocities.data_ptr(), d.vz.data_ptr(), dv, s.charge, None), "currents")
    lk.lk_set_strict(old)
    for o, r in zip(outs, refs):
        assert np.array_equal(_np(o), r)


def test_ke_e_dot(lk, ok):
    import torch
    s = Setup(ok, (12, 9, 14, 10), 4)
    d = Dev(lk, s)
    ext = np.ascontiguousarray(np.random.default_rng(2).uniform(-1, 1, size=(2, s.nd[1], s.nd[0])))
    ref = ok.ok_compute_ke_e_dot(C.byref(s.g), s.f.ravel(), s.charge, s.velocities, ext.ravel(), 0.0)
    dext = d.t(ext)
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    chk(lk, lk.lk_ke_e_dot(out.data_ptr(), d.f.data_ptr(), C.byref(d.g), s.charge, d.velocities.data_ptr(), dext.data_ptr(), None), "ke")
    # tree sum vs the reference's sequential sum: tolerance on the sum of magnitudes
    assert abs(_np(out)[0] - ref) <= 1e-13 * max(abs(ref), 1e-300) + 1e-18


# ---------------------------------------------------------------- Poisson
@pytest.mark.parametrize("nx,ny,order", [(32, 32, 4), (10, 10, 6), (16, 7, 4), (12, 5, 6), (64, 128, 4), (256, 16, 6),
                                         (4, 8, 4)])
def test_electric_field(lk, ok, nx, ny, order):
    """power-of-two grids take the shared-memory FFT path in production arithmetic (lk_fft.cu), all others
    and the strict build the direct DFT; both against the oracle's DFT and numpy's FFT"""
    import torch
    ng = 2 if order == 4 else 3
    Lx, Ly = 18.85, 31.4
    rng = np.random.default_rng(9)
    n1d, n2d = nx + 2 * ng, ny + 2 * ng
    rho = np.zeros((n2d, n1d))
    rho[ng:-ng, ng:-ng] = rng.uniform(-1, 1, size=(ny, nx))
    dx = np.array([Lx / nx, Ly / ny, 1.0, 1.0])
    # oracle: EMSolverBase::electricField sequence
    r = rho.copy()
    ok.ok_neutralize_charge(r.ravel(), nx, ny, ng)
    sx, sy = np.zeros(nx), np.zeros(ny // 2 + 1)
    ok.ok_poisson_symbols(nx, ny, Lx, Ly, order, sx, sy)
    phi = np.zeros((n2d, n1d))
    ok.ok_poisson_fft_solve(phi.ravel(), r.ravel(), nx, ny, ng, sx, sy)
    ok.ok_periodic_fill_2d(phi.ravel(), nx, ny, ng, 1, 1, 1)
    em = np.zeros((2, n2d, n1d))
    ok.ok_efield_from_potential(em.ravel(), phi.ravel(), nx, ny, ng, order, 2, dx)
    ok.ok_periodic_fill_2d(em.ravel(), nx, ny, ng, 2, 1, 1)
    # independent check of the oracle's DFT against numpy's FFT (FFTW stand-in)
    rh = np.fft.rfft2(r[ng:-ng, ng:-ng].T)          # [i (x), j (y)]
    den = sx[:, None] + sy[None, :]
    den[0, 0] = 1.0
    phi_np = np.fft.irfft2(rh / den, s=(nx, ny)) * (nx * ny)
    assert rel_err(phi[ng:-ng, ng:-ng].T, phi_np) < 1e-12
    plan = C.c_void_p()
    chk(lk, lk.lk_poisson_plan_create(C.byref(plan), nx, ny, ng, order, Lx, Ly), "plan")
    for strict_mode, tol in ((1, 0.0), (0, 1e-13)):
        old = lk.lk_set_strict(strict_mode)
        drho = torch.from_numpy(rho.copy()).cuda()
        dphi = torch.zeros_like(drho)
        dem = torch.full((2, n2d, n1d), 3.0, dtype=torch.float64, device="cuda")
        cdx = (C.c_double * 4)(*dx)
        chk(lk, lk.lk_electric_field(plan, drho.data_ptr(), dphi.data_ptr(), dem.data_ptr(), cdx, None), "efield")
        lk.lk_set_strict(old)
        if strict_mode:
            assert np.array_equal(_np(drho), r) and np.array_equal(_np(dphi), phi) and np.array_equal(_np(dem), em)
        else:
            # the (0,0) mode is left as it is (LokiPoissonSolveFFT.C:150-158): phi carries the rounding residue
            # of sum(rho - mean), N*eps-sized and dependent on the summation order, as a constant offset
            dp = _np(dphi) - phi
            dp[ng:-ng, ng:-ng] -= dp[ng:-ng, ng:-ng].mean()
            assert np.max(np.abs(dp[ng:-ng, ng:-ng])) < tol * np.max(np.abs(phi))
            assert rel_err(_np(dem), em) < tol
    lk.lk_poisson_plan_destroy(plan)


# ---------------------------------------------------------------- halo slabs == periodic fill
@pytest.mark.parametrize("n,order", CASES[:3])
def test_halo_pack_unpack_equals_periodic(lk, ok, n, order):
    """exchanging packed slabs with oneself (x then y) must reproduce the periodic wrap"""
    import torch
    s = Setup(ok, n, order)
    d = Dev(lk, s)
    ref = s.f.copy()
    ok.ok_periodic_fill_4d(ref.ravel(), C.byref(s.g), 1, 1)
    for direction in (0, 1):
        cnt = lk.lk_halo_count(C.byref(d.g), direction)
        lo = torch.zeros(cnt, dtype=torch.float64, device="cuda")
        hi = torch.zeros(cnt, dtype=torch.float64, device="cuda")
        chk(lk, lk.lk_halo_pack(lo.data_ptr(), d.f.data_ptr(), C.byref(d.g), direction, 0, None), "pack")
        chk(lk, lk.lk_halo_pack(hi.data_ptr(), d.f.data_ptr(), C.byref(d.g), direction, 1, None), "pack")
        # my low interior layers are my high neighbour's... periodic with one rank: low layers -> high ghosts
        chk(lk, lk.lk_halo_unpack(d.f.data_ptr(), lo.data_ptr(), C.byref(d.g), direction, 1, None), "unpack")
        chk(lk, lk.lk_halo_unpack(d.f.data_ptr(), hi.data_ptr(), C.byref(d.g), direction, 0, None), "unpack")
    out = _np(d.f)
    ng = s.ng
    # x ghosts of interior y rows and all y ghosts (incl. their x ghosts) match the reference wrap;
    # identical on everything a star stencil can read
    assert np.array_equal(out[:, :, ng:-ng, :], ref[:, :, ng:-ng, :])
    assert np.array_equal(out[:, :, :, ng:-ng], ref[:, :, :, ng:-ng])


@pytest.mark.parametrize("n,order", [((40, 12, 20, 11), 4), ((9, 17, 8, 21), 4), ((12, 10, 16, 10), 6)])
@pytest.mark.parametrize("nmom", [1, 3])
def test_fused_stage_moments(lk, ok, fast, n, order, nmom):
    """lk_vlasov_stage leaves the velocity moments of its predictor behind: charge density, the vx / vy
    moments and ke_e_dot computed from them equal the stand-alone reductions of the same array
    (ReductionSchedule.C:421-444, KineticSpeciesF.f:2400-2443, 2563-2602) up to summation order"""
    import torch
    import loki_b200 as lkm
    s = Setup(ok, n, order)
    d = Dev(lk, s)
    dfold = d.t(s.f * 1.01)
    ddelta = d.t(0.001 * s.f)
    pred = d.t(s.f * 0.0)
    u = lkm.RkUpdate()
    u.f_old, u.delta_in, u.delta_out, u.pred = dfold.data_ptr(), ddelta.data_ptr(), ddelta.data_ptr(), pred.data_ptr()
    u.w_delta, u.c_pred, u.use_delta = 0.0123, 0.05, 0
    parts = lk.lk_stage_moment_parts(C.byref(d.g))
    assert parts >= 1
    part = torch.full((nmom * parts * n[0] * n[1],), float("nan"), dtype=torch.float64, device="cuda")
    m = lkm.StageMoments()
    m.nmom, m.partial, m.capacity = nmom, part.data_ptr(), part.numel()
    chk(lk, lk.lk_vlasov_stage(None, d.f.data_ptr(), C.byref(d.g), d.velocities.data_ptr(), C.byref(d.accel), C.byref(u), C.byref(m), None), "stage")
    # the same stage without moments gives the same predictor
    pred2 = d.t(s.f * 0.0)
    ddelta2 = d.t(0.001 * s.f)
    u.delta_in, u.delta_out, u.pred = ddelta2.data_ptr(), ddelta2.data_ptr(), pred2.data_ptr()
    chk(lk, lk.lk_vlasov_rhs(None, d.f.data_ptr(), C.byref(d.g), d.velocities.data_ptr(), C.byref(d.accel), C.byref(u), None), "stage2")
    assert torch.equal(pred, pred2)
    n1d, n2d = s.nd[0], s.nd[1]
    dv, wgt = s.dx[2] * s.dx[3], -1.0
    outs = [torch.full((n2d, n1d), 7.0, dtype=torch.float64, device="cuda") for _ in range(3)]
    chk(lk, lk.lk_moments_finish(outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), C.byref(m), C.byref(d.g), dv, wgt, None), "finish")
    ref0 = torch.zeros((n2d, n1d), dtype=torch.float64, device="cuda")
    chk(lk, lk.lk_reduce_4d_to_2d(ref0.data_ptr(), pred.data_ptr(), C.byref(d.g), dv, wgt, None), "rho")
    assert rel_err(_np(outs[0]), _np(ref0)) < 1e-13
    ng = s.ng
    assert np.all(_np(outs[0])[:ng] == 0.0) and np.all(_np(outs[0])[:, :ng] == 0.0)
    if nmom == 3:
        J = [torch.zeros((n2d, n1d), dtype=torch.float64, device="cuda") for _ in range(3)]
        ones = torch.ones((n2d, n1d), dtype=torch.float64, device="cuda")
        chk(lk, lk.lk_current_density(J[0].data_ptr(), J[1].data_ptr(), J[2].data_ptr(), pred.data_ptr(), C.byref(d.g),
                                      d.velocities.data_ptr(), ones.data_ptr(), dv, wgt, None), "J")
        assert rel_err(_np(outs[1]), _np(J[0])) < 1e-13
        assert rel_err(_np(outs[2]), _np(J[1])) < 1e-13
        ext = d.t(np.random.default_rng(3).uniform(-1, 1, size=(2, n2d, n1d)))
        a, b = torch.zeros(1, dtype=torch.float64, device="cuda"), torch.zeros(1, dtype=torch.float64, device="cuda")
        chk(lk, lk.lk_ke_e_dot_from_moments(a.data_ptr(), C.byref(m), C.byref(d.g), -1.0, ext.data_ptr(), None), "ke_m")
        chk(lk, lk.lk_ke_e_dot(b.data_ptr(), pred.data_ptr(), C.byref(d.g), -1.0, d.velocities.data_ptr(), ext.data_ptr(), None), "ke")
        assert abs(float(a) - float(b)) <= 1e-12 * abs(float(b))


# ---------------------------------------------------------------- properties at a larger size
def test_roll_invariance_large(lk, ok, fast):
    """periodic translation by a non-multiple of the tile size permutes the rhs exactly (tile seams)"""
    import torch
    n, order = (40, 24, 32, 24), 4
    s = Setup(ok, n, order, rough=0.2)
    d = Dev(lk, s)
    ng = s.ng

    def rhs_of(f_int):
        f = torch.zeros_like(d.f)
        f[ng:-ng, ng:-ng, ng:-ng, ng:-ng] = f_int
        chk(lk, lk.lk_periodic_fill_4d(f.data_ptr(), C.byref(d.g), 1, 1, None), "per")
        out = torch.zeros_like(d.f)
        # x,y-uniform acceleration so that the operator commutes with x/y translations
        chk(lk, lk.lk_vlasov_rhs(out.data_ptr(), f.data_ptr(), C.byref(d.g), d.velocities.data_ptr(), C.byref(d.accel), None, None), "rhs")
        return out[ng:-ng, ng:-ng, ng:-ng, ng:-ng]

    d.accel_field.copy_(torch.from_numpy(np.broadcast_to(np.array([0.03, -0.02])[:, None, None], s.accel.shape).copy()))
    f0 = d.f[ng:-ng, ng:-ng, ng:-ng, ng:-ng].clone()
    r0 = rhs_of(f0)
    r1 = rhs_of(torch.roll(f0, shifts=(3, 5), dims=(2, 3)))   # dims: (i4,i3,i2,i1)
    assert torch.equal(torch.roll(r0, shifts=(3, 5), dims=(2, 3)), r1)


def test_uniform_state_has_zero_rhs(lk, ok, fast):
    import torch
    s = Setup(ok, (16, 16, 16, 16), 4, rough=0.0)
    d = Dev(lk, s)
    d.f.fill_(0.125)
    out = torch.ones_like(d.f)
    chk(lk, lk.lk_vlasov_rhs(out.data_ptr(), d.f.data_ptr(), C.byref(d.g), d.velocities.data_ptr(), C.byref(d.accel), None, None), "rhs")
    ng = s.ng
    assert float(out[ng:-ng, ng:-ng, ng:-ng, ng:-ng].abs().max()) < 1e-13


def test_variants_agree_large(lk, ok, strict):
    import torch
    s = Setup(ok, (24, 20, 24, 12), 6, rough=0.2)
    d = Dev(lk, s)
    outs = []
    for variant in (0, 1):
        old = lk.lk_set_rhs_variant(variant)
        out = torch.zeros_like(d.f)
        chk(lk, lk.lk_vlasov_rhs(out.data_ptr(), d.f.data_ptr(), C.byref(d.g), d.velocities.data_ptr(), C.byref(d.accel), None, None), "rhs")
        lk.lk_set_rhs_variant(old)
        outs.append(out)
    assert torch.equal(outs[0], outs[1])


def test_argument_errors(lk):
    import loki_b200 as lkm
    g = lkm.Geom.make((8, 8, 8, 8), 4, (1, 1, 1, 1))
    assert lk.lk_xpby4d(None, None, 1.0, C.byref(g), None) == 1
    g.ng = 3
    assert lk.lk_periodic_fill_4d(1, C.byref(g), 1, 1, None) == 1
    assert b"bad argument" in lk.lk_last_error()


# ---------------------------------------------------------------- time-history diagnostics
@pytest.mark.parametrize("n,order", [((9, 6, 10, 7), 4), ((7, 7, 8, 9), 6), ((40, 12, 24, 16), 4)])
def test_compute_ke_and_field_history(lk, ok, n, order):
    """computeke / computekemaxwell and the Poisson / Maxwell field histories on the device against the
    oracle (pinned to the reference Fortran): tree sums vs the reference's cell-by-cell sums, 1e-13"""
    import torch
    s = Setup(ok, n, order)
    d = Dev(lk, s)
    out = torch.zeros(5, dtype=torch.float64, device="cuda")
    o5 = np.zeros(5)
    ok.ok_compute_ke(C.byref(s.g), s.f.ravel(), 1.7, s.velocities, o5)
    chk(lk, lk.lk_compute_ke(out.data_ptr(), d.f.data_ptr(), C.byref(d.g), 1.7, d.velocities.data_ptr(), None, None), "ke")
    got = _np(out)
    assert np.all(np.abs(got - o5) <= 1e-13 * np.abs(o5).max()) and o5[0] > 0
    o3 = np.zeros(3)
    ok.ok_compute_ke_maxwell(C.byref(s.g), s.f.ravel(), 1.7, s.velocities, s.vz.ravel(), o3)
    dvz = d.t(s.vz)
    chk(lk, lk.lk_compute_ke(out.data_ptr(), d.f.data_ptr(), C.byref(d.g), 1.7, d.velocities.data_ptr(), dvz.data_ptr(), None), "kem")
    got = _np(out)
    assert np.all(np.abs(got[:3] - o3) <= 1e-13 * np.abs(o3).max()) and got[3] == 0.0 and got[4] == 0.0
    dx = np.array(s.dx)
    cdx = (C.c_double * 4)(*dx)
    for ncomp in (2, 6):
        em = np.ascontiguousarray(s.em[:ncomp])
        ref = np.zeros(12)
        ok.ok_field_history(em.ravel(), n[0], n[1], s.ng, ncomp, dx, ref)
        dem = d.t(em)
        o = torch.zeros(12, dtype=torch.float64, device="cuda")
        chk(lk, lk.lk_field_history(o.data_ptr(), dem.data_ptr(), n[0], n[1], s.ng, ncomp, cdx, None), "fh")
        cnt = 5 if ncomp == 2 else 12
        got = _np(o)[:cnt]
        assert np.all(np.abs(got - ref[:cnt]) <= 1e-13 * np.abs(ref[:cnt])) and np.all(ref[:cnt] > 0)
