"""Pins the hand-written oracle (oracle/loki_oracle.c) against the REFERENCE'S OWN Fortran source,
transliterated statement by statement to C (oracle/f77toc.py -> oracle/_ref, built by oracle/Makefile
when /root/reference is mounted) and against the golden vectors committed under tests/golden/ that
were generated from that same library (tests/golden/make_golden.py).  Bit-for-bit."""
import ctypes as C
import os

import numpy as np
import pytest

import ref_binding
from util import Setup

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kinetic_f77_golden.npz")
needs_ref = pytest.mark.skipif(not ref_binding.available(), reason="oracle/_ref not built (reference tree absent)")


@pytest.fixture(scope="module")
def ref():
    return ref_binding.Ref()


def _stencils(order, count, seed):
    rng = np.random.default_rng(seed)
    u = rng.uniform(-1, 1, size=(count, order))
    u[::3] = np.exp(-(rng.uniform(0, 6, size=(len(u[::3]), 1)) + 0.05 * np.arange(order)[None, :]) ** 2)
    u[::5] = 0.25
    vel = rng.uniform(-1, 1, size=count)
    vel[::7] = 0.0
    return u, vel


@needs_ref
@pytest.mark.parametrize("order", [4, 6])
def test_pin_weno_fits(ok, ref, order):
    u, vel = _stencils(order, 3000, order)
    for k in range(len(vel)):
        if order == 4:
            assert ok.ok_weno43_fit(*u[k], vel[k]) == ref.weno43(u[k], vel[k])
        else:
            assert ok.ok_weno65_fit(*u[k], vel[k]) == ref.weno65(u[k], vel[k])


CASES = [((7, 6, 9, 8), 4, (0, 0, 0, 0)), ((6, 7, 8, 7), 6, (3, -2, 5, 1))]


@needs_ref
@pytest.mark.parametrize("n,order,lo", CASES)
def test_pin_kinetic_kernels(ok, ref, n, order, lo):
    s = Setup(ok, n, order, bz=0.3)
    R = ref
    db, ib, data, inter = R.boxes(s, lo)
    n1d, n2d, n3d, n4d = s.nd
    # xpby4d
    y = np.random.default_rng(1).uniform(-1, 1, size=s.f.shape)
    x1, x2 = s.f.copy(), s.f.copy()
    ok.ok_xpby4d(x1.ravel(), y.ravel(), 0.37, C.byref(s.g))
    R.L.xpby4d_(R._p(x2), R._p(y), R._d(0.37), *db, *ib)
    assert np.array_equal(x1, x2)
    # setphasespacevel4D / maxwell4D
    for maxwell in (False, True):
        vel3, vel4, ax, ay = s.vel34(ok, maxwell)
        r3, r4 = np.zeros_like(vel3), np.zeros_like(vel4)
        rax, ray = C.c_double(), C.c_double()
        if maxwell:
            R.L.setphasespacevelmaxwell4d_(R._p(r3), R._p(r4), *db, *ib, R._p(s.vxface), R._p(s.vyface), R._d(s.norm),
                                           R._d(s.bz), R._p(s.em), R._p(s.vz), C.byref(rax), C.byref(ray))
        else:
            a2 = [R._i(data[0]), R._i(data[1]), R._i(data[2]), R._i(data[3])]
            R.L.setphasespacevel4d_(R._p(r3), R._p(r4), *db, *ib, R._p(s.vxface), R._p(s.vyface), R._d(s.norm),
                                    R._d(s.bz), R._p(s.accel), *a2, C.byref(rax), C.byref(ray))
        assert np.array_equal(vel3, r3) and np.array_equal(vel4, r4) and (ax, ay) == (rax.value, ray.value)
    vel3, vel4, _, _ = s.vel34(ok, False)
    # setAccelerationBCs4D: global box == data box interior -> touches all four velocity boundaries
    cb = s.ic_callback(0.7, 0.9)
    u1, u2 = s.f.copy(), s.f.copy()
    ok.ok_set_acceleration_bcs_4d(u1.ravel(), C.byref(s.g), vel3, vel4, 1, 1, 1, 1, cb, None)
    lower = (C.c_int * 4)(*[data[2 * k] for k in range(4)])
    R.L.loki_ref_set_ic(cb, None, C.byref(lower))
    R.L.setaccelerationbcs4d_(R._p(u2), *db, *db, *ib, R._i(order), R._p(vel3), R._p(vel4), C.byref(C.c_int64(0)))
    assert np.array_equal(u1, u2)
    # setAdvectionBCs4D: non-periodic x and y, box touching all four configuration-space boundaries; the
    # velocity tables change sign inside the box, so both the outflow and the inflow branch run
    for xper, yper in ((0, 0), (1, 0), (0, 1)):
        u1, u2 = s.f.copy(), s.f.copy()
        ok.ok_set_advection_bcs_4d(u1.ravel(), C.byref(s.g), s.vel1, s.vel2, 1, 1, 1, 1, xper, yper, cb, None)
        R.L.loki_ref_set_ic(cb, None, C.byref(lower))
        R.L.setadvectionbcs4d_(R._p(u2), *db, *db, *ib, R._i(order), R._p(s.vel1), R._p(s.vel2), R._i(xper), R._i(yper),
                               C.byref(C.c_int64(0)))
        assert np.array_equal(u1, u2) and np.any(u1 != s.f)
    # the "JB" boundary conditions (use_new_bcs).  The Fortran finds a global boundary from the cell coordinate
    # (KineticSpeciesF.f:1368-1371, lower + index * dx): a box is at the lower boundary iff its first
    # interior index is 0, and here at the upper one because the domain is made to end with the box
    dxs_ = np.array(s.dx)
    xlo_ = np.zeros(4)
    xhi_ = np.array([(inter[2 * k + 1] + 1) * s.dx[k] for k in range(4)])
    atl = [int(inter[2 * k] == 0) for k in range(4)]
    u1, u2 = s.f.copy(), s.f.copy()
    ok.ok_set_acceleration_bcs_4d_jb(u1.ravel(), C.byref(s.g), vel3, vel4, atl[2], 1, atl[3], 1, cb, None)
    R.L.loki_ref_set_ic(cb, None, C.byref(lower))
    R.L.setaccelerationbcs4djb_(R._p(u2), *db, *db, *ib, R._i(order), R._i(s.ng), R._p(vel3), R._p(vel4), R._p(xlo_), R._p(xhi_),
                                R._p(dxs_), C.byref(C.c_int64(0)))
    assert np.array_equal(u1, u2) and np.any(u1 != s.f)
    for xper, yper in ((0, 0), (1, 0), (0, 1)):
        u1, u2 = s.f.copy(), s.f.copy()
        ok.ok_set_advection_bcs_4d_jb(u1.ravel(), C.byref(s.g), s.vel1, s.vel2, atl[0], 1, atl[1], 1, xper, yper, cb, None)
        R.L.loki_ref_set_ic(cb, None, C.byref(lower))
        R.L.setadvectionbcs4djb_(R._p(u2), *db, *ib, R._i(order), R._i(s.ng), R._p(s.vel1), R._p(s.vel2), R._p(xlo_), R._p(xhi_),
                                 R._p(dxs_), R._i(xper), R._i(yper), C.byref(C.c_int64(0)))
        assert np.array_equal(u1, u2) and np.any(u1 != s.f)
    # derivatives
    dxs = np.array(s.dx)
    r1, r2 = np.zeros_like(s.f), np.zeros_like(s.f)
    ok.ok_advection_derivatives_4d(r1.ravel(), s.f.ravel(), C.byref(s.g), s.vel1, s.vel2)
    R.L.computeadvectionderivatives4d_(R._p(r2), R._p(s.f), *db, *ib, R._p(s.vel1), R._p(s.vel2), R._p(dxs), R._i(order))
    assert np.array_equal(r1, r2)
    ok.ok_acceleration_derivatives_4d(r1.ravel(), s.f.ravel(), C.byref(s.g), vel3, vel4)
    R.L.computeaccelerationderivatives4d_(R._p(r2), R._p(s.f), *db, *ib, R._p(vel3), R._p(vel4), R._p(dxs), R._i(order))
    assert np.array_equal(r1, r2) and np.any(r1 != 0)
    # currents + ke_e_dot
    J1 = [np.zeros_like(s.f) for _ in range(3)]
    J2 = [np.zeros_like(s.f) for _ in range(3)]
    ok.ok_compute_currents(C.byref(s.g), s.velocities, s.f.ravel(), s.vz.ravel(), *[j.ravel() for j in J1])
    R.L.computecurrents_(*db, *ib, R._p(s.velocities), R._p(s.f), R._p(s.vz), *[R._p(j) for j in J2])
    assert all(np.array_equal(a, b) for a, b in zip(J1, J2))
    ext = np.ascontiguousarray(np.random.default_rng(2).uniform(-1, 1, size=(2, n2d, n1d)))
    k1 = ok.ok_compute_ke_e_dot(C.byref(s.g), s.f.ravel(), s.charge, s.velocities, ext.ravel(), 0.0)
    k2 = C.c_double(0.0)
    xlo = np.zeros(4)
    R.L.computekeedot_(*db, *ib, R._p(xlo), R._p(xlo), R._p(dxs), R._p(s.f), R._d(s.charge), R._p(s.velocities), R._p(ext), C.byref(k2))
    assert k1 == k2.value
    # appendkrook: a layer that covers part of configuration space
    nu = np.zeros((n2d, n1d))
    nu[:, : n1d // 3] = np.random.default_rng(4).uniform(0.1, 1.0, size=(n2d, n1d // 3))
    r1k = np.random.default_rng(5).uniform(-1, 1, size=s.f.shape)
    r2k = r1k.copy()
    ok.ok_append_krook(r1k.ravel(), s.f.ravel(), C.byref(s.g), nu.ravel(), 0.037, cb, None)
    R.L.loki_ref_set_ic(cb, None, C.byref(lower))
    R.L.appendkrook_(*db, *ib, R._d(0.037), C.byref(C.c_int64(0)), R._p(nu), R._p(s.f), R._p(r2k))
    assert np.array_equal(r1k, r2k)
    # flux-form diagnostics: computeadvectionfluxes4D / computeaccelerationfluxes4D (face fits + computeFlux4D),
    # accumfluxdiv4D, computekeflux on all eight boundaries, computekevelspaceflux on the four velocity boundaries
    vels = [s.vel1, s.vel2, vel3, vel4]
    fl1, fc1 = [np.zeros_like(v) for v in vels], [np.zeros_like(v) for v in vels]
    fl2, fc2 = [np.zeros_like(v) for v in vels], [np.zeros_like(v) for v in vels]
    for d in range(4):
        ok.ok_face_fluxes_4d(fl1[d], fc1[d], s.f.ravel(), C.byref(s.g), vels[d], d)
    R.L.computeadvectionfluxes4d_(R._p(fl2[0]), R._p(fl2[1]), *db, R._p(vels[0]), R._p(vels[1]), R._p(fc2[0]), R._p(fc2[1]),
                                  R._p(s.f), R._p(dxs), R._i(order))
    R.L.computeaccelerationfluxes4d_(R._p(fl2[2]), R._p(fl2[3]), *db, R._p(vels[2]), R._p(vels[3]), R._p(fc2[2]), R._p(fc2[3]),
                                     R._p(s.f), R._p(dxs), R._i(order))
    for d in range(4):
        assert np.array_equal(fc1[d], fc2[d]) and np.array_equal(fl1[d], fl2[d]) and np.any(fl1[d] != 0), d
    r1, r2 = np.zeros_like(s.f), np.zeros_like(s.f)
    ok.ok_accum_flux_div_4d(r1.ravel(), C.byref(s.g), *fl1)
    R.L.accumfluxdiv4d_(R._p(r2), *db, *ib, *[R._p(a) for a in fl2], R._p(dxs))
    assert np.array_equal(r1, r2) and np.any(r1 != 0)
    for dr in range(4):
        for side in range(2):
            k1 = ok.ok_compute_ke_flux(C.byref(s.g), *fl1, s.velocities, s.vxface, s.vyface, dr, side, 1.7)
            k2 = C.c_double(0.0)
            R.L.computekeflux_(*db, *ib, *ib, R._p(dxs), *[R._p(a) for a in fl2], R._p(s.velocities), R._p(s.vxface),
                               R._p(s.vyface), R._i(dr), R._i(side), R._d(1.7), C.byref(k2))
            assert k1 == k2.value and k1 != 0.0, (dr, side)
            if dr >= 2:
                q1, q2 = np.zeros((n2d, n1d)), np.zeros((n2d, n1d))
                ok.ok_compute_ke_vel_space_flux(q1.ravel(), C.byref(s.g), fl1[2], fl1[3], s.vxface, s.vyface, dr, side, 1.7)
                R.L.computekevelspaceflux_(*db, *ib, *ib, R._p(dxs), R._p(fl2[2]), R._p(fl2[3]), R._p(q2), R._d(1.7),
                                           R._p(s.vxface), R._p(s.vyface), R._i(side), R._i(dr))
                assert np.array_equal(q1, q2) and np.any(q1 != 0)
    # time-history kinetic energies: computeke, computekemaxwell
    o5 = np.zeros(5)
    ok.ok_compute_ke(C.byref(s.g), s.f.ravel(), 1.7, s.velocities, o5)
    r5 = [C.c_double(0.0) for _ in range(5)]
    R.L.computeke_(*db, *ib, R._p(xlo), R._p(xlo), R._p(dxs), R._p(s.f), R._d(1.7), R._p(s.velocities), *[C.byref(v) for v in r5])
    assert [v.value for v in r5] == list(o5) and o5[0] > 0.0
    o3 = np.zeros(3)
    ok.ok_compute_ke_maxwell(C.byref(s.g), s.f.ravel(), 1.7, s.velocities, s.vz.ravel(), o3)
    r3k = [C.c_double(0.0) for _ in range(3)]
    R.L.computekemaxwell_(*db, *ib, R._p(xlo), R._p(xlo), R._p(dxs), R._p(s.f), R._d(1.7), R._p(s.velocities), R._p(s.vz),
                          *[C.byref(v) for v in r3k])
    assert [v.value for v in r3k] == list(o3) and o3[0] > o5[0]


@needs_ref
@pytest.mark.parametrize("order", [4, 6])
def test_pin_field_kernels(ok, ref, order):
    R = ref
    ng = 2 if order == 4 else 3
    n1, n2 = 9, 7
    n1d, n2d = n1 + 2 * ng, n2 + 2 * ng
    rng = np.random.default_rng(3)
    b = lambda lo, n: [R._i(lo - ng), R._i(lo + n - 1 + ng)]
    bi = lambda lo, n: [R._i(lo), R._i(lo + n - 1)]
    db = b(0, n1) + b(0, n2)
    ib = bi(0, n1) + bi(0, n2)
    rho = rng.uniform(-1, 1, size=(n2d, n1d))
    r1, r2 = rho.copy(), rho.copy()
    ok.ok_neutralize_charge(r1.ravel(), n1, n2, ng)
    R.L.neutralizecharge4d_(*db, *ib, R._p(r2), R._i(0))
    assert np.array_equal(r1, r2)
    phi = rng.uniform(-1, 1, size=(n2d, n1d))
    dx = np.array([0.3, 0.7, 1.0, 1.0])
    e1, e2 = np.zeros((2, n2d, n1d)), np.zeros((2, n2d, n1d))
    ok.ok_efield_from_potential(e1.ravel(), phi.ravel(), n1, n2, ng, order, 2, dx)
    R.L.computeefieldfrompotential_(*db, *ib, R._i(order), R._i(2), R._p(dx), R._p(e2), R._p(phi))
    assert np.array_equal(e1, e2)
    # Maxwell rhs (no supergrid layer: SG bounds at the domain ends), with and without dissipation
    em = rng.uniform(-1, 1, size=(6, n2d, n1d))
    J = [rng.uniform(-1, 1, size=(n2d, n1d)) for _ in range(3)]
    xlo, xhi = np.array([0.0, 0.0, 0, 0]), np.array([n1 * dx[0], n2 * dx[1], 0, 0])
    sglo, sghi = xlo[:2].copy(), xhi[:2].copy()
    for avw, avs in ((0.0, 0.0), (0.1, 1.6 / 22.36)):
        m1, m2 = np.zeros_like(em), np.zeros_like(em)
        ok_fn = ok.ok_maxwell_eval_rhs
        ok_fn.argtypes = None
        ok_fn(R._p(m1), R._p(em), R._p(J[0]), R._p(J[1]), R._p(J[2]), n1, n2, ng, order, R._p(dx), C.c_double(22.36),
              C.c_double(avw), C.c_double(avs))
        R.L.maxwellevalrhs_(*db, *ib, R._p(xlo), R._p(xhi), R._p(dx), R._d(22.36), R._d(avw), R._d(avs), R._i(order),
                            R._p(sglo), R._p(sghi), R._p(em), R._p(J[0]), R._p(J[1]), R._p(J[2]), R._p(m2))
        assert np.array_equal(m1, m2)
    x1, x2 = em.copy(), em.copy()
    ok.ok_xpby2d.argtypes = None
    ok.ok_xpby2d(R._p(x1), R._p(m1), C.c_double(0.01), n1, n2, ng, 6)
    R.L.xpby2d_(R._p(x2), R._p(m1), R._d(0.01), *db, *ib, R._i(6))
    assert np.array_equal(x1, x2)


def test_oracle_matches_committed_golden_vectors(ok):
    """golden vectors generated from the transliterated reference Fortran (tests/golden/make_golden.py)"""
    gold = np.load(GOLD)
    for order in (4, 6):
        u, vel, face = gold["weno%d_u" % order], gold["weno%d_vel" % order], gold["weno%d_face" % order]
        out = np.empty(len(vel))
        (ok.ok_weno43_fit_v if order == 4 else ok.ok_weno65_fit_v)(np.ascontiguousarray(u).ravel(), vel, out, len(vel))
        assert np.array_equal(out, face)
        n = tuple(int(v) for v in gold["rhs%d_n" % order])
        s = Setup(ok, n, order, bz=0.3, seed=int(gold["rhs%d_seed" % order]))
        assert np.array_equal(s.f, gold["rhs%d_f" % order])          # same seeded input
        vel3, vel4, ax, ay = s.vel34(ok)
        r = np.zeros_like(s.f)
        ok.ok_advection_derivatives_4d(r.ravel(), s.f.ravel(), C.byref(s.g), s.vel1, s.vel2)
        ok.ok_acceleration_derivatives_4d(r.ravel(), s.f.ravel(), C.byref(s.g), vel3, vel4)
        assert np.array_equal(r, gold["rhs%d_rhs" % order])
        assert (ax, ay) == tuple(gold["rhs%d_amax" % order])
        u1 = s.f.copy()
        ok.ok_set_acceleration_bcs_4d(u1.ravel(), C.byref(s.g), vel3, vel4, 1, 1, 1, 1, s.ic_callback(0.7, 0.9), None)
        assert np.array_equal(u1, gold["rhs%d_bc" % order])
        if "rhs%d_ke" % order in gold.files:
            o5 = np.zeros(5)
            ok.ok_compute_ke(C.byref(s.g), s.f.ravel(), 1.7, s.velocities, o5)
            assert np.array_equal(o5, gold["rhs%d_ke" % order])
            o3 = np.zeros(3)
            ok.ok_compute_ke_maxwell(C.byref(s.g), s.f.ravel(), 1.7, s.velocities, s.vz.ravel(), o3)
            assert np.array_equal(o3, gold["rhs%d_kem" % order])


@needs_ref
def test_fortran_abi_golden_file_is_current(ok, ref):
    """tests/golden/f77abi_golden.npz (what the GPU tests compare the CUDA library with on the box) equals what
    the transliterated reference Fortran returns today for the calls of tests/f77_cases.py"""
    import f77_cases
    gold = np.load(os.path.join(os.path.dirname(GOLD), "f77abi_golden.npz"))
    B = f77_cases.HostBackend(ref.L, ref.L.loki_ref_set_ic)
    n = 0
    for order in (4, 6):
        for k, v in f77_cases.kinetic_cases(B, ok, order).items():
            assert np.array_equal(v, gold["k%d_%s" % (order, k)]), k
            n += 1
        for k, v in f77_cases.field_cases(B, order).items():
            assert np.array_equal(v, gold["f%d_%s" % (order, k)]), k
            n += 1
        for k, v in f77_cases.collision_cases(B, ok, order).items():
            assert np.array_equal(v, gold["c%d_%s" % (order, k)]), k
            n += 1
    assert n == len(gold.files)


def _pitch_case(ok, order, seed=5):
    """a small box, a rough positive distribution, flow / thermal fields of the oracle, a collisional range that
    puts cells into every branch of evaluateCollisionality (inside, both roll-offs, outside)"""
    n = (5, 4, 14, 12) if order == 4 else (4, 5, 16, 14)
    s = Setup(ok, n, order, seed=seed, rough=0.3, vmax=(5.0, 4.0))
    n1d, n2d, n3d, n4d = s.nd
    iv = np.zeros((3, n2d, n1d))
    ok.ok_pitch_angle_fields(iv[0].ravel(), iv[1].ravel(), iv[2].ravel(), s.f.ravel(), C.byref(s.g), s.velocities)
    xlo = np.array([0.0, 0.0, s.vlo[0], s.vlo[1]])
    xhi = np.array([s.L[0], s.L[1], -s.vlo[0], -s.vlo[1]])
    rlo, rhi = np.array([-1.2, -0.9]), np.array([1.1, 1.3])
    return s, iv, xlo, xhi, rlo, rhi


@needs_ref
@pytest.mark.parametrize("order", [4, 6])
def test_pin_pitch_angle_operator(ok, ref, order):
    """PitchAngleCollisionOperatorF.f: collisionality, the four moment routines and the non-conservative operator bit
    for bit; the two Maple-generated conservative operators to rounding (the oracle writes them in operator form)"""
    R = ref
    s, iv, xlo, xhi, rlo, rhi = _pitch_case(ok, order)
    db, ib, data, inter = R.boxes(s)
    n1d, n2d, n3d, n4d = s.nd
    rng = np.random.default_rng(11)
    # evaluateCollisionality over every branch
    vr = 3 if order == 4 else 4
    for _ in range(2000):
        vxg, vyg = rng.uniform(-5.5, 5.5), rng.uniform(-4.5, 4.5)
        wx, wy = vxg - rng.uniform(-0.3, 0.3), vyg - rng.uniform(-0.3, 0.3)
        args = (wx, wy, vxg, vyg, rlo, rhi, -5.0 + vr * s.dx[2], 5.0 - vr * s.dx[2], -4.0 + vr * s.dx[3], 4.0 - vr * s.dx[3],
                0.05, 0.9, 0.37)
        out = C.c_double()
        R.L.evaluatecollisionality_(C.byref(out), *[R._d(a) for a in args[:4]], R._p(rlo), R._p(rhi),
                                    *[R._d(a) for a in args[6:]], R._i(order))
        assert ok.ok_pitch_angle_collisionality(*args, order) == out.value
    # moments -> flow -> kec -> vthermal: the reduction on one rank is the local sum times dvx*dvy
    rn, rgx, rgy, rk = (np.zeros((n2d, n1d)) for _ in range(4))
    R.L.computepitchanglespeciesmoments_(R._p(rn), R._p(rgx), R._p(rgy), R._p(s.f), *db, *ib, R._p(s.velocities))
    m = s.dx[2] * s.dx[3]
    N, Gx, Gy = rn * m, rgx * m, rgy * m
    vx, vy, vth = (np.zeros((n2d, n1d)) for _ in range(3))
    R.L.computepitchanglespeciesreducedfields_(R._p(vx), R._p(vy), R._p(N), R._p(Gx), R._p(Gy), *db)
    R.L.computepitchanglespecieskec_(R._p(rk), R._p(vx), R._p(vy), R._p(s.f), *db, *ib, R._p(s.velocities))
    K = rk * m
    R.L.computepitchanglespeciesvthermal_(R._p(vth), R._p(K), R._p(N), *db)
    assert np.array_equal(iv[0], vx) and np.array_equal(iv[1], vy) and np.array_equal(iv[2], vth)
    # appendPitchAngleCollision: dparams(1) = vfloor, dparams(3) = nuCoeff; iparams = (icons, order, relativity)
    for cons in (0, 1):
        r1 = rng.uniform(-1, 1, size=s.f.shape) * 1e-3
        r2 = r1.copy()
        base = r1.copy()
        ok.ok_append_pitch_angle_collision(r1.ravel(), s.f.ravel(), C.byref(s.g), s.velocities, iv[0].ravel(),
                                           iv[1].ravel(), iv[2].ravel(), xlo[2:].copy(), xhi[2:].copy(), rlo, rhi, 0.05, 0.37, cons)
        dpar = np.array([0.05, 1.0, 0.37])
        ipar = (C.c_int * 3)(cons, order, 0)
        R.L.appendpitchanglecollision_(R._p(r2), R._p(s.f), R._p(s.velocities), R._p(iv[0]), R._p(iv[1]), R._p(iv[2]),
                                       *db, *ib, R._p(xlo), R._p(xhi), R._p(np.array(s.dx)), R._p(rlo), R._p(rhi),
                                       R._p(dpar), ipar)
        if cons == 0:
            assert np.array_equal(r1, r2)
            assert (order == 4) == bool(np.any(r2 != base))      # order 6 has no non-conservative form
        else:
            assert np.any(r2 != base)
            scale = np.abs(r2 - base).max()
            assert np.abs(r1 - r2).max() <= 1e-13 * scale, np.abs(r1 - r2).max() / scale
            ng = s.ng
            g = np.ones(s.f.shape, bool)
            g[ng:-ng, ng:-ng, ng:-ng, ng:-ng] = False
            assert np.array_equal(r1[g], base[g]) and np.array_equal(r2[g], base[g])


@needs_ref
@pytest.mark.parametrize("n,order,lo", CASES)
def test_pin_trig_tz_source(ok, ref, n, order, lo):
    """settrigtzsource / computetrigtzsourceerror (TZSourceF.f:10-137): the oracle's restatement against the
    transliterated Fortran, bit for bit, on a box whose data-box lower bound is not zero"""
    s = Setup(ok, n, order, bz=0.0)
    R = ref
    db, ib, data, inter = R.boxes(s, lo)
    xlo = np.array([-2 * np.pi, -1.5, -7.0, -7.0])
    xhi = -xlo
    dx = np.array(s.dx)
    lo2 = (C.c_int * 2)(data[0], data[2])
    pairs = ((ok.ok_set_trig_tz_source, ok.ok_compute_trig_tz_source_error, R.L.settrigtzsource_, R.L.computetrigtzsourceerror_),
             # ElectronTrigTZSource (ElectronTZSourceF.f): kx = ky = 4
             (ok.ok_set_electron_trig_tz_source, ok.ok_compute_electron_trig_tz_source_error, R.L.setelectrontrigtzsource_,
              R.L.computeelectrontrigtzsourceerror_))
    for ok_set, ok_err, ref_set, ref_err in pairs:
        for time, amp in ((0.0, 1.0), (0.37, 0.1), (2.5, 1.0)):
            base = np.random.default_rng(3).uniform(-1, 1, size=s.f.shape)
            r1, r2 = base.copy(), base.copy()
            ok_set(r1.ravel(), C.byref(s.g), lo2, xlo, dx, time, s.velocities, amp)
            ref_set(R._p(r2), *db, R._p(xlo), R._p(xhi), R._p(dx), R._d(time), R._p(s.velocities), R._p(np.array([amp])))
            assert np.array_equal(r1, r2) and not np.array_equal(r1, base)
            e1, e2 = np.zeros_like(base), np.zeros_like(base)
            ok_err(e1.ravel(), s.f.ravel(), C.byref(s.g), lo2, xlo, dx, time, s.velocities, amp)
            ref_err(R._p(e2), R._p(s.f), *db, R._p(xlo), R._p(xhi), R._p(dx), R._d(time), R._p(s.velocities), R._p(np.array([amp])))
            assert np.array_equal(e1, e2)
    # the two-species sources of the IAWTZ deck (TwoSpecies_ElectronTZSourceF.f, TwoSpecies_IonTZSourceF.f): dparams =
    # {amp, electron_mass, ion_mass}
    two = ((0, R.L.settwoelectrontrigtzsource_, R.L.computetwoelectrontrigtzsourceerror_),
           (1, R.L.settwoiontrigtzsource_, R.L.computetwoiontrigtzsourceerror_))
    for species, ref_set, ref_err in two:
        for time, dparams in ((0.0, [0.1, 1.0, 4.0]), (0.37, [0.1, 1.0, 4.0]), (2.5, [0.7, 0.5, 25.0])):
            dp = np.array(dparams)
            base = np.random.default_rng(3).uniform(-1, 1, size=s.f.shape)
            r1, r2 = base.copy(), base.copy()
            ok.ok_set_two_species_trig_tz_source(r1.ravel(), C.byref(s.g), lo2, xlo, dx, time, s.velocities, dp, species)
            ref_set(R._p(r2), *db, R._p(xlo), R._p(xhi), R._p(dx), R._d(time), R._p(s.velocities), R._p(dp))
            assert np.array_equal(r1, r2) and not np.array_equal(r1, base)
            e1, e2 = np.zeros_like(base), np.zeros_like(base)
            ok.ok_compute_two_species_trig_tz_source_error(e1.ravel(), s.f.ravel(), C.byref(s.g), lo2, xlo, dx, time, s.velocities, dp, species)
            ref_err(R._p(e2), R._p(s.f), *db, R._p(xlo), R._p(xhi), R._p(dx), R._d(time), R._p(s.velocities), R._p(dp))
            assert np.array_equal(e1, e2)


@needs_ref
@pytest.mark.parametrize("order", [4, 6])
def test_pin_maxwell_boundary_routines(ok, ref, order):
    """zeroghost2d, maxwelladdantennasource, maxwellsetembcs, maxwellsetvzbcs (MaxwellF.f:10-58, 359-389, 473-731): a
    box in every position of a 3 x 3 decomposition (touching no, one or two physical boundaries), every periodicity"""
    R = ref
    ng = 2 if order == 4 else 3
    n1, n2 = 9, 7
    n1d, n2d = n1 + 2 * ng, n2 + 2 * ng
    rng = np.random.default_rng(60 + order)
    nx, ny = 3 * n1, 3 * n2
    for bx in range(3):
        for by in range(3):
            lo1, lo2 = bx * n1, by * n2
            db = [R._i(lo1 - ng), R._i(lo1 + n1 - 1 + ng), R._i(lo2 - ng), R._i(lo2 + n2 - 1 + ng)]
            ib = [R._i(lo1), R._i(lo1 + n1 - 1), R._i(lo2), R._i(lo2 + n2 - 1)]
            at = (C.c_int * 4)(int(lo1 == 0), int(lo1 + n1 == nx), int(lo2 == 0), int(lo2 + n2 == ny))
            for xper, yper in ((0, 0), (1, 0), (0, 1), (1, 1)):
                em = rng.uniform(-1, 1, size=(6, n2d, n1d))
                e1, e2 = em.copy(), em.copy()
                ok.ok_maxwell_set_em_bcs(e1.ravel(), n1, n2, order, at, xper, yper, 22.36)
                R.L.maxwellsetembcs_(*db, *ib, R._p(e2), R._i(nx), R._i(ny), R._i(xper), R._i(yper), R._i(order), R._d(22.36))
                assert np.array_equal(e1, e2)
                touched = (not xper and (at[0] or at[1])) or (not yper and (at[2] or at[3]))
                assert np.array_equal(e1, em) != bool(touched)
                vz = rng.uniform(-1, 1, size=(n2d, n1d))
                v1, v2 = vz.copy(), vz.copy()
                ok.ok_maxwell_set_vz_bcs(v1.ravel(), n1, n2, order, at, xper, yper)
                R.L.maxwellsetvzbcs_(*db, *ib, R._p(v2), R._i(nx), R._i(ny), R._i(xper), R._i(yper), R._i(order))
                assert np.array_equal(v1, v2)
    db = [R._i(-ng), R._i(n1 - 1 + ng), R._i(-ng), R._i(n2 - 1 + ng)]
    ib = [R._i(0), R._i(n1 - 1), R._i(0), R._i(n2 - 1)]
    u = rng.uniform(-1, 1, size=(6, n2d, n1d))
    u1, u2 = u.copy(), u.copy()
    ok.ok_zero_ghost_2d(u1.ravel(), n1, n2, ng, 6)
    R.L.zeroghost2d_(R._p(u2), *ib, *db, R._i(6))
    assert np.array_equal(u1, u2) and np.array_equal(u1[:, ng:-ng, ng:-ng], u[:, ng:-ng, ng:-ng]) and u1[0, 0, 0] == 0.0
    src = rng.uniform(-1, 1, size=(6, n2d, n1d))
    d1, d2 = u.copy(), u.copy()
    z4 = np.zeros(4)
    ok.ok_maxwell_add_antenna_source(d1.ravel(), src.ravel(), n1, n2, ng)
    R.L.maxwelladdantennasource_(*db, *ib, R._p(z4), R._p(z4), R._p(z4), R._p(src), R._p(d2))
    assert np.array_equal(d1, d2) and not np.array_equal(d1, u)
