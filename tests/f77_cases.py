"""Seeded calls of the reference's Fortran-ABI routines (KineticSpeciesF.H, PoissonF.H, MaxwellF.H), written ONCE
against a small backend interface so that the very same argument lists drive

  * oracle/_ref/libloki_ref.so -- the reference's own Fortran, transliterated to C (HostBackend), and
  * loki_b200/libloki_b200.so  -- the CUDA kernels behind the same symbols on device arrays (DeviceBackend,
    include/loki_b200_f77.h).

tests/golden/make_golden.py stores the HostBackend outputs in tests/golden/f77abi_golden.npz;
tests/test_gpu_f77abi.py compares the DeviceBackend outputs (strict arithmetic) with them bit for bit.
Test infrastructure only."""
import ctypes as C

import numpy as np

from util import Setup


class HostBackend:
    """arrays stay numpy; the routine works in place"""
    name = "ref"

    def __init__(self, lib, set_ic):
        self.L, self._set_ic, self._keep = lib, set_ic, []

    def arr(self, a):
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(C.c_void_p)

    def meta(self, a):
        return a.ctypes.data_as(C.c_void_p)

    def i(self, v):
        return C.byref(C.c_int(int(v)))

    def d(self, v):
        return C.byref(C.c_double(float(v)))

    def ic(self, s, lower):
        cb = s.ic_callback(0.7, 0.9)
        self._keep.append(cb)
        self._set_ic(cb, None, C.byref((C.c_int * 4)(*lower)))
        return C.byref(C.c_int64(0))

    def call(self, name, *args):
        getattr(self.L, name)(*args)

    def finish(self):
        pass


class DeviceBackend:
    """arrays are uploaded on first use and copied back by finish(); scalars and boxes stay host references"""
    name = "cuda"

    def __init__(self, lib):
        import torch
        import loki_b200 as lkm
        self.torch, self.lkm, self.L = torch, lkm, lib
        self._dev, self._keep = [], []

    def arr(self, a):
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
        for host, t in self._dev:
            if host is a:
                return C.c_void_p(t.data_ptr())
        t = self.torch.from_numpy(a).cuda()
        self._dev.append((a, t))
        return C.c_void_p(t.data_ptr())

    def meta(self, a):
        """domain metadata (xlo, xhi, dx: PROBLEMDOMAIN_TO_FORT, ProblemDomain.H:318-321) stays on the host"""
        self._keep.append(a)
        return a.ctypes.data_as(C.c_void_p)

    def i(self, v):
        return C.byref(C.c_int(int(v)))

    def d(self, v):
        return C.byref(C.c_double(float(v)))

    def ic(self, s, lower):
        # the same initial condition as Setup.ic_callback(0.7, 0.9): fnorm * fv(i3,i4) * fx(i1,i2) * frac
        fx = self.torch.from_numpy(s.fx).cuda()
        fv = self.torch.from_numpy(s.fv).cuda()
        I = self.lkm.Inflow()
        I.kind, I.fx, I.fv, I.fnorm, I.frac = 1, fx.data_ptr(), fv.data_ptr(), 0.7, 0.9
        self._keep += [fx, fv, I]
        return C.byref(C.c_int64(C.addressof(I)))

    def call(self, name, *args):
        fn = getattr(self.L, name)
        fn.restype = None
        fn(*args)
        st = self.L.lk_f77_status()
        assert st == 0, "%s: status %d: %s" % (name, st, self.L.lk_last_error().decode())

    def finish(self):
        self.torch.cuda.synchronize()
        for host, t in self._dev:
            host[...] = t.cpu().numpy()
        self._dev = []


def _boxes(B, s, lo=(0, 0, 0, 0)):
    ng = s.ng
    data, inter = [], []
    for k in range(4):
        data += [lo[k] - ng, lo[k] + s.n[k] - 1 + ng]
        inter += [lo[k], lo[k] + s.n[k] - 1]
    return [B.i(v) for v in data], [B.i(v) for v in inter], data, inter


GRIDS = {4: (6, 5, 7, 6), 6: (7, 7, 8, 7)}
LOWER = {4: (0, 0, 0, 0), 6: (3, -2, 5, 1)}


def kinetic_cases(B, ok, order):
    """every 4D routine of the path on one seeded box; returns name -> output array"""
    n, lo = GRIDS[order], LOWER[order]
    s = Setup(ok, n, order, bz=0.3, seed=300 + order)
    db, ib, data, inter = _boxes(B, s, lo)
    n1d, n2d, n3d, n4d = s.nd
    out = {}
    rng = np.random.default_rng(17 + order)
    # xpby4d
    x, y = s.f.copy(), np.ascontiguousarray(rng.uniform(-1, 1, size=s.f.shape))
    B.call("xpby4d_", B.arr(x), B.arr(y), B.d(0.37), *db, *ib)
    B.finish()
    out["xpby4d"] = x
    # setphasespacevel4d / setphasespacevelmaxwell4d
    for maxwell in (False, True):
        v3 = np.zeros((n3d + 1) * n4d * n1d * n2d)
        v4 = np.zeros((n4d + 1) * n1d * n2d * n3d)
        ax, ay = C.c_double(), C.c_double()
        if maxwell:
            B.call("setphasespacevelmaxwell4d_", B.arr(v3), B.arr(v4), *db, *ib, B.arr(s.vxface), B.arr(s.vyface), B.d(s.norm),
                   B.d(s.bz), B.arr(s.em), B.arr(s.vz), C.byref(ax), C.byref(ay))
        else:
            a2 = [B.i(data[0]), B.i(data[1]), B.i(data[2]), B.i(data[3])]
            B.call("setphasespacevel4d_", B.arr(v3), B.arr(v4), *db, *ib, B.arr(s.vxface), B.arr(s.vyface), B.d(s.norm),
                   B.d(s.bz), B.arr(s.accel), *a2, C.byref(ax), C.byref(ay))
        B.finish()
        tag = "m" if maxwell else ""
        out["vel3" + tag], out["vel4" + tag], out["amax" + tag] = v3, v4, np.array([ax.value, ay.value])
    vel3, vel4 = out["vel3"], out["vel4"]
    # setaccelerationbcs4d: the box touches all four velocity boundaries
    u = s.f.copy()
    B.call("setaccelerationbcs4d_", B.arr(u), *db, *db, *ib, B.i(order), B.arr(vel3), B.arr(vel4), B.ic(s, [data[2 * k] for k in range(4)]))
    B.finish()
    out["accel_bcs"] = u
    # setadvectionbcs4d: non-periodic x and / or y
    for xper, yper in ((0, 0), (1, 0), (0, 1)):
        u = s.f.copy()
        B.call("setadvectionbcs4d_", B.arr(u), *db, *db, *ib, B.i(order), B.arr(s.vel1), B.arr(s.vel2), B.i(xper), B.i(yper),
               B.ic(s, [data[2 * k] for k in range(4)]))
        B.finish()
        out["adv_bcs_%d%d" % (xper, yper)] = u
    # derivatives: advection assigns, acceleration accumulates
    dxs = np.array(s.dx)
    rhs = np.zeros_like(s.f)
    B.call("computeadvectionderivatives4d_", B.arr(rhs), B.arr(s.f), *db, *ib, B.arr(s.vel1), B.arr(s.vel2), B.meta(dxs), B.i(order))
    B.finish()
    out["rhs_adv"] = rhs.copy()
    B.call("computeaccelerationderivatives4d_", B.arr(rhs), B.arr(s.f), *db, *ib, B.arr(vel3), B.arr(vel4), B.meta(dxs), B.i(order))
    B.finish()
    out["rhs_full"] = rhs
    # currents
    J = [np.zeros_like(s.f) for _ in range(3)]
    B.call("computecurrents_", *db, *ib, B.arr(s.velocities), B.arr(s.f), B.arr(s.vz), *[B.arr(j) for j in J])
    B.finish()
    out["jx"], out["jy"], out["jz"] = J
    # computekeedot
    ext = np.ascontiguousarray(rng.uniform(-1, 1, size=(2, n2d, n1d)))
    k = C.c_double(0.0)
    xlo = np.zeros(4)
    B.call("computekeedot_", *db, *ib, B.meta(xlo), B.meta(xlo), B.meta(dxs), B.arr(s.f), B.d(s.charge), B.arr(s.velocities), B.arr(ext),
           C.byref(k))
    B.finish()
    out["ke_e_dot"] = np.array([k.value])
    # flux-form diagnostics: the face / flux arrays of the four directions, their divergence, the kinetic-energy flux
    # through the eight boundaries (box == domain: every boundary is touched) and through the velocity boundaries as
    # a field over (x,y)
    vels = [s.vel1, s.vel2, vel3, vel4]
    face = [np.zeros_like(v) for v in vels]
    flux = [np.zeros_like(v) for v in vels]
    B.call("computeadvectionfluxes4d_", B.arr(flux[0]), B.arr(flux[1]), *db, B.arr(vels[0]), B.arr(vels[1]), B.arr(face[0]),
           B.arr(face[1]), B.arr(s.f), B.meta(dxs), B.i(order))
    B.call("computeaccelerationfluxes4d_", B.arr(flux[2]), B.arr(flux[3]), *db, B.arr(vels[2]), B.arr(vels[3]), B.arr(face[2]),
           B.arr(face[3]), B.arr(s.f), B.meta(dxs), B.i(order))
    div = np.zeros_like(s.f)
    B.call("accumfluxdiv4d_", B.arr(div), *db, *ib, *[B.arr(a) for a in flux], B.meta(dxs))
    kef = np.zeros(8)
    kev = [np.zeros((n2d, n1d)) for _ in range(4)]
    for dr in range(4):
        for side in range(2):
            kk = C.c_double(0.0)
            B.call("computekeflux_", *db, *ib, *ib, B.meta(dxs), *[B.arr(a) for a in flux], B.arr(s.velocities), B.arr(s.vxface),
                   B.arr(s.vyface), B.i(dr), B.i(side), B.d(1.7), C.byref(kk))
            kef[2 * dr + side] = kk.value
            if dr >= 2:
                B.call("computekevelspaceflux_", *db, *ib, *ib, B.meta(dxs), B.arr(flux[2]), B.arr(flux[3]),
                       B.arr(kev[2 * (dr - 2) + side]), B.d(1.7), B.arr(s.vxface), B.arr(s.vyface), B.i(side), B.i(dr))
    B.finish()
    for d in range(4):
        out["face%d" % (d + 1)], out["flux%d" % (d + 1)] = face[d], flux[d]
    out["flux_div"], out["ke_flux"] = div, kef
    for k in range(4):
        out["ke_vel_flux%d" % k] = kev[k]
    # appendkrook: a layer over part of configuration space
    nu = np.zeros((n2d, n1d))
    nu[:, : n1d // 3] = rng.uniform(0.1, 1.0, size=(n2d, n1d // 3))
    rk = np.ascontiguousarray(rng.uniform(-1, 1, size=s.f.shape))
    B.call("appendkrook_", *db, *ib, B.d(0.037), B.ic(s, [data[2 * k] for k in range(4)]), B.arr(nu), B.arr(s.f), B.arr(rk))
    B.finish()
    out["krook"] = rk
    # TrigTZSource (TZSourceF.f:10-137): the twilight-zone source added to a right-hand side over the data box, and the
    # error of a state against the exact solution; only the data box is passed, the domain arrays stay on the host
    xlo4 = np.array([-2 * np.pi, -1.5, -7.0, -7.0])
    xhi4 = -xlo4
    tz = np.ascontiguousarray(rng.uniform(-1, 1, size=s.f.shape))
    amp = np.array([0.7])
    B.call("settrigtzsource_", B.arr(tz), *db, B.meta(xlo4), B.meta(xhi4), B.meta(dxs), B.d(0.37), B.arr(s.velocities), B.meta(amp))
    err = np.zeros_like(s.f)
    B.call("computetrigtzsourceerror_", B.arr(err), B.arr(s.f), *db, B.meta(xlo4), B.meta(xhi4), B.meta(dxs), B.d(0.37),
           B.arr(s.velocities), B.meta(amp))
    B.finish()
    out["tz_source"], out["tz_error"] = tz, err
    # ElectronTrigTZSource (ElectronTZSourceF.f:10-143): the same lists, kx = ky = 4
    tze = np.ascontiguousarray(rng.uniform(-1, 1, size=s.f.shape))
    B.call("setelectrontrigtzsource_", B.arr(tze), *db, B.meta(xlo4), B.meta(xhi4), B.meta(dxs), B.d(0.37), B.arr(s.velocities), B.meta(amp))
    erre = np.zeros_like(s.f)
    B.call("computeelectrontrigtzsourceerror_", B.arr(erre), B.arr(s.f), *db, B.meta(xlo4), B.meta(xhi4), B.meta(dxs), B.d(0.37),
           B.arr(s.velocities), B.meta(amp))
    B.finish()
    out["etz_source"], out["etz_error"] = tze, erre
    # the two-species sources (TwoSpecies_ElectronTZSourceF.f, TwoSpecies_IonTZSourceF.f): dparams = {amp, me, mi}
    dpar = np.array([0.7, 0.5, 25.0])
    for tag, set_name, err_name in (("tz2e", "settwoelectrontrigtzsource_", "computetwoelectrontrigtzsourceerror_"),
                                    ("tz2i", "settwoiontrigtzsource_", "computetwoiontrigtzsourceerror_")):
        src2 = np.ascontiguousarray(rng.uniform(-1, 1, size=s.f.shape))
        B.call(set_name, B.arr(src2), *db, B.meta(xlo4), B.meta(xhi4), B.meta(dxs), B.d(0.37), B.arr(s.velocities), B.meta(dpar))
        err2 = np.zeros_like(s.f)
        B.call(err_name, B.arr(err2), B.arr(s.f), *db, B.meta(xlo4), B.meta(xhi4), B.meta(dxs), B.d(0.37), B.arr(s.velocities),
               B.meta(dpar))
        B.finish()
        out[tag + "_source"], out[tag + "_error"] = src2, err2
    return out


def field_cases(B, order):
    """the 2D routines (Poisson pieces, Maxwell rhs, vz rhs, xpby2d)"""
    ng = 2 if order == 4 else 3
    n1, n2 = 9, 7
    n1d, n2d = n1 + 2 * ng, n2 + 2 * ng
    rng = np.random.default_rng(40 + order)
    db = [B.i(-ng), B.i(n1 - 1 + ng), B.i(-ng), B.i(n2 - 1 + ng)]
    ib = [B.i(0), B.i(n1 - 1), B.i(0), B.i(n2 - 1)]
    out = {}
    rho = np.ascontiguousarray(rng.uniform(-1, 1, size=(n2d, n1d)))
    B.call("neutralizecharge4d_", *db, *ib, B.arr(rho), B.i(0))
    B.finish()
    out["neutral"] = rho
    phi = np.ascontiguousarray(rng.uniform(-1, 1, size=(n2d, n1d)))
    dx = np.array([0.3, 0.7, 1.0, 1.0])
    e = np.zeros((2, n2d, n1d))
    B.call("computeefieldfrompotential_", *db, *ib, B.i(order), B.i(2), B.meta(dx), B.arr(e), B.arr(phi))
    B.finish()
    out["efield"] = e
    em = np.ascontiguousarray(rng.uniform(-1, 1, size=(6, n2d, n1d)))
    J = [np.ascontiguousarray(rng.uniform(-1, 1, size=(n2d, n1d))) for _ in range(3)]
    xlo, xhi = np.array([0.0, 0.0, 0, 0]), np.array([n1 * dx[0], n2 * dx[1], 0, 0])
    sglo, sghi = xlo[:2].copy(), xhi[:2].copy()
    for tag, avw, avs in (("", 0.0, 0.0), ("_av", 0.1, 1.6 / 22.36)):
        m = np.zeros_like(em)
        B.call("maxwellevalrhs_", *db, *ib, B.meta(xlo), B.meta(xhi), B.meta(dx), B.d(22.36), B.d(avw), B.d(avs), B.i(order),
               B.meta(sglo), B.meta(sghi), B.arr(em), B.arr(J[0]), B.arr(J[1]), B.arr(J[2]), B.arr(m))
        B.finish()
        out["maxwell" + tag] = m
    dvz = np.zeros((n2d, n1d))
    B.call("maxwellevalvzrhs_", *db, *ib, B.d(-1.25), B.arr(em), B.arr(dvz))
    B.finish()
    out["vzrhs"] = dvz
    x = em.copy()
    B.call("xpby2d_", B.arr(x), B.arr(out["maxwell_av"]), B.d(0.01), *db, *ib, B.i(6))
    B.finish()
    out["xpby2d"] = x
    # the boundary routines a non-periodic Vlasov-Maxwell run reaches (MaxwellF.f:10-58, 359-389, 473-731): the box in the
    # low-x / high-y corner of a 2 x 2 decomposition and in the other corner, every periodicity
    nx, ny = 2 * n1, 2 * n2
    for tag, (l1, l2) in (("lo_hi", (0, n2)), ("hi_lo", (n1, 0))):
        dbb = [B.i(l1 - ng), B.i(l1 + n1 - 1 + ng), B.i(l2 - ng), B.i(l2 + n2 - 1 + ng)]
        ibb = [B.i(l1), B.i(l1 + n1 - 1), B.i(l2), B.i(l2 + n2 - 1)]
        for xper, yper in ((0, 0), (1, 0), (0, 1)):
            e = np.ascontiguousarray(rng.uniform(-1, 1, size=(6, n2d, n1d)))
            B.call("maxwellsetembcs_", *dbb, *ibb, B.arr(e), B.i(nx), B.i(ny), B.i(xper), B.i(yper), B.i(order), B.d(22.36))
            v = np.ascontiguousarray(rng.uniform(-1, 1, size=(n2d, n1d)))
            B.call("maxwellsetvzbcs_", *dbb, *ibb, B.arr(v), B.i(nx), B.i(ny), B.i(xper), B.i(yper), B.i(order))
            B.finish()
            out["embcs_%s_%d%d" % (tag, xper, yper)], out["vzbcs_%s_%d%d" % (tag, xper, yper)] = e, v
    z = np.ascontiguousarray(rng.uniform(-1, 1, size=(6, n2d, n1d)))
    B.call("zeroghost2d_", B.arr(z), *ib, *db, B.i(6))
    src = np.ascontiguousarray(rng.uniform(-1, 1, size=(6, n2d, n1d)))
    dem = np.ascontiguousarray(rng.uniform(-1, 1, size=(6, n2d, n1d)))
    z4 = np.zeros(4)
    B.call("maxwelladdantennasource_", *db, *ib, B.meta(z4), B.meta(z4), B.meta(z4), B.arr(src), B.arr(dem))
    B.finish()
    out["zeroghost"], out["antenna"] = z, dem
    return out


COLL_GRIDS = {4: (5, 4, 14, 12), 6: (4, 5, 16, 14)}


def collision_cases(B, ok, order):
    """PitchAngleCollisionOperator::evaluate piece by piece through the Fortran ABI (PitchAngleCollisionOperatorF.H): raw
    moments -> (times dvx dvy, the one-rank ReductionSchedule4D) -> flow -> kec -> thermal speed -> the operator,
    conservative and not.  The collisional range leaves cells in every branch of evaluateCollisionality."""
    s = Setup(ok, COLL_GRIDS[order], order, seed=500 + order, rough=0.3, vmax=(5.0, 4.0))
    db, ib, data, inter = _boxes(B, s)
    n1d, n2d, n3d, n4d = s.nd
    out = {}
    rn, rgx, rgy, rk = (np.zeros((n2d, n1d)) for _ in range(4))
    B.call("computepitchanglespeciesmoments_", B.arr(rn), B.arr(rgx), B.arr(rgy), B.arr(s.f), *db, *ib, B.arr(s.velocities))
    B.finish()
    m = s.dx[2] * s.dx[3]
    N, Gx, Gy = rn * m, rgx * m, rgy * m
    vx, vy, vth = (np.zeros((n2d, n1d)) for _ in range(3))
    B.call("computepitchanglespeciesreducedfields_", B.arr(vx), B.arr(vy), B.arr(N), B.arr(Gx), B.arr(Gy), *db)
    B.call("computepitchanglespecieskec_", B.arr(rk), B.arr(vx), B.arr(vy), B.arr(s.f), *db, *ib, B.arr(s.velocities))
    B.finish()
    K = rk * m
    B.call("computepitchanglespeciesvthermal_", B.arr(vth), B.arr(K), B.arr(N), *db)
    B.finish()
    out["moments"] = np.stack([rn, rgx, rgy, rk])
    out["fields"] = np.stack([vx, vy, vth])
    xlo = np.array([0.0, 0.0, s.vlo[0], s.vlo[1]])
    xhi = np.array([s.L[0], s.L[1], -s.vlo[0], -s.vlo[1]])
    dxs = np.array(s.dx)
    rlo, rhi = np.array([-1.2, -0.9]), np.array([1.1, 1.3])
    dpar = np.array([0.05, 1.0, 0.37])
    rng = np.random.default_rng(23 + order)
    for cons in (0, 1):
        r = np.ascontiguousarray(rng.uniform(-1, 1, size=s.f.shape) * 1e-3)
        ipar = np.array([cons, order, 0], dtype=np.int32)
        B.call("appendpitchanglecollision_", B.arr(r), B.arr(s.f), B.arr(s.velocities), B.arr(vx), B.arr(vy), B.arr(vth), *db, *ib,
               B.meta(xlo), B.meta(xhi), B.meta(dxs), B.meta(rlo), B.meta(rhi), B.meta(dpar), B.meta(ipar))
        B.finish()
        out["coll_cons" if cons else "coll_noncons"] = r
    return out
