"""Shared helpers for the parity tests: synthetic species set-ups and device plumbing (torch is used
only to own device memory)."""
import ctypes as C

import numpy as np

from oracle_binding import OkGeom, IC_FN


def nd_of(n, order):
    ng = 2 if order == 4 else 3
    return tuple(k + 2 * ng for k in n), ng


class Setup:
    """A single-species box with analytic velocity tables and a smooth + rough distribution."""

    def __init__(self, ok, n, order, seed=1234, rough=0.3, vmax=(7.0, 7.0), L=(18.85, 490.0), bz=0.0,
                 charge=-1.0, mass=1.0, relativistic_like=False):
        self.n = tuple(n)
        self.order = order
        self.nd, self.ng = nd_of(n, order)
        ng = self.ng
        n1d, n2d, n3d, n4d = self.nd
        self.dx = (L[0] / n[0], L[1] / n[1], 2 * vmax[0] / n[2], 2 * vmax[1] / n[3])
        self.L = L
        self.vlo = (-vmax[0], -vmax[1])
        self.charge, self.mass, self.bz = charge, mass, bz
        self.norm = charge / mass
        self.g = OkGeom.make(n, order, self.dx)
        rng = np.random.default_rng(seed)
        self.velocities = np.zeros(n3d * n4d * 2)
        self.vxface = np.zeros((n3d + 1) * n4d * 2)
        self.vyface = np.zeros(n3d * (n4d + 1) * 2)
        lo = (C.c_int * 2)(-ng, -ng)
        ok.ok_build_velocity_tables(C.byref(self.g), C.byref(lo), self.vlo[0], self.vlo[1], self.velocities,
                                    self.vxface, self.vyface)
        if relativistic_like:
            # make the tables non-separable (as with Simulation::s_DO_RELATIVITY) to exercise the general path
            self.velocities *= 1.0 / np.sqrt(1.0 + 0.01 * rng.random(self.velocities.shape))
            self.vxface *= 1.0 / np.sqrt(1.0 + 0.01 * rng.random(self.vxface.shape))
            self.vyface *= 1.0 / np.sqrt(1.0 + 0.01 * rng.random(self.vyface.shape))
        vx = self.velocities[: n3d * n4d].reshape(n4d, n3d)
        vy = self.velocities[n3d * n4d:].reshape(n4d, n3d)
        x = (np.arange(n1d) - ng + 0.5) * self.dx[0]
        y = (np.arange(n2d) - ng + 0.5) * self.dx[1]
        fx = 1.0 + 0.1 * np.cos(2 * np.pi * x / L[0])[None, :] * np.cos(2 * np.pi * y / L[1])[:, None]
        fv = np.exp(-0.5 * (vx ** 2 + vy ** 2)) / (2 * np.pi)
        f = fv[:, :, None, None] * fx[None, None, :, :]
        if rough:
            f = f * (1.0 + rough * rng.uniform(-1, 1, size=f.shape))
        self.f = np.ascontiguousarray(f, dtype=np.float64)   # C order (i4,i3,i2,i1) == Fortran (i1..i4)
        self.fx = np.ascontiguousarray(fx)
        self.fv = np.ascontiguousarray(fv)
        a = 0.05 * rng.uniform(-1, 1, size=(2, n2d, n1d))
        self.accel = np.ascontiguousarray(a * self.norm)     # already scaled by q/m
        self.em = np.ascontiguousarray(0.05 * rng.uniform(-1, 1, size=(6, n2d, n1d)))
        self.vz = np.ascontiguousarray(0.1 * rng.uniform(-1, 1, size=(n2d, n1d)))
        self.vel1 = np.zeros((n1d + 1) * n2d * n3d * n4d)
        self.vel2 = np.zeros((n2d + 1) * n3d * n4d * n1d)
        ok.ok_initialize_velocity(C.byref(self.g), self.velocities, self.vel1, self.vel2)

    def vel34(self, ok, maxwell=False):
        n1d, n2d, n3d, n4d = self.nd
        vel3 = np.zeros((n3d + 1) * n4d * n1d * n2d)
        vel4 = np.zeros((n4d + 1) * n1d * n2d * n3d)
        ax, ay = C.c_double(), C.c_double()
        if maxwell:
            ok.ok_set_phase_space_vel_maxwell_4d(vel3, vel4, C.byref(self.g), self.vxface, self.vyface, self.norm,
                                                 self.bz, self.em.ravel(), self.vz.ravel(), C.byref(ax), C.byref(ay))
        else:
            ok.ok_set_phase_space_vel_4d(vel3, vel4, C.byref(self.g), self.vxface, self.vyface, self.norm, self.bz,
                                         self.accel.ravel(), C.byref(ax), C.byref(ay))
        return vel3, vel4, ax.value, ay.value

    def ic_callback(self, fnorm=0.7, frac=0.9):
        n1d, n2d, n3d, n4d = self.nd
        fx, fv = self.fx, self.fv

        def cb(ctx, i1, i2, i3, i4):
            return fnorm * fv[i4, i3] * fx[i2, i1] * frac
        return IC_FN(cb)


class Dev:
    """Device-side mirror of a Setup, through torch tensors (memory plumbing only)."""

    def __init__(self, lkmod, s, maxwell=False):
        import torch
        import loki_b200 as lkm
        self.torch = torch
        self.s = s
        dev = torch.device("cuda:0")
        self.t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self.g = lkm.Geom.make(s.n, s.order, s.dx)
        self.f = self.t(s.f)
        self.velocities = self.t(s.velocities)
        self.vxface = self.t(s.vxface)
        self.vyface = self.t(s.vyface)
        self.accel_field = self.t(s.em if maxwell else s.accel)
        self.vz = self.t(s.vz)
        a = lkm.Accel()
        a.kind = 1 if maxwell else 0
        a.field = self.accel_field.data_ptr()
        a.vz = self.vz.data_ptr()
        a.vxface_velocities = self.vxface.data_ptr()
        a.vyface_velocities = self.vyface.data_ptr()
        a.normalization = s.norm
        a.bz_const = s.bz
        self.accel = a

    def zeros_like_f(self):
        return self.torch.zeros_like(self.f)

    def sync(self):
        self.torch.cuda.synchronize()


def rel_err(test, base):
    """max |t-b| / max|b| : field-level relative error"""
    scale = np.max(np.abs(base))
    return 0.0 if scale == 0 else float(np.max(np.abs(test - base)) / scale)


def cell_rel_err(test, base):
    """checkTests.C:345-358: per-cell |t-b|/|t| (|b| when t == 0), max over cells"""
    t = np.asarray(test).ravel()
    b = np.asarray(base).ravel()
    den = np.where(t != 0.0, np.abs(t), np.abs(b))
    diff = np.abs(t - b)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.where(den > 0, diff / den, 0.0)
    return float(np.max(r))


def star_rel_err(test, base, scale_from, ng):
    """max over interior cells of |t-b| / (largest |scale_from| in the cell's 4D star stencil of radius ng).
    One RK step updates a cell from its star neighbourhood, so this is the accuracy a cell can be held
    to when the data are rough enough for f + dt*rhs to cancel (where the per-cell metric of
    checkTests.C:345-358 is unbounded).  Arrays include the ghost layers."""
    a = np.abs(np.asarray(scale_from))
    m = a.copy()
    for ax in range(4):
        for k in range(1, ng + 1):
            m = np.maximum(m, np.roll(a, k, axis=ax))
            m = np.maximum(m, np.roll(a, -k, axis=ax))
    I = (slice(ng, -ng),) * 4
    d = np.abs(np.asarray(test)[I] - np.asarray(base)[I])
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.where(m[I] > 0, d / m[I], 0.0)
    return float(np.max(r))
