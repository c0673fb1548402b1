"""Deck-shaped problem set-ups: the numbers of the reference's regression decks (test/*/*.pp) with the
configuration-space and velocity grids adjustable, the analytic initial conditions they name
(PerturbedMaxwellianIC.C:95-289 factorable branch / ic_option 1, MaxwellianThermal.C:16-60) as factor
tables, and the descriptor the C++ host mirror (include/loki_b200_host.h) is created from.
Host-side set-up only (numpy, O(N^2) tables); no compute path.  Used by bench.py, __graft_entry__.smoke()
and the tests (tests/decks.py adds the oracle-side adapter)."""
import math

import numpy as np


class Species:
    def __init__(self, name, nv, vlim, mass, charge, tx=1.0, ty=1.0, A=0.0, B=0.0, Cc=0.0, kx1=0.5, ky1=0.5,
                 kx2=0.5, ky2=0.5, frac=1.0, driver=None, bz=0.0, vx0=0.0, vy0=0.0, x_wave_number=0.0,
                 y_wave_number=0.0, flow_phase=0.0, stream=None):
        # defaults as PerturbedMaxwellianIC's constructor sets them (PerturbedMaxwellianIC.C:44-61): the wave numbers 0.5
        self.name, self.nv, self.vlim, self.mass, self.charge = name, nv, vlim, mass, charge
        self.tx, self.ty, self.A, self.B, self.Cc = tx, ty, A, B, Cc
        self.kx1, self.ky1, self.kx2, self.ky2, self.frac = kx1, ky1, kx2, ky2, frac
        self.driver, self.bz = driver, bz
        # ShapedRampedCosineDriver.C:262-268, 293-304: phase in radians; shape_type 0 = "sin2", 1 = "exp"
        self.driver_phase, self.driver_shape_type = 0.0, 0
        # flow-velocity wave of the Perturbed Maxwellian (PerturbedMaxwellianIC.C:393-410): a non-zero vx0/vy0
        # makes the initial condition non-factorable (:395-397)
        self.vx0, self.vy0, self.x_wave_number, self.y_wave_number, self.flow_phase = vx0, vy0, x_wave_number, y_wave_number, flow_phase

        # options beyond the benchmark decks: a Krook layer and a pitch-angle collision operator (see Deck.apply_options)
        # ic.vflowinitx / ic.vflowinity: the constant drift of the Maxwellian (MaxwellianThermal.C:48-49)
        self.vflowinitx, self.vflowinity = 0.0, 0.0
        # the three variants of PerturbedMaxwellianIC (PerturbedMaxwellianIC.C:351-362): 1 "Perturbed Maxwellian",
        # 2 "Landau damping", 3 "Maxwellian with noise" (noise_amp / noise_phase per mode, :380-386); ic.spatial_phase (:411)
        self.ic_option, self.spatial_phase, self.noise_amp, self.noise_phase = 1, 0.0, (), ()
        self.krook, self.collision = None, None
        # "External 2D" initial condition (External2DIC.C): `external` = the file's "2D dist" dataset, the spatial factor
        # of the WHOLE configuration space with its ghost layers, (Ny + 2 ng, Nx + 2 ng); `external_frac` = ic.frac
        self.external, self.external_frac = None, 1.0
        # a twilight-zone source: Species.tz = dict(amp=, kind= 1 TrigTZSource | 2 ElectronTrigTZSource |
        # 3 TwoSpecies_ElectronTrigTZSource | 4 TwoSpecies_IonTrigTZSource[, electron_mass=, ion_mass=]) adds the
        # manufactured-solution forcing to the rhs
        self.tz = None
        # "Interpenetrating Stream" initial condition, half-plane syntax (InterpenetratingStreamIC.C:496-540):
        # dict(tl, tt, theta, d, beta, floor, frac, frac2, two_sided, centered)
        self.stream = stream

    @property
    def factorable(self):
        return self.vx0 == 0.0 and self.vy0 == 0.0


def driver_params(xwidth, ywidth, shape, omega, E0, t_ramp, t_off, x_shape, lwidth, x0, t0=0.0):
    """ShapedRampedCosineDriver parameter vector in enum order, old t_ramp/t_off syntax
    (ShapedRampedCosineDriver.C:306-311)."""
    p = [0.0] * 16
    p[0], p[1], p[2], p[3], p[4], p[5] = xwidth, ywidth, shape, omega, E0, t0
    p[6], p[7], p[8] = t_ramp, 0.0, t_off
    p[9], p[10], p[11], p[12], p[13] = x_shape, lwidth, x0, 0.0, 0.0
    return p


class Deck:
    def __init__(self, name, n, xlim, species, order=4, rk=4, cfl=0.9):  # Simulation.C:174
        self.name, self.n, self.xlim, self.species, self.order, self.rk, self.cfl = name, n, xlim, species, order, rk, cfl
        self.ng = 2 if order == 4 else 3
        self.dx = ((xlim[1] - xlim[0]) / n[0], (xlim[3] - xlim[2]) / n[1])
        # options beyond the benchmark decks: periodic_dir (ProblemDomain), use_new_bcs (VPSystem.C:819-821); a species'
        # Krook layer is Species.krook = dict(x1a=, x1b=, x2a=, x2b=, coefficient=) with absent ends = the domain's
        self.periodic = (True, True)
        self.use_new_bcs = False
        # probe locations as fractions of the domain (Simulation.C:393-412; the default's y fraction is 0 there)
        self.probes = [(0.5, 0.0)]

    def krook_nu(self, sp, tile_lo=(0, 0), tile_n=None):
        """KrookLayer::initialize (KrookLayer.C:54-160): nu (n2d, n1d) of a tile, ghosts zero; None without a layer"""
        kr = getattr(sp, "krook", None)
        if not kr:
            return None
        tile_n = tile_n or self.n
        ng = self.ng
        xmin, xmax, ymin, ymax = self.xlim
        dx, dy = self.dx
        xlo, xhi = kr.get("x1a", xmin), kr.get("x1b", xmax)
        ylo, yhi = kr.get("x2a", ymin), kr.get("x2b", ymax)
        coef = kr.get("coefficient", 1.0)

        def ramp(e):
            if self.order == 4:
                return -pow(e, 4) * (+20.0 * pow(e, 3) - 70.0 * pow(e, 2) + 84.0 * e - 35.0)
            return -pow(e, 6) * (+252.0 * pow(e, 5) - 1386.0 * pow(e, 4) + 3080.0 * pow(e, 3) - 3465.0 * pow(e, 2) + 1980.0 * e - 462.0)

        def frac(c, lo, hi, cmin, cmax):
            if c < lo:
                return (c - lo) / (cmin - lo)
            if c > hi:
                return (c - hi) / (cmax - hi)
            return 0.0

        nu = np.zeros((tile_n[1] + 2 * ng, tile_n[0] + 2 * ng))
        for j in range(tile_n[1]):
            i2 = tile_lo[1] + j
            nuy = ramp(frac(ymin + (0.5 + i2) * dy, ylo, yhi, ymin, ymax))
            for i in range(tile_n[0]):
                i1 = tile_lo[0] + i
                nux = ramp(frac(xmin + (0.5 + i1) * dx, xlo, xhi, xmin, xmax))
                nu[j + ng, i + ng] = coef * ((1.0 - nuy) * nux + nuy)
        return nu

    def apply_options(self, H, sys_, tile_lo=(0, 0), tile_n=None):
        """hand the deck's boundary options and Krook layers to a lk_vp_system (after lk_vp_set_inflow)"""
        st = H.lk_vp_set_boundary_options(sys_, int(not self.periodic[0]), int(not self.periodic[1]), int(self.use_new_bcs))
        for s, sp in enumerate(self.species):
            nu = self.krook_nu(sp, tile_lo, tile_n)
            if st == 0 and nu is not None:
                st = H.lk_vp_set_krook(sys_, s, nu.ctypes.data)
            co = getattr(sp, "collision", None)
            if st == 0 and co:
                # a "Pitch Angle Collision Operator" (PitchAngleCollisionOperator.C): Species.collision = dict(range_lo=,
                # range_hi=, vfloor=, vthermal_dt=, nu_coef=, conservative=)
                from .capi import PitchAngle
                pa = PitchAngle.make(co["range_lo"], co["range_hi"], co["vfloor"], co["vthermal_dt"], co["nu_coef"],
                                     co.get("conservative", 1))
                st = H.lk_vp_set_pitch_angle(sys_, s, pa)
            tz = getattr(sp, "tz", None)
            if st == 0 and tz:
                st = H.lk_vp_set_trig_tz(sys_, s, int(tz.get("kind", 1)), float(tz["amp"]), float(tz.get("electron_mass", 1.0)),
                                         float(tz.get("ion_mass", 1.0)))
        return st

    def geom_of(self, sp):
        dvx = (sp.vlim[1] - sp.vlim[0]) / sp.nv[0]
        dvy = (sp.vlim[3] - sp.vlim[2]) / sp.nv[1]
        return (self.n[0], self.n[1], sp.nv[0], sp.nv[1]), (self.dx[0], self.dx[1], dvx, dvy)

    def _periodic_image(self, x1, x2, Lx, Ly):
        """a cell centre outside a PERIODIC direction's limits is moved by the box length; a non-periodic direction's
        ghost cells keep their own coordinates (PerturbedMaxwellianIC.C:123-143, External2DIC.C:168-187,
        InterpenetratingStreamIC.C:153-172)"""
        per = getattr(self, "periodic", (True, True))
        if per[0]:
            x1 = np.where(x1 < self.xlim[0], x1 + Lx, np.where(x1 > self.xlim[1], x1 - Lx, x1))
        if per[1]:
            x2 = np.where(x2 < self.xlim[2], x2 + Ly, np.where(x2 > self.xlim[3], x2 - Ly, x2))
        return x1, x2

    def ic_tables(self, sp, tile_lo=(0, 0), tile_n=None):
        """fx (n2d,n1d), fv (n4d,n3d), fnorm -- PerturbedMaxwellianIC::cache, factorable branch, the three ic_options"""
        ng = self.ng
        tile_n = tile_n or self.n
        n, dx = self.geom_of(sp)
        xlo, ylo = self.xlim[0], self.xlim[2]
        Lx, Ly = self.n[0] * dx[0], self.n[1] * dx[1]
        i1 = np.arange(-ng, tile_n[0] + ng) + tile_lo[0]
        i2 = np.arange(-ng, tile_n[1] + ng) + tile_lo[1]
        x1 = xlo + (i1 + 0.5) * dx[0]
        x2 = ylo + (i2 + 0.5) * dx[1]
        x1, x2 = self._periodic_image(x1, x2, Lx, Ly)
        phi = float(getattr(sp, "spatial_phase", 0.0))
        option = int(getattr(sp, "ic_option", 1))
        if option == 1:      # PerturbedMaxwellianIC.C:145-150
            fx = (1.0 + sp.A * np.cos(sp.kx1 * x1 + phi)[None, :] * np.cos(sp.ky1 * x2 + phi)[:, None] +
                  sp.B * np.cos(sp.kx2 * x1 + phi)[None, :] + sp.Cc * np.cos(sp.ky2 * x2 + phi)[:, None])
        elif option == 2:    # "Landau damping" (:151-154)
            fx = 1.0 + sp.A * np.cos((sp.kx1 * x1)[None, :] + (sp.ky1 * x2)[:, None] + phi)
        elif option == 3:    # "Maxwellian with noise" (:155-162): a sum over the modes of the box length, in mode order
            pi = 4.0 * math.atan(1.0)
            fx1 = np.ones_like(x1)
            for k in range(1, len(sp.noise_amp) + 1):
                fx1 = fx1 + sp.noise_amp[k - 1] * np.cos(2.0 * pi * k * (x1 + sp.noise_phase[k - 1]) / Lx + phi)
            fx = np.broadcast_to(fx1[None, :], (x2.size, x1.size)).copy()
        else:
            raise ValueError("ic_option %r" % option)
        x3 = sp.vlim[0] + (np.arange(-ng, sp.nv[0] + ng) + 0.5) * dx[2]
        x4 = sp.vlim[2] + (np.arange(-ng, sp.nv[1] + ng) + 0.5) * dx[3]
        thx, thy = sp.tx / sp.mass, sp.ty / sp.mass
        # MaxwellianThermal::thermalFactor(0, x3, x4) (MaxwellianThermal.C:40-59): the drift is m_x0 * 0 + m_flowinitx
        d3 = x3 - (sp.vx0 * 0.0 + getattr(sp, "vflowinitx", 0.0))
        d4 = x4 - (sp.vy0 * 0.0 + getattr(sp, "vflowinity", 0.0))
        fv = np.exp(-0.5 * ((d3 ** 2)[None, :] / thx + (d4 ** 2)[:, None] / thy))
        fnorm = sp.mass / (2.0 * math.pi * math.sqrt(sp.tx * sp.ty))
        if option == 3:
            # getIC_At_Pt multiplies m_fx * fnorm * m_fv * m_frac for this variant (PerturbedMaxwellianIC.C:276-278, and
            # :240 for the cached m_f) instead of fnorm * m_fv * m_fx * m_frac: the first product is taken here and
            # fnorm handed on as 1, so that every consumer's fnorm * fv * fx * frac (host and device, lk_device.cuh
            # inflow kind 1) has the reference's bits -- 1 * fv is exact and IEEE products commute
            fx, fnorm = fx * fnorm, 1.0
        if getattr(sp, "external", None) is not None:
            # External2DIC::cache, factorable branch (External2DIC.C:147-198): m_fx(i1, i2) = the file's value at the cell,
            # a ghost cell of a periodic direction taking its periodic image; getIC_At_Pt multiplies
            # m_frac * fnorm * m_fv * m_fx in this order (:282-284), i.e. the Perturbed Maxwellian's product with
            # fnorm -> frac * fnorm and frac -> 1 (Species.frac stays 1 for such a species)
            ext = np.asarray(sp.external, dtype=np.float64)
            if ext.shape != (self.n[1] + 2 * ng, self.n[0] + 2 * ng):
                raise ValueError("Distribution size does not match configuration space.")      # External2DIC.C:101-103
            idx1 = np.where(i1 < 0, i1 + self.n[0], np.where(i1 > self.n[0] - 1, i1 - self.n[0], i1)) if self.periodic[0] else i1
            idx2 = np.where(i2 < 0, i2 + self.n[1], np.where(i2 > self.n[1] - 1, i2 - self.n[1], i2)) if self.periodic[1] else i2
            fx = ext[(idx2 + ng)[:, None], (idx1 + ng)[None, :]]
            fnorm = sp.external_frac * fnorm
        return np.ascontiguousarray(fx), np.ascontiguousarray(fv), fnorm

    def stream_tables(self, sp, tile_lo=(0, 0), tile_n=None):
        """InterpenetratingStreamIC::cache, half-plane syntax (InterpenetratingStreamIC.C:112-252):
        fx, fx2 (n2d,n1d) and fv, fv2 (n4d,n3d) with fnorm folded into fv; fx2 / fv2 None when unused"""
        st = sp.stream
        ng = self.ng
        tile_n = tile_n or self.n
        n, dx = self.geom_of(sp)
        Lx, Ly = self.n[0] * dx[0], self.n[1] * dx[1]
        x1 = self.xlim[0] + (np.arange(-ng, tile_n[0] + ng) + tile_lo[0] + 0.5) * dx[0]
        x2 = self.xlim[2] + (np.arange(-ng, tile_n[1] + ng) + tile_lo[1] + 0.5) * dx[1]
        x1, x2 = self._periodic_image(x1, x2, Lx, Ly)
        erf = np.vectorize(math.erf)
        th, beta, floor = st["theta"], st["beta"], st.get("floor", 0.0)
        xi0 = -st["d"]
        xi = x1[None, :] * math.cos(th) + x2[:, None] * math.sin(th)
        fx2 = None
        if st.get("two_sided"):
            fx = floor / 2.0 + (st["frac"] - floor) * 0.5 * (1.0 + erf(-beta * (xi - xi0)))
            fx2 = floor / 2.0 + (st["frac2"] - floor) * 0.5 * (1.0 + erf(beta * (xi + xi0)))
        elif st.get("centered"):
            fx = floor + (math.sqrt(st["frac"]) - floor) * 0.5 * (1.0 + erf(-beta * (xi - xi0)))
            fx2 = floor + (math.sqrt(st["frac"]) - floor) * 0.5 * (1.0 + erf(beta * (xi + xi0)))
        else:
            fx = floor + (st["frac"] - floor) * 0.5 * (1.0 + erf(-beta * (xi - xi0)))
        x3 = sp.vlim[0] + (np.arange(-ng, sp.nv[0] + ng) + 0.5) * dx[2]
        x4 = sp.vlim[2] + (np.arange(-ng, sp.nv[1] + ng) + 0.5) * dx[3]
        vl = x3[None, :] * math.cos(th) + x4[:, None] * math.sin(th)
        vt = -x3[None, :] * math.sin(th) + x4[:, None] * math.cos(th)
        thl, tht = st["tl"] / sp.mass, st["tt"] / sp.mass
        fnorm = sp.mass / (2.0 * math.pi * math.sqrt(st["tl"] * st["tt"]))
        # vl0 = vt0 = 0 in the decks: both thermal objects are the same Maxwellian (MaxwellianThermal.C:40-59)
        fv = fnorm * np.exp(-0.5 * ((vl - 0.0) * (vl - 0.0) / thl + (vt - 0.0) * (vt - 0.0) / tht))
        fv2 = fv.copy() if st.get("two_sided") else None
        c = np.ascontiguousarray
        return c(fx), (c(fx2) if fx2 is not None else None), c(fv), (c(fv2) if fv2 is not None else None)

    def inflow_kind(self, sp):
        if sp.stream is None:
            return 1 if sp.factorable else 3
        return 2 if sp.stream.get("two_sided") else (4 if sp.stream.get("centered") else 1)

    def initial_state(self, sp, tile_lo=(0, 0), tile_n=None):
        if sp.stream is not None:
            # getIC_At_Pt (InterpenetratingStreamIC.C:265-286); returns fnorm = 1 (already inside fv)
            fx, fx2, fv, fv2 = self.stream_tables(sp, tile_lo, tile_n)
            X, V = fx[None, None, :, :], fv[:, :, None, None]
            if sp.stream.get("two_sided"):
                f = X * V + fx2[None, None, :, :] * fv2[:, :, None, None]
            elif sp.stream.get("centered"):
                f = (V * X) * fx2[None, None, :, :]
            else:
                f = V * X
            return np.ascontiguousarray(f), fx, fv, 1.0
        fx, fv, fnorm = self.ic_tables(sp, tile_lo, tile_n)
        if not sp.factorable:
            return self.initial_state_full(sp, fx, fnorm), fx, fv, fnorm
        # getIC_At_Pt: fnorm*fv*fx*frac in this order (PerturbedMaxwellianIC.C:279-281)
        f = ((fnorm * fv)[:, :, None, None] * fx[None, None, :, :]) * sp.frac
        return np.ascontiguousarray(f), fx, fv, fnorm

    def initial_state_full(self, sp, fx, fnorm):
        """PerturbedMaxwellianIC::cache, non-factorable branch (PerturbedMaxwellianIC.C:176-246): the drift
        of the Maxwellian follows cos(x_wave_number x + y_wave_number y + phase); whole configuration space"""
        ng = self.ng
        n, dx = self.geom_of(sp)
        Lx, Ly = self.n[0] * dx[0], self.n[1] * dx[1]
        x1 = self.xlim[0] + (np.arange(-ng, self.n[0] + ng) + 0.5) * dx[0]
        x2 = self.xlim[2] + (np.arange(-ng, self.n[1] + ng) + 0.5) * dx[1]
        x1, x2 = self._periodic_image(x1, x2, Lx, Ly)
        x3 = sp.vlim[0] + (np.arange(-ng, sp.nv[0] + ng) + 0.5) * dx[2]
        x4 = sp.vlim[2] + (np.arange(-ng, sp.nv[1] + ng) + 0.5) * dx[3]
        sf = np.cos(sp.x_wave_number * x1[None, :] + sp.y_wave_number * x2[:, None] + sp.flow_phase)  # (n2d,n1d)
        thx, thy = sp.tx / sp.mass, sp.ty / sp.mass
        c3 = sp.vx0 * sf + getattr(sp, "vflowinitx", 0.0)   # m_x0*spatial_factor + m_flowinitx (MaxwellianThermal.C:48-49)
        c4 = sp.vy0 * sf + getattr(sp, "vflowinity", 0.0)
        d3 = x3[None, :, None, None] - c3[None, None, :, :]
        d4 = x4[:, None, None, None] - c4[None, None, :, :]
        fv = np.exp(-0.5 * ((d3 * d3) / thx + (d4 * d4) / thy))
        return np.ascontiguousarray(((fnorm * fv) * fx[None, None, :, :]) * sp.frac)

    def inflow_ghost_tables(self, f_ic):
        """velocity-ghost layers of a cached IC array in the layout of lk_inflow kind 3:
        ghost3 (2ng,n4d -> stored [n4d][2ng][n2d][n1d]), ghost4 ([2ng][n3d][n2d][n1d])"""
        ng = self.ng
        g3 = np.concatenate([f_ic[:, :ng], f_ic[:, -ng:]], axis=1)
        g4 = np.concatenate([f_ic[:ng], f_ic[-ng:]], axis=0)
        return np.ascontiguousarray(g3), np.ascontiguousarray(g4)

    # ---- product side ----
    def set_inflow(self, H, sys_, s, tile_lo=(0, 0), tile_n=None):
        """hand species s's inflow (initial-condition) tables to a lk_vp_system in the form its IC class has"""
        sp = self.species[s]
        kind = self.inflow_kind(sp)
        if sp.stream is not None and kind in (2, 4):
            fx, fx2, fv, fv2 = self.stream_tables(sp, tile_lo, tile_n)
            return H.lk_vp_set_inflow2(sys_, s, kind, fx.ctypes.data, fv.ctypes.data, fx2.ctypes.data,
                                       fv2.ctypes.data if fv2 is not None else None)
        if sp.stream is not None:
            fx, fx2, fv, fv2 = self.stream_tables(sp, tile_lo, tile_n)
            return H.lk_vp_set_inflow(sys_, s, fx.ctypes.data, fv.ctypes.data, 1.0, 1.0)
        if kind == 3:
            # the cached m_f of a drifting Maxwellian (vx0 / vy0 != 0) at the velocity ghosts: lk_vm_set_inflow_ghosts
            # carries it for a Vlasov-Maxwell system; the Vlasov-Poisson host mirror has no such entry point, and the
            # factored tables would put a different Maxwellian into the velocity ghosts
            raise ValueError("species %r: a non-factorable initial condition (ic.vx0 / ic.vy0 != 0) is not supported "
                             "in a Vlasov-Poisson system" % sp.name)
        fx, fv, fnorm = self.ic_tables(sp, tile_lo, tile_n)
        return H.lk_vp_set_inflow(sys_, s, fx.ctypes.data, fv.ctypes.data, fnorm, sp.frac)

    def product_desc(self, tile_lo=(0, 0), tile_n=None, ntiles=1):
        from .host import SpeciesDesc, VPDesc
        tile_n = tile_n or self.n
        sd = (SpeciesDesc * len(self.species))()
        for k, sp in enumerate(self.species):
            sd[k].nv[0], sd[k].nv[1] = sp.nv
            sd[k].vlo[0], sd[k].vlo[1] = sp.vlim[0], sp.vlim[2]
            sd[k].vhi[0], sd[k].vhi[1] = sp.vlim[1], sp.vlim[3]
            sd[k].mass, sd[k].charge, sd[k].bz_const = sp.mass, sp.charge, sp.bz
            sd[k].has_driver = 1 if sp.driver else 0
            if sp.driver:
                for j in range(16):
                    sd[k].driver[j] = sp.driver[j]
            sd[k].driver_phase = float(getattr(sp, "driver_phase", 0.0))
            sd[k].driver_shape_type = int(getattr(sp, "driver_shape_type", 0))
        d = VPDesc()
        d.nspecies, d.species, d.order, d.rk_order = len(self.species), sd, self.order, self.rk
        d.nglobal[0], d.nglobal[1] = self.n
        d.xlo[0], d.xlo[1] = self.xlim[0], self.xlim[2]
        d.xhi[0], d.xhi[1] = self.xlim[1], self.xlim[3]
        d.tile_lo[0], d.tile_lo[1] = tile_lo
        d.tile_n[0], d.tile_n[1] = tile_n
        d.ntiles = ntiles
        d._keep = sd
        return d


PI = 3.1415926535897932384626


def perl15(x):
    """deck constants reach Loki through Perl string interpolation: 15 significant digits
    (LokiParser.C:104-118).  Perl keeps the constants themselves in full precision; only their use in a
    parameter line rounds, so the mirrors below compute in full precision and round at the point of use
    (loki_b200.pp does the same from the deck text; tests compare the two on the reference's own decks)."""
    return float("%.15g" % x)


P = perl15


def plane_epw(n=(32, 32), nv=(128, 32), A=0.0, ky1=None):
    """test/planeEPW_fixedIons/planeEPW_fixedIons.pp: one electron species, driven, order 4 / RK4.
    The deck has A = 0 (spatially uniform start); tests that switch the spatial mode on pass a ky1 that is
    periodic over the box."""
    xa, xb, ya, yb = -3 * PI, 3 * PI, -78 * PI, 78 * PI
    drv = driver_params(xwidth=P(3 * PI), ywidth=P(144 * PI), shape=0.0, omega=1.2001, E0=0.01, t_ramp=10.0,
                        t_off=100.0, x_shape=0.0, lwidth=50.0, x0=0.0)
    e = Species("electron", nv, (-7.0, 7.0, -7.0, 7.0), 1.0, -1.0, A=A, kx1=P(1.0 / 3), ky1=P(1.0 / 12) if ky1 is None else ky1, driver=drv)
    return Deck("planeEPW_fixedIons", n, (P(xa), P(xb), P(ya), P(yb)), [e], order=4, rk=4, cfl=1.0)


def plane_iaw(n=(32, 32), nv=(64, 32), order=4, rk=4, A=0.0, ky1=None):
    """test/planeIAW/planeIAW.pp (order 4 / RK4) and test/planeIAW_6 (order 6 / RK6): electrons + ions"""
    klde = 1.0 / 3
    ialpha = math.sqrt(10.0) * math.sqrt(100.0)
    xa, xb, ya, yb = -PI / klde, PI / klde, -78 * PI / klde, 78 * PI / klde
    drv = driver_params(xwidth=P((xb - xa) / 2.0), ywidth=P(300 * PI), shape=0.0, omega=0.0381, E0=0.1, t_ramp=1.0,
                        t_off=2.0, x_shape=0.0, lwidth=50.0, x0=0.0)
    e = Species("electron", nv, (-7.0, 7.0, -7.0, 7.0), 1.0, -1.0, A=A, kx1=P(klde), ky1=P(klde) if ky1 is None else ky1, driver=drv)
    vi = P(10 / ialpha)
    i = Species("ion", nv, (-vi, vi, -vi, vi), 100.0, 1.0, tx=0.1, ty=0.1, A=A, kx1=P(klde), ky1=P(klde) if ky1 is None else ky1)
    return Deck("planeIAW" + ("_6" if order == 6 else ""), n, (P(xa), P(xb), P(ya), P(yb)), [e, i], order=order, rk=rk, cfl=1.0)


def interpenetrating_streams(n=(128, 7), nv=(24, 16), order=6, rk=6):
    """test/InterpenetratingStreams/InterpenetratingStreams.pp: electrons, He, C; order 6 / RK6, cfl 0.95;
    y limits scale with Ny so that dx = dy as in the deck"""
    mp_over_me = 1836.0
    m_he, m_c = 4.0 * mp_over_me, 12.0 * mp_over_me
    vth_he, vth_c = math.sqrt(1.0 / m_he), math.sqrt(1.0 / m_c)
    xa, xb = -62.5, 62.5
    dx = (xb - xa) / n[0]
    ya, yb = -0.5 * n[1] * dx, 0.5 * n[1] * dx
    common = dict(tl=1.0, tt=1.0, theta=0.0, d=31.25)
    e = Species("electron", nv, (-7.5, 7.5, -7.5, 7.5), 1.0, -1.0,
                stream=dict(common, beta=0.768, floor=0.05, frac=10.0, two_sided=True, frac2=10.0))
    lim_he = (P(-7.5 * vth_he), P(7.5 * vth_he), P(-7.5 * vth_he), P(7.5 * vth_he))
    he = Species("He", nv, lim_he, P(m_he), 2.0, stream=dict(common, beta=-0.768, floor=0.0, frac=0.025, centered=True))
    lim_c = (P(-7.5 * vth_c), P(7.5 * vth_c), P(-7.5 * vth_c), P(7.5 * vth_c))
    cfrac = P(5.0 / 3.0)
    c = Species("C", nv, lim_c, P(m_c), 6.0, stream=dict(common, beta=0.768, floor=0.0, frac=cfrac, two_sided=True, frac2=cfrac))
    return Deck("InterpenetratingStreams", n, (P(xa), P(xb), P(ya), P(yb)), [e, he, c], order=order, rk=rk, cfl=0.95)


def pitch_angle_collisions(n=(32, 5), nv=(64, 64), order=4, rk=4, A=0.0001, conservative=1):
    """test/pitchAngleCollisions/pitchAngleCollisions.pp: one electron species with a "Pitch Angle Collision Operator"
    (PitchAngleCollisionOperator.C), no driver, order 4 / RK4, cfl 0.8, six probes"""
    klde = 0.3
    e = Species("electron", nv, (-7.0, 7.0, -7.0, 7.0), 1.0, -1.0, A=A, kx1=klde, ky1=klde)
    e.collision = dict(range_lo=(-5.25, -5.25), range_hi=(5.25, 5.25), vfloor=0.01, vthermal_dt=1.0, nu_coef=0.1,
                       conservative=conservative)
    d = Deck("pitchAngleCollisions", n, (P(-PI / klde), P(PI / klde), -0.1, 0.1), [e], order=order, rk=rk, cfl=0.8)
    d.probes = [(0.5, 0.5), (0.25, 0.5), (0.75, 0.5), (0.95, 0.5), (0.5, 0.25), (0.5, 0.75)]
    return d


class VMDeck(Deck):
    """a Vlasov-Maxwell deck: Deck + light speed, Maxwell hyper-dissipation and the field initial conditions
    (SimpleEMIC / SimpleVELIC single waves)"""

    def __init__(self, name, n, xlim, species, light_speed, av_weak, av_strong, em_ics, vel_ics, order=4, cfl=0.9, rk=4):
        Deck.__init__(self, name, n, xlim, species, order=order, rk=rk, cfl=cfl)
        self.light_speed, self.av_weak, self.av_strong = light_speed, av_weak, av_strong
        self.em_ics, self.vel_ics = em_ics, vel_ics

    def initial_fields(self):
        """em_vars (6,n2d,n1d) and vz per species (n2d,n1d): setsimpleemic / setsimplevelic over the whole
        data box, ghost cells included (SimpleEMICF.f:10-47, SimpleVELICF.f:10-40)"""
        ng = self.ng
        x1 = self.xlim[0] + (np.arange(-ng, self.n[0] + ng) + 0.5) * self.dx[0]
        x2 = self.xlim[2] + (np.arange(-ng, self.n[1] + ng) + 0.5) * self.dx[1]
        em = np.zeros((6, x2.size, x1.size))
        for ic in self.em_ics:
            env = np.cos(ic["kx"] * x1[None, :] + ic["ky"] * x2[:, None] + ic["phase"])
            o = 0 if ic["field"] == "E" else 3
            em[o] = em[o] + ic["xamp"] * env
            em[o + 1] = em[o + 1] + ic["yamp"] * env
            em[o + 2] = em[o + 2] + ic["zamp"] * env
        vz = []
        for k in range(len(self.species)):
            v = np.zeros((x2.size, x1.size))
            if k < len(self.vel_ics):
                ic = self.vel_ics[k]
                v = v + ic["amp"] * np.cos(ic["kx"] * x1[None, :] + ic["ky"] * x2[:, None] + ic["phase"])
            vz.append(np.ascontiguousarray(v))
        return np.ascontiguousarray(em), vz

    def product_vm_desc(self):
        from .host import VMDesc
        d = VMDesc()
        d.base = self.product_desc()
        d._keep = d.base._keep if hasattr(d.base, "_keep") else None
        d.light_speed, d.av_weak, d.av_strong = self.light_speed, self.av_weak, self.av_strong
        return d


def em_damping(n=(32, 5), nv=(64, 64), order=4, rk=4):
    """test/emDamping/emDamping.pp: one electron species, Vlasov-Maxwell, order 4 / RK4, cfl 0.8"""
    omega, clight = 3.16, 22.36
    klde = math.sqrt(omega ** 2 - 1) / clight
    Ey = 1.0e-4
    Bz = klde * Ey / omega
    uy = -Ey / omega
    av_strong = 1.6 / clight
    xa, xb = -PI / klde, PI / klde
    e = Species("electron", nv, (-7.0, 7.0, -7.0, 7.0), 1.0, -1.0, kx1=0.0, ky1=0.0, vy0=P(uy), x_wave_number=P(klde),
                flow_phase=P(PI / 2.0))
    em_ics = [dict(field="E", xamp=0.0, yamp=P(Ey), zamp=0.0, kx=P(klde), ky=0.0, phase=0.0),
              dict(field="B", xamp=0.0, yamp=0.0, zamp=P(Bz), kx=P(klde), ky=0.0, phase=0.0)]
    vel_ics = [dict(amp=0.0, kx=0.0, ky=0.0, phase=0.0)]
    return VMDeck("emDamping", n, (P(xa), P(xb), -10.0, 10.0), [e], P(clight), 0.0, P(av_strong), em_ics, vel_ics, order=order,
                  cfl=0.8, rk=rk)
