"""Run a `.pp` deck on one GPU through the host mirror: `python -m loki_b200.run deck.pp [--max-steps N]`.

The loop is Simulation::advance (Simulation.C:302-351): dt = cfl * stableDt snapped to the next multiple of
save_times (selectTimeStep, :464-485), System::advance, and a time-history record every sequence_write_times
(VPSystem::accumulateSequences / the Maxwell analogue) printed as a JSON line.  Vlasov-Poisson (RK4/RK6) and
Vlasov-Maxwell (RK4) decks with the initial conditions and the driver the five benchmark decks use.  No
restart / plot files (HDF5 is out of scope)."""
import argparse
import ctypes as C
import json
import math
import sys

import numpy as np

from . import capi, host, pp


def select_dt(time, dt_stable_cfl, last_save, save_times, max_time):
    """Simulation::selectTimeStep (Simulation.C:464-485)"""
    final_time = min((last_save + 1) * save_times, max_time)
    remaining = final_time - time
    if dt_stable_cfl <= remaining:
        return remaining / int(math.ceil(remaining / dt_stable_cfl))
    return remaining * (1.0 + 10 * np.finfo(float).eps)


class Runner:
    def __init__(self, deck, stream=None):
        self.deck, self.L, self.H = deck, capi.load(), host.lib()
        if self.L.lk_device_count() < 1:
            raise capi.LokiError("no CUDA device: the hot path has no CPU fallback")
        self.vm = hasattr(deck, "light_speed")
        self.sys = C.c_void_p()
        H = self.H
        ns = len(deck.species)
        if self.vm:
            self.desc = deck.product_vm_desc()
            capi.check(H.lk_vm_create(C.byref(self.sys), C.byref(self.desc), stream), "lk_vm_create")
        else:
            self.desc = deck.product_desc()
            capi.check(H.lk_vp_create(C.byref(self.sys), C.byref(self.desc), stream), "lk_vp_create")
        self.shapes = []
        for s, sp in enumerate(deck.species):
            f, fx, fv, fnorm = deck.initial_state(sp)
            self.shapes.append(f.shape)
            if self.vm:
                capi.check(H.lk_vm_set_state(self.sys, s, f.ctypes.data), "lk_vm_set_state")
                if sp.factorable:
                    capi.check(H.lk_vm_set_inflow(self.sys, s, fx.ctypes.data, fv.ctypes.data, fnorm, sp.frac), "inflow")
                else:
                    g3, g4 = deck.inflow_ghost_tables(f)
                    capi.check(H.lk_vm_set_inflow_ghosts(self.sys, s, g3.ctypes.data, g4.ctypes.data), "inflow")
            else:
                capi.check(H.lk_vp_set_state(self.sys, s, f.ctypes.data), "lk_vp_set_state")
                capi.check(deck.set_inflow(H, self.sys, s), "inflow")
        if not self.vm:
            capi.check(deck.apply_options(H, self.sys), "boundary options / Krook layers")
        if self.vm:
            em, vz = deck.initial_fields()
            capi.check(H.lk_vm_set_fields(self.sys, em.ctypes.data), "lk_vm_set_fields")
            for s in range(ns):
                capi.check(H.lk_vm_set_vz(self.sys, s, vz[s].ctypes.data), "lk_vm_set_vz")
        self.time, self.last_save, self.last_seq, self.step = 0.0, 0, 0, 0
        self._seed()

    def _seed(self):
        """the throw-away evalRHS that seeds lambda_max (VPSystem.C:227-230, VMSystem.C:262-265)"""
        L, H = self.L, self.H
        bufs = []

        def dev(count):
            p = C.c_void_p()
            capi.check(L.lk_malloc(C.byref(p), 8 * count), "lk_malloc")
            bufs.append(p)
            return p
        ns = len(self.deck.species)
        rhs = (C.c_void_p * ns)(*[dev(int(np.prod(sh))) for sh in self.shapes])
        if self.vm:
            n2d, n1d = self.shapes[0][2], self.shapes[0][3]
            rvz = (C.c_void_p * ns)(*[dev(n1d * n2d) for _ in range(ns)])
            capi.check(H.lk_vm_eval_rhs(self.sys, rhs, dev(6 * n1d * n2d), rvz, 0.0), "lk_vm_eval_rhs")
        else:
            capi.check(H.lk_vp_eval_rhs(self.sys, rhs, 0.0), "lk_vp_eval_rhs")
        L.lk_sync(None)
        for p in bufs:
            L.lk_free(p)

    def history(self):
        ns = len(self.deck.species)
        out = np.zeros((12 + 5 * ns) if self.vm else (5 + 6 * ns))
        fn = self.H.lk_vm_time_history if self.vm else self.H.lk_vp_time_history
        got = C.c_int()
        capi.check(fn(self.sys, out.ctypes.data, out.size, C.byref(got)), "time_history")
        assert got.value == out.size
        return out

    def probes(self):
        """the probe time histories (Poisson.C:852-887): (Ex, Ey) of the last field solve at every probe of the deck
        (Simulation.C:393-412; default: one probe at the fractions (0.5, 0)); Vlasov-Poisson systems"""
        if self.vm:
            raise NotImplementedError("probe histories of the Vlasov-Maxwell system")
        locs = getattr(self.deck, "probes", None) or [(0.5, 0.0)]
        fx = np.array([p[0] for p in locs], dtype=np.float64)
        fy = np.array([p[1] for p in locs], dtype=np.float64)
        out = np.zeros(2 * len(locs))
        capi.check(self.H.lk_vp_probe_history(self.sys, len(locs), fx.ctypes.data, fy.ctypes.data, out.ctypes.data), "probe_history")
        return out.reshape(len(locs), 2)

    def flux_history(self):
        """the `*_flux` time histories (KineticSpecies.C:2052-2097): per species the kinetic-energy flux through the
        eight phase-space boundaries, [8 s + 2 dir + side]; Vlasov-Poisson systems"""
        if self.vm:
            raise NotImplementedError("flux histories of the Vlasov-Maxwell system")
        out = np.zeros(8 * len(self.deck.species))
        got = C.c_int()
        capi.check(self.H.lk_vp_flux_history(self.sys, out.ctypes.data, out.size, C.byref(got)), "flux_history")
        assert got.value == out.size
        return out

    def advance(self):
        H, run = self.H, self.deck.run
        dt = C.c_double()
        capi.check((H.lk_vm_stable_dt if self.vm else H.lk_vp_stable_dt)(self.sys, C.byref(dt)), "stable_dt")
        step = select_dt(self.time, self.deck.cfl * dt.value, self.last_save, run["save_times"], run["final_time"])
        if self.vm:
            capi.check(H.lk_vm_set_time(self.sys, self.time), "set_time")
            capi.check(H.lk_vm_advance(self.sys, step), "lk_vm_advance")
        else:
            capi.check(H.lk_vp_set_time(self.sys, self.time), "set_time")
            capi.check(H.lk_vp_advance(self.sys, step), "lk_vp_advance")
        self.time += step
        self.step += 1
        # exact comparisons as in Simulation::advance (Simulation.C:318-327): when rounding leaves the time one ulp
        # short of a save time, selectTimeStep's remaining * (1 + 10 eps) micro-step closes the gap next
        self.record = False
        if self.time >= (self.last_seq + 1) * run.get("sequence_write_times", 1.0):
            self.record = True          # accumulateSequences()
            self.last_seq += 1
        if self.time >= (self.last_save + 1) * run["save_times"]:
            self.last_save += 1         # writePlotFile()
        return step

    def done(self):
        """!Simulation::notDone (Simulation.H:126-129)"""
        run = self.deck.run
        return not (self.step < run["max_step"] and self.time < run["final_time"])

    def state(self, s):
        out = np.empty(self.shapes[s])
        capi.check((self.H.lk_vm_get_state if self.vm else self.H.lk_vp_get_state)(self.sys, s, out.ctypes.data), "get_state")
        return out

    def close(self):
        if self.sys:
            (self.H.lk_vm_destroy if self.vm else self.H.lk_vp_destroy)(self.sys)
            self.sys = C.c_void_p()


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("deck")
    ap.add_argument("--max-steps", type=int, default=None)
    ap.add_argument("--final-time", type=float, default=None)
    ap.add_argument("--every-step", action="store_true", help="emit the time histories after every step, not every sequence_write_times")
    a = ap.parse_args(argv)
    deck = pp.load(a.deck)
    if a.max_steps is not None:
        deck.run["max_step"] = a.max_steps
    if a.final_time is not None:
        deck.run["final_time"] = a.final_time
    r = Runner(deck)
    names = (["e_max", "e_tot", "ex_max", "ey_max", "ez_max", "e_sum_tot", "b_max", "b_tot", "bx_max", "by_max", "bz_max",
              "b_sum_tot"] if r.vm else ["e_max", "e_tot", "ex_max", "ey_max", "e_sum_tot"])
    per = ["ke", "ke_x", "ke_y", "px", "py"] + ([] if r.vm else ["ke_e_dot"])
    for sp in deck.species:
        names += ["%s_%s" % (sp.name, k) for k in per]
    ap_every = a.every_step
    while not r.done():
        dt = r.advance()
        rec = dict(step=r.step, time=r.time, dt=dt)
        if r.record or ap_every:   # time histories are collected every sequence_write_times (Simulation.C:318-321)
            rec.update(zip(names, r.history().tolist()))
            if not r.vm:
                # probes (Poisson.C:852-887) and the eight kinetic-energy fluxes per species (KineticSpecies.C:2052-2097)
                for k, (ex, ey) in enumerate(r.probes().tolist()):
                    rec["probe%d_ex" % (k + 1)], rec["probe%d_ey" % (k + 1)] = ex, ey
                fl = r.flux_history().tolist()
                for s_, sp in enumerate(deck.species):
                    for d_, dn in enumerate(("x", "y", "vx", "vy")):
                        for side, sn in enumerate(("lo", "hi")):
                            rec["%s_ke_flux_%s_%s" % (sp.name, dn, sn)] = fl[8 * s_ + 2 * d_ + side]
        print(json.dumps(rec))
    r.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
