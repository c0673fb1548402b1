"""Run a `.pp` deck on one GPU through the host mirror: `python -m loki_b200.run deck.pp [--max-steps N]`.

The loop is Simulation::advance (Simulation.C:302-351): dt = cfl * stableDt snapped to the next multiple of
save_times (selectTimeStep, :464-485), System::advance, and a time-history record every sequence_write_times
(VPSystem::accumulateSequences / the Maxwell analogue) printed as a JSON line.  Vlasov-Poisson (RK4/RK6) and
Vlasov-Maxwell (RK4) decks with the initial conditions and the driver the five benchmark decks use.

With --write-directory (or the deck's restart.write_directory under --save-data) a Vlasov-Poisson run also writes
what Simulation writes: <dir>.time_hists_<n>.hdf and <dir>.fields_<k>.hdf at every save time and restart dumps
<dir>/dist_<n>.hdf + .g0 on the deck's restart.time_interval / step_interval (outputs.py); --restart resumes from
the newest dump of the deck's restart.read_directory."""
import argparse
import ctypes as C
import json
import math
import sys

import numpy as np

from . import capi, host, outputs, pp


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
        locs = getattr(self.deck, "probes", None)
        if locs is None:
            locs = [(0.5, 0.0)]
        if not locs:                       # number_of_probes = 0 (e.g. the External2D deck): no probe histories
            return np.zeros((0, 2))
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
        if getattr(self, "out", None):
            self.out["dt"] = step
        # exact comparisons as in Simulation::advance (Simulation.C:318-327): when rounding leaves the time one ulp
        # short of a save time, selectTimeStep's remaining * (1 + 10 eps) micro-step closes the gap next
        self.record = False
        if self.time >= (self.last_seq + 1) * run.get("sequence_write_times", 1.0):
            self.record = True          # accumulateSequences()
            self.last_seq += 1
        self.save = False
        if self.time >= (self.last_save + 1) * run["save_times"]:
            self.save = True            # writePlotFile()
            self.last_save += 1
        if getattr(self, "out", None):
            if self.record:
                self.accumulate_sequences()
            if self.save:
                self.write_plot_file()
            self.write_checkpoint_file()
        return step

    # ---- Simulation's output side (Simulation.C:66-160, :262-290) ----

    def sequence_record(self):
        """one column of VPSystem::accumulateSequences (VPSystem.C:591-636) in Poisson::buildTimeHistoryNames order:
        5 field histories, (Ex, Ey) per probe, then 16 per species"""
        if self.vm:
            # the histories the device computes today: Maxwell's twelve field histories and computekemaxwell's five per
            # species (VMSystem::accumulateSequences; probes, flux and driver histories of the Maxwell system are not built)
            return list(self.history())
        ns = len(self.deck.species)
        h = self.history()
        pr = self.probes().reshape(-1)
        fl = self.flux_history()
        out = list(h[:5]) + list(pr)
        for s in range(ns):
            ked, env = C.c_double(), C.c_double()
            capi.check(self.H.lk_vp_driver_history(self.sys, s, self.time, C.byref(ked), C.byref(env)), "driver_history")
            out += list(h[5 + 6 * s:5 + 6 * s + 5]) + list(fl[8 * s:8 * s + 8]) + [ked.value, h[5 + 6 * s + 5], env.value]
        return out

    def em_vars(self):
        """(2, n2d, n1d): Ex, Ey of the last field solve with their ghost layers (EMSolverBase::plotCommon's m_em_vars)"""
        d = self.deck
        n1d, n2d = d.n[0] + 2 * d.ng, d.n[1] + 2 * d.ng
        if self.vm:
            out = np.empty((6, n2d, n1d))                      # Ex, Ey, Ez, Bx, By, Bz of the state
            capi.check(self.H.lk_vm_get_fields(self.sys, out.ctypes.data), "lk_vm_get_fields")
            return out
        out = np.empty((2, n2d, n1d))
        capi.check(self.L.lk_sync(None), "lk_sync")
        capi.check(self.L.lk_memcpy_d2h(out.ctypes.data, self.H.lk_vp_em_vars_ptr(self.sys), out.nbytes), "lk_memcpy_d2h")
        return out

    def vz(self, s):
        d = self.deck
        out = np.empty((d.n[1] + 2 * d.ng, d.n[0] + 2 * d.ng))
        capi.check(self.H.lk_vm_get_vz(self.sys, s, out.ctypes.data), "lk_vm_get_vz")
        return out

    def open_outputs(self, write_dir, restart_time_interval=None, restart_step_interval=None, max_files=16,
                     restart_index=0):
        """what Simulation's constructor sets up (Simulation.C:262-290): the sequences, the field writer, the restart
        cadence; then the time-0 history, plot and dump"""
        d = self.deck
        # after a restore the next dump is due at once (RestartManager::resetNextWriteTime, Simulation.C:251-254)
        self.out = dict(dir=write_dir, saved_seq=0, saved_save=0, time_seq=[], restart_index=restart_index,
                        max_files=max_files, t_int=restart_time_interval, s_int=restart_step_interval,
                        next_write=self.time, dt=0.0)
        names = [sp.name for sp in d.species]
        if self.vm:
            self.out["names"] = outputs.MAXWELL_FIELD_HISTORIES + [n + "_" + k for n in names for k in ("ke", "ke_x", "ke_y", "px", "py")]
        else:
            self.out["names"] = outputs.poisson_time_history_names(len(d.probes), 0, names)
        self.out["seq"] = [[] for _ in self.out["names"]]
        self.out["fields"] = outputs.FieldWriter(write_dir, (d.xlim[0], d.xlim[2]), d.dx, d.n, d.order, 1)
        self.accumulate_sequences()
        self.write_plot_file()
        self.write_checkpoint_file()

    def accumulate_sequences(self):
        o = self.out
        o["time_seq"].append(self.time)
        rec = self.sequence_record()
        assert len(rec) == len(o["seq"]), "time-history names and values out of step"
        for seq, v in zip(o["seq"], rec):
            seq.append(v)
        o["saved_seq"] += 1

    def write_plot_file(self):
        """Poisson::plot (Poisson.C:687-784): one time slice of EX, EY, then the time histories so far"""
        o, d = self.out, self.deck
        fw = o["fields"]
        npr = len(d.probes)
        plot_names = (outputs.maxwell_plot_names(False, [sp.name for sp in d.species]) if self.vm else
                      outputs.poisson_plot_names(False, []))
        fw.start_time_slice(self.time, o["dt"], plot_names, 0, npr, ([p[0] for p in d.probes], [p[1] for p in d.probes]), d.n)
        em = self.em_vars()
        planes = list(em) + ([self.vz(s) for s in range(len(d.species))] if self.vm else [])
        for name, plane in zip(plot_names, planes):
            fw.write_field(name, plane, (-d.ng, -d.ng), (-d.ng, -d.ng), (d.n[0] + 2 * d.ng, d.n[1] + 2 * d.ng), d.ng)
        fw.end_time_slice()
        outputs.write_time_histories(o["dir"] + ".time_hists", o["saved_save"], o["names"], o["seq"], o["time_seq"],
                                     o["saved_seq"], npr, 0)
        o["saved_save"] += 1

    def write_checkpoint_file(self):
        """Simulation::writeCheckpointFile (Simulation.C:133-160) with RestartManager::requiresAction's rule"""
        o, d = self.out, self.deck
        if o["t_int"] is not None:
            if not self.time >= o["next_write"]:
                return None
        elif o["s_int"] is not None:
            if self.step % o["s_int"] != 0:
                return None
        else:
            return None
        # m_system->updateGhosts(): the dump holds the ghost cells the next step would start from
        if not self.vm:
            capi.check(self.H.lk_vp_update_ghosts(self.sys), "lk_vp_update_ghosts")
        items = []
        for s, sp in enumerate(d.species):
            f = self.state(s)
            if getattr(sp, "tz", None):
                # with a twilight zone the dump holds the error against the exact solution (KineticSpecies.C:987-1004)
                capi.check(self.H.lk_vp_trig_tz_error(self.sys, s, self.time, f.ctypes.data), "lk_vp_trig_tz_error")
            vlim = sp.vlim
            ncell = [d.n[0], d.n[1], sp.nv[0], sp.nv[1]]
            x_lo = [d.xlim[0], d.xlim[2], vlim[0], vlim[2]]
            x_hi = [d.xlim[1], d.xlim[3], vlim[1], vlim[3]]
            dx = [(x_hi[k] - x_lo[k]) / ncell[k] for k in range(4)]
            item = dict(sp=dict(name=sp.name, mass=sp.mass, charge=sp.charge, bz_const=getattr(sp, "bz", 0.0)),
                        domain=(ncell, x_lo, x_hi, dx, d.periodic), tiles={0: f},
                        krook=outputs.krook_state(getattr(sp, "krook", None), x_lo, x_hi),
                        info=outputs.distrib_info(0, 0, d.ng, ncell, [1, 1, 1, 1]))
            if sp.driver and not self.vm:
                v = C.c_double()
                capi.check(self.H.lk_vp_ke_e_dot(self.sys, s, C.byref(v)), "lk_vp_ke_e_dot")
                item["integrated_e_dot_j"] = {0: v.value}
                item["sp"]["driver_state"] = (0, float(getattr(sp, "driver_phase", 0.0)), 0.0)
            items.append(item)
        if self.vm:
            n2 = [d.n[0], d.n[1]]
            info = (outputs.distrib_info(0, 0, d.ng, n2 + [6], [1, 1, 1]), outputs.distrib_info(0, 0, d.ng, n2, [1, 1]))
            name = outputs.write_vm_restart(o["dir"], o["restart_index"], items, d.ng, self.em_vars(),
                                            [self.vz(s) for s in range(len(d.species))], info, self.time, o["dt"], d.cfl,
                                            d.run["final_time"], max_files=o["max_files"])
        else:
            name = outputs.write_vp_restart(o["dir"], o["restart_index"], items, d.ng, self.time, o["dt"], d.cfl,
                                            d.run["final_time"], max_files=o["max_files"])
        o["restart_index"] += 1
        if o["t_int"] is not None:
            o["next_write"] += o["t_int"]
        return name

    def restore(self, read_dir):
        """RestartManager::restore (RestartManager.C:197-228): resume from the newest dist_<n>.hdf of read_dir"""
        import os
        idx = -1
        while os.path.exists(os.path.join(read_dir, "dist_%d.hdf" % (idx + 1))):
            idx += 1
        if idx < 0:
            raise FileNotFoundError("No distributions found from which to restart ... quitting")
        dump = outputs.read_vp_restart(os.path.join(read_dir, "dist_%d.hdf" % idx))
        if dump["num_procs"] != 1:
            raise NotImplementedError("restart dumps of more than one generating process")
        for s, item in enumerate(dump["species"]):
            f = np.ascontiguousarray(item["distribution"], dtype=np.float64)
            if f.shape != tuple(self.shapes[s]):
                raise ValueError("dump of %s has extents %s, the deck %s" % (item["name"], f.shape, self.shapes[s]))
            if self.vm:
                capi.check(self.H.lk_vm_set_state(self.sys, s, f.ctypes.data), "lk_vm_set_state")
                vz = np.ascontiguousarray(dump["vz"][s], dtype=np.float64)
                capi.check(self.H.lk_vm_set_vz(self.sys, s, vz.ctypes.data), "lk_vm_set_vz")
                continue
            capi.check(self.H.lk_vp_set_state(self.sys, s, f.ctypes.data), "lk_vp_set_state")
            if item["integrated_e_dot_j"] is not None:
                capi.check(self.H.lk_vp_set_ke_e_dot(self.sys, s, item["integrated_e_dot_j"]), "lk_vp_set_ke_e_dot")
        if self.vm:
            em = np.ascontiguousarray(dump["em_vars"], dtype=np.float64)
            capi.check(self.H.lk_vm_set_fields(self.sys, em.ctypes.data), "lk_vm_set_fields")
        self.time = dump["time"]
        run = self.deck.run
        # Simulation's constructor after a restore (Simulation.C:246-260): the counters follow from the time
        self.last_save = int(self.time / run["save_times"])
        self.last_seq = int(self.time / run.get("sequence_write_times", 1.0))
        capi.check((self.H.lk_vm_set_time if self.vm else self.H.lk_vp_set_time)(self.sys, self.time), "set_time")
        self._seed()
        return idx

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
    ap.add_argument("--write-directory", default=None, help="write time-history, field and restart files under this "
                    "base path (default with --save-data: the deck's restart.write_directory)")
    ap.add_argument("--save-data", action="store_true", help="write the output files the deck asks for")
    ap.add_argument("--restart", action="store_true", help="resume from the newest dump of the deck's restart.read_directory")
    a = ap.parse_args(argv)
    deck = pp.load(a.deck)
    if a.max_steps is not None:
        deck.run["max_step"] = a.max_steps
    if a.final_time is not None:
        deck.run["final_time"] = a.final_time
    r = Runner(deck)
    rs = deck.run.get("restart", {})
    restart_index = 0
    if a.restart or rs.get("start_from_restart"):
        restart_index = r.restore(rs.get("read_directory") or rs.get("write_directory")) + 1
    wdir = a.write_directory or (rs.get("write_directory") if a.save_data else None)
    if wdir:
        r.open_outputs(wdir, rs.get("time_interval"), rs.get("step_interval"), rs.get("max_files", 16), restart_index)
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
    if wdir:
        r.write_checkpoint_file()       # Simulation::finalize (Simulation.C:354-359)
    r.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
