"""ctypes binding of the CPU oracle (oracle/libloki_oracle.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")
_L = None

dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
IC_FN = C.CFUNCTYPE(C.c_double, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int)


class OkGeom(C.Structure):
    _fields_ = [("n", C.c_int * 4), ("ng", C.c_int), ("order", C.c_int), ("dx", C.c_double * 4)]

    @staticmethod
    def make(n, order, dx):
        g = OkGeom()
        for k in range(4):
            g.n[k] = int(n[k])
            g.dx[k] = float(dx[k])
        g.order = int(order)
        g.ng = 2 if order == 4 else 3
        return g

    @property
    def nd(self):
        return tuple(self.n[k] + 2 * self.ng for k in range(4))


class OkIcTables(C.Structure):
    _fields_ = [("kind", C.c_int), ("n1d", C.c_int), ("n2d", C.c_int), ("n3d", C.c_int), ("n4d", C.c_int),
                ("fx", C.c_void_p), ("fv", C.c_void_p), ("fx2", C.c_void_p), ("fv2", C.c_void_p), ("full", C.c_void_p),
                ("fnorm", C.c_double), ("frac", C.c_double)]


class OkSpecies(C.Structure):
    _fields_ = [("g", OkGeom), ("mass", C.c_double), ("charge", C.c_double), ("bz_const", C.c_double),
                ("vlo", C.c_double * 2), ("vhi", C.c_double * 2), ("ic", IC_FN), ("ic_ctx", C.c_void_p),
                ("has_driver", C.c_int), ("driver", C.c_double * 16), ("driver_phase", C.c_double),
                ("driver_shape_type", C.c_int)]


def load():
    global _L
    if _L is not None:
        return _L
    so = os.path.join(ODIR, "libloki_oracle.so")
    srcs = [os.path.join(ODIR, f) for f in ("loki_oracle.c", "loki_oracle_vp.c", "loki_oracle_vm.c", "loki_oracle_coll.c", "loki_oracle.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", ODIR, "libloki_oracle.so"], stdout=subprocess.DEVNULL)
    L = C.CDLL(so)
    G = C.POINTER(OkGeom)
    d = C.c_double
    i = C.c_int
    L.ok_weno43_fit.restype = d
    L.ok_weno43_fit.argtypes = [d] * 5
    L.ok_weno65_fit.restype = d
    L.ok_weno65_fit.argtypes = [d] * 7
    L.ok_weno43_fit_v.argtypes = [dp, dp, dp, C.c_int64]
    L.ok_weno65_fit_v.argtypes = [dp, dp, dp, C.c_int64]
    L.ok_xpby4d.argtypes = [dp, dp, d, G]
    L.ok_set_phase_space_vel_4d.argtypes = [dp, dp, G, dp, dp, d, d, dp, C.POINTER(d), C.POINTER(d)]
    L.ok_set_phase_space_vel_maxwell_4d.argtypes = [dp, dp, G, dp, dp, d, d, dp, dp, C.POINTER(d), C.POINTER(d)]
    L.ok_set_acceleration_bcs_4d.argtypes = [dp, G, dp, dp, i, i, i, i, IC_FN, C.c_void_p]
    L.ok_set_advection_bcs_4d.argtypes = [dp, G, dp, dp, i, i, i, i, i, i, IC_FN, C.c_void_p]
    L.ok_set_acceleration_bcs_4d_jb.argtypes = [dp, G, dp, dp, i, i, i, i, IC_FN, C.c_void_p]
    L.ok_set_advection_bcs_4d_jb.argtypes = [dp, G, dp, dp, i, i, i, i, i, i, IC_FN, C.c_void_p]
    L.ok_advection_derivatives_4d.argtypes = [dp, dp, G, dp, dp]
    L.ok_acceleration_derivatives_4d.argtypes = [dp, dp, G, dp, dp]
    L.ok_compute_currents.argtypes = [G, dp, dp, dp, dp, dp, dp]
    L.ok_compute_ke_e_dot.restype = d
    L.ok_compute_ke_e_dot.argtypes = [G, dp, d, dp, dp, d]
    L.ok_reduce_4d_to_2d.argtypes = [dp, dp, G, d, d]
    L.ok_append_krook.argtypes = [dp, dp, G, dp, d, IC_FN, C.c_void_p]
    L.ok_compute_ke.argtypes = [G, dp, d, dp, dp]
    L.ok_face_fluxes_4d.argtypes = [dp, dp, dp, G, dp, i]
    L.ok_accum_flux_div_4d.argtypes = [dp, G, dp, dp, dp, dp]
    L.ok_compute_ke_flux.restype = d
    L.ok_compute_ke_flux.argtypes = [G, dp, dp, dp, dp, dp, dp, dp, i, i, d]
    L.ok_compute_ke_vel_space_flux.argtypes = [dp, G, dp, dp, dp, dp, i, i, d]
    L.ok_compute_ke_maxwell.argtypes = [G, dp, d, dp, dp, dp]
    L.ok_field_history.argtypes = [dp, i, i, i, i, dp, dp]
    L.ok_periodic_fill_4d.argtypes = [dp, G, i, i]
    L.ok_periodic_fill_2d.argtypes = [dp, i, i, i, i, i, i]
    L.ok_build_velocity_tables.argtypes = [G, C.POINTER(i * 2), d, d, dp, dp, dp]
    L.ok_initialize_velocity.argtypes = [G, dp, dp, dp]
    L.ok_neutralize_charge.argtypes = [dp, i, i, i]
    L.ok_poisson_symbols.argtypes = [i, i, d, d, i, dp, dp]
    L.ok_poisson_fft_solve.argtypes = [dp, dp, i, i, i, dp, dp]
    L.ok_efield_from_potential.argtypes = [dp, dp, i, i, i, i, i, dp]
    L.ok_vp_work_create.restype = C.c_void_p
    L.ok_vp_work_create.argtypes = [i, C.POINTER(OkSpecies), C.POINTER(d * 2), C.POINTER(d * 2)]
    L.ok_vp_work_destroy.argtypes = [C.c_void_p]
    L.ok_vp_eval_rhs.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), d, dp, dp, dp]
    L.ok_vp_em_vars.restype = C.POINTER(d)
    L.ok_vp_em_vars.argtypes = [C.c_void_p]
    L.ok_vp_rho.restype = C.POINTER(d)
    L.ok_vp_rho.argtypes = [C.c_void_p]
    L.ok_vp_rk4_step.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), d, d, dp]
    L.ok_vp_rk6_step.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), d, d, dp]
    L.ok_vp_last_accel_max.argtypes = [C.c_void_p, dp, dp]
    L.ok_vp_set_options.argtypes = [C.c_void_p, i, i, i]
    L.ok_vp_set_krook.argtypes = [C.c_void_p, i, dp]
    L.ok_vp_set_dt.argtypes = [C.c_void_p, d]
    L.ok_vp_set_pitch_angle.argtypes = [C.c_void_p, i, C.c_void_p]
    L.ok_vp_set_trig_tz.argtypes = [C.c_void_p, i, i, d, d, d]
    L.ok_set_two_species_trig_tz_source.argtypes = [dp, G, C.c_void_p, dp, dp, d, dp, dp, i]
    L.ok_compute_two_species_trig_tz_source_error.argtypes = [dp, dp, G, C.c_void_p, dp, dp, d, dp, dp, i]
    L.ok_set_trig_tz_source.argtypes = [dp, G, C.c_void_p, dp, dp, d, dp, d]
    L.ok_compute_trig_tz_source_error.argtypes = [dp, dp, G, C.c_void_p, dp, dp, d, dp, d]
    L.ok_set_electron_trig_tz_source.argtypes = [dp, G, C.c_void_p, dp, dp, d, dp, d]
    L.ok_compute_electron_trig_tz_source_error.argtypes = [dp, dp, G, C.c_void_p, dp, dp, d, dp, d]
    L.ok_pitch_angle_collisionality.restype = d
    L.ok_pitch_angle_collisionality.argtypes = [d, d, d, d, dp, dp, d, d, d, d, d, d, d, i]
    L.ok_pitch_angle_fields.argtypes = [dp, dp, dp, dp, G, dp]
    L.ok_append_pitch_angle_collision.argtypes = [dp, dp, G, dp, dp, dp, dp, dp, dp, dp, dp, d, d, i]
    L.ok_pitch_angle_real_lam.restype = d
    L.ok_pitch_angle_real_lam.argtypes = [G, d, d, d]
    L.ok_vp_ke_flux_history.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), dp]
    L.ok_vp_stable_dt.restype = d
    L.ok_vp_stable_dt.argtypes = [C.c_void_p, dp, dp, i]
    L.ok_shaped_ramped_driver.argtypes = [dp, dp, i, i, i, i, dp, dp, i, d, dp, d, i]
    vp, pvp = C.c_void_p, C.POINTER(C.c_void_p)
    L.ok_xpby2d.argtypes = [dp, dp, d, i, i, i, i]
    L.ok_maxwell_eval_rhs.argtypes = [dp, dp, dp, dp, dp, i, i, i, i, dp, d, d, d]
    L.ok_maxwell_eval_vz_rhs.argtypes = [dp, dp, d, i, i, i]
    L.ok_zero_ghost_2d.argtypes = [dp, i, i, i, i]
    L.ok_maxwell_add_antenna_source.argtypes = [dp, dp, i, i, i]
    L.ok_maxwell_set_em_bcs.argtypes = [dp, i, i, i, C.c_void_p, i, i, d]
    L.ok_maxwell_set_vz_bcs.argtypes = [dp, i, i, i, C.c_void_p, i, i]
    L.ok_vm_work_create.restype = vp
    L.ok_vm_work_create.argtypes = [i, C.POINTER(OkSpecies), C.POINTER(d * 2), C.POINTER(d * 2), d, d, d]
    L.ok_vm_work_destroy.argtypes = [vp]
    L.ok_vm_eval_rhs.argtypes = [vp, pvp, dp, pvp, pvp, dp, pvp, d, dp, dp]
    L.ok_vm_net_current.restype = C.POINTER(d)
    L.ok_vm_net_current.argtypes = [vp, i]
    L.ok_vm_rk4_step.argtypes = [vp, pvp, pvp, dp, dp, pvp, pvp, d, d]
    L.ok_vm_rk6_step.argtypes = [vp, pvp, pvp, dp, dp, pvp, pvp, d, d]
    L.ok_vm_last_accel_max.argtypes = [vp, dp, dp]
    L.ok_vm_stable_dt.restype = d
    L.ok_vm_stable_dt.argtypes = [vp, dp, dp, i]
    L.ok_simple_em_ic.argtypes = [dp, i, i, i, dp, dp, i, d, d, d, d, d, d]
    L.ok_simple_vel_ic.argtypes = [dp, i, i, i, dp, dp, d, d, d, d]
    L.ok_time_rk4_stage_reference_style.restype = d
    L.ok_time_rk4_stage_reference_style.argtypes = [G, i, i]
    _L = L
    return L
