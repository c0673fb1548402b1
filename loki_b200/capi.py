"""ctypes binding of include/loki_b200.h (one-to-one; no logic)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class LokiError(RuntimeError):
    pass


class Geom(C.Structure):
    _fields_ = [("n", C.c_int * 4), ("ng", C.c_int), ("order", C.c_int), ("dx", C.c_double * 4)]

    @staticmethod
    def make(n, order, dx):
        g = Geom()
        for k in range(4):
            g.n[k] = int(n[k])
            g.dx[k] = float(dx[k])
        g.order = int(order)
        g.ng = 2 if order == 4 else 3
        return g

    @property
    def nd(self):
        return tuple(self.n[k] + 2 * self.ng for k in range(4))


class Accel(C.Structure):
    _fields_ = [("kind", C.c_int), ("field", C.c_void_p), ("vz", C.c_void_p),
                ("vxface_velocities", C.c_void_p), ("vyface_velocities", C.c_void_p),
                ("normalization", C.c_double), ("bz_const", C.c_double)]


class Inflow(C.Structure):
    _fields_ = [("kind", C.c_int), ("fx", C.c_void_p), ("fv", C.c_void_p), ("fx2", C.c_void_p),
                ("fv2", C.c_void_p), ("fnorm", C.c_double), ("frac", C.c_double),
                ("ghost3", C.c_void_p), ("ghost4", C.c_void_p)]


class RkUpdate(C.Structure):
    _fields_ = [("f_old", C.c_void_p), ("delta_in", C.c_void_p), ("delta_out", C.c_void_p),
                ("pred", C.c_void_p), ("w_delta", C.c_double), ("c_pred", C.c_double),
                ("use_delta", C.c_int), ("n_prev", C.c_int), ("k_prev", C.c_void_p * 7),
                ("c_prev", C.c_double * 7), ("wrap", C.c_int), ("accel_bcs", C.c_void_p), ("inflow_preset", C.c_int),
                ("krook_nu", C.c_void_p), ("krook_dt", C.c_double), ("krook_ic", C.c_void_p),
                ("tile_set", C.c_int), ("cut_dirs", C.c_int)]


class PitchAngle(C.Structure):
    """lk_pitch_angle: the keys of PitchAngleCollisionOperator::parseParameters"""
    _fields_ = [("range_lo", C.c_double * 2), ("range_hi", C.c_double * 2), ("vfloor", C.c_double),
                ("vthermal_dt", C.c_double), ("nu_coef", C.c_double), ("conservative", C.c_int)]

    @staticmethod
    def make(range_lo, range_hi, vfloor, vthermal_dt, nu_coef, conservative=1):
        p = PitchAngle()
        for k in range(2):
            p.range_lo[k] = float(range_lo[k])
            p.range_hi[k] = float(range_hi[k])
        p.vfloor, p.vthermal_dt, p.nu_coef, p.conservative = float(vfloor), float(vthermal_dt), float(nu_coef), int(conservative)
        return p


class StageMoments(C.Structure):
    _fields_ = [("nmom", C.c_int), ("partial", C.c_void_p), ("capacity", C.c_int64)]


def library_path():
    # LOKI_B200_LIB: development aid for A/B kernel experiments (loki_b200.build.build_variant)
    return os.environ.get("LOKI_B200_LIB") or os.path.join(_HERE, "libloki_b200.so")


_vp = C.c_void_p
_PROTOS = {
    "lk_version": (C.c_int, []),
    "lk_last_error": (C.c_char_p, []),
    "lk_set_strict": (C.c_int, [C.c_int]),
    "lk_get_strict": (C.c_int, []),
    "lk_device_count": (C.c_int, []),
    "lk_set_rhs_variant": (C.c_int, [C.c_int]),
    "lk_weno_fit": (C.c_int, [C.c_int, _vp, _vp, _vp, C.c_int64, _vp]),
    "lk_xpby4d": (C.c_int, [_vp, _vp, C.c_double, C.POINTER(Geom), _vp]),
    "lk_max_accel": (C.c_int, [C.POINTER(Geom), C.POINTER(Accel), _vp, _vp]),
    "lk_set_phase_space_vel_4d": (C.c_int, [_vp, _vp, C.POINTER(Geom), C.POINTER(Accel), _vp, _vp]),
    "lk_set_acceleration_bcs_4d": (C.c_int, [_vp, C.POINTER(Geom), C.POINTER(Accel), C.POINTER(Inflow),
                                             C.POINTER(C.c_int * 4), _vp]),
    "lk_face_fluxes_4d": (C.c_int, [_vp, _vp, _vp, C.POINTER(Geom), _vp, C.c_int, _vp]),
    "lk_accum_flux_div_4d": (C.c_int, [_vp, C.POINTER(Geom), _vp, _vp, _vp, _vp, _vp]),
    "lk_ke_flux_from_fluxes": (C.c_int, [_vp, C.POINTER(Geom), _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_double, _vp]),
    "lk_ke_vel_space_flux": (C.c_int, [_vp, C.POINTER(Geom), _vp, _vp, _vp, C.c_int, C.c_int, C.c_double, _vp]),
    "lk_ke_flux_boundaries": (C.c_int, [_vp, _vp, C.POINTER(Geom), _vp, C.POINTER(Accel), C.c_double, C.POINTER(C.c_int * 8), _vp]),
    "lk_preset_inflow_ghosts_4d": (C.c_int, [_vp, C.POINTER(Geom), C.POINTER(Inflow), _vp]),
    "lk_rk_stage_update": (C.c_int, [_vp, C.POINTER(Geom), C.POINTER(RkUpdate), _vp]),
    "lk_vlasov_stage_can_split": (C.c_int, [_vp, C.POINTER(Geom), C.POINTER(Accel), C.POINTER(RkUpdate)]),
    "lk_vlasov_stage_folds_bcs": (C.c_int, [_vp, C.POINTER(Geom), C.POINTER(Accel), C.POINTER(RkUpdate)]),
    "lk_periodic_fill_4d": (C.c_int, [_vp, C.POINTER(Geom), C.c_int, C.c_int, _vp]),
    "lk_set_acceleration_bcs_4d_jb": (C.c_int, [_vp, C.POINTER(Geom), C.POINTER(Accel), C.POINTER(Inflow),
                                                C.POINTER(C.c_int * 4), _vp]),
    "lk_set_advection_bcs_4d_jb": (C.c_int, [_vp, C.POINTER(Geom), _vp, C.POINTER(Inflow), C.POINTER(C.c_int * 4), C.c_int,
                                             C.c_int, _vp]),
    "lk_set_advection_bcs_4d": (C.c_int, [_vp, C.POINTER(Geom), _vp, C.POINTER(Inflow), C.POINTER(C.c_int * 4), C.c_int,
                                          C.c_int, _vp]),
    "lk_halo_count": (C.c_int64, [C.POINTER(Geom), C.c_int]),
    "lk_halo_pack": (C.c_int, [_vp, _vp, C.POINTER(Geom), C.c_int, C.c_int, _vp]),
    "lk_halo_unpack": (C.c_int, [_vp, _vp, C.POINTER(Geom), C.c_int, C.c_int, _vp]),
    "lk_advection_derivatives_4d": (C.c_int, [_vp, _vp, C.POINTER(Geom), _vp, _vp]),
    "lk_acceleration_derivatives_4d": (C.c_int, [_vp, _vp, C.POINTER(Geom), C.POINTER(Accel), _vp]),
    "lk_vlasov_rhs": (C.c_int, [_vp, _vp, C.POINTER(Geom), _vp, C.POINTER(Accel), C.POINTER(RkUpdate), _vp]),
    "lk_stage_moment_parts": (C.c_int, [C.POINTER(Geom)]),
    "lk_vlasov_stage": (C.c_int, [_vp, _vp, C.POINTER(Geom), _vp, C.POINTER(Accel), C.POINTER(RkUpdate),
                                  C.POINTER(StageMoments), _vp]),
    "lk_moments_finish": (C.c_int, [_vp, _vp, _vp, C.POINTER(StageMoments), C.POINTER(Geom), C.c_double, C.c_double, _vp]),
    "lk_ke_e_dot_from_moments": (C.c_int, [_vp, C.POINTER(StageMoments), C.POINTER(Geom), C.c_double, _vp, _vp]),
    "lk_reduce_4d_to_2d": (C.c_int, [_vp, _vp, C.POINTER(Geom), C.c_double, C.c_double, _vp]),
    "lk_current_density": (C.c_int, [_vp, _vp, _vp, _vp, C.POINTER(Geom), _vp, _vp, C.c_double, C.c_double, _vp]),
    "lk_ke_e_dot": (C.c_int, [_vp, _vp, C.POINTER(Geom), C.c_double, _vp, _vp, _vp]),
    "lk_poisson_plan_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]),
    "lk_poisson_plan_destroy": (None, [_vp]),
    "lk_electric_field": (C.c_int, [_vp, _vp, _vp, _vp, C.POINTER(C.c_double), _vp]),
    "lk_periodic_fill_2d": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "lk_xpby2d": (C.c_int, [_vp, _vp, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "lk_form_accel": (C.c_int, [_vp, _vp, _vp, C.c_double, C.c_int, C.c_int, C.c_int, _vp]),
    "lk_maxwell_rhs": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.POINTER(C.c_double), C.c_double, C.c_double, C.c_double, _vp]),
    "lk_neutralize_charge": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp]),
    "lk_efield_from_potential": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _vp]),
    "lk_maxwell_vz_rhs": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_double, _vp]),
    "lk_f77_status": (C.c_int, []),
    "lk_trig_tz_table_count": (C.c_int, [C.POINTER(Geom), C.POINTER(C.c_int64)]),
    "lk_trig_tz_tables": (C.c_int, [_vp, C.POINTER(Geom), C.POINTER(C.c_int * 2), C.POINTER(C.c_double * 2), _vp, C.c_int,
                                    C.POINTER(C.c_double * 3), _vp]),
    "lk_set_trig_tz_source": (C.c_int, [_vp, C.POINTER(Geom), _vp, _vp, C.c_double, C.c_int, C.POINTER(C.c_double * 3), _vp]),
    "lk_compute_trig_tz_source_error": (C.c_int, [_vp, _vp, C.POINTER(Geom), _vp, _vp, C.c_double, C.c_int,
                                                  C.POINTER(C.c_double * 3), _vp]),
    "lk_zero_ghost_2d": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "lk_maxwell_add_antenna_source": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp]),
    "lk_maxwell_set_em_bcs": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int * 4), C.c_int, C.c_int, C.c_double, _vp]),
    "lk_maxwell_set_vz_bcs": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int * 4), C.c_int, C.c_int, _vp]),
    "lk_append_krook": (C.c_int, [_vp, _vp, C.POINTER(Geom), _vp, C.c_double, C.POINTER(Inflow), _vp]),
    "lk_pitch_angle_fields": (C.c_int, [_vp, _vp, _vp, _vp, C.POINTER(Geom), _vp, _vp]),
    "lk_append_pitch_angle_collision": (C.c_int, [_vp, _vp, C.POINTER(Geom), _vp, _vp, _vp, _vp, C.POINTER(C.c_double * 2),
                                                  C.POINTER(C.c_double * 2), C.POINTER(PitchAngle), _vp]),
    "lk_pitch_angle_real_lam": (C.c_double, [C.POINTER(Geom), C.POINTER(PitchAngle)]),
    "lk_pitch_angle_check": (C.c_int, [C.POINTER(Geom), C.POINTER(C.c_double * 2), C.POINTER(C.c_double * 2),
                                       C.POINTER(PitchAngle)]),
    "lk_compute_ke": (C.c_int, [_vp, _vp, C.POINTER(Geom), C.c_double, _vp, _vp, _vp]),
    "lk_field_history": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), _vp]),
    "lk_malloc": (C.c_int, [C.POINTER(_vp), C.c_int64]),
    "lk_free": (C.c_int, [_vp]),
    "lk_memcpy_h2d": (C.c_int, [_vp, _vp, C.c_int64]),
    "lk_memcpy_d2h": (C.c_int, [_vp, _vp, C.c_int64]),
    "lk_memset": (C.c_int, [_vp, C.c_int, C.c_int64]),
    "lk_sync": (C.c_int, [_vp]),
    "lk_launch_count": (C.c_int64, []),
    "lk_pipe_launch_count": (C.c_int64, []),
    "lk_profile_enable": (C.c_int, [C.c_int]),
    "lk_profile_summary": (C.c_int, [C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
}


def load():
    """Load libloki_b200.so (building is __graft_entry__.build()'s job).  Raises if it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise LokiError("libloki_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(path)
    for name, (res, args) in _PROTOS.items():
        if not hasattr(L, name):
            continue  # host-level symbols are bound lazily by their users
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = L
    return L


def lib():
    return load()


def check(status, what=""):
    if status != 0:
        raise LokiError("%s failed (%d): %s" % (what, status, lib().lk_last_error().decode()))
