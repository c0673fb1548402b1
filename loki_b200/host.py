"""ctypes binding of include/loki_b200_host.h (the C++ host mirror of VPSystem / KineticSpecies / RK
integrators inside libloki_b200.so).  One-to-one, no logic."""
import ctypes as C

from .capi import Geom, PitchAngle, load, check

_vp = C.c_void_p


class SpeciesDesc(C.Structure):
    _fields_ = [("nv", C.c_int * 2), ("vlo", C.c_double * 2), ("vhi", C.c_double * 2), ("mass", C.c_double),
                ("charge", C.c_double), ("bz_const", C.c_double), ("has_driver", C.c_int),
                ("driver", C.c_double * 16), ("driver_phase", C.c_double), ("driver_shape_type", C.c_int)]


class VPDesc(C.Structure):
    _fields_ = [("nspecies", C.c_int), ("species", C.POINTER(SpeciesDesc)), ("order", C.c_int),
                ("rk_order", C.c_int), ("nglobal", C.c_int * 2), ("xlo", C.c_double * 2), ("xhi", C.c_double * 2),
                ("tile_lo", C.c_int * 2), ("tile_n", C.c_int * 2), ("ntiles", C.c_int)]


class VMDesc(C.Structure):
    _fields_ = [("base", VPDesc), ("light_speed", C.c_double), ("av_weak", C.c_double), ("av_strong", C.c_double)]


_PROTOS = {
    "lk_vm_create": (C.c_int, [C.POINTER(_vp), C.POINTER(VMDesc), _vp]),
    "lk_vm_destroy": (None, [_vp]),
    "lk_vm_species_geom": (C.c_int, [_vp, C.c_int, C.POINTER(Geom)]),
    "lk_vm_set_state": (C.c_int, [_vp, C.c_int, _vp]),
    "lk_vm_get_state": (C.c_int, [_vp, C.c_int, _vp]),
    "lk_vm_state_ptr": (_vp, [_vp, C.c_int]),
    "lk_vm_set_fields": (C.c_int, [_vp, _vp]),
    "lk_vm_get_fields": (C.c_int, [_vp, _vp]),
    "lk_vm_set_vz": (C.c_int, [_vp, C.c_int, _vp]),
    "lk_vm_get_vz": (C.c_int, [_vp, C.c_int, _vp]),
    "lk_vm_fields_ptr": (_vp, [_vp]),
    "lk_vm_current_ptr": (_vp, [_vp, C.c_int]),
    "lk_vm_set_inflow": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_double, C.c_double]),
    "lk_vm_set_inflow_ghosts": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "lk_vm_set_time": (C.c_int, [_vp, C.c_double]),
    "lk_vm_time": (C.c_double, [_vp]),
    "lk_vm_advance": (C.c_int, [_vp, C.c_double]),
    "lk_vm_stable_dt": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "lk_vm_lambda_max": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_double * 2)]),
    "lk_vm_eval_rhs": (C.c_int, [_vp, C.POINTER(_vp), _vp, C.POINTER(_vp), C.c_double]),
    "lk_vp_create": (C.c_int, [C.POINTER(_vp), C.POINTER(VPDesc), _vp]),
    "lk_vp_destroy": (None, [_vp]),
    "lk_vp_species_geom": (C.c_int, [_vp, C.c_int, C.POINTER(Geom)]),
    "lk_vp_set_state": (C.c_int, [_vp, C.c_int, _vp]),
    "lk_vp_get_state": (C.c_int, [_vp, C.c_int, _vp]),
    "lk_vp_state_ptr": (_vp, [_vp, C.c_int]),
    "lk_vp_eval_ptr": (_vp, [_vp, C.c_int]),
    "lk_vp_set_inflow": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_double, C.c_double]),
    "lk_vp_set_inflow2": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "lk_vp_set_time": (C.c_int, [_vp, C.c_double]),
    "lk_vp_time": (C.c_double, [_vp]),
    "lk_vp_stable_dt": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "lk_vp_lambda_max": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_double * 2)]),
    "lk_vp_set_lambda_max": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_double * 2)]),
    "lk_vp_advance": (C.c_int, [_vp, C.c_double]),
    "lk_vp_nstages": (C.c_int, [_vp]),
    "lk_vp_begin_step": (C.c_int, [_vp, C.c_double]),
    "lk_vp_stage_moments": (C.c_int, [_vp, C.c_int]),
    "lk_vp_set_comm_buffers": (C.c_int, [_vp, _vp, _vp]),
    "lk_vp_rho_tile_ptr": (_vp, [_vp]),
    "lk_vp_rho_gather_ptr": (_vp, [_vp]),
    "lk_vp_stage_field": (C.c_int, [_vp, C.c_int, _vp]),
    "lk_vp_local_fill": (C.c_int, [_vp, C.c_int, C.c_int]),
    "lk_vp_local_fill_needed": (C.c_int, [_vp, C.c_int, C.c_int]),
    "lk_vp_stage_finish": (C.c_int, [_vp, C.c_int]),
    "lk_vp_stage_finish_species": (C.c_int, [_vp, C.c_int, C.c_int]),
    "lk_vp_stage_finish_species_part": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int]),
    "lk_vp_wait_faces": (C.c_int, [_vp, C.c_int, _vp]),
    "lk_vp_end_step": (C.c_int, [_vp]),
    "lk_vp_eval_rhs": (C.c_int, [_vp, C.POINTER(_vp), C.c_double]),
    "lk_vp_em_vars_ptr": (_vp, [_vp]),
    "lk_vp_rho_ptr": (_vp, [_vp]),
    "lk_vp_ke_e_dot": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_double)]),
    "lk_vp_set_trig_tz": (C.c_int, [_vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]),
    "lk_vp_trig_tz_error": (C.c_int, [_vp, C.c_int, C.c_double, _vp]),
    "lk_vp_update_ghosts": (C.c_int, [_vp]),
    "lk_vp_set_ke_e_dot": (C.c_int, [_vp, C.c_int, C.c_double]),
    "lk_vp_driver_history": (C.c_int, [_vp, C.c_int, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "lk_vp_time_history": (C.c_int, [_vp, _vp, C.c_int, C.POINTER(C.c_int)]),
    "lk_vp_set_boundary_options": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int]),
    "lk_vp_set_krook": (C.c_int, [_vp, C.c_int, _vp]),
    "lk_vp_set_pitch_angle": (C.c_int, [_vp, C.c_int, C.POINTER(PitchAngle)]),
    "lk_vp_download_state": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "lk_vp_upload_next": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "lk_vp_adopt_next": (C.c_int, [_vp]),
    "lk_vp_probe_history": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp]),
    "lk_vp_flux_history": (C.c_int, [_vp, _vp, C.c_int, C.POINTER(C.c_int)]),
    "lk_vm_time_history": (C.c_int, [_vp, _vp, C.c_int, C.POINTER(C.c_int)]),
}
_bound = False


def lib():
    global _bound
    L = load()
    if not _bound:
        for name, (res, args) in _PROTOS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _bound = True
    return L
