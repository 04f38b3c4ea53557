"""ctypes binding of liboceananigans_b200.so (include/oceananigans_b200.h).

There is NO CPU path: ``load()`` raises if the CUDA library has not been built (``__graft_entry__.build()``)
or cannot be loaded.  ``Library(path)`` with an explicit path exists so that tests can drive the *same*
C ABI of the host-simulation build (tests/hostsim) — the package itself never does that.
"""
import ctypes as C
import os

OC_ABI_VERSION = 6
OC_MAX_TRACERS = 8
OC_MAX_FIELDS = 3 + OC_MAX_TRACERS
OC_TIMER_NAMES = ("tendency", "halo", "poisson_rhs", "fft", "poisson_mid", "projection", "aux", "substep", "comm")

# enums
OC_F64, OC_F32 = 0, 1
OC_PERIODIC, OC_BOUNDED, OC_FLAT = 0, 1, 2
OC_CENTERED2, OC_WENO5, OC_CENTERED4, OC_UPWIND3, OC_UPWIND5, OC_WENO3, OC_UPWIND1, OC_ADVECTION_NONE = range(8)
OC_WENO7, OC_WENO9 = 9, 10
OC_RK3, OC_AB2 = 0, 1
OC_CORIOLIS_NONE, OC_CORIOLIS_FPLANE, OC_CORIOLIS_BETAPLANE, OC_CORIOLIS_CARTESIAN, OC_CORIOLIS_NONTRADITIONAL_BETAPLANE = 0, 1, 2, 3, 4
OC_BUOYANCY_NONE, OC_BUOYANCY_TRACER, OC_BUOYANCY_SEAWATER_LINEAR = 0, 1, 2
OC_BC_DEFAULT, OC_BC_PERIODIC, OC_BC_FLUX, OC_BC_VALUE, OC_BC_GRADIENT, OC_BC_OPEN, OC_BC_NONE = range(7)
OC_FIELD_U, OC_FIELD_V, OC_FIELD_W, OC_FIELD_TRACER0 = 0, 1, 2, 3
OC_FIELD_PNHS, OC_FIELD_PHY, OC_FIELD_NU_E, OC_FIELD_KAPPA_E0, OC_FIELD_GN0, OC_FIELD_GM0 = 32, 33, 34, 40, 64, 96


class oc_bc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("has_value", C.c_int32), ("value", C.c_double)]


class oc_config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("float_type", C.c_int32),
        ("N", C.c_int32 * 3), ("H", C.c_int32 * 3), ("topology", C.c_int32 * 3),
        ("delta", C.c_double * 3), ("extent", C.c_double * 3),
        ("advection", C.c_int32), ("timestepper", C.c_int32), ("ab2_chi", C.c_double),
        ("n_tracers", C.c_int32),
        ("has_scalar_diffusivity", C.c_int32), ("nu", C.c_double), ("kappa", C.c_double * OC_MAX_TRACERS),
        ("has_amd", C.c_int32), ("amd_Cnu", C.c_double), ("amd_Ckappa", C.c_double * OC_MAX_TRACERS),
        ("buoyancy", C.c_int32), ("gravity", C.c_double), ("thermal_expansion", C.c_double), ("haline_contraction", C.c_double),
        ("tracer_T", C.c_int32), ("tracer_S", C.c_int32), ("tracer_b", C.c_int32),
        ("has_coriolis", C.c_int32), ("coriolis_f", C.c_double),
        ("bcs", (oc_bc * 6) * OC_MAX_FIELDS),
        ("device", C.c_int32), ("dist_rank", C.c_int32), ("dist_nranks", C.c_int32),
        ("z_stretched", C.c_int32), ("z_faces", C.POINTER(C.c_double)),
        ("smagorinsky", C.c_int32), ("amd_has_Cb", C.c_int32), ("smag_C", C.c_double), ("smag_Cb", C.c_double),
        ("smag_Pr", C.c_double * OC_MAX_TRACERS),
        ("coriolis_beta", C.c_double), ("origin_y", C.c_double), ("coriolis_fxyz", C.c_double * 3),
        ("tilted_gravity", C.c_int32), ("reserved2", C.c_int32), ("gravity_unit_vector", C.c_double * 3),
        ("amd_Cb", C.c_double),
        ("coriolis_gamma", C.c_double), ("coriolis_radius", C.c_double), ("origin_z", C.c_double),
        ("has_advection_dir", C.c_int32), ("advection_dir", C.c_int32 * 3),
        ("array_diffusivity", C.c_int32), ("dist_ranks_x", C.c_int32),
    ]


class oc_field_info(C.Structure):
    _fields_ = [("location", C.c_int32 * 3), ("interior_size", C.c_int32 * 3), ("parent_size", C.c_int32 * 3),
                ("device_ptr", C.c_void_p), ("stride_y", C.c_int64), ("stride_z", C.c_int64)]


class oc_diagnostics(C.Structure):
    _fields_ = [("cell_advection_timescale", C.c_double), ("max_abs_u", C.c_double), ("max_abs_v", C.c_double),
                ("max_abs_w", C.c_double), ("has_nan", C.c_int32), ("pad", C.c_int32)]


class oc_clock(C.Structure):
    _fields_ = [("time", C.c_double), ("iteration", C.c_int64), ("stage", C.c_int32),
                ("last_dt", C.c_double), ("last_stage_dt", C.c_double)]


EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                          C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t))

# every symbol include/oceananigans_b200.h declares: name -> (restype, argtypes)
_M = C.c_void_p
SYMBOLS = {
    "oc_last_error": (C.c_char_p, []),
    "oc_abi_version": (C.c_int, []),
    "oc_config_init": (None, [C.POINTER(oc_config)]),
    "oc_model_create": (C.c_int, [C.POINTER(oc_config), C.POINTER(_M)]),
    "oc_model_destroy": (C.c_int, [_M]),
    "oc_sync": (C.c_int, [_M]),
    "oc_field_info_get": (C.c_int, [_M, C.c_int, C.POINTER(oc_field_info)]),
    "oc_upload_interior": (C.c_int, [_M, C.c_int, C.c_void_p, C.c_size_t]),
    "oc_download_interior": (C.c_int, [_M, C.c_int, C.c_void_p, C.c_size_t]),
    "oc_upload_parent": (C.c_int, [_M, C.c_int, C.c_void_p, C.c_size_t]),
    "oc_download_parent": (C.c_int, [_M, C.c_int, C.c_void_p, C.c_size_t]),
    "oc_fill_halo_regions": (C.c_int, [_M, C.POINTER(C.c_int), C.c_int, C.c_int]),
    "oc_update_state": (C.c_int, [_M, C.c_int]),
    "oc_compute_tendencies": (C.c_int, [_M]),
    "oc_compute_flux_bc_tendencies": (C.c_int, [_M]),
    "oc_rk3_substep": (C.c_int, [_M, C.c_double, C.c_int]),
    "oc_ab2_step": (C.c_int, [_M, C.c_double, C.c_double]),
    "oc_cache_previous_tendencies": (C.c_int, [_M]),
    "oc_compute_pressure_correction": (C.c_int, [_M, C.c_double]),
    "oc_make_pressure_correction": (C.c_int, [_M, C.c_double]),
    "oc_poisson_solve": (C.c_int, [_M, C.c_void_p, C.c_void_p, C.c_size_t]),
    "oc_set_finalize": (C.c_int, [_M, C.c_int]),
    "oc_time_step_rk3": (C.c_int, [_M, C.c_double]),
    "oc_time_step_ab2": (C.c_int, [_M, C.c_double, C.c_int]),
    "oc_get_clock": (C.c_int, [_M, C.POINTER(oc_clock)]),
    "oc_set_clock": (C.c_int, [_M, C.POINTER(oc_clock)]),
    "oc_restore_previous_tendency": (C.c_int, [_M, C.c_int, C.c_void_p, C.c_size_t]),
    "oc_output_begin": (C.c_int, [_M, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.c_size_t, C.POINTER(C.c_int)]),
    "oc_upload_begin": (C.c_int, [_M, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_int)]),
    "oc_output_wait": (C.c_int, [_M, C.c_int]),
    "oc_output_test": (C.c_int, [_M, C.c_int, C.POINTER(C.c_int)]),
    "oc_compute_diagnostics": (C.c_int, [_M, C.POINTER(oc_diagnostics)]),
    "oc_field_maximum_abs": (C.c_int, [_M, C.c_int, C.POINTER(C.c_double)]),
    "oc_set_bc_array": (C.c_int, [_M, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "oc_set_diffusivity_bc": (C.c_int, [_M, C.c_int, C.c_int, C.c_int, C.c_double]),
    "oc_dist_unique_id": (C.c_int, [C.c_void_p]),
    "oc_dist_attach_nccl": (C.c_int, [_M, C.c_void_p]),
    "oc_dist_attach_host": (C.c_int, [_M, C.c_void_p, C.c_void_p]),
    "oc_timers_enable": (C.c_int, [_M, C.c_int]),
    "oc_timers_reset": (C.c_int, [_M]),
    "oc_timers_get": (C.c_int, [_M, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "oc_stopwatch_start": (C.c_int, [_M]),
    "oc_stopwatch_stop": (C.c_int, [_M, C.POINTER(C.c_double)]),
    "oc_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "oc_host_free": (C.c_int, [C.c_void_p]),
    "oc_launch_count": (C.c_int64, [_M]),
    "oc_device_bytes": (C.c_int, [_M, C.POINTER(C.c_int64)]),
}

PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIBRARY = os.path.join(PACKAGE_DIR, "lib", "liboceananigans_b200.so")


class OceananigansB200Error(RuntimeError):
    pass


class Library:
    """A loaded C-ABI library with typed entry points."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise OceananigansB200Error(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback.")
        self.path = path
        self.dll = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(self.dll, name)          # AttributeError if the library does not export the symbol
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)
        if self.oc_abi_version() != OC_ABI_VERSION:
            raise OceananigansB200Error("ABI version mismatch between the Python binding and the library")

    def check(self, status):
        if status != 0:
            msg = self.oc_last_error()
            raise OceananigansB200Error(f"liboceananigans_b200 error {status}: {msg.decode() if msg else ''}")


_default = None


def load():
    """The product library (CUDA, sm_100a).  Raises when it is missing or cannot be loaded."""
    global _default
    if _default is None:
        _default = Library(DEFAULT_LIBRARY)
    return _default
