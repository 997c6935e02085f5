"""ctypes binding of libomc_b200.so (include/omc_b200.h).

The library is the product; this module only declares its C ABI.  There is no CPU
fallback: ``load()`` raises if the shared object is missing and every compute entry
returns OMC_ERR_CUDA when no sm_100 device is usable.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OMC_B200_LIB") or os.path.join(_HERE, "libomc_b200.so")   # (override: A/B builds during development)

OK = 0
STATUS_OPTIMAL, STATUS_ITERATION_LIMIT, STATUS_INFEASIBLE, STATUS_TIME_LIMIT, STATUS_CUTOFF, STATUS_NUMERICAL = 0, 1, 2, 3, 4, 5
CUT_TYPES = {"linear": 0, "linear2": 1, "linear3": 2}
ENGINES = {"auto": 0, "persistent": 1, "batched": 2}


class RelaxOpts(C.Structure):
    _fields_ = [
        ("eps_abs", C.c_double), ("eps_rel", C.c_double),
        ("max_iter", C.c_int32), ("check_every", C.c_int32), ("adapt_every", C.c_int32),
        ("fix_linear3_right", C.c_int32),
        ("rho0", C.c_double), ("sigma", C.c_double), ("alpha", C.c_double),
        ("cutoff", C.c_double), ("time_limit_s", C.c_double), ("jacobi_tol", C.c_double),
        ("reortho_every", C.c_int32), ("exact_projection", C.c_int32),
    ]


_p = C.POINTER
_vp, _i32, _i64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
_pi32, _pf64, _pu8, _pu64, _pf32 = _p(C.c_int32), _p(C.c_double), _p(C.c_uint8), _p(C.c_uint64), _p(C.c_float)

# name -> (restype, argtypes): exactly the declarations of include/omc_b200.h
SIGNATURES = {
    "omc_init": (_i32, [_i32]),
    "omc_shutdown": (_i32, []),
    "omc_last_error": (C.c_char_p, []),
    "omc_device_info": (_i32, [_pi32, _pi32, _pi32, _p(C.c_int64)]),
    "omc_build_flags": (_i32, []),
    "omc_stream": (_vp, []),
    "omc_problem_create": (_i32, [_i32, _i32, _i32, _pf64, _pu64, _f64, _i32, _i32, _p(_vp)]),
    "omc_problem_destroy": (_i32, [_vp]),
    "omc_problem_get_csr": (_i32, [_vp, _pi32, _pi32, _pi32, _pi32, _p(C.c_int64)]),
    "omc_cutpool_add": (_i32, [_vp, _pf64, _pf64, _pi32]),
    "omc_cutpool_size": (_i32, [_vp, _pi32]),
    "omc_relax_default_opts": (None, [_p(RelaxOpts)]),
    "omc_frontier_create": (_i32, [_vp, _i32, _pi32, _pi32, _pu8, _pi32, _pi32, _p(_vp)]),
    "omc_frontier_create_ex": (_i32, [_vp, _i32, _pi32, _pi32, _pu8, _pi32, _pi32, _i32, _p(_vp)]),
    "omc_frontier_stats": (_i32, [_vp, _p(C.c_int64)]),
    "omc_shor_score_minors": (_i32, [_vp, _pf64, _i32, _i64, _pi32, _i64, _pi32, _i64, _p(C.c_int64), _pi32, _pf64]),
    "omc_profile_kernels": (_i32, [_vp, _i32, _pf32]),
    "omc_problem_set_shor": (_i32, [_vp, _i64, _pi32, _i64, _pi32]),
    "omc_frontier_fetch_shor": (_i32, [_vp, _pf64, _pf64]),
    "omc_frontier_set_tuning": (_i32, [_vp, _i32, _i32, _f64, _f64]),
    "omc_frontier_debug_fetch": (_i64, [_vp, _i32, _i32, _pf64, _i64]),
    "omc_frontier_relax": (_i32, [_vp, _p(RelaxOpts), _pf32]),
    "omc_frontier_fetch": (_i32, [_vp, _pi32, _pf64, _pf64, _pi32, _pf64, _pf64, _pf64, _pf64, _pf64]),
    "omc_frontier_fetch_profile": (_i32, [_vp, _pf64]),
    "omc_frontier_destroy": (_i32, [_vp]),
    "omc_relax_batch": (_i32, [_vp, _i32, _pi32, _pi32, _pu8, _pi32, _pi32, _p(RelaxOpts), _pi32, _pf64, _pf64,
                               _pi32, _pf64, _pf64, _pf64, _pf64, _pf64, _pf32]),
    "omc_smallest_eigvecs_batch": (_i32, [_i32, _i32, _i32, _pf64, _pf64, _i32, _pf64, _pf64, _pf64, _pi32]),
    "omc_altmin": (_i32, [_vp, _pf64, _i32, _pi32, _pu8, _f64, _i32, _f64, _pf64, _pf64, _pi32, _pi32, _pf64, _pf64]),
    "omc_altmin_batch": (_i32, [_vp, _i32, _pf64, _pi32, _pi32, _pu8, _f64, _i32, _f64, _pf64, _pf64, _pi32, _pi32, _pf64, _pf64]),
    "omc_objective_mse": (_i32, [_vp, _pf64, _pf64]),
    "omc_shor_indexes": (_i32, [_vp, _pi32, _i32, _p(C.c_int64), _pi32, _i64, _pi32, _p(C.c_int64)]),
    "omc_comm_unique_id": (_i32, [_pu8]),
    "omc_comm_init": (_i32, [_i32, _i32, _pu8]),
    "omc_comm_info": (_i32, [_pi32, _pi32]),
    "omc_allreduce_min": (_i32, [_pf64, _i32]),
    "omc_allgather": (_i32, [_pf64, _i32, _pf64]),
    "omc_comm_destroy": (_i32, []),
    "omc_comm_last_error": (C.c_char_p, []),
    "omc_debug_psd_project_batch": (_i32, [_i32, _i32, _pf64, _pf64, _pf64, _pi32, _pf32]),
    "omc_measure_fp64_peak": (_i32, [_pf64]),
}

_lib = None


def load():
    """dlopen the in-tree library and attach the prototypes.  Raises when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `make` (nvcc, sm_100a). "
                           "omc_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class OmcError(RuntimeError):
    pass


def check(rc):
    if rc != OK:
        msg = load().omc_last_error()
        raise OmcError(f"libomc_b200 error {rc}: {msg.decode() if msg else ''}")
