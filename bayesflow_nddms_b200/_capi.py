"""ctypes binding of include/ddm_b200.h (libddm_b200.so).

There is no CPU fallback: if the library is missing, or no B200 is visible, creating a
simulator raises.  Nothing here imports the oracle.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

# enum ddm_model
MODEL_BASIC = 0
MODEL_ALPHA = 1
MODEL_ALPHA_DC = 2
MODEL_ALPHA_SCALE = 3
MODEL_ALPHA_SCALE2 = 4
MODEL_TRIALWISE = 5
MODEL_ETA = 6
MODEL_GENERAL = 7
N_PARAMS = {MODEL_BASIC: 5, MODEL_ALPHA: 7, MODEL_ALPHA_DC: 7, MODEL_ALPHA_SCALE: 8, MODEL_ALPHA_SCALE2: 7,
            MODEL_TRIALWISE: 4, MODEL_ETA: 6, MODEL_GENERAL: 24}
N_COLS = {MODEL_GENERAL: 3}  # output columns per trial; every other model has 2

# enum ddm_prior: name -> (id, columns)
PRIORS = {"basic": (0, 5), "alpha": (1, 7), "alpha_dc": (2, 7), "alpha_scale": (3, 8), "alpha_scale2": (4, 7),
          "eta": (6, 6), "sweep": (7, 5), "evidence": (8, 6)}

# enum ddm_status
OK = 0
ERR_INVALID = -1
ERR_CUDA = -2
ERR_NOMEM = -3
ERR_NEGATIVE_BOUND = -4
ERR_STATE = -5

# enum ddm_flags
FLAG_TIMEOUT_CHOICE_ONE = 1
FLAG_OUT_F32 = 2
FLAG_KEEP_STEPS = 4
FLAG_FORCE_GENERIC = 8
FLAG_OUT_STATE = 16
FLAG_F32_NORMALS = 32

MB_NAMES = ["ffma", "imad_wide", "lop3", "iadd3", "mufu_lg2", "mufu_sin", "mix_fma_alu", "fsetp", "philox",
            "sim_block", "mix_imadw_lop3", "mix_mufu_lop3", "mix_mufu_imadw", "mix_blocklike", "ffma_3reg", "fadd_2reg",
            "imad_wide_noacc", "imad_hi", "ffma2_packed", "philox7", "sim_block_x2"]


class Stats(C.Structure):
    _fields_ = [("n_trials", C.c_uint64), ("total_steps", C.c_uint64), ("n_timeouts", C.c_uint64),
                ("n_upper", C.c_uint64), ("reject_cap_hits", C.c_uint64), ("kernel_ms", C.c_double),
                ("kernel_launches", C.c_int32), ("used_persistent", C.c_int32), ("grid", C.c_int32),
                ("block", C.c_int32), ("refill_threshold", C.c_int32), ("tile", C.c_int32),
                ("debug_overruns", C.c_uint64), ("d2h_bytes", C.c_uint64), ("host_decode_threads", C.c_int32),
                ("scheduler", C.c_int32)]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_}


class DLDevice(C.Structure):
    _fields_ = [("device_type", C.c_int), ("device_id", C.c_int32)]


class DLDataType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class DLTensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("device", DLDevice), ("ndim", C.c_int32), ("dtype", DLDataType),
                ("shape", C.POINTER(C.c_int64)), ("strides", C.POINTER(C.c_int64)), ("byte_offset", C.c_uint64)]


class DLManagedTensor(C.Structure):
    pass


DLManagedTensor._fields_ = [("dl_tensor", DLTensor), ("manager_ctx", C.c_void_p),
                            ("deleter", C.CFUNCTYPE(None, C.POINTER(DLManagedTensor)))]

_lib = None

# name -> (restype, argtypes); every function include/ddm_b200.h declares
_dp, _vp = C.POINTER(C.c_double), C.c_void_p
SIGNATURES = {
    "ddm_version": (C.c_int, []),
    "ddm_philox_rounds": (C.c_int, []),
    "ddm_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "ddm_destroy": (C.c_int, [_vp]),
    "ddm_last_error": (C.c_char_p, [_vp]),
    "ddm_set_stream": (C.c_int, [_vp, _vp]),
    "ddm_synchronize": (C.c_int, [_vp]),
    "ddm_set_tuning": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int]),
    "ddm_set_kernel_variant": (C.c_int, [_vp, C.c_int]),
    "ddm_set_pipeline": (C.c_int, [_vp, C.c_int64, C.c_int64]),
    "ddm_set_host_decode": (C.c_int, [_vp, C.c_int]),
    "ddm_pipeline_chunks": (C.c_int64, [C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int64]),
    "ddm_histogram_chunks": (C.c_int64, [C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int64]),
    "ddm_wire_decode_host": (C.c_int, [C.c_void_p, C.c_void_p, _dp, C.c_int, C.c_int64, C.c_int64, C.c_double, C.c_int,
                                       C.c_int, C.c_int]),
    "ddm_simulate": (C.c_int, [_vp, C.c_int, _dp, C.c_int64, C.c_int, C.c_int64, C.c_double, C.c_int, C.c_uint64,
                               C.c_uint64, C.c_int, C.c_int, _vp]),
    "ddm_draw_prior": (C.c_int, [_vp, C.c_int, C.c_int64, C.c_uint64, C.c_uint64, _dp]),
    "ddm_upload_params": (C.c_int, [_vp, C.c_int, _dp, C.c_int64, C.c_int]),
    "ddm_run": (C.c_int, [_vp, C.c_int64, C.c_double, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_int]),
    "ddm_download": (C.c_int, [_vp, _vp]),
    "ddm_simulate_trialwise": (C.c_int, [_vp, C.POINTER(C.c_int32), _dp, _dp, C.c_int64, C.c_int, C.c_double, C.c_int,
                                         C.c_uint64, C.c_uint64, C.c_int, C.c_int, _vp]),
    "ddm_simulate_evidence": (C.c_int, [_vp, _dp, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_int, C.c_uint64,
                                        C.c_uint64, C.c_int, C.c_int, _vp]),
    "ddm_last_steps": (C.c_int, [_vp, C.POINTER(C.c_int32)]),
    "ddm_simulate_exact": (C.c_int, [_vp, _dp, C.c_int64, C.c_int64, C.c_uint64, C.c_uint64, _dp]),
    "ddm_last_output_histogram": (C.c_int, [_vp, C.c_int, C.c_double, C.POINTER(C.c_uint64)]),
    "ddm_simulate_histogram": (C.c_int, [_vp, C.c_int, _dp, C.c_int64, C.c_int, C.c_int64, C.c_double, C.c_int, C.c_uint64,
                                         C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_uint64)]),
    "ddm_host_stream_peak": (C.c_int, [C.c_int, C.c_size_t, _dp]),
    "ddm_last_stats": (C.c_int, [_vp, C.POINTER(Stats)]),
    "ddm_last_output_dlpack": (C.c_int, [_vp, C.POINTER(C.POINTER(DLManagedTensor))]),
    "ddm_training_batch": (C.c_int, [_vp, C.c_int, C.c_int64, C.c_int64, C.c_double, C.c_int, C.c_uint64, C.c_uint64, C.c_int, _dp,
                                     C.POINTER(C.POINTER(DLManagedTensor))]),
    "ddm_last_output_device_ptr": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(C.c_size_t)]),
    "ddm_set_normals_debug": (C.c_int, [_vp, _dp, C.c_size_t, C.POINTER(C.c_int64), C.c_int64]),
    "ddm_export_normals": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                     C.c_int, _dp]),
    "ddm_normals_histogram": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_uint64), _dp]),
    "ddm_philox4x32": (C.c_int, [_vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int64]),
    "ddm_microbench": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp]),
    "ddm_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(_vp)]),
    "ddm_host_free": (C.c_int, [_vp]),
}


def library_path() -> str:
    # DDM_B200_LIB selects an experimental build of the same library (A/B measurements only);
    # DDM_PHILOX_ROUNDS=7 the 7-round build of the generator (1.16x the steps/s; same counters and maps, a different
    # stream: the known-answer tests and the CPU oracle are 10-round, its distribution suite passes -- DESIGN.md section 5)
    if os.environ.get("DDM_B200_LIB"):
        return os.environ["DDM_B200_LIB"]
    if os.environ.get("DDM_PHILOX_ROUNDS", "10") == "7":
        return _build.PHILOX7_LIB_PATH
    return _build.LIB_PATH


def load() -> C.CDLL:
    """dlopen libddm_b200.so and declare every prototype.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or bayesflow_nddms_b200._build.build()).  There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib
