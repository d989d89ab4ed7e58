"""ctypes binding of oracle/ddm_oracle.c -- TEST INFRASTRUCTURE ONLY.

See oracle/__init__.py for who may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libddm_oracle.so")

MODEL_BASIC = 0
MODEL_ALPHA = 1
MODEL_ALPHA_DC = 2
MODEL_ALPHA_SCALE = 3
MODEL_ALPHA_SCALE2 = 4
MODEL_TRIALWISE = 5
MODEL_ETA = 6
MODEL_GENERAL = 7

FLAG_TIMEOUT_CHOICE_ONE = 1

_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "ddm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp = C.POINTER(C.c_double)
        i64p = C.POINTER(C.c_int64)
        i32p = C.POINTER(C.c_int32)
        L.orc_n_params.restype = C.c_int
        L.orc_n_params.argtypes = [C.c_int]
        L.orc_n_cols.restype = C.c_int
        L.orc_n_cols.argtypes = [C.c_int]
        L.orc_mt_normals.restype = None
        L.orc_mt_normals.argtypes = [C.c_uint32, dp, C.c_size_t]
        L.orc_philox4x32_10.restype = None
        L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.orc_philox_normals6.restype = None
        L.orc_philox_normals6.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, dp]
        L.orc_simulate_buffer.restype = C.c_int
        L.orc_simulate_buffer.argtypes = [C.c_int, dp, C.c_int64, dp, C.c_double, C.c_double, C.c_int,
                                          dp, C.c_int64, dp, i64p, i32p, dp, dp, i64p]
        L.orc_simulate_mt.restype = C.c_int
        L.orc_simulate_mt.argtypes = [C.c_int, dp, C.c_int64, dp, C.c_double, C.c_double, C.c_int,
                                      C.c_uint32, dp, i64p, i32p, dp, dp]
        L.orc_simulate_philox.restype = C.c_int
        L.orc_simulate_philox.argtypes = [C.c_int, dp, C.c_int64, dp, C.c_double, C.c_double, C.c_int,
                                          C.c_uint64, C.c_uint32, C.c_uint32, dp, i64p, i32p, dp, dp]
        L.orc_simulate_batch_mt.restype = C.c_int
        L.orc_simulate_batch_mt.argtypes = [C.c_int, dp, C.c_int64, C.c_int64, C.c_double, C.c_double,
                                            C.c_int, C.c_uint32, C.c_int, dp, i64p, i64p]
        L.orc_simulate_evidence.restype = C.c_int
        L.orc_simulate_evidence.argtypes = [dp, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, dp,
                                            C.c_int64, C.c_uint64, C.c_uint32, C.c_uint32, dp, i64p, i64p]
        _lib = L
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


@dataclass
class TrialTable:
    """Per-trial results of one dataset."""
    sim_data: np.ndarray   # (n, 2) f64 -- the reference's stacked output
    n_steps: np.ndarray    # (n,) i64
    choice: np.ndarray     # (n,) i32  (+1/-1/0)
    evidence: np.ndarray   # (n,) f64 final evidence
    bound: np.ndarray      # (n,) f64 boundary the trial used
    consumed: np.ndarray | None = None  # (n,) i64 normals consumed (buffer source)


def n_params(model: int) -> int:
    return lib().orc_n_params(model)


def n_cols(model: int) -> int:
    return lib().orc_n_cols(model)


def _prep(model, params, n_trials, bound_in):
    params = np.ascontiguousarray(params, dtype=np.float64).ravel()
    if params.size != n_params(model):
        raise ValueError(f"model {model} takes {n_params(model)} parameters, got {params.size}")
    if bound_in is not None:
        bound_in = np.ascontiguousarray(bound_in, dtype=np.float64).ravel()
        if bound_in.size != n_trials:
            raise ValueError("bound_in must have n_trials entries")
    out = np.empty((n_trials, n_cols(model)), np.float64)
    ns = np.empty(n_trials, np.int64)
    ch = np.empty(n_trials, np.int32)
    ev = np.empty(n_trials, np.float64)
    bd = np.empty(n_trials, np.float64)
    return params, bound_in, out, ns, ch, ev, bd


def _check(rc):
    if rc == -1:
        raise IndexError("normals buffer exhausted")
    if rc == -2:
        raise ValueError("Trial-level boundary cannot be less than zero")
    if rc != 0:
        raise RuntimeError(f"oracle error {rc}")


def simulate_buffer(model, params, n_trials, normals, dt=0.01, max_steps=400.0, flags=0, bound_in=None):
    """Reference loop fed an explicit array of standard normals (shared increments)."""
    params, bound_in, out, ns, ch, ev, bd = _prep(model, params, n_trials, bound_in)
    normals = np.ascontiguousarray(normals, dtype=np.float64)
    cons = np.empty(n_trials, np.int64)
    rc = lib().orc_simulate_buffer(model, _dp(params), n_trials, _dp(bound_in), dt, float(max_steps), flags,
                                   _dp(normals), normals.size, _dp(out),
                                   ns.ctypes.data_as(C.POINTER(C.c_int64)),
                                   ch.ctypes.data_as(C.POINTER(C.c_int32)), _dp(ev), _dp(bd),
                                   cons.ctypes.data_as(C.POINTER(C.c_int64)))
    _check(rc)
    return TrialTable(out, ns, ch, ev, bd, cons)


def simulate_mt(model, params, n_trials, seed, dt=0.01, max_steps=400.0, flags=0, bound_in=None):
    """Reference loop on the MT19937/polar stream numba uses (seeded in-jit)."""
    params, bound_in, out, ns, ch, ev, bd = _prep(model, params, n_trials, bound_in)
    rc = lib().orc_simulate_mt(model, _dp(params), n_trials, _dp(bound_in), dt, float(max_steps), flags,
                               int(seed) & 0xFFFFFFFF, _dp(out),
                               ns.ctypes.data_as(C.POINTER(C.c_int64)),
                               ch.ctypes.data_as(C.POINTER(C.c_int32)), _dp(ev), _dp(bd))
    _check(rc)
    return TrialTable(out, ns, ch, ev, bd)


def simulate_philox(model, params, n_trials, seed, dataset=0, trial_offset=0, dt=0.01, max_steps=400.0,
                    flags=0, bound_in=None):
    """Reference loop on the fp64-ideal normals of the CUDA kernels' Philox stream."""
    params, bound_in, out, ns, ch, ev, bd = _prep(model, params, n_trials, bound_in)
    rc = lib().orc_simulate_philox(model, _dp(params), n_trials, _dp(bound_in), dt, float(max_steps), flags,
                                   int(seed), int(dataset), int(trial_offset), _dp(out),
                                   ns.ctypes.data_as(C.POINTER(C.c_int64)),
                                   ch.ctypes.data_as(C.POINTER(C.c_int32)), _dp(ev), _dp(bd))
    _check(rc)
    return TrialTable(out, ns, ch, ev, bd)


def simulate_batch_mt(model, params, n_trials, seed=0, dt=0.01, max_steps=400.0, flags=0, n_threads=1,
                      keep_output=True):
    """B datasets x n_trials on n_threads host threads (the CPU baseline).

    Returns (sim_data (B, n_trials, 2) or None, total_steps, total_timeouts)."""
    params = np.ascontiguousarray(params, dtype=np.float64)
    B = params.shape[0]
    if params.shape[1] != n_params(model):
        raise ValueError("bad parameter count")
    out = np.empty((B, n_trials, n_cols(model)), np.float64) if keep_output else None
    steps = C.c_int64(0)
    touts = C.c_int64(0)
    rc = lib().orc_simulate_batch_mt(model, _dp(params), B, n_trials, dt, float(max_steps), flags,
                                     int(seed) & 0xFFFFFFFF, int(n_threads), _dp(out),
                                     C.byref(steps), C.byref(touts))
    _check(rc)
    return out, steps.value, touts.value


def mt_normals(seed: int, n: int) -> np.ndarray:
    out = np.empty(n, np.float64)
    lib().orc_mt_normals(int(seed) & 0xFFFFFFFF, _dp(out), n)
    return out


def philox4x32_10(ctr, key) -> np.ndarray:
    c = (C.c_uint32 * 4)(*[int(x) & 0xFFFFFFFF for x in ctr])
    k = (C.c_uint32 * 2)(*[int(x) & 0xFFFFFFFF for x in key])
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return np.array(list(o), dtype=np.uint32)


def philox_normals(seed, dataset, trial, stream, first, count) -> np.ndarray:
    """fp64-ideal normals first..first+count-1 of one (dataset, trial, stream)."""
    out = np.empty(count, np.float64)
    z = (C.c_double * 6)()
    cached = -1
    # 64-bit global dataset index: low word = counter word 3, bits 32..55 ride in the stream word (word 0) (ddm_rng.cuh)
    dataset = int(dataset)
    stream = int(stream) | ((dataset >> 32) << 8)
    dataset &= 0xFFFFFFFF
    for i in range(count):
        idx = first + i
        if idx // 6 != cached:
            cached = idx // 6
            lib().orc_philox_normals6(int(seed), cached, int(trial), int(dataset), int(stream), z)
        out[i] = z[idx % 6]
    return out


def simulate_evidence(params, n_trials, n_obs=200, mode=1, dt=0.001, max_steps=4000.0, *, normals=None, mt_seed=None,
                      philox_seed=None, dataset=0, trial_offset=0):
    """Evidence-path variants (retired_models/basic_ddm_dc_evidence*.py).  params = [drift, boundary, beta,
    tau, dc, sigma1]; mode 0 raw noisy path, 1 per-trial z-score, 2 dataset-level standardisation.
    Exactly one normal source: ``normals`` (buffer), ``mt_seed`` or ``philox_seed``.
    Returns (sim_data (n_trials, 2 + n_obs), n_steps, consumed)."""
    params = np.ascontiguousarray(params, dtype=np.float64).ravel()
    if params.size != 6:
        raise ValueError("evidence model takes 6 parameters")
    out = np.empty((n_trials, 2 + n_obs), np.float64)
    ns = np.empty(n_trials, np.int64)
    cons = np.empty(n_trials, np.int64)
    if normals is not None:
        normals = np.ascontiguousarray(normals, dtype=np.float64)
        kind, seed, nn = 0, 0, normals.size
    elif mt_seed is not None:
        kind, seed, nn = 1, int(mt_seed) & 0xFFFFFFFF, 0
    else:
        kind, seed, nn = 2, int(philox_seed), 0
    rc = lib().orc_simulate_evidence(_dp(params), n_trials, int(n_obs), int(mode), dt, float(max_steps), kind,
                                     _dp(normals), nn, seed, int(dataset), int(trial_offset), _dp(out),
                                     ns.ctypes.data_as(C.POINTER(C.c_int64)), cons.ctypes.data_as(C.POINTER(C.c_int64)))
    _check(rc)
    return out, ns, cons
