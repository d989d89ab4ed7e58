"""Host-side handle on the CUDA trial simulator (one ddm_ctx per instance).

This is the only place that talks to the C ABI.  The model modules
(basic_ddm_dc, single_trial_alpha_not_scaled, imputation_from_stahl_not_scaled) keep the
reference's call signatures and delegate here.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _capi
from .dlpack import DeviceBatch


class DDMError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"[ddm status {status}] {message}")
        self.status = status


class _PinnedBlock:
    """One page-locked host block handed out as a result array's base object; goes back to its pool when the
    last array over it is collected."""

    def __init__(self, pool, ptr: int, cap: int):
        self._pool, self._ptr, self._cap = pool, ptr, cap
        self.__array_interface__ = {"shape": (cap,), "typestr": "|u1", "data": (ptr, False), "version": 3}

    def __del__(self):
        try:
            self._pool._give_back(self._ptr, self._cap)
        except Exception:  # interpreter shutdown
            pass


class _FreeingPool:
    """Owner of a single page-locked block (``DDMSimulator.pinned_empty``): the block is freed when the last
    array over it goes away."""

    def __init__(self, lib):
        self._lib = lib

    def _give_back(self, ptr: int, cap: int):
        self._lib.ddm_host_free(C.c_void_p(ptr))


class _PinnedResultPool:
    """Result arrays of mid-size host-destined batches live in page-locked memory: a device-to-host copy into a
    fresh pageable numpy array runs at a third of the PCIe rate (1.1 ms instead of 0.45 ms for the 16 MB of a
    1024 x 1000 batch).  Blocks are power-of-two sized and reused once the previous result has been dropped;
    the pool never pins more than ``limit`` bytes, beyond that results are ordinary numpy arrays."""

    MIN_BYTES, MAX_BYTES = 256 << 10, 512 << 20

    def __init__(self, lib, limit: int = 2 << 30):
        self._lib, self._limit = lib, int(limit)
        self._free: dict[int, list[int]] = {}
        self._pinned_bytes = 0
        self._lock = threading.RLock()  # re-entrant: a garbage collection inside empty() may run a block's __del__
        self._closed = False

    def empty(self, shape, dtype) -> np.ndarray | None:
        dtype = np.dtype(dtype)
        count = int(np.prod(shape, dtype=np.int64))
        nbytes = count * dtype.itemsize
        if not (self.MIN_BYTES <= nbytes <= self.MAX_BYTES) or self._closed:
            return None
        cap = 1 << (nbytes - 1).bit_length()
        with self._lock:
            blocks = self._free.get(cap)
            ptr = blocks.pop() if blocks else None
            if ptr is None:
                if self._pinned_bytes + cap > self._limit:
                    return None
                self._pinned_bytes += cap
        if ptr is None:
            p = C.c_void_p()
            if self._lib.ddm_host_alloc(cap, C.byref(p)) != _capi.OK or not p.value:
                with self._lock:
                    self._pinned_bytes -= cap
                return None
            ptr = p.value
        raw = np.asarray(_PinnedBlock(self, ptr, cap))
        return raw[:nbytes].view(dtype).reshape(shape)

    def _give_back(self, ptr: int, cap: int):
        with self._lock:
            if not self._closed:
                self._free.setdefault(cap, []).append(ptr)
                return
            self._pinned_bytes -= cap
        self._lib.ddm_host_free(C.c_void_p(ptr))

    def close(self):
        with self._lock:
            self._closed = True
            blocks = [(p, cap) for cap, ps in self._free.items() for p in ps]
            self._free = {}
            self._pinned_bytes -= sum(cap for _, cap in blocks)
        for p, _ in blocks:
            self._lib.ddm_host_free(C.c_void_p(p))


class DDMSimulator:
    """Owns a device context.  ``seed`` keys Philox; ``dataset_counter`` is the global
    index of the next dataset, so successive batches draw from disjoint counter ranges
    and any batch can be regenerated from (seed, its first dataset index)."""

    def __init__(self, device: int = 0, seed: int = 2023):
        self._lib = _capi.load()
        self._ctx = C.c_void_p()
        rc = self._lib.ddm_create(int(device), C.byref(self._ctx))
        if rc != _capi.OK:
            msg = self._lib.ddm_last_error(None)
            self._ctx = C.c_void_p()
            raise DDMError(rc, (msg or b"").decode())
        self.device = int(device)
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.dataset_counter = 0
        self.trial_counter = 0  # trialwise (Stahl) path: global index of the next trial, see simulate_trialwise
        self._pinned = {}
        self._results = _PinnedResultPool(self._lib)

    # ---- plumbing ---------------------------------------------------------------------
    def close(self):
        ctx, self._ctx = getattr(self, "_ctx", None), C.c_void_p()
        if ctx:
            self._pinned = {}   # blocks are freed as the last arrays over them are collected
            self._results.close()
            self._lib.ddm_destroy(ctx)

    def __del__(self):
        try:  # the interpreter may already be tearing the library down
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc == _capi.OK:
            return
        msg = (self._lib.ddm_last_error(self._ctx) or b"").decode()
        if rc == _capi.ERR_NEGATIVE_BOUND:
            # imputation_from_stahl_not_scaled.py:124-125 raises ValueError
            raise ValueError(msg)
        if rc == _capi.ERR_INVALID:
            raise ValueError(msg)
        raise DDMError(rc, msg)

    def set_tuning(self, refill_threshold: int = 0, blocks_per_sm: int = 0, tile: int = 0):
        self._check(self._lib.ddm_set_tuning(self._ctx, refill_threshold, blocks_per_sm, tile))

    def set_kernel_variant(self, variant: int = -1):
        """-1 (default): the tile-staged persistent kernel, and the latency kernel (one thread per trial, speculative
        six-step blocks) for launches of at most 256 Ki trials; 0: the tile kernel at any size; 1: the round-1
        persistent kernel; 2: the latency kernel at any size.  Bit-identical results (A/B measurements, tests)."""
        self._check(self._lib.ddm_set_kernel_variant(self._ctx, int(variant)))

    def set_stream(self, cuda_stream_ptr: int | None):
        self._check(self._lib.ddm_set_stream(self._ctx, C.c_void_p(cuda_stream_ptr or 0)))

    def synchronize(self):
        self._check(self._lib.ddm_synchronize(self._ctx))

    def bind_host_thread_near_gpu(self) -> bool:
        """Pin the calling thread to the CPUs closest to this simulator's GPU (NVML's ideal affinity), so that
        pinned staging buffers allocated afterwards are first-touched on the GPU's own NUMA node.  On a
        multi-GPU box this keeps every rank's PCIe traffic off the inter-socket link.  Returns False where
        NVML or the container's cpuset does not allow it."""
        try:
            import pynvml

            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(self.device))
            return True
        except Exception:
            return False

    def pinned_empty(self, shape, dtype=np.float64, slot: str = "out") -> np.ndarray:
        """A numpy array over page-locked host memory, reused per slot (full-rate D2H).  The block belongs to
        the arrays handed out over it: asking the slot for more than it holds allocates a new block and the
        old one is freed when the last array over it is collected (never under a live array)."""
        dtype = np.dtype(dtype)
        count = int(np.prod(shape, dtype=np.int64))
        nbytes = count * dtype.itemsize
        blk = self._pinned.get(slot)
        if blk is None or blk._cap < nbytes:
            p = C.c_void_p()
            rc = self._lib.ddm_host_alloc(max(nbytes, 1), C.byref(p))
            if rc != _capi.OK:
                why = (self._lib.ddm_last_error(None) or b"").decode()
                raise DDMError(rc, f"pinned host allocation of {nbytes} bytes failed: {why}")
            blk = _PinnedBlock(_FreeingPool(self._lib), p.value, max(nbytes, 1))
            self._pinned[slot] = blk
        return np.asarray(blk)[:nbytes].view(dtype).reshape(shape)

    # ---- the hot path -----------------------------------------------------------------
    @staticmethod
    def _roll(counter: int, n: int) -> int:
        """A launch must not straddle a multiple of 2^32 of the global index (its low 32 bits are one Philox
        counter word, the high part is launch-uniform): skip to the next multiple when it would."""
        if (counter & 0xFFFFFFFF) + int(n) > 1 << 32:
            counter = ((counter >> 32) + 1) << 32
        return counter

    def _next_offset(self, n_datasets: int, dataset_offset):
        if dataset_offset is None:
            dataset_offset = self._roll(self.dataset_counter, n_datasets)
            self.dataset_counter = dataset_offset + int(n_datasets)
        return int(dataset_offset)

    def draw_prior(self, prior: str, n_draws: int, *, seed=None, draw_offset=None, to_host: bool = True):
        """Device-side batched prior (``ddm_draw_prior``): (n_draws, P) float64 in the reference's column
        order.  The draws stay in the simulator's parameter arena, so ``run_uploaded`` can simulate them
        without a host round trip; ``to_host=False`` skips the copy and returns None."""
        pid, cols = _capi.PRIORS[prior]
        off = self._next_offset(n_draws, draw_offset)
        out = np.empty((int(n_draws), cols), dtype=np.float64) if to_host else None
        self._check(self._lib.ddm_draw_prior(self._ctx, pid, int(n_draws),
                                             self.seed if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF, off,
                                             out.ctypes.data_as(_capi._dp) if to_host else None))
        self._prior_offset = off
        return out

    def run_uploaded(self, n_trials: int, dt: float = 0.01, max_steps: int = 400, *, seed=None, dataset_offset=None,
                     precision: int = 32, flags: int = 0):
        """Launch on the parameters already resident on the device (after ``draw_prior``); the datasets
        are keyed by the same global indices as the draws unless ``dataset_offset`` is given."""
        off = getattr(self, "_prior_offset", 0) if dataset_offset is None else int(dataset_offset)
        self._check(self._lib.ddm_run(self._ctx, int(n_trials), float(dt), int(max_steps),
                                      self.seed if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF, off,
                                      int(precision), int(flags)))

    def training_batch(self, prior: str, n_draws: int, n_trials: int, dt: float = 0.01, max_steps: int = 400, *, seed=None,
                       draw_offset=None, flags: int = _capi.FLAG_OUT_F32, to_host: bool = True):
        """One online-training batch in one call (``ddm_training_batch``): device prior -> simulation on the resident
        draws -> DLPack hand-off, with the (n_draws, P) draws copied out for the trainer's targets.  Returns
        ``(prior_draws, DeviceBatch)``; the same bits as ``draw_prior`` + ``run_uploaded`` + ``last_output_dlpack``."""
        pid, cols = _capi.PRIORS[prior]
        off = self._next_offset(n_draws, draw_offset)
        draws = np.empty((int(n_draws), cols), dtype=np.float64) if to_host else None
        m = C.POINTER(_capi.DLManagedTensor)()
        self._check(self._lib.ddm_training_batch(
            self._ctx, pid, int(n_draws), int(n_trials), float(dt), int(max_steps),
            self.seed if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF, off, int(flags),
            draws.ctypes.data_as(_capi._dp) if to_host else None, C.byref(m)))
        self._prior_offset = off
        t = m.contents.dl_tensor
        return draws, DeviceBatch(m, [t.shape[i] for i in range(t.ndim)], t.dtype.bits, t.device.device_id)

    def run(self, model: int, params, n_trials: int, dt: float = 0.01, max_steps: int = 400, *, seed=None,
            dataset_offset=None, precision: int = 32, flags: int = 0):
        """Upload (B, P) parameters and launch; results stay on the device."""
        params = np.ascontiguousarray(params, dtype=np.float64)
        if params.ndim == 1:
            params = params[None, :]
        if params.ndim != 2:
            raise ValueError("params must be (P,) or (B, P)")
        B, P = params.shape
        off = self._next_offset(B, dataset_offset)
        self._check(self._lib.ddm_upload_params(self._ctx, int(model), params.ctypes.data_as(_capi._dp), B, P))
        self._check(self._lib.ddm_run(self._ctx, int(n_trials), float(dt), int(max_steps),
                                      self.seed if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF, off,
                                      int(precision), int(flags)))
        return B

    def download(self, shape, f32: bool, out: np.ndarray | None = None) -> np.ndarray:
        dtype = np.float32 if f32 else np.float64
        if out is None:
            out = np.empty(shape, dtype=dtype)
        elif out.dtype != dtype or out.shape != tuple(shape) or not out.flags.c_contiguous:
            raise ValueError("out must be C-contiguous with the result's shape and dtype")
        self._check(self._lib.ddm_download(self._ctx, out.ctypes.data_as(C.c_void_p)))
        return out

    def simulate(self, model: int, params, n_trials: int, dt: float = 0.01, max_steps: int = 400, *, seed=None,
                 dataset_offset=None, precision: int = 32, flags: int = 0, out: np.ndarray | None = None) -> np.ndarray:
        """B datasets x n_trials -> numpy (B, n_trials, 2), float64 unless FLAG_OUT_F32.

        One ``ddm_simulate`` call: upload, launch, copy back.  Large batches are streamed to the host
        chunk by chunk with the PCIe copy of one chunk overlapping the kernel of the next; pass a
        ``pinned_empty`` array as ``out`` for full-rate asynchronous copies."""
        params = np.ascontiguousarray(params, dtype=np.float64)
        if params.ndim == 1:
            params = params[None, :]
        if params.ndim != 2:
            raise ValueError("params must be (P,) or (B, P)")
        B, P = params.shape
        shape = (B, int(n_trials), _capi.N_COLS.get(int(model), 2))
        dtype = np.float32 if flags & _capi.FLAG_OUT_F32 else np.float64
        if out is None:
            out = self._results.empty(shape, dtype)
            if out is None:
                out = np.empty(shape, dtype=dtype)
        elif out.dtype != dtype or out.shape != shape or not out.flags.c_contiguous:
            raise ValueError("out must be C-contiguous with the result's shape and dtype")
        off = self._next_offset(B, dataset_offset)
        self._check(self._lib.ddm_simulate(self._ctx, int(model), params.ctypes.data_as(_capi._dp), B, P, int(n_trials),
                                           float(dt), int(max_steps),
                                           self.seed if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF, off,
                                           int(precision), int(flags), out.ctypes.data_as(C.c_void_p)))
        return out

    def set_pipeline(self, min_rows: int = -1, chunk_rows: int = -1):
        """Host-destined batches of at least ``min_rows`` trials are streamed in chunks of about
        ``chunk_rows`` trials (values < 0: defaults)."""
        self._check(self._lib.ddm_set_pipeline(self._ctx, int(min_rows), int(chunk_rows)))

    def set_host_decode(self, n_threads: int = 0):
        """Streamed two-column batches cross PCIe as 4- or 8-byte (steps, choice[, draw]) records that
        ``n_threads`` host threads expand into the float64 rows (0: automatic, < 0: ship float64 rows)."""
        self._check(self._lib.ddm_set_host_decode(self._ctx, int(n_threads)))

    def simulate_device(self, model: int, params, n_trials: int, dt: float = 0.01, max_steps: int = 400, *,
                        seed=None, dataset_offset=None, precision: int = 32, flags: int = _capi.FLAG_OUT_F32) -> DeviceBatch:
        """Same, but the batch stays in HBM and is returned as a DLPack producer."""
        self.run(model, params, n_trials, dt, max_steps, seed=seed, dataset_offset=dataset_offset,
                 precision=precision, flags=flags)
        return self.last_output_dlpack()

    def simulate_trialwise(self, group, bound, group_params, dt: float = 0.01, max_steps: int = 400, *, seed=None,
                           trial_offset: int | None = None, precision: int = 32, flags: int = 0, device: bool = False):
        """Per-trial supplied boundary, parameters gathered by group (Stahl imputation).

        Trial i is keyed by the global trial index ``trial_offset + i``.  With ``trial_offset=None`` the
        simulator's ``trial_counter`` supplies it and advances by n, so successive calls -- the reference calls
        ``diffusion_trial`` once per CSV row (imputation_from_stahl_not_scaled.py:205-213) -- draw fresh noise,
        as ``dataset_counter`` does for the dataset-wise models; pass an explicit offset to regenerate a call."""
        group = np.ascontiguousarray(group, dtype=np.int32).ravel()
        bound = np.ascontiguousarray(bound, dtype=np.float64).ravel()
        group_params = np.ascontiguousarray(group_params, dtype=np.float64)
        if group_params.ndim == 1:
            group_params = group_params[None, :]
        if group_params.ndim != 2 or group_params.shape[1] != 4:
            raise ValueError("group_params must be (G, 4) = (drift, beta, ter, dc)")
        if group.size != bound.size:
            raise ValueError("group and bound must have one entry per trial")
        n = group.size
        if trial_offset is None:
            trial_offset = self._roll(self.trial_counter, n)
            self.trial_counter = trial_offset + int(n)
        f32 = bool(flags & _capi.FLAG_OUT_F32)
        out = None if device else np.empty((n, 2), dtype=np.float32 if f32 else np.float64)
        self._check(self._lib.ddm_simulate_trialwise(
            self._ctx, group.ctypes.data_as(C.POINTER(C.c_int32)), bound.ctypes.data_as(_capi._dp),
            group_params.ctypes.data_as(_capi._dp), n, group_params.shape[0], float(dt), int(max_steps),
            self.seed if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF, int(trial_offset), int(precision),
            int(flags), None if device else out.ctypes.data_as(C.c_void_p)))
        return self.last_output_dlpack() if device else out

    def simulate_evidence(self, params, n_trials: int, n_obs: int = 200, standardize: int = 1, dt: float = 0.001,
                          max_steps: int = 4000, *, seed=None, dataset_offset=None, precision: int = 32, flags: int = 0,
                          device: bool = False):
        """Evidence-path variants: (B, 6) parameters -> (B, n_trials, 2 + n_obs): rt, choice, observed path."""
        params = np.ascontiguousarray(params, dtype=np.float64)
        if params.ndim == 1:
            params = params[None, :]
        if params.ndim != 2 or params.shape[1] != 6:
            raise ValueError("params must be (6,) or (B, 6) = drift, boundary, beta, tau, dc, sigma1")
        B = params.shape[0]
        f32 = bool(flags & _capi.FLAG_OUT_F32)
        out = None if device else np.empty((B, int(n_trials), 2 + int(n_obs)), dtype=np.float32 if f32 else np.float64)
        off = self._next_offset(B, dataset_offset)
        self._check(self._lib.ddm_simulate_evidence(
            self._ctx, params.ctypes.data_as(_capi._dp), B, int(n_trials), int(n_obs), int(standardize), float(dt),
            int(max_steps), self.seed if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF, off, int(precision), int(flags),
            None if device else out.ctypes.data_as(C.c_void_p)))
        return self.last_output_dlpack() if device else out

    # ---- results of the last run --------------------------------------------------------
    def last_output_dlpack(self) -> DeviceBatch:
        m = C.POINTER(_capi.DLManagedTensor)()
        self._check(self._lib.ddm_last_output_dlpack(self._ctx, C.byref(m)))
        t = m.contents.dl_tensor
        shape = [t.shape[i] for i in range(t.ndim)]
        return DeviceBatch(m, shape, t.dtype.bits, t.device.device_id)

    def last_output_device_ptr(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self._lib.ddm_last_output_device_ptr(self._ctx, C.byref(p), C.byref(n)))
        return p.value, n.value

    def simulate_exact(self, params, n_trials: int, *, seed=None, dataset_offset=None, device: bool = False):
        """Exact first-passage sampler (``ddm_simulate_exact``; pyhddmjagsutils.py:47-176): params (B, 8) or (8,)
        = [Alpha, Tau, Nu, Beta, rangeTau, rangeBeta, Eta, Varsigma] -> (B, n_trials) float64 signed response
        times.  ``device=True`` leaves the batch on the GPU and returns a DLPack producer of shape (B, n_trials, 1)."""
        params = np.ascontiguousarray(params, dtype=np.float64)
        if params.ndim == 1:
            params = params[None, :]
        if params.ndim != 2 or params.shape[1] != 8:
            raise ValueError("params must be (8,) or (B, 8): Alpha, Tau, Nu, Beta, rangeTau, rangeBeta, Eta, Varsigma")
        B = params.shape[0]
        out = None if device else np.empty((B, int(n_trials)), dtype=np.float64)
        off = self._next_offset(B, dataset_offset)
        self._check(self._lib.ddm_simulate_exact(self._ctx, params.ctypes.data_as(_capi._dp), B, int(n_trials),
                                                 self.seed if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF, off,
                                                 None if device else out.ctypes.data_as(_capi._dp)))
        return self.last_output_dlpack() if device else out

    def last_steps(self, n: int) -> np.ndarray:
        out = np.empty(int(n), dtype=np.int32)
        self._check(self._lib.ddm_last_steps(self._ctx, out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def last_stats(self) -> dict:
        st = _capi.Stats()
        self._check(self._lib.ddm_last_stats(self._ctx, C.byref(st)))
        return st.as_dict()

    def last_histogram(self, n_bins: int = 400, rt_max: float = 4.0) -> dict:
        """Response-time histogram of the last resident batch, reduced on the device: counts per
        |rt| bin for upper- and lower-boundary responses, trials without a response, and responses
        beyond ``rt_max``.  Additive over datasets, shards and GPUs."""
        h = np.zeros(2 * int(n_bins) + 2, dtype=np.uint64)
        self._check(self._lib.ddm_last_output_histogram(self._ctx, int(n_bins), float(rt_max),
                                                        h.ctypes.data_as(C.POINTER(C.c_uint64))))
        return {"upper": h[:n_bins].copy(), "lower": h[n_bins:2 * n_bins].copy(), "missing": int(h[2 * n_bins]),
                "overflow": int(h[2 * n_bins + 1]), "edges": np.linspace(0.0, float(rt_max), int(n_bins) + 1)}

    def simulate_histogram(self, model: int, params, n_trials: int, dt: float = 0.01, max_steps: int = 400, *, seed=None,
                           dataset_offset=None, precision: int = 32, flags: int = 0, n_bins: int = 400,
                           rt_max: float = 4.0) -> dict:
        """Simulate (B, P) host parameters and return only the response-time histogram, reduced on the device
        (``ddm_simulate_histogram``): the throughput sweep's own output.  The float32
        rows stay resident (``last_output_dlpack``)."""
        params = np.ascontiguousarray(params, dtype=np.float64)
        if params.ndim == 1:
            params = params[None, :]
        if params.ndim != 2:
            raise ValueError("params must be (P,) or (B, P)")
        B, P = params.shape
        off = self._next_offset(B, dataset_offset)
        h = np.zeros(2 * int(n_bins) + 2, dtype=np.uint64)
        self._check(self._lib.ddm_simulate_histogram(
            self._ctx, int(model), params.ctypes.data_as(_capi._dp), B, P, int(n_trials), float(dt), int(max_steps),
            self.seed if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF, off, int(precision), int(flags), int(n_bins),
            float(rt_max), h.ctypes.data_as(C.POINTER(C.c_uint64))))
        return {"upper": h[:n_bins].copy(), "lower": h[n_bins:2 * n_bins].copy(), "missing": int(h[2 * n_bins]),
                "overflow": int(h[2 * n_bins + 1]), "edges": np.linspace(0.0, float(rt_max), int(n_bins) + 1)}

    def host_stream_peak(self, n_threads: int = 0, nbytes: int = 1 << 30) -> float:
        """Bytes per second this box's host threads reach with the compact wire decode's streaming stores."""
        v = C.c_double()
        self._check(self._lib.ddm_host_stream_peak(int(n_threads), int(nbytes), C.byref(v)))
        return v.value

    # ---- parity hooks ---------------------------------------------------------------------
    def set_normals_debug(self, z, offsets):
        """Shared-increment mode: trial t consumes z[offsets[t]:] in the reference's order."""
        if z is None:
            self._check(self._lib.ddm_set_normals_debug(self._ctx, None, 0, None, 0))
            return
        z = np.ascontiguousarray(z, dtype=np.float64).ravel()
        offsets = np.ascontiguousarray(offsets, dtype=np.int64).ravel()
        self._check(self._lib.ddm_set_normals_debug(self._ctx, z.ctypes.data_as(_capi._dp), z.size,
                                                    offsets.ctypes.data_as(C.POINTER(C.c_int64)), offsets.size))

    def export_normals(self, dataset: int, trial: int, stream: int, first: int, count: int, *, seed=None,
                       precision: int = 32) -> np.ndarray:
        out = np.empty(int(count), dtype=np.float64)
        self._check(self._lib.ddm_export_normals(self._ctx, self.seed if seed is None else int(seed), int(dataset),
                                                 int(trial), int(stream), int(first), int(count), int(precision),
                                                 out.ctypes.data_as(_capi._dp)))
        return out

    def normals_histogram(self, n_normals: int, n_bins_abs: int = 112, z_max: float = 5.6, n_bins_angle: int = 256, *, seed=None):
        """|z| and pair-angle histograms and raw moments of ``n_normals`` production-map normals, reduced on the device."""
        h = np.zeros(int(n_bins_abs) + 1 + int(n_bins_angle), dtype=np.uint64)
        m = np.zeros(4, dtype=np.float64)
        self._check(self._lib.ddm_normals_histogram(self._ctx, self.seed if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF,
                                                    int(n_normals), int(n_bins_abs), float(z_max), int(n_bins_angle),
                                                    h.ctypes.data_as(C.POINTER(C.c_uint64)), m.ctypes.data_as(_capi._dp)))
        n = 6 * ((int(n_normals) + 5) // 6)
        return {"n": n, "abs": h[:n_bins_abs].copy(), "beyond": int(h[n_bins_abs]), "angle": h[n_bins_abs + 1:].copy(),
                "edges": np.linspace(0.0, float(z_max), int(n_bins_abs) + 1), "moments": m / n}

    def philox4x32(self, ctr, key) -> np.ndarray:
        ctr = np.ascontiguousarray(ctr, dtype=np.uint32).reshape(-1, 4)
        key = np.ascontiguousarray(key, dtype=np.uint32).reshape(-1, 2)
        if key.shape[0] != ctr.shape[0]:
            raise ValueError("one key per counter")
        out = np.empty_like(ctr)
        u32p = C.POINTER(C.c_uint32)
        self._check(self._lib.ddm_philox4x32(self._ctx, ctr.ctypes.data_as(u32p), key.ctypes.data_as(u32p),
                                             out.ctypes.data_as(u32p), ctr.shape[0]))
        return out

    def microbench(self, which: int, iters: int = 4096):
        ips, hz = C.c_double(), C.c_double()
        self._check(self._lib.ddm_microbench(self._ctx, int(which), int(iters), C.byref(ips), C.byref(hz)))
        return ips.value, hz.value


_default = None
_default_lock = threading.Lock()


def default_simulator() -> DDMSimulator:
    """Process-wide simulator on the current rank's GPU (LOCAL_RANK, else device 0)."""
    import os

    global _default
    with _default_lock:
        if _default is None:
            _default = DDMSimulator(device=int(os.environ.get("LOCAL_RANK", "0")),
                                    seed=int(os.environ.get("DDM_SEED", "2023")))
            # one process per GPU: every rank draws its own batches, so each starts in its own 2^40-dataset
            # range of the global index (same seed, disjoint Philox counters); rank 0 starts at 0 as before
            rank = int(os.environ.get("RANK", "0"))
            _default.dataset_counter = rank << 40
            _default.trial_counter = rank << 40
        return _default


def set_default_simulator(sim: DDMSimulator | None):
    global _default
    with _default_lock:
        _default = sim
