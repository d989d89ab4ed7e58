#!/usr/bin/env python
"""bench.py -- throughput of the DDM trial simulator hot path (BASELINE.json metric).

A "step" = one pass of the hot path over one batch of synthetic prior draws:
the throughput-sweep configuration C5: D datasets x 1000 trials of the
basic dcDDM, broad prior, tau = 0, dt = 1e-3, max_steps = 4000 (mean ~258 Euler steps/trial,
heavy-tailed).  Default D = 1 000 000 per GPU (1e9 trials per step per GPU), weak scaling:
every rank simulates its own D datasets at disjoint global dataset indices (disjoint Philox
counter ranges), no data-path collective.

  value     Euler steps/s, whole job, parameters resident in HBM, output left in HBM
  e2e       same metric through the reference-facing call with HOST buffers, as SURVEY.md section 8d specifies C5
            ("outputs reduced on device"): basic_ddm_dc.batch_simulate_histogram(params_host, n_trials)
            = H2D of the parameters, simulator kernel, histogram kernel, D2H of the histogram, all timed
  e2e_host_rows  the same batch delivered as (B, N, 2) host arrays (float64 as the reference returns them, and
            float32 as its configurator casts them), against this host's measured streaming-store rate
  roofline  issue-slot roofline of the stepping kernel (not HBM, not tensor: no contraction)
  configs   BASELINE.json's other configurations (C1, C3, C4) through the reference-facing calls and device-
            resident, with their kernels' roofline fractions; fp64 validation-mode throughput on C5
  allgather_batch  (N > 1) the one collective of the path: C3-sized training-batch shards -> every rank
  cpu_baseline  the CPU oracle port of the reference's numba loop on this box's host cores

`--impl reference` times that CPU implementation alone (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TRIALS = 1000
DT = 1e-3
MAX_STEPS = 4000
# Algorithmic issue slots per Euler step per lane of the algorithm as implemented:
# Philox4x32-10 block (20 IMAD.WIDE + 20 LOP3 + 2 counter adds) / 6 normals = 7, field extraction 12/6 = 2,
# Box-Muller 8 per pair = 4 (2 of them MUFU), Euler update FFMA + FADD + IADD + FSETP = 4.
# (SURVEY.md section 8d budgeted 28 for a 4-normals-per-block design.)
I_STEP = 17
I_STEP_SURVEY = 28
NCU_DRAM_BYTES_PER_LAUNCH_1E9 = 131_922_944 + 7_955_124_480  # profiles/r02_ncu_sweep_fullsize_details.txt (tile kernel)
MODEL_BASIC = 0
FLAG_OUT_F32 = 2


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--datasets", type=int, default=1_000_000, help="datasets per GPU per step")
    ap.add_argument("--e2e-datasets", type=int, default=0, help="datasets per GPU per e2e step (0 = same as --datasets if host memory allows)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 3)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-chunk-rows", type=int, default=-1, help="e2e leg: trials per streamed chunk (-1 = library default)")
    ap.add_argument("--host-decode", type=int, default=0,
                    help="e2e leg: host threads expanding the compact PCIe records (0 auto, < 0 float64 rows over PCIe)")
    ap.add_argument("--microbench", action="store_true", help="also measure per-pipe issue rates")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1/C3/C4/fp64 legs")
    ap.add_argument("--no-host-rows", action="store_true", help="skip the host-row legs of the end-to-end section")
    ap.add_argument("--kernel-variant", type=int, default=-1, help="-1 automatic (default: latency kernel up to 256 Ki trials, tile kernel above), 0 tile kernel, 1 the round-1 persistent kernel, 2 latency kernel")
    return ap.parse_args()


def sweep_params(n_datasets: int, seed: int = 2023) -> np.ndarray:
    from bayesflow_nddms_b200 import priors

    return priors.draw_prior_batch("sweep", n_datasets, np.random.default_rng(seed))


# ------------------------------------------------------------------------------------------------
# clocks during the timed region (pynvml; same fields as the nvidia-smi recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.power = [], set(), []
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    self.nv, "nvmlDeviceGetCurrentClocksEventReasons") else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
                self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self) -> dict:
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "power_w_median": float(np.median(self.power)) if self.power else None, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's loop, all host threads
# ------------------------------------------------------------------------------------------------
def cpu_run(n_datasets: int, threads: int, seed: int = 7):
    from oracle import cpu as orc

    params = sweep_params(n_datasets, seed=seed)
    t0 = time.perf_counter()
    _, steps, _ = orc.simulate_batch_mt(MODEL_BASIC, params, N_TRIALS, seed=seed, dt=DT, max_steps=float(MAX_STEPS),
                                        n_threads=threads, keep_output=True)
    dt = time.perf_counter() - t0
    return steps, n_datasets * N_TRIALS, dt


def cpu_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_sample_size(threads: int, target_s: float) -> int:
    """Datasets whose simulation takes about target_s on `threads` threads (calibrated)."""
    probe = max(threads * 2, 16)
    cpu_run(probe, threads)  # warm (page-in, thread start)
    steps, _, dt = cpu_run(probe, threads)
    rate = probe / max(dt, 1e-6)
    return int(max(probe, min(rate * target_s, 2_000_000)))


def cpu_baseline(target_s: float) -> dict:
    threads = cpu_threads()
    n = cpu_sample_size(threads, target_s)
    steps, trials, dt = cpu_run(n, threads)
    if dt < 0.6 * target_s:  # the small calibration probe under-estimates the rate: size the sample again
        n = int(min(n * target_s / max(dt, 1e-3), 4_000_000))
        steps, trials, dt = cpu_run(n, threads)
    out = {"value": steps / dt, "unit": "steps/s", "trials_per_s": trials / dt, "cores": threads, "kind": "port",
           "sample": f"{n} datasets x {N_TRIALS} trials of the same prior (C oracle of the numba loop: MT19937+polar normals, "
                     f"fp64), {threads} pthreads, {dt:.1f} s",
           "mean_steps_per_trial": steps / trials}
    # the same loop under the reference's own engine (numba), one thread as the reference runs it and numba-parallel
    try:
        from oracle import numba_loop as nl

        p1 = sweep_params(300, seed=11)
        s1, t1 = nl.time_numba(p1, N_TRIALS, DT, MAX_STEPS, parallel=False)
        pn = sweep_params(300 * min(threads, 16), seed=12)
        sn, tn = nl.time_numba(pn, N_TRIALS, DT, MAX_STEPS, parallel=True)
        out["numba_restatement"] = {"steps_per_s_1thread": s1 / t1, "steps_per_s_numba_parallel": sn / tn,
                                    "sample": f"300 datasets x {N_TRIALS} trials on one thread ({t1:.1f} s); "
                                              f"{pn.shape[0]} datasets numba-parallel ({tn:.1f} s); JIT compile excluded"}
    except Exception as e:
        out["numba_restatement"] = {"unavailable": repr(e)[:200]}
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu as orc

    orc.build()
    threads = cpu_threads()
    per_step_s = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    n = cpu_sample_size(threads, per_step_s)
    _, _, dt0 = cpu_run(n, threads)
    if dt0 < 0.6 * per_step_s:
        n = int(min(n * per_step_s / max(dt0, 1e-3), 4_000_000))
    for _ in range(args.warmup):
        cpu_run(n, threads)
    tot_steps = tot_trials = 0
    t0 = time.perf_counter()
    for k in range(args.steps):
        s, t, _ = cpu_run(n, threads, seed=100 + k)
        tot_steps += s
        tot_trials += t
    el = time.perf_counter() - t0
    val = tot_steps / el
    line = {
        "impl": "reference", "metric": "euler_steps_per_sec", "value": val, "unit": "steps/s",
        "trials_per_s": tot_trials / el, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(n, note="bounded sample per step; CPU arm runs on rank 0 only"),
        "cpu_baseline": {"value": val, "unit": "steps/s", "cores": threads, "kind": "port",
                         "sample": f"{n} datasets x {N_TRIALS} trials per step, {threads} pthreads"},
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(datasets_per_gpu: int, note: str = "") -> dict:
    c = {"workload": "C5 throughput sweep: basic dcDDM, broad prior (drift~N(0,2), alpha~TN(1,.5;0,10), beta~Beta(2,2), "
                     "dc~TN(1,.5;0,10), tau=0)",
         "datasets_per_gpu": datasets_per_gpu, "trials_per_dataset": N_TRIALS, "dt": DT, "max_steps": MAX_STEPS,
         "parallelism": "datasets sharded by rank, disjoint Philox counter ranges, no collective"}
    if note:
        c["note"] = note
    return c



# ------------------------------------------------------------------------------------------------
# BASELINE.json's other configurations (rank 0)
# ------------------------------------------------------------------------------------------------
def _median_ms(fn, sim, reps):
    fn()
    sim.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        sim.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


def other_configs(sim, issue_peak, want_cpu=True) -> dict:
    """C1 (1 x 300), C3 (1024 x 1000, per-trial boundary), C4 (Stahl-shaped, 19 374 trials) through the reference-facing
    calls (host arrays in and out) and device-resident, each with its kernel's time and roofline fraction; and C5's
    sweep in the fp64 validation mode (the reference's own precision)."""
    from bayesflow_nddms_b200 import _capi as capi
    from bayesflow_nddms_b200 import basic_ddm_dc as m0
    from bayesflow_nddms_b200 import imputation_from_stahl_not_scaled as stahl
    from bayesflow_nddms_b200 import priors
    from bayesflow_nddms_b200 import single_trial_alpha_not_scaled as m1

    def kernel_fields(st):
        k_s = st["kernel_ms"] * 1e-3
        d = {"kernel_ms": st["kernel_ms"], "euler_steps": st["total_steps"],
             "scheduler": {1: "round-1 persistent kernel", 2: "tile kernel", 3: "latency kernel"}.get(st["scheduler"], "generic")}
        if k_s > 0:
            d["kernel_steps_per_s"] = st["total_steps"] / k_s
            d["kernel_roofline_frac"] = st["total_steps"] * I_STEP / 32.0 / k_s / issue_peak
        return d

    out = {}
    orc = None
    if want_cpu:
        from oracle import cpu as orc  # the checker, timed beside (cpu_baseline leg)

    # ---- C1: one prior draw x 300 trials (reference defaults dt=.01/400, and the dt=1e-3 BASELINE.json names)
    p = np.array([3.0, 1.5, 0.5, 0.4, 1.0])          # single_trial_alpha_not_scaled.py:853-883's fixed vector, basic layout
    for dt, ms, key in ((0.01, 400, "C1_dt0.01"), (0.001, 4000, "C1_dt0.001")):
        g = _median_ms(lambda: m0.batch_simulate_trials(p[None], 300, sim, dt=dt, max_steps=ms)[0], sim, 200)
        st = sim.last_stats()
        dev_ms = _median_ms(lambda: sim.run(0, p[None], 300, dt, ms, flags=capi.FLAG_OUT_F32), sim, 200)
        out[key] = {"workload": f"basic_ddm_dc: one parameter vector x 300 trials, dt={dt}, max_steps={ms}",
                    "call_ms": g, "call": "simulate_trials(params (5,), 300) -> (300, 2) float64 host array",
                    "device_resident_ms": dev_ms, "trials_per_s_call": 300 / (g * 1e-3), **kernel_fields(st)}
        if orc is not None:
            t0 = time.perf_counter()
            for _ in range(20):
                orc.simulate_batch_mt(0, p[None], 300, dt=dt, max_steps=float(ms))
            out[key]["cpu_port_1thread_ms"] = (time.perf_counter() - t0) / 20 * 1e3

    # ---- C3: single_trial_alpha_not_scaled, 1024 datasets x 1000 trials, dt=.01
    P = priors.draw_prior_batch("alpha", 1024, np.random.default_rng(2023))
    g = _median_ms(lambda: m1.batch_simulate_trials(P, 1000, sim), sim, 30)

    def c3_device():
        b = m1.batch_simulate_trials_device(P, 1000, sim)
        del b

    dev_ms = _median_ms(c3_device, sim, 50)
    st = sim.last_stats()   # of the resident launch (the host call streams the rows in two chunks: its event window spans copies)
    out["C3"] = {"workload": "single_trial_alpha_not_scaled: 1024 datasets x 1000 trials, per-trial boundary, dt=.01, max_steps=400",
                 "call_ms": g, "call": "batch_simulate_trials(params (1024,7), 1000) -> (1024, 1000, 2) float64 host array (16 MB, pinned pool)",
                 "device_resident_ms": dev_ms, "device_call": "batch_simulate_trials_device -> DLPack (1024, 1000, 2) float32",
                 "steps_per_s_call": st["total_steps"] / (g * 1e-3), "steps_per_s_device_resident": st["total_steps"] / (dev_ms * 1e-3),
                 **kernel_fields(st)}
    # the same model at a launch size that fills the GPU (20 000 datasets): the kernel's own rate in this regime
    Pb = priors.draw_prior_batch("alpha", 20_000, np.random.default_rng(2))
    best = None
    for _ in range(3):
        sim.run(1, Pb, 1000, 0.01, 400, flags=capi.FLAG_OUT_F32)
        stb = sim.last_stats()
        best = stb if best is None or stb["kernel_ms"] < best["kernel_ms"] else best
    out["C3_model_at_2e7_trials"] = {"workload": "the same model and dt, 20 000 datasets x 1000 trials (short trials: 43 steps each)",
                                     "mean_steps_per_trial": best["total_steps"] / best["n_trials"], **kernel_fields(best)}
    if orc is not None:
        t0 = time.perf_counter()
        orc.simulate_batch_mt(1, P[:128], 1000)
        out["C3"]["cpu_port_1thread_ms"] = (time.perf_counter() - t0) * 8 * 1e3
        out["C3"]["cpu_port_note"] = "128 of the 1024 datasets timed on one thread, scaled by 8"

    # ---- C4: Stahl-shaped imputation (19 374 trials, 89 participants), fed to training via DLPack
    subj, pe = stahl.synthetic_stahl_like()
    pp = stahl.draw_participant_params(89, np.random.default_rng(2024))

    def c4_device():
        t, _, _ = stahl.impute_dataset(subj, pe, pp, simulator=sim, device=True)
        del t

    g_dev = _median_ms(c4_device, sim, 50)
    st = sim.last_stats()
    g_host = _median_ms(lambda: stahl.impute_dataset(subj, pe, pp, simulator=sim), sim, 50)
    out["C4"] = {"workload": "imputation_from_stahl_not_scaled: 19 374 Stahl-shaped trials, 89 participants, per-trial boundary from the EEG "
                             "channel, dt=.01 (synthetic data with the CSV's shape)",
                 "call_ms": g_host, "call": "impute_dataset(subj_idx, pre_Pe, participant params) -> (19374, 2) float64 host array",
                 "device_resident_ms": g_dev, "device_call": "impute_dataset(..., device=True) -> torch tensor over the DLPack capsule",
                 "trials_per_s_device": 19374 / (g_dev * 1e-3), **kernel_fields(st)}
    # where the call's time goes: the reference's own NumPy preprocessing (the participant index of the subject column, the Pe
    # standardisation; imputation_from_stahl_not_scaled.py:59-105) against the library call that replaces its per-row loop
    _, alphas4 = stahl.boundaries_from_pe(pe)
    _, idx4 = stahl.unique_inverse(subj)
    out["C4"]["simulate_call_ms"] = _median_ms(lambda: sim.simulate_trialwise(idx4, alphas4, pp), sim, 50)
    out["C4"]["simulate_call"] = "ddm_simulate_trialwise(group (n,) i32, bound (n,) f64, params (89,4)) -> (n, 2) float64 host array"
    t0 = time.perf_counter()
    for _ in range(50):
        stahl.boundaries_from_pe(pe)
        stahl.unique_inverse(subj)
    out["C4"]["host_numpy_preprocessing_ms"] = (time.perf_counter() - t0) / 50 * 1e3
    if orc is not None:
        _, alphas = stahl.boundaries_from_pe(pe)
        _, idx = np.unique(subj, return_inverse=True)
        t0 = time.perf_counter()
        for i in range(0, 19374, 97):   # every 97th trial through the scalar oracle, scaled up
            orc.simulate_mt(5, pp[idx[i]], 1, seed=i, bound_in=[alphas[i]])
        out["C4"]["cpu_port_1thread_ms"] = (time.perf_counter() - t0) * 97 * 1e3

    # ---- evidence-path variant: 200 observed states per trial, per-trial z-score, device-resident float32 rows
    from bayesflow_nddms_b200 import basic_ddm_dc_evidence as mev
    Pe = mev.batch_draw_prior(2048)
    best = None
    for _ in range(4):
        b = sim.simulate_evidence(Pe, 1000, 200, 1, flags=capi.FLAG_OUT_F32, device=True)
        del b
        ste = sim.last_stats()
        best = ste if best is None or ste["kernel_ms"] < best["kernel_ms"] else best
    out["evidence_path"] = {
        "workload": "retired_models/basic_ddm_dc_evidence: 2048 datasets x 1000 trials, dt=.001, 200 observed evidence states + noise "
                    "per trial, z-scored, (rt, choice, path[200]) float32 rows device-resident (808 B per trial)",
        "kernel_ms": best["kernel_ms"], "kernels": "record_kernel (step + record) + evidence_post_kernel (noise, z-score, rows)",
        "trials_per_s": best["n_trials"] / (best["kernel_ms"] * 1e-3), "steps_per_s": best["total_steps"] / (best["kernel_ms"] * 1e-3),
        "output_gb_per_s": best["n_trials"] * 808 / (best["kernel_ms"] * 1e-3) / 1e9}

    # ---- C5 in the fp64 validation mode (the reference's own precision): bounded sample
    Ps = sweep_params(2000, seed=99)
    best = None
    for _ in range(2):
        sim.run(0, Ps, N_TRIALS, DT, MAX_STEPS, precision=64, flags=0)
        st64 = sim.last_stats()
        best = st64 if best is None or st64["kernel_ms"] < best["kernel_ms"] else best
    sim.run(0, Ps, N_TRIALS, DT, MAX_STEPS, precision=32, flags=capi.FLAG_OUT_F32)
    st32 = sim.last_stats()
    out["C5_fp64_validation_mode"] = {
        "workload": "the sweep's prior, 2000 datasets x 1000 trials, precision 64: reference operation order in double, fp64 "
                    "Box-Muller (libdevice log/sincospi), one thread per trial (generic kernel), float64 rows",
        "kernel_ms": best["kernel_ms"], "steps_per_s": best["total_steps"] / (best["kernel_ms"] * 1e-3),
        "fp32_production_steps_per_s_same_batch": st32["total_steps"] / (st32["kernel_ms"] * 1e-3),
        "note": "B200 issues fp64 at a fraction of the fp32 rate and the validation kernel has no lane refill; it exists to check "
                "the production kernel, not to be fast"}
    return out


PHILOX7_PROBE = r"""
import json, sys
import numpy as np
sys.path.insert(0, %r)
import bayesflow_nddms_b200 as pkg
from bayesflow_nddms_b200 import priors
sim = pkg.DDMSimulator(device=%d, seed=2023)
rounds = sim._lib.ddm_philox_rounds()
P = priors.draw_prior_batch("sweep", 100_000, np.random.default_rng(2023))
best = None
for _ in range(4):
    sim.run(0, P, 1000, 1e-3, 4000, flags=2)
    st = sim.last_stats()
    best = st if best is None or st["kernel_ms"] < best["kernel_ms"] else best
print(json.dumps({"rounds": rounds, "kernel_ms": best["kernel_ms"], "steps": best["total_steps"],
                  "steps_per_s": best["total_steps"] / (best["kernel_ms"] * 1e-3)}))
"""


def philox_rounds_leg(device: int) -> dict:
    """What the generator's round count costs: the sweep kernel (1e8 trials) from the shipped 10-round library and from
    the measurement-only 7-round build (ddm_rng.cuh: DDM_PHILOX_ROUNDS), each in its own process (DDM_B200_LIB)."""
    import subprocess

    from bayesflow_nddms_b200 import _build

    out = {}
    for name, lib in (("philox10_shipped", _build.LIB_PATH), ("philox7_variant", _build.PHILOX7_LIB_PATH)):
        if not os.path.exists(lib):
            out[name] = {"unavailable": os.path.basename(lib) + " not built"}
            continue
        env = dict(os.environ, DDM_B200_LIB=lib)
        r = subprocess.run([sys.executable, "-c", PHILOX7_PROBE % (ROOT, device)], env=env, capture_output=True, text=True, timeout=300)
        try:
            out[name] = json.loads(r.stdout.strip().split("\n")[-1])
        except Exception:
            out[name] = {"error": (r.stderr or r.stdout)[-300:]}
    a, b = out.get("philox10_shipped", {}), out.get("philox7_variant", {})
    if "steps_per_s" in a and "steps_per_s" in b:
        out["speedup_7_over_10"] = b["steps_per_s"] / a["steps_per_s"]
    out["note"] = ("10 rounds is the shipped generator (Random123 / cuRAND default) and what every other figure uses; 7 rounds is the "
                   "fewest that pass BigCrush (Salmon et al. 2011) and is built only to measure the cost of the 20 IMAD.WIDE per block")
    return out


def allgather_leg(sim, rank, world, dev, stream, barrier, max_over_ranks) -> dict:
    """the path's only collective.  Every rank simulates its shard of a C3-sized training batch
    (single_trial_alpha_not_scaled: 1024 x 1000 x 2 float32) and the shards are all-gathered so that the training
    rank holds the batch (bayesflow_nddms_b200.distributed.all_gather_batch, NCCL).  Timed on the device, max over ranks."""
    import torch
    import torch.distributed as dist

    from bayesflow_nddms_b200 import distributed as D
    from bayesflow_nddms_b200 import priors
    from bayesflow_nddms_b200 import single_trial_alpha_not_scaled as m1

    B, N = 1024, 1000
    P = priors.draw_prior_batch("alpha", B, np.random.default_rng(77))     # same draws on every rank
    lo, hi = D.shard_range(B, rank, world)
    local = torch.from_dlpack(m1.batch_simulate_trials_device(P[lo:hi], N, sim, dataset_offset=5000 + lo))
    torch.cuda.synchronize()
    full = D.all_gather_batch(local, B, world)                              # warm-up (communicator set-up)
    ok = True
    if rank == 0:   # the assembled batch is what one GPU produces alone, bit for bit
        alone = torch.from_dlpack(m1.batch_simulate_trials_device(P, N, sim, dataset_offset=5000))
        ok = bool(torch.equal(alone, full))
        del alone
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    barrier()
    e0.record()
    for _ in range(reps):
        full = D.all_gather_batch(local, B, world)
    e1.record()
    torch.cuda.synchronize()
    us = max_over_ranks(e0.elapsed_time(e1)) / reps * 1e3
    # one step of the data-parallel generator: simulate the shard + gather (the simulator call blocks the host until
    # its batch is ready for hand-off, so this one is a host clock around the loop, device drained on both sides)
    for _ in range(5):
        local = torch.from_dlpack(m1.batch_simulate_trials_device(P[lo:hi], N, sim, dataset_offset=5000 + lo))
        full = D.all_gather_batch(local, B, world)
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        local = torch.from_dlpack(m1.batch_simulate_trials_device(P[lo:hi], N, sim, dataset_offset=5000 + lo))
        full = D.all_gather_batch(local, B, world)
    torch.cuda.synchronize()
    step_us = max_over_ranks((time.perf_counter() - t0) * 1e6) / reps
    nbytes = B * N * 2 * 4
    recv = nbytes * (world - 1) / world
    return {"workload": "C3-sized training batch: (1024/R, 1000, 2) float32 shards -> (1024, 1000, 2) on every rank",
            "collective": "torch.distributed.all_gather_into_tensor (NCCL)", "batch_bytes": nbytes,
            "bytes_received_per_rank": int(recv), "us_per_allgather": us,
            "algorithm_gbs_per_rank": recv / (us * 1e-6) / 1e9, "bus_gbs": nbytes * (world - 1) / world / (us * 1e-6) / 1e9,
            "us_per_simulate_plus_allgather": step_us, "assembled_batch_equals_single_gpu_batch": ok,
            "note": "8 MB: latency-bound on NVLink 5 / NVSwitch, as SURVEY.md section 5 expects"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    import bayesflow_nddms_b200 as pkg
    from bayesflow_nddms_b200 import basic_ddm_dc

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def min_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    D = args.datasets
    sim = pkg.DDMSimulator(device=local_rank, seed=2023)
    sim.set_kernel_variant(args.kernel_variant)
    numa_bound = sim.bind_host_thread_near_gpu() if world > 1 else False  # host staging on the GPU's own NUMA node
    stream = torch.cuda.Stream(device=dev)
    sim.set_stream(stream.cuda_stream)
    params = sweep_params(D, seed=2023 + rank)
    ds_base = rank * D

    # ---- device-resident throughput ("value") ----------------------------------------------------
    lib, ctx = sim._lib, sim._ctx
    sim._check(lib.ddm_upload_params(ctx, MODEL_BASIC, params.ctypes.data_as(pkg._capi._dp), D, 5))

    def step(k):
        sim._check(lib.ddm_run(ctx, N_TRIALS, DT, MAX_STEPS, sim.seed, ds_base + 0, 32, FLAG_OUT_F32))

    for k in range(args.warmup):
        step(k)
    sim.synchronize()
    st0 = sim.last_stats()
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, launches = [], 0
    with torch.cuda.stream(stream):
        e0.record(stream)
        for k in range(args.steps):
            step(k)
        e1.record(stream)
    torch.cuda.synchronize()
    barrier()
    clk = clocks.stop()
    elapsed_ms = max_over_ranks(e0.elapsed_time(e1))
    st = sim.last_stats()  # every step simulates the same trials: per-step counts are identical
    launches = st["kernel_launches"] * args.steps
    steps_per_step = sum_over_ranks(float(st["total_steps"]))
    trials_per_step = float(D) * N_TRIALS * world
    value = steps_per_step * args.steps / (elapsed_ms * 1e-3)
    trials_per_s = trials_per_step * args.steps / (elapsed_ms * 1e-3)

    # dominant kernel, timed live with CUDA events on its own stream (events inside ddm_run)
    for k in range(3):
        step(k)
        kernel_ms.append(sim.last_stats()["kernel_ms"])
    k_ms = float(np.mean(kernel_ms))

    # the batch is reduced on the device (RT histogram by choice + missing count), nothing is downloaded
    sim.last_histogram(401, 4.01)
    t_h = time.perf_counter()
    hist = sim.last_histogram(401, 4.01)
    hist_ms = (time.perf_counter() - t_h) * 1e3
    hist_total = int(hist["upper"].sum() + hist["lower"].sum()) + hist["missing"] + hist["overflow"]
    reduction = {"kind": "RT histogram, 401 bins of 10 ms x 2 boundaries + missing, over the resident float32 rows",
                 "ms": hist_ms, "accounts_for_every_trial": bool(hist_total == D * N_TRIALS),
                 "agrees_with_kernel_counters": bool(int(hist["upper"].sum()) == st["n_upper"] and hist["missing"] == st["n_timeouts"])}

    # ---- roofline -----------------------------------------------------------------------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    f_hz = (clk.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)) * 1e6
    issue_peak = sm_count * 4 * f_hz                           # warp-instructions/s: 1 per SMSP per clock
    ach = st["total_steps"] * I_STEP / 32.0 / (k_ms * 1e-3)    # algorithmic warp-instructions/s of this rank's kernel
    out_bytes = D * N_TRIALS * 8 + D * 40 + D * 32             # f32 pair out + fp64 params in + fp32 constants
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    roofline = {
        "bound": "issue", "achieved": ach / 1e12, "peak": issue_peak / 1e12, "unit": "T warp-inst/s",
        "frac": ach / issue_peak, "frac_at_survey_28_slots": ach / issue_peak * I_STEP_SURVEY / I_STEP,
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch at this very size, from the ncu --set full capture
        # profiles/r02_ncu_sweep_fullsize_details.txt (scripts/r02_gpu_ncu_full.sh); other sizes: not captured
        "traffic": NCU_DRAM_BYTES_PER_LAUNCH_1E9 if (D == 1_000_000 and N_TRIALS == 1000 and st["scheduler"] == 2) else None,
        "traffic_note": "ncu --set full of this kernel at the bench's launch size (1e9 trials): dram read 0.132 GB + write "
                        "7.955 GB = 8.087 GB against 8.072 GB algorithmic (8 B/trial out + 72 B/dataset in): 1.00x (the tile "
                        "kernel writes rows a tile at a time, lane-contiguous; ~45 MB of the last rows are still in L2 at kernel end)",
        "kernel": ("ddm::persistent_kernel<KIND_FIXED, f32 out>" if st["scheduler"] == 1 else "ddm::tile_kernel<KIND_FIXED, f32 out>"),
        "kernel_ms": k_ms,
        "steps_per_launch": st["total_steps"], "issue_slots_per_step": I_STEP,
        "peak_how": f"{sm_count} SMs x 4 schedulers x {f_hz / 1e6:.0f} MHz (median SM clock sampled during the timed region)",
        "steps_per_s_kernel": st["total_steps"] / (k_ms * 1e-3),
        "hbm": {"algorithmic_bytes_per_launch": out_bytes, "achieved_gbs": out_bytes / (k_ms * 1e-3) / 1e9,
                "peak_gbs": hbm_peak, "frac": out_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
    }
    # the stepping loop's own ceiling on this chip: one Philox block + three Box-Muller pairs + six Euler steps per warp in
    # a bare loop (no refill, no output), measured now.  The loop is bound by the FMA-heavy pipe (20 IMAD.WIDE of ~4.6
    # cycles each per block): this rate does not depend on occupancy or on instruction-level parallelism.
    try:
        blocks_per_s, _hz = sim.microbench(pkg._capi.MB_NAMES.index("sim_block"), 2048)
        ceil_steps = blocks_per_s * 32 * 6
        roofline["bare_loop_ceiling_steps_per_s"] = ceil_steps
        roofline["frac_of_bare_loop_ceiling"] = st["total_steps"] / (k_ms * 1e-3) / ceil_steps
    except Exception as e:
        roofline["bare_loop_ceiling_steps_per_s"] = None
        roofline["bare_loop_note"] = repr(e)[:120]
    if args.microbench and rank == 0:
        mb = {}
        for i, name in enumerate(pkg._capi.MB_NAMES):
            ips, hz = sim.microbench(i, 2048)
            mb[name] = {"warp_inst_per_s": ips, "sm_mhz": hz / 1e6, "per_smsp_per_clk": ips / (sm_count * 4 * hz)}
        roofline["microbench"] = mb

    # ---- end to end through the reference-facing calls, host buffers on both sides ----------------------
    e2e, e2e_rows = None, None
    if not args.no_e2e:
        import psutil

        N_BINS, RT_MAX = 401, 4.01
        pe = sweep_params(D, seed=4000 + rank)   # host (pageable) parameters, as a caller holds them
        ke = args.e2e_steps or min(args.steps, 3)

        # (1) C5 as specified: parameters in, batch simulated and reduced on the device, histogram out
        def hist_step():
            return basic_ddm_dc.batch_simulate_histogram(pe, N_TRIALS, sim, dt=DT, max_steps=MAX_STEPS, dataset_offset=ds_base,
                                                         n_bins=N_BINS, rt_max=RT_MAX)

        h = hist_step()
        st_h = sim.last_stats()
        h_total = int(h["upper"].sum() + h["lower"].sum()) + h["missing"] + h["overflow"]
        barrier()
        with torch.cuda.stream(stream):
            e0.record(stream)
            for k in range(ke):
                h = hist_step()
            e1.record(stream)
        torch.cuda.synchronize()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        tot = sum_over_ranks(float(st_h["total_steps"]))
        e2e = {"value": tot * ke / (ms * 1e-3), "unit": "steps/s", "trials_per_s": D * N_TRIALS * world * ke / (ms * 1e-3),
               "h2d_bytes_per_step": int(pe.nbytes), "d2h_bytes_per_step": int((2 * N_BINS + 2) * 8), "steps": ke,
               "datasets_per_gpu": D, "ms_per_step": ms / ke,
               "result_accounts_for_every_trial": bool(h_total == D * N_TRIALS),
               "result_agrees_with_kernel_counters": bool(int(h["upper"].sum()) == st_h["n_upper"] and h["missing"] == st_h["n_timeouts"]),
               "api": "basic_ddm_dc.batch_simulate_histogram(params (B,5) f64 host, n_trials) -> RT histogram by boundary (401 bins "
                      "of 10 ms x 2 + missing + overflow) as host arrays; one ddm_simulate_histogram call: H2D of the parameters, "
                      "prep + stepping kernel (float32 rows stay in HBM), histogram kernel, D2H of 804 counters -- C5 as SURVEY.md "
                      "section 8d specifies it (outputs reduced on device).  From 64 Mi trials on the call works through a few "
                      "chunks of datasets (ddm_histogram_chunks: six at 1e9 trials) so that uploads and reductions run beside "
                      "the neighbouring chunks' kernels; same resident rows, same histogram"}
        # the chunked call counts its histogram launches itself; the single-launch form adds one
        hist_call_launches = st_h["kernel_launches"] + (1 if st_h["kernel_launches"] <= 2 else 0)
        e2e["launches_per_call"] = hist_call_launches
        launches += hist_call_launches * ke

        # (2) the same batch delivered as host rows: float64 (what simulate_trials returns) and float32 (what the
        # reference's configurator turns it into at once, basic_ddm_dc.py:146), against the host's streaming-store rate
        if not args.no_host_rows:
            De = args.e2e_datasets or D
            local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
            avail = psutil.virtual_memory().available
            while De > 1000 and De * N_TRIALS * 16 * local_world * 3 > avail:
                De //= 2
            pr = pe[:De]
            sim.set_host_decode(args.host_decode)
            sim.set_pipeline(-1, args.e2e_chunk_rows)
            e2e_rows = {"datasets_per_gpu": De, "h2d_bytes_per_step": int(pr.nbytes)}
            for label, dtype, fl in (("float64", np.float64, 0), ("float32", np.float32, FLAG_OUT_F32)):
                # page-locked destination; every rank must get one (the legs below contain barriers): agree, else halve
                out_host, why = None, ""
                for _attempt in range(4):
                    try:
                        out_host = sim.pinned_empty((De, N_TRIALS, 2), dtype, slot="rows_" + label)
                    except Exception as e:
                        out_host, why = None, repr(e)[:200]
                    if min_over_ranks(0.0 if out_host is None else 1.0) > 0.5:
                        break
                    out_host = None
                    sim._pinned.pop("rows_" + label, None)
                    De //= 2
                    pr = pe[:De]
                if out_host is None:
                    e2e_rows[label] = {"error": "no rank-wide page-locked destination: " + why}
                    continue
                e2e_rows["datasets_per_gpu"] = De
                basic_ddm_dc.batch_simulate_trials(pr, N_TRIALS, sim, dt=DT, max_steps=MAX_STEPS, dataset_offset=ds_base, out=out_host,
                                                   flags=fl)
                st_r = sim.last_stats()
                barrier()
                with torch.cuda.stream(stream):
                    e0.record(stream)
                    for k in range(ke):
                        basic_ddm_dc.batch_simulate_trials(pr, N_TRIALS, sim, dt=DT, max_steps=MAX_STEPS, dataset_offset=ds_base,
                                                           out=out_host, flags=fl)
                    e1.record(stream)
                torch.cuda.synchronize()
                barrier()
                ms_r = max_over_ranks(e0.elapsed_time(e1))
                tot_r = sum_over_ranks(float(st_r["total_steps"]))
                e2e_rows[label] = {"value": tot_r * ke / (ms_r * 1e-3), "unit": "steps/s", "ms_per_step": ms_r / ke,
                                   "datasets_per_gpu": De, "d2h_bytes_per_step": int(st_r["d2h_bytes"] or out_host.nbytes),
                                   "host_rows_bytes_per_step": int(out_host.nbytes),
                                   "host_decode_threads": int(st_r["host_decode_threads"]),
                                   "host_row_write_gbs_all_ranks": out_host.nbytes * world * ke / (ms_r * 1e-3) / 1e9}
                launches += st_r["kernel_launches"] * ke
                del out_host
                sim._pinned.pop("rows_" + label, None)   # give the page-locked block back before the next layout takes its own
            # the ceiling of the row-writing side: every rank's decode threads filling memory with the decode's own
            # streaming stores, all ranks at once (they share the host's memory controllers)
            barrier()
            t0 = time.perf_counter()
            peak = sim.host_stream_peak(args.host_decode if args.host_decode > 0 else 0, 1 << 30)
            peak_all = sum_over_ranks(peak)
            e2e_rows["host_stream_store_peak_gbs_all_ranks"] = peak_all / 1e9
            e2e_rows["host_stream_store_peak_how"] = ("ddm_host_stream_peak: each rank's decode threads fill 1 GiB with non-temporal 32-byte "
                                                     "stores, best of 5 passes, all ranks concurrently; sum over ranks")
            for label in ("float64", "float32"):
                if "host_row_write_gbs_all_ranks" in e2e_rows.get(label, {}):
                    e2e_rows[label]["frac_of_host_store_peak"] = e2e_rows[label]["host_row_write_gbs_all_ranks"] / (peak_all / 1e9)
            e2e_rows["api"] = ("basic_ddm_dc.batch_simulate_trials(params (B,5) f64 host, n_trials[, flags=OUT_F32]) -> (B, n_trials, 2) pinned "
                               "host array; one ddm_simulate call: chunked kernels overlapped with the D2H of the previous chunk; trials "
                               "cross PCIe as 4-byte (steps, choice) records and host threads write the rows (rt = n*dt + ndt, choice)")
            e2e_rows["host_thread_bound_near_gpu"] = bool(numa_bound)
            sim.set_pipeline(-1, -1)

    # ---- BASELINE.json's other configurations, through the reference-facing calls and device-resident ----------
    configs = None
    if rank == 0 and not args.no_configs:
        try:
            configs = other_configs(sim, issue_peak, want_cpu=(not args.no_cpu_baseline and world == 1))
        except Exception as e:  # secondary figures must not take the headline down
            configs = {"error": repr(e)[:300]}

    if rank == 0 and world == 1 and not args.no_configs and isinstance(configs, dict):
        try:
            configs["philox_rounds"] = philox_rounds_leg(local_rank)
        except Exception as e:
            configs["philox_rounds"] = {"error": repr(e)[:300]}

    # ---- the one collective of the path: all-gather of a training batch's shards (N > 1) ------------------------
    allgather = None
    if world > 1:
        try:
            allgather = allgather_leg(sim, rank, world, dev, stream, barrier, max_over_ranks)
        except Exception as e:
            allgather = {"error": repr(e)[:300]}

    # ---- BASELINE config 2: one online-training batch (64 datasets x 500 trials, dt=.01), latency -------------
    training_batch = None
    if rank == 0:
        try:
            from bayesflow_nddms_b200 import _capi as capi

            B2, N2, reps = 64, 500, 200
            for _ in range(10):
                sim.draw_prior("basic", B2)
                sim.run_uploaded(N2, 0.01, 400, flags=capi.FLAG_OUT_F32)
            sim.synchronize()
            blocks = []
            for _ in range(5):                                                    # five blocks of 40 batches: the median block
                t0 = time.perf_counter()
                for _ in range(reps // 5):
                    pd, batch = sim.training_batch("basic", B2, N2, 0.01, 400)    # device prior, simulate, (64,5) draws out, hand-off
                    del batch
                blocks.append((time.perf_counter() - t0) / (reps // 5))
            lat = float(np.median(blocks))
            blocks3 = []
            for _ in range(5):                                                    # the same batch through the three separate calls
                t0 = time.perf_counter()
                for _ in range(reps // 5):
                    pd = sim.draw_prior("basic", B2)                              # device prior + (64,5) copy out
                    sim.run_uploaded(N2, 0.01, 400, flags=capi.FLAG_OUT_F32)      # simulate on the resident draws
                    batch = sim.last_output_dlpack()                              # hand-off (syncs the stream)
                    del batch
                blocks3.append((time.perf_counter() - t0) / (reps // 5))
            training_batch = {"workload": "C2 basic_ddm_dc online-training batch: 64 datasets x 500 trials, dt=.01, max_steps=400",
                              "path": "ddm_training_batch: device prior -> simulate -> draws to the host + DLPack hand-off of the "
                                      "device-resident f32 batch, one call and one stream synchronisation",
                              "ms_per_batch": lat * 1e3, "ms_per_batch_blocks_of_40": [b * 1e3 for b in blocks],
                              "ms_per_batch_three_calls": float(np.median(blocks3)) * 1e3,
                              "trials_per_s": B2 * N2 / lat, "launches_per_batch": 2}
            if not args.no_cpu_baseline and world == 1:
                from oracle import cpu as orc

                pp = sweep_params(B2, seed=5)
                pp[:, 3] = 0.3
                t0 = time.perf_counter()
                for _ in range(5):
                    orc.simulate_batch_mt(MODEL_BASIC, pp, N2, seed=1, dt=0.01, max_steps=400.0, n_threads=1)
                training_batch["cpu_port_1thread_ms_per_batch"] = (time.perf_counter() - t0) / 5 * 1e3
        except Exception as e:  # a secondary figure must not take the headline down
            training_batch = {"error": repr(e)}

    line = {
        "metric": "euler_steps_per_sec", "value": value, "unit": "steps/s", "trials_per_s": trials_per_s,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(D, note="output 8 GB per step per GPU streams through L2 (>> 126 MB), parameters 40 MB; "
                                          "no L2 flush needed: the kernel is issue-bound, not memory-bound"),
        "mean_steps_per_trial": steps_per_step / trials_per_step,
        "timeout_frac": sum_over_ranks(float(st["n_timeouts"])) / trials_per_step,
        "kernel": {"grid": st["grid"], "block": st["block"], "refill_threshold": st["refill_threshold"], "tile": st["tile"],
                   "persistent": bool(st["used_persistent"]), "scheduler": {1: "round-1 persistent kernel", 2: "tile kernel", 3: "latency kernel"}.get(st["scheduler"], "generic")},
        "clocks": clk, "roofline": roofline, "gpu_launches": launches,
    }
    line["device_reduction"] = reduction
    if e2e is not None:
        line["e2e"] = e2e
    if e2e_rows is not None:
        line["e2e_host_rows"] = e2e_rows
    if configs is not None:
        line["configs"] = configs
    if allgather is not None:
        line["allgather_batch"] = allgather
    if training_batch is not None:
        line["training_batch"] = training_batch
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(args.cpu_seconds)
    if rank == 0:
        print(json.dumps(line), flush=True)
    sim.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
