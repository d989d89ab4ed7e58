"""Latency of the C3-sized host-array call (1024 datasets x 1000 trials, single_trial_alpha_not_scaled) by transfer mode."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesflow_nddms_b200 as pkg
from bayesflow_nddms_b200 import priors, single_trial_alpha_not_scaled as m1

sim = pkg.DDMSimulator(0, seed=2023)
P = priors.draw_prior_batch("alpha", 1024, np.random.default_rng(2023))
out_pageable = np.empty((1024, 1000, 2))
out_pinned = sim.pinned_empty((1024, 1000, 2), np.float64)


def timeit(fn, reps=30):
    fn(); fn()
    sim.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return np.median(ts) * 1e3, np.min(ts) * 1e3


order = sys.argv[1:] or ["auto", "off", "auto"]
for mode in order:
    sim.set_host_decode(0 if mode == "auto" else -1)
    for name, kw in (("fresh result array", {}), ("pageable out=", {"out": out_pageable}), ("pinned out=", {"out": out_pinned})):
        med, mn = timeit(lambda: m1.batch_simulate_trials(P, 1000, sim, **kw))
        st = sim.last_stats()
        print(f"host_decode={mode:5s} {name:20s} median {med:.3f} ms  min {mn:.3f} ms  kernel_ms {st['kernel_ms']:.3f} d2h {st['d2h_bytes']} threads {st['host_decode_threads']}", flush=True)
med, mn = timeit(lambda: m1.batch_simulate_trials_device(P, 1000, sim))
print(f"device-resident (DLPack)                median {med:.3f} ms  min {mn:.3f} ms")
