"""Wall-clock of BASELINE.json's configurations C1-C4 through the reference-facing Python API
(host arrays in, host arrays out), next to the CPU oracle port on one thread (how the reference runs)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesflow_nddms_b200 as pkg  # noqa: E402
from bayesflow_nddms_b200 import basic_ddm_dc as m0  # noqa: E402
from bayesflow_nddms_b200 import imputation_from_stahl_not_scaled as stahl  # noqa: E402
from bayesflow_nddms_b200 import priors  # noqa: E402
from bayesflow_nddms_b200 import single_trial_alpha_not_scaled as m1  # noqa: E402
from oracle import cpu as orc  # noqa: E402

sim = pkg.DDMSimulator(0, seed=2023)
rng = np.random.default_rng(2023)


def timeit(fn, reps):
    fn()
    sim.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    sim.synchronize()
    return (time.perf_counter() - t0) / reps


rows = []


def add(name, gpu_s, cpu_s, trials, steps):
    rows.append(dict(config=name, gpu_ms=gpu_s * 1e3, cpu_port_1thread_ms=cpu_s * 1e3, speedup=cpu_s / gpu_s, trials=trials,
                     euler_steps=steps, gpu_trials_per_s=trials / gpu_s))
    print(json.dumps(rows[-1]), flush=True)


# C1: one prior draw x 300 trials, reference defaults and the dt = 1e-3 variant BASELINE.json names
p = np.array([3.0, 1.5, 0.5, 0.4, 1.0])
for dt, ms in ((0.01, 400), (0.001, 4000)):
    g = timeit(lambda: m0.batch_simulate_trials(p[None], 300, sim, dt=dt, max_steps=ms), 200)
    steps = sim.last_stats()["total_steps"]
    t0 = time.perf_counter()
    for _ in range(20):
        orc.simulate_batch_mt(0, p[None], 300, dt=dt, max_steps=float(ms))
    add(f"C1 basic_ddm_dc 1x300 dt={dt}", g, (time.perf_counter() - t0) / 20, 300, steps)

# C2: 64 datasets x 500 trials from the prior
P = priors.draw_prior_batch("basic", 64, rng)
g = timeit(lambda: m0.batch_simulate_trials(P, 500, sim), 200)
steps = sim.last_stats()["total_steps"]
t0 = time.perf_counter()
for _ in range(5):
    orc.simulate_batch_mt(0, P, 500)
add("C2 basic_ddm_dc 64x500 dt=0.01", g, (time.perf_counter() - t0) / 5, 64 * 500, steps)

# C3: single_trial_alpha_not_scaled 1024 x 1000
P = priors.draw_prior_batch("alpha", 1024, rng)
g = timeit(lambda: m1.batch_simulate_trials(P, 1000, sim), 20)
steps = sim.last_stats()["total_steps"]
t0 = time.perf_counter()
orc.simulate_batch_mt(1, P, 1000)
cpu3 = time.perf_counter() - t0
add("C3 single_trial_alpha_not_scaled 1024x1000 dt=0.01", g, cpu3, 1024 * 1000, steps)
sim.set_host_decode(-1)   # the same call with float64 rows copied straight into the (pageable) result array
g = timeit(lambda: m1.batch_simulate_trials(P, 1000, sim), 20)
sim.set_host_decode(0)
add("C3 as above, float64 rows over PCIe (host decode off)", g, cpu3, 1024 * 1000, steps)
out3 = sim.pinned_empty((1024, 1000, 2), np.float64)
g = timeit(lambda: m1.batch_simulate_trials(P, 1000, sim, out=out3), 20)
add("C3 as above, compact wire into a pinned result array", g, cpu3, 1024 * 1000, steps)

# C4: Stahl-shaped imputation, 19 374 trials / 89 participants, device hand-off
subj, pe = stahl.synthetic_stahl_like()
pp = stahl.draw_participant_params(89, np.random.default_rng(2024))
g = timeit(lambda: stahl.impute_dataset(subj, pe, pp, simulator=sim, device=True), 50)
steps = sim.last_stats()["total_steps"]
_, alphas = stahl.boundaries_from_pe(pe)
_, idx = np.unique(subj, return_inverse=True)
t0 = time.perf_counter()
for i in range(0, 19374, 97):   # every 97th trial through the scalar oracle, scaled up
    orc.simulate_mt(5, pp[idx[i]], 1, seed=i, bound_in=[alphas[i]])
cpu = (time.perf_counter() - t0) * 97
add("C4 imputation_from_stahl 19374 trials (preprocess + simulate + DLPack)", g, cpu, 19374, steps)

json.dump(rows, open(os.path.join("gpurun_out", "config_timings.json"), "w"), indent=1)

# evidence-path variant (retired zoo): 256 datasets x 1000 trials, 200 observed samples, float32 rows left on the device
from bayesflow_nddms_b200 import basic_ddm_dc_evidence as mev  # noqa: E402

Pe = mev.batch_draw_prior(256)
for mode, name in ((1, "per-trial z-score"), (2, "dataset-level standardisation")):
    def go():
        b = sim.simulate_evidence(Pe, 1000, 200, mode, flags=2, device=True)
        del b
    g = timeit(go, 10)
    st = sim.last_stats()
    rows.append(dict(config=f"evidence 256x1000, n_obs=200, {name}, device-resident f32", gpu_ms=g * 1e3, trials=256000,
                     euler_steps=st["total_steps"], gpu_trials_per_s=256000 / g, out_gbs=256000 * 202 * 4 / g / 1e9,
                     kernel_ms=st["kernel_ms"]))
    print(json.dumps(rows[-1]), flush=True)
json.dump(rows, open(os.path.join("gpurun_out", "config_timings.json"), "w"), indent=1)
