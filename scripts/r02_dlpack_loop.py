"""Time the simulate -> DLPack -> torch loop of the all-gather leg piece by piece (one GPU)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch

import bayesflow_nddms_b200 as pkg
from bayesflow_nddms_b200 import priors
from bayesflow_nddms_b200 import single_trial_alpha_not_scaled as m1

sim = pkg.DDMSimulator(device=0, seed=2023)
P = priors.draw_prior_batch("alpha", 512, np.random.default_rng(77))


def t(label, fn, reps=50):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    print(f"{label:50s} {(time.perf_counter() - t0) / reps * 1e3:.4f} ms", flush=True)


keep = {}


def a():
    b = m1.batch_simulate_trials_device(P, 1000, sim, dataset_offset=5000)
    del b


def b_():
    keep["x"] = torch.from_dlpack(m1.batch_simulate_trials_device(P, 1000, sim, dataset_offset=5000))


def c():
    sim.run(1, P, 1000, 0.01, 400, flags=2, dataset_offset=5000)
    sim.synchronize()


t("run only (resident, sync)", c)
t("simulate_device + del capsule", a)
t("simulate_device + torch.from_dlpack (kept)", b_)
big = torch.from_dlpack(m1.batch_simulate_trials_device(priors.draw_prior_batch("alpha", 1024, np.random.default_rng(1)), 1000, sim, dataset_offset=0))
t("same, with an unrelated 8 MB batch alive", b_)
del big
t("same, after freeing it", b_)
print(sim.last_stats())
sim.close()
