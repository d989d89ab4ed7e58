"""C3 (1024 x 1000 per-trial-boundary trials) through the reference-facing host call: one launch + one copy against the
streamed path (compact wire + host decode) at several chunk sizes and thread counts."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bayesflow_nddms_b200 as pkg
from bayesflow_nddms_b200 import priors
from bayesflow_nddms_b200 import single_trial_alpha_not_scaled as m1
from bayesflow_nddms_b200 import basic_ddm_dc as m0

sim = pkg.DDMSimulator(device=0, seed=2023)


def med(fn, reps=40):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


for name, mod, prior, B, N in (("C3 alpha 1024x1000", m1, "alpha", 1024, 1000), ("basic 1024x1000", m0, "basic", 1024, 1000),
                               ("alpha 4096x1000", m1, "alpha", 4096, 1000), ("alpha 256x1000", m1, "alpha", 256, 1000)):
    P = priors.draw_prior_batch(prior, B, np.random.default_rng(2023))
    for f32 in (0, 2):
        sim.set_pipeline(1 << 60, -1)
        sim.set_host_decode(0)
        base = med(lambda: mod.batch_simulate_trials(P, N, sim, flags=f32))
        row = [f"{name} {'f32' if f32 else 'f64'}: one launch + copy {base:.3f} ms |"]
        for chunk in (B * N // 2, B * N // 4, B * N // 8):
            for thr in (2, 4, 8):
                sim.set_pipeline(1, chunk)
                sim.set_host_decode(thr)
                t = med(lambda: mod.batch_simulate_trials(P, N, sim, flags=f32))
                row.append(f"c{B * N // chunk}/t{thr} {t:.3f}")
        sim.set_host_decode(-1)
        for chunk in (B * N // 2, B * N // 4):
            sim.set_pipeline(1, chunk)
            t = med(lambda: mod.batch_simulate_trials(P, N, sim, flags=f32))
            row.append(f"plain-c{B * N // chunk} {t:.3f}")
        sim.set_pipeline(-1, -1)
        sim.set_host_decode(0)
        row.append(f"| defaults {med(lambda: mod.batch_simulate_trials(P, N, sim, flags=f32)):.3f}")
        print(" ".join(row), flush=True)
sim.close()
