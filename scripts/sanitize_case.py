"""Small end-to-end case for compute-sanitizer: every kernel family once."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesflow_nddms_b200 as pkg  # noqa: E402
from bayesflow_nddms_b200 import priors  # noqa: E402

sim = pkg.DDMSimulator(0, seed=1)
rng = np.random.default_rng(0)
for model, name in [(0, "basic"), (1, "alpha"), (2, "alpha_dc"), (3, "alpha_scale"), (6, "eta")]:
    p = priors.draw_prior_batch(name, 9, rng)
    a = sim.simulate(model, p, 77, flags=4)
    b = sim.simulate(model, p, 77, flags=8, precision=32)
    c = sim.simulate(model, p, 77, precision=64)
    assert a.shape == b.shape == c.shape == (9, 77, 2)
# round 2: both schedulers, tiles that force recycling with trials still running (stragglers write their own rows),
# the general model, the histogram call, the generator histogram
from bayesflow_nddms_b200 import two_channel  # noqa: E402

for variant, thr, bps, tile in ((0, 2, 1, 8), (0, 0, 0, 0), (1, 0, 0, 0)):
    sim.set_kernel_variant(variant)
    sim.set_tuning(thr, bps, tile)
    for model, name in [(0, "basic"), (1, "alpha"), (2, "alpha_dc")]:
        sim.simulate(model, priors.draw_prior_batch(name, 23, rng), 301, dt=0.01, max_steps=400)
    raw = np.column_stack([rng.normal(0, 2, 7), 1.0 + rng.random(7), np.full(7, 0.5), np.full(7, 0.3), 0.5 + rng.random(7),
                           0.6 + rng.random(7), np.full(7, 0.3), np.full(7, 0.5), np.full(7, 0.6), np.full(7, 0.2), np.full(7, 0.1)])
    sim.simulate(7, two_channel.canonical_drift_dc5(raw), 130)
sim.set_kernel_variant(-1)
sim.set_tuning(0, 0, 0)
sim.simulate_histogram(0, priors.draw_prior_batch("sweep", 11, rng), 500, 1e-3, 4000, n_bins=64, rt_max=4.0)
sim.normals_histogram(600_000, 56, 5.6, 64)
sim.set_pipeline(1, 77 * 3)
sim.simulate(0, priors.draw_prior_batch("basic", 9, rng), 77)
sim.set_pipeline(-1, -1)
sim.simulate_trialwise(np.arange(50) % 3, rng.uniform(0, 2, 50), priors.draw_prior_batch("stahl", 3, rng))
ev = np.concatenate([priors.draw_prior_batch("basic", 3, rng), np.full((3, 1), 0.5)], axis=1)
for mode in (0, 1, 2):
    sim.simulate_evidence(ev, 45, 200, mode)
    sim.simulate_evidence(ev, 45, 200, mode, precision=64)
sim.draw_prior("alpha_scale", 100)
sim.export_normals(0, 0, 0, 0, 100)
d = sim.simulate_device(0, priors.draw_prior_batch("basic", 4, rng), 33)
import torch  # noqa: E402

t = torch.from_dlpack(d)
assert t.shape == (4, 33, 2)
del t
sim.close()
print("sanitize case ok")
