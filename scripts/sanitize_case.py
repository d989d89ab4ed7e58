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
