"""A/B of ddm_simulate_histogram at the bench's size (1e6 datasets x 1000 trials, host parameters in, histogram out):
one upload + one launch + one reduction against the chunked schedule (ddm_histogram_chunks)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesflow_nddms_b200 as pkg
from bayesflow_nddms_b200 import basic_ddm_dc, priors

D = int(os.environ.get("AB_DATASETS", "1000000"))
sim = pkg.DDMSimulator(0, seed=2023)
P = priors.draw_prior_batch("sweep", D, np.random.default_rng(2023))
ref = None
for name, args in (("one launch", (1 << 62, -1)), ("chunked (default)", (-1, -1)), ("chunked, min chunk 16 Mi", (-1, 16 << 20)),
                   ("chunked, min chunk 1 Mi", (-1, 1 << 20)), ("one launch", (1 << 62, -1)), ("chunked (default)", (-1, -1))):
    sim.set_pipeline(*args)
    ts = []
    for rep in range(4):
        t0 = time.perf_counter()
        h = basic_ddm_dc.batch_simulate_histogram(P, 1000, sim, dt=1e-3, max_steps=4000, seed=7, dataset_offset=0, n_bins=401, rt_max=4.01)
        ts.append((time.perf_counter() - t0) * 1e3)
    st = sim.last_stats()
    key = (h["upper"].tobytes(), h["lower"].tobytes(), h["missing"], h["overflow"], st["total_steps"])
    ref = key if ref is None else ref
    print(f"{name:28s} call ms {['%.2f' % t for t in ts]}  best {min(ts[1:]):.2f}  event window {st['kernel_ms']:.2f}  launches {st['kernel_launches']}"
          f"  steps/s (best call) {st['total_steps'] / min(ts[1:]) * 1e3:.4e}  same result {key == ref}", flush=True)
