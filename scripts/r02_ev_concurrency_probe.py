"""Is there anything to gain from running the evidence path's two kernels (record: issue-bound; post: latency-bound)
beside each other?  Two simulator contexts on one GPU, each given half of a batch from its own host thread, against one
context given the whole batch (wall clock, device-resident rows)."""
import os, sys, time, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesflow_nddms_b200 as pkg
from bayesflow_nddms_b200 import basic_ddm_dc_evidence as mev

B = int(os.environ.get("EV_B", "16384"))
sims = [pkg.DDMSimulator(0, seed=1) for _ in range(4)]
Pe = mev.batch_draw_prior(B)

def one(sim, P, off):
    b = sim.simulate_evidence(P, 1000, 200, 1, flags=2, device=True, dataset_offset=off); del b
    sim.synchronize()

def timed(nctx, reps=6):
    parts = np.array_split(np.arange(B), nctx)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        th = [threading.Thread(target=one, args=(sims[i], Pe[p], int(p[0]))) for i, p in enumerate(parts)]
        for t in th: t.start()
        for t in th: t.join()
        ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts[1:]), np.median(ts[1:])

for nctx in (1, 2, 4, 1, 2):
    best, med = timed(nctx)
    print(f"{nctx} context(s): best {best:.3f} ms  median {med:.3f} ms  -> {B * 1000 / best * 1e3:.3e} trials/s", flush=True)
