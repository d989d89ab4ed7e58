"""A/B of the host-array path: plain float64 rows over PCIe against the compact wire format with
1..N host decode threads (ddm_set_host_decode), on the bench workload.  Run on a GPU box:
    python scripts/e2e_ab.py [datasets] > gpurun_out/e2e_ab.json"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bayesflow_nddms_b200 import basic_ddm_dc, default_simulator  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
sim = default_simulator()
bound = sim.bind_host_thread_near_gpu()
pe = bench.sweep_params(D, seed=4000)
out = sim.pinned_empty((D, bench.N_TRIALS, 2), np.float64)
res = {"cpus_affinity": len(os.sched_getaffinity(0)), "cpus_total": os.cpu_count(), "numa_bound": bool(bound), "datasets": D,
       "runs": []}
ref = None
for threads, chunk in [(-1, -1), (0, -1), (4, -1), (8, -1), (12, -1), (16, -1), (24, -1), (32, -1), (0, 8 << 20), (0, 16 << 20),
                       (0, 64 << 20), (-1, -1)]:
    sim.set_host_decode(threads)
    sim.set_pipeline(-1, chunk)
    best = 1e9
    for rep in range(3):
        t = time.perf_counter()
        basic_ddm_dc.batch_simulate_trials(pe, bench.N_TRIALS, sim, dt=bench.DT, max_steps=bench.MAX_STEPS, dataset_offset=0, out=out)
        best = min(best, time.perf_counter() - t)
    steps = float(sim.last_stats()["total_steps"])
    chk = float(out[::997, ::7].sum())
    if ref is None:
        ref = chk
    res["runs"].append({"threads": threads, "chunk_rows": chunk, "s": best, "steps_per_s": steps / best,
                        "trials_per_s": D * bench.N_TRIALS / best, "same_checksum": chk == ref})
    print(res["runs"][-1], file=sys.stderr)
print(json.dumps(res))
