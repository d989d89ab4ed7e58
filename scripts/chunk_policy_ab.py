"""Same-box A/B of the streamed path's chunk schedules (set_pipeline chunk_rows codes: -2 a quarter of the batch,
-3 half of what is left, -1 the smaller of the two; all within 2-32 Mi trials), compact wire and float64 rows."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bayesflow_nddms_b200 import basic_ddm_dc, default_simulator

sim = default_simulator()
for D in (8192, 16384, 32768, 65536, 262144, 1000000):
    pe = bench.sweep_params(D, seed=1)
    out = sim.pinned_empty((D, 1000, 2), np.float64)
    row = {}
    for rnd in range(2):
        for name, hd, chunk in (("quarter", 0, -2), ("half-left", 0, -3), ("min", 0, -1), ("f64 quarter", -1, -2), ("f64 half-left", -1, -3),
                                ("f64 min", -1, -1)):
            sim.set_host_decode(hd)
            sim.set_pipeline(1, chunk)
            ts = []
            for rep in range(6 if D < 1000000 else 4):
                t = time.perf_counter()
                basic_ddm_dc.batch_simulate_trials(pe, 1000, sim, dt=1e-3, max_steps=4000, dataset_offset=0, out=out)
                ts.append(time.perf_counter() - t)
            row.setdefault(name, []).append(round(float(np.median(ts[1:]) * 1e3), 2))
    print(D * 1000, row, flush=True)
