"""Host-array call at mid sizes (1 Mi .. 64 Mi trials) into a pinned result array: one launch + one copy of float64
rows, against the streamed paths (float64 rows / compact wire) at several chunk sizes.  Decides the streaming defaults."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bayesflow_nddms_b200 import basic_ddm_dc, default_simulator

sim = default_simulator()
for D in (1024, 4096, 8192, 16384, 32768, 65536, 262144)[: int(os.environ.get("MIDSIZE_N", "7"))]:
    pe = bench.sweep_params(D, seed=1)
    out = sim.pinned_empty((D, 1000, 2), np.float64)
    row = {}
    cases = [("one launch + copy", -1, 1 << 60, -1), ("defaults", 0, -1, -1)]
    for chunk in (1 << 20, 2 << 20, 4 << 20, 8 << 20, 32 << 20):
        if chunk < D * 1000 or chunk == 1 << 20:
            cases += [(f"f64 {chunk >> 20}Mi", -1, 1, chunk), (f"compact {chunk >> 20}Mi", 0, 1, chunk), (f"compact4t {chunk >> 20}Mi", 4, 1, chunk)]
    for name, hd, min_rows, chunk in cases:
        sim.set_host_decode(hd)
        sim.set_pipeline(min_rows, chunk)
        ts = []
        for rep in range(7):
            t = time.perf_counter()
            basic_ddm_dc.batch_simulate_trials(pe, 1000, sim, dt=1e-3, max_steps=4000, dataset_offset=0, out=out)
            ts.append(time.perf_counter() - t)
        row[name] = round(float(np.median(ts[2:]) * 1e3), 2)
    print(D * 1000, row, flush=True)
