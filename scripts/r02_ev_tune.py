"""Evidence path (record_kernel + evidence_post_kernel): kernel time over refill thresholds and claim tiles.
kernel_ms covers both kernels; the post kernel's share does not depend on the knobs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesflow_nddms_b200 as pkg
from bayesflow_nddms_b200 import basic_ddm_dc_evidence as mev
sim = pkg.DDMSimulator(0, seed=1)
print("lib", os.environ.get("DDM_B200_LIB", "default"))
for B in (2048, 16384):
    Pe = mev.batch_draw_prior(B)
    for thr, tile in ((0, 0), (1, 64), (2, 64), (3, 64), (4, 64), (6, 64), (8, 64), (12, 64), (3, 32), (3, 128), (6, 128), (3, 256)):
        sim.set_tuning(thr, 0, tile)
        best = 1e9
        for _ in range(4):
            b = sim.simulate_evidence(Pe, 1000, 200, 1, flags=2, device=True); del b
            st = sim.last_stats()
            best = min(best, st["kernel_ms"])
        print(B, "thr", thr, "tile", tile, "kernel_ms %.4f" % best, "trials/s %.4e" % (B * 1000 / best * 1e3), "grid", st["grid"], flush=True)
