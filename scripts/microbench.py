"""Per-pipe issue rates and pipe-mix experiments (ddm_microbench)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesflow_nddms_b200 as pkg  # noqa: E402

sim = pkg.DDMSimulator(0)
out = {}
for i, name in enumerate(pkg._capi.MB_NAMES):
    ips, hz = sim.microbench(i, 4096)
    per = ips / (148 * 4 * 1.965e9)
    out[name] = dict(warp_inst_per_s=ips, per_smsp_per_clk_at_1965=per, cycles_per_inst=1 / per if per else None)
    print(f"{name:16s} {ips:.4e} warp-inst/s  {per:.4f} /clk/SMSP  {1/per:8.2f} cycles each", flush=True)
json.dump(out, open("gpurun_out/microbench.json", "w"), indent=1)
