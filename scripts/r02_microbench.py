"""Per-pipe issue rates on this B200 (ddm_microbench), round-2 set: python scripts/r02_microbench.py > out.json"""
import json
import sys

sys.path.insert(0, ".")
import bayesflow_nddms_b200 as pkg
from bayesflow_nddms_b200 import _capi

sim = pkg.DDMSimulator(device=0)
out = {}
for i, name in enumerate(_capi.MB_NAMES):
    ips, hz = sim.microbench(i, 2048)
    out[name] = {"warp_inst_per_s": ips, "sm_mhz": hz / 1e6, "cycles_per_inst_per_smsp": 148 * 4 * hz / ips}
print(json.dumps(out, indent=1))
sim.close()
