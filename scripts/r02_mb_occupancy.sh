#!/bin/bash
# the simulator block (one and two trials per lane) at 2..8 resident 256-thread blocks per SM
for b in 2 3 4 5 6 8; do
  DDM_MB_BLOCKS_PER_SM=$b python - <<PY
import sys
sys.path.insert(0, ".")
import bayesflow_nddms_b200 as pkg
from bayesflow_nddms_b200 import _capi
sim = pkg.DDMSimulator(device=0)
row = {}
for name in ("philox", "philox7", "sim_block", "sim_block_x2"):
    ips, hz = sim.microbench(_capi.MB_NAMES.index(name), 2048)
    row[name] = round(148 * 4 * hz / ips, 1)
print("blocks/SM $b warps/SMSP", $b * 2, row)
PY
done
