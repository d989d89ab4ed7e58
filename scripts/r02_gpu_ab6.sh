#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/r02_ab_kernels.py > gpurun_out/r02_ab_kernels_v5.jsonl 2>gpurun_out/r02_ab_kernels_v5.err; tail -2 gpurun_out/r02_ab_kernels_v5.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_ab_kernels_v5.jsonl"):
    d = json.loads(l)
    if "variant" in d:
        print(d["case"].ljust(22), d["variant"].ljust(7), "thr", str(d["thr"]).ljust(3), "tile", str(d["tile"]).ljust(4), "ms %8.4f" % d["kernel_ms"], "steps/s %.4g" % d["steps_per_s"])
PY
for cfg in "sweep 0 3 128" "sweep 0 4 128" "sweep 0 5 128" "basic01 0 8 128" "basic01 0 10 128"; do
  echo -n "6 blocks: "; DDM_B200_LIB=$PWD/bayesflow_nddms_b200/libddm_b200_ab6.so python scripts/r02_probe.py $cfg 4
done
