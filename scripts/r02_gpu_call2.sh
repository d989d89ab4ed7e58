#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_call2_pytest.txt
tail -5 gpurun_out/r02_call2_pytest.txt
timeout 600 python scripts/r02_ab_kernels.py > gpurun_out/r02_ab_kernels_v3.jsonl 2> gpurun_out/r02_ab_kernels_v3.err
tail -3 gpurun_out/r02_ab_kernels_v3.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_ab_kernels_v3.jsonl"):
    d = json.loads(l)
    if "variant" in d:
        print(d["case"].ljust(22), d["variant"].ljust(7), "thr", str(d["thr"]).ljust(3), "tile", str(d["tile"]).ljust(4), "ms %8.4f" % d["kernel_ms"], "steps/s %.4g" % d["steps_per_s"])
    else:
        print(d["case"], "identical:", d["identical_output_checksums"])
PY
bash scripts/r02_mb_occupancy.sh 2>&1 | grep blocks
