#!/bin/bash
# counter-layout change: parity first, then the kernel A/B (compare with profiles/r02_ab_kernels.jsonl)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python scripts/r02_ab_kernels.py > gpurun_out/r02_ab_kernels_layout.jsonl 2> gpurun_out/r02_ab_kernels_layout.err
tail -3 gpurun_out/r02_ab_kernels_layout.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_ab_kernels_layout.jsonl"):
    d = json.loads(l)
    if "kernel_ms" in d: print(d["case"], d["variant"], d["thr"], d["tile"], d["kernel_ms"], "%.4e" % d["steps_per_s"])
    else: print(d["case"], d["identical_output_checksums"])
PY
