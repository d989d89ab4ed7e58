#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_call3_pytest.txt
tail -4 gpurun_out/r02_call3_pytest.txt
timeout 600 python scripts/r02_ab_kernels.py > gpurun_out/r02_ab_kernels_v4.jsonl 2> gpurun_out/r02_ab_kernels_v4.err
tail -3 gpurun_out/r02_ab_kernels_v4.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_ab_kernels_v4.jsonl"):
    d = json.loads(l)
    if "variant" in d:
        print(d["case"].ljust(22), d["variant"].ljust(7), "thr", str(d["thr"]).ljust(3), "tile", str(d["tile"]).ljust(4), "ms %8.4f" % d["kernel_ms"], "steps/s %.4g" % d["steps_per_s"])
    else:
        print(d["case"], "identical:", d["identical_output_checksums"])
PY
timeout 900 python bench.py --steps 3 --warmup 3 --datasets 100000 --cpu-seconds 2 --no-host-rows > gpurun_out/r02_bench_small.json 2> gpurun_out/r02_bench_small.err
tail -5 gpurun_out/r02_bench_small.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_small.json").read().strip().split("\n")[-1])
for k in ("value", "ms_per_step", "gpu_launches"):
    print(k, d[k])
print("roofline", {k: d["roofline"][k] for k in ("frac", "kernel_ms", "bare_loop_ceiling_steps_per_s", "frac_of_bare_loop_ceiling") if k in d["roofline"]})
print("e2e", {k: v for k, v in d["e2e"].items() if k != "api"})
print("rows", json.dumps(d.get("e2e_host_rows"), indent=0)[:1500])
print("configs", json.dumps(d.get("configs"), indent=0)[:4000])
print("training", d.get("training_batch"))
PY
