#!/bin/bash
# bench.py on N GPUs of one box (usage: r02_gpu_scale_n.sh N), plus the multi-GPU pytest
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | head -2
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N --steps 5 --warmup 3 2>gpurun_out/r02_bench_${N}gpu.err | tail -1 > gpurun_out/r02_bench_${N}gpu.json
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_${N}gpu.json"))
print("value %.4e e2e %.4e (%.1f ms/step)" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["clocks"])
r = d.get("e2e_host_rows") or {}
for k in ("float64", "float32"):
    if k in r:
        print("rows", k, "%.4e steps/s %.1f ms %.1f GB/s frac_of_peak %.2f" % (r[k]["value"], r[k]["ms_per_step"], r[k]["host_row_write_gbs_all_ranks"], r[k].get("frac_of_host_store_peak", 0)))
print("host peak GB/s", r.get("host_stream_store_peak_gbs_all_ranks"))
print("allgather", d.get("allgather_batch"))
PY
tail -3 gpurun_out/r02_bench_${N}gpu.err
