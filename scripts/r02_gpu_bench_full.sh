#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_full_1gpu.json 2> gpurun_out/r02_bench_full_1gpu.err
tail -3 gpurun_out/r02_bench_full_1gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_full_1gpu.json").read().strip().split("\n")[-1])
print("value %.4e  ms/step %.2f  launches %d" % (d["value"], d["ms_per_step"], d["gpu_launches"]), d["clocks"])
print("roofline", {k: d["roofline"][k] for k in ("frac", "kernel_ms", "bare_loop_ceiling_steps_per_s", "frac_of_bare_loop_ceiling")})
print("e2e %.4e %.2f ms" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]))
r = d["e2e_host_rows"]
for k in ("float64", "float32"):
    print("rows", k, "%.4e %.1f ms, %.1f GB/s = %.2f of host peak" % (r[k]["value"], r[k]["ms_per_step"], r[k]["host_row_write_gbs_all_ranks"], r[k]["frac_of_host_store_peak"]))
print("host peak", r["host_stream_store_peak_gbs_all_ranks"])
print("training", d["training_batch"])
c = d["configs"]
for k, v in c.items():
    print(k, {kk: vv for kk, vv in v.items() if kk not in ("workload", "call", "device_call", "note", "cpu_port_note")} if isinstance(v, dict) else v)
print("cpu", d["cpu_baseline"])
PY
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; cut -c1-300 gpurun_out/r02_bench_reference_arm.json
