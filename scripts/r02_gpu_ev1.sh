#!/bin/bash
# evidence path with the period-aligned recording kernel: tests, timing, launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_evidence.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python scripts/ev_probe.py 2>&1 | tail -5
CMD="python scripts/ev_probe_small.py"
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_write.sum,dram__bytes_read.sum --clock-control none -k regex:"evidence_post_kernel|record_kernel" -c 4 --csv --log-file gpurun_out/r02_ev_launches.csv $CMD > gpurun_out/r02_ncu_ev_l.log 2>&1
tail -28 gpurun_out/r02_ev_launches.csv | cut -d, -f5,13-
