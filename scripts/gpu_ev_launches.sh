# evidence path: per-kernel durations and instruction counts (ncu launch list)
set -x
mkdir -p gpurun_out
CMD="python scripts/ev_probe_small.py"
$CMD > gpurun_out/ev_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"evidence_post_kernel|persistent_kernel|prep" -c 12 --csv --log-file gpurun_out/ev_launches.csv $CMD > gpurun_out/ncu_ev_l.log 2>&1
tail -40 gpurun_out/ev_launches.csv | cut -d, -f5,12-
