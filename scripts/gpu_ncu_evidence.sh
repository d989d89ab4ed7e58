set -x
mkdir -p gpurun_out
CMD="python scripts/config_timings.py"
$CMD > gpurun_out/ev_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:evidence_warp_kernel -s 2 -c 1 -f -o gpurun_out/prof_evidence $CMD > gpurun_out/ncu_ev.log 2>&1
tail -4 gpurun_out/ev_plain.log | cut -c1-400; tail -3 gpurun_out/ncu_ev.log
