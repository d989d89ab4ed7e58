# ncu captures of the two kernels of the evidence path (stepping kernel in RECORD form, post kernel)
set -x
mkdir -p gpurun_out
CMD="python scripts/ev_probe_small.py"
$CMD > gpurun_out/ev_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"evidence_post_kernel|persistent_kernel" -s 4 -c 2 -f -o gpurun_out/prof_evidence $CMD > gpurun_out/ncu_ev.log 2>&1
tail -2 gpurun_out/ev_plain.log | cut -c1-400; tail -3 gpurun_out/ncu_ev.log
