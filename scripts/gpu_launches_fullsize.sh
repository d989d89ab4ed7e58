# ncu launch list (durations only) of the bench command at its own size, e2e leg included
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$CMD > gpurun_out/lf_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fullsize.csv $CMD > gpurun_out/lf_ncu.log 2>&1
tail -1 gpurun_out/lf_plain.log | cut -c1-160; wc -l gpurun_out/launches_fullsize.csv
