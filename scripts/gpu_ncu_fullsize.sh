# one ncu --set full capture of the persistent kernel at the bench's own launch size (1e9 trials) for roofline.traffic
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/ncu_fs_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:persistent_kernel -s 3 -c 1 -f -o gpurun_out/prof_persistent_fullsize $CMD > gpurun_out/ncu_fs.log 2>&1
tail -1 gpurun_out/ncu_fs_plain.log | cut -c1-200; tail -3 gpurun_out/ncu_fs.log
