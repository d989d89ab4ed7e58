#!/bin/bash
# Round 2 profiling: (1) plain run of the bench command, (2) its ncu launch list, (3) ncu --set full of the stepping kernel
# at the bench's own launch size (1e9 trials) and (4) of the per-trial-boundary kernel at the reference's dt = .01.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-host-rows --no-configs"
$CMD > gpurun_out/r02_lf_plain.json 2> gpurun_out/r02_lf_plain.err || { tail -5 gpurun_out/r02_lf_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_fullsize.csv $CMD > gpurun_out/r02_lf_ncu.log 2>&1
wc -l gpurun_out/r02_launches_fullsize.csv
CMD2="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
ncu --set full --clock-control none --import-source on -k regex:"tile_kernel|persistent_kernel" -s 3 -c 1 -f -o gpurun_out/r02_prof_sweep_fullsize $CMD2 > gpurun_out/r02_ncu_fs.log 2>&1
tail -2 gpurun_out/r02_ncu_fs.log
ncu -i gpurun_out/r02_prof_sweep_fullsize.ncu-rep --page details > gpurun_out/r02_ncu_sweep_fullsize_details.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tile_kernel|persistent_kernel" -s 1 -c 1 -f -o gpurun_out/r02_prof_c3 python scripts/r02_probe.py c3 ${1:-0} 0 0 3 > gpurun_out/r02_ncu_c3.log 2>&1
tail -2 gpurun_out/r02_ncu_c3.log
ncu -i gpurun_out/r02_prof_c3.ncu-rep --page details > gpurun_out/r02_ncu_c3_details.txt 2>&1
grep -E "Duration|Issue Slots Busy|No Eligible|Registers Per|Achieved Occupancy|Avg. Active Threads|dram__bytes|DRAM Throughput|Executed Ipc" gpurun_out/r02_ncu_sweep_fullsize_details.txt gpurun_out/r02_ncu_c3_details.txt | head -40
ncu -i gpurun_out/r02_prof_sweep_fullsize.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum 2>/dev/null | tail -3
