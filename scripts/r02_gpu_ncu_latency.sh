#!/bin/bash
# ncu --set full of the latency kernel and of the tile kernel on the same C2-sized launch (64 x 500 trials, dt = .01)
mkdir -p gpurun_out
python scripts/r02_c2_kernels.py -1 > gpurun_out/r02_c2_kernels_plain.txt 2>&1 || { cat gpurun_out/r02_c2_kernels_plain.txt; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"latency_kernel" -s 2 -c 1 -f -o gpurun_out/r02_prof_latency_c2 python scripts/r02_c2_kernels.py -1 > gpurun_out/r02_ncu_lat.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tile_kernel" -s 2 -c 1 -f -o gpurun_out/r02_prof_tile_c2 python scripts/r02_c2_kernels.py 0 > gpurun_out/r02_ncu_tile_c2.log 2>&1
ncu -i gpurun_out/r02_prof_latency_c2.ncu-rep --page details > gpurun_out/r02_ncu_latency_c2_details.txt 2>&1
ncu -i gpurun_out/r02_prof_tile_c2.ncu-rep --page details > gpurun_out/r02_ncu_tile_c2_details.txt 2>&1
grep -E "Duration|SM Frequency|Issue Slots Busy|No Eligible|Registers Per|Achieved Occupancy|Avg. Active Threads|Executed Ipc Active|Warp Cycles Per Issued|One or More Eligible|Executed Instructions  " gpurun_out/r02_ncu_latency_c2_details.txt gpurun_out/r02_ncu_tile_c2_details.txt
