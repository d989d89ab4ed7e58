#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_evidence.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python scripts/ev_probe.py 2>&1 | tail -3
CMD="python scripts/ev_probe_small.py"
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_write.sum,dram__bytes_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct
for lib in libddm_b200.so; do
DDM_B200_LIB=bayesflow_nddms_b200/$lib timeout 300 ncu --metrics $M --clock-control none -k regex:"evidence_post_kernel|record_kernel" -c 2 --csv --log-file gpurun_out/r02_ev_launches_$lib.csv $CMD > gpurun_out/r02_ncu_ev_l.log 2>&1
echo $lib; tail -22 gpurun_out/r02_ev_launches_$lib.csv | cut -d, -f5,13- | sed 's/"Command line profiler metrics",//; s/"(256, 1, 1)","0","10.0",//'
done
