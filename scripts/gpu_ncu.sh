# ncu evidence for the persistent kernel: launch list + one full capture (1 GPU, small case)
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --datasets 20000 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --metrics smsp__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed_pipe_fmalite.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_uniform.sum,smsp__inst_executed_pipe_cbu.sum,smsp__inst_executed_pipe_adu.sum,smsp__inst_executed_pipe_lsu.sum --clock-control none --import-source on -k regex:persistent_kernel -s 3 -c 1 -f -o gpurun_out/prof_persistent $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_plain.log | cut -c1-300; tail -5 gpurun_out/ncu_full.log | cut -c1-300; ls -la gpurun_out
