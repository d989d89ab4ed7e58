#!/bin/bash
# first GPU call of round 2: the suite, then the kernel A/B
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_call1_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r02_call1_pytest.txt
tail -5 gpurun_out/r02_call1_pytest.txt
timeout 600 python scripts/r02_ab_kernels.py > gpurun_out/r02_ab_kernels.jsonl 2> gpurun_out/r02_ab_kernels.err
tail -3 gpurun_out/r02_ab_kernels.err
cat gpurun_out/r02_ab_kernels.jsonl | cut -c1-250
