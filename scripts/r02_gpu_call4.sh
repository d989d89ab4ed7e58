#!/bin/bash
mkdir -p gpurun_out
python scripts/r02_c2_latency.py 2>&1 | tee gpurun_out/r02_c2_latency.txt
timeout 600 python -m pytest tests/test_generator.py -m gpu -x -q -s 2>&1 | tail -12 | tee gpurun_out/r02_generator_test.txt
bash scripts/r02_mb_occupancy.sh 2>&1 | grep blocks | tee gpurun_out/r02_microbench_occupancy.txt
