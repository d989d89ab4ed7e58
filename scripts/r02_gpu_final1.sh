#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "check1_at_scale or trialwise_check1" 2>&1 | grep -E "crossing steps|distance of|passed|failed" > gpurun_out/r02_tie_classification.txt; cat gpurun_out/r02_tie_classification.txt
python -c "import __graft_entry__ as g; g.smoke()"
bash scripts/r02_gpu_bench_full.sh
