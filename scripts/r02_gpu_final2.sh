#!/bin/bash
# final-library checks beside the main suite: bounds-checked build, 7-round build's distribution suite, generator test
mkdir -p gpurun_out
DDM_B200_LIB=$PWD/bayesflow_nddms_b200/libddm_b200_checked.so timeout 900 python scripts/r02_checked_run.py > gpurun_out/r02_checked_build_run.txt 2>&1; tail -4 gpurun_out/r02_checked_build_run.txt
DDM_B200_LIB=$PWD/bayesflow_nddms_b200/libddm_b200_checked.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_two_channel.py tests/test_evidence.py tests/test_gpu_distribution.py -m gpu -q 2>&1 | tail -2 >> gpurun_out/r02_checked_build_run.txt; tail -2 gpurun_out/r02_checked_build_run.txt
DDM_B200_LIB=$PWD/bayesflow_nddms_b200/libddm_b200_philox7.so timeout 900 python -m pytest tests/test_gpu_distribution.py tests/test_generator.py tests/test_gpu_exact.py -m gpu -s -q > gpurun_out/r02_philox7_distribution_suite.txt 2>&1; tail -3 gpurun_out/r02_philox7_distribution_suite.txt
timeout 600 python -m pytest tests/test_generator.py -m gpu -s -q > gpurun_out/r02_generator_test.txt 2>&1; tail -6 gpurun_out/r02_generator_test.txt
