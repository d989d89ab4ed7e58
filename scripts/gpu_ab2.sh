# A/B of an experimental build (AB_LIB) against the shipped library: parity tests on the experimental build, then
# the C5 tuning points and the dt = .01 probe on both, interleaved, same box
set -x
mkdir -p gpurun_out
DDM_B200_LIB=$PWD/$AB_LIB timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_distribution.py -m gpu -x -q 2>&1 | tail -4
for rep in 1 2; do
  for lib in bayesflow_nddms_b200/libddm_b200.so $AB_LIB; do
    echo "== $lib"
    DDM_B200_LIB=$PWD/$lib python scripts/tune.py 4,0,0 5,0,0 6,0,0 8,0,0 2>&1 | tail -5
    DDM_B200_LIB=$PWD/$lib python scripts/short_trials.py 2>&1 | tail -6
  done
done
