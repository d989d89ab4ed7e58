# A/B: the shipped library vs experimental builds under build/, same box, same run
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for rep in 1 2; do
  for lib in bayesflow_nddms_b200/libddm_b200.so $AB_LIBS; do
    echo "== $lib"
    DDM_B200_LIB=$PWD/$lib python scripts/tune.py 0,0,0 5,0,0 6,0,0 2>&1 | tail -4
  done
done
python scripts/short_trials.py 2>&1 | tail -10
