# A/B: shipped (5 blocks x 256 threads, 48 regs) vs 4 x 256 (<= 64 regs) vs 3 x 384 (<= 56 regs)
OUT=gpurun_out/r02_occ_ab.txt
: > $OUT
for rep in 1 2; do
  for lib in bayesflow_nddms_b200/libddm_b200.so build/lib_mb4.so build/lib_b384.so; do
    echo "== $lib rep $rep" >> $OUT
    DDM_B200_LIB=$PWD/$lib python scripts/tune.py 0,0,0 2>&1 | tail -2 | head -1 >> $OUT
    DDM_B200_LIB=$PWD/$lib python scripts/short_trials.py 2>&1 | grep "thr= 0" >> $OUT
  done
done
cat $OUT
