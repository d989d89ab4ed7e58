# A/B of the tile kernel's refill forms (branchy = committed, branch-free selects, single loop), same box, same run
set -x
mkdir -p gpurun_out
OUT=gpurun_out/r02_refill_ab.txt
: > $OUT
for rep in 1 2; do
  for lib in build/lib_branchy.so build/lib_branchfree.so build/lib_single.so; do
    echo "== $lib rep $rep" >> $OUT
    DDM_B200_LIB=$PWD/$lib python scripts/tune.py 0,0,0 2>&1 | tail -2 >> $OUT
    DDM_B200_LIB=$PWD/$lib python scripts/short_trials.py 2>&1 | grep "thr= 0\|thr=12" >> $OUT
  done
done
for lib in build/lib_branchfree.so build/lib_single.so; do
  echo "== parity $lib" >> $OUT
  DDM_B200_LIB=$PWD/$lib timeout 600 python -m pytest tests -m gpu -x -q -k "bitwise or parity or tile" 2>&1 | tail -3 >> $OUT
done
cat $OUT
