# threads per block of the latency kernel: 128 (shipped) vs 64 vs 32
for lib in bayesflow_nddms_b200/libddm_b200.so build/lib_lt64.so build/lib_lt32.so; do
  echo "== $lib"
  DDM_B200_LIB=$PWD/$lib python scripts/r02_latency_probe.py 2>&1 | sed 's/one-thread-per-trial.*latency kernel/latency kernel/' | cut -c1-150
done
