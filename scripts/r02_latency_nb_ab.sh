# blocks per iteration of the latency kernel: 2 (shipped) vs 3 vs 4
for lib in bayesflow_nddms_b200/libddm_b200.so build/lib_lat3.so build/lib_lat4.so; do
  echo "== $lib"
  DDM_B200_LIB=$PWD/$lib python scripts/r02_latency_probe.py 2>&1 | grep -v "4096 x\|1024 x" | sed 's/one-thread-per-trial.*latency kernel/latency kernel/' | cut -c1-150
done
