#!/bin/bash
for lib in libddm_b200.so libddm_b200_ab6.so; do
  for cfg in "sweep 0 5 128" "sweep 0 4 128" "sweep 1 5 64" "basic01 0 12 128" "basic01 1 16 64"; do
    echo -n "$lib: "; DDM_B200_LIB=$PWD/bayesflow_nddms_b200/$lib python scripts/r02_probe.py $cfg 4
  done
done
