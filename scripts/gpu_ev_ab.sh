# evidence path: parity tests, then the probe (kernel times at 256 / 2048 / 16384 datasets x 1000 trials)
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_evidence.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/ev_probe.py 2>&1 | tee gpurun_out/ev_probe.log | tail -20
