set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python -m pytest tests -m gpu -q 2>&1 | tail -15
python scripts/tune.py 2,0,64 3,0,64 4,0,64 5,0,64 6,0,64 8,0,64 4,0,16 4,0,32 4,0,128 4,4,64 4,5,64 2>&1 | tail -14 | tee gpurun_out/tune.log
python bench.py --steps 3 --warmup 3 --datasets 100000 --microbench --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_small.json
