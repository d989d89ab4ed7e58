set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -20
python -m pytest tests -m gpu -x -q 2>&1 | tail -40
python bench.py --steps 2 --warmup 3 --datasets 100000 --microbench 2>&1 | tail -5 | tee gpurun_out/bench_small.json
