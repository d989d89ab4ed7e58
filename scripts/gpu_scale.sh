set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
free -g | head -2
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 5 --warmup 3 2>gpurun_out/bench_${n}gpu.err | tail -1 > gpurun_out/bench_${n}gpu.json
  tail -c 600 gpurun_out/bench_${n}gpu.json; echo
  tail -3 gpurun_out/bench_${n}gpu.err
done
