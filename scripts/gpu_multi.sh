set -x
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -12
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 2>gpurun_out/b2.err | tail -1 > gpurun_out/bench_2gpu.json
python -c "
import json
d=json.load(open('gpurun_out/bench_2gpu.json'))
print('value %.4e e2e %.4e ms %.0f bound %s' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('host_thread_bound_near_gpu')))"
tail -2 gpurun_out/b2.err
