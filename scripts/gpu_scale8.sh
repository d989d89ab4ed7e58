set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | head -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --steps 5 --warmup 3 2>gpurun_out/bench_8gpu.err | tail -1 > gpurun_out/bench_8gpu.json
tail -c 400 gpurun_out/bench_8gpu.json; echo
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29539 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline --host-decode -1 2>gpurun_out/bench_8gpu_plain.err | tail -1 > gpurun_out/bench_8gpu_plain.json
python -c "
import json
for f in ('bench_8gpu','bench_8gpu_plain'):
    d=json.load(open('gpurun_out/%s.json'%f)); e=d['e2e']
    print(f,'value %.4e e2e %.4e ms %.0f threads %s' % (d['value'], e['value'], e['ms_per_step'], e.get('host_decode_threads')))"
