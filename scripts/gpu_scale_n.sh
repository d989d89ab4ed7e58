# bench.py on N GPUs of one box (usage: gpu_scale_n.sh N)
set -x
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | head -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N --steps 5 --warmup 3 2>gpurun_out/bench_${N}gpu.err | tail -1 > gpurun_out/bench_${N}gpu.json
python -c "
import json,sys
d=json.load(open(sys.argv[1])); e=d['e2e']
print('value %.4e e2e %.4e ms %.0f threads %s' % (d['value'], e['value'], e['ms_per_step'], e.get('host_decode_threads')), d['clocks'])" gpurun_out/bench_${N}gpu.json
tail -2 gpurun_out/bench_${N}gpu.err
