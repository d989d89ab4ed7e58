# e2e leg of bench.py on N GPUs, plain float64 rows against the compact wire format (usage: gpu_multi_ab.sh N)
set -x
N=${1:-2}
mkdir -p gpurun_out
nproc; free -g | head -2
for hd in -1 0; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --host-decode $hd 2>gpurun_out/ab_${N}gpu_hd${hd}.err | tail -1 > gpurun_out/ab_${N}gpu_hd${hd}.json
  python -c "
import json,sys
d=json.load(open(sys.argv[1]))
e=d['e2e']
print('hd', sys.argv[2], 'value %.4e e2e %.4e ms %.0f threads %s d2h %.2e' % (d['value'], e['value'], e['ms_per_step'], e.get('host_decode_threads'), e['d2h_bytes_per_step']))" gpurun_out/ab_${N}gpu_hd${hd}.json $hd
  tail -2 gpurun_out/ab_${N}gpu_hd${hd}.err
done
