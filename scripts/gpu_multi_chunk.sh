# e2e leg on N GPUs at several chunk sizes (is the staging traffic served from the last-level cache when chunks are small?)
set -x
N=${1:-2}
mkdir -p gpurun_out
for cr in -1 2097152 8388608; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$N bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --e2e-chunk-rows $cr 2>gpurun_out/mc.err | tail -1 > gpurun_out/mc_${N}_${cr}.json
  python -c "
import json,sys
d=json.load(open(sys.argv[1])); e=d['e2e']
print('chunk', sys.argv[2], 'value %.4e e2e %.4e ms %.0f threads %s' % (d['value'], e['value'], e['ms_per_step'], e.get('host_decode_threads')))" gpurun_out/mc_${N}_${cr}.json $cr
done
