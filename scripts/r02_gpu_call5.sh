#!/bin/bash
mkdir -p gpurun_out
for extra in "--no-configs" "" ; do
python bench.py --steps 2 --warmup 3 --datasets 50000 --no-cpu-baseline --no-host-rows $extra 2>gpurun_out/r02_b5.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('training_batch', d['training_batch'].get('ms_per_batch'), 'configs' , list((d.get('configs') or {}).keys()))
"
tail -2 gpurun_out/r02_b5.err
done
