#!/bin/bash
# One bench line per BASELINE.json config that is not the headline (configs[0], [1], a 3-point [3] sweep + uniform ids),
# 1 GPU.  The [3] points use a 10 M row cap (shorter set-up; the cache geometry is what is swept).
mkdir -p gpurun_out
run() {
  name=$1; shift
  T0=$(date +%s)
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r2_$name.json 2> gpurun_out/r2_$name.err; echo "$name rc=$? $(( $(date +%s) - T0 )) s"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_$name.json').read().strip().splitlines()[-1])
    k=d.get('kernels') or {}
    print('$name', 'ms/step', round(d['ms_per_step'],4), 'samples/s', round(d['value']), 'e2e', round(d['e2e']['ms_per_step'],4) if d.get('e2e') else None,
          {n:(k[n]['us_per_launch'], k[n].get('misses_per_step')) for n in ('embed_fwd','embed_miss','bwd_sgd') if n in k})
except Exception as e:
    print('$name failed', e); print(open('gpurun_out/r2_$name.err').read()[-1500:])
PY
}
run cfg0_small --workload small
run cfg1_kaggle --workload kaggle
run cfg3_s50k_w4 --row-cap 10000000 --cache-size 50000 --num-ways 4
run cfg3_s300k_w8 --row-cap 10000000 --cache-size 300000 --num-ways 8
run cfg3_s600k_w16 --row-cap 10000000 --cache-size 600000 --num-ways 16
run cfg3_uniform --row-cap 10000000 --dist uniform
sleep 3
timeout 60 python -c "import torch; x=torch.zeros(8,device='cuda:0'); torch.cuda.synchronize(); print('gpu alive')"
