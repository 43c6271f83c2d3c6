#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(( $(date +%s) - T0 )) s"
tail -8 gpurun_out/pytest_gpu.log
SMALL="--row-cap 2000000 --lookahead 200 --e2e-steps -1 --no-cpu-baseline --steps 20 --warmup 5"
for c in 0 4 8; do
  CDLRM_PLAN_CLUSTER=$c timeout 300 python bench.py $SMALL > gpurun_out/ab_cluster$c.json 2> gpurun_out/ab_cluster$c.err; echo "cluster $c rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/ab_cluster$c.json').read().strip().splitlines()[-1])
k=d['kernels']
print('cluster $c ms/step', round(d['ms_per_step'],4), {n:(k[n]['us_per_launch'],k[n].get('frac_of_peak')) for n in ('embed_fwd','embed_miss','bwd_plan','bwd_sgd','interact_fwd','interact_bwd') if n in k})
PY
done
