#!/bin/bash
# 1-GPU visit: the whole GPU test suite, the default bench line, and one bench line per other BASELINE.json config
# (configs[0] small, configs[1] Kaggle shape, configs[3]: small 4-way cache and uniform ids at a 10 M row cap).
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(( $(date +%s) - T0 )) s"
tail -3 gpurun_out/pytest_gpu.log
T0=$(date +%s)
timeout 1200 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$? $(( $(date +%s) - T0 )) s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['pcie'], d.get('cpu_baseline',{}).get('value'))
fw=d['full_window']; print({k:fw[k] for k in fw if k!='ms_per_step_series'})
print(d['roofline']); print(d['step_roofline'])
print({n:(k['us_per_launch'],k.get('frac_of_peak'),k.get('dram_frac_of_peak')) for n,k in (d['kernels'] or {}).items()})
PY
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
          {n:(k[n]['us_per_launch'], k[n].get('misses_per_step')) for n in ('embed_fwd','embed_miss','bwd_sgd') if n in k}, d['pcie'].get('prefetch_rows'), d['pcie'].get('prefetch_GB/s'))
except Exception as e:
    print('$name failed', e); print(open('gpurun_out/r2_$name.err').read()[-1500:])
PY
}
run cfg0_small --workload small
run cfg1_kaggle --workload kaggle
run cfg3_s50k_w4 --row-cap 10000000 --cache-size 50000 --num-ways 4
run cfg3_uniform --row-cap 10000000 --dist uniform
sleep 2
timeout 60 python -c "import torch; x=torch.zeros(8,device='cuda:0'); torch.cuda.synchronize(); print('gpu alive')"
