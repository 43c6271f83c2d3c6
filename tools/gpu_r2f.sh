#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q -x --durations=3 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(( $(date +%s) - T0 )) s"
tail -6 gpurun_out/pytest_gpu.log
SMALL="--row-cap 2000000 --lookahead 200 --e2e-steps -1 --no-cpu-baseline --no-kernel-prof --steps 100 --warmup 10"
for v in 0 1; do
  CDLRM_SGD_SPLIT=$v timeout 300 python bench.py $SMALL > gpurun_out/ab_sgdsplit_$v.json 2> gpurun_out/ab_sgdsplit_$v.err; echo "sgd_split=$v rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/ab_sgdsplit_$v.json').read().strip().splitlines()[-1])
print('sgd_split=$v ms/step', round(d['ms_per_step'],4))
PY
done
# copy-engine interference probe: 8 GB of cudaMemcpyAsync H2D beside the steps (compare with the SM-driven prefetch)
timeout 300 python bench.py --row-cap 2000000 --lookahead 2000 --e2e-steps -1 --no-cpu-baseline --no-kernel-prof --steps 1200 --warmup 10 --ce-probe-gb 8 > gpurun_out/ce_probe.json 2> gpurun_out/ce_probe.err; echo "ce probe rc=$?"
grep "copy-engine probe" gpurun_out/ce_probe.err
T0=$(date +%s)
timeout 1200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$? $(( $(date +%s) - T0 )) s"
tail -4 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
fw=d['full_window']; print({k:fw[k] for k in fw if k!='ms_per_step_series'})
s=fw['ms_per_step_series']; print(s['first_40_steps_ms'][:24]); print(s['ms_per_step'][:40])
print({n:(k['us_per_launch'],k.get('frac_of_peak')) for n,k in (d['kernels'] or {}).items()})
PY
sleep 3
timeout 60 python -c "import torch; x=torch.zeros(8,device='cuda:0'); torch.cuda.synchronize(); print('gpu alive')"
