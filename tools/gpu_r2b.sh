#!/bin/bash
# A/B of the side-stream weight gradients, then the default bench line (boundary diagnosis).
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_trainer.py tests/test_gpu_run.py -m gpu -q -x > gpurun_out/pytest_mlp.log 2>&1; echo "pytest rc=$? $(( $(date +%s) - T0 )) s"
tail -5 gpurun_out/pytest_mlp.log
SMALL="--row-cap 2000000 --lookahead 200 --e2e-steps -1 --no-cpu-baseline --no-kernel-prof --steps 100 --warmup 10"
for v in "0 0" "1 0" "1 1"; do
  set -- $v
  CDLRM_WGRAD_SIDE=$1 CDLRM_DEFER_WGRAD=$2 timeout 300 python bench.py $SMALL > gpurun_out/ab_wgrad_$1$2.json 2> gpurun_out/ab_wgrad_$1$2.err; echo "side=$1 defer=$2 rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/ab_wgrad_$1$2.json').read().strip().splitlines()[-1])
print('side=$1 defer=$2 ms/step', round(d['ms_per_step'],4))
PY
done
T0=$(date +%s)
timeout 1200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$? $(( $(date +%s) - T0 )) s"
tail -6 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
fw=d['full_window']; print({k:fw[k] for k in fw if k!='ms_per_step_series'})
print(fw['ms_per_step_series'])
print({n:(k['us_per_launch'],k.get('frac_of_peak')) for n,k in (d['kernels'] or {}).items()})
PY
