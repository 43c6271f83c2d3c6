#!/bin/bash
# GPU visit: parity suite, then the default bench line.  Outputs under gpurun_out/.
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(( $(date +%s) - T0 )) s"
tail -8 gpurun_out/pytest_gpu.log
T0=$(date +%s)
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$? $(( $(date +%s) - T0 )) s"
tail -12 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'roofline', d['roofline'])
fw=d['full_window']; print({k:fw[k] for k in fw if k!='ms_per_step_series'})
print(fw['ms_per_step_series']['ms_per_step'])
print({n:(k['us_per_launch'],k.get('frac_of_peak')) for n,k in (d['kernels'] or {}).items()})
PY
