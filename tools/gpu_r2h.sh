#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "interaction" > gpurun_out/pytest_int.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_int.log
timeout 300 python tools/interact_fwd_time.py 2>&1 | tail -10
T0=$(date +%s)
timeout 1200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$? $(( $(date +%s) - T0 )) s"
grep -E "trainer ready|timed region" gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['pcie'])
fw=d['full_window']; print({k:fw[k] for k in fw if k!='ms_per_step_series'})
s=fw['ms_per_step_series']; print(s['first_40_steps_ms'][:24]); print(s['ms_per_step'][:40])
print({n:(k['us_per_launch'],k.get('frac_of_peak')) for n,k in (d['kernels'] or {}).items()})
PY
sleep 3
timeout 60 python -c "import torch; x=torch.zeros(8,device='cuda:0'); torch.cuda.synchronize(); print('gpu alive')"
