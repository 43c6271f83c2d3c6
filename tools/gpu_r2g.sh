#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q -x --durations=3 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(( $(date +%s) - T0 )) s"
tail -6 gpurun_out/pytest_gpu.log
for mode in ce sm; do
T0=$(date +%s)
CDLRM_PREFETCH=$mode timeout 1200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-prof > gpurun_out/bench_n1_$mode.json 2> gpurun_out/bench_n1_$mode.err; echo "bench $mode rc=$? $(( $(date +%s) - T0 )) s"
tail -3 gpurun_out/bench_n1_$mode.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n1_$mode.json').read().strip().splitlines()[-1])
print('$mode: ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['pcie'])
fw=d['full_window']; print({k:fw[k] for k in fw if k!='ms_per_step_series'})
s=fw['ms_per_step_series']; print(s['first_40_steps_ms'][:24]); print(s['ms_per_step'][:40])
PY
done
sleep 3
timeout 60 python -c "import torch; x=torch.zeros(8,device='cuda:0'); torch.cuda.synchronize(); print('gpu alive')"
