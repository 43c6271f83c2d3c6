#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(( $(date +%s) - T0 )) s"
tail -3 gpurun_out/pytest_gpu.log
for th in 4 8; do
T0=$(date +%s)
CDLRM_HOST_THREADS=$th CDLRM_PREFETCH=ce timeout 1200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-prof > gpurun_out/bench_n1_ce$th.json 2> gpurun_out/bench_n1_ce$th.err; echo "bench ce threads=$th rc=$? $(( $(date +%s) - T0 )) s"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n1_ce$th.json').read().strip().splitlines()[-1])
print('threads $th: ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['pcie']['prefetch_GB/s'])
fw=d['full_window']; print(fw['planner_timeline_ms'], fw['boundary_device_ms'])
s=fw['ms_per_step_series']; print(s['first_40_steps_ms'][:30]); print(s['ms_per_step'][:60])
PY
done
sleep 3
timeout 60 python -c "import torch; x=torch.zeros(8,device='cuda:0'); torch.cuda.synchronize(); print('gpu alive')"
