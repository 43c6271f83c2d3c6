#!/bin/bash
# 2-GPU visit: write-back shared by the ranks (multi-rank check + bench at N=2, default modes)
mkdir -p gpurun_out
MGPU_MARKER=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    tools/mgpu_check.py > gpurun_out/mgpu_check_wbshared.log 2>&1; echo "mgpu_check rc=$?"
grep -E "mgpu_check OK|Error|error|assert" gpurun_out/mgpu_check_wbshared.log | head -4
grep -q "mgpu_check OK" gpurun_out/mgpu_check_wbshared.log || { tail -30 gpurun_out/mgpu_check_wbshared.log; echo "multi-rank check failed: no bench"; exit 1; }
T0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench N=2 rc=$? $(( $(date +%s) - T0 )) s"
grep -v -i "warn" gpurun_out/r2_bench_n2.err | grep "rank 0" | grep -E "trainer ready|window 0|timed region"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'hbm', d['config']['hbm_peak_allocated_gb'], 'first install', d['config']['first_window_install_ms'])
fw=d['full_window']; print({k:fw[k] for k in fw if k!='ms_per_step_series'})
s=fw['ms_per_step_series']; print(s['first_40_steps_ms'][:20]); print(s['ms_per_step'][:70])
print(d['pcie'])
print({n:(k['us_per_launch'],k.get('frac_of_peak'),k.get('misses_per_step')) for n,k in (d['kernels'] or {}).items()})
PY
sleep 3
timeout 60 python -c "import torch; x=torch.zeros(8,device='cuda:0'); torch.cuda.synchronize(); print('gpu alive')"
