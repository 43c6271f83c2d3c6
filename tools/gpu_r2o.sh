#!/bin/bash
# 2-GPU visit: copy-engine prefetch (native chunk loops) against the zero-copy kernels at N=2 with the host cores of an 8-GPU box
# per rank (taskset: 8 cores for 2 ranks = the 4 cores per rank of a 32-core, 8-GPU box).
mkdir -p gpurun_out
echo "nproc $(nproc)"
CDLRM_PREFETCH=ce MGPU_MARKER=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    tools/mgpu_check.py > gpurun_out/mgpu_check_ce.log 2>&1; echo "mgpu_check ce rc=$?"
grep -E "mgpu_check OK|Error|error|assert" gpurun_out/mgpu_check_ce.log | head -4
for mode in ce sm; do
T0=$(date +%s)
CDLRM_PREFETCH=$mode taskset -c 0-7 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$([ $mode = ce ] && echo 1 || echo 2) \
  bench.py --gpus 2 --steps 20 --warmup 5 --no-kernel-prof > gpurun_out/r2o_bench_n2_$mode.json 2> gpurun_out/r2o_bench_n2_$mode.err; echo "bench N=2 $mode rc=$? $(( $(date +%s) - T0 )) s"
grep -v -i "warn" gpurun_out/r2o_bench_n2_$mode.err | grep "rank 0" | grep -E "trainer ready|window 0|timed region" 
python - <<PY
import json
d=json.loads(open('gpurun_out/r2o_bench_n2_$mode.json').read().strip().splitlines()[-1])
print('$mode N', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'hbm', d['config']['hbm_peak_allocated_gb'], 'first install', d['config']['first_window_install_ms'])
fw=d['full_window']; print({k:fw[k] for k in fw if k!='ms_per_step_series'})
s=fw['ms_per_step_series']; print(s['first_40_steps_ms'][:20]); print(s['ms_per_step'][:70])
print(d['pcie'])
PY
done
sleep 3
timeout 60 python -c "import torch; x=torch.zeros(8,device='cuda:0'); torch.cuda.synchronize(); print('gpu alive')"
