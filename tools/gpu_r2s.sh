#!/bin/bash
# last 2-GPU visit of round 2: trainer parity tests, multi-rank check with the sharded fill prefetch, bench at N=2
mkdir -p gpurun_out
T0=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_trainer.py -m gpu -q > gpurun_out/pytest_trainer.log 2>&1; echo "pytest trainer rc=$? $(( $(date +%s) - T0 )) s"
tail -2 gpurun_out/pytest_trainer.log
MGPU_MARKER=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    tools/mgpu_check.py > gpurun_out/mgpu_check_fillshard.log 2>&1; echo "mgpu_check rc=$?"
grep -E "mgpu_check OK|Error|error|assert" gpurun_out/mgpu_check_fillshard.log | head -4
if ! grep -q "mgpu_check OK" gpurun_out/mgpu_check_fillshard.log; then
  tail -25 gpurun_out/mgpu_check_fillshard.log
  echo "sharded fills FAILED the multi-rank check: re-checking with CDLRM_FILL_SHARDED=0"
  CDLRM_FILL_SHARDED=0 MGPU_MARKER=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
    tools/mgpu_check.py > gpurun_out/mgpu_check_nofillshard.log 2>&1; echo "mgpu_check (fills per rank) rc=$?"
  grep -E "mgpu_check OK|Error|error|assert" gpurun_out/mgpu_check_nofillshard.log | head -4
  exit 0
fi
T0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus 2 --steps 20 --warmup 5 --no-kernel-prof > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench N=2 rc=$? $(( $(date +%s) - T0 )) s"
grep -v -i "warn" gpurun_out/r2_bench_n2.err | grep "rank 0" | grep -E "trainer ready|window 0|timed region|rror"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'hbm', d['config']['hbm_peak_allocated_gb'], 'first install', d['config']['first_window_install_ms'])
fw=d['full_window']; print({k:fw[k] for k in fw if k!='ms_per_step_series'})
s=fw['ms_per_step_series']; print(s['first_40_steps_ms'][:12]); print(s['ms_per_step'][:50])
print(d['pcie'])
PY
