#!/bin/bash
# N-GPU visit: multi-rank checks (own-id loser store, whole-window and sharded scan; the NVLink-sharded store too), then the bench at N.
N=${1:-2}
mkdir -p gpurun_out
chk() {
  env $1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 \
    tools/mgpu_check.py > gpurun_out/mgpu_check_$3.log 2>&1; echo "mgpu_check $3 rc=$?"
  grep -E "mgpu_check OK|Error|error|assert" gpurun_out/mgpu_check_$3.log | head -4
  grep -q "mgpu_check OK" gpurun_out/mgpu_check_$3.log || { tail -30 gpurun_out/mgpu_check_$3.log; echo "multi-rank check failed: no bench"; exit 1; }
}
chk "MGPU_MARKER=1" 29541 own_marker
chk "MGPU_MARKER=0" 29542 own_tensor
chk "MGPU_MARKER=1 CDLRM_LOSER_SHARDED=1" 29543 sharded
T0=$(date +%s)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$? $(( $(date +%s) - T0 )) s"
grep -v -i "warn" gpurun_out/bench_n$N.err | grep "rank 0" | tail -7
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,2), 'e2e', round(d['e2e']['ms_per_step'],4), round(d['e2e']['value']/1e6,2), 'hbm', d['config']['hbm_peak_allocated_gb'], 'first install', d['config']['first_window_install_ms'])
fw=d['full_window']; print({k:fw[k] for k in fw if k!='ms_per_step_series'})
s=fw['ms_per_step_series']; print(s['first_40_steps_ms'][:20]); print(s['ms_per_step'][:60])
print({n:(k['us_per_launch'],k.get('frac_of_peak'),k.get('misses_per_step')) for n,k in (d['kernels'] or {}).items()})
print(d['pcie'])
PY
sleep 3
timeout 60 python -c "import torch; x=torch.zeros(8,device='cuda:0'); torch.cuda.synchronize(); print('gpu alive')"
