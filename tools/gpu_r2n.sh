#!/bin/bash
# 1-GPU visit: parity of the new kernels, interaction-forward A/B, bench with the fastest variant, ncu evidence
# (DRAM bytes per launch in ONE pass at the bench configuration -- no replay, so no save / restore of the 96 GB pinned master,
# which is what made `--set full` fail there -- and `--set full` at a 4 M row cap for the stall / source pages).
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_trainer.py tests/test_gpu_edge.py -m gpu -q -x \
  -k "interaction or terabyte or kaggle or trainer or aux or rebind" > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$? $(( $(date +%s) - T0 )) s"
tail -3 gpurun_out/pytest_new.log
timeout 300 python tools/interact_fwd_time.py > gpurun_out/interact_fwd_time.txt 2>&1; cat gpurun_out/interact_fwd_time.txt | tail -9
BEST=$(python - <<'PY'
import re
best={}
for l in open('gpurun_out/interact_fwd_time.txt'):
    m=re.match(r"fwd variant (\d) stagger \d+ ns: [\d.]+ us raw, ([\d.]+) us net", l)
    if m and m.group(1) in '356': best.setdefault(m.group(1), []).append(float(m.group(2)))
avg={k:sum(v)/len(v) for k,v in best.items()}
print(min(avg, key=avg.get) if avg else 3)
PY
)
echo "fastest forward variant: $BEST"
T0=$(date +%s)
CDLRM_INTERACT_FWD=$BEST timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/r2n_bench_n1.json 2> gpurun_out/r2n_bench_n1.err; echo "bench rc=$? $(( $(date +%s) - T0 )) s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2n_bench_n1.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
print({n:(k['us_per_launch'],k.get('frac_of_peak')) for n,k in (d['kernels'] or {}).items()})
PY
COMMON="--warmup 3 --no-graph --no-cpu-baseline --no-kernel-prof --e2e-steps -1"
T0=$(date +%s)
CDLRM_INTERACT_FWD=$BEST CDLRM_BENCH_CUPROF=1 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  --profile-from-start off -c 40 -o gpurun_out/r2_traffic_pass -f python bench.py --steps 1 $COMMON > gpurun_out/ncu_traffic_bench.log 2>&1; echo "ncu traffic rc=$? $(( $(date +%s) - T0 )) s"
tail -4 gpurun_out/ncu_traffic_bench.log
T0=$(date +%s)
CDLRM_INTERACT_FWD=$BEST CDLRM_BENCH_CUPROF=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k 'regex:fwd_fused_kernel|fwd_miss_kernel|bwd_sgd_apply_kernel|bwd_plan|interact_fwd|interact_bwd|gemm3x' -c 12 \
  -o gpurun_out/r2_hot_full -f python bench.py --steps 1 --row-cap 4000000 $COMMON > gpurun_out/ncu_full_bench.log 2>&1; echo "ncu full rc=$? $(( $(date +%s) - T0 )) s"
tail -4 gpurun_out/ncu_full_bench.log
ls -la gpurun_out | grep r2_
sleep 2
timeout 60 python -c "import torch; x=torch.zeros(8,device='cuda:0'); torch.cuda.synchronize(); print('gpu alive')"
