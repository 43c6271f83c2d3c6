#!/bin/bash
# ncu evidence of round 2 at the DEFAULT bench configuration (terabyte shape, 40 M row cap, lookahead 3000; eager steps so
# that every kernel is its own launch): (1) launch list of two steps, (2) --set full of the cache path + interaction +
# small kernels, (3) --set full of the first GEMMs.  Read here with tools/ncu_summary.py.
mkdir -p gpurun_out
COMMON="--warmup 3 --no-graph --no-cpu-baseline --no-kernel-prof --e2e-steps -1"
T0=$(date +%s)
CDLRM_BENCH_CUPROF=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
  --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 $COMMON > gpurun_out/ncu_launch_bench.log 2>&1; echo "ncu launches rc=$? $(( $(date +%s) - T0 )) s"
T0=$(date +%s)
CDLRM_BENCH_CUPROF=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k 'regex:fwd_fused_kernel|fwd_miss_kernel|bwd_sgd_apply_kernel|bwd_plan|interact_fwd|interact_bwd|narrow_|bce_mean|split_' -c 14 \
  -o gpurun_out/r2_hot_full -f python bench.py --steps 1 $COMMON > gpurun_out/ncu_full_bench.log 2>&1; echo "ncu full (cache path) rc=$? $(( $(date +%s) - T0 )) s"
T0=$(date +%s)
CDLRM_BENCH_CUPROF=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k 'regex:gemm3x' -c 6 -o gpurun_out/r2_gemm_full -f python bench.py --steps 1 $COMMON > gpurun_out/ncu_gemm_bench.log 2>&1; echo "ncu full (gemm) rc=$? $(( $(date +%s) - T0 )) s"
ls -la gpurun_out | grep r2_
sleep 3
timeout 60 python -c "import torch; x=torch.zeros(8,device='cuda:0'); torch.cuda.synchronize(); print('gpu alive')"
