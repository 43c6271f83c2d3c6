#!/bin/bash
# One GPU-box visit: parity tests, the default bench line, an ncu launch list of the timed region
# and --set full captures of the hot kernels.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(( $(date +%s) - T0 )) s"
tail -3 gpurun_out/pytest_gpu.log
T0=$(date +%s)
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$? $(( $(date +%s) - T0 )) s"
SMALL="--warmup 4 --lookahead 50 --row-cap 2000000 --no-graph --no-cpu-baseline --e2e-steps 0"
T0=$(date +%s)
CDLRM_BENCH_CUPROF=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 800 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 4 $SMALL > gpurun_out/ncu_launch_bench.log 2>&1; echo "ncu launches rc=$? $(( $(date +%s) - T0 )) s"
T0=$(date +%s)
CDLRM_BENCH_CUPROF=1 timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k 'regex:fwd_fused_kernel|fwd_miss_kernel|bwd_sgd_apply_kernel|bwd_plan_kernel|interact_fwd|interact_bwd|narrow_|bce_mean' -c 12 \
  -o gpurun_out/hot_full -f python bench.py --steps 1 $SMALL > gpurun_out/ncu_full_bench.log 2>&1; echo "ncu full (cache path) rc=$? $(( $(date +%s) - T0 )) s"
T0=$(date +%s)
CDLRM_BENCH_CUPROF=1 timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k 'regex:gemm3x' -c 7 -o gpurun_out/gemm_full -f python bench.py --steps 1 $SMALL > gpurun_out/ncu_gemm_bench.log 2>&1; echo "ncu full (gemm) rc=$? $(( $(date +%s) - T0 )) s"
ls -la gpurun_out
