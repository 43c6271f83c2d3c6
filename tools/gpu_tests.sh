#!/bin/bash
# One GPU-box visit: the GPU parity suite only (no -x: report every failure).  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
T0=$(date +%s)
timeout ${1:-1500} python -m pytest tests -m gpu -q --durations=12 ${@:2} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(( $(date +%s) - T0 )) s"
tail -40 gpurun_out/pytest_gpu.log
