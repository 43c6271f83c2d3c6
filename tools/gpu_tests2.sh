#!/bin/bash
# A/B of the trainer / pipeline tests with private (default) and pooled side streams
mkdir -p gpurun_out
for mode in 0 1; do
  CDLRM_POOL_STREAMS=$mode timeout 600 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_run.py -m gpu -q > gpurun_out/pytest_ab_$mode.log 2>&1
  echo "mode $mode rc=$?"; tail -12 gpurun_out/pytest_ab_$mode.log
done
