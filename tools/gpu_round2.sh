#!/bin/bash
# One GPU-box visit: the GPU parity suite, then the default bench line (both arms).  Outputs under gpurun_out/.
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(( $(date +%s) - T0 )) s"
tail -15 gpurun_out/pytest_gpu.log
T0=$(date +%s)
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "reference arm rc=$? $(( $(date +%s) - T0 )) s"
tail -c 1500 gpurun_out/bench_ref.json
T0=$(date +%s)
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$? $(( $(date +%s) - T0 )) s"
tail -20 gpurun_out/bench_n1.err
