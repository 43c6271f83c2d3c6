#!/bin/bash
# 8-GPU bench of the default configuration (global window of 197 M ids per table) + per-rank HBM use.
# Run as:  gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_n8.sh [N]'
N=${1:-8}
mkdir -p gpurun_out
( while true; do nvidia-smi --query-gpu=index,memory.used --format=csv,noheader,nounits | tr '\n' ' '; echo; sleep 5; done ) > gpurun_out/hbm_n$N.txt &
MON=$!
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"
kill $MON
python - <<PY
import json
r = json.load(open("gpurun_out/bench_n$N.json"))
print(r["n_gpus"], "GPUs:", round(r["ms_per_step"], 4), "ms/step", round(r["value"] / 1e6, 2), "M samples/s; HBM peak allocated",
      r["config"].get("hbm_peak_allocated_gb"), "GB")
print(r["ms_per_step_series"])
print("max memory.used seen (MiB):", max(int(x) for line in open("gpurun_out/hbm_n$N.txt") for x in line.replace(",", " ").split()[1::2] or [0]))
PY
grep -v -i warn gpurun_out/bench_n$N.err | tail -6
