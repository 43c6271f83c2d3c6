#!/bin/bash
# interference of the PCIe prefetch with the training step: in-flight budget of the zero-copy gather
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mlp.py -m gpu -x -q 2>&1 | tail -2
COMMON="--steps 600 --warmup 20 --row-cap 10000000 --no-cpu-baseline --e2e-steps 0"
for v in "$@"; do
  IFS=: read ctas thr <<< "$v"
  CDLRM_PCIE_CTAS=$ctas CDLRM_PCIE_THREADS=$thr timeout 300 python bench.py $COMMON > gpurun_out/pcie_$ctas_$thr.json 2> gpurun_out/pcie.err
  python - <<PY
import json
r = json.load(open("gpurun_out/pcie_$ctas_$thr.json"))
s = r["ms_per_step_series"]["ms_per_step"]
base = sorted(s)[len(s)//4]
print("ctas $ctas thr $thr:", round(r["ms_per_step"], 4), "ms/step; base", base, "; extra ms over the region:", round(sum((x - base) * 25 for x in s), 1), s)
PY
done
