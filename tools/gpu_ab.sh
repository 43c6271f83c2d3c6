#!/bin/bash
# A/B of launch plumbing options on one GPU (short windows, capped master tables to keep set-up short)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
COMMON="--steps 600 --warmup 20 --row-cap 10000000 --no-cpu-baseline --e2e-steps 0"
for v in "${@:-base}"; do
  case $v in
    base) env= ;;
    pdl0) env="CDLRM_PDL=0" ;;
    flat0) env="CDLRM_FLAT_MLP=0" ;;
    torchmlp) env="CDLRM_MLP=torch" ;;
  esac
  env $env timeout 300 python bench.py $COMMON > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err; echo "$v rc=$?"
done
python - "$@" <<'PY'
import json, sys
for n in (sys.argv[1:] or ["base"]):
    try:
        r = json.load(open(f"gpurun_out/ab_{n}.json"))
        print(n, round(r["ms_per_step"], 4), "ms/step", r["ms_per_step_series"]["ms_per_step"])
        print("   ", {k: v["us_per_launch"] for k, v in r["kernels"].items()})
    except Exception as e:
        print(n, "failed", e)
PY
