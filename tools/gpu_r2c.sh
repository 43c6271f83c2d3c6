#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_synth.py -m gpu -q -x -k "interaction or stream or chunked" > gpurun_out/pytest_int.log 2>&1; echo "pytest rc=$? $(( $(date +%s) - T0 )) s"
tail -5 gpurun_out/pytest_int.log
timeout 300 python tools/interact_fwd_time.py 2>&1 | tail -8
SMALL="--row-cap 2000000 --lookahead 200 --e2e-steps -1 --no-cpu-baseline --no-kernel-prof --steps 100 --warmup 10"
for v in 0 3; do
  CDLRM_INTERACT_FWD=$v timeout 300 python bench.py $SMALL > gpurun_out/ab_ifwd_$v.json 2> gpurun_out/ab_ifwd_$v.err; echo "fwd=$v rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/ab_ifwd_$v.json').read().strip().splitlines()[-1])
print('interact fwd=$v ms/step', round(d['ms_per_step'],4))
PY
done
# whole-window leg at a smaller master (10 M row cap: 25 s of set-up): data generation in short CTAs
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-prof --row-cap 10000000 > gpurun_out/bench_w10m.json 2> gpurun_out/bench_w10m.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_w10m.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
fw=d['full_window']; print({k:fw[k] for k in fw if k!='ms_per_step_series'})
s=fw['ms_per_step_series']; print(s['first_40_steps_ms']); print(s['ms_per_step'][:40])
PY
