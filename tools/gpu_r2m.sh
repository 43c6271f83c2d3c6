#!/bin/bash
# 1-GPU visit of round 2: GPU tests, the default bench line (native copy-engine pipeline), then the ncu evidence.
mkdir -p gpurun_out
echo "nproc $(nproc); cpu.max $(cat /sys/fs/cgroup/cpu.max 2>/dev/null); load $(cat /proc/loadavg)"
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(( $(date +%s) - T0 )) s"
tail -3 gpurun_out/pytest_gpu.log
T0=$(date +%s)
timeout 1200 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$? $(( $(date +%s) - T0 )) s"
grep -E "trainer ready|timed region" gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['pcie'], d.get('cpu_baseline'))
fw=d['full_window']; print({k:fw[k] for k in fw if k!='ms_per_step_series'})
s=fw['ms_per_step_series']; print(s['first_40_steps_ms']); print(s.get('first_40_host_iter_ms')); print(s['ms_per_step'][:50])
print({n:(k['us_per_launch'],k.get('frac_of_peak')) for n,k in (d['kernels'] or {}).items()})
PY
bash tools/gpu_profile_r2.sh
