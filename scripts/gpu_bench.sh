#!/bin/bash
# quick perf iteration: bench only (no CPU baseline), prints the per-kernel-class table
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('ms/step', round(d['ms_per_step'],3), 'Gkmers/s', round(d['value']/1e9,2), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'launches', d['gpu_launches'])
for k,v in d['kernel_classes'].items(): print('  ', k, round(v['ms_per_step'],3), v['launches_per_step'], v['gbs'] and round(v['gbs']))
PY
