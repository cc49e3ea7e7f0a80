#!/bin/bash
# usage: gpu_variants.sh "ENV1=.. ENV2=.." "ENV.." ...   -- quick bench (no CPU baseline) once per environment setting
mkdir -p gpurun_out
i=0
for v in "$@"; do
  i=$((i+1))
  env $v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/variant_$i.json 2> gpurun_out/variant_$i.err; echo "variant $i [$v] rc=$?"
  python - "$v" gpurun_out/variant_$i.json <<'PY'
import json, sys
d=json.load(open(sys.argv[2]))
print(' ', sys.argv[1], '| ms/step', round(d['ms_per_step'],3), 'Gkmers/s', round(d['value']/1e9,2), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'launches', d['gpu_launches'])
print('   ', ' '.join(f"{k}={round(v['ms_per_step'],3)}" for k,v in d['kernel_classes'].items()))
PY
done
