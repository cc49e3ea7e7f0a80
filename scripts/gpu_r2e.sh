#!/bin/bash
# TMA variants + small engine threads: correctness, then A/B timing on configs[1]
mkdir -p gpurun_out
export KC_GROUP_TIMEOUT_MS=10000
timeout 900 python -m pytest tests/test_gpu.py tests/test_sharded.py tests/test_gpu_big.py -m gpu -x -q -k "fast or group or big or compute" > gpurun_out/pytest_tma.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_tma.log
tail -6 gpurun_out/pytest_tma.log
for cfg in "1 0" "0 0" "1 256" "1 512"; do
  set -- $cfg
  KC_FAST_TMA=$1 KC_SMALL_THREADS=$2 timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tma$1_st$2.json 2> gpurun_out/bench_tma$1_st$2.err
  python - "$1" "$2" <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/bench_tma{sys.argv[1]}_st{sys.argv[2]}.json'))
print('tma',sys.argv[1],'small_threads',sys.argv[2],'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],3), {k: round(v['ms_per_step'],4) for k,v in d['kernel_classes'].items()})
PY
done
