#!/bin/bash
mkdir -p gpurun_out
export KC_GROUP_TIMEOUT_MS=20000
timeout 900 python -m pytest tests/test_gpu_sig.py -m gpu -q --durations=5 > gpurun_out/pytest_sig.log 2>&1; echo "pytest sig rc=$?"
tail -30 gpurun_out/pytest_sig.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_sig.json 2> gpurun_out/bench_sig.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_sig.err; python -c "
import json; d=json.load(open('gpurun_out/bench_sig.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], json.dumps(d['kernel_classes']))"
