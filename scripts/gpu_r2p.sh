#!/bin/bash
mkdir -p gpurun_out
export KC_GROUP_TIMEOUT_MS=20000
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_full.log
tail -25 gpurun_out/pytest_gpu_full.log | cut -c1-300
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], json.dumps(d['roofline'])[:1500])"
