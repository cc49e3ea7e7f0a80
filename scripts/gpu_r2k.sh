#!/bin/bash
mkdir -p gpurun_out
export KC_GROUP_TIMEOUT_MS=20000
timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_full.log
tail -14 gpurun_out/pytest_gpu_full.log
timeout 300 python profiles/path_stage_profile.py cfg4_human_310M | cut -c1-1300
timeout 300 python profiles/path_stage_profile.py cfg3_reads_10M | cut -c1-1300
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'])"
