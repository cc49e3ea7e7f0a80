#!/bin/bash
# large-engine change: engine / sparse / group / CLI / big-config tests, then the path stage on the 310 Mbp repeat-model genome and the reads config
mkdir -p gpurun_out
export KC_GROUP_TIMEOUT_MS=20000
timeout 1500 python -m pytest tests/test_gpu.py tests/test_gpu_sparse.py tests/test_sharded.py tests/test_cli.py tests/test_gpu_big.py -m gpu -x -q > gpurun_out/pytest_engine.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_engine.log | cut -c1-300
timeout 600 python profiles/path_stage_profile.py cfg4_human_310M 2> gpurun_out/levels.err | cut -c1-1400
timeout 600 python profiles/path_stage_profile.py cfg3_reads_10M 2>> gpurun_out/levels.err | cut -c1-1400
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_e.json')); kc=d['kernel_classes']; print(round(d['ms_per_step'],4), round(d['ms_per_step_with_kernel_timers'],4), 'e2e', round(d['e2e']['ms_per_step'],3), {k: round(v['ms_per_step'],4) for k,v in kc.items()})"
