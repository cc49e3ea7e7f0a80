#!/bin/bash
# parity tests of the k-mer set paths, then the bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sig.py tests/test_gpu.py -m gpu -x -q > gpurun_out/pytest_sig.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_sig.log | cut -c1-200
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_y.json 2> gpurun_out/bench_y.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_y.json')); kc=d['kernel_classes']; print(round(d['ms_per_step'],4), round(d['ms_per_step_with_kernel_timers'],4), 'e2e', round(d['e2e']['ms_per_step'],3), {k: round(v['ms_per_step'],4) for k,v in kc.items()})"
