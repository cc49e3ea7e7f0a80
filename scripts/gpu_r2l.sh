#!/bin/bash
# signature-bucket construction: memcheck on small inputs, its tests, then the bench
mkdir -p gpurun_out
export KC_GROUP_TIMEOUT_MS=20000
timeout 900 compute-sanitizer --tool memcheck --target-processes all --log-file gpurun_out/sig_memcheck.log python -m pytest tests/test_gpu_sig.py -m gpu -q -x -k "edge_inputs or (matches_exact and (31-True or 63-False or 127-True))" > gpurun_out/sig_sanitizer_pytest.log 2>&1; echo "sanitizer rc=$?"
tail -5 gpurun_out/sig_sanitizer_pytest.log; grep "ERROR SUMMARY" gpurun_out/sig_memcheck.log | sort | uniq -c | head -5; grep -m3 -A12 "Invalid\|out of bounds" gpurun_out/sig_memcheck.log | head -50
timeout 900 python -m pytest tests/test_gpu_sig.py -m gpu -q --durations=5 > gpurun_out/pytest_sig.log 2>&1; echo "pytest sig rc=$?"
tail -40 gpurun_out/pytest_sig.log | cut -c1-400
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_sig.json 2> gpurun_out/bench_sig.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_sig.err; python -c "
import json; d=json.load(open('gpurun_out/bench_sig.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], json.dumps(d['kernel_classes']))"
