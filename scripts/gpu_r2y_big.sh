#!/bin/bash
# signature path on all word widths: parity tests, then the full-size configs (sig vs fixed-slot superstrings must be identical)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sig.py -m gpu -x -q > gpurun_out/pytest_sig.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_sig.log | cut -c1-200
timeout 900 python profiles/sig_big_check.py cfg2_k63u cfg2_k127u cfg4_human_310M 2> gpurun_out/sig_big.err | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); s = r['signature_buckets']; f = r['fixed_slots_or_exact']
    print(r['config'], '| sig', s['stage_ms'], s['set_kernels_ms'], '| other count', f['stage_ms']['count'], '| identical', r['identical_superstrings'])
"
