#!/bin/bash
mkdir -p gpurun_out
export KC_GROUP_TIMEOUT_MS=20000
timeout 900 python -m pytest tests/test_gpu_sig.py -m gpu -q -x --durations=3 > gpurun_out/pytest_sig.log 2>&1; echo "pytest sig rc=$?"
tail -25 gpurun_out/pytest_sig.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_sig.json 2> gpurun_out/bench_sig.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_sig.err; python -c "
import json; d=json.load(open('gpurun_out/bench_sig.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], json.dumps(d['kernel_classes']))"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"kc_sig_scan|kc_sig_resolve" --launch-skip 2 -c 2 -o gpurun_out/r02_sig -f python profiles/step_for_ncu.py 1 1 > gpurun_out/ncu_sig.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py gpurun_out/r02_sig.ncu-rep gpurun_out/r02_sig_ncu.md "round 2: signature-bucket kernels on configs[1]" | grep -v "^$" | grep "duration\|DRAM\|SM thr\|occupancy %\|stall\|##\|conflict"
