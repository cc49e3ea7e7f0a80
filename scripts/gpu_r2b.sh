#!/bin/bash
# round 2, second call: multi-rank protocol on one GPU (virtual ranks, IPC processes), sparse switch, CLI, full suite, bench
mkdir -p gpurun_out
export KC_GROUP_TIMEOUT_MS=10000
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests/test_sharded.py -m gpu -x -q --durations=10 > gpurun_out/pytest_sharded.log 2>&1; echo "pytest sharded rc=$?" | tee -a gpurun_out/pytest_sharded.log
tail -40 gpurun_out/pytest_sharded.log
timeout 600 python -m pytest tests/test_gpu_sparse.py tests/test_cli.py -m gpu -q --durations=5 > gpurun_out/pytest_sparse_cli.log 2>&1; echo "pytest sparse+cli rc=$?" | tee -a gpurun_out/pytest_sparse_cli.log
tail -30 gpurun_out/pytest_sparse_cli.log
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_sharded.py --deselect tests/test_gpu_sparse.py --deselect tests/test_cli.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json | cut -c1-1500
