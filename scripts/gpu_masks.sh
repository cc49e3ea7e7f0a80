#!/bin/bash
# New-feature check on the GPU box: masks tests first (fast feedback), then the whole gpu suite, then a quick bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_masks.py -x -q > gpurun_out/pytest_masks.log 2>&1; echo "masks rc=$?"
tail -25 gpurun_out/pytest_masks.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_masks.py --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
bash scripts/gpu_bench.sh
