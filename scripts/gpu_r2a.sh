#!/bin/bash
# round 2, first call: new tests (big configs vs reference, sparse switch), whole GPU suite, bench, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
timeout 900 python -m pytest tests/test_gpu_big.py tests/test_gpu_sparse.py -m gpu -x -q --durations=15 > gpurun_out/pytest_new.log 2>&1; echo "pytest new rc=$?" | tee -a gpurun_out/pytest_new.log
tail -25 gpurun_out/pytest_new.log
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_big.py --deselect tests/test_gpu_sparse.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
