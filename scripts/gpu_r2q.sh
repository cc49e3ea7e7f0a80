#!/bin/bash
mkdir -p gpurun_out
export KC_GROUP_TIMEOUT_MS=20000
timeout 1200 python -m pytest tests/test_sharded.py -m gpu -x -q --durations=5 > gpurun_out/pytest_sharded.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_sharded.log
tail -30 gpurun_out/pytest_sharded.log | cut -c1-400
