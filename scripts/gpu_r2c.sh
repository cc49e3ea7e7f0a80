#!/bin/bash
mkdir -p gpurun_out
export KC_GROUP_TIMEOUT_MS=10000
timeout 900 python -m pytest tests/test_sharded.py -m gpu -q --durations=10 > gpurun_out/pytest_sharded.log 2>&1; echo "pytest sharded rc=$?" | tee -a gpurun_out/pytest_sharded.log
tail -40 gpurun_out/pytest_sharded.log
timeout 600 python -m pytest tests/test_cli.py -m gpu -q -k several > gpurun_out/pytest_cli.log 2>&1; echo "pytest cli rc=$?" | tee -a gpurun_out/pytest_cli.log
tail -30 gpurun_out/pytest_cli.log
