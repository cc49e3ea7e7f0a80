#!/bin/bash
mkdir -p gpurun_out
export KC_GROUP_TIMEOUT_MS=20000
timeout 600 python -m pytest tests/test_sharded.py -m gpu -q -x > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_2gpu.log
tail -4 gpurun_out/pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"; tail -3 gpurun_out/bench_2gpu.err | cut -c1-300
python - <<'PY'
import json
for f in ("gpurun_out/bench_2gpu.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["ms_per_step"],3), "ms", round(d["value"]/1e9,2), "G/s e2e", round(d["e2e"]["ms_per_step"],3), d.get("parity_n"), d["config"])
        for k,v in (d.get("kernel_classes") or d.get("kernel_classes_rank0")).items(): print("   ",k, round(v["ms_per_step"],3), v["launches_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
