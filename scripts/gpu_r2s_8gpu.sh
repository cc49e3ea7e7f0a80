#!/bin/bash
mkdir -p gpurun_out
export KC_GROUP_TIMEOUT_MS=60000
for N in ${KC_NS:-8}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench$N rc=$?"
done
python - <<'PY'
import json
for n in (int(__import__("os").environ.get("KC_NS", "8")),):
    f=f"gpurun_out/bench_{n}gpu.json"
    try:
        lines=[l for l in open(f).read().strip().splitlines() if l.startswith("{")]
        d=json.loads(lines[-1])
        print(f, d["n_gpus"], round(d["ms_per_step"],3), "ms", round(d["value"]/1e9,2), "G/s e2e", round(d["e2e"]["ms_per_step"],3), d.get("parity_n"))
        for k,v in (d.get("kernel_classes") or d.get("kernel_classes_rank0")).items(): print("   ",k, round(v["ms_per_step"],3), v["launches_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
