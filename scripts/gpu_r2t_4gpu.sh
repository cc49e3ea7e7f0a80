#!/bin/bash
mkdir -p gpurun_out
export KC_GROUP_TIMEOUT_MS=60000
for N in 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench$N rc=$?"
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench1 rc=$?"; tail -2 gpurun_out/bench_1gpu.err
python - <<'PY'
import json
for n in (1,2,4):
    f=f"gpurun_out/bench_{n}gpu.json"
    try:
        lines=[l for l in open(f).read().strip().splitlines() if l.startswith("{")]
        d=json.loads(lines[-1])
        print(f, d["n_gpus"], round(d["ms_per_step"],3), "ms", round(d["value"]/1e9,2), "G/s e2e", round(d["e2e"]["ms_per_step"],3), d.get("parity_n"))
        for k,v in (d.get("kernel_classes") or d.get("kernel_classes_rank0")).items(): print("   ",k, round(v["ms_per_step"],3), v["launches_per_step"])
        if n==1: print(json.dumps(d["roofline"])[:1800])
        else: print(json.dumps(d["exchange"]))
    except Exception as e: print(f, "ERR", e)
PY
