#!/bin/bash
# 8 x B200: the north-star run (3.1 Gbp repeat-model genome through the C++ CLI on 8 GPUs, verified) and the 4 / 8 GPU bench lines
mkdir -p gpurun_out
export KC_GROUP_TIMEOUT_MS=60000
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/ns_smi.txt 2>&1
nproc > gpurun_out/ns_host.txt; free -g >> gpurun_out/ns_host.txt
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "bench8 rc=$?" ) &
BENCH=$!
python scripts/northstar_input.py /dev/shm/h.fa > gpurun_out/ns_input.log 2>&1; cat gpurun_out/ns_input.log
wait $BENCH
tail -c 1200 gpurun_out/bench_8gpu.json; tail -3 gpurun_out/bench_8gpu.err
ls -la /dev/shm/h.fa
T0=$(date +%s%N)
timeout 600 host/kmercamel compute -k 31 -g 0-7 -o /dev/shm/out8.msfa /dev/shm/h.fa 2> gpurun_out/ns_cli_8gpu.log; echo "cli8 rc=$? wall $(( ($(date +%s%N) - T0) / 1000000 )) ms" | tee -a gpurun_out/ns_cli_8gpu.log
cat gpurun_out/ns_cli_8gpu.log
T0=$(date +%s%N)
timeout 900 host/kmercamel compute -k 31 -g 0-7 -V -o /dev/shm/out8v.msfa /dev/shm/h.fa 2> gpurun_out/ns_cli_8gpu_verify.log; echo "cli8 -V rc=$? wall $(( ($(date +%s%N) - T0) / 1000000 )) ms" | tee -a gpurun_out/ns_cli_8gpu_verify.log
tail -6 gpurun_out/ns_cli_8gpu_verify.log
T0=$(date +%s%N)
timeout 600 host/kmercamel compute -k 31 -g 0 -o /dev/shm/out1.msfa /dev/shm/h.fa 2> gpurun_out/ns_cli_1gpu.log; echo "cli1 rc=$? wall $(( ($(date +%s%N) - T0) / 1000000 )) ms" | tee -a gpurun_out/ns_cli_1gpu.log
tail -5 gpurun_out/ns_cli_1gpu.log
md5sum /dev/shm/out8.msfa /dev/shm/out8v.msfa /dev/shm/out1.msfa | tee gpurun_out/ns_md5.txt
ls -la /dev/shm/*.msfa >> gpurun_out/ns_md5.txt
rm -f /dev/shm/out8v.msfa /dev/shm/out1.msfa
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err; echo "bench4 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_4gpu.json","gpurun_out/bench_8gpu.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["ms_per_step"],3), "ms", round(d["value"]/1e9,2), "G/s e2e", round(d["e2e"]["ms_per_step"],3), d.get("parity_n"), d["exchange"]["nvlink_gbs_during_level0_rank0"])
        for k,v in (d.get("kernel_classes") or d.get("kernel_classes_rank0")).items(): print("   ",k, round(v["ms_per_step"],3), v["launches_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
rm -f /dev/shm/h.fa /dev/shm/out8.msfa
