#!/bin/bash
# ncu capture of the signature kernels after the branch-free resolve
mkdir -p gpurun_out
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_z.json 2> gpurun_out/bench_z.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_z.json')); kc=d['kernel_classes']; print(round(d['ms_per_step'],4), round(d['ms_per_step_with_kernel_timers'],4), 'e2e', round(d['e2e']['ms_per_step'],3), {k: round(v['ms_per_step'],4) for k,v in kc.items()})"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"kc_sig_scan|kc_sig_resolve" --launch-skip 2 -c 2 -o gpurun_out/r02z_full -f python profiles/step_for_ncu.py 1 1 > gpurun_out/ncu_full_z.log 2>&1; echo "ncu full rc=$?"
