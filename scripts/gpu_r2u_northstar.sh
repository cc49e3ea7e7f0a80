#!/bin/bash
# 8 x B200: the north-star run with the signature-bucket construction (3.1 Gbp repeat-model genome through the C++ CLI, verified)
mkdir -p gpurun_out/ns
export KC_GROUP_TIMEOUT_MS=60000
nproc > gpurun_out/ns/ns_host.txt; free -g >> gpurun_out/ns/ns_host.txt
python scripts/northstar_input.py /dev/shm/h.fa > gpurun_out/ns/ns_input.log 2>&1; cat gpurun_out/ns/ns_input.log
T0=$(date +%s%N)
timeout 600 host/kmercamel compute -k 31 -g 0-7 -o /dev/shm/out8.msfa /dev/shm/h.fa 2> gpurun_out/ns/ns_cli_8gpu.log; echo "cli8 rc=$? wall $(( ($(date +%s%N) - T0) / 1000000 )) ms" | tee -a gpurun_out/ns/ns_cli_8gpu.log
tail -9 gpurun_out/ns/ns_cli_8gpu.log
T0=$(date +%s%N)
timeout 900 host/kmercamel compute -k 31 -g 0-7 -V -o /dev/shm/out8v.msfa /dev/shm/h.fa 2> gpurun_out/ns/ns_cli_8gpu_verify.log; echo "cli8 -V rc=$? wall $(( ($(date +%s%N) - T0) / 1000000 )) ms" | tee -a gpurun_out/ns/ns_cli_8gpu_verify.log
tail -5 gpurun_out/ns/ns_cli_8gpu_verify.log
T0=$(date +%s%N)
timeout 600 host/kmercamel compute -k 31 -g 0 -o /dev/shm/out1.msfa /dev/shm/h.fa 2> gpurun_out/ns/ns_cli_1gpu.log; echo "cli1 rc=$? wall $(( ($(date +%s%N) - T0) / 1000000 )) ms" | tee -a gpurun_out/ns/ns_cli_1gpu.log
tail -5 gpurun_out/ns/ns_cli_1gpu.log
md5sum /dev/shm/out8.msfa /dev/shm/out8v.msfa /dev/shm/out1.msfa | tee gpurun_out/ns/ns_md5.txt
ls -la /dev/shm/*.msfa >> gpurun_out/ns/ns_md5.txt
rm -f /dev/shm/h.fa /dev/shm/out8.msfa /dev/shm/out8v.msfa /dev/shm/out1.msfa
