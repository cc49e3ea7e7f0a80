#!/bin/bash
# usage: gpu_prof.sh <tag> <kernel-regex> [count]   -- ncu --set full of the matching kernels on the bench workload (2 passes)
TAG=${1:-prof}; RX=${2:-kc_}; CNT=${3:-24}
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$RX" -c $CNT -o gpurun_out/$TAG -f python profiles/step_for_ncu.py 1 1 > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${TAG}_ncu.log
