#!/bin/bash
# ncu full captures of the final signature kernels on the wide-word configs (k = 63: L = 2, k = 127: L = 4, 500 Mbp)
mkdir -p gpurun_out
for C in cfg2_k63u cfg2_k127u; do
  timeout 75 ncu --set full --clock-control none --import-source on -k regex:"kc_sig_scan|kc_sig_resolve" --launch-skip 2 -c 2 -o gpurun_out/r02z_$C -f python profiles/step_for_ncu_cfg.py $C > gpurun_out/ncu_$C.log 2>&1; echo "ncu $C rc=$?"; grep "pass 1" gpurun_out/ncu_$C.log | cut -c1-300
  python scripts/ncu_summary.py gpurun_out/r02z_$C.ncu-rep gpurun_out/r02z_${C}_ncu.md "round 2 (final state): signature-bucket kernels on $C (one GPU)" > /dev/null 2>&1
  grep "duration\|DRAM read\|DRAM write\|SM throughput\|occupancy %\|registers/\|##" gpurun_out/r02z_${C}_ncu.md | cut -c1-150
done
