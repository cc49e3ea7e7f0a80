#!/bin/bash
# ncu full capture of the two signature kernels (second step: warm)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"kc_sig_scan|kc_sig_resolve" --launch-skip 2 -c 2 -o gpurun_out/r02_sig -f python profiles/step_for_ncu.py 1 1 > gpurun_out/ncu_sig.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py gpurun_out/r02_sig.ncu-rep gpurun_out/r02_sig_ncu.md "round 2: signature-bucket kernels on configs[1]" | grep -v "^$" | head -70
