#!/bin/bash
# One gpurun call: variant sweep -> GPU tests under the best variant -> bench line -> ncu launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/sweep_gpu.txt 2>&1
timeout 600 python profiles/variant_sweep.py --steps 20 > gpurun_out/sweep_choice.txt 2> gpurun_out/sweep.log
tail -n 12 gpurun_out/sweep.log
CHOICE=$(tail -n 1 gpurun_out/sweep_choice.txt)
echo "choice: $CHOICE"
export $CHOICE
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_variant.log 2>&1
echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_variant.log
timeout 200 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_variant.json 2> gpurun_out/bench_variant.err
echo "bench exit $?"; cut -c1-400 gpurun_out/bench_variant.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_variant.csv python profiles/step_for_ncu.py 1 1 > gpurun_out/ncu_bench.log 2>&1
echo "ncu exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"kc_ksf_(scatter0|scatter_pf|scatter|resolve)" --launch-skip 3 -c 3 -o gpurun_out/r01h_full -f python profiles/step_for_ncu.py 1 1 > gpurun_out/r01h_full_ncu.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/*.ncu-rep
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_variant.json 2> gpurun_out/bench_ref_variant.err
echo "ref exit $?"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_variant.log 2>&1
echo "smoke exit $?"; tail -n 2 gpurun_out/smoke_variant.log
