#!/bin/bash
# round 2 profiles: reference wall clock on the box's host core (background), ncu launch list + full captures, compute-sanitizer,
# CLI wall clocks
mkdir -p gpurun_out
export KC_GROUP_TIMEOUT_MS=20000
nproc > gpurun_out/prof_host.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/prof_host.txt
# --- the reference on the 310 Mbp scale model of the north-star input, one host core (the reference is single-threaded)
( python scripts/northstar_input.py /dev/shm/h310.fa small > gpurun_out/ref310_input.log 2>&1
  T0=$(date +%s%N)
  oracle/_ref/kmercamel compute -k 31 -o /dev/shm/ref310.msfa /dev/shm/h310.fa 2> gpurun_out/ref310_cli.log
  echo "reference rc=$? wall_ms=$(( ($(date +%s%N) - T0) / 1000000 ))" >> gpurun_out/ref310_cli.log
  T0=$(date +%s%N)
  host/kmercamel compute -k 31 -V -o /dev/shm/ours310.msfa /dev/shm/h310.fa 2> gpurun_out/ours310_cli.log
  echo "ours rc=$? wall_ms=$(( ($(date +%s%N) - T0) / 1000000 ))" >> gpurun_out/ours310_cli.log
  ls -l /dev/shm/ref310.msfa /dev/shm/ours310.msfa >> gpurun_out/ours310_cli.log
  rm -f /dev/shm/h310.fa /dev/shm/ref310.msfa /dev/shm/ours310.msfa ) &
REFPID=$!
# --- launch list of one step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv python profiles/step_for_ncu.py 1 1 > gpurun_out/ncu_step.log 2>&1; echo "ncu list rc=$?"
N=$(grep -c "step 1" gpurun_out/ncu_step.log); L=$(grep "step 1" gpurun_out/ncu_step.log | sed 's/.*launches=\([0-9]*\).*/\1/')
python scripts/summarize_launches.py gpurun_out/launches.csv ${L:-20} gpurun_out/launches.md | tail -25
# --- full captures of the k-mer set kernels, the small engine and the emission (second step: warm)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"kc_ksf_scatter0|kc_ksf_scatter_pf|kc_ksf_resolve2|kc_small_engine|kc_emit" --launch-skip-before-match 0 -c 12 -o gpurun_out/r02_full -f python profiles/step_for_ncu.py 1 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
python scripts/ncu_summary.py gpurun_out/r02_full.ncu-rep gpurun_out/r02_full_ncu.md "round 2: configs[1] step, TMA leaf / tile streaming" > /dev/null 2>&1; grep -c "##" gpurun_out/r02_full_ncu.md
# --- compute-sanitizer memcheck over small-input tests of every path (fixed-slot + exact construction, engine, emission, group protocol)
timeout 1500 compute-sanitizer --tool memcheck --target-processes all --log-file gpurun_out/sanitizer_memcheck.log python -m pytest tests/test_gpu.py tests/test_sharded.py tests/test_gpu_sparse.py -m gpu -q -x -k "edge_inputs or kats or global_sparse or small_and_empty or test_fa or sparse_switch or overflow_falls_back" > gpurun_out/sanitizer_pytest.log 2>&1; echo "sanitizer rc=$?"
tail -3 gpurun_out/sanitizer_pytest.log; grep -c "ERROR SUMMARY" gpurun_out/sanitizer_memcheck.log; grep "ERROR SUMMARY" gpurun_out/sanitizer_memcheck.log | sort | uniq -c | head
# --- CLI wall clocks on configs[1]
timeout 600 python profiles/cli_wallclock.py > gpurun_out/cli_wallclock.json 2> gpurun_out/cli_wallclock.err; echo "cli wallclock rc=$?"; cut -c1-900 gpurun_out/cli_wallclock.json
wait $REFPID
cat gpurun_out/ref310_cli.log | tail -4; tail -4 gpurun_out/ours310_cli.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], json.dumps(d['roofline'])[:900])"
