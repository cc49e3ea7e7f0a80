#!/bin/bash
# round 2, final: full GPU suite, memcheck on the small-input tests of the kernels changed last (signature scan / resolve, small engine,
# runs, emission), bench (both arms), launch list + ncu captures of the signature kernels and the small engine
mkdir -p gpurun_out
export KC_GROUP_TIMEOUT_MS=20000
timeout 2400 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_full.log
tail -12 gpurun_out/pytest_gpu_full.log | cut -c1-200
timeout 900 compute-sanitizer --tool memcheck --target-processes all --log-file gpurun_out/z_memcheck.log python -m pytest tests/test_gpu_sig.py tests/test_gpu.py tests/test_sharded.py -m gpu -q -x -k "sig_edge_inputs or sig_overflow or (sig_matches_exact and (31-True or 29-True or 63-False or 127-True)) or in_process_matches_single_gpu or overlap_path_kats or overlap_path_sparse_kats or compute_S_fuzz or small_engine_and_host_levels or compute_small_cases or random_vs_oracle[1]" > gpurun_out/z_sanitizer_pytest.log 2>&1; echo "sanitizer rc=$?"
tail -3 gpurun_out/z_sanitizer_pytest.log; grep "ERROR SUMMARY" gpurun_out/z_memcheck.log | sort | uniq -c | head -5
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['ms_per_step'], d['ms_per_step_with_kernel_timers'], d['e2e']['ms_per_step'], d['value']/1e9, json.dumps(d['roofline'])[:900]); print(json.dumps(d['cpu_baseline'])[:300]); r=json.load(open('gpurun_out/bench_ref.json')); print(r['value'], r['ms_per_step'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python profiles/step_for_ncu.py 1 1 > gpurun_out/ncu_step.log 2>&1; echo "ncu list rc=$?"
L=$(grep "step 1" gpurun_out/ncu_step.log | sed 's/.*launches=\([0-9]*\).*/\1/'); python scripts/summarize_launches.py gpurun_out/launches.csv ${L:-17} gpurun_out/launches.md | tail -22
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"kc_sig_scan|kc_sig_resolve|kc_small_engine|kc_small_level" --launch-skip 4 -c 4 -o gpurun_out/r02z_final -f python profiles/step_for_ncu.py 1 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
python scripts/ncu_summary.py gpurun_out/r02z_final.ncu-rep gpurun_out/r02z_final_ncu.md "round 2 (final state): signature-bucket kernels and the small engine on configs[1]" > /dev/null 2>&1; grep -c "##" gpurun_out/r02z_final_ncu.md
