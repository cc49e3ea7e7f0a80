#!/bin/bash
# per-level statistics of the overlap stage on the 310 Mbp repeat-model genome
mkdir -p gpurun_out
KC_TRACE=1 timeout 600 python profiles/path_stage_profile.py cfg4_human_310M 2> gpurun_out/levels_trace.log | cut -c1-1500
grep "level d=\|small engine" gpurun_out/levels_trace.log | tail -34 | cut -c1-400
