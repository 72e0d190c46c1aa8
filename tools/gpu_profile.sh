#!/bin/bash
# Runs ON the GPU box (gpurun -- bash tools/gpu_profile.sh): bench lines first (never under a profiler), then the ncu
# captures that profiles/ is condensed from (tools/launch_table.py, tools/ncu_summary.py, tools/ncu_lines.py).
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2>> gpurun_out/bench.err
python bench.py --workload c1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c1.json 2>> gpurun_out/bench.err
# launch list of bench.py itself (warm-up, 2 timed steps, e2e + CLI-stage solves)
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
# launch list of plain solves (tools/prof_run.py c2 2)
python tools/prof_run.py c2 2 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python tools/prof_run.py c2 2 > gpurun_out/ncu_l.log 2>&1
# full sections + source counters of the per-contig / per-segment kernels of the second solve
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:FnRelaxSeg|FnRelaxSweep|FnTopoSeg|FnHeaps|FnEnum>|FnMainTrace|FnTasksA0|FnTasksA1Solo|FnTasksB|FnParts' \
    -s 12 -c 12 -f -o gpurun_out/prof_r1 python tools/prof_run.py c2 2 > gpurun_out/ncu_f.log 2>&1
AA_TRACE=1 python tools/prof_run.py c2 2 > gpurun_out/trace_c2.log 2>&1
AA_TRACE=1 python tools/prof_run.py c3 2 > gpurun_out/trace_c3.log 2>&1
AA_SEG_DEBUG=1 python tools/prof_run.py c2 1 2> gpurun_out/seg_c2.log > /dev/null
AA_SEG_DEBUG=1 python tools/prof_run.py c3 1 2> gpurun_out/seg_c3.log > /dev/null
tail -c 600 gpurun_out/bench.json
