#!/bin/bash
# Runs ON the GPU box (gpurun -- bash tools/gpu_profile.sh): bench lines first (never under a profiler), then the ncu
# captures that profiles/ is condensed from (tools/launch_table.py, tools/ncu_summary.py, tools/ncu_lines.py).
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err || exit 1
[ -n "$SKIP_REF" ] || python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2>> gpurun_out/bench.err
python bench.py --workload c1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c1.json 2>> gpurun_out/bench.err
# launch list of bench.py itself (the default job: c5 on one GPU, then the c2 sub-object; warm-up, 2 timed steps, e2e solves)
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
# launch list of plain solves (tools/prof_run.py c2 2)
python tools/prof_run.py c2 2 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python tools/prof_run.py c2 2 > gpurun_out/ncu_l.log 2>&1
# full sections + source counters of the serial / per-contig kernels (both solves of prof_run; the summary keeps the second)
[ -n "$SKIP_FULL" ] || ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:FnHeaps>|FnHeapsLeaf|FnEnum>|FnXRec|FnOpsFill|FnRootFill|FnMainTrace|FnTasksA0|FnParts|FnRelaxSeg' \
    -c 24 -f -o gpurun_out/prof_r2 python tools/prof_run.py c2 2 > gpurun_out/ncu_f.log 2>&1
[ -n "$SKIP_TRACE" ] || AA_TRACE=1 python tools/prof_run.py c2 2 > gpurun_out/trace_c2.log 2>&1
[ -n "$SKIP_TRACE" ] || AA_TRACE=1 python tools/prof_run.py c3 2 > gpurun_out/trace_c3.log 2>&1
tail -c 600 gpurun_out/bench.json
