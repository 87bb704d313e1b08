#!/bin/bash
# Runs on the GPU box (under gpurun): ncu launch list + one full capture of the main kernels for the default bench
# command.  Outputs land in gpurun_out/; `python profiles/ingest.py rNN` turns them into the committed summaries.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
# every launch of the timed region (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -s 18 -c 18 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/launches_bench.log 2>&1
# the three main kernels, once each
ncu --set full --clock-control none --import-source on -k regex:"aggregate_views_kernel|march_neus_kernel|fill_rows_tma_kernel" -s 9 -c 3 -o gpurun_out/prof_main $CMD > gpurun_out/prof_main.log 2>&1
# the real (un-profiled) bench line with clocks
python bench.py > gpurun_out/bench_line.json 2> gpurun_out/bench_line.err
tail -c 400 gpurun_out/bench_line.json
ls -la gpurun_out
