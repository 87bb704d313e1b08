#!/bin/bash
# Runs on the GPU box (under gpurun): ncu launch list + full captures of the main kernels for the bench command, and the
# un-profiled bench line.  Outputs land in gpurun_out/$1 (default r02); `python profiles/ingest.py r02 gpurun_out/r02`
# turns them into the committed summaries.
set -u
O=gpurun_out/${1:-r02}
mkdir -p $O
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --scenes 1"
# every launch of the timed region (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -s 18 -c 18 --csv --log-file $O/launches.csv $CMD > $O/launches_bench.log 2>&1
# the three main kernels of the headline configuration, once each
ncu --set full --clock-control none --import-source on -k regex:"aggregate_views_kernel|march_neus_kernel|fill_rows_tma_kernel" -s 9 -c 3 -o $O/prof_main $CMD > $O/prof_main.log 2>&1
$CMD > $O/bench_one_scene.json 2> /dev/null
# Stage A of the fine grid (cfg 4, 256 channels: column units with view culling) and of the long-ray configuration
# (cfg 5: the long-list kernel)
ncu --set full --clock-control none --import-source on -k regex:"aggregate_views" -s 4 -c 1 -o $O/prof_cfg4 python bench.py --config cfg4 --steps 2 --warmup 3 --stage a --no-e2e --no-cpu-baseline --scenes 1 > $O/prof_cfg4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"aggregate_views" -s 4 -c 1 -o $O/prof_cfg5 python bench.py --config cfg5 --steps 2 --warmup 3 --stage a --no-e2e --no-cpu-baseline --scenes 1 > $O/prof_cfg5.log 2>&1
# the real (un-profiled) bench line with clocks, both arms
python bench.py > $O/bench_line.json 2> $O/bench_line.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err
# one kernel-only line per BASELINE configuration
for c in cfg1 cfg3 cfg4 cfg4_c32 cfg5 ref_test; do
  python bench.py --config $c --steps 20 --no-e2e --no-cpu-baseline --scenes 2 > $O/bench_$c.json 2> $O/bench_$c.err
done
tail -c 300 $O/bench_line.json
ls -la $O
