#!/bin/bash
# Runs on the GPU box: one bench line per BASELINE config (kernel-only; e2e / CPU legs off) -> gpurun_out/configs.jsonl
mkdir -p gpurun_out
: > gpurun_out/configs.jsonl
for c in cfg1 cfg2 cfg3 cfg4 cfg4_c32 cfg5 ref_test; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/configs.jsonl
done
python - <<'PY'
import json
print("| config | workload | ms/scene | stage A ms | march ms | fill ms | rows M | voxel*views/s | stage A % HBM | fill % HBM |")
print("|---|---|---|---|---|---|---|---|---|---|")
for line in open("gpurun_out/configs.jsonl"):
    d = json.loads(line)
    w = d["config"]["workload"]
    s = d["stage_ms"]
    k = d["kernels"]
    fill = [v for n, v in k.items() if n.startswith("fill")][0]
    print(f"| {w.split(':')[0]} | {w.split(':')[1].split(', NeuS')[0].strip()} | {d['ms_per_step']:.3f} | {s['stage_a']:.3f} | {s['march']:.3f} | {s['fill']:.3f} | "
          f"{d['config']['rows_per_scene']} | {d['value']:.3g} | {100 * d['roofline']['stage_a_frac']:.1f} | {100 * fill['gbs'] / d['roofline']['peak']:.1f} |")
PY
