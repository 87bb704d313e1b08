#!/usr/bin/env python
"""Benchmark of the ray-marching aggregation path (BASELINE.json metric: voxel*views/s, scenes/s, % HBM peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg2] [--impl reference]

One "step" = one synthetic scene through the whole path the reference runs per forward:
    Stage A  V x aggregate_2d_features + clear_3d_features      (rm.py:220-257)  -> volume, count, valid
    Stage B  aggregate_2d_features_ray_marching (NeuS, N=300)   (rm.py:260-307)  -> points [M, 3+C]
on the configuration BASELINE.json's metric is quoted on (configs[1]: 50 views, 256 ch, 160x120, 80x80x32).
N > 1 (torchrun, one rank per GPU): every rank lifts its own scene -- scene data-parallelism, no data-path
collective, weak scaling; `value` = all ranks' voxel*views / max-over-ranks device time.

`--impl reference` times the CPU oracle port of the reference algorithm (oracle/, OpenMP over all host cores)
on a bounded sample of the same workload; rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--threshold", type=float, default=0.05)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-views", type=int, default=0, help="views in the CPU sample (0 = auto)")
    ap.add_argument("--stage", default="both", choices=["both", "a", "b"], help="profiling aid: run one stage only")
    ap.add_argument("--mode", default="scene-dp", choices=["scene-dp", "view-sharded", "view-p2p"],
                    help="N>1: independent scenes per GPU (default) or ONE scene with its views sharded + one all-reduce")
    return ap.parse_args()


def bind_to_gpu_numa_node(torch, index):
    """Pins this rank to the CPUs next to its GPU (sysfs local_cpulist of the GPU's PCI device) so that the pinned
    host buffers of the end-to-end leg are first-touched on the GPU's own NUMA node; without it every rank's copies
    cross the socket interconnect and the N-GPU end-to-end number is capped by it.  Best effort: returns the CPU
    list used, or None."""
    try:
        props = torch.cuda.get_device_properties(index)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return spec
    except (OSError, AttributeError, ValueError):
        pass
    return None


def load_traffic(config):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the main kernels from the committed
    `ncu --set full` capture of this same command (profiles/traffic.json); {} if the config was not profiled."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(config, {})
    return {}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax = float(parts[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # median of the upper half: samples taken between steps see idle clocks
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": (load[len(load) // 2] if load else None), "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm
# ----------------------------------------------------------------------------------------------------

def cpu_step(oracle, sc, feats_nchw, views, threshold):
    """One pass of the path on the CPU oracle over the first `views` views.  Returns (seconds_a, seconds_b, M)."""
    t0 = time.perf_counter()
    oracle.aggregate_views(sc.projections[:views], feats_nchw[:views], sc.voxel_dim, sc.voxel_size, sc.origin,
                           sc.stride, mean=True)
    t1 = time.perf_counter()
    pts = oracle.aggregate_2d_features_ray_marching(sc.projections[:views], feats_nchw[:views], sc.tsdf, sc.voxel_dim,
                                                    sc.voxel_size, sc.origin, sc.stride, grids=sc.grids,
                                                    neus_threshold=threshold)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1, (0 if pts is None else pts.shape[0])


def cpu_features(sc, views):
    import numpy as np
    rng = np.random.default_rng(1000 + sc.seed)
    return rng.standard_normal((views, sc.channels, sc.height, sc.width), dtype=np.float32)


def cpu_sample_views(oracle, sc, threshold, target_s, requested):
    """How many of the scene's views one CPU pass should cover to take about `target_s` seconds (probe: 2 views)."""
    if requested:
        return min(requested, sc.views)
    probe = cpu_features(sc, 2)
    cpu_step(oracle, sc, probe, 2, threshold)               # page in, spin up the thread pool
    a, b, _ = cpu_step(oracle, sc, probe, 2, threshold)
    per_view = max((a + b) / 2.0, 1e-4)
    return int(max(2, min(sc.views, target_s / per_view)))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    from cnrma_b200 import synthetic
    oracle.build()
    cores = oracle.num_threads()
    sc = synthetic.make_scene(args.config, seed=0, with_features=False)
    # bounded sample per step: ~2 s of CPU work, less when many steps are requested, so that the whole run stays
    # within a few minutes whatever --steps says
    per_step_s = max(0.05, min(2.0, 150.0 / max(args.steps + min(args.warmup, 2), 1)))
    views = cpu_sample_views(oracle, sc, args.threshold, per_step_s, args.cpu_views)
    feats = cpu_features(sc, views)
    for _ in range(min(args.warmup, 2)):
        cpu_step(oracle, sc, feats, views, args.threshold)
    t0 = time.perf_counter()
    ta = tb = 0.0
    for _ in range(args.steps):
        a, b, m = cpu_step(oracle, sc, feats, views, args.threshold)
        ta += a
        tb += b
    dt = time.perf_counter() - t0
    vv = views * sc.nvox * args.steps
    value = vv / dt
    sample = f"first {views} of {sc.views} views of the {args.config} scene per step, both stages, C oracle + OpenMP"
    line = {
        "impl": "reference", "metric": "voxel_views_per_s", "value": value, "unit": "voxel*views/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 2), "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.config, sc), "sample": sample},
        "scenes_per_s": (views / sc.views) * args.steps / dt,
        "stage_ms": {"stage_a": 1e3 * ta / args.steps, "stage_b": 1e3 * tb / args.steps},
        "cpu_baseline": {"value": value, "unit": "voxel*views/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "voxel*views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_name(config, sc):
    nx, ny, nz = sc.voxel_dim
    return (f"{config}: {sc.views} views x {sc.channels} ch @ {sc.width}x{sc.height}, grid {nx}x{ny}x{nz} "
            f"@ {sc.voxel_size} m, NeuS march N={sc.grids}, Stage A + Stage B per scene")


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------

def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import cnrma_b200 as cn
    from cnrma_b200 import functional as F, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the aggregation path has no CPU implementation "
                         "(use --impl reference for the CPU oracle arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cn.load()

    hbm_peak, peak_src = load_peaks()
    if args.mode in ("view-sharded", "view-p2p"):
        run_view_sharded(args, cn, dev, rank, world)
        return
    sc = cn.synthetic.make_scene(args.config, seed=rank, with_features=False)
    feats = cn.synthetic.device_features(sc, dev, channels_last=True)           # [V,1,C,H,W] logical
    if sc.meta.get("dtype") == "bf16":
        feats = feats.to(torch.bfloat16)
    proj_host = torch.from_numpy(sc.projections).unsqueeze(1)     # cameras originate on the host (dataloader);
    proj = proj_host.to(dev)                                      # Stage A reads the device copy
    tsdf = torch.from_numpy(sc.tsdf).to(dev)[None, None]
    V, C, H, W = sc.views, sc.channels, sc.height, sc.width
    esz = feats.element_size()
    thr = args.threshold

    # --- one step, with CUDA events between the phases (all on the current stream) ---------------------
    fs = F._FeatureStack(F._as_view_list(feats), need_vector_layout=True)
    grid = _lib.make_grid(sc.voxel_dim, sc.voxel_size, sc.origin)
    fill_desc = fs.descriptor(0)

    def step(ev=None):
        # host half of the ray set-up first (4x4 LAPACK inverses, rm.py:96-102), so that it overlaps with the
        # GPU work already queued instead of delaying the march launch
        pinv = F.prepare_pinv(F.scale_projections(proj_host, sc.stride)[:, 0], dev) if args.stage in ("both", "b") else None
        if ev:
            ev[0].record()
        out_a = None
        if args.stage in ("both", "a"):
            out_a = F.aggregate_views(proj, feats, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride, mean=True)
        if ev:
            ev[1].record()
        rows, m_rows = None, 0
        if args.stage in ("both", "b"):
            m = F._march(fs, 0, None, tsdf[0, 0], grid, sc.voxel_dim, sc.voxel_size, sc.grids, "neus", thr, None,
                         pinv=pinv)
            if ev:
                ev[2].record()
            # fill is queued speculatively (row count of the previous scene + headroom) and the read-back of M,
            # the path's one host sync, overlaps with it; see functional._march_and_fill
            rows, res = F._march_and_fill(fs, 0, m, grid, True, None, ("bench", args.config), fill_desc)
            m_rows = int(res.rows)
        elif ev:
            ev[2].record()
        if ev:
            ev[3].record()
        return out_a, rows, m_rows

    for _ in range(max(args.warmup, 3)):
        out_a, rows, m_rows = step()
    torch.cuda.synchronize()

    # --- timed region -----------------------------------------------------------------------------------
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for k in range(args.steps):
        out_a, rows, m_rows = step(evs[k])
    t_end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = t_start.elapsed_time(t_end)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    ms_a = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    ms_march = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    ms_fill = float(np.mean([e[2].elapsed_time(e[3]) for e in evs]))

    # --- algorithmic bytes (DESIGN.md "Measurement") --------------------------------
    nvox = sc.nvox
    bytes_a = V * H * W * C * esz + nvox * C * 4 + nvox * 4 + nvox + V * 48
    bytes_fill = m_rows * (3 + C) * 4 + V * H * W * C * esz + m_rows * 8 + V * H * W * 4
    kernels = {
        "aggregate_views_kernel": {"ms": ms_a, "bytes": bytes_a, "gbs": bytes_a / ms_a / 1e6 if ms_a else None},
        "tsdf_prepare+march_neus+scan_blocks": {"ms": ms_march, "ray_steps": sc.ray_steps,
                                                "ray_steps_per_s": sc.ray_steps / ms_march * 1e3 if ms_march else None},
        "fill_rows_tma_kernel(+host read of M)": {"ms": ms_fill, "bytes": bytes_fill,
                                                  "gbs": bytes_fill / ms_fill / 1e6 if ms_fill else None},
    }
    traffic = load_traffic(args.config)
    if ms_fill >= ms_a:
        dom, ach, dbytes = "fill_rows_tma_kernel", kernels["fill_rows_tma_kernel(+host read of M)"]["gbs"], bytes_fill
    else:
        dom, ach, dbytes = "aggregate_views_kernel", kernels["aggregate_views_kernel"]["gbs"], bytes_a
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                "frac": (ach / hbm_peak) if ach else None, "traffic": traffic.get(dom), "peak_source": peak_src,
                "algorithmic_bytes": dbytes,
                "timing": "CUDA events on the launch stream around the phase, inside the timed region (includes the "
                          "launch / host-sync gaps of the phase)",
                "stage_a_frac": kernels["aggregate_views_kernel"]["gbs"] / hbm_peak if ms_a else None,
                "stage_a_traffic": traffic.get("aggregate_views_kernel")}

    value = world * sc.voxel_views / (ms_step * 1e-3)

    # --- end to end through the public API with host buffers ----------------------------------------------
    e2e = None
    if not args.no_e2e and args.stage == "both":
        e2e = run_e2e(args, cn, sc, feats, proj, tsdf, dev, world, m_rows)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # --- the lift as the detector consumes it (reported beside the headline, not part of it): Stage B fused with the
    # point-cloud hand-off (rm.py:339-407, max_points = 500000 as in the shipped config), mask drawn on the device
    handoff = None
    if args.stage == "both" and world == 1:
        def lift_and_handoff():
            n_rows = m_rows
            mask = cn.sample_points_device(n_rows, 500000, 1234, dev)
            return cn.rma_points_selected(proj_host, feats, tsdf, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride,
                                          offsets=[[0.0, 0.0, 0.0]], masks=[mask], grids=sc.grids, threshold=thr)
        for _ in range(3):
            lift_and_handoff()
        torch.cuda.synchronize()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        for _ in range(10):
            sel = lift_and_handoff()
        h1.record()
        torch.cuda.synchronize()
        handoff = {"ms_per_scene": h0.elapsed_time(h1) / 10, "rows_kept": int(sel[0][0].shape[0]), "rows_total": m_rows,
                   "what": "sample mask (device) + march + fill of the kept rows only, offset added; excludes Stage A"}
        if not args.no_e2e:
            del sel
            handoff["e2e"] = run_e2e(args, cn, sc, feats, proj, tsdf, dev, world, m_rows, handoff_rows=500000)
            handoff["e2e"]["what"] = ("the e2e leg with the lift fused with the hand-off: host features in, Stage A volume "
                                      "+ the 500000 kept point rows out (the un-sampled cloud is never materialised)")

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        cpu_baseline = run_cpu_baseline(args, sc)

    line = {
        "metric": "voxel_views_per_s", "value": value, "unit": "voxel*views/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if esz == 4 else "bf16", "data": "synthetic",
        "config": {"workload": workload_name(args.config, sc), "layout": "channels-last feature maps (NHWC physical)",
                   "l2": ("inputs (%.0f MB features/scene) exceed the 126 MB L2; no explicit flush" if V * H * W * C * esz > 126e6 else
                          "inputs (%.0f MB features/scene) FIT the 126 MB L2 and no flush is done: this configuration's "
                          "numbers include L2 reuse across steps") % (V * H * W * C * esz / 1e6),
                   "parallelism": f"scene-dp{world}", "rows_per_scene": m_rows, "threshold": thr,
                   "host_cpus": numa},
        "scenes_per_s": world / (ms_step * 1e-3),
        "ray_steps_per_s": world * sc.ray_steps / (ms_step * 1e-3),
        # the dense (Stage A) lift alone, the quantity SURVEY.md section 8d's 60 % target is stated on
        "stage_a_voxel_views_per_s": (world * sc.voxel_views / (ms_a * 1e-3)) if ms_a else None,
        "stage_ms": {"stage_a": ms_a, "march": ms_march, "fill": ms_fill},
        "kernels": kernels, "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clocks,
        "handoff": handoff,
        # aggregate_views | tsdf_prepare_slab, dist_pass (x), march_neus, scan_blocks | fill_rows_tma
        "gpu_launches": (6 if args.stage == "both" else (1 if args.stage == "a" else 5)) * args.steps,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_view_sharded(args, cn, dev, rank, world):
    """BASELINE config 4 style: ONE scene, views sharded over the ranks, dense Stage A sums + counts combined by one
    NCCL all-reduce (strong scaling; `value` = the scene's voxel*views / max-over-ranks time)."""
    import torch
    import torch.distributed as dist
    from cnrma_b200 import distributed as D
    sc = cn.synthetic.make_scene(args.config, seed=0, with_features=False)
    lo, hi = D.view_shard(sc.views, rank, world)
    full = cn.synthetic.device_features(sc, dev, channels_last=True)      # same seed on every rank
    feats = full[lo:hi]
    proj = torch.from_numpy(sc.projections).to(dev).unsqueeze(1)[lo:hi]
    del full
    p2p = args.mode == "view-p2p"      # exchange fused into the gather kernel (peer stores) instead of an all-reduce
    if p2p:
        call = lambda: D.aggregate_views_p2p(proj, feats, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride)
    else:
        call = lambda: D.aggregate_views_sharded(proj, feats, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride)
    for _ in range(max(args.warmup, 3)):
        call()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        call()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms /= args.steps
    if rank == 0:
        nbytes = sc.nvox * (sc.channels + 1) * 4
        print(json.dumps({
            "metric": "voxel_views_per_s", "value": sc.voxel_views / (ms * 1e-3), "unit": "voxel*views/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.config, sc) + " -- Stage A only, views sharded",
                       "parallelism": (f"view-shard{world} + peer stores of {nbytes * (world - 1) / world / 1e6:.0f} MB per rank "
                                       "(output stays sharded by voxel range)") if p2p else
                                      f"view-shard{world} + 1 all-reduce of {nbytes / 1e6:.0f} MB"},
            "scenes_per_s": 1e3 / ms, "gpu_launches": 2 * args.steps}))
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, cn, sc, feats, proj, tsdf, dev, world, m_rows, handoff_rows=None):
    """Same step through the public host API with HOST buffers: every step copies its inputs from pinned host
    memory to the device and its results (volume, count, points) back to pinned host memory.  Three streams
    (copy-in, compute, copy-out) with double-buffered device inputs, so the next scene's upload and the previous
    scene's download overlap with the kernels; timed with CUDA events from the first upload to the last download,
    max over ranks.  With `handoff_rows` the lift is fused with the detector's point-cloud hand-off (rm.py:339-407):
    only the kept rows are produced and downloaded."""
    import torch
    import torch.distributed as dist
    V, C, H, W = sc.views, sc.channels, sc.height, sc.width
    h_feats = torch.empty(feats.permute(0, 1, 3, 4, 2).shape, dtype=feats.dtype).pin_memory()
    h_feats.copy_(feats.permute(0, 1, 3, 4, 2))
    h_proj = proj.cpu().pin_memory()
    h_tsdf = tsdf.cpu().pin_memory()
    nx, ny, nz = sc.voxel_dim
    h_vol = torch.empty((1, nx, ny, nz, C), dtype=torch.float32).pin_memory()
    h_cnt = torch.empty((1, 1, nx, ny, nz), dtype=torch.int32).pin_memory()
    h_pts = torch.empty((handoff_rows or int(m_rows * 1.05) + 1024, 3 + C), dtype=torch.float32).pin_memory()
    ag = cn.RayMarchingAggregator(sc.voxel_size, sc.voxel_dim, origin=sc.origin.tolist(), backbone2d_stride=sc.stride,
                                  neus_threshold=args.threshold)
    s_in, s_cmp, s_out = (torch.cuda.Stream(device=dev) for _ in range(3))
    slots = [dict(feats=torch.empty_like(h_feats, device=dev), proj=torch.empty_like(h_proj, device=dev),
                  tsdf=torch.empty_like(h_tsdf, device=dev), loaded=torch.cuda.Event(), used=torch.cuda.Event())
             for _ in range(2)]

    def upload(k):
        sl = slots[k % 2]
        with torch.cuda.stream(s_in):
            s_in.wait_event(sl["used"])                      # the kernels that read this slot two steps ago are done
            sl["feats"].copy_(h_feats, non_blocking=True)
            sl["proj"].copy_(h_proj, non_blocking=True)
            sl["tsdf"].copy_(h_tsdf, non_blocking=True)
            sl["loaded"].record(s_in)

    def compute_and_download(k):
        sl = slots[k % 2]
        with torch.cuda.stream(s_cmp):
            s_cmp.wait_event(sl["loaded"])
            d_feats = sl["feats"].permute(0, 1, 4, 2, 3)
            ag.initialize_volume()
            for v in range(V):
                ag.aggregate_2d_features(sl["proj"][v], d_feats[v])
            ag.clear_3d_features()
            if handoff_rows is None:
                ag.aggregate_2d_features_ray_marching(h_proj, d_feats, sl["tsdf"])   # cameras also known on the host
                pts = ag.points_detection[0]
            else:
                draw = lambda n: cn.sample_points_device(n, handoff_rows, 1234 + k, dev)
                co, fe = cn.rma_points_selected(h_proj, d_feats, sl["tsdf"], sc.voxel_dim, sc.voxel_size, sc.origin,
                                                sc.stride, offsets=[[0.0, 0.0, 0.0]], masks=[draw], grids=sc.grids,
                                                threshold=args.threshold)
                pts = co[0]._base if co[0]._base is not None else torch.cat((co[0], fe[0]), 1)
            sl["used"].record(s_cmp)
            vol, cnt = ag.volume, ag._sum[1]
        with torch.cuda.stream(s_out):
            s_out.wait_stream(s_cmp)
            h_vol.copy_(vol.permute(0, 2, 3, 4, 1), non_blocking=True)
            h_cnt.copy_(cnt, non_blocking=True)
            h_pts[: pts.shape[0]].copy_(pts, non_blocking=True)
            for t in (pts, vol, cnt):
                t.record_stream(s_out)
        return pts.shape[0]

    steps = max(2, min(args.steps, 6))
    upload(0)
    compute_and_download(0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(s_in)
    upload(0)
    for k in range(steps):
        if k + 1 < steps:
            upload(k + 1)
        rows = compute_and_download(k)
    s_out.wait_stream(s_cmp)
    s_out.wait_stream(s_in)
    t1.record(s_out)
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms /= steps
    h2d = h_feats.numel() * h_feats.element_size() + h_proj.numel() * 4 + h_tsdf.numel() * 4
    d2h = h_vol.numel() * 4 + h_cnt.numel() * 4 + rows * (3 + C) * 4
    return {"value": world * sc.voxel_views / (ms * 1e-3), "unit": "voxel*views/s", "ms_per_step": ms,
            "scenes_per_s": world / (ms * 1e-3), "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": steps,
            "pcie_gbs": (h2d + d2h) / ms / 1e6,
            "api": "RayMarchingAggregator (host mirror of the reference detector's aggregation methods)"
                   + ("" if handoff_rows is None else " + rma_points_selected (lift fused with switch_pointcloud)")
                   + "; copy-in / compute / copy-out on three streams"}


def run_cpu_baseline(args, sc):
    """The CPU oracle port on this box's host cores, on a bounded sample of the same workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    oracle.build()
    cores = oracle.num_threads()
    views = cpu_sample_views(oracle, sc, args.threshold, 12.0, args.cpu_views)    # ~10-30 s of CPU work
    feats = cpu_features(sc, views)
    a, b, m = cpu_step(oracle, sc, feats, views, args.threshold)
    vv = views * sc.nvox
    return {"value": vv / (a + b), "unit": "voxel*views/s", "cores": cores, "kind": "port",
            "sample": f"first {views} of {sc.views} views of the {args.config} scene, both stages, one pass "
                      f"({a + b:.2f} s: stage A {a:.2f} s, stage B {b:.2f} s, {m} rows)",
            "scenes_per_s": (views / sc.views) / (a + b)}


if __name__ == "__main__":
    main()
