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
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--threshold", type=float, default=0.05)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-views", type=int, default=0, help="views in the CPU sample (0 = auto)")
    ap.add_argument("--stage", default="both", choices=["both", "a", "b"], help="profiling aid: run one stage only")
    ap.add_argument("--mode", default="scene-dp", choices=["scene-dp", "view-suite"],
                    help="scene-dp (default): independent scenes per GPU, plus at N>1 the view-sharded record of ONE scene; "
                         "view-suite: only that record")
    ap.add_argument("--no-view-sharded", action="store_true", help="N>1: skip the view-sharded record")
    ap.add_argument("--scenes", type=int, default=4, help="seeded scenes per rank, rotated step by step")
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
# CPU arms: the UNMODIFIED reference (oracle/_ref through oracle/ref_shim.py) and the C oracle port
# ----------------------------------------------------------------------------------------------------

def host_threads():
    """All the host cores this process may use: the affinity mask, not OMP_NUM_THREADS (torchrun exports 1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def sample_view_ids(total, n):
    """`n` view indices spread evenly over the scene's views (the cost of a view depends on where its camera looks)."""
    n = max(1, min(int(n), total))
    return sorted({int(round(i * (total - 1) / max(n - 1, 1))) for i in range(n)}) if n > 1 else [total // 2]


def cpu_features(sc, view_ids):
    """fp32 NCHW maps for the sampled views (the reference's layout), seeded by the scene."""
    import numpy as np
    rng = np.random.default_rng(1000 + sc.seed)
    return rng.standard_normal((len(view_ids), sc.channels, sc.height, sc.width), dtype=np.float32)


def port_step(oracle, sc, feats_nchw, view_ids, threshold):
    """One pass of the path on the C oracle over the views `view_ids`.  Returns (seconds_a, seconds_b, M)."""
    proj = sc.projections[view_ids]
    t0 = time.perf_counter()
    oracle.aggregate_views(proj, feats_nchw, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride, mean=True)
    t1 = time.perf_counter()
    pts = oracle.aggregate_2d_features_ray_marching(proj, feats_nchw, sc.tsdf, sc.voxel_dim, sc.voxel_size, sc.origin,
                                                    sc.stride, grids=sc.grids, neus_threshold=threshold)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1, (0 if pts is None else pts.shape[0])


def load_cpu_arms():
    """(oracle module, ref_runner module or None, threads): both CPU implementations set to all host threads."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    oracle.build()
    threads = host_threads()
    oracle.set_threads(threads)
    ref_runner = None
    try:
        import ref_runner as rr
        if rr.available():
            ref_runner = rr
    except Exception:
        ref_runner = None
    return oracle, ref_runner, threads


def port_record(oracle, sc, threshold, threads):
    """The C oracle port (OpenMP) on ALL views of the scene, one pass: the second CPU number beside the reference's."""
    ids = list(range(sc.views))
    feats = cpu_features(sc, ids)
    port_step(oracle, sc, feats[:2], ids[:2], threshold)                 # page in, spin up the thread pool
    a, b, m = port_step(oracle, sc, feats, ids, threshold)
    return {"value": sc.voxel_views / (a + b), "unit": "voxel*views/s", "cores": threads, "kind": "port",
            "sample": f"all {sc.views} views, both stages, one pass of the C oracle + OpenMP ({a + b:.2f} s: stage A {a:.2f} s, "
                      f"stage B {b:.2f} s, {m} rows)", "scenes_per_s": 1.0 / (a + b)}


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation of the path (projects/mvsdetection/models/
    ray_marching.py, unmodified, staged under oracle/_ref by oracle/build_ref.py and imported through
    oracle/ref_shim.py) on all host threads.  Each step is a bounded sample of the workload -- a few evenly spaced views
    of the scene through both stages -- sized from a probe so that `--steps K --warmup W` ends within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cnrma_b200 import synthetic
    oracle, ref_runner, threads = load_cpu_arms()
    sc = synthetic.make_scene(args.config, seed=0, with_features=False, **scene_override(args.config))
    passes = args.steps + args.warmup
    budget_s = 170.0
    if ref_runner is not None:
        kind = "reference"
        probe_ids = sample_view_ids(sc.views, 2)
        probe = cpu_features(sc, probe_ids)
        ref_runner.reference_pass(sc, probe, probe_ids, args.threshold, threads)          # import, page in
        a, b, _ = ref_runner.reference_pass(sc, probe, probe_ids, args.threshold, threads)
        per_view = max((a + b) / 2.0, 1e-3)
        target = budget_s / max(passes, 1)
        n_views = args.cpu_views or int(max(1, min(sc.views, target / per_view)))
        # the reference's cost per view GROWS with the number of views (rm.py:289-293 re-copies the growing point cloud at
        # every view), so the two-view probe flatters it: calibrate on the chosen sample and shrink it until a pass fits
        for _ in range(2):
            if args.cpu_views or n_views <= 1:
                break
            ids = sample_view_ids(sc.views, n_views)
            a, b, _ = ref_runner.reference_pass(sc, cpu_features(sc, ids), ids, args.threshold, threads)
            if a + b <= 1.2 * target:
                break
            n_views = max(1, int(n_views * (target / (a + b)) ** 0.75))
        ids = sample_view_ids(sc.views, n_views)
        feats = cpu_features(sc, ids)
        step = lambda: ref_runner.reference_pass(sc, feats, ids, args.threshold, threads)
        what = "the reference's own Python (ray_marching.py, unmodified, torch CPU)"
    else:
        kind = "port"
        ids = list(range(sc.views))
        feats = cpu_features(sc, ids)
        step = lambda: port_step(oracle, sc, feats, ids, args.threshold)
        what = "C oracle port + OpenMP (oracle/_ref is not staged on this box)"
    for _ in range(args.warmup):
        step()
    ta = tb = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        a, b, m = step()
        ta += a
        tb += b
    dt = time.perf_counter() - t0
    value = len(ids) * sc.nvox * args.steps / dt
    sample = (f"{len(ids)} of {sc.views} views (evenly spaced) of the {args.config} scene per step, both stages, {what}, "
              f"{threads} threads")
    base = {"value": value, "unit": "voxel*views/s", "cores": threads, "kind": kind, "sample": sample}
    line = {
        "impl": "reference", "metric": "voxel_views_per_s", "value": value, "unit": "voxel*views/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, sc, args.gpus),
        "scenes_per_s": (len(ids) / sc.views) * args.steps / dt,
        "stage_ms": {"stage_a": 1e3 * ta / args.steps, "stage_b": 1e3 * tb / args.steps},
        "cpu_baseline": base,
        "port": port_record(oracle, sc, args.threshold, threads) if kind == "reference" else None,
        "e2e": {"value": value, "unit": "voxel*views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_cpu_baseline(args, sc):
    """`cpu_baseline` of the GPU arm's line: the reference itself on this box's host cores over a bounded sample of the
    same workload (about 10-30 s), with the C port's full-scene number beside it."""
    oracle, ref_runner, threads = load_cpu_arms()
    port = port_record(oracle, sc, args.threshold, threads)
    if ref_runner is None:
        return port
    probe_ids = sample_view_ids(sc.views, 2)
    probe = cpu_features(sc, probe_ids)
    ref_runner.reference_pass(sc, probe, probe_ids, args.threshold, threads)
    a, b, _ = ref_runner.reference_pass(sc, probe, probe_ids, args.threshold, threads)
    per_view = max((a + b) / 2.0, 1e-3)
    ids = sample_view_ids(sc.views, args.cpu_views or int(max(2, min(sc.views, 15.0 / per_view))))
    a, b, m = ref_runner.reference_pass(sc, cpu_features(sc, ids), ids, args.threshold, threads)
    return {"value": len(ids) * sc.nvox / (a + b), "unit": "voxel*views/s", "cores": threads, "kind": "reference",
            "sample": f"{len(ids)} of {sc.views} views (evenly spaced) of the {args.config} scene, both stages, one pass of "
                      f"the reference's own Python (ray_marching.py, unmodified, torch CPU; {a + b:.2f} s: stage A {a:.2f} s, "
                      f"stage B {b:.2f} s, {m} rows)",
            "scenes_per_s": (len(ids) / sc.views) / (a + b), "port": port}


def scene_override(config):
    """Bench scenes rotate seeds; room scenes use the seed-dependent room (walls / truncation vary) so that the number of
    rows the march keeps differs from scene to scene."""
    from cnrma_b200 import synthetic
    return {"tsdf": "room_var"} if synthetic.CONFIGS[config]["tsdf"] == "room" else {}


def workload_name(config, sc):
    nx, ny, nz = sc.voxel_dim
    return (f"{config}: {sc.views} views x {sc.channels} ch @ {sc.width}x{sc.height}, grid {nx}x{ny}x{nz} "
            f"@ {sc.voxel_size} m, NeuS march N={sc.grids}, Stage A + Stage B per scene")


def bench_config(args, sc, world):
    """The `config` object of the JSON line -- identical in the GPU arm and the reference arm."""
    feat_mb = sc.views * sc.height * sc.width * sc.channels * (2 if sc.meta.get("dtype") == "bf16" else 4) / 1e6
    return {"workload": workload_name(args.config, sc),
            "layout": "channels-last feature maps (NHWC physical)",
            "l2": ("inputs (%.0f MB features/scene, %d scenes rotated) exceed the 126 MB L2; no explicit flush" if feat_mb > 126 else
                   "inputs (%.0f MB features/scene, %d scenes rotated) may FIT the 126 MB L2 and no flush is done: this "
                   "configuration's numbers include L2 reuse across steps") % (feat_mb, args.scenes),
            "parallelism": f"scene-dp{world}", "threshold": args.threshold,
            "scenes": f"{args.scenes} seeded scenes (the same on every rank, rank-dependent start), rotated step by step "
                      "(cameras and room differ: M varies)"}


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
    from cnrma_b200 import functional as F

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the aggregation path has no CPU implementation "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cn.load()

    hbm_peak, peak_src = load_peaks()
    if args.mode == "view-suite":
        run_view_sharded(args, cn, dev, rank, world)
        return

    # --- the scenes of this rank: `--scenes` seeds, rotated step by step, inputs resident in HBM ---------------
    thr = args.thr = args.threshold
    override = scene_override(args.config)
    scenes = []
    for i in range(args.scenes):
        # the same seeds on every rank (weak scaling: identical work per rank), visited from a rank-dependent start, so
        # that at any step the ranks are on different scenes
        sc = cn.synthetic.make_scene(args.config, seed=i, with_features=False, **override)
        feats = cn.synthetic.device_features(sc, dev, channels_last=True)           # [V,1,C,H,W] logical
        if sc.meta.get("dtype") == "bf16":
            feats = feats.to(torch.bfloat16)
        proj_host = torch.from_numpy(sc.projections).unsqueeze(1)     # cameras originate on the host (dataloader);
        scenes.append(dict(sc=sc, feats=feats, proj_host=proj_host, proj=proj_host.to(dev),   # Stage A reads the device copy
                           tsdf=torch.from_numpy(sc.tsdf).to(dev)[None, None]))
    sc = scenes[0]["sc"]
    V, C, H, W = sc.views, sc.channels, sc.height, sc.width
    esz = scenes[0]["feats"].element_size()
    grid_args = (sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride)

    # --- one step = one scene through the PUBLIC API: cn.aggregate_views + cn.rma_points ------------------------
    marks = {}
    F.set_phase_hook(lambda label: marks["ev"][2].record() if marks.get("ev") else None)

    def step(k, ev=None):
        s_ = scenes[(k + rank) % len(scenes)]
        marks["ev"] = ev
        if ev:
            ev[0].record()
        out_a = None
        if args.stage in ("both", "a"):
            out_a = cn.aggregate_views(s_["proj"], s_["feats"], *grid_args, mean=True)
        if ev:
            ev[1].record()
        rows, m_rows = None, 0
        if args.stage in ("both", "b"):
            # rma_points inverts the cameras on the host (LAPACK, like the reference), queues the march, then the fill --
            # speculatively, sized from recent scenes -- and reads M back while the fill runs (the path's one host sync)
            rows = cn.rma_points(s_["proj_host"], s_["feats"], s_["tsdf"], *grid_args, grids=sc.grids, threshold=thr)[0]
            m_rows = int(rows.shape[0])
        elif ev:
            ev[2].record()
        if ev:
            ev[3].record()
        return out_a, rows, m_rows

    for k in range(max(args.warmup, 3, len(scenes))):
        out_a, rows, m_rows = step(k)
    torch.cuda.synchronize()
    F.fill_stats(reset=True)

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
    rows_seen = []
    for k in range(args.steps):
        out_a, rows, m_rows = step(k, evs[k])
        rows_seen.append(m_rows)
    t_end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    F.set_phase_hook(None)
    fill = F.fill_stats()
    ms_total = t_start.elapsed_time(t_end)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    ms_a = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    ms_march = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    ms_fill = float(np.mean([e[2].elapsed_time(e[3]) for e in evs]))
    m_mean = float(np.mean(rows_seen)) if rows_seen else 0.0

    # --- algorithmic bytes (DESIGN.md "Measurement") --------------------------------
    nvox = sc.nvox
    bytes_a = V * H * W * C * esz + nvox * C * 4 + nvox * 4 + nvox + V * 48
    bytes_fill = m_mean * (3 + C) * 4 + V * H * W * C * esz + m_mean * 8 + V * H * W * 4
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
                          "launch / host-sync gaps of the phase); bytes use the mean row count of the rotated scenes",
                "stage_a_frac": kernels["aggregate_views_kernel"]["gbs"] / hbm_peak if ms_a else None,
                "stage_a_traffic": traffic.get("aggregate_views_kernel")}
    if traffic.get("_fill_rows") and dom == "fill_rows_tma_kernel":
        # the ncu capture ran ONE scene (seed 0): its row count and the algorithmic bytes of that launch, so that
        # `traffic` is compared with the bytes of the launch it was measured on, not with the mean over the rotated scenes
        m0 = traffic["_fill_rows"]
        roofline["traffic_rows"] = m0
        roofline["traffic_algorithmic_bytes"] = m0 * (3 + C) * 4 + V * H * W * C * esz + m0 * 8 + V * H * W * 4

    value = world * sc.voxel_views / (ms_step * 1e-3)

    # --- end to end through the public API with host buffers ----------------------------------------------
    e2e = None
    if not args.no_e2e and args.stage == "both":
        e2e = run_e2e(args, cn, scenes, dev, world, max(rows_seen))
        if world > 1:                                       # one probe of the host's copy ceiling beside the number
            e2e["host_copy_ceiling"] = host_copy_probe(dev, world)

    # --- the view-sharded forms of ONE large scene (BASELINE config 4), N > 1 only ------------------------------
    view_sharded = None
    if world > 1 and not args.no_view_sharded and args.stage == "both":
        del scenes[1:]
        torch.cuda.empty_cache()
        try:
            view_sharded = view_sharded_suite(cn, dev, rank, world, config="cfg4", steps=10, warmup=3)
        except Exception as exc:   # e.g. no peer access on this box: keep the scene-parallel line (all ranks fail alike)
            print(f"view-sharded record skipped on rank {rank}: {exc!r}", file=sys.stderr)
            view_sharded = {"error": repr(exc)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    extras = {}
    if args.stage == "both" and world == 1:
        s0 = scenes[0]
        # --- the lift as the detector consumes it (reported beside the headline, not part of it): Stage B fused with the
        # point-cloud hand-off (rm.py:339-407, max_points = 500000 as in the shipped config), mask drawn on the device
        def lift_and_handoff(k):
            s_ = scenes[k % len(scenes)]
            return cn.rma_points_selected(s_["proj_host"], s_["feats"], s_["tsdf"], *grid_args, offsets=[[0.0, 0.0, 0.0]],
                                          max_points=500000, device_seed=1234 + k, grids=sc.grids, threshold=thr)
        for k in range(len(scenes) + 2):
            lift_and_handoff(k)
        torch.cuda.synchronize()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        for k in range(12):
            sel = lift_and_handoff(k)
        h1.record()
        torch.cuda.synchronize()
        handoff = {"ms_per_scene": h0.elapsed_time(h1) / 12, "rows_kept": int(sel[0][0].shape[0]),
                   "what": "march + device-side sample mask + fill of the kept rows only, offset added, queued without a "
                           "host stall (M is read back last); excludes Stage A"}
        if not args.no_e2e:
            del sel
            handoff["e2e"] = run_e2e(args, cn, scenes, dev, world, max(rows_seen), handoff_rows=500000)
            handoff["e2e"]["what"] = ("the e2e leg with the lift fused with the hand-off: host features in, Stage A volume "
                                      "+ the 500000 kept point rows out (the un-sampled cloud is never materialised)")
        extras["handoff"] = handoff
        # --- the reference's NCHW backbone output instead of channels-last maps: conversion kernel + Stage A
        nchw = s0["feats"].contiguous()
        for _ in range(3):
            cn.aggregate_views(s0["proj"], nchw, *grid_args, mean=True)
        torch.cuda.synchronize()
        h0.record()
        for _ in range(10):
            cn.aggregate_views(s0["proj"], nchw, *grid_args, mean=True)
        h1.record()
        torch.cuda.synchronize()
        extras["nchw_input_ms"] = {"stage_a_ms": h0.elapsed_time(h1) / 10,
                                   "what": "Stage A from NCHW maps (the reference backbone's layout): "
                                           "cnrma_to_channels_last + the gather kernel"}
        del nchw
        # --- worst case for the march: a random TSDF of the same size (no empty space to skip, many kept samples)
        rnd = torch.rand(s0["tsdf"].shape, device=dev, generator=torch.Generator(device=dev).manual_seed(7)) * 2 - 1
        for _ in range(2):
            F.rma_dense_weights(s0["proj_host"], H, W, rnd, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride, grids=sc.grids,
                                threshold=thr)
        torch.cuda.synchronize()
        marks["ev"] = None
        t_m = []
        for _ in range(5):
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            fsn = type("G", (), dict(device=dev, V=V, H=H, W=W))()
            P_scaled = F.scale_projections(s0["proj_host"], sc.stride)
            pinv = F.prepare_pinv(P_scaled[:, 0], dev)
            a_.record()
            F._march(fsn, 0, None, rnd[0, 0], cn._lib.make_grid(sc.voxel_dim, sc.voxel_size, sc.origin), sc.voxel_dim,
                     sc.voxel_size, sc.grids, "neus", thr, None, pinv=pinv)
            b_.record()
            torch.cuda.synchronize()
            t_m.append(a_.elapsed_time(b_))
        extras["march_worst_case"] = {"ms": float(np.median(t_m)), "ray_steps_per_s": sc.ray_steps / np.median(t_m) * 1e3,
                                      "what": "pre-pass + march + scan on a uniform random TSDF of the same grid (nothing "
                                              "to skip); the headline scenes are rooms"}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        cpu_baseline = run_cpu_baseline(args, sc)

    line = {
        "metric": "voxel_views_per_s", "value": value, "unit": "voxel*views/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3, len(scenes)), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if esz == 4 else "bf16", "data": "synthetic",
        "config": bench_config(args, sc, world),
        "api": "cnrma_b200.aggregate_views + cnrma_b200.rma_points (the package's public functions), one call each per scene",
        "rows_per_scene": {"mean": m_mean, "min": min(rows_seen), "max": max(rows_seen)},
        "speculative_fill": {**fill, "miss_rate": fill["misses"] / max(fill["calls"], 1)},
        "host_cpus": numa,
        "scenes_per_s": world / (ms_step * 1e-3),
        "ray_steps_per_s": world * sc.ray_steps / (ms_step * 1e-3),
        # the dense (Stage A) lift alone, the quantity SURVEY.md section 8d's 60 % target is stated on
        "stage_a_voxel_views_per_s": (world * sc.voxel_views / (ms_a * 1e-3)) if ms_a else None,
        "stage_ms": {"stage_a": ms_a, "march": ms_march, "fill": ms_fill},
        "kernels": kernels, "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clocks,
        "view_sharded": view_sharded,
        # aggregate_views | tsdf_prepare_slab, dist_pass (x), march_neus, scan_blocks | fill_rows_tma
        "gpu_launches": (6 if args.stage == "both" else (1 if args.stage == "a" else 5)) * args.steps + fill["misses"],
    }
    line.update(extras)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


NVLINK_GBS = 770.0   # measured peer-copy bandwidth per direction per GPU on this pool (B200_PROFILING.md); nominal 900


def view_sharded_suite(cn, dev, rank, world, config="cfg4", steps=20, warmup=3, modes=None):
    """ONE scene (BASELINE config 4: 0.04 m voxels, 160x160x64, 50 views x 256 ch), Stage A, its views sharded over the
    ranks -- every multi-GPU form the package has, each timed with CUDA events (max over ranks) against the same
    scene on one GPU.  Returns the `view_sharded` record of the bench line: per mode ms, the bytes one rank moves
    over NVLink, the NVLink floor (bytes / measured 770 GB/s) and the speed-up over one GPU."""
    import torch
    import torch.distributed as dist
    from cnrma_b200 import distributed as D
    sc = cn.synthetic.make_scene(config, seed=0, with_features=False)
    lo, hi = D.view_shard(sc.views, rank, world)
    full = cn.synthetic.device_features(sc, dev, channels_last=True)      # same seed on every rank
    proj = torch.from_numpy(sc.projections).to(dev).unsqueeze(1)
    feats, projs = full[lo:hi], proj[lo:hi]
    args = (sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride)
    vol_bytes = sc.nvox * (sc.channels + 1) * 4
    feat_bytes = sc.views * sc.height * sc.width * sc.channels * full.element_size()
    ex = D.ViewExchange(sc.views, sc.channels, sc.height, sc.width, full.dtype, dev) if world > 1 else None
    if ex is not None:
        ex.local_features().copy_(feats)

    def timed(call, n=steps):
        for _ in range(max(warmup, 3)):
            call()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(n):
            call()
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / n

    single = timed(lambda: cn.aggregate_views(proj, full, *args))
    out = {"config": workload_name(config, sc) + " -- Stage A only, ONE scene, views sharded over the ranks",
           "single_gpu_ms": single, "nvlink_gbs": NVLINK_GBS, "steps": steps, "modes": {}}
    if world == 1:
        return out
    g = world
    table = {
        "all_reduce": (lambda: D.aggregate_views_sharded(projs, feats, *args), 2.0 * (g - 1) / g * vol_bytes,
                       "one NCCL all-reduce of the packed [sums | counts] volume (north_star's form); output replicated"),
        "all_reduce_pipelined": (lambda: D.aggregate_views_sharded(projs, feats, *args, chunks=8),
                                 2.0 * (g - 1) / g * vol_bytes,
                                 "8 x-ranges, each range's all-reduce overlapped with the next range's gather kernel"),
        "reduce_scatter_pipelined": (lambda: D.aggregate_views_sharded(projs, feats, *args, chunks=2,
                                                                       collective="reduce_scatter"),
                                     (g - 1) / g * vol_bytes,
                                     "each x-range reduced to its owner only (ncclReduce), overlapped; output sharded by x"),
        "peer_store": (lambda: D.aggregate_views_p2p(projs, feats, *args), (g - 1) / g * vol_bytes,
                       "partial sums stored by the gather kernel straight into the owners' symmetric memory; output "
                       "sharded by voxel range"),
        "feature_exchange_exact": (lambda: ex.aggregate(proj, *args, overlap=False), None,
                                   "every rank owns a box and lifts ALL views: only the feature rows its voxels need "
                                   "are pulled from the peers (TMA puller), one launch in view order: bit-identical to "
                                   "one GPU; output sharded by box"),
        "feature_exchange_parts": (lambda: ex.aggregate(proj, *args, overlap=True), None,
                                   "same, the box served in 4 parts: the rows of the next part are pulled while the "
                                   "current part is gathered (still bit-identical); slower in practice: a GPU that is "
                                   "gathering saturates its own L2 and serves its peers' reads slowly"),
    }
    if os.environ.get("CNRMA_SUITE_EXPLORE"):      # tuning aid: part counts and puller grids of the exchange mode
        for parts, ctas, path in ((1, 148, "tma"), (1, 0, "lsu"), (4, 0, "tma"), (2, 0, "lsu"), (8, 0, "lsu")):
            table[f"explore_parts{parts}_ctas{ctas}_{path}"] = (
                (lambda parts=parts, ctas=ctas, path=path: ex.aggregate(proj, *args, overlap=parts > 1, parts=parts,
                                                                        pull_ctas=ctas, pull_path=path)),
                None, "exploration")
    for name, (call, wire, what) in table.items():
        if modes and name not in modes:
            continue
        ms = timed(call)
        if wire is None:
            t = torch.tensor([float(ex.pulled_bytes())], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wire = float(t.item())
        floor = wire / (NVLINK_GBS * 1e6)          # ms
        out["modes"][name] = {"ms": ms, "speedup_vs_1gpu": single / ms, "nvlink_bytes_per_rank": wire,
                              "nvlink_floor_ms": floor, "nvlink_roofline_frac": floor / ms, "what": what}
    # where the time of the exchange modes goes: device times of one call's phases on rank 0 (ms since the call began;
    # mark / pull run on the side stream, gather on the main one) and this rank's box gathered from local data alone
    prof = {}
    for name, overlap in (("feature_exchange_exact", False), ("feature_exchange_parts", True)):
        if name in out["modes"]:
            ph = {}
            ex.aggregate(proj, *args, overlap=overlap, profile=ph)
            prof[name] = {k: round(v, 4) for k, v in ph.items()}
    blo, bdim = D.box_shard(sc.voxel_dim, rank, world)
    box_ms = timed(lambda: cn.aggregate_views(proj, full, *args, box=(blo, bdim)), n=10)
    out["exchange_phases_ms"] = prof
    out["box_gather_alone_ms"] = box_ms
    # the ray-driven point form shards by view without any exchange of rows: every rank marches its own views and scales
    # its rows by the GLOBAL mean weight (two scalars all-reduced); rank order == view order
    tsdf = torch.from_numpy(sc.tsdf).to(dev)[None, None]
    proj_host = torch.from_numpy(sc.projections).unsqueeze(1)
    thr = 0.05
    pts_single = timed(lambda: cn.rma_points(proj_host, full, tsdf, *args, grids=sc.grids, threshold=thr), n=10)
    pts_sharded = timed(lambda: D.rma_points_sharded(proj_host[lo:hi], feats, tsdf, *args, grids=sc.grids, threshold=thr), n=10)
    out["points_sharded"] = {"single_gpu_ms": pts_single, "ms": pts_sharded, "speedup_vs_1gpu": pts_single / pts_sharded,
                             "what": "rma_points (march + fill of [M,3+C] rows) with the views sharded; one all-reduce of "
                                     "(sum of weights, M); rows stay on the rank that made them"}
    out["volume_bytes"] = vol_bytes
    out["feature_bytes"] = feat_bytes
    best = min(out["modes"].items(), key=lambda kv: kv[1]["ms"]) if out["modes"] else None
    if best:
        out["best"] = {"mode": best[0], "ms": best[1]["ms"], "speedup_vs_1gpu": best[1]["speedup_vs_1gpu"]}
    return out


def run_view_sharded(args, cn, dev, rank, world):
    """`--mode view-suite`: only the view-sharded record (profiles/r02_multi_gpu.md is made from these lines)."""
    import torch.distributed as dist
    rec = view_sharded_suite(cn, dev, rank, world, config=args.config if args.config != "cfg2" else "cfg4",
                             steps=min(args.steps, 30), warmup=args.warmup)
    if rank == 0:
        print(json.dumps({"view_sharded": rec, "n_gpus": world}))
    if world > 1:
        dist.destroy_process_group()


def host_copy_probe(dev, world, mb=512):
    """What the host lets this rank copy while every rank copies at once: pinned H2D alone, D2H alone and both together
    (GB/s, two streams).  The e2e leg moves ~1 GB in and ~7 GB out per scene and rank, so at N ranks it is bounded by
    these figures, not by the kernels (profiles/r02_multi_gpu.md holds the 1/2/4/8-rank table of this pool's box)."""
    import torch
    import torch.distributed as dist
    n = mb * 1024 * 1024
    h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in, d_out = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def run(h2d, d2h, reps=3):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s1.wait_event(a)
        s2.wait_event(a)
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        b.record()
        torch.cuda.synchronize()
        return (int(h2d) + int(d2h)) * reps * n / a.elapsed_time(b) / 1e6

    run(True, True, 1)
    return {"ranks_copying": world, "h2d_gbs": run(True, False), "d2h_gbs": run(False, True),
            "both_total_gbs": run(True, True), "what": "per-rank pinned-copy rates of rank 0 while all ranks copy"}


def run_e2e(args, cn, scenes, dev, world, max_rows, handoff_rows=None):
    """The same step through the stateful host mirror of the reference detector's aggregation methods
    (cn.RayMarchingAggregator) with HOST buffers: every step copies its scene's inputs from pinned host memory to the
    device and its results (volume, count, points) back to pinned host memory.  Three streams (copy-in, compute,
    copy-out); device inputs are double-buffered so that the next scene's upload and the previous scene's download overlap
    with the kernels and with each other; timed with CUDA events from the first upload to the last download, max over
    ranks.  With `handoff_rows` the lift is fused with the detector's point-cloud hand-off
    (rm.py:339-407): only the kept rows are produced and downloaded."""
    import torch
    import torch.distributed as dist
    sc = scenes[0]["sc"]
    V, C, H, W = sc.views, sc.channels, sc.height, sc.width
    nx, ny, nz = sc.voxel_dim
    hosts, outs, ok = [], [], 1
    try:
        for s_ in scenes[:2 if world < 4 else 1]:               # host scenes alternate (pinned, like a dataloader's)
            f = s_["feats"]
            h_feats = torch.empty(f.permute(0, 1, 3, 4, 2).shape, dtype=f.dtype).pin_memory()
            h_feats.copy_(f.permute(0, 1, 3, 4, 2))
            hosts.append(dict(feats=h_feats, proj=s_["proj_host"].pin_memory(), tsdf=s_["tsdf"].cpu().pin_memory()))
        # one pinned result slot: the downloads of consecutive steps are ordered on the copy-out stream, and nobody reads
        # the host copy in between (8 ranks x 8 GB of pinned rows is what the host can spare)
        outs = [dict(vol=torch.empty((1, nx, ny, nz, C), dtype=torch.float32).pin_memory(),
                     cnt=torch.empty((1, 1, nx, ny, nz), dtype=torch.int32).pin_memory(),
                     pts=torch.empty((handoff_rows or int(max_rows) + 4096, 3 + C), dtype=torch.float32).pin_memory())]
    except RuntimeError:
        ok = 0
    if world > 1:                                               # every rank or none: the leg has collectives
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = int(flag.item())
    if not ok:
        return {"value": None, "unit": "voxel*views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "error": "the host could not pin the buffers of the end-to-end leg on every rank"}
    ag = cn.RayMarchingAggregator(sc.voxel_size, sc.voxel_dim, origin=sc.origin.tolist(), backbone2d_stride=sc.stride,
                                  neus_threshold=args.threshold)
    s_in, s_cmp, s_out = (torch.cuda.Stream(device=dev) for _ in range(3))
    h0 = hosts[0]
    slots = [dict(feats=torch.empty_like(h0["feats"], device=dev), proj=torch.empty_like(h0["proj"], device=dev),
                  tsdf=torch.empty_like(h0["tsdf"], device=dev), loaded=torch.cuda.Event(), used=torch.cuda.Event())
             for _ in range(2)]

    def upload(k):
        sl, hs = slots[k % 2], hosts[k % len(hosts)]
        with torch.cuda.stream(s_in):
            s_in.wait_event(sl["used"])                      # the kernels that read this slot two steps ago are done
            sl["feats"].copy_(hs["feats"], non_blocking=True)
            sl["proj"].copy_(hs["proj"], non_blocking=True)
            sl["tsdf"].copy_(hs["tsdf"], non_blocking=True)
            sl["loaded"].record(s_in)

    def compute_and_download(k):
        sl, hs, out = slots[k % 2], hosts[k % len(hosts)], outs[0]
        with torch.cuda.stream(s_cmp):
            s_cmp.wait_event(sl["loaded"])
            d_feats = sl["feats"].permute(0, 1, 4, 2, 3)
            ag.initialize_volume()
            for v in range(V):
                ag.aggregate_2d_features(sl["proj"][v], d_feats[v])
            ag.clear_3d_features()
            if handoff_rows is None:
                ag.aggregate_2d_features_ray_marching(hs["proj"], d_feats, sl["tsdf"])   # cameras also known on the host
                pts = ag.points_detection[0]
            else:
                co, fe = cn.rma_points_selected(hs["proj"], d_feats, sl["tsdf"], sc.voxel_dim, sc.voxel_size, sc.origin,
                                                sc.stride, offsets=[[0.0, 0.0, 0.0]], max_points=handoff_rows,
                                                device_seed=1234 + k, grids=sc.grids, threshold=args.threshold)
                pts = co[0]._base if co[0]._base is not None else torch.cat((co[0], fe[0]), 1)
                pts = pts[: co[0].shape[0]]
            sl["used"].record(s_cmp)
            vol, cnt = ag.volume, ag._sum[1]
        with torch.cuda.stream(s_out):
            s_out.wait_stream(s_cmp)
            out["vol"].copy_(vol.permute(0, 2, 3, 4, 1), non_blocking=True)
            out["cnt"].copy_(cnt, non_blocking=True)
            out["pts"][: pts.shape[0]].copy_(pts, non_blocking=True)
            for t in (pts, vol, cnt):
                t.record_stream(s_out)
        return pts.shape[0]

    steps = max(2, min(args.steps, 6))
    # warm-up: three full steps, so that the caching allocator owns every block the pipeline cycles through (a result
    # block stays busy until its download is done; a cudaMalloc in the timed region would synchronise the device)
    upload(0)
    for k in range(3):
        if k < 2:
            upload(k + 1)
        compute_and_download(k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(s_in)
    upload(0)
    d2h = 0
    for k in range(steps):
        if k + 1 < steps:
            upload(k + 1)
        rows = compute_and_download(k)
        d2h += outs[0]["vol"].numel() * 4 + outs[0]["cnt"].numel() * 4 + rows * (3 + C) * 4
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
    h2d = h0["feats"].numel() * h0["feats"].element_size() + h0["proj"].numel() * 4 + h0["tsdf"].numel() * 4
    d2h //= steps
    return {"value": world * sc.voxel_views / (ms * 1e-3), "unit": "voxel*views/s", "ms_per_step": ms,
            "scenes_per_s": world / (ms * 1e-3), "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": steps,
            "pcie_gbs": (h2d + d2h) / ms / 1e6,
            "api": "RayMarchingAggregator (host mirror of the reference detector's aggregation methods)"
                   + ("" if handoff_rows is None else " + rma_points_selected (lift fused with switch_pointcloud)")
                   + "; copy-in / compute / copy-out on three streams"}


if __name__ == "__main__":
    main()
