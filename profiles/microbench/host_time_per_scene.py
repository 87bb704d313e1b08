"""Host time per scene of the public calls (aggregate_views + rma_points) against the GPU time of the same loop: where the host
thread is busy, where it waits for M (the march result), and whether anything is left between steps.  Run under gpurun."""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch, cnrma_b200 as cn
from cnrma_b200 import functional as F
dev = torch.device("cuda", 0)
scenes = []
for i in range(4):
    sc = cn.synthetic.make_scene("cfg2", seed=i, with_features=False)   # fixed room (bench.py rotates "room_var" scenes: more rows)
    feats = cn.synthetic.device_features(sc, dev, channels_last=True)
    ph = torch.from_numpy(sc.projections).unsqueeze(1)
    scenes.append(dict(sc=sc, feats=feats, proj_host=ph, proj=ph.to(dev), tsdf=torch.from_numpy(sc.tsdf).to(dev)[None, None]))
sc = scenes[0]["sc"]; ga = (sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride)
T = {"rr_in": [], "rr_out": []}
orig = F._read_result
def rr(m):
    T["rr_in"].append(time.perf_counter()); r = orig(m); T["rr_out"].append(time.perf_counter()); return r
F._read_result = rr
def run(n, tag):
    T["rr_in"].clear(); T["rr_out"].clear()
    t_a0, t_a1, t_b1 = [], [], []
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
    for k in range(n):
        s_ = scenes[k % 4]
        t_a0.append(time.perf_counter())
        cn.aggregate_views(s_["proj"], s_["feats"], *ga, mean=True)
        t_a1.append(time.perf_counter())
        rows = cn.rma_points(s_["proj_host"], s_["feats"], s_["tsdf"], *ga, grids=sc.grids, threshold=0.05)[0]
        t_b1.append(time.perf_counter())
    e1.record(); torch.cuda.synchronize()
    a0, a1, b1, ri, ro = map(np.array, (t_a0, t_a1, t_b1, T["rr_in"], T["rr_out"]))
    print(f"{tag}: gpu ms/step {e0.elapsed_time(e1)/n:.3f} | host: aggregate_views {1e3*(a1-a0)[5:].mean():.3f} ms, rma_points until the wait for M "
          f"{1e3*(ri-a1)[5:].mean():.3f} ms, waiting for M {1e3*(ro-ri)[5:].mean():.3f} ms, after M {1e3*(b1-ro)[5:].mean():.3f} ms, "
          f"between steps {1e3*(a0[1:]-b1[:-1])[5:].mean():.3f} ms; host busy per step {1e3*((a1-a0)+(ri-a1)+(b1-ro))[5:].mean():.3f} ms", flush=True)
run(10, "warm"); run(100, "default threads")
torch.set_num_threads(1); run(100, "torch.set_num_threads(1)")
