"""Stage A time over the sweep knobs (slab thickness, (x, y) tile, work-unit shape) in ONE process: the knobs are flipped with
os.environ + cn.reload_tuning() between timings, every result is compared bit for bit with the first.  Run under gpurun."""
import os, sys
sys.path.insert(0, ".")
import torch, cnrma_b200 as cn
def run(cfg, combos, iters=10):
    sc = cn.synthetic.make_scene(cfg, seed=0, with_features=False)
    f = cn.synthetic.device_features(sc, "cuda")
    p = torch.from_numpy(sc.projections).cuda().unsqueeze(1)
    args = (sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride)
    ref = None
    for slab, tile, cull in combos:
        for k, v in (("CNRMA_AGG_SLAB", slab), ("CNRMA_AGG_TILE", tile), ("CNRMA_AGG_CULL", cull)):
            if v is None: os.environ.pop(k, None)
            else: os.environ[k] = str(v)
        cn.reload_tuning()
        for _ in range(3): vol, cnt, _v = cn.aggregate_views(p, f, *args)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters): vol, cnt, _v = cn.aggregate_views(p, f, *args)
        b.record(); torch.cuda.synchronize()
        if ref is None: ref = vol.clone()
        same = bool(torch.equal(vol, ref))
        print(f"{cfg} slab={slab} tile={tile} cull={cull}: {a.elapsed_time(b) / iters:.4f} ms  identical={same}", flush=True)
combos4 = [(None, None, None)] + [(s, t, c) for c in (1, 0) for s in (4, 8, 16, 32, 64) for t in (None, 16, 32, 80)]
run("cfg4", combos4)
combos3 = [(None, None, None)] + [(s, t, 0) for s in (5, 8, 20, 40) for t in (None, 16, 32)]
run("cfg3", combos3, iters=5)
