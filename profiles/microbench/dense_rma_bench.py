import sys, torch, time
sys.path.insert(0, '/root/repo')
import cnrma_b200 as cn
sc = cn.synthetic.make_scene('cfg2', seed=0, with_features=False)
dev = torch.device('cuda')
feats = cn.synthetic.device_features(sc, dev, channels_last=True)
proj = torch.from_numpy(sc.projections).to(dev).unsqueeze(1)
tsdf = torch.from_numpy(sc.tsdf).to(dev)[None, None]
for _ in range(3):
    wsum, wtot = cn.dense_rma(proj, feats, tsdf, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride, threshold=0.05)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    wsum, wtot = cn.dense_rma(proj, feats, tsdf, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride, threshold=0.05)
e1.record(); torch.cuda.synchronize()
print('dense_rma ms', e0.elapsed_time(e1)/10, 'occupied voxels', int((wtot>0).sum()), 'of', sc.nvox)
