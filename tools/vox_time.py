import sys, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxactb_b200 import VoxelGrid, synth
B=16
dev=torch.device('cuda')
obs = synth.make_observation(1234, B, 4, 128, 128)
coords, feats = synth.flatten_cameras(obs)
vg = VoxelGrid(synth.SCENE_BOUNDS, 100, dev, B, 3, coords.shape[1])
bnd = obs['bounds'].cuda()
def t(c, f, n=200):
    for _ in range(5): vg.coords_to_bounding_voxel_grid(c, f, bnd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): vg.coords_to_bounding_voxel_grid(c, f, bnd)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n*1000
c, f = coords.cuda(), feats.cuda()
print('normal us', t(c, f))
print('all outside us', t(c + 100.0, f))
half = c.clone(); half[:, ::2] += 100.0
print('half outside us', t(half, f))
one = torch.zeros_like(c) + torch.tensor([0.2, 0.0, 1.0], device=dev)
print('all in one voxel us', t(one, f))
