"""Diagnostic at the headline geometry (100^3, 4 cameras, 2048 latents, depth 6): error of the two tensor-core math modes against
the library's own fp32 FFMA mode on fresh seeds (python tools/mode_agreement_v100.py on a GPU box).  The CPU oracle is too slow for
more than the committed V=100 goldens; this widens the sample."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'golden')):
    sys.path.insert(0, p)
import torch
import util, make_golden
from voxactb_b200 import QFunction, VoxelGrid, _lib, synth

dev = torch.device('cuda')


def run(c, obs, enc, mode):
    enc.math_mode = mode
    vg = VoxelGrid(synth.SCENE_BOUNDS, c['V'], dev, c['B'], 3, c['cameras'] * c['H'] * c['W'])
    q = QFunction(enc, vg, 0.15, 5, dev, False, c['arm']).to(dev).eval()
    rgb = [t.cuda() for t in obs['rgb']]; pcd = [t.cuda() for t in obs['pcd']]
    o = q([[r, p] for r, p in zip(rgb, pcd)], obs['proprio'].cuda(), pcd, obs['lang_goal_emb'].cuda(), obs['lang_token_embs'].cuda(),
          obs['bounds'].cuda(), None, None)
    torch.cuda.synchronize()
    return [t.clone() for t in o[:3]]


worst = {1: [0, 0, 0], 2: [0, 0, 0]}
for seed in range(6):
    c = dict(make_golden.QNET_CASES['qnet_v100_b1'], B=2, seed=1000 + seed, crop=bool(seed & 1))
    obs, enc, sd = util.make_case(c)
    ref = run(c, obs, enc, _lib.MATH_FP32_SIMT)
    for mode in (_lib.MATH_F16X3, _lib.MATH_F16F8C):
        out = run(c, obs, enc, mode)
        errs = [util.rel_err(a, b) for a, b in zip(out, ref)]
        same = bool(torch.equal(out[0].reshape(2, -1).argmax(-1), ref[0].reshape(2, -1).argmax(-1)))
        worst[mode] = [max(w, e) for w, e in zip(worst[mode], errs)]
        print('seed %d mode %d  trans %.2e  rot_grip %.2e  collision %.2e  argmax voxel equal %s' % (seed, mode, *errs, same), flush=True)
for mode, w in worst.items():
    print('WORST mode %d  trans %.2e  rot_grip %.2e  collision %.2e' % (mode, *w))
