"""Diagnostic: relative error of every Q-network output against the reference goldens, per math mode
(python tools/report_errors.py on a GPU box)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'golden')):
    sys.path.insert(0, p)
import torch, util, make_golden
from voxactb_b200 import QFunction, QFunction2Robots, VoxelGrid, _lib, synth
dev = torch.device('cuda')
def run(c, obs, enc, mode, two=False):
    enc.math_mode = mode
    vg = VoxelGrid(synth.SCENE_BOUNDS, c['V'], dev, c['B'], 3, c['cameras'] * c['H'] * c['W'])
    rgb = [t.cuda() for t in obs['rgb']]; pcd = [t.cuda() for t in obs['pcd']]
    rp = [[r, p] for r, p in zip(rgb, pcd)]
    if two:
        q = QFunction2Robots(enc, vg, 0.15, 5, dev, False).to(dev).eval()
        o = q(rp, obs['proprio'].cuda(), obs['proprio_left'].cuda(), pcd, obs['lang_goal_emb'].cuda(), obs['lang_token_embs'].cuda(), obs['bounds'].cuda(), None, None)
        return dict(trans=o[0], rot_grip=o[1], collision=o[2], trans_left=o[4], rot_grip_left=o[5], collision_left=o[6])
    q = QFunction(enc, vg, 0.15, 5, dev, False, c['arm']).to(dev).eval()
    o = q(rp, obs['proprio'].cuda(), pcd, obs['lang_goal_emb'].cuda(), obs['lang_token_embs'].cuda(), obs['bounds'].cuda(), None, None)
    return dict(trans=o[0], rot_grip=o[1], collision=o[2])
for name in ['qnet_v20', 'qnet_v20_arm_crop', 'qnet_v32_config1']:
    c = make_golden.QNET_CASES[name]; g = util.golden(name)
    for mode in (0, 1, 2):
        for rep in range(2):
            obs, enc, sd = util.make_case(c)
            o = run(c, obs, enc, mode)
            print(name, 'mode', mode, {k: '%.2e' % util.rel_err(v, g[k]) for k, v in o.items()})
c = make_golden.QNET2_CASES['qnet2_v20']; g = util.golden('qnet2_v20')
for mode in (0, 1, 2):
    for rep in range(3):
        obs, enc, sd = util.make_case_two_robots(c)
        o = run(c, obs, enc, mode, True)
        print('qnet2_v20 mode', mode, {k: '%.2e' % util.rel_err(v, g[k]) for k, v in o.items()})
