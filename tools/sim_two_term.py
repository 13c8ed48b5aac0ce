"""CPU experiment: what a 2-term split-fp16 product would cost in accuracy, per layer.
Runs the fp32 oracle with one operand of chosen conv layers rounded to fp16 (11 significant bits) -- the effect of dropping
the A_lo.W_hi term (activations rounded) or the A_hi.W_lo term (weights rounded) -- and reports the change of every output
relative to the unrounded oracle, with the tests' metric (max|a-b| / max|b|)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'golden')):
    sys.path.insert(0, p)
import torch
import torch.nn.functional as F
import util, make_golden
from oracle import qnet_oracle, voxel_oracle

MODE = {}
orig_conv = F.conv3d


def h(x):
    return x.half().float()


def conv_patched(x, w, b=None, stride=1, **kw):
    tag = None
    if w.shape[1] == 128 and w.shape[-1] == 3:
        tag = 'final'
    elif w.shape[-1] == 5 and w.shape[0] == 64 and w.shape[1] == 64 and stride == 1 and x.shape[-1] > 25:
        tag = 'upconv'
    m = MODE.get(tag)
    if m == 'A':
        x = h(x)
    elif m == 'W':
        w = h(w)
    return orig_conv(x, w, b, stride=stride, **kw)


F.conv3d = conv_patched
names = sys.argv[1:] or ['qnet_v20', 'qnet_v20_arm_crop', 'qnet_v32_config1']
for name in names:
    c = make_golden.QNET_CASES[name]
    obs, enc, sd = util.make_case(c)
    def run():
        return qnet_oracle.qfunction_forward(sd, util.oracle_cfg(c), voxel_oracle.voxelize, obs['rgb'], obs['pcd'], obs['proprio'],
                                             obs['lang_token_embs'], obs['bounds'], c['V'])
    MODE.clear()
    ref = run()
    for mode in ({'final': 'A'}, {'final': 'W'}, {'upconv': 'A'}, {'upconv': 'W'}, {'final': 'A', 'upconv': 'A'}, {'final': 'W', 'upconv': 'W'}):
        MODE.clear(); MODE.update(mode)
        o = run()
        print(name, mode, {k: '%.2e' % util.rel_err(o[k], ref[k]) for k in ('trans', 'rot_grip', 'collision')},
              'argmax same', bool((o['trans'].flatten(1).argmax(1) == ref['trans'].flatten(1).argmax(1)).all()), flush=True)
