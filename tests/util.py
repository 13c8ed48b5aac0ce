"""Shared helpers for the parity tests."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, 'golden')
sys.path.insert(0, GOLDEN)

# north_star gate: Q-values within 1e-3 relative (fp32).  Measured as the largest absolute
# deviation normalised by the largest reference magnitude of the same tensor.
Q_REL_TOL = 1e-3


def rel_err(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def frac_outside(a, b, rtol=1e-3, floor=1e-2):
    """Stricter, elementwise reading of "within 1e-3 relative": the fraction of elements with
    |a - b| > rtol * (|b| + floor * max|b|) (the floor keeps the ratio defined near zero crossings)."""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    tol = rtol * (b.abs() + floor * b.abs().max())
    return float(((a - b).abs() > tol).double().mean())


def golden(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)


def checksum(t):
    return float(torch.as_tensor(t).double().abs().sum())


def oracle_cfg(c):
    return dict(voxel_patch_size=c['k'], voxel_patch_stride=c['s'], depth=c['depth'], iterations=1,
                cross_heads=1, latent_heads=8, activation='lrelu', num_collision_classes=2,
                arm_pred_loss=c['arm'], no_language=False, no_skip_connection=c.get('no_skip_connection', False),
                no_perceiver=c.get('no_perceiver', False))


def make_case(c):
    """Inputs + seeded state dict for a QNET_CASES entry (same recipe as make_golden.py)."""
    from voxactb_b200 import synth, PerceiverVoxelLangEncoder
    import make_golden
    obs = synth.make_observation(c['seed'], c['B'], c['cameras'], c['H'], c['W'], low_dim=c['low_dim'],
                                 per_sample_crop=c['crop'])
    enc = PerceiverVoxelLangEncoder(**make_golden.encoder_kwargs(c)).eval()
    sd = synth.random_state_dict(enc, make_golden.weight_seed(c))
    missing = enc.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys
    assert all(k.endswith(('pos_x', 'pos_y', 'pos_z')) for k in missing.missing_keys)
    return obs, enc, sd


def make_case_two_robots(c):
    """Inputs + seeded state dict for a QNET2_CASES entry (2-robot encoder)."""
    from voxactb_b200 import synth, PerceiverVoxelLang2RobotsEncoder
    import make_golden
    obs = synth.make_observation(c['seed'], c['B'], c['cameras'], c['H'], c['W'], low_dim=c['low_dim'])
    obs['proprio_left'] = make_golden.proprio_left(c)
    kw = make_golden.encoder_kwargs(c)
    kw.pop('arm_pred_loss')
    enc = PerceiverVoxelLang2RobotsEncoder(**kw).eval()
    sd = synth.random_state_dict(enc, make_golden.weight_seed(c))
    missing = enc.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys
    assert all(k.endswith(('pos_x', 'pos_y', 'pos_z')) for k in missing.missing_keys)
    return obs, enc, sd
