"""The CPU oracle against (a) the reference's own modules, imported live when /root/reference is
mounted, and (b) the committed golden fixtures generated from them (always)."""
import numpy as np
import pytest
import torch

import refimport
import util
from oracle import qnet_oracle, train_oracle, voxel_oracle
from voxactb_b200 import synth

import make_golden

needs_ref = pytest.mark.skipif(not refimport.available(), reason='/root/reference not mounted')


@pytest.mark.parametrize('name', ['voxel_v20', 'voxel_v32_crop', 'voxel_v100'])
def test_voxel_oracle_matches_golden(name):
    c = make_golden.VOXEL_CASES[name]
    g = util.golden(name)
    obs = synth.make_observation(c['seed'], c['B'], c['cameras'], c['H'], c['W'], per_sample_crop=c['crop'])
    coords, feats = synth.flatten_cameras(obs)
    assert np.allclose([util.checksum(coords), util.checksum(feats), util.checksum(obs['bounds'])],
                       g['in_checksum'], rtol=1e-12), 'synthetic input drifted from the golden run'
    idx = voxel_oracle.voxel_indices(coords.numpy(), obs['bounds'].numpy(), c['V'])
    grid = voxel_oracle.voxelize(coords.numpy(), feats.numpy(), obs['bounds'].numpy(), c['V'])
    assert int(idx.astype(np.int64).sum()) == int(g['idx_checksum'][0])
    assert int((grid[..., -1] > 0).sum()) == int(g['occupied'][0])
    if 'grid' in g.files:
        assert np.array_equal(idx, g['idx'].astype(np.int32))       # bit-exact indices
        assert np.array_equal(grid, g['grid'])                       # same sequential fp32 sums
    else:
        assert np.array_equal(idx[:, :4096], g['idx_head'].astype(np.int32))
        flat = grid.reshape(-1, grid.shape[-1])
        assert np.array_equal(flat[g['sample_pos']], g['sample_val'])


@pytest.mark.parametrize('name', ['qnet_v20', 'qnet_v20_arm_crop', 'qnet_v32_config1'])
def test_qnet_oracle_matches_golden(name):
    c = make_golden.QNET_CASES[name]
    g = util.golden(name)
    obs, enc, sd = util.make_case(c)
    coords, feats = synth.flatten_cameras(obs)
    assert np.allclose([util.checksum(coords), util.checksum(feats), util.checksum(obs['proprio']),
                        util.checksum(obs['lang_token_embs']), util.checksum(obs['bounds'])],
                       g['in_checksum'], rtol=1e-12)
    assert np.allclose(sum(util.checksum(v) for v in sd.values()), g['sd_checksum'], rtol=1e-12)
    out = qnet_oracle.qfunction_forward(sd, util.oracle_cfg(c), voxel_oracle.voxelize, obs['rgb'], obs['pcd'],
                                        obs['proprio'], obs['lang_token_embs'], obs['bounds'], c['V'])
    assert util.rel_err(out['trans'], g['trans']) < 1e-5
    assert util.rel_err(out['rot_grip'], g['rot_grip']) < 1e-5
    assert util.rel_err(out['collision'], g['collision']) < 1e-5
    if c['arm']:
        assert util.rel_err(out['arm'], g['arm']) < 1e-5
    assert np.array_equal(out['trans'].reshape(c['B'], -1).argmax(-1).numpy(), g['trans_argmax'])


def test_two_robot_oracle_matches_golden_and_keys():
    """PerceiverVoxelLang2RobotsEncoder: oracle vs the reference-generated fixture; checkpoint keys identical."""
    name = 'qnet2_v20'
    c = make_golden.QNET2_CASES[name]
    g = util.golden(name)
    obs, enc, sd = util.make_case_two_robots(c)
    assert sorted(enc.state_dict().keys()) == list(g['keys'])
    cfg = dict(util.oracle_cfg(c), two_robots=True, proprio_left=obs['proprio_left'])
    out = qnet_oracle.qfunction_forward(sd, cfg, voxel_oracle.voxelize, obs['rgb'], obs['pcd'], obs['proprio'],
                                        obs['lang_token_embs'], obs['bounds'], c['V'])
    for k in ('trans', 'rot_grip', 'collision', 'trans_left', 'rot_grip_left', 'collision_left'):
        assert util.rel_err(out[k], g[k]) < 1e-5, k


@pytest.mark.parametrize('name', ['train_v20', 'train_v20_arm'])
def test_training_step_oracle_matches_reference_golden(name):
    """Row a18: losses, EVERY parameter gradient and the LAMB-updated parameters of one `update` step, against the
    fixture the reference produced (its modules in train mode with zero dropout, torch autograd, its Lamb class)."""
    c = make_golden.TRAIN_CASES[name]
    g = util.golden(name)
    obs, enc, sd = util.make_case(c)
    lab = make_golden.train_labels(c)
    assert int(sum(int(v.long().sum()) for v in lab.values())) == int(g['label_checksum'][0])
    sd = {k: v for k, v in sd.items() if k in set(g['keys'].tolist())}        # parameters only (no pos_x/y/z buffers)
    res = train_oracle.training_step(sd, util.oracle_cfg(c), obs['rgb'], obs['pcd'], obs['proprio'], obs['lang_token_embs'],
                                     obs['bounds'], c['V'], lab)
    assert abs(float(res['total']) - float(g['total'][0])) <= 2e-5 * abs(float(g['total'][0]))
    for t in ('trans', 'rot', 'grip', 'collision') + (('arm',) if c['arm'] else ()):
        np.testing.assert_allclose(res['terms'][t].numpy(), g['loss_' + t], rtol=2e-5, atol=2e-5)
    keys = g['keys'].tolist()
    assert sorted(res['grads'].keys()) == keys and len(g['no_grad_keys']) == 0
    for i, k in enumerate(keys):
        gr = res['grads'][k]
        scale = max(float(g['grad_max'][i]), 1e-12)
        flat = gr.reshape(-1)
        samp = flat if flat.numel() <= 4096 else flat[:: max(1, flat.numel() // 2048)]
        assert float((samp - torch.from_numpy(g['g:' + k])).abs().max()) <= 2e-4 * scale, k
        assert abs(float(gr.double().abs().sum()) - float(g['grad_abs'][i])) <= 2e-4 * max(float(g['grad_abs'][i]), 1e-12), k
        # LAMB step from zero state: the update magnitude |p_new - p| (trust-ratio scaled) and the new parameter sum
        delta = float((res['params'][k].double() - sd[k].double()).abs().sum())
        assert abs(delta - float(g['param_delta_abs'][i])) <= 2e-3 * max(float(g['param_delta_abs'][i]), 1e-12), k
        assert abs(float(res['params'][k].double().sum()) - float(g['param_sum'][i])) <= 1e-4 * max(1.0, abs(float(g['param_sum'][i]))), k


def test_state_dict_keys_match_reference():
    """Checkpoint compatibility: our module exposes exactly the reference's state_dict keys."""
    for name in ('qnet_v20', 'qnet_v20_arm_crop'):
        c = make_golden.QNET_CASES[name]
        _, enc, _ = util.make_case(c)
        assert sorted(enc.state_dict().keys()) == list(util.golden(name)['keys'])


@needs_ref
def test_weight_tied_encoder_matches_the_reference_class_and_the_oracle():
    """weight_tie_layers=True (perceiver_lang_io.py:263-276, cache_fn): same state-dict keys as the reference class, every
    layers.<i> entry aliasing layers.0; the oracle run on that state dict equals the reference forward."""
    _, RefEnc = refimport.load()
    from voxactb_b200 import PerceiverVoxelLangEncoder
    c = dict(make_golden.QNET_CASES['qnet_v20'], depth=3)
    kw = dict(make_golden.encoder_kwargs(c), weight_tie_layers=True)
    ref, ours = RefEnc(**kw).eval(), PerceiverVoxelLangEncoder(**kw).eval()
    assert sorted(ref.state_dict().keys()) == sorted(ours.state_dict().keys())
    sd = synth.random_state_dict(ref, 91)
    ref.load_state_dict(sd, strict=False)
    ours.load_state_dict(sd, strict=False)
    got = ours.state_dict()
    for k, v in ref.state_dict().items():
        if not k.endswith(('pos_x', 'pos_y', 'pos_z')):
            assert torch.equal(v, got[k]), k
    assert got['layers.2.1.fn.net.0.weight'].data_ptr() == got['layers.0.1.fn.net.0.weight'].data_ptr()
    obs = synth.make_observation(c['seed'], c['B'], c['cameras'], c['H'], c['W'], low_dim=c['low_dim'])
    coords, feats = synth.flatten_cameras(obs)
    grid = torch.from_numpy(voxel_oracle.voxelize(coords.numpy(), feats.numpy(), obs['bounds'].numpy(), c['V'])).permute(0, 4, 1, 2, 3)
    with torch.no_grad():
        want = ref(grid, obs['proprio'], obs['lang_goal_emb'], obs['lang_token_embs'], None, obs['bounds'], None)
        have = qnet_oracle.qnet_forward(dict(ref.state_dict()), util.oracle_cfg(c), grid, obs['proprio'], obs['lang_token_embs'])
    for w, k in zip(want[:3], ('trans', 'rot_grip', 'collision')):
        assert util.rel_err(have[k], w) < 2e-5, k


@needs_ref
def test_pos_encoding_without_language_matches_the_reference_class():
    """pos_encoding_with_lang=False (perceiver_lang_io.py:192-194, 392-393): parameter shape [1,S,S,S,C] as in the reference,
    oracle forward == reference forward."""
    _, RefEnc = refimport.load()
    from voxactb_b200 import PerceiverVoxelLangEncoder
    c = make_golden.QNET_CASES['qnet_v20']
    kw = dict(make_golden.encoder_kwargs(c), pos_encoding_with_lang=False)
    ref, ours = RefEnc(**kw).eval(), PerceiverVoxelLangEncoder(**kw).eval()
    assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    sd = synth.random_state_dict(ref, 93)
    ref.load_state_dict(sd, strict=False)
    obs = synth.make_observation(c['seed'], c['B'], c['cameras'], c['H'], c['W'], low_dim=c['low_dim'])
    coords, feats = synth.flatten_cameras(obs)
    grid = torch.from_numpy(voxel_oracle.voxelize(coords.numpy(), feats.numpy(), obs['bounds'].numpy(), c['V'])).permute(0, 4, 1, 2, 3)
    with torch.no_grad():
        want = ref(grid, obs['proprio'], obs['lang_goal_emb'], obs['lang_token_embs'], None, obs['bounds'], None)
        have = qnet_oracle.qnet_forward(dict(ref.state_dict()), util.oracle_cfg(c), grid, obs['proprio'], obs['lang_token_embs'])
    for w, k in zip(want[:3], ('trans', 'rot_grip', 'collision')):
        assert util.rel_err(have[k], w) < 2e-5, k


@needs_ref
@pytest.mark.parametrize('flag', ['no_skip_connection', 'no_perceiver'])
def test_final_conv_ablations_match_the_reference_class(flag):
    """no_skip_connection / no_perceiver (perceiver_lang_io.py:296-306, 456-462): final convolution over 64 channels (u0 or d0
    alone); same parameter shapes as the reference, oracle forward == reference forward."""
    _, RefEnc = refimport.load()
    from voxactb_b200 import PerceiverVoxelLangEncoder
    c = dict(make_golden.QNET_CASES['qnet_v20'], **{flag: True})
    kw = dict(make_golden.encoder_kwargs(c), **{flag: True})
    ref, ours = RefEnc(**kw).eval(), PerceiverVoxelLangEncoder(**kw).eval()
    assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    assert tuple(ours.state_dict()['final.conv3d.weight'].shape) == (64, 64, 3, 3, 3)
    sd = synth.random_state_dict(ref, 95)
    ref.load_state_dict(sd, strict=False)
    obs = synth.make_observation(c['seed'], c['B'], c['cameras'], c['H'], c['W'], low_dim=c['low_dim'])
    coords, feats = synth.flatten_cameras(obs)
    grid = torch.from_numpy(voxel_oracle.voxelize(coords.numpy(), feats.numpy(), obs['bounds'].numpy(), c['V'])).permute(0, 4, 1, 2, 3)
    with torch.no_grad():
        want = ref(grid, obs['proprio'], obs['lang_goal_emb'], obs['lang_token_embs'], None, obs['bounds'], None)
        have = qnet_oracle.qnet_forward(dict(ref.state_dict()), util.oracle_cfg(c), grid, obs['proprio'], obs['lang_token_embs'])
    for w, k in zip(want[:3], ('trans', 'rot_grip', 'collision')):
        assert util.rel_err(have[k], w) < 2e-5, k


@needs_ref
def test_voxel_oracle_vs_live_reference():
    RefVG, _ = refimport.load()
    for seed, V, crop in ((5, 16, False), (6, 24, True)):
        obs = synth.make_observation(seed, 2, 2, 24, 24, per_sample_crop=crop)
        coords, feats = synth.flatten_cameras(obs)
        ref = RefVG(synth.SCENE_BOUNDS, V, 'cpu', 2, 3, coords.shape[1]).coords_to_bounding_voxel_grid(
            coords, feats, obs['bounds'])
        mine = voxel_oracle.voxelize(coords.numpy(), feats.numpy(), obs['bounds'].numpy(), V)
        assert np.array_equal(ref.numpy(), mine)


@needs_ref
def test_qnet_oracle_and_signature_vs_live_reference():
    import inspect
    from voxactb_b200 import PerceiverVoxelLangEncoder, VoxelGrid
    RefVG, RefEnc = refimport.load()
    assert inspect.signature(RefEnc.__init__) == inspect.signature(PerceiverVoxelLangEncoder.__init__)
    assert list(inspect.signature(RefEnc.forward).parameters) == \
        list(inspect.signature(PerceiverVoxelLangEncoder.forward).parameters)
    assert list(inspect.signature(RefVG.__init__).parameters) == list(inspect.signature(VoxelGrid.__init__).parameters)
    c = dict(V=16, k=5, s=4, L=48, depth=1, B=2, cameras=1, H=16, W=16, low_dim=7, arm=True, crop=False, seed=77)
    obs, enc, sd = util.make_case(c)
    net = RefEnc(**make_golden.encoder_kwargs(c)).eval()
    net.load_state_dict(sd, strict=False)
    coords, feats = synth.flatten_cameras(obs)
    grid = torch.from_numpy(voxel_oracle.voxelize(coords.numpy(), feats.numpy(), obs['bounds'].numpy(), 16)).permute(0, 4, 1, 2, 3)
    with torch.no_grad():
        ref = net(grid, obs['proprio'], None, obs['lang_token_embs'], None, obs['bounds'], None)
    out = qnet_oracle.qnet_forward(sd, util.oracle_cfg(c), grid, obs['proprio'], obs['lang_token_embs'])
    for a, b in zip(ref, (out['trans'], out['rot_grip'], out['collision'], out['arm'])):
        assert util.rel_err(b, a) < 1e-6
