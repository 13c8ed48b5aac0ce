"""Row f4: the act() tail on the device and the per-episode language cache."""
import importlib.util
import os
from unittest import mock
import sys

import numpy as np
import pytest
import torch

import util
from oracle import act_oracle, qnet_oracle, voxel_oracle

import make_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UTILS = next((p for p in (os.path.join(ROOT, 'baseline', '_ref', 'peract', 'helpers', 'utils.py'),
                          '/root/reference/peract/helpers/utils.py') if os.path.exists(p)), None)


@pytest.mark.skipif(UTILS is None, reason='reference helpers/utils.py not available')
def test_act_oracle_matches_reference_helper():
    for name in ('pyrender', 'pyrender.trackball', 'trimesh', 'rlbench', 'rlbench.backend', 'rlbench.backend.const',
                 'rlbench.backend.observation_two_robots', 'pyrep', 'pyrep.const'):
        sys.modules.setdefault(name, mock.MagicMock())
    spec = importlib.util.spec_from_file_location('ref_helpers_utils', UTILS)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(0)
    for _ in range(200):
        d = rng.integers(0, 72, 3)
        assert np.array_equal(act_oracle.discrete_euler_to_quaternion(d, 5), ref.discrete_euler_to_quaternion(d, 5))


def test_cached_language_encoder_encodes_once_per_instruction():
    from voxactb_b200.act import CachedLanguageEncoder
    calls = []

    def enc(tok):
        calls.append(1)
        return torch.ones(1, 1024) * float(tok.sum()), torch.ones(1, 77, 512)
    c = CachedLanguageEncoder(enc)
    a = torch.arange(77)[None]
    for _ in range(5):
        e, t = c(a)
    c(a + 1)
    assert len(calls) == 2 and c.hits == 4 and float(e[0, 0]) == float(a.sum())


@pytest.mark.gpu
def test_act_tail_all_bins(cuda_lib):
    from voxactb_b200 import _lib
    rng = np.random.default_rng(1)
    B = 4096
    rg = np.concatenate([rng.integers(0, 72, (B, 3)), rng.integers(0, 2, (B, 1))], 1).astype(np.int32)
    rg[:72, 0] = np.arange(72); rg[72:144, 1] = np.arange(72); rg[144:216, 2] = np.arange(72)
    coll = rng.integers(0, 2, B).astype(np.int32)
    xyz = rng.normal(size=(B, 3)).astype(np.float32)
    out = torch.empty(B, 9, device='cuda')
    rg_d, coll_d, xyz_d = torch.from_numpy(rg).cuda(), torch.from_numpy(coll).cuda(), torch.from_numpy(xyz).cuda()
    rc = cuda_lib.vxb_act_tail_f32(_lib.ptr(rg_d), _lib.ptr(coll_d), _lib.ptr(xyz_d), 5.0, _lib.ptr(out), B, _lib.stream())
    _lib.check(rc, 'vxb_act_tail_f32')
    ref = act_oracle.continuous_action(xyz, rg, coll, 5)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=0, atol=1e-6)


@pytest.mark.gpu
def test_fused_actor_matches_oracle_pipeline(cuda_lib):
    from voxactb_b200 import QFunction, VoxelGrid, synth
    from voxactb_b200.act import FusedActor
    c = make_golden.QNET_CASES['qnet_v20']
    obs, enc, sd = util.make_case(c)
    dev = torch.device('cuda')
    vg = VoxelGrid(synth.SCENE_BOUNDS, c['V'], dev, c['B'], 3, c['cameras'] * c['H'] * c['W'])
    q = QFunction(enc, vg, 0.15, 5, dev, False, False).to(dev).eval()
    rgb = [t.cuda() for t in obs['rgb']]
    pcd = [t.cuda() for t in obs['pcd']]
    action, extra = FusedActor(q, 5).act([[r, p] for r, p in zip(rgb, pcd)], obs['proprio'].cuda(), pcd,
                                         obs['lang_goal_emb'].cuda(), obs['lang_token_embs'].cuda(), obs['bounds'].cuda())
    ref = qnet_oracle.qfunction_forward(sd, util.oracle_cfg(c), voxel_oracle.voxelize, obs['rgb'], obs['pcd'], obs['proprio'],
                                        obs['lang_token_embs'], obs['bounds'], c['V'])
    coords, rg, ic = qnet_oracle.choose_highest_action(ref['trans'], ref['rot_grip'], ref['collision'])
    xyz = qnet_oracle.attention_coordinate(coords, obs['bounds'].expand(c['B'], 6), c['V']).numpy()
    expect = act_oracle.continuous_action(xyz, rg.numpy(), ic.numpy(), 5)
    np.testing.assert_allclose(action.numpy(), expect, rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
def test_graphed_actor_replays_one_graph_for_new_observations(cuda_lib):
    """The CUDA-graph actor (batch-1 closed loop): one capture, then every replay on NEW observations matches the oracle
    pipeline; a parameter update forces a re-capture."""
    from voxactb_b200 import QFunction, VoxelGrid, synth
    from voxactb_b200.act import GraphedActor
    c = dict(make_golden.QNET_CASES['qnet_v20'], B=1)
    obs, enc, sd = util.make_case(c)
    dev = torch.device('cuda')
    vg = VoxelGrid(synth.SCENE_BOUNDS, c['V'], dev, c['B'], 3, c['cameras'] * c['H'] * c['W'])
    q = QFunction(enc, vg, 0.15, 5, dev, False, False).to(dev).eval()
    actor = GraphedActor(q, 5)
    for seed in (c['seed'], c['seed'] + 1, c['seed'] + 2):
        o = synth.make_observation(seed, 1, c['cameras'], c['H'], c['W'], low_dim=c['low_dim'], per_sample_crop=c['crop'])
        rgb = [t.cuda() for t in o['rgb']]
        pcd = [t.cuda() for t in o['pcd']]
        action, extra = actor.act([[r, p] for r, p in zip(rgb, pcd)], o['proprio'].cuda(), pcd, o['lang_goal_emb'].cuda(),
                                  o['lang_token_embs'].cuda(), o['bounds'].cuda())
        ref = qnet_oracle.qfunction_forward(sd, util.oracle_cfg(c), voxel_oracle.voxelize, o['rgb'], o['pcd'], o['proprio'],
                                            o['lang_token_embs'], o['bounds'], c['V'])
        assert util.rel_err(extra['q_trans'], ref['trans']) < util.Q_REL_TOL
        coords, rg, ic = qnet_oracle.choose_highest_action(ref['trans'], ref['rot_grip'], ref['collision'])
        xyz = qnet_oracle.attention_coordinate(coords, o['bounds'].expand(1, 6), c['V']).numpy()
        np.testing.assert_allclose(action.numpy(), act_oracle.continuous_action(xyz, rg.numpy(), ic.numpy(), 5), rtol=1e-5, atol=1e-5)
    assert actor.captures == 1
    with torch.no_grad():
        next(q.parameters()).mul_(1.0)             # bumps the version: prepared weights and the graph are stale
    actor.act([[r, p] for r, p in zip(rgb, pcd)], o['proprio'].cuda(), pcd, o['lang_goal_emb'].cuda(), o['lang_token_embs'].cuda(),
              o['bounds'].cuda())
    assert actor.captures == 2
