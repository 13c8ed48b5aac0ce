"""Agent-level drop-in: the REFERENCE's own agent classes -- PreprocessAgent(QAttentionStackAgent(
[QAttentionPerActBCAgent])) with its own QFunction, built the way launch_utils.create_agent builds them (launch_utils.py:
744-829) -- run on this library through voxactb_b200.install_shims(): `from voxel.voxel_grid import VoxelGrid` and
`from agents.peract_bc.perceiver_lang_io import PerceiverVoxelLangEncoder` resolve to this package, everything else is the
unmodified reference code (vendored into baseline/_ref/ by tools/vendor_reference.py; only the simulator / renderer / CLIP
weight packages, which the path never calls, are stubbed).

  * eval:  agent.build(training=False) -> load_weights of a checkpoint WRITTEN BY THE REFERENCE encoder class -> act();
           the 9-D continuous action equals the CPU oracle's (voxel oracle + Q-net oracle + the reference's own helpers);
  * train: agent.build(training=True) (the reference wraps our encoder in DistributedDataParallel) -> update(): the
           reference's losses, `total_loss.backward()` through this library's autograd node, the reference's Lamb step."""
import importlib
import importlib.util
import os
import sys
import types
from unittest import mock

import numpy as np
import pytest
import torch

import util
import make_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFS = [os.path.join(ROOT, 'baseline', '_ref'), '/root/reference']
REF = next((r for r in REFS if os.path.isdir(os.path.join(r, 'peract', 'agents', 'peract_bc'))), None)

pytestmark = pytest.mark.skipif(REF is None, reason='reference files not vendored (python tools/vendor_reference.py)')

CAMERAS = ['front', 'wrist']


class FakeClip:
    """Stands in for the CLIP RN50 text encoder (weights are not available offline): seeded embeddings per instruction."""

    def float(self):
        return self

    def to(self, device):
        self.device = device
        return self

    def eval(self):
        return self

    def state_dict(self):
        return {}

    def encode_text_with_embeddings(self, tokens):
        g = torch.Generator().manual_seed(int(tokens.sum()) % 100003)
        return (torch.randn(tokens.shape[0], 1024, generator=g).to(tokens.device),
                torch.randn(tokens.shape[0], 77, 512, generator=g).to(tokens.device))


def import_reference_agent():
    """Import the reference agent modules with this package shimmed in.  Returns (agent module, stack module,
    PreprocessAgent, reference utils module, reference encoder class)."""
    import voxactb_b200
    for p in (os.path.join(REF, 'peract'), os.path.join(REF, 'YARR')):
        if p not in sys.path:
            sys.path.insert(0, p)
    for name in ('pytorch3d', 'pytorch3d.transforms', 'pyrender', 'pyrender.trackball', 'trimesh', 'matplotlib',
                 'matplotlib.pyplot', 'ftfy', 'rlbench', 'rlbench.backend', 'rlbench.backend.const',
                 'rlbench.backend.observation_two_robots', 'pyrep', 'pyrep.const'):
        sys.modules.setdefault(name, mock.MagicMock())
    sys.modules['rlbench.backend.const'].DEPTH_SCALE = 2 ** 24 - 1
    clip = types.ModuleType('helpers.clip.core.clip')
    clip.load_clip = lambda *a, **k: (FakeClip(), None)
    clip.build_model = lambda sd: FakeClip()
    for name in ('helpers.clip', 'helpers.clip.core'):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules.setdefault(name, m)
    sys.modules['helpers.clip.core.clip'] = clip
    for name in ('agents', 'agents.peract_bc'):      # the packages' __init__ import the launcher (hydra, rlbench): bypass
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(REF, 'peract', name.replace('.', '/'))]
            sys.modules[name] = m
    # the reference encoder class itself (to write a reference checkpoint) under a private name, BEFORE the shim
    spec = importlib.util.spec_from_file_location('ref_perceiver_lang_io',
                                                  os.path.join(REF, 'peract', 'agents', 'peract_bc', 'perceiver_lang_io.py'))
    ref_plio = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_plio)
    voxactb_b200.install_shims()
    agent_mod = importlib.import_module('agents.peract_bc.qattention_peract_bc_agent')
    stack_mod = importlib.import_module('agents.peract_bc.qattention_stack_agent')
    pre_mod = importlib.import_module('helpers.preprocess_agent')
    utils_mod = importlib.import_module('helpers.utils')
    assert agent_mod.VoxelGrid is voxactb_b200.VoxelGrid                       # agent:17 now resolves to this package
    from agents.peract_bc.perceiver_lang_io import PerceiverVoxelLangEncoder     # launch_utils.py:21
    assert PerceiverVoxelLangEncoder is voxactb_b200.PerceiverVoxelLangEncoder
    return agent_mod, stack_mod, pre_mod.PreprocessAgent, utils_mod, ref_plio.PerceiverVoxelLangEncoder


def build_agent(mods, c, training, device, dropout):
    """launch_utils.create_agent (launch_utils.py:744-829) by hand, with the shimmed encoder class."""
    agent_mod, stack_mod, PreprocessAgent, _, _ = mods
    from agents.peract_bc.perceiver_lang_io import PerceiverVoxelLangEncoder
    kw = make_golden.encoder_kwargs(c) if dropout else make_golden.train_encoder_kwargs(c)
    enc = PerceiverVoxelLangEncoder(**kw)
    from voxactb_b200 import synth
    qagent = agent_mod.QAttentionPerActBCAgent(
        layer=0, coordinate_bounds=list(synth.SCENE_BOUNDS), perceiver_encoder=enc, camera_names=CAMERAS,
        voxel_size=c['V'], bounds_offset=0.15, image_crop_size=64, lr=5e-4, training_iterations=100, lr_scheduler=False,
        num_warmup_steps=10, trans_loss_weight=1.0, rot_loss_weight=1.0, grip_loss_weight=1.0, collision_loss_weight=1.0,
        include_low_dim_state=True, image_resolution=[c['H'], c['W']], batch_size=c['B'], voxel_feature_size=3,
        lambda_weight_l2=1e-6, num_rotation_classes=72, rotation_resolution=5, transform_augmentation=False,
        transform_augmentation_xyz=[0.125, 0.125, 0.125], transform_augmentation_rpy=[0.0, 0.0, 45.0],
        transform_augmentation_rot_resolution=5, optimizer_type='lamb', num_devices=1)
    stack = stack_mod.QAttentionStackAgent(qattention_agents=[qagent], rotation_resolution=5, camera_names=CAMERAS)
    agent = PreprocessAgent(pose_agent=stack)
    agent.build(training=training, device=device)
    return agent, qagent, enc


def test_reference_agent_imports_with_shims():
    """CPU: the reference agent module imports against this package (incl. voxel.augmentation) and builds
    an agent whose Q-function holds this package's encoder and voxelizer."""
    import voxactb_b200
    mods = import_reference_agent()
    c = make_golden.QNET_CASES['qnet_v20']
    agent, qagent, enc = build_agent(mods, dict(c, cameras=2), False, torch.device('cpu'), True)
    assert isinstance(qagent._q._qnet, voxactb_b200.PerceiverVoxelLangEncoder)
    assert isinstance(qagent._voxelizer, voxactb_b200.VoxelGrid)
    assert type(qagent._q).__module__ == 'agents.peract_bc.qattention_peract_bc_agent'      # the reference's QFunction
    assert all(not p.requires_grad for p in qagent._q.parameters())                          # agent:320-321
    import voxel.augmentation                                                               # noqa: F401  (agent:18)


@pytest.mark.gpu
def test_reference_agent_act_on_this_library(cuda_lib, tmp_path):
    from oracle import qnet_oracle, voxel_oracle
    from voxactb_b200 import synth
    mods = import_reference_agent()
    utils_mod, RefEnc = mods[3], mods[4]
    c = dict(make_golden.QNET_CASES['qnet_v20'], B=1, cameras=2)
    dev = torch.device('cuda:0')
    agent, qagent, enc = build_agent(mods, c, False, dev, True)
    # a checkpoint written by the REFERENCE encoder class, in the agent's own format (agent:878-880)
    ref_enc = RefEnc(**make_golden.encoder_kwargs(c))
    sd = synth.random_state_dict(ref_enc, 4242)
    ref_enc.load_state_dict(sd, strict=False)
    torch.save({'_qnet.' + k: v for k, v in ref_enc.state_dict().items()}, tmp_path / ('%s.pt' % qagent._name))
    qagent.load_weights(str(tmp_path))
    # one observation as rollout_generator feeds it: [time=1, batch=1, ...], rgb as 0..255
    obs_f = synth.make_observation(77, 1, 2, c['H'], c['W'], low_dim=4)
    observation = {'lang_goal_tokens': torch.randint(0, 1000, (1, 1, 77)), 'low_dim_state': obs_f['proprio'][None]}
    for n, rgb, pcd in zip(CAMERAS, obs_f['rgb'], obs_f['pcd']):
        observation['%s_rgb' % n] = torch.round((rgb + 1.0) / 2.0 * 255.0)[None]
        observation['%s_point_cloud' % n] = pcd[None]
        observation['%s_camera_extrinsics' % n] = torch.eye(4)[None, None]
        observation['%s_camera_intrinsics' % n] = torch.tensor([[[[100., 0., c['W'] / 2], [0., 100., c['H'] / 2], [0., 0., 1.]]]])
    res = agent.act(0, dict(observation), deterministic=True)
    torch.cuda.synchronize()
    # oracle on the same inputs (CPU): the reference's preprocessing, voxel + Q-net oracle, the reference's own helpers
    rgb = [(observation['%s_rgb' % n][0].float() / 255.0) * 2.0 - 1.0 for n in CAMERAS]
    lang_emb, lang_tok = FakeClip().encode_text_with_embeddings(observation['lang_goal_tokens'][0].long())
    cfg = util.oracle_cfg(c)
    sd_cpu = {k: v.cpu() for k, v in ref_enc.state_dict().items()}
    ref = qnet_oracle.qfunction_forward(sd_cpu, cfg, voxel_oracle.voxelize, rgb, obs_f['pcd'], obs_f['proprio'], lang_tok,
                                        obs_f['bounds'], c['V'])
    coords, rg, ic = qnet_oracle.choose_highest_action(ref['trans'], ref['rot_grip'], ref['collision'])
    xyz = qnet_oracle.attention_coordinate(coords, obs_f['bounds'], c['V'])[0].numpy()
    expect = np.concatenate([xyz, utils_mod.discrete_euler_to_quaternion(rg[0, :3].numpy(), 5), rg[0, 3:].numpy(),
                             [float(ic[0, 0])]])
    np.testing.assert_allclose(np.asarray(res.action, dtype=np.float64), expect, rtol=1e-5, atol=1e-5)
    assert np.array_equal(res.observation_elements['trans_action_indicies'], coords[0].numpy())
    # the Q-values the reference agent kept for its summaries came from this library and match the oracle
    assert util.rel_err(qagent._act_qvalues, torch.softmax(ref['trans'].reshape(1, -1), 1).reshape(ref['trans'].shape)[0]) < 1e-3


@pytest.mark.gpu
def test_reference_agent_update_on_this_library(cuda_lib):
    """The reference's update() (losses :517-578, backward :581, its Lamb) drives this library's training step."""
    import torch.distributed as dist
    from oracle import train_oracle
    from voxactb_b200 import synth, _lib
    mods = import_reference_agent()
    c = dict(make_golden.TRAIN_CASES['train_v20'])
    if not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29533')
        dist.init_process_group('nccl', rank=0, world_size=1)
    agent, qagent, enc = build_agent(mods, c, True, 0, False)
    enc.math_mode = _lib.MATH_FP32_SIMT
    sd = synth.random_state_dict(enc, make_golden.weight_seed(c))
    enc.load_state_dict(sd, strict=False)
    assert type(qagent._q._qnet).__name__ == 'DistributedDataParallel'                    # agent:50-54
    obs = synth.make_observation(c['seed'], c['B'], c['cameras'], c['H'], c['W'], low_dim=c['low_dim'])
    lab = make_golden.train_labels(c)
    dev = torch.device('cuda:0')
    B = c['B']
    sample = {'trans_action_indicies': lab['trans'].float(), 'rot_grip_action_indicies': lab['rot_grip'].float(),
              'gripper_pose': torch.zeros(B, 7), 'ignore_collisions': lab['collision'].float(),
              'lang_goal_emb': obs['lang_goal_emb'], 'lang_token_embs': obs['lang_token_embs'],
              'low_dim_state': obs['proprio'], 'demo': torch.ones(B, dtype=torch.bool)}
    for n, rgb, pcd in zip(CAMERAS, obs['rgb'], obs['pcd']):
        sample['%s_rgb' % n] = torch.round((rgb + 1.0) / 2.0 * 255.0)      # PreprocessAgent normalises 0..255 -> [-1, 1]
        sample['%s_point_cloud' % n] = pcd
    # replay samples carry a task axis at dim 1 for tensors with more than 2 dims (preprocess_agent.py:24-25)
    sample = {k: (v[:, None] if v.dim() > 2 else v).to(dev) for k, v in sample.items()}
    before = {k: v.detach().clone() for k, v in enc.state_dict().items() if k in sd}
    out = agent.update(0, sample)
    torch.cuda.synchronize()
    # oracle: same quantised images
    rgb_q = [(torch.round((r + 1.0) / 2.0 * 255.0) / 255.0) * 2.0 - 1.0 for r in obs['rgb']]
    sdp = {k: v for k, v in sd.items() if not k.endswith(('pos_x', 'pos_y', 'pos_z'))}
    res = train_oracle.training_step(sdp, util.oracle_cfg(c), rgb_q, obs['pcd'], obs['proprio'], obs['lang_token_embs'],
                                     obs['bounds'], c['V'], lab)
    assert abs(float(out['total_losses']) - float(res['total'])) <= 5e-5 * abs(float(res['total']))
    # parameters moved the way the oracle's LAMB step moves them (the first step is sign-like: compare directions)
    moved, worst = 0, (1.0, None)
    for k, new in res['params'].items():
        ours = enc.state_dict()[k].detach().cpu().double() - before[k].cpu().double()
        ref_d = new.double() - sdp[k].double()
        if float(ref_d.norm()) == 0 or k == 'trans_decoder.conv3d.bias':
            continue          # d loss / d bias = sum_v (softmax - onehot) = 0 analytically: the sign-like step follows rounding noise
        cos = float((ours * ref_d).sum() / (ours.norm() * ref_d.norm()).clamp_min(1e-30))
        worst = min(worst, (cos, k))
        moved += 1
    print('reference-agent update: %d parameters moved, worst direction cosine vs the oracle step %.4f (%s)' % (moved, *worst))
    assert moved > 50 and worst[0] > 0.98, (moved, worst)
