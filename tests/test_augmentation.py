"""Row f2: SE(3) augmentation -- oracle pinned to the reference, device kernel against the oracle."""
import importlib.util
import os
import sys
import types
from unittest import mock

import numpy as np
import pytest
import torch
from scipy.spatial.transform import Rotation

from oracle import aug_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = next((p for p in (os.path.join(ROOT, 'baseline', '_ref', 'peract'), '/root/reference/peract')
            if os.path.exists(os.path.join(p, 'voxel', 'augmentation.py'))), None)
needs_ref = pytest.mark.skipif(REF is None, reason='reference voxel/augmentation.py not available')


def load_reference():
    """The reference module by file path; pytorch3d.transforms -> the restated conversions (the package is not installed)."""
    from voxactb_b200 import augmentation as ours
    for name in ('pyrender', 'pyrender.trackball', 'trimesh', 'rlbench', 'rlbench.backend', 'rlbench.backend.const',
                 'rlbench.backend.observation_two_robots', 'pyrep', 'pyrep.const'):
        sys.modules.setdefault(name, mock.MagicMock())
    tf = types.ModuleType('pytorch3d.transforms')
    tf.quaternion_to_matrix, tf.euler_angles_to_matrix, tf.matrix_to_quaternion = (
        ours.quaternion_to_matrix, ours.euler_angles_to_matrix, ours.matrix_to_quaternion)
    p3 = types.ModuleType('pytorch3d')
    p3.transforms = tf
    saved = {k: sys.modules.get(k) for k in ('pytorch3d', 'pytorch3d.transforms', 'helpers', 'helpers.utils')}
    sys.modules['pytorch3d'], sys.modules['pytorch3d.transforms'] = p3, tf
    spec = importlib.util.spec_from_file_location('_vxb_ref_helpers_utils', os.path.join(REF, 'helpers', 'utils.py'))
    hu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(hu)
    helpers = types.ModuleType('helpers')
    helpers.utils = hu
    sys.modules['helpers'], sys.modules['helpers.utils'] = helpers, hu
    try:
        spec = importlib.util.spec_from_file_location('_vxb_ref_augmentation', os.path.join(REF, 'voxel', 'augmentation.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def make_inputs(seed, bs=3, cams=2, H=12, W=16):
    g = torch.Generator().manual_seed(seed)
    pcd = [torch.rand(bs, 3, H, W, generator=g) * torch.tensor([1.0, 1.0, 1.0]).view(1, 3, 1, 1)
           + torch.tensor([-0.3, -0.5, 0.6]).view(1, 3, 1, 1) for _ in range(cams)]
    bounds = torch.tensor([[-0.3, -0.5, 0.6, 0.7, 0.5, 1.6]]).repeat(bs, 1)
    pos = torch.rand(bs, 3, generator=g) * 0.5 + torch.tensor([-0.05, -0.25, 0.85])
    quat = torch.from_numpy(Rotation.random(bs, random_state=seed).as_quat()).float()      # xyzw
    pose = torch.cat([pos, quat], 1)
    trans_idx = torch.randint(0, 100, (bs, 3), generator=g)
    rot_grip = torch.cat([torch.randint(0, 72, (bs, 3), generator=g), torch.randint(0, 2, (bs, 1), generator=g)], 1)
    return pcd, pose, trans_idx, rot_grip, bounds


def test_rotation_conversions_match_scipy():
    from voxactb_b200 import augmentation as ours
    rot = Rotation.random(64, random_state=3)
    q_xyzw = torch.from_numpy(rot.as_quat()).float()
    q_wxyz = torch.cat([q_xyzw[:, 3:], q_xyzw[:, :3]], 1)
    m = ours.quaternion_to_matrix(q_wxyz)
    np.testing.assert_allclose(m.numpy(), rot.as_matrix(), atol=2e-6)
    q = ours.matrix_to_quaternion(m)
    q = q * torch.sign(q[:, :1]) * torch.sign(q_wxyz[:, :1])
    np.testing.assert_allclose(q.numpy(), q_wxyz.numpy(), atol=2e-6)
    e = torch.from_numpy(np.random.default_rng(0).uniform(-1, 1, (64, 3))).float()
    np.testing.assert_allclose(ours.euler_angles_to_matrix(e, 'XYZ').numpy(), Rotation.from_euler('XYZ', e.numpy()).as_matrix(),
                               atol=2e-6)


@needs_ref
def test_oracle_perturb_se3_matches_reference():
    ref = load_reference()
    pcd, pose, _, _, bounds = make_inputs(0)
    bs = pose.shape[0]
    eye = torch.eye(4).repeat(bs, 1, 1)
    t4, r4, a4 = eye.clone(), eye.clone(), eye.clone()
    t4[:, :3, 3] = torch.tensor([[0.05, -0.02, 0.6], [-0.9, 0.1, 0.0], [0.0, 0.0, 0.03]])   # sample 0 / 1 hit the clamp
    r4[:, :3, :3] = torch.from_numpy(Rotation.from_euler('XYZ', [[0, 0, 0.3], [0.1, 0, -0.5], [0, 0.2, 0.7]]).as_matrix()).float()
    a4[:, :3, 3] = pose[:, :3]
    want = ref.perturb_se3(pcd, t4, r4, a4, bounds)
    got = aug_oracle.perturb_se3([p.numpy() for p in pcd], t4.numpy(), r4.numpy(), a4.numpy(), bounds.numpy())
    for w, g in zip(want, got):
        np.testing.assert_allclose(g, w.numpy(), rtol=0, atol=3e-7)


@needs_ref
@pytest.mark.gpu
def test_apply_se3_augmentation_matches_reference_under_the_same_seed(cuda_lib):
    """Same torch seed -> the reference (CPU tensors) and this module (CUDA point clouds) draw the same perturbation, return the
    same discretised actions and the same perturbed point clouds."""
    from voxactb_b200 import augmentation as ours
    ref = load_reference()
    for seed in (1, 2, 3):
        pcd, pose, trans_idx, rot_grip, bounds = make_inputs(seed)
        args = (0, torch.tensor([0.125, 0.125, 0.125]), [0.0, 0.0, 45.0], 5, 100, 5)
        torch.manual_seed(100 + seed)
        wt, wr, wp = ref.apply_se3_augmentation(pcd, pose, trans_idx, rot_grip, bounds, *args, 'cpu')
        torch.manual_seed(100 + seed)
        gt, gr, gp = ours.apply_se3_augmentation([p.cuda() for p in pcd], pose.cuda(), trans_idx.cuda(), rot_grip.cuda(), bounds.cuda(),
                                                 *args, torch.device('cuda'))
        assert torch.equal(gt.cpu(), wt) and torch.equal(gr.cpu(), wr)
        for w, g in zip(wp, gp):
            np.testing.assert_allclose(g.cpu().numpy(), w.numpy(), rtol=0, atol=5e-7)
        pr, pl = pose, pose.roll(1, 0)
        torch.manual_seed(200 + seed)
        w5 = ref.apply_se3_augmentation_2Robots(pcd, pr, trans_idx, rot_grip, pl, trans_idx, rot_grip, bounds, *args, 'cpu')
        torch.manual_seed(200 + seed)
        g5 = ours.apply_se3_augmentation_2Robots([p.cuda() for p in pcd], pr.cuda(), trans_idx.cuda(), rot_grip.cuda(), pl.cuda(),
                                                 trans_idx.cuda(), rot_grip.cuda(), bounds.cuda(), *args, torch.device('cuda'))
        for w, g in zip(w5[:4], g5[:4]):
            assert torch.equal(g.cpu(), w)
        for w, g in zip(w5[4], g5[4]):
            np.testing.assert_allclose(g.cpu().numpy(), w.numpy(), rtol=0, atol=5e-7)


@pytest.mark.gpu
def test_device_perturb_se3_matches_oracle_and_voxel_indices(cuda_lib):
    """Full-size clouds (4 cameras 128x128, B=4): the kernel against the numpy restatement, and the voxel indices of the
    perturbed clouds against the oracle's (equal except for points within float rounding of a voxel face)."""
    from oracle import voxel_oracle
    from voxactb_b200 import augmentation as ours, synth
    obs = synth.make_observation(7, 4, 4, 128, 128)
    bs = 4
    eye = torch.eye(4).repeat(bs, 1, 1)
    t4, r4, a4 = eye.clone(), eye.clone(), eye.clone()
    rng = np.random.default_rng(5)
    t4[:, :3, 3] = torch.from_numpy(rng.uniform(-0.1, 0.1, (bs, 3))).float()
    r4[:, :3, :3] = torch.from_numpy(Rotation.from_euler('XYZ', rng.uniform(-0.6, 0.6, (bs, 3))).as_matrix()).float()
    a4[:, :3, 3] = torch.tensor([[0.2, 0.0, 1.0]]).repeat(bs, 1) + torch.from_numpy(rng.uniform(-0.1, 0.1, (bs, 3))).float()
    got = ours.perturb_se3([p.cuda() for p in obs['pcd']], t4, r4, a4, obs['bounds'])
    want = aug_oracle.perturb_se3([p.numpy() for p in obs['pcd']], t4.numpy(), r4.numpy(), a4.numpy(), obs['bounds'].numpy())
    for w, g in zip(want, got):
        np.testing.assert_allclose(g.cpu().numpy(), w, rtol=0, atol=5e-7)
    flat_w = np.concatenate([w.reshape(bs, 3, -1).transpose(0, 2, 1) for w in want], 1)
    flat_g = np.concatenate([g.cpu().numpy().reshape(bs, 3, -1).transpose(0, 2, 1) for g in got], 1)
    iw = voxel_oracle.voxel_indices(flat_w, obs['bounds'].numpy(), 100)
    ig = voxel_oracle.voxel_indices(flat_g, obs['bounds'].numpy(), 100)
    assert (iw != ig).any(-1).mean() < 1e-4
