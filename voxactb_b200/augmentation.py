"""SE(3) augmentation on the device -- drop-in for the reference's ``voxel.augmentation``
(peract/voxel/augmentation.py): same function names, arguments and return values.

What moves: ``perturb_se3`` -- the only part that touches the 3 MB/sample point clouds -- is ONE streaming CUDA kernel per
camera (vxb_se3_perturb_f32) instead of ~10 full-size torch temporaries.  What stays on the host: sampling and rejecting the
perturbation (augmentation.py:112-176) works on a few floats per sample; the reference runs it on device tensors with a
``.cpu()`` round trip per attempt and per sample, here the poses and bounds cross PCIe once.  The random draws are the
reference's own (``torch.rand`` / ``torch.randint`` on the global CPU generator, helpers/utils.py:501-508) in the same order,
so a seeded run perturbs identically.  pytorch3d (quaternion / Euler conversions, not installed in this image) is restated
from its published formulas below.
"""
import numpy as np
import torch
from scipy.spatial.transform import Rotation

from . import _lib


# ---- pytorch3d.transforms restated (rotation_conversions.py: quaternion_to_matrix, euler_angles_to_matrix, matrix_to_quaternion)
def quaternion_to_matrix(q):
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def _axis_rotation(axis, angle):
    c, s = torch.cos(angle), torch.sin(angle)
    one, zero = torch.ones_like(angle), torch.zeros_like(angle)
    flat = {'X': (one, zero, zero, zero, c, -s, zero, s, c),
            'Y': (c, zero, s, zero, one, zero, -s, zero, c),
            'Z': (c, -s, zero, s, c, zero, zero, zero, one)}[axis]
    return torch.stack(flat, -1).reshape(angle.shape + (3, 3))


def euler_angles_to_matrix(angles, convention):
    mats = [_axis_rotation(c, e) for c, e in zip(convention, torch.unbind(angles, -1))]
    return torch.matmul(torch.matmul(mats[0], mats[1]), mats[2])


def matrix_to_quaternion(m):
    """Real part first; the branch with the largest denominator is selected (pytorch3d >= 0.6)."""
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(m.reshape(m.shape[:-2] + (9,)), -1)
    q_abs = torch.sqrt(torch.clamp(torch.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22,
                                                1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], -1), min=0.0))
    cand = torch.stack([torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], -1),
                        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], -1),
                        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], -1),
                        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], -1)], -2)
    cand = cand / (2.0 * q_abs[..., None].clamp(min=0.1))
    best = q_abs.argmax(-1)
    return torch.gather(cand, -2, best[..., None, None].expand(best.shape + (1, 4))).squeeze(-2)


# ---- helpers/utils.py:63-64, 92-97, 104-116, 501-508
def _point_to_voxel_index(point, voxel_size, coord_bounds):
    bb_mins, bb_maxs = np.array(coord_bounds[0:3]), np.array(coord_bounds[3:])
    dims_m_one = np.array([voxel_size] * 3) - 1
    res = (bb_maxs - bb_mins) / (np.array([voxel_size] * 3) + 1e-12)
    return np.minimum(np.floor((point - bb_mins) / (res + 1e-12)).astype(np.int32), dims_m_one)


def _quaternion_to_discrete_euler(quaternion, resolution):
    euler = Rotation.from_quat(quaternion).as_euler('xyz', degrees=True) + 180
    assert np.min(euler) >= 0 and np.max(euler) <= 360
    disc = np.around((euler / resolution)).astype(int)
    disc[disc == int(360 / resolution)] = 0
    return disc


def _rand_dist(size, lo=-1.0, hi=1.0):
    return (hi - lo) * torch.rand(size) + lo


def _rand_discrete(size, lo=0, hi=1):
    if lo == hi:
        return torch.zeros(size)
    return torch.randint(lo, hi + 1, size)


def perturb_se3(pcd, trans_shift_4x4, rot_shift_4x4, action_gripper_4x4, bounds):
    """Reference augmentation.py:7-65.  pcd: list of [bs,3,H,W] (or [bs,3,N]) CUDA tensors; the 4x4 matrices and bounds may
    live on either device (they are a few floats).  Returns new tensors of the same shapes."""
    bs = pcd[0].shape[0]
    dev = pcd[0].device
    if not pcd[0].is_cuda:
        raise RuntimeError('voxactb_b200.augmentation.perturb_se3 runs on CUDA only (no CPU fallback); got %s' % dev)
    bounds = bounds.detach().float().cpu().reshape(-1, 6)
    if bounds.shape[0] != bs:
        bounds = bounds.repeat(bs, 1)
    a = action_gripper_4x4.detach().float().cpu()[:, 0:3, 3]
    t = trans_shift_4x4.detach().float().cpu()[:, 0:3, 3]
    lo = torch.stack([bounds[:, 0].min(), bounds[:, 1].min(), bounds[:, 2].min()])          # :44-46 (batch-wide extrema)
    hi = torch.stack([bounds[:, 3].max(), bounds[:, 4].max(), bounds[:, 5].max()])
    c = torch.minimum(torch.maximum(a + t, lo), hi)                                         # :48-54
    xform = torch.cat([a, rot_shift_4x4.detach().float().cpu()[:, :3, :3].reshape(bs, 9), c], 1).contiguous()
    xform = xform.to(dev, non_blocking=True)
    L = _lib.lib()
    out = []
    for p in pcd:
        p = _lib.f32(p)
        o = torch.empty_like(p)
        n = p.numel() // (bs * 3)
        _lib.check(L.vxb_se3_perturb_f32(_lib.ptr(p), _lib.ptr(xform), _lib.ptr(o), bs, n, _lib.stream()), 'vxb_se3_perturb_f32')
        out.append(o)
    return out


def _augment(pcd, poses, trans, rot_grips, bounds, layer, trans_aug_range, rot_aug_range, rot_aug_resolution, voxel_size,
             rot_resolution, device, max_attempts):
    bs = pcd[0].shape[0]
    identity = torch.eye(4).unsqueeze(0).repeat(bs, 1, 1)
    bounds_h = bounds.detach().float().cpu()
    grip_h = [rg.detach().cpu() for rg in rot_grips]
    a4 = []
    for pose in poses:
        pose = pose.detach().float().cpu()
        m = identity.clone()
        m[:, :3, :3] = quaternion_to_matrix(torch.cat((pose[:, 6].unsqueeze(1), pose[:, 3:6]), dim=1))      # :102-107
        m[:, 0:3, 3] = pose[:, :3]
        a4.append(m)
    done = False
    attempts = 0
    tar = torch.as_tensor(trans_aug_range, dtype=torch.float32).cpu()
    while not done:
        attempts += 1
        if attempts > max_attempts:
            raise Exception('Failing to perturb action and keep it within bounds.')
        trans_range = (bounds_h[:, 3:] - bounds_h[:, :3]) * tar                                            # :121
        trans_shift = trans_range * _rand_dist((bs, 3))
        trans_shift_4x4 = identity.clone()
        trans_shift_4x4[:, 0:3, 3] = trans_shift
        steps = [int(rot_aug_range[i] // rot_aug_resolution) for i in range(3)]                             # :127-129
        rpy = [_rand_discrete((bs, 1), -s, s) * np.deg2rad(rot_aug_resolution) for s in steps]
        rot_shift_4x4 = identity.clone()
        rot_shift_4x4[:, :3, :3] = euler_angles_to_matrix(torch.cat(rpy, dim=1).float(), 'XYZ')
        new_trans, new_rg = [], []
        done = True
        for m, rg in zip(a4, grip_h):
            pm = torch.bmm(m, rot_shift_4x4)                                                                # :145
            pm[:, 0:3, 3] += trans_shift
            p_trans = pm[:, 0:3, 3].numpy()
            q_wxyz = matrix_to_quaternion(pm[:, :3, :3])
            q_xyzw = torch.cat([q_wxyz[:, 1:], q_wxyz[:, 0].unsqueeze(1)], dim=1).numpy()
            ti, ri = [], []
            for b in range(bs):
                bnd = bounds_h[b if layer > 0 else 0].numpy()
                ti.append(_point_to_voxel_index(p_trans[b], voxel_size, bnd).tolist())
                quat = np.array(q_xyzw[b]) / np.linalg.norm(q_xyzw[b], axis=-1, keepdims=True)
                if quat[-1] < 0:
                    quat = -quat
                ri.append(_quaternion_to_discrete_euler(quat, rot_resolution).tolist() + [int(rg[b, 3].numpy())])
            ti, ri = torch.from_numpy(np.array(ti)), torch.from_numpy(np.array(ri))
            done = done and not bool(torch.any(ti < 0))
            new_trans.append(ti)
            new_rg.append(ri)
    pcd = perturb_se3(pcd, trans_shift_4x4, rot_shift_4x4, a4[0], bounds_h)                                # :181 / 2 robots: right arm
    return [t.to(device) for t in new_trans], [r.to(device) for r in new_rg], pcd


def apply_se3_augmentation(pcd, action_gripper_pose, action_trans, action_rot_grip, bounds, layer, trans_aug_range,
                           rot_aug_range, rot_aug_resolution, voxel_size, rot_resolution, device):
    """Reference augmentation.py:68-183 (same arguments, same returns: action_trans, action_rot_grip, pcd)."""
    t, r, pcd = _augment(pcd, [action_gripper_pose], [action_trans], [action_rot_grip], bounds, layer, trans_aug_range,
                         rot_aug_range, rot_aug_resolution, voxel_size, rot_resolution, device, 100)
    return t[0], r[0], pcd


def apply_se3_augmentation_2Robots(pcd, action_gripper_pose_right, action_trans_right, action_rot_grip_right,
                                   action_gripper_pose_left, action_trans_left, action_rot_grip_left, bounds, layer,
                                   trans_aug_range, rot_aug_range, rot_aug_resolution, voxel_size, rot_resolution, device):
    """Reference augmentation.py:186-348: both keyframe poses get the SAME perturbation; the point clouds turn about the right
    arm's pose."""
    t, r, pcd = _augment(pcd, [action_gripper_pose_right, action_gripper_pose_left], [action_trans_right, action_trans_left],
                         [action_rot_grip_right, action_rot_grip_left], bounds, layer, trans_aug_range, rot_aug_range,
                         rot_aug_resolution, voxel_size, rot_resolution, device, 400)
    return t[0], r[0], t[1], r[1], pcd
