"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the SE(3) point-cloud perturbation.

Follows /root/reference/peract/voxel/augmentation.py:7-65 (perturb_se3) in numpy float32, step by step.  Pinned live against
the reference function (loaded by file path; its pytorch3d import is stubbed -- perturb_se3 itself is pure torch) in
tests/test_augmentation.py.  The sampling half (apply_se3_augmentation, :68-183) is checked by running the REFERENCE function
with a pytorch3d stub next to voxactb_b200.augmentation under the same torch seed (pytorch3d is an un-vendored dependency that
is not installed in this image; its three conversions are restated from the published formulas in the product module and
checked here against scipy)."""
import numpy as np


def perturb_se3(pcd, trans_shift_4x4, rot_shift_4x4, action_gripper_4x4, bounds):
    """pcd: list of [bs,3,H,W] float32 arrays -> list of perturbed arrays (augmentation.py:25-65)."""
    bs = pcd[0].shape[0]
    bounds = np.asarray(bounds, np.float32).reshape(-1, 6)
    if bounds.shape[0] != bs:
        bounds = np.repeat(bounds, bs, 0)
    a = np.asarray(action_gripper_4x4, np.float32)[:, 0:3, 3]                       # :30
    t = np.asarray(trans_shift_4x4, np.float32)[:, 0:3, 3]                          # :31
    R = np.asarray(rot_shift_4x4, np.float32)
    lo = np.array([bounds[:, 0].min(), bounds[:, 1].min(), bounds[:, 2].min()], np.float32)   # :44-46
    hi = np.array([bounds[:, 3].max(), bounds[:, 4].max(), bounds[:, 5].max()], np.float32)
    c = np.clip((a + t).astype(np.float32), lo, hi)                                 # :48-54
    out = []
    for p in pcd:
        p = np.asarray(p, np.float32)
        flat = p.reshape(bs, 3, -1)
        p4 = np.ones((bs, 4, flat.shape[-1]), np.float32)
        p4[:, :3, :] = flat - a[:, :, None]                                          # :37
        rot = np.einsum('bnj,bjk->bnk', p4.transpose(0, 2, 1), R).transpose(0, 2, 1).astype(np.float32)   # :40-41
        out.append((rot[:, :3, :] + c[:, :, None]).astype(np.float32).reshape(p.shape))               # :61-63
    return out
