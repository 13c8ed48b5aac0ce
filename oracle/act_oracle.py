"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the act() tail.

Follows /root/reference/peract/agents/peract_bc/qattention_stack_agent.py:78-89 (continuous_action = concat of the attention
coordinate, the quaternion of the discrete Euler bins, the gripper bin and the ignore-collision bin) and
/root/reference/peract/helpers/utils.py:103-105 (discrete_euler_to_quaternion, scipy Rotation, extrinsic 'xyz', degrees).
Pinned live against the reference helper when /root/reference (or baseline/_ref) is present (tests/test_act.py)."""
import numpy as np
from scipy.spatial.transform import Rotation


def discrete_euler_to_quaternion(discrete_euler, resolution):
    euler = (np.asarray(discrete_euler) * resolution) - 180                       # utils.py:104
    return Rotation.from_euler('xyz', euler, degrees=True).as_quat()              # utils.py:105


def continuous_action(attention_xyz, rot_grip_idx, coll_idx, rotation_resolution):
    """attention_xyz [B,3], rot_grip_idx [B,4] (3 Euler bins + gripper), coll_idx [B] -> [B,9] (stack_agent.py:83-88)."""
    xyz = np.asarray(attention_xyz, dtype=np.float64)
    rg = np.asarray(rot_grip_idx)
    out = np.empty((xyz.shape[0], 9))
    for b in range(xyz.shape[0]):
        out[b] = np.concatenate([xyz[b], discrete_euler_to_quaternion(rg[b, :3], rotation_resolution), rg[b, 3:4],
                                 [float(np.asarray(coll_idx).reshape(-1)[b])]])
    return out
