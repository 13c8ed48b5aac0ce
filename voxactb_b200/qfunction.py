"""``QFunction`` -- the drop-in boundary (reference qattention_peract_bc_agent.py:31-135).

Same constructor and ``forward`` contract as the reference class the agent builds
(``QAttentionPerActBCAgent.build``, :236-242): it flattens the per-camera point clouds and RGB
images, voxelises them (``VoxelGrid``), and runs the Q-network (``PerceiverVoxelLangEncoder``);
``choose_highest_action`` / ``_argmax_3d`` run as one fused device kernel
(``vxb_select_action_f32``) instead of a chain of torch ops.
"""
import torch
from torch import nn

from . import _lib
from .voxel_grid import VoxelGrid


class QFunction(nn.Module):

    def __init__(self, perceiver_encoder: nn.Module, voxelizer: VoxelGrid, bounds_offset: float,
                 rotation_resolution: float, device, training, arm_pred_loss=False):
        super().__init__()
        self._rotation_resolution = rotation_resolution
        self._voxelizer = voxelizer
        self._bounds_offset = bounds_offset
        self._qnet = perceiver_encoder.to(device)
        self._arm_pred_loss = arm_pred_loss
        self._is_training = training
        self._select_ws = None

    def _select(self, q_trans, q_rot_grip, q_collision, bounds=None):
        B, V = q_trans.shape[0], q_trans.shape[-1]
        dev = q_trans.device
        L = _lib.lib()
        ws_bytes = L.vxb_select_action_workspace_bytes(B, V)
        if self._select_ws is None or self._select_ws.numel() < ws_bytes or self._select_ws.device != dev:
            self._select_ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        coords = torch.empty(B, 3, dtype=torch.int32, device=dev)
        rg = torch.empty(B, 4, dtype=torch.int32, device=dev) if q_rot_grip is not None else None
        coll = torch.empty(B, dtype=torch.int32, device=dev) if q_collision is not None else None
        xyz = torch.empty(B, 3, dtype=torch.float32, device=dev) if bounds is not None else None
        R = int(360 // self._rotation_resolution)
        # the kernel reads rows of 3R + 2 rotation/gripper logits and 2 collision logits (reference :65-80)
        if q_rot_grip is not None and tuple(q_rot_grip.shape) != (B, 3 * R + 2):
            raise ValueError('q_rot_grip must be [%d,%d] (3 x %d rotation bins + 2 gripper), got %s'
                             % (B, 3 * R + 2, R, tuple(q_rot_grip.shape)))
        if q_collision is not None and tuple(q_collision.shape) != (B, 2):
            raise ValueError('q_collision must be [%d,2], got %s' % (B, tuple(q_collision.shape)))
        bnd = _lib.f32(bounds.reshape(-1, 6)) if bounds is not None else None
        rc = L.vxb_select_action_f32(
            _lib.ptr(_lib.f32(q_trans)), _lib.ptr(_lib.f32(q_rot_grip)) if q_rot_grip is not None else None,
            _lib.ptr(_lib.f32(q_collision)) if q_collision is not None else None, _lib.ptr(bnd),
            bnd.shape[0] if bnd is not None else 1, B, V, R, _lib.ptr(coords), _lib.ptr(rg), _lib.ptr(coll),
            _lib.ptr(xyz), _lib.ptr(self._select_ws), ws_bytes, _lib.stream())
        _lib.check(rc, 'vxb_select_action_f32')
        return coords, rg, coll, xyz

    def _argmax_3d(self, tensor_orig):
        """[B,1,D,H,W] -> int64 [B,3] voxel index of the maximum (reference :57-63)."""
        return self._select(tensor_orig, None, None)[0].long()

    def choose_highest_action(self, q_trans, q_rot_grip, q_collision):
        """Reference :65-80: argmax voxel, per-axis rotation bins + gripper, collision flag."""
        coords, rg, coll, _ = self._select(q_trans, q_rot_grip, q_collision)
        return (coords.long(), rg.long() if rg is not None else None,
                coll.long().unsqueeze(-1) if coll is not None else None)

    def select_action(self, q_trans, q_rot_grip, q_collision, bounds):
        """Fused act() tail (reference :709-724): also returns the metric attention coordinate
        bounds_min + res*idx + res/2."""
        return self._select(q_trans, q_rot_grip, q_collision, bounds)

    def forward(self, rgb_pcd, proprio, pcd, lang_goal_emb, lang_token_embs, bounds=None,
                prev_bounds=None, prev_layer_voxel_grid=None):
        """Reference :82-135.  rgb_pcd: list of [rgb, pcd] per camera, pcd: list of [B,3,H,W]."""
        b = rgb_pcd[0][0].shape[0]
        pcd_flat = torch.cat([p.permute(0, 2, 3, 1).reshape(b, -1, 3) for p in pcd], 1)
        rgb = [rp[0] for rp in rgb_pcd]
        feat_size = rgb[0].shape[1]
        flat_imag_features = torch.cat(
            [p.permute(0, 2, 3, 1).reshape(b, -1, feat_size) for p in rgb], 1)
        voxel_grid = self._voxelizer.coords_to_bounding_voxel_grid(
            pcd_flat, coord_features=flat_imag_features, coord_bounds=bounds)
        voxel_grid = voxel_grid.permute(0, 4, 1, 2, 3).detach()   # channels-first VIEW, no copy
        if bounds.shape[0] != b:
            bounds = bounds.repeat(b, 1)
        out = self._qnet(voxel_grid, proprio, lang_goal_emb, lang_token_embs, prev_layer_voxel_grid,
                         bounds, prev_bounds)
        if self._arm_pred_loss:
            q_trans, q_rot_and_grip, q_ignore_collisions, arm_out = out
            if self._is_training:
                return q_trans, q_rot_and_grip, q_ignore_collisions, voxel_grid, arm_out
            return q_trans, q_rot_and_grip, q_ignore_collisions, voxel_grid
        q_trans, q_rot_and_grip, q_ignore_collisions = out
        return q_trans, q_rot_and_grip, q_ignore_collisions, voxel_grid


class QFunction2Robots(QFunction):
    """Drop-in for reference qattention_peract_bc_agent.py:882-963 (the 2-robot / "one policy, more heads"
    variant): same voxelisation, one ``PerceiverVoxelLang2RobotsEncoder`` pass, two head sets."""

    def __init__(self, perceiver_encoder: nn.Module, voxelizer: VoxelGrid, bounds_offset: float,
                 rotation_resolution: float, device, training):
        super().__init__(perceiver_encoder, voxelizer, bounds_offset, rotation_resolution, device, training, False)

    def forward(self, rgb_pcd, proprio_right, proprio_left, pcd, lang_goal_emb, lang_token_embs, bounds=None,
                prev_bounds=None, prev_layer_voxel_grid=None):
        b = rgb_pcd[0][0].shape[0]
        pcd_flat = torch.cat([p.permute(0, 2, 3, 1).reshape(b, -1, 3) for p in pcd], 1)
        rgb = [rp[0] for rp in rgb_pcd]
        feat_size = rgb[0].shape[1]
        flat_imag_features = torch.cat(
            [p.permute(0, 2, 3, 1).reshape(b, -1, feat_size) for p in rgb], 1)
        voxel_grid = self._voxelizer.coords_to_bounding_voxel_grid(
            pcd_flat, coord_features=flat_imag_features, coord_bounds=bounds)
        voxel_grid = voxel_grid.permute(0, 4, 1, 2, 3).detach()
        if bounds.shape[0] != b:
            bounds = bounds.repeat(b, 1)
        (q_trans_right, q_rot_and_grip_right, q_ignore_collisions_right,
         q_trans_left, q_rot_and_grip_left, q_ignore_collisions_left) = self._qnet(
            voxel_grid, proprio_right, proprio_left, lang_goal_emb, lang_token_embs, prev_layer_voxel_grid,
            bounds, prev_bounds)
        return (q_trans_right, q_rot_and_grip_right, q_ignore_collisions_right, voxel_grid,
                q_trans_left, q_rot_and_grip_left, q_ignore_collisions_left)
