"""B200-native ``VoxelGrid`` -- host-side mirror of reference peract/voxel/voxel_grid.py:15-198.

Same constructor and ``coords_to_bounding_voxel_grid(coords, coord_features, coord_bounds)``
contract; the scatter-mean, crop, occupancy and index-grid channels are produced by
``vxb_voxelize_f32`` (voxactb_b200/csrc/voxelize.cu).  Differences that are deliberate:
  * no per-(B,N) helper buffers are materialised (the reference registers ~12, 12.7 MB x B for the
    index grid alone); ``state_dict`` therefore has no ``_voxelizer.*`` tensors, which the
    agent's ``load_weights`` ignores by design (qattention_peract_bc_agent.py:850-854);
  * the runtime batch may differ from the constructor's ``batch_size`` (the reference fails on
    the ``cat`` at voxel_grid.py:172-173 in that case).
"""
import ctypes

import torch
from torch import nn

from . import _lib


class VoxelGrid(nn.Module):

    def __init__(self, coord_bounds, voxel_size: int, device, batch_size, feature_size,
                 max_num_coords: int):
        super().__init__()
        self._device = device
        self._voxel_size = int(voxel_size)
        self._voxel_shape = [self._voxel_size] * 3
        self._voxel_d = float(self._voxel_size)
        self._voxel_feature_size = 4 + feature_size
        self._feature_size = int(feature_size)
        self._batch_size = batch_size
        self._num_coords = int(max_num_coords)
        self.register_buffer('_coord_bounds',
                             torch.tensor(coord_bounds, dtype=torch.float).reshape(1, 6),
                             persistent=False)
        self._workspace = None
        self.last_indices = None

    def __getstate__(self):
        state = self.__dict__.copy()
        state['_workspace'] = None
        state['last_indices'] = None
        return state

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__getstate__().items():
            setattr(new, k, copy.deepcopy(v, memo))
        return new

    def coords_to_bounding_voxel_grid(self, coords, coord_features=None, coord_bounds=None,
                                      return_indices=False):
        """coords [B,N,3], coord_features [B,N,F], coord_bounds [1|B,6] -> [B,V,V,V,3+F+3+1]
        (reference voxel_grid.py:148-198).  ``return_indices`` additionally returns the clamped
        (V+2)-grid int32 index of every point (voxel_grid.py:159-163) for the bit-exact test."""
        if not coords.is_cuda:
            raise RuntimeError('voxactb_b200.VoxelGrid runs on CUDA only (no CPU fallback); got %s' % coords.device)
        coords = _lib.f32(coords)
        B, N, three = coords.shape
        if three != 3:
            raise ValueError('coords must be [B,N,3]')
        F = 0
        feats = None
        if coord_features is not None:
            feats = _lib.f32(coord_features)
            F = feats.shape[-1]
            if feats.shape[:2] != (B, N):
                raise ValueError('coord_features must be [B,N,F]')
        if F != self._feature_size:
            raise ValueError('feature size %d != constructor feature_size %d' % (F, self._feature_size))
        bounds = self._coord_bounds if coord_bounds is None else coord_bounds
        bounds = _lib.f32(bounds.to(coords.device).reshape(-1, 6))
        Bb = bounds.shape[0]
        if Bb not in (1, B):
            raise ValueError('coord_bounds batch must be 1 or %d, got %d' % (B, Bb))
        V = self._voxel_size
        L = _lib.lib()
        ws_bytes = L.vxb_voxelize_workspace_bytes(B, N, V, F)
        if self._workspace is None or self._workspace.numel() < ws_bytes or self._workspace.device != coords.device:
            self._workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=coords.device)
        out = torch.empty(B, V, V, V, 7 + F, dtype=torch.float32, device=coords.device)
        idx = torch.empty(B, N, 3, dtype=torch.int32, device=coords.device) if return_indices else None
        rc = L.vxb_voxelize_f32(_lib.ptr(coords), _lib.ptr(feats), _lib.ptr(bounds), Bb, B, N, F, V,
                                _lib.ptr(out), 0, _lib.ptr(idx), _lib.ptr(self._workspace), ws_bytes,
                                _lib.stream())
        _lib.check(rc, 'vxb_voxelize_f32')
        if return_indices:
            return out, idx
        return out

    def depth_to_bounding_voxel_grid(self, depth, intrinsics, extrinsics, rgb=None, coord_bounds=None, return_points=False,
                                     return_indices=False):
        """Raw-depth entry: depth [B,cams,H,W] (CUDA fp32, metres), intrinsics [B,cams,3,3] and
        extrinsics [B,cams,4,4] (HOST arrays / CPU tensors, as they sit in every observation, launch_utils.py:84-87), rgb
        [B,cams,F,H,W] (CUDA fp32, already normalised) -> the same [B,V,V,V,3+F+3+1] grid coords_to_bounding_voxel_grid
        gives for the reference's host-side back-projection (PyRep vision_sensor.py:155-175) of those images, with the
        back-projection fused into the scatter kernel.  The per-camera 4x4 inverse stays on the host in float64, as in the
        reference."""
        import numpy as np
        if not depth.is_cuda:
            raise RuntimeError('voxactb_b200.VoxelGrid runs on CUDA only (no CPU fallback); got %s' % depth.device)
        depth = _lib.f32(depth)
        B, cams, H, W = depth.shape
        Kh = np.asarray(intrinsics.cpu() if torch.is_tensor(intrinsics) else intrinsics, dtype=np.float64).reshape(B, cams, 3, 3)
        Eh = np.asarray(extrinsics.cpu() if torch.is_tensor(extrinsics) else extrinsics, dtype=np.float64).reshape(B, cams, 4, 4)
        minv = np.empty((B, cams, 3, 4), dtype=np.float64)
        for b in range(B):
            for c in range(cams):
                # vision_sensor.py:166-172
                Cc = np.expand_dims(Eh[b, c, :3, 3], 0).T
                R_inv = Eh[b, c, :3, :3].T
                ext = np.concatenate((R_inv, -np.matmul(R_inv, Cc)), -1)
                homo = np.concatenate([np.matmul(Kh[b, c], ext), [np.array([0, 0, 0, 1])]])
                minv[b, c] = np.linalg.inv(homo)[0:3]
        minv_d = torch.from_numpy(minv).to(depth.device)
        F = 0
        if rgb is not None:
            rgb = _lib.f32(rgb)
            F = rgb.shape[2]
            if rgb.shape != (B, cams, F, H, W):
                raise ValueError('rgb must be [B,cams,F,H,W]')
        if F != self._feature_size:
            raise ValueError('feature size %d != constructor feature_size %d' % (F, self._feature_size))
        bounds = self._coord_bounds if coord_bounds is None else coord_bounds
        bounds = _lib.f32(bounds.to(depth.device).reshape(-1, 6))
        Bb = bounds.shape[0]
        if Bb not in (1, B):
            raise ValueError('coord_bounds batch must be 1 or %d, got %d' % (B, Bb))
        V, N = self._voxel_size, cams * H * W
        L = _lib.lib()
        ws_bytes = L.vxb_voxelize_workspace_bytes(B, N, V, F)
        if self._workspace is None or self._workspace.numel() < ws_bytes or self._workspace.device != depth.device:
            self._workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=depth.device)
        out = torch.empty(B, V, V, V, 7 + F, dtype=torch.float32, device=depth.device)
        pts = torch.empty(B, N, 3, dtype=torch.float32, device=depth.device) if return_points else None
        idx = torch.empty(B, N, 3, dtype=torch.int32, device=depth.device) if return_indices else None
        rc = L.vxb_voxelize_depth_f32(_lib.ptr(depth), _lib.ptr(minv_d), _lib.ptr(rgb), _lib.ptr(bounds), Bb, B, cams, H, W, F, V,
                                      _lib.ptr(out), 0, _lib.ptr(pts), _lib.ptr(idx), _lib.ptr(self._workspace), ws_bytes,
                                      _lib.stream())
        _lib.check(rc, 'vxb_voxelize_depth_f32')
        res = (out,) + ((pts,) if return_points else ()) + ((idx,) if return_indices else ())
        return res if len(res) > 1 else out
