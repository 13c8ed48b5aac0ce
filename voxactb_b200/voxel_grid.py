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
