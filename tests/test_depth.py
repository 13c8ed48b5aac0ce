"""Row f1: depth -> world point cloud back-projection fused into the voxelizer.
CPU: the numpy oracle against the fixture produced by the reference's own PyRep functions.  GPU: vxb_voxelize_depth_f32
against the oracle -- back-projected points within 1 fp32 ulp-class tolerance, voxel indices bit-exact, identical grid."""
import numpy as np
import pytest
import torch

import util
from oracle import depth_oracle, voxel_oracle
from voxactb_b200 import synth

import make_golden


def oracle_case(c):
    o = synth.make_depth_observation(c['seed'], c['B'], c['cameras'], c['H'], c['W'])
    pts = depth_oracle.batch_pointcloud(o['depth'].numpy(), o['extrinsics'], o['intrinsics'])
    feats = o['rgb'].permute(0, 1, 3, 4, 2).reshape(c['B'], -1, 3).numpy()
    return o, pts, feats


def test_depth_oracle_matches_reference_fixture():
    c = make_golden.DEPTH_CASES['depth_v20']
    g = util.golden('depth_v20')
    o, pts, feats = oracle_case(c)
    assert np.array_equal(pts, g['points'])                      # same numpy operations in the same order: bit-identical
    idx = voxel_oracle.voxel_indices(pts, o['bounds'].numpy(), c['V'])
    assert np.array_equal(idx, g['idx'].astype(np.int32))
    grid = voxel_oracle.voxelize(pts, feats, o['bounds'].numpy(), c['V'])
    np.testing.assert_allclose(grid, g['grid'], rtol=2e-6, atol=2e-6)


@pytest.mark.gpu
@pytest.mark.parametrize('case', ['depth_v20', 'full'])
def test_voxelize_depth_matches_oracle(cuda_lib, case):
    from voxactb_b200 import VoxelGrid
    c = make_golden.DEPTH_CASES['depth_v20'] if case == 'depth_v20' else dict(V=100, B=2, cameras=4, H=128, W=128, seed=52)
    o, pts, feats = oracle_case(c)
    vg = VoxelGrid(synth.SCENE_BOUNDS, c['V'], 'cuda', c['B'], 3, c['cameras'] * c['H'] * c['W'])
    grid, gp, idx = vg.depth_to_bounding_voxel_grid(o['depth'].cuda(), o['intrinsics'], o['extrinsics'], o['rgb'].cuda(),
                                                    o['bounds'].cuda(), return_points=True, return_indices=True)
    torch.cuda.synchronize()
    gp, idx, grid = gp.cpu().numpy(), idx.cpu().numpy(), grid.cpu().numpy()
    # float64 dot product rounded to fp32: the summation order inside numpy's matmul is not specified -> allow 1 ulp
    ulp = np.spacing(np.abs(pts).astype(np.float32))
    assert np.all(np.abs(gp - pts) <= ulp), float(np.abs(gp - pts).max())
    same = np.all(gp == pts, axis=-1)
    assert same.mean() > 0.999
    ref_idx = voxel_oracle.voxel_indices(gp, o['bounds'].numpy(), c['V'])      # indices of OUR points: bit-exact arithmetic
    assert np.array_equal(idx, ref_idx)
    assert np.array_equal(idx[same], voxel_oracle.voxel_indices(pts, o['bounds'].numpy(), c['V'])[same])
    ref = voxel_oracle.voxelize(gp, feats, o['bounds'].numpy(), c['V'])
    assert np.array_equal(grid[..., -4:], ref[..., -4:])
    np.testing.assert_allclose(grid[..., :-4], ref[..., :-4], rtol=2e-6, atol=2e-6)
    # and equal to the two-step path (host back-projection + vxb_voxelize_f32) on the same points
    two, _ = vg.coords_to_bounding_voxel_grid(torch.from_numpy(gp).cuda(), torch.from_numpy(feats).cuda(), o['bounds'].cuda(),
                                              return_indices=True)
    assert np.array_equal(two.cpu().numpy()[..., -4:], grid[..., -4:])
