"""K1 voxelizer on the GPU (through the C ABI) against the CPU oracle and the reference goldens."""
import numpy as np
import pytest
import torch

import util
from oracle import voxel_oracle
from voxactb_b200 import VoxelGrid, synth

import make_golden

pytestmark = pytest.mark.gpu


def run_cuda(coords, feats, bounds, V):
    vg = VoxelGrid(synth.SCENE_BOUNDS, V, 'cuda', coords.shape[0], 0 if feats is None else feats.shape[-1],
                   coords.shape[1])
    grid, idx = vg.coords_to_bounding_voxel_grid(coords.cuda(), None if feats is None else feats.cuda(),
                                                 None if bounds is None else bounds.cuda(), return_indices=True)
    torch.cuda.synchronize()
    return grid.cpu().numpy(), idx.cpu().numpy()


def compare(grid, idx, coords, feats, bounds, V):
    ref_idx = voxel_oracle.voxel_indices(coords.numpy(), bounds.numpy(), V)
    ref = voxel_oracle.voxelize(coords.numpy(), None if feats is None else feats.numpy(), bounds.numpy(), V)
    assert np.array_equal(idx, ref_idx), 'voxel indices must be bit-exact'
    assert np.array_equal(grid[..., -1], ref[..., -1]), 'occupancy must be exact'
    assert np.array_equal(grid[..., -4:-1], ref[..., -4:-1]), 'index-grid channels must be exact'
    # means: fp32 sums in a different order (atomics) -> tolerance, not bit equality
    np.testing.assert_allclose(grid[..., :-4], ref[..., :-4], rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize('name', ['voxel_v20', 'voxel_v32_crop', 'voxel_v100'])
def test_voxelize_matches_oracle_and_golden(cuda_lib, name):
    c = make_golden.VOXEL_CASES[name]
    g = util.golden(name)
    obs = synth.make_observation(c['seed'], c['B'], c['cameras'], c['H'], c['W'], per_sample_crop=c['crop'])
    coords, feats = synth.flatten_cameras(obs)
    grid, idx = run_cuda(coords, feats, obs['bounds'], c['V'])
    compare(grid, idx, coords, feats, obs['bounds'], c['V'])
    assert int(idx.astype(np.int64).sum()) == int(g['idx_checksum'][0])
    assert int((grid[..., -1] > 0).sum()) == int(g['occupied'][0])
    if 'grid' in g.files:
        assert np.array_equal(idx, g['idx'].astype(np.int32))
        np.testing.assert_allclose(grid, g['grid'], rtol=2e-6, atol=2e-6)
    else:
        flat = grid.reshape(-1, grid.shape[-1])
        np.testing.assert_allclose(flat[g['sample_pos']], g['sample_val'], rtol=2e-6, atol=2e-6)


def test_voxelize_edge_cases(cuda_lib):
    gen = torch.Generator().manual_seed(3)
    bounds = torch.tensor([[-1., -1., -1., 1., 1., 1.]])
    # ragged N (not a multiple of the block), points exactly on the bounds, far outside, duplicates
    coords = torch.rand(2, 1000 + 37, 3, generator=gen) * 2.4 - 1.2
    coords[0, :8] = torch.tensor([[-1., -1., -1.], [1., 1., 1.], [0., 0., 0.], [1., -1., 0.],
                                  [0.999999, 0.999999, 0.999999], [-1.000001, 0., 0.], [50., 50., 50.],
                                  [-50., 0.3, 0.3]])
    coords[1, 100:400] = coords[1, 100]          # 300 points in one voxel (contention)
    feats = torch.rand(2, coords.shape[1], 3, generator=gen)
    for V in (1, 7, 16):
        grid, idx = run_cuda(coords, feats, bounds, V)
        compare(grid, idx, coords, feats, bounds, V)
    # everything out of bounds -> empty grid except the index channels
    far = torch.full((1, 64, 3), 9.0)
    grid, idx = run_cuda(far, torch.ones(1, 64, 3), bounds, 8)
    assert grid[..., :6].sum() == 0 and grid[..., -1].sum() == 0
    compare(grid, idx, far, torch.ones(1, 64, 3), bounds, 8)
    # single point, no features (F=0 path)
    one = torch.tensor([[[0.1, 0.2, 0.3]]])
    grid, idx = run_cuda(one, None, bounds, 10)
    compare(grid, idx, one, None, bounds, 10)
    # per-sample bounds
    b2 = torch.tensor([[-1., -1., -1., 1., 1., 1.], [-0.5, -0.2, 0., 0.7, 0.9, 1.]])
    grid, idx = run_cuda(coords, feats, b2, 12)
    compare(grid, idx, coords, feats, b2, 12)


def test_voxelize_non_finite_and_huge_coordinates(cuda_lib):
    """|q| >= 2^31, +-inf and NaN coordinates.  The reference converts floor(q) with `.int()` (voxel_grid.py:160-163):
    on CUDA (where train.py / eval.py run it) that conversion saturates (+huge -> V+1 after the min, -huge and NaN -> 0),
    on x86 every invalid conversion gives INT_MIN -> 0.  Both cells are border cells that voxel_grid.py:184 crops, so
    the GRID is identical either way; the kernel follows the CUDA semantics for the raw index."""
    gen = torch.Generator().manual_seed(5)
    bounds = torch.tensor([[-1., -1., -1., 1., 1., 1.]])
    V = 16
    coords = torch.rand(1, 512, 3, generator=gen) * 2 - 1
    feats = torch.rand(1, 512, 3, generator=gen)
    clean_grid, _ = run_cuda(coords[:, 8:].contiguous(), feats[:, 8:].contiguous(), bounds, V)
    inf, nan = float('inf'), float('nan')
    bad = torch.tensor([[1e30, 0., 0.], [-1e30, 0., 0.], [0., inf, 0.], [0., -inf, 0.], [0., 0., nan],
                        [3e9, -3e9, 0.], [nan, nan, nan], [inf, -inf, nan]])
    coords[0, :8] = bad
    grid, idx = run_cuda(coords, feats, bounds, V)
    expect = torch.tensor([[V + 1, -1, -1], [0, -1, -1], [-1, V + 1, -1], [-1, 0, -1], [-1, -1, 0],
                           [V + 1, 0, -1], [0, 0, 0], [V + 1, 0, 0]])
    got = torch.from_numpy(idx[0, :8]).long()
    assert torch.equal(got[expect >= 0], expect[expect >= 0])
    # every such point falls in a cropped border cell: the grid is the one of the remaining points
    assert np.isfinite(grid).all()
    assert np.array_equal(grid[..., -4:], clean_grid[..., -4:])
    np.testing.assert_allclose(grid[..., :-4], clean_grid[..., :-4], rtol=2e-6, atol=2e-6)


def test_voxelize_full_size_properties(cuda_lib):
    """BASELINE.json size (B=16, V=100, 4 cameras): size-independent properties."""
    B, V = 16, 100
    obs = synth.make_observation(99, B, 4, 128, 128, per_sample_crop=True)
    coords, feats = synth.flatten_cameras(obs)
    vg = VoxelGrid(synth.SCENE_BOUNDS, V, 'cuda', B, 3, coords.shape[1])
    cc, ff, bb = coords.cuda(), feats.cuda(), obs['bounds'].cuda()
    grid, idx = vg.coords_to_bounding_voxel_grid(cc, ff, bb, return_indices=True)
    # (1) indices bit-exact against the oracle at full size (index math is cheap on the CPU)
    assert np.array_equal(idx.cpu().numpy(), voxel_oracle.voxel_indices(coords.numpy(), obs['bounds'].numpy(), V))
    # (2) occupancy == set of in-range voxel ids
    inr = ((idx >= 1) & (idx <= V)).all(-1)
    flat = ((idx[..., 0] - 1) * V + (idx[..., 1] - 1)) * V + (idx[..., 2] - 1)
    for b in range(B):
        occ = torch.zeros(V ** 3, device='cuda')
        occ[flat[b][inr[b]].long()] = 1
        assert torch.equal(occ, grid[b, ..., -1].reshape(-1))
    # (3) mean xyz of an occupied voxel lies inside that voxel's cell (+- fp slack)
    res = (bb[:, 3:] - bb[:, :3]) / V
    lo = bb[:, None, None, None, :3] + grid[..., 6:9] * V * res[:, None, None, None, :]
    occm = grid[..., -1] > 0
    eps = 1e-4
    assert bool(((grid[..., :3] >= lo - eps) & (grid[..., :3] <= lo + res[:, None, None, None, :] + eps))[occm].all())
    # (4) permutation of the points leaves the grid unchanged up to fp32 summation order
    perm = torch.randperm(coords.shape[1], device='cuda')
    grid2 = vg.coords_to_bounding_voxel_grid(cc[:, perm].contiguous(), ff[:, perm].contiguous(), bb)
    assert torch.equal(grid2[..., 6:], grid[..., 6:])
    torch.testing.assert_close(grid2[..., :6], grid[..., :6], rtol=2e-6, atol=2e-6)
    # (5) batch position does not matter
    g0 = vg.coords_to_bounding_voxel_grid(cc[3:4].contiguous(), ff[3:4].contiguous(), bb[3:4].contiguous())
    assert torch.equal(g0[0, ..., 6:], grid[3, ..., 6:])
    torch.testing.assert_close(g0[0, ..., :6], grid[3, ..., :6], rtol=2e-6, atol=2e-6)
