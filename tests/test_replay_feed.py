"""Row f3: the replay -> device feeder returns exactly the sampled batches, in order."""
import numpy as np
import pytest
import torch

from voxactb_b200.replay_feed import DeviceFeeder


def batches(n, seed=0):
    rng = np.random.default_rng(seed)
    for i in range(n):
        yield {'front_rgb': rng.integers(0, 255, (4, 3, 16, 16), dtype=np.uint8),
               'front_point_cloud': rng.normal(size=(4, 3, 16, 16)).astype(np.float32),
               'low_dim_state': torch.from_numpy(rng.normal(size=(4, 4)).astype(np.float32)),
               'task': 'open_drawer', 'index': i}


def test_feeder_preserves_batches_on_cpu():
    want = list(batches(7))
    got = list(DeviceFeeder(batches(7), 'cpu', depth=2))
    assert len(got) == 7
    for w, g in zip(want, got):
        assert g['task'] == w['task'] and g['index'] == w['index']
        assert g['front_rgb'].dtype == torch.uint8 and np.array_equal(g['front_rgb'].numpy(), w['front_rgb'])
        assert torch.equal(g['low_dim_state'], w['low_dim_state'])


def test_feeder_propagates_iterator_errors():
    def bad():
        yield {'x': np.zeros(3, np.float32)}
        raise ValueError('replay buffer exhausted')
    f = DeviceFeeder(bad(), 'cpu')
    next(f)
    with pytest.raises(ValueError):
        next(f)


@pytest.mark.gpu
def test_feeder_on_cuda_overlaps_and_reuses_buffers(cuda_lib):
    want = list(batches(9, seed=3))
    feeder = DeviceFeeder(batches(9, seed=3), 'cuda', depth=2)
    sums = []
    for w, g in zip(want, feeder):
        assert g['front_point_cloud'].is_cuda and g['front_rgb'].dtype == torch.uint8
        # a consumer kernel on the compute stream; the slot is only recycled after it
        sums.append((g['front_point_cloud'].double().sum() + g['front_rgb'].double().sum(), w))
        torch.cuda._sleep(2_000_000)
    for s, w in sums:
        assert abs(float(s) - (w['front_point_cloud'].astype(np.float64).sum() + w['front_rgb'].astype(np.float64).sum())) < 1e-6
