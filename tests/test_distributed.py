"""N > 1 host logic on CPU: world_size-2 gloo processes shard a batch, each 'evaluates' its slice, results are
gathered back in batch order and the step time is the max over ranks (what bench.py does with NCCL)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from voxactb_b200 import distributed as vd
from voxactb_b200 import synth


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        obs = synth.make_observation(99, n, 1, 8, 8, low_dim=4)
        mine = vd.shard_observation(obs, rank, world)
        b, e = vd.shard_range(n, rank, world)
        assert mine['proprio'].shape[0] == e - b and mine['rgb'][0].shape[0] == e - b
        assert torch.equal(mine['lang_token_embs'], obs['lang_token_embs'][b:e])
        # stand-in for the per-sample result of the hot path: a row that identifies the sample
        rows = mine['proprio'].sum(1, keepdim=True) + torch.arange(b, e).float().unsqueeze(1)
        allrows = vd.gather_rows(rows, n)
        ref = obs['proprio'].sum(1, keepdim=True) + torch.arange(n).float().unsqueeze(1)
        assert torch.allclose(allrows, ref)
        t = vd.max_over_ranks(10.0 + rank)
        assert t == 10.0 + world - 1
        torch.save(allrows, os.path.join(out_dir, 'rank%d.pt' % rank))
    finally:
        dist.destroy_process_group()


def _grad_worker(rank, world, port, out_dir):
    """DDP-equivalent gradient averaging of the training step (voxactb_b200.train.GradientReducer, the flat-arena reducer of PerActTrainer.update; the reference
    wraps the Q-network in DDP over gloo, agent:50-54): several buckets, ragged sizes, a parameter without a gradient."""
    from voxactb_b200 import train
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(7)
        shapes = [(3, 5), (17,), (2, 2, 2), (1,), (64, 9)]
        params = [torch.nn.Parameter(torch.zeros(*s)) for s in shapes] + [torch.nn.Parameter(torch.zeros(4))]
        base = [torch.randn(*s, generator=g) for s in shapes]
        for p, b in zip(params, base):
            p.grad = b * (rank + 1)                           # rank-dependent gradient; the last parameter has none
        red = train.GradientReducer(params, bucket_bytes=128)   # tiny buckets: several slices of the flat arena, ragged tail
        red.allreduce()
        assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(params[:-1], red.views))   # grads now live in the arena
        mean_factor = sum(r + 1 for r in range(world)) / world
        for p, b in zip(params, base):
            assert torch.allclose(p.grad, b * mean_factor, rtol=1e-6, atol=1e-6)
        assert params[-1].grad is None
        torch.save([p.grad for p in params[:-1]], os.path.join(out_dir, 'grads%d.pt' % rank))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gradient_allreduce(tmp_path):
    world = 2
    mp.spawn(_grad_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    a = torch.load(tmp_path / 'grads0.pt')
    b = torch.load(tmp_path / 'grads1.pt')
    assert all(torch.equal(x, y) for x, y in zip(a, b))      # every rank ends with the same averaged gradients


def test_shard_range_partitions():
    for n in (1, 5, 16, 17, 64):
        for world in (1, 2, 3, 8):
            cuts = [vd.shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [e - b for b, e in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_sharding(tmp_path):
    world, n = 2, 5
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    a = torch.load(tmp_path / 'rank0.pt')
    b = torch.load(tmp_path / 'rank1.pt')
    assert torch.equal(a, b) and a.shape == (n, 1)
