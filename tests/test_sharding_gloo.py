"""CPU, world_size 2 over gloo: the N>1 path of the framework is burst sharding + an output gather.  A stand-in
per-burst function replaces the CUDA forward (no GPU here); the sharded + gathered result must equal the single-process
result bit for bit, including ragged shards, and max_over_ranks must reduce correctly."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deep_rawburst_sr_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _standin_forward(b):          # independent per burst, like the real path
    return (b.mean(dim=(1, 2)) * 3.0 + b.amax(dim=(1, 2, 3, 4)).view(-1, 1, 1)).unsqueeze(1)


def _worker(rank, world, port, total, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    bursts = torch.rand(total, 4, 4, 6, 6, generator=g)
    full = sharding.sharded_forward(_standin_forward, bursts, gather=True)
    mx = sharding.max_over_ranks(float(rank + 1), torch.device('cpu'))
    # the overlapped gatherer (side-stream NCCL on the GPU box) runs the same slot / padding logic synchronously on CPU
    gat = sharding.OutputGatherer(total, depth=2)
    lo, hi = sharding.shard_range(total, rank, world)
    for rep in range(3):     # slots are recycled
        got, ev = gat.submit(_standin_forward(bursts[lo:hi]) + rep)
        assert ev is None and torch.equal(got, _standin_forward(bursts) + rep)
    # int16 predictions (the 14-bit quantised output format; NCCL has no int16 -> the gatherer moves bytes)
    gat16 = sharding.OutputGatherer(total, depth=1)
    q = (_standin_forward(bursts) * 1000).short()
    got16, _ = gat16.submit(q[lo:hi])
    assert got16.dtype == torch.int16 and torch.equal(got16, q)
    # metric scalars: one all_reduce of [sums | counts] instead of gathering images; inf / nan images are dropped
    per_image = _standin_forward(bursts).flatten(1).mean(1)
    per_image[0] = float('inf')
    two = torch.stack([per_image, per_image * 2 + 1], dim=1)
    mean1 = sharding.reduce_metric_means(per_image[lo:hi])
    mean2 = sharding.reduce_metric_means(two[lo:hi])
    assert torch.allclose(mean1, per_image[1:].mean(), rtol=1e-6, atol=0)
    assert mean2.shape == (2,) and torch.allclose(mean2[0], mean1) and torch.allclose(mean2[1], two[1:, 1].mean(), rtol=1e-6, atol=0)
    torch.save({'full': full, 'mx': mx, 'range': sharding.shard_range(total, rank, world)}, os.path.join(out_dir, f'r{rank}.pt'))
    dist.destroy_process_group()


@pytest.mark.parametrize('total', [4, 5])
def test_sharded_forward_matches_single_process(tmp_path, total):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    g = torch.Generator().manual_seed(5)
    ref = _standin_forward(torch.rand(total, 4, 4, 6, 6, generator=g))
    ranges = []
    for r in range(world):
        d = torch.load(os.path.join(tmp_path, f'r{r}.pt'))
        assert torch.equal(d['full'], ref)
        assert d['mx'] == float(world)
        ranges.append(d['range'])
    assert ranges[0][0] == 0 and ranges[-1][1] == total and ranges[0][1] == ranges[1][0]


def test_shard_ranges_cover_everything():
    for total in (1, 7, 16, 33):
        for world in (1, 2, 4, 8):
            r = [sharding.shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert sum(sharding.shard_sizes(total, world)) == total
            assert max(sharding.shard_sizes(total, world)) - min(sharding.shard_sizes(total, world)) <= 1
