"""Sharded SyntheticBurst scoring over NCCL (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/eval_sharded.py

Every rank scores its contiguous shard of a seeded in-memory validation set (37 bursts: ragged shards) and the ranks
all-reduce `[sums | counts]`; rank 0 also scores the whole set alone and prints both reports (they must agree)."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_rawburst_sr_b200.evaluation.synburst.compute_score import TensorBurstSet, score_dataset  # noqa: E402
from deep_rawburst_sr_b200.models.dbsr.dbsrnet import dbsrnet_default_synthetic  # noqa: E402

rank, local = int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0'))
dev = torch.device('cuda', local)
torch.cuda.set_device(dev)
dist.init_process_group('nccl', device_id=dev)
torch.manual_seed(0)
net = dbsrnet_default_synthetic().to(dev).eval().set_precision('bf16')
g = torch.Generator().manual_seed(5)
n = 37
data = TensorBurstSet(torch.rand(n, 14, 4, 48, 48, generator=g).pin_memory(), torch.rand(n, 3, 384, 384, generator=g).pin_memory())
score_dataset(net, data, batch_size=8, device=dev)            # warm-up
torch.cuda.synchronize()
dist.barrier()
t0 = time.perf_counter()
sharded = score_dataset(net, data, batch_size=8, device=dev)
dt = time.perf_counter() - t0
if rank == 0:
    alone = score_dataset(net, data, batch_size=8, device=dev, shard=False)
    ok = abs(alone['psnr'] - sharded['psnr']) <= 1e-4 and abs(alone['ssim'] - sharded['ssim']) <= 1e-6
    print(json.dumps({'world': dist.get_world_size(), 'bursts': n, 'sharded': sharded, 'single_rank': alone, 'agree': ok,
                      'sharded_s': dt, 'bursts_per_s': n / dt}))
    assert ok
dist.barrier()
dist.destroy_process_group()
