"""2-GPU check of the overlapped output gather with RAGGED shards over NCCL (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/gather_ragged_check.py

total = 5 bursts over 2 ranks (3 + 2): every rank submits its shard `depth + 3` times with different contents while a long
kernel keeps the current stream busy, and checks -- after waiting on the returned event only -- that the gathered batch is
the full batch of THAT submission (a compaction that ran before its collective would return the previous contents)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_rawburst_sr_b200 import sharding  # noqa: E402

rank, local = int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0'))
dev = torch.device('cuda', local)
torch.cuda.set_device(dev)
dist.init_process_group('nccl', device_id=dev)
world = dist.get_world_size()
total = 2 * world + 1
lo, hi = sharding.shard_range(total, rank, world)
g = torch.Generator().manual_seed(3)
full = torch.rand(total, 3, 384, 384, generator=g)
gat = sharding.OutputGatherer(total, depth=2)
busy = torch.empty(64 << 20, device=dev)
bad = 0
for rep in range(5):
    local_pred = (full[lo:hi] + rep).to(dev)
    torch.cuda.synchronize()
    busy.normal_()                                  # the current stream is busy while the gather runs on the side stream
    got, ev = gat.submit(local_pred)
    ev.synchronize()
    if not torch.equal(got.cpu(), full + rep):
        bad += 1
t = torch.tensor([bad], device=dev)
dist.all_reduce(t)
if rank == 0:
    print(json.dumps({'world': world, 'total': total, 'sizes': gat.sizes, 'mismatching_submissions': int(t.item())}))
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if int(t.item()) else 0)
