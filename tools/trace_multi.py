"""GPU timeline of a few sharded builds on rank 0 (torch profiler / CUPTI, no nsys needed). torchrun, N >= 2."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from schicexplorer_b200 import _capi, synth  # noqa: E402
from schicexplorer_b200.dist import ShardedMinHash, partition_rows  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
api = _capi.cuda_api()
cfg = synth.CONFIGS["s10"]
bounds = partition_rows(cfg.n_cells, world)
lo, hi = bounds[rank], bounds[rank + 1]
ip, ix, dt, _ = synth.generate_csr(cfg, lo, hi, workers=max(1, 16 // world))
d_ip, d_ix, d_dt = torch.from_numpy(ip).to(dev), torch.from_numpy(ix.view(np.int32)).to(dev), torch.from_numpy(dt).to(dev)
kw = dict(n_neighbors=100, number_of_hash_functions=800, fast=False, minimal_blocks_in_common=100, excess_factor=1,
          max_bin_size=100000, shingle_size=5)
t = torch.zeros(world, dtype=torch.int64, device=dev)
dist.all_gather_into_tensor(t, torch.tensor([ix.size], dtype=torch.int64, device=dev))
sh = ShardedMinHash(api, cfg.n_cells, bounds, dev, n_features=cfg.chroms.nbins ** 2, shard_nnz=t.tolist(), **kw)
k = 100
out = (torch.empty((hi - lo, k), dtype=torch.int32, device=dev), torch.empty((hi - lo, k), dtype=torch.float32, device=dev),
       torch.empty(hi - lo, dtype=torch.int32, device=dev))


def step():
    sh.fit_local(d_ip, d_ix, d_dt)
    sh.gather()
    sh.kneighbors_local(k, out)


for _ in range(4):
    step()
torch.cuda.synchronize(); dist.barrier()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    print(f"{'start_us':>10} {'dur_us':>9} {'gap_us':>8}  name")
    prev_end = t0
    for e in evs:
        s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
        print(f"{s:10.0f} {d:9.0f} {e.time_range.start - prev_end:8.0f}  {e.name[:90]}")
        prev_end = max(prev_end, e.time_range.end)
sh.close()
dist.barrier()
dist.destroy_process_group()
