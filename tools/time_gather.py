"""How long does the CSR exchange take on its own? (run under torchrun, N >= 2)"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

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
sh._keep["csr"] = (d_ip, d_ix, d_dt)
total_bytes = int(t.sum().item()) * 8
for name, fn in (("gather_csr (2 padded all-gathers + compaction)", sh._gather_csr),):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    if rank == 0:
        print(f"{name}: {ms:.3f} ms for {total_bytes/1e9:.3f} GB gathered -> {total_bytes/ms/1e6:.0f} GB/s algbw, "
              f"busbw {total_bytes*(world-1)/world/ms/1e6:.0f} GB/s")
# raw NCCL all-gather of the same volume, equal shards
n_el = int(t.max().item())
a = torch.empty(n_el, dtype=torch.int32, device=dev)
b = torch.empty(world * n_el, dtype=torch.int32, device=dev)
for _ in range(2):
    dist.all_gather_into_tensor(b, a)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    dist.all_gather_into_tensor(b, a)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
if rank == 0:
    print(f"raw all_gather_into_tensor {world}x{n_el*4/1e6:.1f} MB: {ms:.3f} ms, busbw {world*n_el*4*(world-1)/world/ms/1e6:.0f} GB/s")
sh.close()
dist.barrier()
dist.destroy_process_group()
