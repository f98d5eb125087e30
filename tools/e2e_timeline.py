"""Where does the end-to-end (host buffers in, lists out) step spend its wall time? Times each host call of
bench.py's e2e step, once with a device synchronize after every call (exclusive times) and once without."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from schicexplorer_b200 import _capi, synth  # noqa: E402
from schicexplorer_b200.dist import ShardedMinHash, partition_rows  # noqa: E402

api = _capi.cuda_api()
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
cfg = synth.CONFIGS["s10"]
n = int(sys.argv[1]) if len(sys.argv) > 1 else cfg.n_cells
cfg = synth.SynthConfig(**{**cfg.__dict__, "n_cells": n})
ip, ix, dt, _ = synth.generate_csr(cfg, 0, n, workers=16)
kw = dict(n_neighbors=100, number_of_hash_functions=800, fast=False, minimal_blocks_in_common=100, excess_factor=1,
          max_bin_size=100000, shingle_size=5)
p_ip = api.pinned_empty(ip.shape, np.int64); p_ip[:] = ip
p_ix = api.pinned_empty(ix.shape, np.uint32); p_ix[:] = ix
p_dt = api.pinned_empty(dt.shape, np.float32); p_dt[:] = dt
k = 100
out = (torch.empty((n, k), dtype=torch.int32, device=dev), torch.empty((n, k), dtype=torch.float32, device=dev),
       torch.empty(n, dtype=torch.int32, device=dev))
h_idx = torch.empty((n, k), dtype=torch.int32).pin_memory()
sh = ShardedMinHash(api, n, partition_rows(n, 1), dev, **kw)


def step(sync_each):
    t = [time.perf_counter()]

    def mark():
        if sync_each:
            torch.cuda.synchronize()
        t.append(time.perf_counter())
    sh.fit_local_host(p_ip, p_ix, p_dt); mark()
    sh.gather(); mark()
    sh.kneighbors_local(k, out); mark()
    h_idx.copy_(out[0], non_blocking=True); mark()
    torch.cuda.synchronize(); mark()
    return [round((b - a) * 1e3, 2) for a, b in zip(t[:-1], t[1:])], round((t[-1] - t[0]) * 1e3, 2)


for _ in range(2):
    step(False)
print("calls: fit_local_host, gather, kneighbors_local, d2h copy, final sync")
for s in (True, False, False):
    print("sync_each" if s else "async    ", *step(s), api.stage_ms(sh.h))
