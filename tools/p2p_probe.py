"""Probe: copy speed into a peer process's memory mapped through CUDA IPC (torchrun, 2 ranks)."""
import ctypes
import os

import torch
import torch.distributed as dist
from torch.multiprocessing.reductions import reduce_tensor

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
n = 256 << 20
mine = torch.empty(n, dtype=torch.uint8, device=dev)
src = torch.empty(n, dtype=torch.uint8, device=dev)
objs = [None] * world
dist.all_gather_object(objs, reduce_tensor(mine))
peer = 1 - rank
f, a = objs[peer]
pt = f(*a)
print(rank, "peer tensor device", pt.device, "can access", torch.cuda.can_device_access_peer(rank, peer), flush=True)


def timeit(name, fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(rank); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record(); torch.cuda.synchronize(rank)
    ms = e0.elapsed_time(e1) / 10
    if rank == 0:
        print(f"{name}: {ms:.3f} ms = {n/ms/1e6:.0f} GB/s", flush=True)


timeit("torch copy_ into IPC peer tensor", lambda: pt.copy_(src, non_blocking=True))
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
s = torch.cuda.current_stream(dev).cuda_stream
timeit("cudaMemcpyAsync(default kind) into IPC peer pointer", lambda: rt.cudaMemcpyAsync(pt.data_ptr(), src.data_ptr(), n, 4, s))
rt.cudaDeviceEnablePeerAccess.argtypes = [ctypes.c_int, ctypes.c_uint]
print(rank, "enable peer access rc", rt.cudaDeviceEnablePeerAccess(peer, 0), flush=True)
timeit("cudaMemcpyAsync after cudaDeviceEnablePeerAccess", lambda: rt.cudaMemcpyAsync(pt.data_ptr(), src.data_ptr(), n, 4, s))
timeit("torch copy_ after enable", lambda: pt.copy_(src, non_blocking=True))
dist.barrier()
del pt
dist.destroy_process_group()
