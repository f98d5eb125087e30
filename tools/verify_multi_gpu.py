"""Multi-GPU parity check (run under torchrun on N >= 2 GPUs of one box):
sharded build (NCCL all-gather of signatures and CSR shards) == single-GPU build, bit for bit on
indices / counts, and identical fp32 distances (same kernels, same summation order per pair).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/verify_multi_gpu.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from schicexplorer_b200 import _capi, synth  # noqa: E402
from schicexplorer_b200.dist import ShardedMinHash, partition_rows_by_nnz  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    api = _capi.cuda_api()
    n, k = int(os.environ.get("VERIFY_CELLS", "900")), 20
    cfg = synth.small_config(n, config_id=31)
    ip, ix, dt, _ = synth.generate_csr(cfg, workers=1)
    ok = True
    # (fast, exchange): signatures only / exact rerank with the NCCL all-gather of the CSR shards / exact rerank with
    # the peer-memory exchange (two builds in a row: both buffer sets of the double buffering are used)
    # "packed": the ranks exchange the packed count stream of the owning rank's preparation pass instead of ids + values
    oracle_ref = {}
    for fast, exchange in ((True, "nccl"), (False, "nccl"), (False, "packed"), (False, "p2p"), (False, "p2p")):
        os.environ["SCHIC_CSR_EXCHANGE"] = "p2p" if exchange == "p2p" else "nccl"
        kw = dict(n_neighbors=k, number_of_hash_functions=800, fast=fast, minimal_blocks_in_common=100,
                  excess_factor=1, max_bin_size=100000)
        bounds = partition_rows_by_nnz(ip, world)
        lo, hi = bounds[rank], bounds[rank + 1]
        a, b = int(ip[lo]), int(ip[hi])
        shard_nnz = [int(ip[bounds[r + 1]] - ip[bounds[r]]) for r in range(world)]
        if exchange == "p2p" and "sh_p2p" in globals():
            sh = globals()["sh_p2p"]                     # second build on the same object
        else:
            sh = ShardedMinHash(api, n, bounds, dev, n_features=cfg.chroms.nbins ** 2, shard_nnz=shard_nnz,
                                integer_counts=(exchange == "packed"), **kw)
            if exchange == "p2p":
                globals()["sh_p2p"] = sh
                assert sh._peer is not None, "peer-memory exchange was not set up"
        sh.fit_local(torch.from_numpy((ip[lo:hi + 1] - ip[lo]).copy()).to(dev),
                     torch.from_numpy(ix[a:b].view(np.int32).copy()).to(dev),
                     None if fast else torch.from_numpy(dt[a:b].copy()).to(dev))
        sh.gather()
        idx, dst, nv = sh.kneighbors_local(k)
        torch.cuda.synchronize()
        # single-GPU reference on every rank (same library, local database)
        h = api.create(device=local, **kw)
        api.fit_csr(h, ip, ix, None if fast else dt, False)
        ridx, rdst, rnv = api.kneighbors(h, k)
        rsig = api.signatures(h, 800)
        same = (np.array_equal(idx.cpu().numpy(), ridx[lo:hi]) and np.array_equal(nv.cpu().numpy(), rnv[lo:hi])
                and np.array_equal(dst.cpu().numpy(), rdst[lo:hi])
                and np.array_equal(sh._keep["sig_all"].cpu().numpy().view(np.uint32), rsig))
        if not fast and exchange == "nccl":      # the exact-kNN sibling shards the same way (brute-force K4)
            eidx, edst, env = sh.kneighbors_exact_local(k)
            torch.cuda.synchronize()
            xidx, xdst, xnv = api.kneighbors_exact(h, k, False)
            same = (same and np.array_equal(eidx.cpu().numpy(), xidx[lo:hi]) and np.array_equal(edst.cpu().numpy(), xdst[lo:hi])
                    and np.array_equal(env.cpu().numpy(), xnv[lo:hi]))
        api.destroy(h)
        if exchange == "packed":
            same = same and api.rerank_kernel_name(sh.h) in ("k4c_rerank_direct_kernel", "k4c_rerank_kernel")
        # ... and against the CPU oracle (rank 0 computes it once per mode, everybody compares its own rows)
        if fast not in oracle_ref:
            from oracle import oracle_api
            ora = oracle_api()
            ho = ora.create(number_of_cores=max(1, (os.cpu_count() or 1) // world), **kw)
            ora.fit_csr(ho, ip, ix, None if fast else dt, False)
            oracle_ref[fast] = ora.kneighbors(ho, k)
            ora.destroy(ho)
        oidx, odst, onv = oracle_ref[fast]
        same = (same and np.array_equal(idx.cpu().numpy(), oidx[lo:hi]) and np.array_equal(nv.cpu().numpy(), onv[lo:hi])
                and np.allclose(dst.cpu().numpy(), odst[lo:hi], rtol=1e-5))
        flag = torch.tensor([int(same)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = ok and bool(flag.item())
        if rank == 0:
            print(f"fast={fast} exchange={exchange}: world={world} rows/rank={[bounds[i+1]-bounds[i] for i in range(world)]} identical={bool(flag.item())}")
        if sh is not globals().get("sh_p2p"):
            sh.close()
    if "sh_p2p" in globals():
        globals()["sh_p2p"].close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_PARITY_OK" if ok else "MULTI_GPU_PARITY_FAILED")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
