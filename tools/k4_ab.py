"""A/B timing of K4 (exact rerank) variants on ONE generated workload, one process: every variant is a (library build,
environment) pair; the data is generated and uploaded once, each variant does warm-up builds and timed builds on the
device-resident inputs and reports the step time, the per-stage CUDA-event times and the parity checksum of its
neighbour lists (all variants must print the same checksum).

  python tools/k4_ab.py [--workload s10] [--steps 4] name[@variant_build][:ENV=V[,ENV=V...]] ...

`@variant_build` loads schicexplorer_b200/csrc/build_<v>/libschic_mh_<v>.so (build.py build_cuda(variant=...)); without it
the product library. Example:
  python tools/k4_ab.py iter:SCHIC_K4_KERNEL=iter direct6 direct4:SCHIC_K4_U=4 t416@t416c2
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="s10")
    ap.add_argument("--cells", type=int, default=None)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("variants", nargs="+")
    args = ap.parse_args()

    import bench
    import torch
    from schicexplorer_b200 import _capi, synth
    from schicexplorer_b200.dist import ShardedMinHash, partition_rows

    cfg = synth.CONFIGS[args.workload]
    if args.cells:
        cfg = synth.SynthConfig(**{**cfg.__dict__, "n_cells": args.cells})
    kw = dict(n_neighbors=args.k, number_of_hash_functions=800, fast=False, minimal_blocks_in_common=100, excess_factor=1,
              max_bin_size=100000, shingle_size=5, hash_width=32)
    ip, ix, dt, _ = synth.generate_csr(cfg, workers=min(32, bench.host_cores()))
    n = cfg.n_cells
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    d_ip = torch.from_numpy(ip).to(dev)
    d_ix = torch.from_numpy(ix.view(np.int32)).to(dev)
    d_dt = torch.from_numpy(dt).to(dev)
    kk = min(args.k, n)
    g_out = (torch.empty(n + 1, dtype=torch.int64, device=dev), torch.empty(n * kk, dtype=torch.int32, device=dev),
             torch.empty(n * kk, dtype=torch.float32, device=dev))
    bounds = partition_rows(n, 1)
    apis = {}
    for spec in args.variants:
        head, _, envs = spec.partition(":")
        name, _, build = head.partition("@")
        env = dict(e.split("=", 1) for e in envs.split(",") if e)
        lib = _capi.CUDA_LIB_PATH if not build else os.path.join(ROOT, "schicexplorer_b200", "csrc", f"build_{build}",
                                                                 f"libschic_mh_{build}.so")
        try:
            if lib not in apis:
                apis[lib] = _capi.CApi(lib)
            api = apis[lib]
            for k_, v_ in env.items():
                os.environ[k_] = v_
            sh = ShardedMinHash(api, n, bounds, dev, n_features=cfg.chroms.nbins ** 2, shard_nnz=[int(ip[-1])],
                                integer_counts=True, **kw)

            def step():
                sh.fit_local(d_ip, d_ix, d_dt, indptr0=0)
                sh.gather()
                sh.kneighbors_graph_local(kk, g_out)

            for _ in range(args.warmup):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            stage = api.stage_ms(sh.h)
            kern = api.rerank_kernel_name(sh.h)
            idx_t, dst_t, _nv = sh.kneighbors_local(kk)
            torch.cuda.synchronize()
            cs = bench.fold_checksum(np.arange(n), idx_t.cpu().numpy(), dst_t.cpu().numpy().view(np.int32))
            sh.close()
            print(json.dumps({"variant": name, "build": build or "product", "env": env, "ms_per_step": round(ms, 3),
                              "rerank_ms": round(stage.get("rerank", 0.0), 3), "kernel": kern, "checksum": cs,
                              "stage_ms": {k_: round(v_, 3) for k_, v_ in stage.items() if v_}}), flush=True)
        except Exception as e:  # a variant that does not load / launch must not end the sweep
            print(json.dumps({"variant": name, "build": build or "product", "env": env, "error": repr(e)[:300]}), flush=True)
        finally:
            for k_ in env:
                os.environ.pop(k_, None)


if __name__ == "__main__":
    main()
