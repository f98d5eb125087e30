"""Dense-graph PCA post-step: device time by stage for the graph of n cells (k = n, fast distances)
against scikit-learn's PCA(n_components=n-1).fit_transform on the host (the reference's call,
scHicClusterMinHash.py:358-366) on a bounded size. One JSON object on stdout.

usage: python tools/bench_pca.py [n_cells ...]      (default 2000 10000; sklearn is timed on sizes <= 4000)"""
import json
import os
import sys
import time

import numpy as np
from scipy.sparse import csr_matrix

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from schicexplorer_b200 import MinHash, _capi, synth  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [2000, 10000]
api = _capi.cuda_api()
out = {"dim": 100, "runs": []}
for n in sizes:
    cfg = synth.SynthConfig(**{**synth.CONFIGS["s10"].__dict__, "n_cells": n})
    ip, ix, dt, _ = synth.generate_csr(cfg, 0, n, workers=min(32, os.cpu_count() or 1))
    X = synth.to_scipy(cfg, ip, ix, dt)
    mh = MinHash(n_neighbors=n, number_of_hash_functions=800, fast=True, minimal_blocks_in_common=100, excess_factor=1,
                 max_bin_size=100000, shingle_size=5)
    t0 = time.perf_counter()
    G = mh.fit(X).kneighbors_graph(mode="distance")
    t_graph = time.perf_counter() - t0
    mh.close()
    G = csr_matrix((G.data.astype(np.float32), G.indices, G.indptr), shape=G.shape)
    run = {"n_cells": n, "graph_nnz": int(G.nnz), "graph_build_ms": round(t_graph * 1e3, 1)}
    for rep in range(2):
        t0 = time.perf_counter()
        emb, sv, _ = api.pca_fit_transform(G, 100)
        run["device_call_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
        run["device_stage_ms"] = {k: round(v, 2) for k, v in api.pca_stage_ms().items()}
    n3 = float(n) ** 3
    st = run["device_stage_ms"]
    run["covariance_tflops_fp64"] = round(n3 / (st["covariance"] * 1e-3) / 1e12, 2)      # n^3 FMA-flops of the upper half x 2
    # lower-triangle passes: 16 B (read + write) per element of the trailing lower triangle and step = 8/3 n^3 bytes
    run["tridiag_GBps"] = round(8.0 / 3.0 * n3 / (st["tridiagonalise"] * 1e-3) / 1e9, 1)
    if n <= 4000:
        from sklearn.decomposition import PCA as SkPCA
        dense = np.asarray(G.todense(), dtype=np.float64)
        t0 = time.perf_counter()
        ref = SkPCA(n_components=n - 1).fit_transform(dense)[:, :100]
        run["sklearn_s"] = round(time.perf_counter() - t0, 2)
        run["sklearn_cores"] = os.cpu_count()
        run["max_abs_diff_over_scale"] = float(np.abs(ref - emb).max() / np.abs(ref).max())
    out["runs"].append(run)
print(json.dumps(out))
