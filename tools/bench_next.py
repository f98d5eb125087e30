"""Measurement of the SURVEY 8(f) rows built on the hot path's kernels (one JSON object on stdout):

* exact kNN (scHicCluster -drm knn, row 3): brute-force K4 over all n^2 pairs of the S10 workload (10 000 cells,
  1.43e8 non-zeros, k = 100), timed on the device (stage timers) and end to end through NearestNeighbors (host CSR in,
  host graph out); pairs/s, and the algorithmic stream rate 8 B x nnz(candidate) per pair against the HBM peak;
  beside it the reference's own call - sklearn NearestNeighbors(algorithm='ball_tree').fit(X).kneighbors_graph() - on a
  bounded sample of the same cells on the host cores;
* hyperopt fast path (row 4): the per-trial graph (switch hash count + query side, CSR resident) against a trial that
  re-fits with that hash count (what every trial of the reference's tool does, without even counting its reload).

usage: python tools/bench_next.py [n_cells] [sklearn_sample_cells]"""
import json
import os
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from schicexplorer_b200 import MinHash, NearestNeighbors, _capi, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
n_sk = int(sys.argv[2]) if len(sys.argv) > 2 else 1_500
k = 100
cfg = synth.SynthConfig(**{**synth.CONFIGS["s10"].__dict__, "n_cells": n})
ip, ix, dt, _ = synth.generate_csr(cfg, 0, n, workers=min(32, os.cpu_count() or 1))
X = synth.to_scipy(cfg, ip, ix, dt)
api = _capi.cuda_api()
out = {"workload": "s10", "n_cells": n, "nnz": int(ip[-1]), "k": k}

# ---- row 3: exact kNN ------------------------------------------------------------------------------------------------
h = api.create(n_neighbors=k, number_of_hash_functions=1, fast=False)
api.set_feature_range(h, X.shape[1])
api.fit_csr(h, ip, ix, dt, False)
times, stage = [], None
for rep in range(4):
    t0 = time.perf_counter()
    idx, dist, nv = api.kneighbors_exact(h, k, False)
    times.append(time.perf_counter() - t0)
    stage = api.stage_ms(h)
api.destroy(h)
rerank_ms = float(stage["rerank"])
pairs = n * n
mean_row = ip[-1] / n
out["exact_knn"] = {
    "call_ms": [round(t * 1e3, 1) for t in times], "k4_ms": round(rerank_ms, 2), "prep_ms": round(float(stage["rerank_prep"]), 2),
    "pairs_per_s": pairs / (rerank_ms * 1e-3), "cells_per_s": n / min(times),
    "stream_GBps": 8.0 * mean_row * pairs / (rerank_ms * 1e-3) / 1e9,
    "note": "stream = 8 B (id + value) x nnz(candidate) per pair; all 148 CTAs walk the same 1024-row candidate chunk, "
            "so most of it is served by L2, not HBM",
}
t0 = time.perf_counter()
nn = NearestNeighbors(n_neighbors=k, algorithm="ball_tree").fit(X)
g = nn.kneighbors_graph(mode="distance")
out["exact_knn"]["class_fit_plus_graph_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
nn.close()
assert g.shape == (n, n) and (np.diff(g.indptr) == k).all()

from sklearn.neighbors import NearestNeighbors as SkNN  # noqa: E402
Xs = X[:n_sk]
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    t0 = time.perf_counter()
    gs = SkNN(n_neighbors=min(k, n_sk - 1), algorithm="ball_tree", n_jobs=os.cpu_count()).fit(Xs).kneighbors_graph(mode="distance")
    t_sk = time.perf_counter() - t0
out["exact_knn"]["sklearn"] = {"sample_cells": n_sk, "seconds": round(t_sk, 2), "pairs_per_s": n_sk * n_sk / t_sk,
                               "cores": os.cpu_count()}
# the sample's graph from the GPU for a parity spot check (same rows, same k)
nn = NearestNeighbors(n_neighbors=min(k, n_sk - 1)).fit(Xs)
gg = nn.kneighbors_graph(mode="distance")
nn.close()
gs.sort_indices(); gg.sort_indices()
# identical up to ties: where the neighbour sets differ, the swapped entries must sit at the row's k-th distance
differing, beyond_ties = 0, 0
for r in range(n_sk):
    si, sd = gs.indices[gs.indptr[r]:gs.indptr[r + 1]], gs.data[gs.indptr[r]:gs.indptr[r + 1]]
    gi, gd = gg.indices[gg.indptr[r]:gg.indptr[r + 1]], gg.data[gg.indptr[r]:gg.indptr[r + 1]]
    if np.array_equal(si, gi):
        beyond_ties += not np.allclose(sd, gd, rtol=1e-5)
        continue
    differing += 1
    only_s, only_g = ~np.isin(si, gi), ~np.isin(gi, si)
    beyond_ties += not (np.allclose(sd[only_s], sd.max(), rtol=1e-5) and np.allclose(gd[only_g], sd.max(), rtol=1e-5))
out["exact_knn"]["sample_rows_with_tie_swaps_vs_sklearn"] = int(differing)
out["exact_knn"]["sample_rows_differing_beyond_ties"] = int(beyond_ties)
out["exact_knn"]["speedup_vs_sklearn_pairs_per_s"] = out["exact_knn"]["pairs_per_s"] / out["exact_knn"]["sklearn"]["pairs_per_s"]

# ---- row 4: hyperopt sweep over -nh ----------------------------------------------------------------------------------
kw = dict(n_neighbors=k, fast=True, minimal_blocks_in_common=100, excess_factor=1, max_bin_size=100000, shingle_size=5)
sweep = [400, 800, 1200, 1600, 2000]
mh = MinHash(number_of_hash_functions=max(sweep), **kw)
t0 = time.perf_counter(); mh.fit(X); t_fit = time.perf_counter() - t0
resident, refit = {}, {}
for H in sweep + sweep:
    t0 = time.perf_counter()
    g1 = mh.set_number_of_hash_functions(H).kneighbors_graph(mode="distance")
    resident[H] = time.perf_counter() - t0
mh.close()
for H in sweep:
    m2 = MinHash(number_of_hash_functions=H, **kw)
    m2.fit(X); m2.kneighbors_graph(mode="distance")          # warm-up of this handle's buffers
    t0 = time.perf_counter()
    m2.fit(X)
    g2 = m2.kneighbors_graph(mode="distance")
    refit[H] = time.perf_counter() - t0
    m2.close()
out["hyperopt_sweep"] = {
    "hashes": sweep, "fit_at_max_ms": round(t_fit * 1e3, 1),
    "resident_trial_ms": [round(resident[H] * 1e3, 1) for H in sweep],
    "refit_trial_ms": [round(refit[H] * 1e3, 1) for H in sweep],
    "note": "trial = graph of 10k cells, k = 100, fast distances, host CSR graph out; refit additionally uploads and "
            "hashes the CSR (the reference's tool also reloads the scool and rebuilds the CSR per trial)",
}
print(json.dumps(out))
