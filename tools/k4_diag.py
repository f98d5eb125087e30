"""Diagnostic: the matrix of tests/test_gpu_parity.py::test_counts_rerank_kernel through every (lanes per slice, chunks per
round) build of the direct counts kernel, against the oracle; prints the worst relative distance error per variant."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import Fitted  # noqa: E402
from test_gpu_parity import _count_rows  # noqa: E402
from oracle import oracle_api  # noqa: E402
from schicexplorer_b200 import _capi  # noqa: E402

cuda, oracle = _capi.cuda_api(), oracle_api()
rng = np.random.default_rng(41)
ncols = 1 << 22
pool = np.sort(rng.choice(1 << 17, size=20000, replace=False)) + (3 << 17)
base = _count_rows(rng, 70, ncols, [3000, 4500, 6000, 800, 2049, 2048], pool=pool)
scat = _count_rows(rng, 70, ncols, [700, 50, 1200])
X = (base + scat).tocsr()
X.sum_duplicates()
X.sort_indices()
kw = dict(fast=False, n_neighbors=25, number_of_hash_functions=128, minimal_blocks_in_common=1, excess_factor=2)
with Fitted(oracle, X, **kw) as o:
    oi, od, on = oracle.kneighbors(o.h, 25)
for lps, u in (("8", "2"), ("8", "3"), ("4", "4"), ("4", "6")):
    for threads in ("416", "512"):
        os.environ.update(SCHIC_K4_KERNEL="direct", SCHIC_K4_LPS=lps, SCHIC_K4_U=u, SCHIC_K4_THREADS=threads)
        with Fitted(cuda, X, **kw) as g:
            cuda.set_feature_range(g.h, X.shape[1])
            gi, gd, gn = cuda.kneighbors(g.h, 25)
        m = np.isfinite(od) & np.isfinite(gd) & (od > 0)
        rel = np.abs(gd[m] - od[m]) / od[m]
        print(f"lps={lps} u={u} threads={threads} max_rel={rel.max():.3e} n_bad={int((rel > 1e-5).sum())}", flush=True)
