"""Small end-to-end runs of every kernel family for compute-sanitizer (memcheck): shared-table / tile / chunk K1,
packed + plain K3, both selects, windowed + hash K4 (incl. sub-parts), graph emit, narrow counts, async fit."""
import os
import sys

import numpy as np
from scipy.sparse import csr_matrix, vstack

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from schicexplorer_b200 import _capi, synth  # noqa: E402

api = _capi.cuda_api()
cfg = synth.small_config(72, config_id=3, median_nnz=900, min_nnz=1, max_nnz=3000)
ip, ix, dt, _ = synth.generate_csr(cfg, workers=1)
X = synth.to_scipy(cfg, ip, ix, dt)
rng = np.random.default_rng(1)
dense = csr_matrix((np.ones(6000), (np.zeros(6000, dtype=int), np.sort(rng.choice(1 << 19, 6000, replace=False)))),
                   shape=(1, X.shape[1]))                       # > 4096 keys in one window: K4 sub-parts
X = vstack([X, dense, dense, csr_matrix((1, X.shape[1]))]).tocsr()
ipx, ixx = X.indptr.astype(np.int64), np.ascontiguousarray(X.indices.astype(np.uint32))
f32 = np.ascontiguousarray(X.data.astype(np.float32))
u8 = np.ascontiguousarray(np.minimum(X.data, 255).astype(np.uint8))
ref = None
for k1 in ("table", "tile", "direct", "chunk"):
    for k3 in ("packed", "plain"):
        for k4 in ("window", "hash"):
            os.environ.update(SCHIC_K1_MODE=k1, SCHIC_K3_MODE=k3, SCHIC_K4_MODE=k4)
            h = api.create(n_neighbors=9, number_of_hash_functions=96, fast=False, minimal_blocks_in_common=1,
                           excess_factor=2, max_bin_size=100000)
            api.set_feature_range(h, X.shape[1])
            api.fit_csr(h, ipx, ixx, u8 if k3 == "packed" else f32.clip(max=255), False, async_upload=(k4 == "hash"))
            res = api.kneighbors(h, 9)
            g = api.kneighbors_graph(h, 9)
            sig = api.signatures(h, 96)
            cnt = api.collision_counts(h, 3, 41)
            api.destroy(h)
            cur = (sig, cnt, res[0], res[2], g[0], g[1])
            if ref is None:
                ref = cur
            assert all(np.array_equal(a, b) for a, b in zip(cur, ref)), (k1, k3, k4)
print("sanitize_small ok: 16 kernel combinations agree")
