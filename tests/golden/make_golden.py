"""Regenerates tests/golden/golden20.npz: signatures, collision counts and neighbour lists of the
20-cell fixture computed by the numpy restatement (oracle/oracle_np.py) with the default HashSpec
(width 32) and the alternative (width 64). The reference ships no golden vectors for this path
(parity unpinned); these pin our two oracles and the GPU against each other."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(__file__), "..", "..")
sys.path.insert(0, ROOT)
from oracle import oracle_np  # noqa: E402
from schicexplorer_b200.loader import cells_to_csr, load_cells_npz  # noqa: E402

if __name__ == "__main__":
    here = os.path.dirname(__file__)
    cells, chroms = load_cells_npz(os.path.join(here, "fixture20.npz"))
    X, names = cells_to_csr(cells, chroms)
    H = 800
    out = {}
    for width in (32, 64):
        sig = oracle_np.signatures(X, H, width=width)
        cnt = oracle_np.collision_counts(sig)
        out[f"sig{width}"] = sig
        out[f"cnt{width}"] = cnt.astype(np.uint16)
    sig = out["sig32"]
    # the kwargs of scHicClusterMinHash.py:402-404 (k = ncells = 20)
    for fast in (True, False):
        idx, dist, nv = oracle_np.kneighbors(X, sig, 20, fast=fast, min_blocks=100, excess_factor=1)
        tag = "fast" if fast else "exact"
        out[f"idx_{tag}"], out[f"dist_{tag}"], out[f"nv_{tag}"] = idx, dist, nv
    np.savez_compressed(os.path.join(here, "golden20.npz"), **out)
    print({k: v.shape for k, v in out.items()})
