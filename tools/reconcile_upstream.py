#!/usr/bin/env python
"""Pin the oracle to upstream the day upstream appears (SURVEY.md H1 / VERDICT r01 "parity unpinned").

The arithmetic of the hot path lives in the third-party package ``sparse-neighbors-search`` (``requirements.txt:3`` of
the reference), which is neither vendored in /root/reference nor installable offline. This tool looks for a REAL
install of it -- ``baseline/_ref/`` first, then the interpreter's own site-packages, never this repository's import
shim ``sparse_neighbors_search/`` -- and, when one is importable, runs the reference's own call sequence
(``scHicClusterMinHash.py:402-406``: MinHash(...).fit(X); kneighbors_graph(mode='distance')) on the reference's 20-cell
test matrix and diffs the result against the committed golden vectors of the oracle (``tests/golden/golden20.npz``) for
BOTH HashSpecs (width 32 / 64, SURVEY 8(c) decision point D1):

* fast mode: upstream's distances are 1 - collisions/H (decision point D5), so the collision counts are recovered from
  them and compared with ``cnt32`` / ``cnt64`` entry by entry; neighbour indices are compared up to ties;
* exact mode (fast=False): euclidean distances against ``dist_exact`` (1e-5 relative), indices up to ties;
* if the upstream object exposes its signatures (attribute names tried below), they are compared with ``sig32`` /
  ``sig64`` directly.

Exit status: 0 = upstream found and at least one HashSpec reproduces it (printed which); 1 = upstream found, neither
matches (the report says where they differ -- the decision points D1-D5 to revisit); 3 = upstream not importable
(nothing to reconcile; the state today). ``--json`` prints the report as one JSON object.

  python tools/reconcile_upstream.py [--json] [--path DIR_WITH_sparse_neighbors_search]
"""
from __future__ import annotations

import argparse
import importlib
import importlib.util
import json
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
GOLDEN = os.path.join(ROOT, "tests", "golden")
SHIM = os.path.join(ROOT, "sparse_neighbors_search")
# kwargs of the reference's default branch (scHicClusterMinHash.py:402-404), H = 800 as in its tests
REF_KW = dict(number_of_hash_functions=800, number_of_cores=4, shingle_size=5, absolute_numbers=False, max_bin_size=100000,
              minimal_blocks_in_common=100, excess_factor=1, prune_inverse_index=False)


def find_upstream(extra_path=None):
    """Import the real package or return (None, reason). The repository's shim re-exports OUR classes and must not
    be mistaken for upstream."""
    candidates = [p for p in (extra_path, os.path.join(ROOT, "baseline", "_ref")) if p and os.path.isdir(p)]
    saved = list(sys.path)
    try:
        sys.path[:] = candidates + [p for p in saved if os.path.abspath(p or ".") != ROOT]
        sys.modules.pop("sparse_neighbors_search", None)
        spec = importlib.util.find_spec("sparse_neighbors_search")
        if spec is None or not spec.origin:
            return None, "sparse_neighbors_search is not importable (not in baseline/_ref, not installed)"
        if os.path.abspath(os.path.dirname(spec.origin)) == SHIM:
            return None, "only this repository's import shim is on the path"
        mod = importlib.import_module("sparse_neighbors_search")
        if not hasattr(mod, "MinHash"):
            return None, f"{spec.origin} has no MinHash class"
        return mod, spec.origin
    except Exception as e:          # a broken install is "not importable", with the reason kept
        return None, f"import failed: {e!r}"
    finally:
        sys.path[:] = saved


def fixture_matrix():
    sys.path.insert(0, ROOT)
    from schicexplorer_b200.loader import cells_to_csr, load_cells_npz
    cells, chroms = load_cells_npz(os.path.join(GOLDEN, "fixture20.npz"))
    X, _ = cells_to_csr(cells, chroms)
    return X


def ties_only(idx, dist, ridx, rdist, rtol):
    """True when the two neighbour lists differ only inside groups of (nearly) equal distance."""
    for a, da, b, db in zip(idx, dist, ridx, rdist):
        ok = a >= 0
        if not np.array_equal(np.sort(a[ok]), np.sort(b[b >= 0])):
            return False
        bad = a != b
        if bad.any() and not np.allclose(da[bad], db[bad], rtol=max(rtol, 1e-7)):
            return False
    return True


def dense_lists(graph, n, k):
    """kneighbors_graph(mode='distance') -> ([n, k] indices by (distance, index), [n, k] distances), -1 / inf padded."""
    g = graph.tocsr()
    idx = np.full((n, k), -1, dtype=np.int64)
    dist = np.full((n, k), np.inf, dtype=np.float64)
    for r in range(n):
        c, d = g.indices[g.indptr[r]:g.indptr[r + 1]], g.data[g.indptr[r]:g.indptr[r + 1]]
        order = np.lexsort((c, d))[:k]
        idx[r, :len(order)], dist[r, :len(order)] = c[order], d[order]
    return idx, dist


def reconcile(mod, origin):
    golden = np.load(os.path.join(GOLDEN, "golden20.npz"))
    X = fixture_matrix()
    n, H, k = X.shape[0], REF_KW["number_of_hash_functions"], 20
    report = {"upstream": origin, "version": getattr(mod, "__version__", None), "checks": {}}
    # ---- fast mode: distances -> collision counts -------------------------------------------------------------
    mh = mod.MinHash(n_neighbors=k, fast=True, maxFeatures=int(max(X.getnnz(1))), **REF_KW)
    mh.fit(X)
    g = mh.kneighbors_graph(mode="distance")
    idx, dist = dense_lists(g, n, k)
    cnt_up = np.rint((1.0 - dist) * H)
    for width in (32, 64):
        gold = golden[f"cnt{width}"].astype(np.int64)
        want = np.where(idx >= 0, np.take_along_axis(gold, np.maximum(idx, 0), axis=1), -1)
        have = np.where(idx >= 0, cnt_up, -1)
        report["checks"][f"fast_counts_width{width}"] = {
            "equal": bool(np.array_equal(want, have)), "entries": int((idx >= 0).sum()),
            "mismatches": int((want != have).sum()), "max_abs_diff": float(np.abs(want - have).max())}
    gi, gd = golden["idx_fast"], golden["dist_fast"].astype(np.float64)
    report["checks"]["fast_neighbours_vs_golden(width32)"] = {"identical_up_to_ties": bool(ties_only(idx, dist, gi, gd, 0))}
    # ---- signatures, when the object shows them ------------------------------------------------------------------
    sig = None
    for name in ("signatures", "get_signatures", "_signatures", "signature_"):
        obj = getattr(mh, name, None)
        try:
            sig = np.asarray(obj() if callable(obj) else obj) if obj is not None else None
        except Exception:
            sig = None
        if sig is not None and sig.shape == (n, H):
            break
        sig = None
    if sig is not None:
        for width in (32, 64):
            report["checks"][f"signatures_width{width}"] = {"equal": bool(np.array_equal(sig.astype(np.uint32), golden[f"sig{width}"]))}
    else:
        report["checks"]["signatures"] = {"skipped": "upstream object exposes no [n, H] signature array under the names tried"}
    # ---- exact mode ------------------------------------------------------------------------------------------------
    mh = mod.MinHash(n_neighbors=k, fast=False, maxFeatures=int(max(X.getnnz(1))), **REF_KW)
    mh.fit(X)
    idx, dist = dense_lists(mh.kneighbors_graph(mode="distance"), n, k)
    gi, gd = golden["idx_exact"], golden["dist_exact"].astype(np.float64)
    ok = np.isfinite(gd) & np.isfinite(dist)
    report["checks"]["exact_distances"] = {
        "within_1e-5": bool(np.array_equal(np.isfinite(gd), np.isfinite(dist)) and
                            np.allclose(np.sort(dist[ok]), np.sort(gd[ok]), rtol=1e-5)),
        "identical_up_to_ties": bool(ties_only(idx, dist, gi, gd, 1e-5))}
    matches = [w for w in (32, 64) if report["checks"][f"fast_counts_width{w}"]["equal"]]
    report["hash_width_matching_upstream"] = matches
    report["pinned"] = bool(matches) and report["checks"]["exact_distances"]["within_1e-5"]
    return report


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--json", action="store_true")
    ap.add_argument("--path", default=None, help="directory that contains the sparse_neighbors_search package")
    args = ap.parse_args(argv)
    mod, origin = find_upstream(args.path)
    if mod is None:
        report = {"upstream": None, "reason": origin, "pinned": False,
                  "status": "parity unpinned: nothing to reconcile against (SURVEY.md F2); golden vectors are the oracle's own"}
        print(json.dumps(report) if args.json else f"upstream sparse_neighbors_search not found: {origin}\n{report['status']}")
        return 3
    report = reconcile(mod, origin)
    if args.json:
        print(json.dumps(report))
    else:
        print(f"upstream: {origin} (version {report['version']})")
        for name, res in report["checks"].items():
            print(f"  {name}: {res}")
        print("hash width matching upstream:", report["hash_width_matching_upstream"] or "NONE (revisit D1-D5)")
    return 0 if report["pinned"] else 1


if __name__ == "__main__":
    sys.exit(main())
