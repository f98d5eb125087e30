"""Second, independent restatement of the MinHash kNN path in numpy -- TEST INFRASTRUCTURE ONLY.

Written as all-pairs array arithmetic (no inverse index) directly from SURVEY.md 8(c) rules 1-11, so
that it shares no code path with oracle.cpp (bucket walk) and no layout with the CUDA kernels. It
generates the committed golden vectors (tests/golden/make_golden.py) and cross-checks the C++
oracle. PARITY UNPINNED: see the header of oracle/oracle.cpp -- upstream `sparse-neighbors-search`
is absent and the reference's tests pin no arithmetic on this path.
"""
import numpy as np

M = np.uint32(2147483647)


def wang32(k):
    """Thomas Wang's 32-bit mix (docs/content/tools/scHicClusterMinHash.rst:11), uint32 wraparound."""
    k = k.astype(np.uint32, copy=True)
    k = ~k + (k << np.uint32(15))
    k ^= k >> np.uint32(12)
    k += k << np.uint32(2)
    k ^= k >> np.uint32(4)
    k *= np.uint32(2057)
    k ^= k >> np.uint32(16)
    return k


def wang64(k):
    k = k.astype(np.uint64, copy=True)
    k = ~k + (k << np.uint64(15))
    k ^= k >> np.uint64(12)
    k += k << np.uint64(2)
    k ^= k >> np.uint64(4)
    k *= np.uint64(2057)
    k ^= k >> np.uint64(16)
    return k


def signatures(X, n_hash, width=32, block=64):
    """Rules 2-3. X: scipy CSR. Returns uint32 [n, H]; empty rows get the sentinel M."""
    n = X.shape[0]
    sig = np.full((n, n_hash), M, dtype=np.uint32)
    with np.errstate(over="ignore"):
        for c in range(n):
            f = X.indices[X.indptr[c]:X.indptr[c + 1]].astype(np.uint64)
            if len(f) == 0:
                continue
            for j0 in range(0, n_hash, block):
                j = np.arange(j0, min(n_hash, j0 + block), dtype=np.uint64)
                if width == 32:
                    k = ((f + np.uint64(1)).astype(np.uint32)[:, None] * (j + np.uint64(1)).astype(np.uint32)[None, :])
                    h = wang32(k) % M
                else:
                    k = (f + np.uint64(1))[:, None] * (j + np.uint64(1))[None, :]
                    h = (wang64(k) % np.uint64(M)).astype(np.uint32)
                sig[c, j0:j0 + len(j)] = h.min(axis=0)
    return sig


def combine(sig, block, width=32, empty_rows=None):
    """Rule 4 (block combine / "shingle", decision point D2; off unless asked for): fold the minima of `block`
    consecutive hash functions into one key, key = sig[b*s]; key = h'(sig[b*s+t] + 1, key + 1), t = 1..s-1, where h' is
    the family's mix of the product of its arguments. [n, H] -> [n, H // block]; empty rows keep the sentinel M."""
    n, H = sig.shape
    T = H // block
    out = np.empty((n, T), dtype=np.uint32)
    with np.errstate(over="ignore"):
        for b in range(T):
            key = sig[:, b * block].astype(np.uint32)
            for t in range(1, block):
                v = sig[:, b * block + t]
                if width == 32:
                    key = wang32((v + np.uint32(1)) * (key + np.uint32(1))) % M
                else:
                    k = (v.astype(np.uint64) + np.uint64(1)) * (key.astype(np.uint64) + np.uint64(1))
                    key = (wang64(k) % np.uint64(M)).astype(np.uint32)
            out[:, b] = key
    if empty_rows is not None:
        out[empty_rows] = M
    return out


def collision_counts(sig, max_bin_size=None):
    """Rules 5-6, all pairs: cnt[c][d] = #{j: sig[c][j] == sig[d][j], value admissible}."""
    n, H = sig.shape
    ok = (sig != 0) & (sig != M)
    if max_bin_size is not None and n >= max_bin_size:
        for j in range(H):
            vals, inv, counts = np.unique(sig[:, j], return_inverse=True, return_counts=True)
            ok[:, j] &= counts[inv] < max_bin_size
    cnt = np.zeros((n, n), dtype=np.int64)
    for c in range(n):
        cnt[c] = ((sig == sig[c][None, :]) & ok[c][None, :] & ok).sum(axis=1)
    return cnt


def euclid(X, a, b):
    """Rule 9: dense-free euclidean distance in float64, rounded to float32."""
    d = (X[a] - X[b])
    return np.float32(np.sqrt(np.float64((d.multiply(d)).sum())))


def kneighbors(X, sig, k, fast=True, min_blocks=1, excess_factor=1, absolute=False, max_bin_size=None):
    """Rules 7-9. Returns idx [n,k] (-1 pad), dist [n,k] (inf pad), n_valid [n]."""
    n, H = sig.shape
    cnt = collision_counts(sig, max_bin_size)
    idx = np.full((n, k), -1, dtype=np.int32)
    dist = np.full((n, k), np.inf, dtype=np.float32)
    nv = np.zeros(n, dtype=np.int32)
    Xf = X.astype(np.float64)
    for c in range(n):
        cand = np.nonzero((cnt[c] >= max(min_blocks, 1)))[0]
        order = np.lexsort((cand, -cnt[c][cand]))          # count desc, index asc
        cand = cand[order][:k * max(1, excess_factor)]
        if fast:
            d = cnt[c][cand].astype(np.float32) if absolute else \
                np.float32(1.0) - cnt[c][cand].astype(np.float32) / np.float32(H)
        else:
            d = np.array([euclid(Xf, c, int(x)) for x in cand], dtype=np.float32)
            order = np.lexsort((cand, d))                  # distance asc, index asc
            cand, d = cand[order], d[order]
        cand, d = cand[:k], d[:k]
        nv[c] = len(cand)
        idx[c, :len(cand)] = cand
        dist[c, :len(cand)] = d
    return idx, dist, nv
