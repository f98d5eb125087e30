"""Oracle of the dense-graph PCA post-step -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product path
(schicexplorer_b200/pca.py -> csrc/pca_dense.cu) never does.

What it restates: the reference's
    PCA(n_components=min(shape) - 1).fit_transform(graph.todense())[:, :dimensionsPCA]
(/root/reference/schicexplorer/scHicClusterMinHash.py:358-366, scHicCluster.py:183-188; scikit-learn's full-SVD PCA
with svd_flip(u_based_decision=False)), written as the SAME numerical steps the CUDA kernels take -- centring,
covariance, Householder tridiagonalisation (LAPACK dsytd2 conventions), Sturm bisection for the leading eigenvalues,
inverse iteration with Gram-Schmidt for their vectors, back-transformation, sign convention, projection -- so that a
stage of the device pipeline can be compared with the matching stage here.

PINNED: scikit-learn (the reference's dependency) is importable in the build and test image, and tests/test_pca.py
holds this restatement to sklearn.decomposition.PCA itself (1e-9 relative) on the reference's 20-cell graph and on
seeded dense graphs, so parity for this row is pinned to the reference's own arithmetic."""
from __future__ import annotations

import numpy as np


def tridiagonalize(C):
    """Householder reduction of symmetric C to tridiagonal (d, e). Reflector j (acting on indices j+1..n-1, v[j+1] = 1)
    is returned in V[j+1:, j] with its tau."""
    S = np.array(C, dtype=np.float64)
    n = S.shape[0]
    d, e, tau = np.zeros(n), np.zeros(max(n - 1, 0)), np.zeros(max(n - 1, 0))
    V = np.zeros((n, n))
    for j in range(n - 1):
        x = S[j + 1:, j].copy()
        alpha, xnorm = x[0], np.linalg.norm(x[1:])
        if xnorm == 0.0:
            t, beta = 0.0, alpha
            v = np.zeros_like(x)
        else:
            beta = -np.copysign(np.hypot(alpha, xnorm), alpha)
            t = (beta - alpha) / beta
            v = x / (alpha - beta)
        v[0] = 1.0
        e[j], tau[j], d[j] = beta, t, S[j, j]
        V[j + 1:, j] = v
        if t != 0.0:
            S22 = S[j + 1:, j + 1:]
            p = t * (S22 @ v)
            w = p - (0.5 * t * (p @ v)) * v
            S22 -= np.outer(v, w) + np.outer(w, v)
    if n:
        d[n - 1] = S[n - 1, n - 1]
    return d, e, V, tau


def tridiagonalize_blocked(C, nb=32):
    """The same reduction in panels of nb reflectors (LAPACK dlatrd / dsytrd scheme): inside a panel the trailing matrix
    is only READ (S v, corrected by the panel's V W^T + W V^T so far); it is updated once per panel by a rank-2nb
    product. Same d, e, V, tau as :func:`tridiagonalize` up to rounding. This is the formulation the device kernel should
    move to next (halves the HBM traffic of the tridiagonalisation again); kept here, tested against the
    unblocked version, so that the CUDA port has a stage-by-stage checker."""
    S = np.array(C, dtype=np.float64)
    n = S.shape[0]
    d, e, tau = np.zeros(n), np.zeros(max(n - 1, 0)), np.zeros(max(n - 1, 0))
    V = np.zeros((n, n))
    j0 = 0
    while j0 < n - 1:
        jb = min(nb, n - 1 - j0)
        Vp, Wp = np.zeros((n, jb)), np.zeros((n, jb))
        for t in range(jb):
            j = j0 + t
            col = S[j:, j] - Vp[j:, :t] @ Wp[j, :t] - Wp[j:, :t] @ Vp[j, :t]      # column j with the panel applied
            d[j] = col[0]
            x = col[1:].copy()
            alpha, xnorm = x[0], np.linalg.norm(x[1:])
            if xnorm == 0.0:
                tj, beta = 0.0, alpha
                v = np.zeros_like(x)
            else:
                beta = -np.copysign(np.hypot(alpha, xnorm), alpha)
                tj = (beta - alpha) / beta
                v = x / (alpha - beta)
            v[0] = 1.0
            e[j], tau[j] = beta, tj
            V[j + 1:, j] = v
            Vp[j + 1:, t] = v
            if tj != 0.0:
                p = S[j + 1:, j + 1:] @ v
                p -= Vp[j + 1:, :t] @ (Wp[j + 1:, :t].T @ v) + Wp[j + 1:, :t] @ (Vp[j + 1:, :t].T @ v)
                p *= tj
                Wp[j + 1:, t] = p - (0.5 * tj * (p @ v)) * v
        r = j0 + jb
        S[r:, r:] -= Vp[r:, :] @ Wp[r:, :].T + Wp[r:, :] @ Vp[r:, :].T
        j0 = r
    if n:
        d[n - 1] = S[n - 1, n - 1]
    return d, e, V, tau


def sturm_count(d, e2, x):
    """Number of eigenvalues of the tridiagonal matrix below x (vectorised over x)."""
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    cnt = np.zeros(x.shape, dtype=np.int64)
    q = d[0] - x
    cnt += q < 0
    tiny = 1e-300
    for i in range(1, len(d)):
        q = np.where(np.abs(q) < tiny, -tiny, q)
        q = d[i] - x - e2[i - 1] / q
        cnt += q < 0
    return cnt


def top_eigenvalues(d, e, m):
    """The m largest eigenvalues, descending, by bisection."""
    n = len(d)
    e2 = e * e
    bound = np.max(np.abs(d) + np.abs(np.r_[e, 0.0]) + np.abs(np.r_[0.0, e]))
    lo, hi = np.full(m, -bound), np.full(m, bound)
    target = n - 1 - np.arange(m)
    for _ in range(110):
        mid = 0.5 * (lo + hi)
        right = sturm_count(d, e2, mid) <= target
        lo, hi = np.where(right, mid, lo), np.where(right, hi, mid)
    return 0.5 * (lo + hi)


def tridiag_solve_shifted(d, e, theta, b, pivmin):
    """(T - theta I) x = b: LU with partial pivoting (dgtsv layout), one right-hand side."""
    n = len(d)
    x = np.array(b, dtype=np.float64)
    DD, DU, DU2 = np.zeros(n), np.zeros(n), np.zeros(n)
    dd = d[0] - theta
    du = e[0] if n > 1 else 0.0
    xi = x[0]
    for i in range(n - 1):
        dl, dnext = e[i], d[i + 1] - theta
        unext = e[i + 1] if i + 2 < n else 0.0
        xn = x[i + 1]
        if abs(dd) >= abs(dl):
            if dd == 0.0:
                dd = pivmin
            f = dl / dd
            DD[i], DU[i], DU2[i] = dd, du, 0.0
            x[i] = xi
            xi = xn - f * xi
            dd, du = dnext - f * du, unext
        else:
            f = dd / dl
            DD[i], DU[i], DU2[i] = dl, dnext, unext
            x[i] = xn
            xi = xi - f * xn
            dd, du = du - f * dnext, -f * unext
    if dd == 0.0:
        dd = pivmin
    x1, x2 = xi / dd, 0.0
    x[n - 1] = x1
    for i in range(n - 2, -1, -1):
        x0 = (x[i] - DU[i] * x1 - DU2[i] * x2) / DD[i]
        x[i] = x0
        x2, x1 = x1, x0
    return x


def tridiag_vectors(d, e, theta, sweeps=3, seed=0):
    """Eigenvectors for the eigenvalues theta: inverse iteration + classical Gram-Schmidt (twice), in order."""
    n, m = len(d), len(theta)
    tn = np.max(np.abs(d) + np.abs(np.r_[e, 0.0]) + np.abs(np.r_[0.0, e])) if n else 0.0
    pivmin = (tn if tn > 0 else 1.0) * np.finfo(np.float64).eps      # tn == 0: the all-zero matrix (any basis is right)
    X = np.random.default_rng(seed).random((n, m)) - 0.5
    for _ in range(sweeps):
        for i in range(m):
            X[:, i] = tridiag_solve_shifted(d, e, theta[i], X[:, i], pivmin)
            X[:, i] /= np.linalg.norm(X[:, i])
        for i in range(m):
            for _ in range(2):
                if i:
                    X[:, i] -= X[:, :i] @ (X[:, :i].T @ X[:, i])
            nrm = np.linalg.norm(X[:, i])
            X[:, i] = X[:, i] / nrm if nrm > 0 else 0.0
    return X


def back_transform(V, tau, X):
    Z = np.array(X, dtype=np.float64)
    for j in range(V.shape[0] - 2, -1, -1):
        if tau[j] == 0.0:
            continue
        v = V[j + 1:, j]
        Z[j + 1:, :] -= tau[j] * np.outer(v, v @ Z[j + 1:, :])
    return Z


def pca_fit_transform(X, n_components):
    """(embedding [n_rows, m], singular values [m], components [m, n_cols]) of the dense matrix X."""
    X = np.array(X, dtype=np.float64)
    X[~np.isfinite(X)] = 0.0
    m = int(n_components)
    Xc = X - X.mean(axis=0)
    d, e, V, tau = tridiagonalize(Xc.T @ Xc)
    theta = top_eigenvalues(d, e, m)
    Z = back_transform(V, tau, tridiag_vectors(d, e, theta))
    flip = np.sign(Z[np.argmax(np.abs(Z), axis=0), np.arange(m)])
    flip[flip == 0] = 1.0
    Z = Z * flip
    return Xc @ Z, np.sqrt(np.maximum(theta, 0.0)), Z.T.copy()
