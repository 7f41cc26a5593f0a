#!/usr/bin/env python
"""numpy prototype for the next eigensolver step: whitening matrix W = S^-1/2 (pseudo-inverse on the range of S)
without an eigendecomposition, from GEMMs only, so that the C <= 128 problem can be spread over many SMs instead of one.

  S = L L^T           rank-revealing pivoted Cholesky (already the first phase of jacobi_chol_kernel), L: n x r
  B = L^T L           r x r, full rank, cond(B) = cond(S on its range)
  Z -> B^-1/2         coupled Newton-Schulz:  T = (3I - Z Y)/2,  Y <- Y T,  Z <- T Z   (Y0 = B/s, Z0 = I, s >= lambda_max)
  W = L Z^3 L^T       since (L B^-3/2 L^T)^2 = L B^-2 L^T = S^+

Prints iterations / GEMM counts and the error against the LAPACK pseudo-inverse square root on the covariances of the
reference-generated golden features (incl. the rank-deficient stage-5 ones) and on synthetic spectra.  No GPU, no oracle."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def chol_pivot(S, thr_rel=1e-14):
    n = len(S); A = S.copy(); d = np.diag(A).copy(); cols = []; thr = thr_rel * d.max(); done = np.zeros(n, bool)
    for _ in range(n):
        dm = np.where(done, -np.inf, d); p = int(np.argmax(dm))
        if dm[p] <= thr:
            break
        l = np.where(done, 0.0, A[:, p]) / np.sqrt(d[p]); l[p] = np.sqrt(d[p]); done[p] = True
        nd = ~done
        A[np.ix_(nd, nd)] -= np.outer(l[nd], l[nd]); d[nd] -= l[nd] ** 2
        cols.append(l)
    return np.array(cols).T          # n x r


def ns_invsqrt(B, tol=1e-14, maxit=60):
    n = len(B); s = np.abs(B).sum(0).max(); Y = B / s; Z = np.eye(n); I = np.eye(n); it = 0
    for it in range(1, maxit + 1):
        T = 0.5 * (3 * I - Z @ Y); Y = Y @ T; Z = T @ Z
        if np.abs(I - Z @ Y).max() < tol:
            break
    return Z / np.sqrt(s), it


def whiten_ns(S):
    L = chol_pivot(S); Z, it = ns_invsqrt(L.T @ L)
    return L @ (Z @ Z @ Z) @ L.T, L.shape[1], it


def whiten_ref(S, tau=1e-10):
    w, v = np.linalg.eigh(S); keep = w > tau * w.max()
    return (v[:, keep] * w[keep] ** -0.5) @ v[:, keep].T


if __name__ == "__main__":
    mats = {}
    W = np.load(os.path.join(ROOT, "tests", "golden", "golden_16x.npz"))
    for s in (5, 4, 3):
        for w in ("cF", "sF"):
            f = W["a10.%s%d" % (w, s)].astype(np.float64); f = f.reshape(f.shape[0], -1)
            fc = f - f.mean(1, keepdims=True); S = fc @ fc.T / (fc.shape[1] - 1)
            live = np.diag(S) > 0; mats["golden.%s%d" % (w, s)] = S[np.ix_(live, live)]
    rng = np.random.default_rng(1)
    for C, lo in [(99, 2e-3), (51, 2.4e-2), (59, 3.3e-4), (28, 2.9e-5), (24, 4.4e-4), (100, 1e-7)]:   # cfg3 live sizes / spreads
        Q, _ = np.linalg.qr(rng.standard_normal((C, C))); mats["syn%d_%.0e" % (C, lo)] = (Q * np.logspace(0, np.log10(lo), C)) @ Q.T
    for key, S in mats.items():
        Wn, r, it = whiten_ns(S); Wr = whiten_ref(S)
        print("%-14s n=%3d rank=%3d  NS iterations=%2d (%d GEMMs of r^3)  max|W-Wref|/max|Wref| = %.1e"
              % (key, len(S), r, it, 3 * it + 4, np.abs(Wn - Wr).max() / np.abs(Wr).max()))
