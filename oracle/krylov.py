"""Lanczos eigensolver -- restatement of KrylovKit.jl `eigsolve(A, x0, 1, :SR, Lanczos(...))`
(oracle; test infrastructure only).

KrylovKit (0.9 | 0.10 per /root/reference/Project.toml:18) is a third-party dependency that is
NOT vendored under /root/reference; the reference reaches it from `eig_solver`
(src/base/solver.jl:23-43) with krylovdim=5, maxiter=2, tol=1e-14, eager=false.  What follows
restates its published algorithm: Lanczos factorisation with ModifiedGramSchmidt2
re-orthogonalisation (`initialize` / `lanczosrecurrence`), Rayleigh-Ritz on the tridiagonal,
convergence test on |beta * U[K, i]|, and the Krylov-Schur style thick restart
keep = div(3*krylovdim + 2*converged, 5) with Householder restoration of tridiagonal form.

The vector type only needs: inner(x, y), x.add(y, a) (= x + a*y), x.scale(a), x.norm().
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import numpy as np


class VecOps:
    """Adapter so the same code runs on BSTensor (oracle) and numpy arrays (dense KATs)."""

    @staticmethod
    def inner(x, y):
        if isinstance(x, np.ndarray):
            return np.vdot(x, y)
        from .blocksparse import inner
        return inner(x, y)

    @staticmethod
    def add(x, y, a):
        if isinstance(x, np.ndarray):
            return x + a * y
        return x.add(y, a)

    @staticmethod
    def scale(x, a):
        if isinstance(x, np.ndarray):
            return x * a
        return x.scale(a)

    @staticmethod
    def norm(x):
        if isinstance(x, np.ndarray):
            return float(np.linalg.norm(x))
        return x.norm()


def _householder(x: np.ndarray, i: int):
    """KrylovKit `_householder!`: reflector H = I - beta v v^T with H x = nu e_i, nu = ||x|| >= 0."""
    v = np.array(x, dtype=np.float64)
    sigma = float(np.sum(np.delete(v, i) ** 2))      # summed WITHOUT v[i]: no cancellation
    vi = v[i]
    nu = np.sqrt(vi * vi + sigma)
    if sigma == 0.0 and vi == nu:
        return 0.0, v, nu
    if vi < 0:
        vi = vi - nu
    else:
        vi = -sigma / (vi + nu)
    v = v / vi
    v[i] = 1.0
    beta = -vi / nu
    return beta, v, nu


def _restore_tridiagonal(D: np.ndarray, f: np.ndarray, U: np.ndarray, keep: int):
    """Thick-restart bookkeeping of KrylovKit `eigsolve` (Lanczos): bring [diag(D[:keep]); f[:keep]^T]
    back to Lanczos (tridiagonal) form with Householder reflections applied from column `keep`
    down to 1; the same reflections are accumulated into U.  Returns alphas, betas, U."""
    H = np.zeros((keep + 1, keep))
    for j in range(keep):
        H[j, j] = D[j]
        H[keep, j] = f[j]
    U = U.copy()
    for j in range(keep - 1, -1, -1):            # 0-based column j <-> Julia j+1
        beta, v, nu = _householder(H[j + 1, :j + 1], j)
        H[j + 1, j] = nu
        H[j + 1, :j] = 0.0
        if beta != 0.0:
            # lmul!(h, H): rows 0..j ;  rmul!(view(H, 1:j, :), h') : columns 0..j of rows 0..j
            R = H[:j + 1, :]
            R -= beta * np.outer(v, v @ R)
            C = H[:j + 1, :j + 1]
            C -= beta * np.outer(C @ v, v)
            Uc = U[:, :j + 1]
            Uc -= beta * np.outer(Uc @ v, v)
    alphas = [H[j, j] for j in range(keep)]
    betas = [H[j + 1, j] for j in range(keep)]
    return alphas, betas, U


def eigsolve_lanczos(A: Callable, x0, tol: float = 1e-14, krylovdim: int = 5, maxiter: int = 2,
                     eager: bool = False, which: str = "SR", ops=VecOps):
    """Returns (eigenvalue, eigenvector, info) with info = dict(converged, normres, numiter, numops)."""
    howmany = 1
    # ---- initialize (KrylovKit lanczos.jl `initialize`, orth = ModifiedGramSchmidt2)
    beta0 = ops.norm(x0)
    if beta0 == 0:
        raise ValueError("initial vector should not have norm zero")
    Ax0 = A(x0)
    alpha = ops.inner(x0, Ax0) / (beta0 * beta0)
    v = ops.scale(x0, 1.0 / beta0)
    r = ops.scale(Ax0, 1.0 / beta0)
    r = ops.add(r, v, -alpha)
    beta = ops.norm(r)
    dalpha = ops.inner(v, r)
    alpha = alpha + dalpha
    r = ops.add(r, v, -dalpha)
    beta = ops.norm(r)
    V: List = [v]
    alphas = [float(np.real(alpha))]
    betas = [beta]
    numops, numiter = 1, 1
    converged = 0
    D = U = f = None
    while True:
        beta = betas[-1]
        K = len(alphas)
        if K == krylovdim or beta <= tol or (eager and K >= howmany):
            T = np.diag(alphas)
            for j in range(K - 1):
                T[j, j + 1] = T[j + 1, j] = betas[j]
            if K == 1:
                D = np.array([T[0, 0]])
                U = np.eye(1)
                f = np.array([beta])
                converged = int(beta <= tol)
            else:
                D, U = np.linalg.eigh(T)
                p = np.argsort(D if which == "SR" else -D, kind="stable")
                D, U = D[p], U[:, p]
                f = U[K - 1, :] * beta
                converged = 0
                while converged < K and abs(f[converged]) <= tol:
                    converged += 1
            if converged >= howmany:
                break
        if K < krylovdim:
            # ---- expand! + lanczosrecurrence (ModifiedGramSchmidt2)
            bold = betas[-1]
            vnew = ops.scale(r, 1.0 / bold)
            V.append(vnew)
            w = A(vnew)
            numops += 1
            w = ops.add(w, V[-2], -bold)
            a = ops.inner(vnew, w)
            w = ops.add(w, vnew, -a)
            s = a
            for q in V:
                s = ops.inner(q, w)
                w = ops.add(w, q, -s)
            a = a + s
            b = ops.norm(w)
            alphas.append(float(np.real(a)))
            betas.append(b)
            r = w
        else:
            if numiter == maxiter:
                break
            keep = (3 * krylovdim + 2 * converged) // 5
            al, be, U2 = _restore_tridiagonal(D, f, U, keep)
            # basistransform!(B, U[:, 1:keep]);  B[keep+1] = r / beta
            newV = []
            for j in range(keep):
                acc = ops.scale(V[0], U2[0, j])
                for i in range(1, K):
                    acc = ops.add(acc, V[i], U2[i, j])
                newV.append(acc)
            rn = ops.scale(r, 1.0 / beta)
            V = newV
            alphas, betas = list(al), list(be)
            r = ops.scale(rn, betas[-1])          # shrink!: r <- r * normres
            numiter += 1
    K = len(alphas)
    vec = ops.scale(V[0], U[0, 0])
    for i in range(1, K):
        vec = ops.add(vec, V[i], U[i, 0])
    info = dict(converged=converged, normres=abs(f[0]), numiter=numiter, numops=numops)
    return float(D[0]), vec, info
