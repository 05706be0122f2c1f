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


# ---------------------------------------------------------------------------------------------------
# exponentiate -- restatement of KrylovKit.jl `exponentiate(A, t, x0; Lanczos)` = `expintegrator` with p = 1
# (third-party, not vendored; reached from `exp_solver`, /root/reference/src/base/solver.jl:66-88, with
# krylovdim=30, maxiter=100, tol=1e-12, eager=true).  Published algorithm (Niesen & Wright style phi-function
# integrator): u(t) = u0 + t*phi_1(tA) A u0.  One extra apply w1 = A u0 (beta = ||w1||), Lanczos factorisation
# started from w1 (`initialize` / `expand!` exactly as in eigsolve above), small dense exponential of the
# (K+2)x(K+2) augmented matrix [[s*dt*T, e1, 0], [0, 0, 1], [0, 0, 0]] whose column K+1 holds phi_1 and column
# K+2 holds phi_2, error estimate eps = |dt * beta * normres * expH[K, K+2]| against eta = tol/|t| per unit
# time, adaptive sub-stepping (safety factors delta = 1.2, gamma = 0.8) when the basis is full, eager exit at
# every K, first-correction term residual * expH[K, K+2].
def _aug_exp(alphas, betas, K, s_dt):
    from scipy.linalg import expm
    dt = np.result_type(np.asarray(s_dt).dtype, np.float64)
    H = np.zeros((K + 2, K + 2), dtype=dt)
    for j in range(K):
        H[j, j] = alphas[j] * s_dt
    for j in range(K - 1):
        H[j, j + 1] = H[j + 1, j] = betas[j] * s_dt
    H[0, K] = 1.0
    H[K, K + 1] = 1.0
    return expm(H)


def exponentiate(A: Callable, t, x0, tol: float = 1e-12, krylovdim: int = 30, maxiter: int = 100,
                 eager: bool = True, ops=VecOps):
    """Returns (exp(t*A) x0, info) with info = dict(converged, normres (= total error), numiter, numops)."""
    cplx = isinstance(t, complex) or np.iscomplexobj(t)
    t = complex(t) if cplx else float(t)
    tau = abs(t)
    if tau == 0:
        return ops.scale(x0, 1.0), dict(converged=1, normres=0.0, numiter=0, numops=0)
    sgn = t / tau
    tau0 = 0.0
    dtau = tau - tau0
    delta, gamma = 1.2, 0.8
    eta = tol / tau
    totalerr = 0.0
    w0 = ops.scale(x0, 1.0 + 0.0j if cplx else 1.0)
    w1 = A(w0)
    numops = 1
    beta = ops.norm(w1)
    if beta < tol:
        return w0, dict(converged=1, normres=beta, numiter=0, numops=numops)

    def lanczos_init(x):
        b0 = ops.norm(x)
        Ax = A(x)
        al = ops.inner(x, Ax) / (b0 * b0)
        v = ops.scale(x, 1.0 / b0)
        r = ops.scale(Ax, 1.0 / b0)
        r = ops.add(r, v, -al)
        da = ops.inner(v, r)
        al = al + da
        r = ops.add(r, v, -da)
        return [v], [float(np.real(al))], [ops.norm(r)], r

    V, alphas, betas, r = lanczos_init(w1)
    numops += 1
    numiter = 1

    def step(dt_):
        """small exponential, error estimate"""
        K = len(alphas)
        E = _aug_exp(alphas, betas, K, sgn * dt_)
        eps = abs(dt_ * beta * betas[-1] * E[K - 1, K + 1])
        return E, eps

    def take(E, dt_):
        nonlocal w0
        K = len(alphas)
        y = ops.scale(V[0], E[0, K])
        for i in range(1, K):
            y = ops.add(y, V[i], E[i, K])
        y = ops.add(y, r, E[K - 1, K + 1])           # first correction
        w0 = ops.add(w0, y, beta * sgn * dt_)

    while True:
        K = len(alphas)
        if K == krylovdim:
            dtau = min(dtau, tau - tau0)
            E, eps = step(dtau)
            omega = eps / (dtau * eta)
            q = K / 2
            while omega > 1:
                eps_prev, dtau_prev = eps, dtau
                dtau *= (gamma / omega) ** (1.0 / (q + 1))
                E, eps = step(dtau)
                omega = eps / (dtau * eta)
                q = max(0.0, np.log(eps / eps_prev) / np.log(dtau / dtau_prev) - 1)
            totalerr += eps
            take(E, dtau)
            tau0 += dtau
            if omega < gamma:
                dtau *= (gamma / omega) ** (1.0 / (q + 1)) if omega > 0 else delta
        elif betas[-1] <= (tau - tau0) * eta or eager:
            E, eps = step(tau - tau0)
            omega = eps / ((tau - tau0) * eta)
            if omega < 1:
                totalerr += eps
                take(E, tau - tau0)
                tau0 = tau
        if tau0 >= tau:
            return w0, dict(converged=1, normres=totalerr, numiter=numiter, numops=numops)
        if K < krylovdim:
            # expand! + lanczosrecurrence (ModifiedGramSchmidt2), as in eigsolve_lanczos
            bold = betas[-1]
            vnew = ops.scale(r, 1.0 / bold)
            V.append(vnew)
            w = A(vnew)
            numops += 1
            w = ops.add(w, V[-2], -bold)
            a = ops.inner(vnew, w)
            w = ops.add(w, vnew, -a)
            s = a
            for qv in V:
                s = ops.inner(qv, w)
                w = ops.add(w, qv, -s)
            a = a + s
            alphas.append(float(np.real(a)))
            betas.append(ops.norm(w))
            r = w
        else:
            if numiter == maxiter:
                return w0, dict(converged=0, normres=totalerr, numiter=numiter, numops=numops)
            numiter += 1
            w1 = A(w0)
            numops += 1
            beta = ops.norm(w1)
            if beta < tol:
                return w0, dict(converged=1, normres=beta, numiter=numiter, numops=numops)
            V, alphas, betas, r = lanczos_init(w1)
            numops += 1
