"""Exact diagonalisation helpers for the physics known-answer tests."""
import numpy as np


def spin_ops(S2):
    d = S2 + 1
    S = S2 / 2.0
    m = np.array([S - k for k in range(d)])
    Sp = np.zeros((d, d))
    for k in range(1, d):
        Sp[k - 1, k] = np.sqrt(S * (S + 1) - m[k] * (m[k] + 1))
    return np.diag(m), Sp, Sp.T.copy()


def heisenberg_dense(N, S2):
    Sz, Sp, Sm = spin_ops(S2)
    d = S2 + 1

    def at(op, j):
        out = np.eye(1)
        for k in range(N):
            out = np.kron(out, op if k == j else np.eye(d))
        return out
    H = np.zeros((d ** N, d ** N))
    for j in range(N - 1):
        H += at(Sz, j) @ at(Sz, j + 1) + 0.5 * (at(Sp, j) @ at(Sm, j + 1) + at(Sm, j) @ at(Sp, j + 1))
    return H


def sz0_sector(N, S2):
    """Indices of the total-Sz = 0 basis states (site 1 = most significant digit)."""
    d = S2 + 1
    idx = np.arange(d ** N)
    tot = np.zeros(d ** N)
    for k in range(N):
        digit = (idx // d ** (N - 1 - k)) % d
        tot += (S2 - 2 * digit)
    return np.nonzero(tot == 0)[0]


def lowest_energies(N, S2, k=2):
    H = heisenberg_dense(N, S2)
    sel = sz0_sector(N, S2)
    w = np.linalg.eigvalsh(H[np.ix_(sel, sel)])
    return w[:k]
