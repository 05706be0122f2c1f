"""Synthetic model inputs for benchmarks and smoke tests (host side, one-off; stands in for ITensors
`siteinds` / `MPO(OpSum)` / `MPS(sites, states)`, test/test_MPS_DMRG.jl:7-47).  Conventions: charges in units
of 2*Sz; A_j(l+, s+, r-), W_j(wl+, s'+, s-, wr-), all tensors flux 0, dim-1 boundary links."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from .tensor import HostTensor, Index


def spin_ops(S2: int):
    d = S2 + 1
    S = S2 / 2.0
    m = np.array([S - k for k in range(d)])
    Sp = np.zeros((d, d))
    for k in range(1, d):
        Sp[k - 1, k] = np.sqrt(S * (S + 1) - m[k] * (m[k] + 1))
    return dict(Sz=np.diag(m), Sp=Sp, Sm=Sp.T.copy(), Id=np.eye(d))


def siteinds(kind: str, N: int) -> List[Index]:
    S2 = {"S=1/2": 1, "S=1": 2}[kind]
    return [Index([(S2 - 2 * k,) for k in range(S2 + 1)], [1] * (S2 + 1), dir=+1, tags=f"Site,{kind},n={j + 1}")
            for j in range(N)]


def _from_dense(inds: Sequence[Index], dense: np.ndarray) -> HostTensor:
    t = HostTensor(inds)
    offs = [np.concatenate([[0], np.cumsum(ix.dims)]) for ix in inds]
    for c in np.ndindex(*[ix.nsect for ix in inds]):
        q = sum(ix.dir * ix.qns[k][0] for ix, k in zip(inds, c))
        blk = dense[tuple(slice(o[k], o[k + 1]) for o, k in zip(offs, c))]
        if q == 0 and np.any(blk != 0):
            t.blocks[tuple(c)] = np.array(blk)
    return t


def heisenberg_mpo(sites: Sequence[Index], Jz: float = 1.0, Jxy: float = 1.0) -> List[HostTensor]:
    """H = sum Jz SzSz + Jxy/2 (S+S- + S-S+): exact w=5 automaton, link sectors (-2:1, 0:3, +2:1)."""
    N = len(sites)
    ops = spin_ops(sites[0].nsect - 1)
    F, P, M, Z, I = range(5)
    table = {(F, F): ops["Id"], (P, F): ops["Sp"], (M, F): ops["Sm"], (Z, F): ops["Sz"],
             (I, P): 0.5 * Jxy * ops["Sm"], (I, M): 0.5 * Jxy * ops["Sp"], (I, Z): Jz * ops["Sz"], (I, I): ops["Id"]}
    pos = {P: 0, F: 1, Z: 2, I: 3, M: 4}
    links = [Index([(0,)], [1], tags="Link,l=0")]
    for j in range(1, N):
        links.append(Index([(-2,), (0,), (2,)], [1, 3, 1], tags=f"Link,l={j}"))
    links.append(Index([(0,)], [1], tags=f"Link,l={N}"))
    d = sites[0].dim
    H = []
    for j in range(N):
        wl, wr = links[j], links[j + 1]
        dense = np.zeros((wl.dim, d, d, wr.dim))
        for (a, b), op in table.items():
            if (j == 0 and a != I) or (j == N - 1 and b != F):
                continue
            dense[0 if j == 0 else pos[a], :, :, 0 if j == N - 1 else pos[b]] += op
        s = sites[j]
        H.append(_from_dense([wl.copy(dir=+1), s.prime().copy(dir=+1), s.copy(dir=-1), wr.copy(dir=-1)], dense))
    return H


def product_mps(sites: Sequence[Index], states: Sequence[int]) -> List[HostTensor]:
    q = 0
    links = [Index([(0,)], [1], tags="Link,l=0")]
    for j, s in enumerate(sites):
        q += s.qns[states[j]][0]
        links.append(Index([(q,)], [1], tags=f"Link,l={j + 1}"))
    return [HostTensor([links[j].copy(dir=+1), sites[j].copy(dir=+1), links[j + 1].copy(dir=-1)],
                       {(0, states[j], 0): np.ones((1, 1, 1))}) for j in range(len(sites))]


def neel_mps(sites):
    last = sites[0].nsect - 1
    return product_mps(sites, [0 if j % 2 == 0 else last for j in range(len(sites))])


def gaussian_link_sectors(chi: int, sigma: float, qmax: int, parity_offset: int = 0, step: int = 2):
    """Discretised Gaussian sector profile."""
    ks = np.arange(-qmax, qmax + 1)
    w = np.exp(-ks.astype(float) ** 2 / (2 * sigma ** 2))
    dims = np.maximum(1, np.floor(w / w.sum() * chi)).astype(int)
    dims[len(ks) // 2] += chi - dims.sum()
    return [(int(parity_offset + step * k),) for k in ks], [int(x) for x in dims]


def random_mps_links(sites: Sequence[Index], link_qns, link_dims, total_q: int = 0) -> List[Index]:
    """Link indices of a QN MPS with the given bulk sector profile, clipped to what is reachable from both
    ends (so that near the edges the bond dimension grows like d^j)."""
    N = len(sites)

    def grow(cur, s, sign):
        out = {}
        for q, m in cur.items():
            for qs in s.qns:
                out[q + sign * qs[0]] = out.get(q + sign * qs[0], 0) + m
        return out
    left = [{0: 1}]
    for j in range(N):
        left.append(grow(left[-1], sites[j], +1))
    right = [{total_q: 1}]
    for j in range(N - 1, -1, -1):
        right.append(grow(right[-1], sites[j], -1))
    right = right[::-1]
    links = []
    for j in range(N + 1):
        if j == 0:
            qs, ds = [(0,)], [1]
        elif j == N:
            qs, ds = [(total_q,)], [1]
        else:
            qs, ds = [], []
            for q, d in zip(link_qns, link_dims):
                cap = min(left[j].get(q[0], 0), right[j].get(q[0], 0))
                if cap > 0:
                    qs.append(tuple(q))
                    ds.append(min(d, cap))
        links.append(Index(qs, ds, tags=f"Link,l={j}"))
    return links
