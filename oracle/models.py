"""Site sets, product/random QN MPS and exact Heisenberg MPOs (oracle; test infrastructure only).

Stands in for the model-building layer the reference delegates to ITensors `siteinds`,
`MPS(sites, states)` and `MPO(OpSum, sites)` (test/test_MPS_DMRG.jl:7-47).  Conventions used
throughout oracle/ and the CUDA library (they only need to be self-consistent -- the kernels
never see arrows):

* charges are integers in units of 2*Sz;  site index s: dir +1;
* MPS tensor A_j(l_{j-1}, s_j, l_j): dirs (+1, +1, -1), flux 0, q(l_j) = q(l_{j-1}) + q(s_j);
  l_0 and l_N are dim-1 boundary links (q(l_0) = 0, q(l_N) = total charge);
* MPO tensor W_j(w_{j-1}, s'_j, s_j, w_j): dirs (+1, +1, -1, -1), flux 0; w_0, w_N are dim-1.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from .blocksparse import BSTensor, Index, contract


# ------------------------------------------------------------------------------ sites
def spin_ops(S2: int):
    """Sz, S+, S-, Id for spin S = S2/2 in the basis m = S, S-1, ..., -S (ITensors "Up" first)."""
    d = S2 + 1
    S = S2 / 2.0
    m = np.array([S - k for k in range(d)])
    Sz = np.diag(m)
    Sp = np.zeros((d, d))
    for k in range(1, d):            # S+ |m_k> = c |m_{k-1}>
        Sp[k - 1, k] = np.sqrt(S * (S + 1) - m[k] * (m[k] + 1))
    return dict(Sz=Sz, Sp=Sp, Sm=Sp.T.copy(), Id=np.eye(d))


def siteinds(kind: str, N: int) -> List[Index]:
    """kind in {"S=1/2", "S=1"}: one dim-1 sector per Sz value, charge 2*Sz, highest first."""
    S2 = {"S=1/2": 1, "S=1": 2}[kind]
    qns = [(S2 - 2 * k,) for k in range(S2 + 1)]
    return [Index(qns, [1] * (S2 + 1), dir=+1, tags=f"Site,{kind},n={j + 1}") for j in range(N)]


def site_S2(s: Index) -> int:
    return s.nsect - 1


# ------------------------------------------------------------------------------ MPO
def heisenberg_mpo(sites: Sequence[Index], Jz: float = 1.0, Jxy: float = 1.0) -> List[BSTensor]:
    """H = sum_j Jz Sz_j Sz_{j+1} + (Jxy/2)(S+_j S-_{j+1} + S-_j S+_{j+1}), exact w=5 MPO
    (the Hamiltonian of every reference test, test/test_MPS_DMRG.jl:13-17).

    Automaton states and the charge they carry to the right:  F(0) 'finished', P(-2) 'S- pending
    partner S+', M(+2), Z(0), I(0) 'nothing yet'.  Link sectors are merged by charge, ascending:
    (-2:[P]), (0:[F,Z,I]), (+2:[M])."""
    N = len(sites)
    ops = spin_ops(site_S2(sites[0]))
    # dense automaton  Wd[a, b] = operator taking state a (left) to state b (right)
    F, P, M, Z, I = range(5)
    table = {(F, F): ops["Id"], (P, F): ops["Sp"], (M, F): ops["Sm"], (Z, F): ops["Sz"],
             (I, P): 0.5 * Jxy * ops["Sm"], (I, M): 0.5 * Jxy * ops["Sp"], (I, Z): Jz * ops["Sz"],
             (I, I): ops["Id"]}
    order = [P, F, Z, I, M]                       # position inside the merged link index
    bulk = Index([(-2,), (0,), (2,)], [1, 3, 1], dir=+1, tags="Link")
    pos = {st: k for k, st in enumerate(order)}
    links = [Index([(0,)], [1], dir=+1, tags="Link,l=0")]
    for j in range(1, N):
        links.append(bulk.sim().copy(tags=f"Link,l={j}"))
    links.append(Index([(0,)], [1], dir=+1, tags=f"Link,l={N}"))
    H = []
    d = sites[0].dim
    for j in range(N):
        wl, wr = links[j], links[j + 1]
        dense = np.zeros((wl.dim, d, d, wr.dim))
        for (a, b), op in table.items():
            if j == 0 and a != I:
                continue
            if j == N - 1 and b != F:
                continue
            ia = 0 if j == 0 else pos[a]
            ib = 0 if j == N - 1 else pos[b]
            dense[ia, :, :, ib] += op
        s = sites[j]
        inds = [wl.copy(dir=+1), s.prime().copy(dir=+1), s.copy(dir=-1), wr.copy(dir=-1)]
        H.append(BSTensor.from_dense(inds, dense))
    return H


def mpo_to_dense(H: Sequence[BSTensor]) -> np.ndarray:
    """Full 2^N x 2^N (d^N) matrix of an MPO (small N only; KAT helper)."""
    acc = None
    for W in H:
        Wd = W.to_dense()                          # (wl, s', s, wr)
        if acc is None:
            acc = Wd[0]                            # (s', s, wr)
        else:
            acc = np.tensordot(acc, Wd, axes=([-1], [0]))   # (..., s', s, wr)
    acc = acc[..., 0]
    n = acc.ndim // 2
    perm = [2 * k for k in range(n)] + [2 * k + 1 for k in range(n)]
    acc = np.transpose(acc, perm)
    D = int(np.prod(acc.shape[:n]))
    return acc.reshape(D, D)


# ------------------------------------------------------------------------------ MPS
def product_mps(sites: Sequence[Index], states: Sequence[int]) -> List[BSTensor]:
    """Product state; states[j] = sector number on site j (0 = highest Sz, "Up")."""
    N = len(sites)
    nq = len(sites[0].qns[0])
    q = (0,) * nq
    links = [Index([q], [1], dir=+1, tags="Link,l=0")]
    for j in range(N):
        q = tuple(a + b for a, b in zip(q, sites[j].qns[states[j]]))
        links.append(Index([q], [1], dir=+1, tags=f"Link,l={j + 1}"))
    psi = []
    for j in range(N):
        inds = [links[j].copy(dir=+1), sites[j].copy(dir=+1), links[j + 1].copy(dir=-1)]
        t = BSTensor(inds)
        t.blocks[(0, states[j], 0)] = np.ones((1, 1, 1))
        psi.append(t)
    return psi


def neel_mps(sites: Sequence[Index]) -> List[BSTensor]:
    """"Up" on odd sites, "Dn" on even sites (test/test_MPS_DMRG.jl:20)."""
    last = sites[0].nsect - 1
    return product_mps(sites, [0 if j % 2 == 0 else last for j in range(len(sites))])


def gaussian_link_sectors(chi: int, sigma: float, qmax: int, parity_offset: int = 0, step: int = 2):
    """Discretised Gaussian sector profile in total charge: charges
    q = parity_offset + step*k, |k| <= qmax, dims ~ exp(-(k)^2 / (2 sigma^2)) scaled to sum chi,
    every sector at least 1."""
    ks = np.arange(-qmax, qmax + 1)
    w = np.exp(-ks.astype(float) ** 2 / (2 * sigma ** 2))
    dims = np.maximum(1, np.floor(w / w.sum() * chi)).astype(int)
    dims[len(ks) // 2] += chi - dims.sum()
    return [(int(parity_offset + step * k),) for k in ks], [int(x) for x in dims]


def random_mps(sites: Sequence[Index], link_qns, link_dims, rng: np.random.Generator,
               total_q=(0,)) -> List[BSTensor]:
    """Random QN MPS: bulk links carry the given sectors clipped to what is reachable from both
    ends; all allowed blocks i.i.d. N(0,1).  Not normalised / not canonical."""
    N = len(sites)
    nq = len(sites[0].qns[0])
    # reachable charge sets from the left and from the right, with multiplicities
    def grow(cur, s, sign):
        out = {}
        for q, m in cur.items():
            for qs in s.qns:
                qq = tuple(a + sign * b for a, b in zip(q, qs))
                out[qq] = out.get(qq, 0) + m
        return out
    left = [{(0,) * nq: 1}]
    for j in range(N):
        left.append(grow(left[-1], sites[j], +1))
    right = [{tuple(total_q): 1}]
    for j in range(N - 1, -1, -1):
        right.append(grow(right[-1], sites[j], -1))
    right = right[::-1]
    links = []
    for j in range(N + 1):
        if j == 0:
            qs, ds = [(0,) * nq], [1]
        elif j == N:
            qs, ds = [tuple(total_q)], [1]
        else:
            qs, ds = [], []
            for q, d in zip(link_qns, link_dims):
                q = tuple(q)
                cap = min(left[j].get(q, 0), right[j].get(q, 0))
                if cap > 0:
                    qs.append(q)
                    ds.append(min(d, cap))
        links.append(Index(qs, ds, dir=+1, tags=f"Link,l={j}"))
    psi = []
    for j in range(N):
        inds = [links[j].copy(dir=+1), sites[j].copy(dir=+1), links[j + 1].copy(dir=-1)]
        psi.append(BSTensor.random(inds, rng))
    return psi


def mps_to_dense(psi: Sequence[BSTensor]) -> np.ndarray:
    acc = None
    for A in psi:
        Ad = A.to_dense()
        acc = Ad[0] if acc is None else np.tensordot(acc, Ad, axes=([-1], [0]))
    return acc[..., 0].reshape(-1)


def mps_norm(psi: Sequence[BSTensor]) -> float:
    E = None
    for A in psi:
        Ad = A.dag().prime(1, [A.inds[0], A.inds[2]])
        if E is None:
            # boundary: contract the two dim-1 left links with a 1x1 identity
            l = A.inds[0]
            one = BSTensor([l.dag(), l.prime()], {(0, 0): np.ones((1, 1))})
            E = one
        E = contract(contract(E, A), Ad)
    return float(np.sqrt(abs(E.to_dense().reshape(-1)[0])))


def maxlinkdim(psi: Sequence[BSTensor]) -> int:
    return max(A.inds[2].dim for A in psi[:-1]) if len(psi) > 1 else 1
