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


# =====================================================================================================
# Generic finite-state-automaton MPOs: sums of one-site and two-site terms with (optional) strings in between.
# Stands in for ITensors `MPO(OpSum, sites)` (reference call sites test/test_MPS_DMRG.jl:36-47) including the
# Jordan-Wigner strings the reference inserts with `bosonize` (src/base/fermions.jl:30-96): here the strings are
# explicit operators of the automaton, so every tensor downstream is an ordinary (bosonic) block-sparse tensor.
# BASELINE.json configs[2] (J1-J2 cylinder), configs[3] (Hubbard ladder) and configs[4] (TFI) are built with it.
# =====================================================================================================
def _charge_of_op(site: Index, op: np.ndarray):
    """Charge an operator adds: q(s') - q(s) of its non-zero entries (must be unique); op[s', s]."""
    offs = np.concatenate([[0], np.cumsum(site.dims)])
    dq = None
    for a in range(site.nsect):
        for b in range(site.nsect):
            if np.any(op[offs[a]:offs[a + 1], offs[b]:offs[b + 1]] != 0):
                q = tuple(x - y for x, y in zip(site.qns[a], site.qns[b]))
                if dq is not None and q != dq:
                    raise ValueError("operator does not carry a definite charge")
                dq = q
    return dq if dq is not None else (0,) * len(site.qns[0])


def from_dense(inds: Sequence[Index], dense: np.ndarray) -> HostTensor:
    """Flux-0 blocks of a dense array (any number of charges per sector); zero blocks are dropped."""
    t = HostTensor(inds)
    nq = len(inds[0].qns[0])
    offs = [np.concatenate([[0], np.cumsum(ix.dims)]) for ix in inds]
    for c in np.ndindex(*[ix.nsect for ix in inds]):
        q = [sum(ix.dir * ix.qns[k][a] for ix, k in zip(inds, c)) for a in range(nq)]
        blk = dense[tuple(slice(o[k], o[k + 1]) for o, k in zip(offs, c))]
        if not any(q) and np.any(blk != 0):
            t.blocks[tuple(c)] = np.array(blk)
    return t


def automaton_mpo(sites: Sequence[Index], ops, onsite=(), pairs=()) -> List[HostTensor]:
    """H = sum_(j, c, A) c A_j  +  sum_(i<j, c, A, B, S) c A_i S_{i+1} ... S_{j-1} B_j  as an exact MPO.

    `ops(name, j)` returns the dense d x d matrix [s', s] of operator `name` on site j (0-based);
    `onsite` = [(j, coef, name)], `pairs` = [(i, j, coef, nameA, nameB, nameS)] with i < j.
    Automaton states on the link right of site j: I (nothing yet), F (finished) and, for every channel (A, S) and
    every start site i <= j that still has a partner beyond j, "A applied at i, strings since".  Link states are
    sorted by the charge they carry (= the charge of A) into QN sectors, ascending."""
    N = len(sites)
    nq = len(sites[0].qns[0])
    zero = (0,) * nq
    pairs = [p for p in pairs if p[2] != 0.0]
    # states per link l = 0..N (link l sits left of site l)
    link_states: List[list] = [[] for _ in range(N + 1)]
    for l in range(N + 1):
        st = []
        if l < N:
            st.append(("I",))
        if l > 0:
            st.append(("F",))
        seen = set()
        for (i, j, c, A, B, S) in pairs:
            if i < l <= j and (A, S, i) not in seen:
                seen.add((A, S, i))
                st.append(("P", A, S, i))
        link_states[l] = st
    chan_q = {}
    for (i, j, c, A, B, S) in pairs:
        chan_q[(A, S, i)] = _charge_of_op(sites[i], ops(A, i))

    def state_q(s):
        return zero if s[0] in ("I", "F") else chan_q[(s[1], s[2], s[3])]
    links, order = [], []
    for l in range(N + 1):
        st = link_states[l]
        idx = sorted(range(len(st)), key=lambda k: (state_q(st[k]), k))
        qs, ds = [], []
        for k in idx:
            q = state_q(st[k])
            if qs and qs[-1] == q:
                ds[-1] += 1
            else:
                qs.append(q); ds.append(1)
        links.append(Index(qs, ds, dir=+1, tags=f"Link,l={l}"))
        order.append({st[k]: n for n, k in enumerate(idx)})
    H = []
    for j in range(N):
        d = sites[j].dim
        wl, wr = links[j], links[j + 1]
        dense = np.zeros((wl.dim, d, d, wr.dim))
        pl, pr = order[j], order[j + 1]
        Id = np.eye(d)
        if ("I",) in pl and ("I",) in pr:
            dense[pl[("I",)], :, :, pr[("I",)]] += Id
        if ("F",) in pl and ("F",) in pr:
            dense[pl[("F",)], :, :, pr[("F",)]] += Id
        if ("I",) in pl and ("F",) in pr:
            for (jj, c, A) in onsite:
                if jj == j:
                    dense[pl[("I",)], :, :, pr[("F",)]] += c * ops(A, j)
        started = set()
        for (i, jj, c, A, B, S) in pairs:
            if i == j and (A, S) not in started:                     # I -> pending
                started.add((A, S))
                dense[pl[("I",)], :, :, pr[("P", A, S, i)]] += ops(A, j)
            if jj == j:                                              # pending -> F
                dense[pl[("P", A, S, i)], :, :, pr[("F",)]] += c * ops(B, j)
        for s in pl:                                                 # pending passes through with its string
            if s[0] == "P" and s in pr:
                dense[pl[s], :, :, pr[s]] += ops(s[2], j)
        s = sites[j]
        H.append(from_dense([wl.copy(dir=+1), s.prime().copy(dir=+1), s.copy(dir=-1), wr.copy(dir=-1)], dense))
    return H


# ------------------------------------------------------------------------------ spin models on graphs
def spin_op_table(S2: int):
    t = spin_ops(S2)
    return lambda name, j: t[name]


def heisenberg_bonds_mpo(sites: Sequence[Index], bonds) -> List[HostTensor]:
    """H = sum_(i, j, J) J [Sz_i Sz_j + (S+_i S-_j + S-_i S+_j)/2] for arbitrary bonds (i < j, 0-based)."""
    pairs = []
    for (i, j, J) in bonds:
        i, j = min(i, j), max(i, j)
        pairs += [(i, j, J, "Sz", "Sz", "Id"), (i, j, 0.5 * J, "Sp", "Sm", "Id"), (i, j, 0.5 * J, "Sm", "Sp", "Id")]
    return automaton_mpo(sites, spin_op_table(sites[0].nsect - 1), (), pairs)


def j1j2_cylinder_bonds(Lx: int, Ly: int, J1: float = 1.0, J2: float = 0.5):
    """J1-J2 Heisenberg model on an Lx x Ly cylinder (periodic along y, open along x), site (x, y) -> x*Ly + y
    (column-major snake-free ordering): nearest neighbours J1, next-nearest (diagonal) neighbours J2
    (BASELINE.json configs[2]: width-6 cylinder, MPO bond dimension ~30)."""
    def n(x, y):
        return x * Ly + (y % Ly)
    bonds = {}

    def add(a, b, J):
        if a == b:
            return
        k = (min(a, b), max(a, b))
        bonds[k] = bonds.get(k, 0.0) + J
    for x in range(Lx):
        for y in range(Ly):
            if Ly > 2 or y + 1 < Ly:
                add(n(x, y), n(x, y + 1), J1)
            if x + 1 < Lx:
                add(n(x, y), n(x + 1, y), J1)
                if Ly > 2 or y + 1 < Ly:
                    add(n(x, y), n(x + 1, y + 1), J2)
                if Ly > 2 or y - 1 >= 0:
                    add(n(x, y), n(x + 1, y - 1), J2)
    return [(a, b, J) for (a, b), J in sorted(bonds.items())]


# ------------------------------------------------------------------------------ fermions (Hubbard)
def electron_siteinds(N: int) -> List[Index]:
    """ITensors "Electron" sites with conserve_qns: states |0>, |up>, |dn>, |updn>, charges (Nf, 2 Sz)."""
    return [Index([(0, 0), (1, 1), (1, -1), (2, 0)], [1, 1, 1, 1], dir=+1, tags=f"Site,Electron,n={j + 1}")
            for j in range(N)]


def electron_ops():
    """Local operators of an "Electron" site in the basis (0, up, dn, updn), |updn> = c+_up c+_dn |0>; ITensors'
    "Cup", "Cdn" (= F_up A_dn), "F" (site parity), "Nup", "Ndn", "Nupdn"."""
    Cup = np.zeros((4, 4)); Cup[0, 1] = 1.0; Cup[2, 3] = 1.0
    Cdn = np.zeros((4, 4)); Cdn[0, 2] = 1.0; Cdn[1, 3] = -1.0
    F = np.diag([1.0, -1.0, -1.0, 1.0])
    t = dict(Id=np.eye(4), F=F, Cup=Cup, Cdn=Cdn, Cdagup=Cup.T.copy(), Cdagdn=Cdn.T.copy())
    t["Nup"] = t["Cdagup"] @ Cup
    t["Ndn"] = t["Cdagdn"] @ Cdn
    t["Nupdn"] = t["Nup"] @ t["Ndn"]
    # hopping factors with the Jordan-Wigner string of the LEFT site absorbed (i < j):
    #   c+_i c_j = (C+ F)_i F_{i+1} ... F_{j-1} C_j ,   c+_j c_i = (F C)_i F ... F C+_j
    for s in ("up", "dn"):
        t["CdagF" + s] = t["Cdag" + s] @ F
        t["FC" + s] = F @ t["C" + s]
    return t


def hubbard_mpo(sites: Sequence[Index], bonds, t: float = 1.0, U: float = 4.0) -> List[HostTensor]:
    """H = -t sum_{<ij>, s} (c+_{is} c_{js} + h.c.) + U sum_i n_up n_dn with explicit Jordan-Wigner strings."""
    tb = electron_ops()
    pairs = []
    for (i, j) in bonds:
        i, j = min(i, j), max(i, j)
        for s in ("up", "dn"):
            pairs.append((i, j, -t, "CdagF" + s, "C" + s, "F"))
            pairs.append((i, j, -t, "FC" + s, "Cdag" + s, "F"))
    onsite = [(j, U, "Nupdn") for j in range(len(sites))]
    return automaton_mpo(sites, lambda name, j: tb[name], onsite, pairs)


def ladder_bonds(L: int, legs: int = 2):
    """Nearest-neighbour bonds of an L-rung ladder, site (x, leg) -> x*legs + leg (BASELINE.json configs[3])."""
    b = []
    for x in range(L):
        for a in range(legs):
            if a + 1 < legs:
                b.append((x * legs + a, x * legs + a + 1))
            if x + 1 < L:
                b.append((x * legs + a, (x + 1) * legs + a))
    return b


def product_mps_q(sites: Sequence[Index], states: Sequence[int]) -> List[HostTensor]:
    """Product state for sites with any number of charges."""
    nq = len(sites[0].qns[0])
    q = (0,) * nq
    links = [Index([q], [1], tags="Link,l=0")]
    for j, s in enumerate(sites):
        q = tuple(a + b for a, b in zip(q, s.qns[states[j]]))
        links.append(Index([q], [1], tags=f"Link,l={j + 1}"))
    return [HostTensor([links[j].copy(dir=+1), sites[j].copy(dir=+1), links[j + 1].copy(dir=-1)],
                       {(0, states[j], 0): np.ones((1, 1, 1))}) for j in range(len(sites))]


def random_mps_links_q(sites: Sequence[Index], total_q, dim_of) -> List[Index]:
    """Link indices of a QN MPS for sites with any number of charges: at link j every charge reachable from both
    ends gets dimension min(dim_of(j, q), reachable multiplicity), sectors ascending in charge."""
    N = len(sites)
    nq = len(sites[0].qns[0])

    def grow(cur, s, sign):
        out = {}
        for q, m in cur.items():
            for qs, ds in zip(s.qns, s.dims):
                qq = tuple(a + sign * b for a, b in zip(q, qs))
                out[qq] = min(out.get(qq, 0) + m * ds, 1 << 40)
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
        qs, ds = [], []
        for q in sorted(set(left[j]) & set(right[j])):
            d = min(int(dim_of(j, q)), left[j][q], right[j][q])
            if d > 0:
                qs.append(q); ds.append(d)
        links.append(Index(qs, ds, tags=f"Link,l={j}"))
    return links
