"""CouplingModel + ProjCouplingModel on the CPU (oracle; test infrastructure only).

Restates, term by term ("id" by "id"), the in-tree reference code
  /root/reference/src/base/couplingmodel.jl:14-17,120-230   CouplingModel, _initCouplingModel (merge = true | false)
  /root/reference/src/base/helper_internal_funcs.jl:41-63    _add_oplinks!  (dim-1 OpLinks carrying the flux)
  /root/reference/src/mps/projcouplingmodel.jl:123-196       _makeL!   (212-286 _makeR!, mirror image)
  /root/reference/src/mps/projcouplingmodel.jl:315-383       _contract / product
  /root/reference/src/mps/projcouplingmodel.jl:391-492       noiseterm
with the oracle's block-sparse `contract`.  The model constructor covers what the tests need (products of
one- and two-site spin operators); merged terms are the exact direct sum of the collected terms WITHOUT the
SVD recompression of `_chunksum_cm_terms` (:60-118), which only changes the gauge on the OpLinks, not the
operator.  `ITensors.contract(v, tensors...; sequence)` is evaluated left to right -- the sequence changes
the flop count, not the result.
"""
from __future__ import annotations

import itertools
from typing import Dict, List, Sequence

import numpy as np

from .blocksparse import BSTensor, Index, contract, matricize
from .models import site_S2, spin_ops

_ids = itertools.count(1000)


def gen_id() -> int:
    return next(_ids)


class CouplingModel:
    """sites + terms[site] = {id: tensor}  (IDTensors, src/base/typedef.jl:27)."""

    def __init__(self, sites: Sequence[Index], terms: List[Dict[int, BSTensor]]):
        self.sites = list(sites)
        self.terms = terms

    def __len__(self):
        return len(self.sites)

    def __getitem__(self, j):            # 1-based like the reference
        return self.terms[j - 1]


_DENSE_OPS = {"X": np.array([[0.0, 1.0], [1.0, 0.0]]), "Z": np.array([[1.0, 0.0], [0.0, -1.0]]), "Id": np.eye(2),
              "Sz": np.array([[0.5, 0.0], [0.0, -0.5]]), "Sx": np.array([[0.0, 0.5], [0.5, 0.0]]),
              "S+": np.array([[0.0, 1.0], [0.0, 0.0]]), "S-": np.array([[0.0, 0.0], [1.0, 0.0]])}


def _op_tensor(name: str, s: Index) -> BSTensor:
    if s.nsect == 1:                                        # site without quantum numbers (d = 2): dense operator
        return BSTensor.from_dense([s.prime().copy(dir=+1), s.copy(dir=-1)], _DENSE_OPS[name], keep_zero_blocks=True)
    ops = spin_ops(site_S2(s))
    dense = {"Sz": ops["Sz"], "S+": ops["Sp"], "S-": ops["Sm"], "Id": ops["Id"]}[name]
    step = s.qns[0][0] - s.qns[1][0]                       # charge carried by one spin flip (2 in 2Sz units)
    flux = {"Sz": 0, "Id": 0, "S+": step, "S-": -step}[name]
    return BSTensor.from_dense([s.prime().copy(dir=+1), s.copy(dir=-1)], dense, flux=(flux,))


def _flux(t: BSTensor):
    f = t.flux()
    return tuple(0 for _ in t.inds[0].qns[0]) if f is None else f


def _add_oplinks(tensors: List[BSTensor]) -> List[BSTensor]:
    """helper_internal_funcs.jl:41-63: link b between tensors b and b+1 carries the flux of tensors 1..b."""
    n = len(tensors)
    out = list(tensors)
    lflux = tuple(0 for _ in _flux(tensors[0]))
    for b in range(n - 1):
        lflux = tuple(a + c for a, c in zip(lflux, _flux(tensors[b])))
        link = Index([lflux], [1], dir=-1, tags="OpLink")          # dag(Index(lflux => 1))
        one = BSTensor([link], {(0,): np.ones(1)})
        out[b] = _outer(out[b], one)
        out[b + 1] = _outer(BSTensor([link.copy(dir=+1)], {(0,): np.ones(1)}), out[b + 1])
    return out


def _outer(a: BSTensor, b: BSTensor) -> BSTensor:
    return contract(a, b)                # no common indices: outer product


def _directsum_chain(terms: List[List[BSTensor]]) -> List[BSTensor]:
    """Exact direct sum of terms with the same support over their OpLinks (`_directsum`,
    src/base/helper_internal_funcs.jl:73-110; ITensors `directsum` keeps the sectors of the summands apart): link b of
    the sum has one dim-1 sector per term; tensor b of term k sits in sector (k, k) of its two links.
    Index order from `_add_oplinks`: first (s', s, r), middle (l, s', s, r), last (l, s', s)."""
    n = len(terms[0])
    K = len(terms)
    links = []
    for bnd in range(n - 1):
        qns = [t[bnd].inds[-1].qns[0] for t in terms]
        links.append(Index(qns, [1] * K, dir=-1, tags="OpLink"))
    out = []
    for pos in range(n):
        t0 = terms[0][pos]
        sp = next(ix for ix in t0.inds if "OpLink" not in ix.tags and ix.plev == 1)
        sk = next(ix for ix in t0.inds if "OpLink" not in ix.tags and ix.plev == 0)
        inds = ([links[pos - 1].copy(dir=+1)] if pos > 0 else []) + [sp, sk] + ([links[pos]] if pos < n - 1 else [])
        T = BSTensor(inds)
        for k, term in enumerate(terms):
            for c, blk in term[pos].blocks.items():
                site_c = c[1:3] if pos > 0 else c[0:2]
                cc = ((k,) if pos > 0 else ()) + tuple(site_c) + ((k,) if pos < n - 1 else ())
                T.blocks[cc] = blk.copy()
        out.append(T)
    return out


def coupling_model(os: Sequence, sites: Sequence[Index], merge: bool = True) -> CouplingModel:
    """`CouplingModel(os::OpStrings, sites; merge)`: os = [(coeff, (opname, site), (opname, site), ...), ...]."""
    N = len(sites)
    collected: Dict[tuple, List[List[BSTensor]]] = {}
    for term in os:
        coeff, ops = term[0], term[1:]
        if coeff == 0:
            continue
        pos = tuple(p for _, p in ops)
        assert list(pos) == sorted(set(pos)) and 1 <= pos[0] and pos[-1] <= N
        tensors = [_op_tensor(name, sites[p - 1]) for name, p in ops]
        a = abs(coeff) ** (1.0 / len(ops))
        tensors = [t.scale(a) for t in tensors]
        tensors[0] = tensors[0].scale(np.sign(coeff))
        tensors = _add_oplinks(tensors)
        if pos not in collected:
            collected[pos] = [tensors]
        elif len(pos) == 1:
            collected[pos][0] = [collected[pos][0][0].add(tensors[0])]
        else:
            collected[pos].append(tensors)
    terms: List[Dict[int, BSTensor]] = [dict() for _ in range(N)]
    for pos, lst in collected.items():
        if merge and len(lst) > 1:
            lst = [_directsum_chain(lst)]
        for tensors in lst:
            tid = gen_id()
            for p, t in zip(pos, tensors):
                terms[p - 1][tid] = t
    return CouplingModel(sites, terms)


def heisenberg_coupling_model(sites: Sequence[Index], Jz: float = 1.0, Jxy: float = 1.0, merge: bool = True,
                              field: float = 0.0, j2: float = 0.0) -> CouplingModel:
    """The Hamiltonian of the reference tests (test/test_MPS_DMRG.jl:28-36) as OpStrings; optional Sz field
    (one-site terms) and next-nearest-neighbour coupling j2 (terms that skip a site)."""
    os = []
    N = len(sites)
    for j in range(1, N):
        os += [(Jz, ("Sz", j), ("Sz", j + 1)), (0.5 * Jxy, ("S+", j), ("S-", j + 1)), (0.5 * Jxy, ("S-", j), ("S+", j + 1))]
    for j in range(1, N - 1):
        if j2:
            os += [(j2, ("Sz", j), ("Sz", j + 2)), (0.5 * j2, ("S+", j), ("S-", j + 2)), (0.5 * j2, ("S-", j), ("S+", j + 2))]
    for j in range(1, N + 1):
        if field:
            os += [(field, ("Sz", j))]
    return coupling_model(os, sites, merge)


def tfi_coupling_model(sites: Sequence[Index], h: float = 1.0, J: float = 1.0, merge: bool = True) -> CouplingModel:
    """H = -J sum Z_j Z_{j+1} - h sum X_j on sites without quantum numbers (BASELINE.json configs[4])."""
    os = []
    N = len(sites)
    for j in range(1, N):
        os.append((-J, ("Z", j), ("Z", j + 1)))
    for j in range(1, N + 1):
        os.append((-h, ("X", j)))
    return coupling_model(os, sites, merge)


def coupling_model_to_dense(M: CouplingModel) -> np.ndarray:
    """Dense matrix of the model (small N; KAT helper)."""
    N = len(M)
    d = [s.dim for s in M.sites]
    D = int(np.prod(d))
    H = np.zeros((D, D))
    ids = set()
    for t in M.terms:
        ids |= set(t)
    for tid in ids:
        acc = None                      # running (rows, cols, link) array
        for j in range(N):
            t = M.terms[j].get(tid)
            if t is None:
                op = np.eye(d[j])[None, :, :, None] if acc is None or acc.shape[2] == 1 else None
                if op is None:
                    w = acc.shape[2]
                    op = np.einsum("ab,xy->axyb", np.eye(w), np.eye(d[j]))
            else:
                sp = next(ix for ix in t.inds if "OpLink" not in ix.tags and ix.plev == 1)
                sk = next(ix for ix in t.inds if "OpLink" not in ix.tags and ix.plev == 0)
                links = [ix for ix in t.inds if "OpLink" in ix.tags]
                prev = None if acc is None else acc
                wl = [ix for ix in links if ix.dir == +1]
                wr = [ix for ix in links if ix.dir == -1]
                td = t.permute(wl + [sp, sk] + wr).to_dense()
                if not wl:
                    td = td[None]
                if not wr:
                    td = td[..., None]
                op = td
            if acc is None:
                acc = op[0] if op.shape[0] == 1 else None
                assert acc is not None
            else:
                if acc.shape[2] != op.shape[0]:
                    assert acc.shape[2] == 1 or op.shape[0] == 1
                acc = np.einsum("rcw,wxyv->rxcyv", acc, op).reshape(acc.shape[0] * op.shape[1], acc.shape[1] * op.shape[2], op.shape[3])
        assert acc.shape[2] == 1
        H += acc[:, :, 0]
    return H


class ProjCouplingModel:
    def __init__(self, M: CouplingModel):
        self.M = M
        self.N = len(M)
        self.lpos = 0
        self.rpos = self.N + 1
        self.nsite = 2
        self.LR: List[Dict[int, BSTensor] | None] = [None] * self.N

    def set_nsite(self, n):
        self.nsite = n

    def site_range(self):
        return range(self.lpos + 1, self.rpos)

    def lproj(self) -> Dict[int, BSTensor]:
        return {} if self.lpos <= 0 else self.LR[self.lpos - 1]

    def rproj(self) -> Dict[int, BSTensor]:
        return {} if self.rpos >= self.N + 1 else self.LR[self.rpos - 1]

    # -- environment update (projcouplingmodel.jl:123-196 / 212-286)
    def _step(self, E: Dict[int, BSTensor], site: int, psi, left: bool) -> Dict[int, BSTensor]:
        phi = psi[site - 1]
        Ms = self.M[site]
        new: Dict[int, BSTensor] = {}
        local = None
        out_link = phi.inds[2] if left else phi.inds[0]       # commonind(phi, next tensor along the sweep)
        for tid in list(dict.fromkeys(list(E) + list(Ms))):
            if tid in E and tid in Ms:
                t = phi.prime().dag()
                t = contract(t, E[tid])
                t = contract(t, Ms[tid])
                t = contract(t, phi)
            else:
                unc = E[tid] if tid in E else Ms[tid]
                ind1 = next(ix for ix in phi.inds if ix in set(unc.inds))
                t = phi.prime(1, [ind1, out_link]).dag()
                t = contract(t, unc)
                t = contract(t, phi)
            if t.rank > 2:
                new[tid] = t
            else:
                local = t if local is None else local.add(t)
        if local is not None:
            new[gen_id()] = local
        return new

    def makeL(self, psi, k):
        ll = self.lpos
        if ll >= k:
            self.lpos = k
            return
        ll = max(ll, 0)
        L = self.lproj()
        while ll < k:
            L = self._step(L, ll + 1, psi, True)
            self.LR[ll] = L
            ll += 1
        self.lpos = k

    def makeR(self, psi, k):
        rl = self.rpos
        if rl <= k:
            self.rpos = k
            return
        rl = min(rl, self.N + 1)
        R = self.rproj()
        while rl > k:
            R = self._step(R, rl - 1, psi, False)
            self.LR[rl - 2] = R
            rl -= 1
        self.rpos = k

    def position(self, psi, pos):
        self.makeL(psi, pos - 1)
        self.makeR(psi, pos + self.nsite)
        # outer links of the site range ("Link,l=first-1" / "Link,l=last" in the reference's tag lookup, :430,470)
        self._link_l = psi[self.lpos].inds[0]
        self._link_r = psi[self.rpos - 2].inds[2]

    # -- product (projcouplingmodel.jl:315-383)
    def _group(self, maps):
        idtens: Dict[int, List[BSTensor]] = {}
        for it in maps:
            for tid, t in it.items():
                idtens.setdefault(tid, []).append(t)
        return idtens

    def product(self, v: BSTensor) -> BSTensor:
        maps = [self.lproj()] + [self.M[j] for j in self.site_range()] + [self.rproj()]
        out = None
        for tid, ts in self._group(maps).items():
            Hv = v
            for t in ts:
                Hv = contract(Hv, t)
            Hv = Hv.noprime()
            out = Hv if out is None else out.add(Hv)
        if out.rank != v.rank:
            raise RuntimeError("The order of the ProjCouplingModel-ITensor product P*v is not equal to the order of v")
        return out

    __call__ = product

    # -- noise term (projcouplingmodel.jl:391-492)
    def noiseterm(self, phi: BSTensor, ortho: str) -> BSTensor:
        if self.nsite != 2:
            raise RuntimeError("noise term only defined for 2-site ProjMPO")
        sr = list(self.site_range())
        if ortho == "left":
            maps, site, link = [self.lproj(), self.M[sr[0]]], sr[0], self._link_l
        elif ortho == "right":
            maps, site, link = [self.rproj(), self.M[sr[-1]]], sr[-1], self._link_r
        else:
            raise ValueError(f"In noiseterm, got ortho = {ortho}, only supports `left` and `right`")
        s = self.M.sites[site - 1]
        nt = None
        for tid, ts in self._group(maps).items():
            t = phi
            for x in ts:
                t = contract(t, x)
            # setprime!: OpLinks -> 0, the kept link and the kept site -> 1
            inds = []
            for ix in t.inds:
                if "OpLink" in ix.tags:
                    inds.append(ix.copy(plev=0))
                elif ix.id == link.id or ix.id == s.id:
                    inds.append(ix.copy(plev=1))
                else:
                    inds.append(ix)
            t = t._with_inds(inds)
            d = contract(t, t.noprime().dag())
            nt = d if nt is None else nt.add(d)
        return nt
