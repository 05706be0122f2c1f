"""Projected-MPO environments on the CPU (oracle; test infrastructure only).

Restates ITensorMPS `ProjMPO` (third-party, not vendored; reached from `StateEnvs{ProjMPO}`
at /root/reference/src/mps/state_envs.jl:54-60,364-378) in the exact contraction order the
reference executes:

* `product(v)`   = noprime(v * L * W_j [* W_{j+1}] * R)                 
* `_makeL!`      : L_j = L_{j-1} * psi[j] * H[j] * dag(prime(psi[j]))    
* `_makeR!`      : mirror image;  watermarks lpos / rpos as in
                   src/mps/projcouplingmodel.jl:123-129,212-218 (same contract as ProjMPO)
* `noiseterm`    : nt = L*W_j*phi (ortho left) | phi*W_{j+1}*R (right); nt * dag(noprime(nt)).

The same pattern is visible in-tree in `ProjMPS2.contract` (src/mps/projmps2.jl:159-173).
Boundary environments are explicit 1x1x1 tensors on the dim-1 boundary links (oracle/models.py)
instead of ITensors' `OneITensor`, which changes no arithmetic.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from .blocksparse import BSTensor, Index, contract, matricize


def _edge(A: BSTensor, W: BSTensor, left: bool) -> BSTensor:
    """1x1x1 boundary environment facing tensor A / MPO tensor W on their outer links."""
    k = 0 if left else -1
    a, w = A.inds[k], W.inds[k]
    inds = [a.prime().copy(dir=a.dir), w.copy(dir=-w.dir), a.copy(dir=-a.dir)]
    return BSTensor(inds, {(0, 0, 0): np.ones((1, 1, 1))})


class ProjMPO:
    def __init__(self, H: Sequence[BSTensor]):
        self.H = list(H)
        self.N = len(H)
        self.lpos = 0
        self.rpos = self.N + 1
        self.nsite = 2
        self.LR: List[BSTensor | None] = [None] * self.N     # LR[j-1] holds L_j or R_j (1-based j)
        self._Ledge = None
        self._Redge = None

    # -- accessors (1-based positions as in the reference)
    def site_range(self):
        return range(self.lpos + 1, self.rpos)

    def lproj(self):
        return self._Ledge if self.lpos <= 0 else self.LR[self.lpos - 1]

    def rproj(self):
        return self._Redge if self.rpos >= self.N + 1 else self.LR[self.rpos - 1]

    def set_nsite(self, n: int):
        self.nsite = n

    # -- environment update
    def _ensure_edges(self, psi):
        if self._Ledge is None or self._Ledge.inds[2] != psi[0].inds[0]:
            self._Ledge = _edge(psi[0], self.H[0], True)
        if self._Redge is None or self._Redge.inds[2] != psi[-1].inds[-1]:
            self._Redge = _edge(psi[-1], self.H[-1], False)

    def makeL(self, psi, k: int):
        self._ensure_edges(psi)
        ll = self.lpos
        if ll >= k:
            self.lpos = k
            return
        ll = max(ll, 0)
        L = self.lproj()
        while ll < k:
            A = psi[ll]
            L = contract(contract(contract(L, A), self.H[ll]), A.prime().dag())
            self.LR[ll] = L
            ll += 1
        self.lpos = k

    def makeR(self, psi, k: int):
        self._ensure_edges(psi)
        rl = self.rpos
        if rl <= k:
            self.rpos = k
            return
        rl = min(rl, self.N + 1)
        R = self.rproj()
        while rl > k:
            A = psi[rl - 2]
            R = contract(contract(contract(R, A), self.H[rl - 2]), A.prime().dag())
            self.LR[rl - 2] = R
            rl -= 1
        self.rpos = k

    def position(self, psi, pos: int):
        self.makeL(psi, pos - 1)
        self.makeR(psi, pos + self.nsite)

    # -- H_eff apply
    def contract_v(self, v: BSTensor) -> BSTensor:
        Hv = contract(v, self.lproj())
        for j in self.site_range():
            Hv = contract(Hv, self.H[j - 1])
        return contract(Hv, self.rproj())

    def product(self, v: BSTensor) -> BSTensor:
        Pv = self.contract_v(v)
        if Pv.rank != v.rank:
            raise RuntimeError("The order of the ProjMPO-ITensor product P*v is not equal to the order of v")
        return Pv.noprime()

    __call__ = product

    # -- noise term (density-matrix perturbation), returned per charge group of the kept side
    def noiseterm(self, phi: BSTensor, ortho: str) -> BSTensor:
        if self.nsite != 2:
            raise RuntimeError("noise term only defined for 2-site ProjMPO")
        sr = list(self.site_range())
        if ortho == "left":
            AL = contract(self.lproj(), self.H[sr[0] - 1])
            nt = contract(AL, phi)
        elif ortho == "right":
            AR = contract(self.H[sr[-1] - 1], self.rproj())
            nt = contract(phi, AR)
        else:
            raise ValueError(f"In noiseterm, got ortho = {ortho}, only supports `left` and `right`")
        return contract(nt, nt.noprime().dag())


class ProjMPOSum2:
    """src/mps/projmposum2.jl:15-145: a vector of ProjMPO over the same state; product and noiseterm are sums."""

    def __init__(self, Hs: Sequence[Sequence[BSTensor]]):
        self.PHs = [ProjMPO(H) for H in Hs]
        self.N = self.PHs[0].N

    @property
    def nsite(self):
        return self.PHs[0].nsite

    @property
    def lpos(self):
        return self.PHs[0].lpos

    @property
    def rpos(self):
        return self.PHs[0].rpos

    def set_nsite(self, n: int):
        for p in self.PHs:
            p.set_nsite(n)

    def position(self, psi, pos: int):
        for p in self.PHs:
            p.position(psi, pos)

    def product(self, v: BSTensor) -> BSTensor:
        Pv = None
        for p in self.PHs:
            t = p.product(v)
            Pv = t if Pv is None else Pv.add(t.permute(Pv.inds))
        return Pv

    __call__ = product

    def noiseterm(self, phi: BSTensor, ortho: str) -> BSTensor:
        nt = None
        for p in self.PHs:
            t = p.noiseterm(phi, ortho)
            nt = t if nt is None else nt.add(t.permute(nt.inds))
        return nt


def drho_matrices(drho: BSTensor, scale: float, order: Sequence[Index] | None = None):
    """Per-charge matrices of scale*drho, rows = primed index group (layout of `matricize`).  `order`: the
    (unprimed) indices of the kept side in the order the factorisation groups them -- the noise term of a
    CouplingModel comes out with its indices in another order than the two-site tensor's."""
    if order is None:
        primed = [ix for ix in drho.inds if ix.plev > 0]
    else:
        primed = [next(jx for jx in drho.inds if jx.id == ix.id and jx.plev > 0) for ix in order]
    # columns in the same order as the rows (the noise term's own index order is arbitrary)
    unprimed = [next(jx for jx in drho.inds if jx.id == ix.id and jx.plev == 0) for ix in primed]
    drho = drho.permute(primed + unprimed)
    _, groups, _, _, _ = matricize(drho, primed)
    return {q: scale * M for q, (M, _, _) in groups.items()}


class ProjMPS2:
    """Overlap environments of a fixed MPS `M` with the optimised state (src/mps/projmps2.jl:26-32,77-224):
    L_j = L_{j-1} * psi[j] * dag(prime(M[j], "Link")); product(v) = <m|v> |m> with
    dag(m) = L * dag(M_j') * dag(M_{j+1}') * R  (`proj_mps`, :182-196)."""

    def __init__(self, M: Sequence[BSTensor]):
        self.M = list(M)
        self.N = len(M)
        self.lpos, self.rpos, self.nsite = 0, self.N + 1, 2
        self.LR: List[BSTensor | None] = [None] * self.N
        self._Ledge = self._Redge = None

    def set_nsite(self, n):
        self.nsite = n

    def site_range(self):
        return range(self.lpos + 1, self.rpos)

    def _mdag(self, j):          # dag(prime(M[j], "Link")) : links primed, site index untouched
        Mj = self.M[j]
        return Mj.prime(1, [Mj.inds[0], Mj.inds[2]]).dag()

    def _edges(self, psi):
        def edge(a, m):
            return BSTensor([a.copy(dir=-a.dir), m.prime().copy(dir=m.dir)], {(0, 0): np.ones((1, 1))})
        if self._Ledge is None:
            self._Ledge = edge(psi[0].inds[0], self.M[0].inds[0])
            self._Redge = edge(psi[-1].inds[2], self.M[-1].inds[2])

    def lproj(self):
        return self._Ledge if self.lpos <= 0 else self.LR[self.lpos - 1]

    def rproj(self):
        return self._Redge if self.rpos >= self.N + 1 else self.LR[self.rpos - 1]

    def position(self, psi, pos):
        self._edges(psi)
        k = pos - 1
        if self.lpos >= k:
            self.lpos = k
        else:
            ll, L = max(self.lpos, 0), self.lproj()
            while ll < k:
                L = contract(contract(L, psi[ll]), self._mdag(ll))
                self.LR[ll] = L
                ll += 1
            self.lpos = k
        k = pos + self.nsite
        if self.rpos <= k:
            self.rpos = k
        else:
            rl, R = min(self.rpos, self.N + 1), self.rproj()
            while rl > k:
                R = contract(contract(R, psi[rl - 2]), self._mdag(rl - 2))
                self.LR[rl - 2] = R
                rl -= 1
            self.rpos = k

    def proj_mps(self) -> BSTensor:
        m = self.lproj()
        for j in self.site_range():
            m = contract(m, self._mdag(j - 1))
        return contract(m, self.rproj())

    def contract_v(self, v):
        Mv = contract(v, self.lproj())
        for j in self.site_range():
            Mv = contract(Mv, self._mdag(j - 1))
        return contract(Mv, self.rproj())

    def product(self, v):
        ov = self.contract_v(v).scalar()
        return self.proj_mps().dag().scale(ov)


class ProjMPO_MPS2:
    """PH + weight * sum_M |M><M| (src/mps/projmpo_mps2.jl:94-134); noiseterm forwards to PH.  With `PH` given
    explicitly the same wrapper is ProjMPOSum_MPS (src/mps/projmposum_mps.jl:94-100) and ProjCouplingModel_MPS
    (src/mps/projcouplingmodel_mps.jl:95-101), which differ only in the type of PH."""

    def __init__(self, H, Ms, weight: float, PH=None):
        if weight <= 0.0:
            raise ValueError(f"`weight` parameter should be > 0.0 (value passed was `weight={weight}`)")
        self.PH = ProjMPO(H) if PH is None else PH
        self.pm = [ProjMPS2(M) for M in Ms]
        self.weight = weight
        self.N = self.PH.N

    @property
    def nsite(self):
        return self.PH.nsite

    def set_nsite(self, n):
        self.PH.set_nsite(n)
        for p in self.pm:
            p.set_nsite(n)

    def position(self, psi, pos):
        self.PH.position(psi, pos)
        for p in self.pm:
            p.position(psi, pos)

    def product(self, v):
        Pv = self.PH.product(v)
        for p in self.pm:
            Pv = Pv.add(p.product(v), self.weight)
        return Pv

    __call__ = product

    def noiseterm(self, phi, ortho):
        return self.PH.noiseterm(phi, ortho)
