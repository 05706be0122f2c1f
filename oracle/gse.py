"""Global Subspace Expansion on the CPU (oracle; test infrastructure only) -- SURVEY.md section 8(f) rank 1.

Restates /root/reference/src/mps/sweep.jl:399-555:
  * `krylov_extend!(psi, H)`      : phis[k] = normalize(apply(H, phis[k-1]; maxdim, cutoff)), then `_krylov_addbasis!`
  * `_krylov_addbasis!`           : from the right end, enlarge the right-orthonormal basis B of psi[j] by the dominant
                                    eigenvectors of P rho P, rho = sum_k phi_k[j]^+ phi_k[j] / tr, P = 1 - B^+ B
                                    (the reference credits ITensorTDVP.jl PR 24)
  * `krylov_extend!(sysenv)`      : the same on sysenv.psi followed by a reset of the environments (:446-452)
and ITensorMPS `apply(H::MPO, psi::MPS; maxdim, cutoff)` (third-party, not vendored; default algorithm
"densitymatrix": sequential optimal truncation of H|psi> from the right end with the truncation rule of `truncate!`).
Here the exact product is canonicalised from the left without truncation and then truncated from the right by SVD,
which is the same sequence of optimal truncations (reduced density matrices of the already truncated right part)
in another gauge.  Deviation without effect on any contraction: the enlarged link keeps one sector per charge
(ITensors' `directsum` lists the sectors of both summands separately).
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

from .blocksparse import (BSTensor, Index, _group_sectors, _unmatricize_rows, contract, factorize, matricize, svd_bs,
                          truncate_spectrum)
from .dmrg import FLOAT64_THRESHOLD, MPS, StateEnvs, orthogonalize
from .projmpo import ProjMPO


def _fuse_boundary(T: BSTensor, a: Index, b: Index, tags: str) -> BSTensor:
    """Replace the two dim-1 indices a, b of T by one dim-1 index carrying their combined charge."""
    assert a.dim == 1 and b.dim == 1
    q = tuple(a.dir * x + b.dir * y for x, y in zip(a.qns[0], b.qns[0]))
    # new index with direction of `a`: a.dir * qnew = q
    new = Index([tuple(a.dir * x for x in q)], [1], dir=a.dir, tags=tags)
    F = BSTensor([a.dag(), b.dag(), new], {(0, 0, 0): np.ones((1, 1, 1))})
    return contract(T, F), new


def apply_mpo(H: Sequence[BSTensor], psi: MPS, maxdim=None, cutoff: float = FLOAT64_THRESHOLD) -> MPS:
    """|phi> = H|psi> as an MPS with bond dimension <= maxdim (see the module docstring).  The result shares the site
    indices and the boundary links' charges with psi; its orthogonality centre is site 1."""
    N = len(psi)
    phi: List[BSTensor] = []
    for j in range(N):
        A, W = psi.t[j], H[j]
        T = contract(A, W)                         # (l, r, wl, s', wr) in some order
        T = T.noprime()                            # s' -> s
        phi.append(T)
    # left to right: orthonormalise without truncation; the pair (l_j, w_j) stays a pair of indices
    carry = None
    for j in range(N - 1):
        T = phi[j] if carry is None else contract(carry, phi[j])
        right = [psi.t[j].inds[2], H[j].inds[3]]
        left = [ix for ix in T.inds if ix not in right]
        U, R, _, _ = factorize(T, left, ortho="left", which_decomp="svd", maxdim=None, mindim=1, cutoff=None, tags=f"Link,l={j + 1}")
        phi[j] = U
        carry = R
    phi[N - 1] = contract(carry, phi[N - 1]) if carry is not None else phi[N - 1]
    # boundary links: (l_0, w_0) and (l_N, w_N) are pairs of dim-1 indices -> one dim-1 index each
    phi[0], _ = _fuse_boundary(phi[0], psi.t[0].inds[0], H[0].inds[0], "Link,l=0")
    phi[N - 1], _ = _fuse_boundary(phi[N - 1], psi.t[N - 1].inds[2], H[N - 1].inds[3], f"Link,l={N}")
    # right to left: truncate
    for j in range(N - 1, 0, -1):
        T = phi[j]
        lft = [ix for ix in T.inds if ix in phi[j - 1].inds]
        L, R, _, _ = factorize(T, lft, ortho="right", which_decomp="svd", maxdim=maxdim, mindim=1, cutoff=cutoff,
                               tags=f"Link,l={j}")
        phi[j] = R
        phi[j - 1] = contract(phi[j - 1], L)
    # canonical index order (l, s, r)
    out = []
    for j in range(N):
        s = psi.t[j].inds[1]
        T = phi[j]
        sidx = next(ix for ix in T.inds if ix.id == s.id)
        if j == 0:
            r = next(ix for ix in T.inds if ix in phi[1].inds) if N > 1 else None
            l = next(ix for ix in T.inds if ix is not sidx and ix != sidx and (r is None or ix != r))
        else:
            l = next(ix for ix in T.inds if ix in out[j - 1].inds)
            r = next(ix for ix in T.inds if ix != sidx and ix != l)
        out.append(T.permute([l, sidx, r]))
    return MPS(out, 0, 2)


def mps_normalize(psi: MPS) -> MPS:
    c = psi.orthocenter()
    psi[c] = psi[c].scale(1.0 / psi[c].norm())
    return psi


def _row_matrices(T: BSTensor, rowinds: Sequence[Index], colinds: Sequence[Index]):
    """{column charge: matrix [rows x cols]} of T with the given bipartition, columns laid out by `_group_sectors`
    of colinds (so that matrices of different tensors over the same column indices are compatible)."""
    _, groups, _, _, _ = matricize(T.permute(list(rowinds) + list(colinds)), list(rowinds))
    return {tuple(-x for x in q): M for q, (M, _, _) in groups.items()}


def krylov_addbasis(psi: MPS, phis: List[MPS], extension_cutoff: float) -> MPS:
    """`_krylov_addbasis!` (src/mps/sweep.jl:497-555)."""
    N = len(psi)
    orthogonalize(psi, N)
    for phi in phis:
        orthogonalize(phi, N)
    # the Krylov states must share psi's right boundary link (ITensors MPS have none)
    for phi in phis:
        rb, pb = phi[N].inds[2], psi[N].inds[2]
        assert rb.same_space(pb), "Krylov vector in another charge sector"
        phi[N] = phi[N].replaceinds([rb], [pb])
    for j in range(N, 1, -1):
        A = psi[j]
        l, s, r = A.inds
        U, S, B, _, b = svd_bs(A, [l], truncate=False, tags=f"Link,l={j}")        # B(b+, s, r): right-orthonormal
        cols = [B.inds[1], B.inds[2]]
        gcols, tcols = _group_sectors(cols)
        Bm = _row_matrices(B, [B.inds[0]], cols)
        rho: Dict[tuple, np.ndarray] = {}
        for phi in phis:
            T = phi[j]
            pc = [next(ix for ix in T.inds if ix == c) for c in cols]
            pl = [ix for ix in T.inds if ix not in pc]
            for qc, M in _row_matrices(T, pl, pc).items():
                rho[qc] = rho.get(qc, 0) + M.conj().T @ M
        tr = sum(float(np.real(np.trace(M))) for M in rho.values())
        PrhoP: Dict[tuple, np.ndarray] = {}
        nrm2 = 0.0
        for qc, M in rho.items():
            n = M.shape[0]
            P = np.eye(n, dtype=M.dtype)
            if qc in Bm:
                P = P - Bm[qc].conj().T @ Bm[qc]
            X = P @ (M / tr) @ P
            PrhoP[qc] = X
            nrm2 += float(np.linalg.norm(X) ** 2)
        newrows: Dict[tuple, np.ndarray] = {}
        if np.sqrt(nrm2) > 1e-14:
            evs, vecs = {}, {}
            for qc, X in PrhoP.items():
                w, v = np.linalg.eigh((X + X.conj().T) / 2)
                o = np.argsort(-np.abs(w), kind="stable")
                evs[qc], vecs[qc] = w[o], v[:, o]
            pool = np.sort(np.concatenate([np.abs(w) for w in evs.values()]))[::-1]
            _, _, docut = truncate_spectrum(pool, None, 1, extension_cutoff)
            for qc in evs:
                nk = int(np.sum(np.abs(evs[qc]) > docut))
                if nk > 0:
                    newrows[qc] = vecs[qc][:, :nk].conj().T              # rows = new basis vectors
        # enlarged basis Bx: per column charge the rows of B followed by the new rows
        qcs = sorted(set(Bm) | set(newrows))
        mats, dims = {}, []
        for qc in qcs:
            parts = [m for m in (Bm.get(qc), newrows.get(qc)) if m is not None]
            mats[qc] = np.vstack(parts)
            dims.append(mats[qc].shape[0])
        # index bx on Bx has direction +1 and charge = row charge = -(column charge) of the (s, r) group
        rowq = [tuple(-x for x in qc) for qc in qcs]
        order = sorted(range(len(qcs)), key=lambda i: rowq[i])
        bx = Index([rowq[i] for i in order], [dims[i] for i in order], dir=-1, tags=f"Link,l={j - 1}")
        # _unmatricize_rows builds (cols..., u) from [cols x n] matrices keyed by u's sector charge
        Bx = _unmatricize_rows(cols, bx.dag(), {rowq[i]: mats[qcs[i]].T for i in order},
                               {rowq[i]: gcols[qcs[i]] for i in order}, A.dtype)
        Bx = Bx.permute([Bx.inds[-1]] + Bx.inds[:-1])                    # (bx+, s, r)
        for st in [psi] + list(phis):
            Tj = st[j]
            pc = {c.id: next(ix for ix in Tj.inds if ix == c) for c in cols}
            Tj = Tj.permute([ix for ix in Tj.inds if ix.id not in pc] + [pc[c.id] for c in cols])
            st[j - 1] = contract(st[j - 1], contract(Tj, Bx.dag()))
            st[j] = Bx
            st[j - 1] = st[j - 1].permute(list(st[j - 1].inds[:2]) + [st[j - 1].inds[2]])
    psi.llim, psi.rlim = 0, 2
    for phi in phis:
        phi.llim, phi.rlim = 0, 2
    return psi


def krylov_extend_mps(psi: MPS, H: Sequence[BSTensor], **kw) -> MPS:
    """`krylov_extend!(psi::MPS, H::MPO; kwargs...)` (src/mps/sweep.jl:399-417)."""
    kdim = kw.get("extension_krylovdim", 3)
    acut = kw.get("extension_applyH_cutoff", FLOAT64_THRESHOLD)
    amax = kw.get("extension_applyH_maxdim", max(A.inds[2].dim for A in psi.t[:-1]) + 2)
    ecut = kw.get("extension_cutoff", 1e-7)
    phis: List[MPS] = []
    for k in range(kdim):
        prev = psi if k == 0 else phis[k - 1]
        phi = apply_mpo(H, prev, maxdim=amax, cutoff=acut)
        phis.append(mps_normalize(phi))
    return krylov_addbasis(psi, phis, ecut)


def krylov_extend(sysenv: StateEnvs, **kw) -> None:
    """`krylov_extend!(sysenv::StateEnvs{ProjMPO}; kwargs...)` (src/mps/sweep.jl:432-467)."""
    if type(sysenv.PH) is not ProjMPO:
        raise RuntimeError("krylov_extend! needs a StateEnvs created from a single MPO")
    krylov_extend_mps(sysenv.psi, sysenv.PH.H, **kw)
    sysenv.PH.lpos, sysenv.PH.rpos, sysenv.PH.nsite = 0, sysenv.PH.N + 1, 2
    sysenv.PH.LR = [None] * sysenv.PH.N
