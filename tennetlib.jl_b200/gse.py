"""Global Subspace Expansion on the GPU -- `krylov_extend!` / `_krylov_addbasis!` of
/root/reference/src/mps/sweep.jl:399-555, the first thing `dynamic_fullsweep!` does for
`StateEnvs{ProjMPO}` (:266-282) and what makes one-site TDVP usable.

All tensors stay in HBM; every step is a call of the generic device algebra (itensor.py):
  * `apply(H, psi; maxdim, cutoff)` [3P ITensorMPS, default "densitymatrix" algorithm]: the exact product H|psi> is
    brought to left-canonical form without truncation (gauge moves) and then truncated from the right end with the
    `truncate!` rule -- the same sequence of optimal truncations in another gauge;
  * `_krylov_addbasis!`: from the right end, the right-orthonormal basis B of psi[j] is enlarged by the dominant
    eigenvectors of P rho P, rho = sum_k phi_k[j]^+ phi_k[j] / tr, P = 1 - B^+ B.  With X_k = phi_k[j] (1 - B^+ B) stacked
    along their left links, P rho P = X^+ X / tr: its eigenvectors above the cutoff are the right singular vectors of X,
    so the step is one `factorize` (Gram eigenproblem on the DMMA GEMM) plus a `directsum`.
"""
from __future__ import annotations

import time
from typing import List, Sequence

import numpy as np

from .itensor import ITensor, commoninds, contract, directsum, factorize, same_index
from .tensor import HostTensor, Index

FLOAT64_THRESHOLD = 1e-14


class DeviceMPS:
    """An MPS as a list of device ITensors (l, s, r) with shared link identities."""

    def __init__(self, tensors: List[ITensor], center: int | None = None):
        self.t = list(tensors)
        self.center = center          # 1-based orthogonality centre (None: unknown)

    def __len__(self):
        return len(self.t)

    def maxlinkdim(self) -> int:
        return max([A.inds[2].dim for A in self.t[:-1]] or [1])

    def orthogonalize(self, j: int):
        """orthogonalize!(psi, j) by untruncated gauge moves."""
        N = len(self.t)
        lo = 1 if self.center is None else self.center
        hi = N if self.center is None else self.center
        for k in range(lo, j):                                       # left-orthonormalise k, push into k + 1
            A = self.t[k - 1]
            U, R, _, _ = factorize(A, A.inds[:2], ortho="left", which_decomp="qr", tags=A.inds[2].tags)
            self.t[k - 1] = U
            self.t[k] = _lsr(contract(R, self.t[k]), self.t[k])
        for k in range(hi, j, -1):                                   # right-orthonormalise k, push into k - 1
            A = self.t[k - 1]
            L, Q, _, _ = factorize(A, A.inds[:1], ortho="right", which_decomp="qr", tags=A.inds[0].tags)
            self.t[k - 1] = Q
            self.t[k - 2] = contract(self.t[k - 2], L)
        self.center = j

    def normalize(self):
        c = self.t[self.center - 1]
        self.t[self.center - 1] = c.fresh().scale_(1.0 / c.norm())


def _lsr(T: ITensor, like: ITensor) -> ITensor:
    """T with its indices ordered (left link, site, right link) where `like` supplies site and right link."""
    s, r = like.inds[1], like.inds[2]
    l = next(ix for ix in T.inds if not same_index(ix, s) and not same_index(ix, r))
    return T.permute([l, T.inds[T.find(s)], T.inds[T.find(r)]], 1)


def mps_from_env(sysenv, H: Sequence[ITensor]) -> DeviceMPS:
    """The state of a StateEnvs as ITensors sharing the site identities of the MPO tensors W_j(wl, s', s, wr)."""
    N = sysenv.N
    out, prev = [], None
    for j in range(1, N + 1):
        dt = sysenv.site_tensor(j)
        l, s, r = dt.inds
        left = Index(l.qns, l.dims, dir=l.dir, tags=f"Link,l={j - 1}") if prev is None else prev.copy(dir=l.dir)
        site = H[j - 1].inds[2].copy(dir=s.dir, plev=0)
        right = Index(r.qns, r.dims, dir=r.dir, tags=f"Link,l={j}")
        out.append(ITensor(dt, [left, site, right]))
        prev = right
    c = sysenv.orthocenter() if sysenv.isortho() else None
    return DeviceMPS(out, c)


def _fuse_boundary(ctx, T: ITensor, a, b, tags: str) -> ITensor:
    """Replace the two dim-1 indices a, b of T by one dim-1 index carrying their combined charge."""
    a, b = T.inds[T.find(a)], T.inds[T.find(b)]
    if a.dim != 1 or b.dim != 1:
        raise RuntimeError("apply(H, psi): boundary links must have dimension 1")
    q = tuple(a.dir * x + b.dir * y for x, y in zip(a.qns[0], b.qns[0]))
    new = Index([tuple(a.dir * x for x in q)], [1], dir=a.dir, tags=tags)
    F = ITensor.from_host(ctx, HostTensor([a.dag(), b.dag(), new], {(0, 0, 0): np.ones((1, 1, 1))}), nrow=2)
    return contract(T, F)


def apply_mpo(H: Sequence[ITensor], psi: DeviceMPS, maxdim=None, cutoff: float = FLOAT64_THRESHOLD) -> DeviceMPS:
    """|phi> = H|psi> with bond dimension <= maxdim; site and boundary-link charges of psi; centre at site 1."""
    N = len(psi)
    ctx = psi.t[0].ctx
    phi = [contract(psi.t[j], H[j]).noprime() for j in range(N)]               # (l, r, wl, s, wr)
    carry = None
    for j in range(N - 1):
        T = phi[j] if carry is None else contract(carry, phi[j])
        right = [psi.t[j].inds[2], H[j].inds[3]]
        left = [ix for ix in T.inds if not any(same_index(ix, r) for r in right)]
        U, carry, _, _ = factorize(T, left, ortho="left", which_decomp="qr", tags=f"Link,l={j + 1}")
        phi[j] = U
    if carry is not None:
        phi[N - 1] = contract(carry, phi[N - 1])
    phi[0] = _fuse_boundary(ctx, phi[0], psi.t[0].inds[0], H[0].inds[0], "Link,l=0")
    phi[N - 1] = _fuse_boundary(ctx, phi[N - 1], psi.t[N - 1].inds[2], H[N - 1].inds[3], f"Link,l={N}")
    for j in range(N - 1, 0, -1):
        T = phi[j]
        lft = commoninds(T, phi[j - 1])
        L, R, _, _ = factorize(T, lft, ortho="right", which_decomp="svd", maxdim=maxdim, mindim=1, cutoff=cutoff,
                               tags=f"Link,l={j}")
        phi[j] = R
        phi[j - 1] = contract(phi[j - 1], L)
    out: List[ITensor] = []
    for j in range(N):
        T = phi[j]
        s = T.inds[T.find(psi.t[j].inds[1])]
        r = commoninds(T, phi[j + 1])[0] if j + 1 < N else None
        l = commoninds(T, out[j - 1])[0] if j > 0 else None
        rest = [ix for ix in T.inds if not same_index(ix, s) and (r is None or not same_index(ix, r))
                and (l is None or not same_index(ix, l))]
        if l is None:
            l = rest[0]
        if r is None:
            r = rest[0]
        out.append(T.permute([l, s, r], 1))
    return DeviceMPS(out, 1)


def krylov_addbasis(psi: DeviceMPS, phis: List[DeviceMPS], extension_cutoff: float) -> DeviceMPS:
    """`_krylov_addbasis!` (src/mps/sweep.jl:477-555)."""
    N = len(psi)
    psi.orthogonalize(N)
    for phi in phis:
        phi.orthogonalize(N)
        rb, pb = phi.t[N - 1].inds[2], psi.t[N - 1].inds[2]
        if tuple(rb.qns) != tuple(pb.qns) or tuple(rb.dims) != tuple(pb.dims):
            raise RuntimeError("`_krylov_addbasis!()`: Krylov vector in another charge sector")
        phi.t[N - 1] = phi.t[N - 1].replaceinds([rb], [pb])
    for j in range(N, 1, -1):
        A = psi.t[j - 1]
        _, B, _, b = factorize(A, A.inds[:1], ortho="right", which_decomp="qr", tags=f"Link,l={j - 1}")   # B(b, s, r)
        rinds = B.inds[1:]
        Xs, tr = [], 0.0
        for phi in phis:
            Tj = phi.t[j - 1]
            tr += Tj.norm() ** 2
            ov = contract(Tj, B.dag())                      # (l_k, b)
            Xs.append(Tj.fresh().add_(contract(ov, B), -1.0))
        X, Lx = Xs[0], Xs[0].inds[0]
        for Xk in Xs[1:]:
            X, Lx = directsum(X, Lx, Xk, Xk.inds[0], tags="Link,krylov")
        X = X.fresh().scale_(1.0 / np.sqrt(tr))
        G = contract(X.prime(1, rinds).dag(), X)             # P rho P as a tensor over (s', r'; s, r)
        Bx = B
        if G.norm() > 1e-14:
            _, Bphi, _, bphi = factorize(X, [Lx], ortho="right", which_decomp="svd", cutoff=extension_cutoff,
                                         tags=f"bphi_{j},Link")
            Bx, bx = directsum(B, B.inds[0], Bphi, Bphi.inds[0], tags=f"Link,l={j - 1}")
            Bx = Bx.permute([bx] + list(Bx.inds[:-1]), 1)
        for st in [psi] + list(phis):
            st.t[j - 2] = contract(st.t[j - 2], contract(st.t[j - 1], Bx.dag()))
            st.t[j - 1] = Bx
    psi.center = 1
    for phi in phis:
        phi.center = 1
    return psi


def krylov_extend_mps(psi: DeviceMPS, H: Sequence[ITensor], **kw) -> DeviceMPS:
    """`krylov_extend!(psi::MPS, H::MPO; kwargs...)` (src/mps/sweep.jl:399-417)."""
    kdim = kw.get("extension_krylovdim", 3)
    acut = kw.get("extension_applyH_cutoff", FLOAT64_THRESHOLD)
    amax = kw.get("extension_applyH_maxdim", psi.maxlinkdim() + 2)
    ecut = kw.get("extension_cutoff", 1e-7)
    phis: List[DeviceMPS] = []
    for k in range(kdim):
        prev = psi if k == 0 else phis[k - 1]
        phi = apply_mpo(H, prev, maxdim=amax, cutoff=acut)
        phi.normalize()
        phis.append(phi)
    return krylov_addbasis(psi, phis, ecut)


def krylov_extend(sysenv, **kw) -> None:
    """`krylov_extend!(sysenv::StateEnvs{ProjMPO}; kwargs...)` (src/mps/sweep.jl:432-467): the expansion on sysenv.psi,
    then the environments are reset (lpos = 0, rpos = N + 1, nsite = 2)."""
    if sysenv.nterms != 1 or sysenv.is_coupling_model or sysenv.has_penalty:
        raise RuntimeError("`krylov_extend!()`: the `StateEnvs` must be created by a single MPO")
    t0 = time.time()
    ctx = sysenv.ctx
    H = [ITensor.from_host(ctx, W, nrow=2) for W in sysenv.H_host]
    psi = mps_from_env(sysenv, H)
    krylov_extend_mps(psi, H, **kw)
    for j, A in enumerate(psi.t):
        T = A.materialize().permute(A.inds, 2)
        sysenv.set_site_tensor(j + 1, T.dt if T.dt is not A.dt else T.dt.copy())
    sysenv.llim, sysenv.rlim = 0, 2
    sysenv.set_nsite(2)
    if kw.get("outputlevel", 1) > 0:
        print("-----------------------------------------------------------------------------------")
        print(f"Global Subspace Expansion: KrylovDim={kw.get('extension_krylovdim', 3)}, "
              f"applyH Cutoff={kw.get('extension_applyH_cutoff', FLOAT64_THRESHOLD):.2g}")
        print(f"Global Subspace Expansion: Cutoff={kw.get('extension_cutoff', 1e-7):.2g}, "
              f"MaxLinkDim={max(sysenv.linkdims())}, Time={time.time() - t0:.3f}")
        print("-----------------------------------------------------------------------------------", flush=True)
