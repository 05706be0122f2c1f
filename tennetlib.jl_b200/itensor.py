"""Device-resident tensors with index identities: the ITensor operations the tree-tensor-network path, the Global
Subspace Expansion and the measurements of the reference are written in, served by the generic device algebra of
include/tnl_b200.h (`tnl_tensor_contract / _permute / _dag / _directsum / _factorize`).

The reference manipulates `ITensor`s [3P] through index identities:
  A * B                      contracts the indices the two tensors share     (src/ttn/linktensors.jl:91-94,197-199)
  prime / noprime / dag      relabel / reverse arrows (+ conjugate)           (src/ttn/linktensors.jl:88-90)
  commoninds / uniqueinds    index bookkeeping                                (src/ttn/ttn.jl:281-283)
  svd / qr / factorize       bipartition + truncation                         (src/ttn/ttn.jl:293-305)
  directsum                  enlarge one index                                (src/ttn/update_site_ttn.jl:92-104)
Here an identity is the pair (Index.id, Index.plev); it is mapped to the integer labels of the C ABI.  `dag` is lazy
(a flag handed to the contraction, which conjugates complex operands on the fly); it is materialised only when the
tensor is used as a vector.  Nothing in this module computes on the host: every numerical operation is a library call.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np

from ._lib import check
from .tensor import Context, DeviceTensor, HostTensor, Index

_labels: Dict[Tuple[int, int], int] = {}


def _key(ix) -> Tuple[int, int]:
    return (ix.id, ix.plev)


def _label(ix) -> int:
    k = _key(ix)
    v = _labels.get(k)
    if v is None:
        v = _labels[k] = len(_labels) + 1
    return v


def _as_index(ix) -> Index:
    """Accept foreign index objects (anything with id / plev / qns / dims / dir / tags)."""
    if isinstance(ix, Index):
        return ix
    return Index(ix.qns, ix.dims, dir=ix.dir, tags=ix.tags, plev=ix.plev, id=ix.id)


def same_index(a, b) -> bool:
    return a.id == b.id and a.plev == b.plev


def hastags(ix, tag: str) -> bool:
    return tag in [t.strip() for t in ix.tags.split(",")]


class ITensor:
    """A `tnl_tensor_t` plus the identities of its indices."""

    __slots__ = ("dt", "inds", "conj")

    def __init__(self, dt: DeviceTensor, inds: Sequence[Index], conj: bool = False):
        self.dt = dt
        self.inds = [_as_index(ix) for ix in inds]
        self.conj = conj

    # ------------------------------------------------------------------ host <-> device
    @staticmethod
    def from_host(ctx: Context, t, nrow: int = 1) -> "ITensor":
        """t: any block-sparse host tensor (`.inds` with id / plev / tags / qns / dims / dir, `.blocks`)."""
        inds = [_as_index(ix) for ix in t.inds]
        return ITensor(DeviceTensor.from_host(ctx, HostTensor(inds, t.blocks), nrow=min(nrow, len(inds))), inds)

    def to_host(self, drop_zero_blocks: bool = False) -> HostTensor:
        m = self.materialize()
        h = m.dt.to_host(drop_zero_blocks)
        h.inds = [ix.copy() for ix in m.inds]
        return h

    @staticmethod
    def zeros(ctx: Context, inds: Sequence[Index], nrow: int = 1) -> "ITensor":
        inds = [_as_index(ix) for ix in inds]
        return ITensor(DeviceTensor.zeros(ctx, inds, nrow), inds)

    @staticmethod
    def random(ctx: Context, inds: Sequence[Index], seed: int, nrow: int = 1) -> "ITensor":
        """`randomITensor(QN(), inds...)`: uniform entries in every flux-0 block (counter-based generator on the device)."""
        t = ITensor.zeros(ctx, inds, nrow)
        t.dt.fill_random(seed)
        return t

    # ------------------------------------------------------------------ bookkeeping
    @property
    def ctx(self) -> Context:
        return self.dt.ctx

    @property
    def rank(self) -> int:
        return len(self.inds)

    @property
    def nrow(self) -> int:
        out = C.c_int32()
        check(self.ctx.lib.tnl_tensor_nrow(self.dt.h, C.byref(out)), self.ctx.h)
        return out.value

    def is_complex(self) -> bool:
        return self.dt.is_complex()

    def _with(self, inds, conj=None) -> "ITensor":
        return ITensor(self.dt, inds, self.conj if conj is None else conj)

    def copy(self) -> "ITensor":
        return ITensor(self.dt.copy(), self.inds, self.conj)

    def find(self, ix) -> int:
        for k, jx in enumerate(self.inds):
            if same_index(ix, jx):
                return k
        raise KeyError(f"index {ix!r} not in the tensor")

    def hasind(self, ix) -> bool:
        return any(same_index(ix, jx) for jx in self.inds)

    def prime(self, n: int = 1, which: Iterable | None = None, tags: str | None = None) -> "ITensor":
        if which is not None:
            keys = {_key(ix) for ix in which}
            sel = lambda ix: _key(ix) in keys
        elif tags is not None:
            sel = lambda ix: hastags(ix, tags)
        else:
            sel = lambda ix: True
        return self._with([ix.prime(n) if sel(ix) else ix for ix in self.inds])

    def noprime(self) -> "ITensor":
        return self._with([ix.copy(plev=0) for ix in self.inds])

    def dag(self) -> "ITensor":
        return self._with([ix.dag() for ix in self.inds], not self.conj)

    def replaceinds(self, old: Sequence, new: Sequence) -> "ITensor":
        inds = list(self.inds)
        for o, n in zip(old, new):
            k = self.find(o)
            n = _as_index(n)
            if tuple(n.dims) != tuple(inds[k].dims) or tuple(n.qns) != tuple(inds[k].qns):
                raise ValueError("replaceinds: the new index spans a different space")
            inds[k] = n.copy(dir=inds[k].dir)
        return self._with(inds)

    def fresh(self) -> "ITensor":
        """A private copy with any pending `dag` carried out (safe to modify in place)."""
        m = self.materialize()
        return m.copy() if m is self else m

    def materialize(self) -> "ITensor":
        """Carry out a pending `dag` on the device (arrows reversed, ComplexF64 conjugated)."""
        if not self.conj:
            return self
        h = C.c_void_p()
        check(self.ctx.lib.tnl_tensor_dag(self.dt.h, C.byref(h)), self.ctx.h)
        return ITensor(DeviceTensor(self.ctx, h, self.inds), self.inds, False)

    def permute(self, order: Sequence, nrow: int = 1) -> "ITensor":
        """Same tensor with its indices in `order` and the first `nrow` of them as row group of the layout."""
        perm = [self.find(ix) for ix in order]
        if len(perm) != self.rank or len(set(perm)) != self.rank:
            raise ValueError("permute: not a permutation of the tensor's indices")
        if perm == list(range(self.rank)) and self.nrow == nrow:
            return self
        p = np.ascontiguousarray(perm, dtype=np.int32)
        h = C.c_void_p()
        check(self.ctx.lib.tnl_tensor_permute(self.dt.h, p.ctypes.data, int(nrow), C.byref(h)), self.ctx.h)
        inds = [self.inds[k] for k in perm]
        return ITensor(DeviceTensor(self.ctx, h, inds), inds, self.conj)

    # ------------------------------------------------------------------ algebra
    def __mul__(self, other):
        if isinstance(other, ITensor):
            return contract(self, other)
        return self.fresh().scale_(float(other))

    __rmul__ = __mul__

    def __truediv__(self, a):
        return self * (1.0 / float(a))

    def norm(self) -> float:
        return self.dt.norm()

    def scale_(self, a: float) -> "ITensor":
        self.dt.scale_(float(a))
        return self

    def like(self, other: "ITensor") -> "ITensor":
        """self brought into the index order, arrows and layout of `other` (for flat vector operations)."""
        m = self.materialize()
        o_nrow = other.nrow
        return m.permute(other.inds, o_nrow)

    def add_(self, other: "ITensor", alpha: float = 1.0) -> "ITensor":
        """self += alpha * other (ITensor `+`: same index set, any order)."""
        if self.conj:
            raise ValueError("add_: materialize() the target first")
        o = other.like(self)
        self.dt.axpy_(o.dt, float(alpha))
        return self

    def add(self, other: "ITensor", alpha: float = 1.0) -> "ITensor":
        return self.fresh().add_(other, alpha)

    def inner(self, other: "ITensor"):
        """<self|other> = scalar(dag(self) * other)."""
        a = self.materialize()
        return a.dt.dot(other.like(a).dt)

    def scale_index_(self, ix, values) -> "ITensor":
        self.dt.scale_index_(self.find(ix), values)
        return self


def commoninds(A: ITensor, B: ITensor) -> List[Index]:
    return [ix for ix in A.inds if B.hasind(ix)]


def commonind(A: ITensor, B: ITensor) -> Index:
    c = commoninds(A, B)
    if not c:
        raise ValueError("the tensors share no index")
    return c[0]


def uniqueinds(A: ITensor, B: ITensor) -> List[Index]:
    return [ix for ix in A.inds if not B.hasind(ix)]


def contract(A: ITensor, B: ITensor) -> ITensor:
    """ITensor `A * B`: one grouped FP64 DMMA GEMM over the charge sectors (operands permuted into
    [free | contracted] / [contracted | free] on the device when they are not in that form already)."""
    if not uniqueinds(A, B) and uniqueinds(B, A):
        A, B = B, A                      # the result's row group must not be empty
    la = np.ascontiguousarray([_label(ix) for ix in A.inds], dtype=np.int32)
    lb = np.ascontiguousarray([_label(ix) for ix in B.inds], dtype=np.int32)
    lout = np.zeros(len(la) + len(lb), dtype=np.int32)
    rank = C.c_int32()
    h = C.c_void_p()
    ctx = A.ctx
    check(ctx.lib.tnl_tensor_contract(A.dt.h, la.ctypes.data, int(A.conj), B.dt.h, lb.ctypes.data, int(B.conj), C.byref(h),
                                      lout.ctypes.data, C.byref(rank)), ctx.h)
    by_label = {}
    for ix in list(A.inds) + list(B.inds):
        by_label.setdefault(_label(ix), ix)
    inds = [by_label[int(l)] for l in lout[:rank.value]]
    return ITensor(DeviceTensor(ctx, h, inds), inds, False)


def directsum(A: ITensor, ia, B: ITensor, ib, tags: str | None = None):
    """ITensors `directsum(A => ia, B => ib)`: all other indices shared; returns (tensor, new index).  The summed index
    comes last, the shared ones keep A's order."""
    A, B = A.materialize(), B.materialize()
    oa = [ix for ix in A.inds if not same_index(ix, ia)]
    Ap = A.permute(oa + [A.inds[A.find(ia)]], max(1, len(oa)))
    Bp = B.permute([B.inds[B.find(ix)] for ix in oa] + [B.inds[B.find(ib)]], max(1, len(oa)))
    h = C.c_void_p()
    ctx = A.ctx
    k = len(oa)
    check(ctx.lib.tnl_tensor_directsum(Ap.dt.h, k, Bp.dt.h, k, C.byref(h)), ctx.h)
    xa, xb = Ap.inds[k], Bp.inds[k]
    new = Index(list(xa.qns) + list(xb.qns), list(xa.dims) + list(xb.dims), dir=xa.dir,
                tags=xa.tags if tags is None else tags)
    inds = oa + [new]
    return ITensor(DeviceTensor(ctx, h, inds), inds, False), new


_DECOMP = {None: 0, "svd": 1, "eigen": 2, "qr": 3}
_SVD_ALG = {"divide_and_conquer": 0, "polar": 1, "gram": 2, "qr_iteration": 3, "recursive": 3}


class Spectrum:
    def __init__(self, eigs, truncerr):
        self.eigs = eigs
        self.truncerr = truncerr


def factorize(T: ITensor, left: Sequence, *, ortho: str = "left", which_decomp: str | None = None, maxdim=None,
              mindim: int = 1, cutoff: float | None = None, tags: str = "Link", svd_alg: str = "divide_and_conquer"):
    """ITensors `factorize(T, left...; ortho, which_decomp, maxdim, mindim, cutoff, tags)` -> (L, R, spec, link):
    L carries `left` + the new link, R the link + the rest; with ortho = "left" L is the isometry.  `which_decomp = "qr"`
    is the untruncated gauge move of `qr(T, left)` (src/ttn/ttn.jl:302-305)."""
    T = T.materialize()
    left = [T.inds[T.find(ix)] for ix in left]
    right = [ix for ix in T.inds if not any(same_index(ix, l) for l in left)]
    if not left or not right:
        raise ValueError("factorize needs a proper bipartition")
    Tp = T.permute(left + right, len(left))
    ctx = T.ctx
    cap = min(int(min(np.prod([float(ix.dim) for ix in left]), np.prod([float(ix.dim) for ix in right]))) + 1, 1 << 22)
    eigs = np.zeros(cap)
    terr = C.c_double()
    neigs = C.c_int64()
    hl, hr = C.c_void_p(), C.c_void_p()
    which = _DECOMP[which_decomp] | (_SVD_ALG[svd_alg] << 4)
    check(ctx.lib.tnl_tensor_factorize(Tp.dt.h, len(left), 1 if ortho == "left" else 0, int(maxdim) if maxdim else 0,
                                       int(mindim), 0.0 if cutoff is None else float(cutoff), which, C.byref(hl), C.byref(hr),
                                       C.byref(terr), eigs.ctypes.data, cap, C.byref(neigs)), ctx.h)
    dl = DeviceTensor(ctx, hl)
    dr = DeviceTensor(ctx, hr)
    lk = dl.inds[-1]                       # new link as the device built it (sectors ascending in charge)
    link = Index(lk.qns, lk.dims, dir=lk.dir, tags=tags)
    linds = left + [link]
    rinds = [link.copy(dir=dr.inds[0].dir)] + right
    n = min(cap, neigs.value)
    return (ITensor(dl, linds), ITensor(dr, rinds), Spectrum(eigs[:n].copy(), terr.value), link)
