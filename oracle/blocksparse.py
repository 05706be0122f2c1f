"""Block-sparse U(1)^n tensors on the CPU (oracle; test infrastructure only).

Restates the storage and contraction semantics of NDTensors' `BlockSparse` tensors
that every ITensor in the reference hot path uses (third-party, not vendored under
/root/reference; call sites e.g. src/mps/projcouplingmodel.jl:145-147,342-343,
src/mps/projmps2.jl:86,118,170, src/mps/update_site.jl:46):

* an index is a list of (QN, dim) sectors plus an arrow (`dir`), an id and a prime level;
* a tensor stores one dense block per tuple of sector numbers; a block is present only if
  sum_i dir_i * qn_i(block_i) equals the tensor's flux;
* contraction pairs blocks whose contracted sector numbers agree; output blocks are created in
  first-appearance order (blocks of the first tensor in the outer loop, of the second in the
  inner loop), one GEMM per pair, accumulated when an output block is hit again.

`FLOPS` counts 2*m*n*k per block pair (8*m*n*k for complex): this is the ALGORITHMIC flop
count used by bench.py's roofline.
"""
from __future__ import annotations

import itertools
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np

_id_counter = itertools.count(1)

# running algorithmic flop counter (reference contraction order), see module docstring
FLOPS = [0]


def reset_flops() -> None:
    FLOPS[0] = 0


def get_flops() -> int:
    return FLOPS[0]


QN = Tuple[int, ...]


class Index:
    """QN index: sectors [(qn, dim)], arrow dir (+1 = Out, -1 = In), id, prime level, tags."""

    __slots__ = ("id", "qns", "dims", "dir", "tags", "plev")

    def __init__(self, qns: Sequence[QN], dims: Sequence[int], dir: int = +1, tags: str = "",
                 plev: int = 0, id: int | None = None):
        assert len(qns) == len(dims)
        self.qns = tuple(tuple(int(x) for x in q) for q in qns)
        self.dims = tuple(int(d) for d in dims)
        self.dir = int(dir)
        self.tags = tags
        self.plev = int(plev)
        self.id = next(_id_counter) if id is None else id

    # -- identity ignores the arrow, as ITensors' Index `==` does
    def __eq__(self, other):
        return isinstance(other, Index) and self.id == other.id and self.plev == other.plev

    def __hash__(self):
        return hash((self.id, self.plev))

    def __repr__(self):
        s = ",".join(f"{q}:{d}" for q, d in zip(self.qns, self.dims))
        return f"Index(id={self.id}{chr(39) * self.plev}|{self.tags}|dir={self.dir:+d}|{s})"

    @property
    def nsect(self) -> int:
        return len(self.dims)

    @property
    def dim(self) -> int:
        return int(sum(self.dims))

    def copy(self, **kw) -> "Index":
        d = dict(qns=self.qns, dims=self.dims, dir=self.dir, tags=self.tags, plev=self.plev, id=self.id)
        d.update(kw)
        return Index(**d)

    def dag(self) -> "Index":
        return self.copy(dir=-self.dir)

    def prime(self, n: int = 1) -> "Index":
        return self.copy(plev=self.plev + n)

    def noprime(self) -> "Index":
        return self.copy(plev=0)

    def sim(self) -> "Index":
        return self.copy(id=next(_id_counter))

    def offsets(self) -> np.ndarray:
        return np.concatenate([[0], np.cumsum(self.dims)]).astype(np.int64)

    def same_space(self, other: "Index") -> bool:
        return self.qns == other.qns and self.dims == other.dims


def qn_add(a: QN, b: QN, sa: int = 1, sb: int = 1) -> QN:
    return tuple(sa * x + sb * y for x, y in zip(a, b))


def qn_zero(n: int) -> QN:
    return (0,) * n


class BSTensor:
    """Block-sparse tensor: `inds` + insertion-ordered dict {sector coords -> dense block}."""

    def __init__(self, inds: Sequence[Index], blocks: Dict[Tuple[int, ...], np.ndarray] | None = None,
                 dtype=np.float64):
        self.inds: List[Index] = list(inds)
        self.blocks: Dict[Tuple[int, ...], np.ndarray] = {} if blocks is None else blocks
        self.dtype = np.dtype(dtype)

    # ---------------------------------------------------------------- basic queries
    @property
    def rank(self) -> int:
        return len(self.inds)

    def block_shape(self, coords) -> Tuple[int, ...]:
        return tuple(ix.dims[c] for ix, c in zip(self.inds, coords))

    def block_qn(self, coords) -> QN:
        nq = len(self.inds[0].qns[0]) if self.inds else 0
        tot = (0,) * nq
        for ix, c in zip(self.inds, coords):
            tot = qn_add(tot, ix.qns[c], 1, ix.dir)
        return tot

    def flux(self) -> QN | None:
        for c in self.blocks:
            return self.block_qn(c)
        return None

    def nelem(self) -> int:
        return int(sum(b.size for b in self.blocks.values()))

    def copy(self) -> "BSTensor":
        return BSTensor(self.inds, {c: b.copy() for c, b in self.blocks.items()}, self.dtype)

    def allowed_blocks(self, flux: QN | None = None) -> List[Tuple[int, ...]]:
        """All sector tuples with the given flux, first index fastest (column-major order)."""
        nq = len(self.inds[0].qns[0])
        flux = qn_zero(nq) if flux is None else tuple(flux)
        out = []
        for rc in itertools.product(*[range(ix.nsect) for ix in reversed(self.inds)]):
            c = rc[::-1]
            if self.block_qn(c) == flux:
                out.append(c)
        return out

    # ---------------------------------------------------------------- index manipulation
    def _with_inds(self, inds) -> "BSTensor":
        return BSTensor(inds, self.blocks, self.dtype)

    def prime(self, n: int = 1, which: Iterable[Index] | None = None) -> "BSTensor":
        sel = None if which is None else set(which)
        return self._with_inds([ix.prime(n) if (sel is None or ix in sel) else ix for ix in self.inds])

    def noprime(self) -> "BSTensor":
        return self._with_inds([ix.noprime() for ix in self.inds])

    def replaceinds(self, old: Sequence[Index], new: Sequence[Index]) -> "BSTensor":
        m = {o: n for o, n in zip(old, new)}
        out = []
        for ix in self.inds:
            if ix in m:
                assert m[ix].same_space(ix)
                out.append(m[ix].copy(dir=ix.dir))
            else:
                out.append(ix)
        return self._with_inds(out)

    def dag(self) -> "BSTensor":
        blocks = {c: np.conj(b) for c, b in self.blocks.items()} if self.dtype.kind == "c" else self.blocks
        return BSTensor([ix.dag() for ix in self.inds], blocks, self.dtype)

    def permute(self, new_inds: Sequence[Index]) -> "BSTensor":
        perm = [self.inds.index(ix) for ix in new_inds]
        assert sorted(perm) == list(range(self.rank))
        blocks = {tuple(c[p] for p in perm): np.transpose(b, perm) for c, b in self.blocks.items()}
        return BSTensor([self.inds[p] for p in perm], blocks, self.dtype)

    # ---------------------------------------------------------------- vector-space ops
    def scale(self, a) -> "BSTensor":
        return BSTensor(self.inds, {c: b * a for c, b in self.blocks.items()},
                        np.result_type(self.dtype, np.asarray(a).dtype))

    def __mul__(self, other):
        if isinstance(other, BSTensor):
            return contract(self, other)
        return self.scale(other)

    __rmul__ = __mul__

    def __truediv__(self, a):
        return self.scale(1.0 / a)

    def add(self, other: "BSTensor", alpha=1.0) -> "BSTensor":
        """self + alpha*other (block union, as ITensor `+`; indices matched by identity)."""
        if other.inds != self.inds:
            other = other.permute(self.inds)
        dt = np.result_type(self.dtype, other.dtype, np.asarray(alpha).dtype)
        blocks = {c: b.astype(dt, copy=True) for c, b in self.blocks.items()}
        for c, b in other.blocks.items():
            if c in blocks:
                blocks[c] += alpha * b
            else:
                blocks[c] = (alpha * b).astype(dt)
        return BSTensor(self.inds, blocks, dt)

    def __add__(self, other):
        return self.add(other)

    def __sub__(self, other):
        return self.add(other, -1.0)

    def norm(self) -> float:
        return float(np.sqrt(sum(float(np.vdot(b, b).real) for b in self.blocks.values())))

    def scalar(self):
        assert self.rank == 0
        return self.blocks[()].item() if () in self.blocks else 0.0

    # ---------------------------------------------------------------- dense conversion
    def to_dense(self) -> np.ndarray:
        out = np.zeros([ix.dim for ix in self.inds], dtype=self.dtype)
        offs = [ix.offsets() for ix in self.inds]
        for c, b in self.blocks.items():
            sl = tuple(slice(o[k], o[k + 1]) for o, k in zip(offs, c))
            out[sl] = b
        return out

    @staticmethod
    def from_dense(inds: Sequence[Index], arr: np.ndarray, flux: QN | None = None, tol: float = 0.0,
                   keep_zero_blocks: bool = False) -> "BSTensor":
        t = BSTensor(inds, dtype=arr.dtype)
        offs = [ix.offsets() for ix in inds]
        for c in t.allowed_blocks(flux):
            sl = tuple(slice(o[k], o[k + 1]) for o, k in zip(offs, c))
            blk = np.array(arr[sl])
            if keep_zero_blocks or np.abs(blk).max(initial=0.0) > tol:
                t.blocks[c] = blk
        return t

    @staticmethod
    def random(inds: Sequence[Index], rng: np.random.Generator, flux: QN | None = None,
               dtype=np.float64) -> "BSTensor":
        """All symmetry-allowed blocks i.i.d. N(0,1), block by block in column-major block order
        (SURVEY.md section 8d synthetic inputs)."""
        t = BSTensor(inds, dtype=dtype)
        for c in t.allowed_blocks(flux):
            shp = t.block_shape(c)
            blk = rng.standard_normal(shp)
            if np.dtype(dtype).kind == "c":
                blk = blk + 1j * rng.standard_normal(shp)
            t.blocks[c] = blk.astype(dtype)
        return t

    # ---------------------------------------------------------------- NDTensors-style flat export
    def export_flat(self, sort_blocks: bool = True):
        """(coords[nb, rank] int32, offsets[nb] int64, data[nelem]) with each block dense
        COLUMN-major and 0-based element offsets, the NDTensors `BlockSparse` layout
        (SURVEY.md section 8b).  Blocks are emitted in column-major coordinate order."""
        keys = list(self.blocks.keys())
        if sort_blocks:
            keys.sort(key=lambda c: c[::-1])
        coords = np.zeros((len(keys), self.rank), dtype=np.int32)
        offsets = np.zeros(len(keys), dtype=np.int64)
        chunks = []
        off = 0
        for n, c in enumerate(keys):
            coords[n] = c
            offsets[n] = off
            b = self.blocks[c]
            chunks.append(np.asarray(b).reshape(-1, order="F"))
            off += b.size
        data = np.concatenate(chunks) if chunks else np.zeros(0, dtype=self.dtype)
        return coords, offsets, np.ascontiguousarray(data, dtype=self.dtype)

    @staticmethod
    def import_flat(inds: Sequence[Index], coords, offsets, data) -> "BSTensor":
        t = BSTensor(inds, dtype=np.asarray(data).dtype)
        for c, off in zip(np.asarray(coords).reshape(len(offsets), len(inds)), offsets):
            c = tuple(int(x) for x in c)
            shp = t.block_shape(c)
            n = int(np.prod(shp)) if shp else 1
            t.blocks[c] = np.array(data[off:off + n]).reshape(shp, order="F")
        return t


# =====================================================================================
def commoninds(A: BSTensor, B: BSTensor) -> List[Index]:
    sb = set(B.inds)
    return [ix for ix in A.inds if ix in sb]


def uniqueinds(A: BSTensor, B: BSTensor) -> List[Index]:
    sb = set(B.inds)
    return [ix for ix in A.inds if ix not in sb]


def contract(A: BSTensor, B: BSTensor) -> BSTensor:
    """ITensor `A * B`: contract all common indices (NDTensors block-matching rule)."""
    com = commoninds(A, B)
    ia = [A.inds.index(ix) for ix in com]
    ib = [B.inds.index(ix) for ix in com]
    for p, q in zip(ia, ib):
        assert A.inds[p].dir == -B.inds[q].dir, f"arrow mismatch contracting {A.inds[p]} with {B.inds[q]}"
        assert A.inds[p].same_space(B.inds[q])
    ua = [p for p in range(A.rank) if p not in ia]
    ub = [q for q in range(B.rank) if q not in ib]
    dt = np.result_type(A.dtype, B.dtype)
    fl = 8 if dt.kind == "c" else 2
    groupB: Dict[Tuple[int, ...], list] = {}
    for cb, b in B.blocks.items():
        groupB.setdefault(tuple(cb[q] for q in ib), []).append((cb, b))
    out: Dict[Tuple[int, ...], np.ndarray] = {}
    for ca, a in A.blocks.items():
        key = tuple(ca[p] for p in ia)
        for cb, b in groupB.get(key, ()):
            co = tuple(ca[p] for p in ua) + tuple(cb[q] for q in ub)
            r = np.tensordot(a, b, axes=(ia, ib))
            k = int(np.prod([a.shape[p] for p in ia])) if ia else 1
            FLOPS[0] += fl * (a.size // k) * (b.size // k) * k
            if co in out:
                out[co] += r
            else:
                out[co] = np.asarray(r, dtype=dt)
    return BSTensor([A.inds[p] for p in ua] + [B.inds[q] for q in ub], out, dt)


def inner(A: BSTensor, B: BSTensor):
    """<A|B> = sum conj(A) .* B over common blocks (VectorInterface.inner on ITensors)."""
    if B.inds != A.inds:
        B = B.permute(A.inds)
    s = 0.0
    for c, a in A.blocks.items():
        b = B.blocks.get(c)
        if b is not None:
            s = s + np.vdot(a, b)
    return s if np.iscomplexobj(s) else float(s)


# =====================================================================================
# Truncation and factorisation (NDTensors `truncate!`, block-sparse `svd` / `eigen`;
# ITensors `factorize`, used through ITensorMPS `replacebond!` at src/mps/update_site.jl:64-76)
# =====================================================================================
def truncate_spectrum(P: np.ndarray, maxdim: int | None = None, mindim: int = 1, cutoff: float = 0.0,
                      use_absolute_cutoff: bool = False, use_relative_cutoff: bool = True):
    """NDTensors `truncate!` on a descending spectrum P.  Returns (kept P, truncerr, docut)."""
    P = np.array(P, dtype=np.float64)
    origm = len(P)
    docut = 0.0
    if origm == 0:
        return P, 0.0, 0.0
    if origm == 1:
        return P, 0.0, abs(P[0]) / 2
    maxdim = origm if maxdim is None else min(int(maxdim), origm)
    s = np.sign(P[0])
    if s < 0:
        P = P * s
    for n in range(origm - 1, -1, -1):          # zero out negative weight at the tail
        if P[n] >= 0:
            break
        P[n] = 0.0
    n = origm
    truncerr = 0.0
    while n > maxdim:
        truncerr += P[n - 1]
        n -= 1
    if use_absolute_cutoff:
        while n > mindim and P[n - 1] <= cutoff:
            truncerr += P[n - 1]
            n -= 1
    else:
        scale = 1.0
        if use_relative_cutoff:
            scale = float(P.sum())
            if scale == 0.0:
                scale = 1.0
        while n > mindim and (truncerr + P[n - 1] <= cutoff * scale):
            truncerr += P[n - 1]
            n -= 1
        truncerr /= scale
    if n < 1:
        n = 1
    if n < origm:
        docut = (P[n - 1] + P[n]) / 2
        if abs(P[n - 1] - P[n]) < 1e-3 * P[n - 1]:
            docut += 1e-3 * P[n - 1]
    if s < 0:
        P = P * s
    return P[:n], float(truncerr), float(docut)


class Spectrum:
    def __init__(self, eigs, truncerr):
        self.eigs = np.asarray(eigs, dtype=np.float64)
        self.truncerr = float(truncerr)


def _group_sectors(inds: Sequence[Index]):
    """Combiner semantics: product sectors enumerated first-index-fastest, stably grouped by
    total charge sum_i dir_i*qn_i, groups in ascending charge order.
    Returns {charge: [(multi_sector, dim, row_offset)]} and {charge: total_dim}."""
    nq = len(inds[0].qns[0])
    groups: Dict[QN, list] = {}
    totals: Dict[QN, int] = {}
    for rc in itertools.product(*[range(ix.nsect) for ix in reversed(inds)]):
        c = rc[::-1]
        q = (0,) * nq
        d = 1
        for ix, k in zip(inds, c):
            q = qn_add(q, ix.qns[k], 1, ix.dir)
            d *= ix.dims[k]
        off = totals.get(q, 0)
        groups.setdefault(q, []).append((c, d, off))
        totals[q] = off + d
    return groups, totals


def matricize(T: BSTensor, left: Sequence[Index]):
    """Block-diagonal matrix view of a flux-0 tensor: {row charge q: (M_q, row_layout, col_layout)}.
    Only charge groups holding at least one stored block are returned, ascending in q."""
    left = list(left)
    right = [ix for ix in T.inds if ix not in left]
    Tp = T.permute(left + right)
    fl = Tp.flux()
    nq = len(T.inds[0].qns[0])
    assert fl is None or fl == qn_zero(nq), "matricize expects a flux-0 tensor"
    nl = len(left)
    gl, tl = _group_sectors(Tp.inds[:nl])
    gr, tr = _group_sectors(Tp.inds[nl:])
    rowpos = {q: {c: (d, off) for c, d, off in lst} for q, lst in gl.items()}
    colpos = {q: {c: (d, off) for c, d, off in lst} for q, lst in gr.items()}
    mats: Dict[QN, np.ndarray] = {}
    for c, b in Tp.blocks.items():
        cl, cr = c[:nl], c[nl:]
        q = Tp._with_inds(Tp.inds[:nl]).block_qn(cl) if nl else qn_zero(nq)
        qr = tuple(-x for x in q)
        if q not in mats:
            mats[q] = np.zeros((tl[q], tr[qr]), dtype=T.dtype)
        dr, ro = rowpos[q][cl]
        dc, co = colpos[qr][cr]
        mats[q][ro:ro + dr, co:co + dc] = np.reshape(b, (dr, dc), order="F")
    groups = {q: (mats[q], gl[q], gr[tuple(-x for x in q)]) for q in sorted(mats)}
    return Tp, groups, left, right, (gl, tl, gr, tr)


def _unmatricize_rows(left: Sequence[Index], u: Index, mats: Dict[QN, np.ndarray], layouts, dtype) -> BSTensor:
    """Tensor (left..., u) from per-charge matrices [rows(group q) x kept(q)]; u sector k <-> k-th charge."""
    t = BSTensor(list(left) + [u], dtype=dtype)
    for k, q in enumerate(u.qns):
        M = mats[q]
        for c, d, off in layouts[q]:
            shp = tuple(ix.dims[s] for ix, s in zip(left, c)) + (M.shape[1],)
            t.blocks[tuple(c) + (k,)] = np.reshape(M[off:off + d, :], shp, order="F").copy()
    return t


def svd_bs(T: BSTensor, left: Sequence[Index], maxdim=None, mindim=1, cutoff=None, tags="Link",
           truncate=True):
    """Block-sparse truncated SVD  T = U * diag(S) * V  (NDTensors `svd(::BlockSparseMatrix)`).

    Every charge group of the matricised tensor is decomposed with LAPACK gesdd; all sigma^2 are
    pooled, sorted descending and passed to `truncate_spectrum`; group q keeps sigma^2 > docut and
    is dropped when it keeps nothing.  The new index `u` (dir=-1 on U, +1 on V) has one sector per
    surviving group, ascending in charge.  Returns U, S (dict q->sigma), V, Spectrum, u."""
    Tp, groups, left, right, _ = matricize(T, left)
    Us, Ss, Vs = {}, {}, {}
    for q, (M, _, _) in groups.items():
        try:
            U, s, Vt = np.linalg.svd(M, full_matrices=False)
        except np.linalg.LinAlgError:  # NDTensors falls back gesdd -> gesvd
            import scipy.linalg
            U, s, Vt = scipy.linalg.svd(M, full_matrices=False, lapack_driver="gesvd")
        Us[q], Ss[q], Vs[q] = U, s, Vt
    P = np.sort(np.concatenate([s ** 2 for s in Ss.values()]) if Ss else np.zeros(0))[::-1]
    if truncate:
        Pk, truncerr, docut = truncate_spectrum(P, maxdim, mindim, 0.0 if cutoff is None else cutoff)
    else:
        Pk, truncerr, docut = P, 0.0, -1.0
    kept_q, kept_d = [], []
    for q in groups:
        nk = int(np.sum(Ss[q] ** 2 > docut))
        if nk > 0:
            kept_q.append(q)
            kept_d.append(nk)
    u = Index(kept_q, kept_d, dir=-1, tags=tags)
    Umats = {q: Us[q][:, :n] for q, n in zip(kept_q, kept_d)}
    Vmats = {q: Vs[q][:n, :].T for q, n in zip(kept_q, kept_d)}
    U = _unmatricize_rows(left, u, Umats, {q: groups[q][1] for q in kept_q}, T.dtype)
    Vr = _unmatricize_rows(right, u.dag(), {q: Vmats[q] for q in kept_q},
                           {q: groups[q][2] for q in kept_q}, T.dtype)
    V = Vr.permute([Vr.inds[-1]] + Vr.inds[:-1])
    S = {q: Ss[q][:n] for q, n in zip(kept_q, kept_d)}
    return U, S, V, Spectrum(Pk, truncerr), u


def _scale_link(T: BSTensor, u: Index, S: Dict[QN, np.ndarray]) -> BSTensor:
    """Multiply the `u` leg of T by diag(S)."""
    p = T.inds.index(u)
    out = BSTensor(T.inds, dtype=T.dtype)
    for c, b in T.blocks.items():
        s = S[u.qns[c[p]]]
        shp = [1] * b.ndim
        shp[p] = len(s)
        out.blocks[c] = b * s.reshape(shp)
    return out


def eigen_bs(T: BSTensor, left: Sequence[Index], which_side: str, drho: Dict[QN, np.ndarray] | None,
             maxdim=None, mindim=1, cutoff=None, tags="Link"):
    """Density-matrix factorisation (ITensors `factorize_eigen`): rho = M M^+ (+ drho) on the `left`
    index group when which_side == "left", M^+ M on the complementary group otherwise; Hermitian
    eigendecomposition per charge group, eigenvalues sorted descending by |.|, pooled truncation.
    Returns (Vt tensor with inds (group..., u), Spectrum, u)."""
    Tp, groups, left, right, (gl, tl, gr, tr) = matricize(T, left)
    if drho is not None:                      # the perturbation may open charge groups T lacks
        for q in drho:
            if q not in groups:
                qr = tuple(-x for x in q)
                groups[q] = (np.zeros((tl.get(q, 0), tr.get(qr, 0)), dtype=T.dtype), gl.get(q, []), gr.get(qr, []))
        groups = {q: groups[q] for q in sorted(groups)}
    Ds, Vs = {}, {}
    for q, (M, _, _) in groups.items():
        rho = M @ M.conj().T if which_side == "left" else M.conj().T @ M
        if drho is not None and q in drho:
            rho = rho + drho[q]
        w, v = np.linalg.eigh(rho)
        p = np.argsort(-np.abs(w), kind="stable")
        Ds[q], Vs[q] = w[p], v[:, p]
    P = np.sort(np.concatenate([np.abs(d) for d in Ds.values()]) if Ds else np.zeros(0))[::-1]
    Pk, truncerr, docut = truncate_spectrum(P, maxdim, mindim, 0.0 if cutoff is None else cutoff)
    kept_q, kept_d = [], []
    for q in groups:
        nk = int(np.sum(np.abs(Ds[q]) > docut))
        if nk > 0:
            kept_q.append(q)
            kept_d.append(nk)
    if which_side == "left":
        u = Index(kept_q, kept_d, dir=-1, tags=tags)
        Vt = _unmatricize_rows(left, u, {q: Vs[q][:, :n] for q, n in zip(kept_q, kept_d)},
                               {q: groups[q][1] for q in kept_q}, T.dtype)
    else:
        # u sits on the right factor: charge label of the ROW group (as for svd), arrow -1 on the
        # left factor => the right factor R(u+, right...) carries dir +1.
        u = Index(kept_q, kept_d, dir=-1, tags=tags)
        Vr = _unmatricize_rows(right, u.dag(), {q: np.conj(Vs[q][:, :n]) for q, n in zip(kept_q, kept_d)},
                               {q: groups[q][2] for q in kept_q}, T.dtype)
        Vt = Vr.permute([Vr.inds[-1]] + Vr.inds[:-1])
    return Vt, Spectrum(Pk, truncerr), u


def factorize(T: BSTensor, left: Sequence[Index], ortho: str = "left", maxdim=None, mindim=1, cutoff=None,
              eigen_perturbation: Dict[QN, np.ndarray] | None = None, which_decomp: str | None = None,
              tags="Link"):
    """ITensors `factorize` as called by `replacebond!` (which_decomp=nothing):
    eigen path when an `eigen_perturbation` is given or cutoff > 1e-12, SVD path otherwise.
    ortho == "left":  L = U (isometry), R = S*V ;  ortho == "right": L = U*S, R = V.
    Returns L, R, Spectrum, u."""
    if which_decomp is None:
        if eigen_perturbation is not None:
            which_decomp = "eigen"
        elif cutoff is None or cutoff <= 1e-12:
            which_decomp = "svd"
        else:
            which_decomp = "eigen"
    if which_decomp == "svd":
        U, S, V, spec, u = svd_bs(T, left, maxdim, mindim, cutoff, tags)
        if ortho == "left":
            return U, _scale_link(V, u, S), spec, u
        return _scale_link(U, u, S), V, spec, u
    side = "left" if ortho == "left" else "right"
    Vt, spec, u = eigen_bs(T, left, side, eigen_perturbation, maxdim, mindim, cutoff, tags)
    if side == "left":
        L = Vt                                     # (left..., u)
        R = contract(L.dag(), T)                   # (u, right...)
        return L, R, spec, u
    R = Vt                                         # (u, right...)
    L = contract(T, R.dag())                       # (left..., u)
    return L, R, spec, u
