"""Measurements on a device-resident MPS -- mirror of /root/reference/src/mps/measure.jl:4-343 (`entropy`,
`bond_spectrum`, `measure` for one operator tensor, an operator name at one / all sites, a list of single-site operator
tensors): the state never leaves HBM, every contraction is a call of the generic device algebra, the host receives the
spectrum or one scalar.

Device tensors carry flux 0.  A single-site operator with a non-zero flux (S+, Cdag, ...) is therefore given a dim-1
index that carries its flux and links it to the next operator of the product -- the same device `_add_oplinks!` uses for
CouplingModel terms (src/base/helper_internal_funcs.jl:41-63); a product whose total flux does not vanish has
expectation value zero in a charge eigenstate and returns 0 without touching the device."""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

from .gse import DeviceMPS, mps_from_env
from .itensor import ITensor, contract, factorize, same_index
from .tensor import HostTensor, Index

FLOAT64_THRESHOLD = 1e-14


def _entropy(p) -> float:
    p = np.asarray(p, dtype=np.float64)
    s = p.sum()
    if abs(s - 1.0) > 100 * FLOAT64_THRESHOLD:
        p = p / s
    p = p[p > 0]
    return float(-np.sum(p * np.log(p)))


def as_device_mps(psi, sites=None) -> DeviceMPS:
    """DeviceMPS as is; a StateEnvs (with the site indices of its Hamiltonian) is wrapped without copying tensors."""
    if isinstance(psi, DeviceMPS):
        return DeviceMPS(psi.t, psi.center)
    if sites is None:
        raise ValueError("measure on a StateEnvs needs the site indices")
    return mps_from_env(psi, [_SiteOnly(s) for s in sites])


class _SiteOnly:
    """stands in for an MPO tensor W(wl, s', s, wr) where only the site identity is needed"""

    def __init__(self, s):
        self.inds = [None, s, s, None]


def _bond_svd(psi: DeviceMPS, bond: int):
    if not (0 < bond < len(psi)):
        raise ValueError("bond out of range")
    psi.orthogonalize(bond)
    A = psi.t[bond - 1]
    return factorize(A, A.inds[:2], ortho="left", which_decomp="svd", cutoff=None, tags="Link,u")


def bond_spectrum(psi, bond=None, *, bonds=None, by_charge: bool = False, sites=None):
    """measure.jl:24-138: squared Schmidt values at `bond` (descending), or per charge sector of the bond."""
    psi = as_device_mps(psi, sites)
    if bond is None:
        bonds = list(range(1, len(psi))) if bonds is None else bonds
        return [bond_spectrum(psi, b, by_charge=by_charge) for b in bonds]
    U, R, spec, link = _bond_svd(psi, bond)
    if not by_charge:
        return np.asarray(spec.eigs)
    if link.nsect == 1 and all(q == 0 for q in link.qns[0]) and psi.t[bond - 1].inds[1].nsect == 1:
        raise RuntimeError("`bond_spectrum()`: `by_charge` cannot be `true` for QN non-conserving MPS !!")
    # per sector: the squared singular values are the squared row norms of R = S V in that sector
    G = contract(R, R.prime(1, [R.inds[0]]).dag()).to_host()          # (m, m') block diagonal
    out = []
    for k, q in enumerate(link.qns):
        blk = G.blocks.get((k, k))
        w = np.sort(np.real(np.diag(blk)))[::-1] if blk is not None else np.zeros(0)
        qn = tuple(q) if link.dir == -1 else tuple(-x for x in q)
        out.append((qn, w))
    return out


def entropy(psi, bond=None, *, bonds=None, sites=None):
    """measure.jl:61-97: von Neumann entropy at one bond / at all (or the given) bonds."""
    psi = as_device_mps(psi, sites)
    if bond is not None:
        return _entropy(bond_spectrum(psi, bond))
    bonds = list(range(1, len(psi))) if bonds is None else bonds
    return [_entropy(bond_spectrum(psi, b)) for b in bonds]


def _op_flux(op) -> tuple:
    """flux of a host operator tensor (all its blocks carry the same)"""
    nq = len(op.inds[0].qns[0])
    for c, b in op.blocks.items():
        if np.any(np.asarray(b) != 0):
            return tuple(sum(ix.dir * ix.qns[k][a] for ix, k in zip(op.inds, c)) for a in range(nq))
    return (0,) * nq


def _site_pos(psi: DeviceMPS, op) -> int:
    ids = {ix.id for ix in op.inds}
    for j, A in enumerate(psi.t):
        if A.inds[1].id in ids:
            return j + 1
    raise RuntimeError("`measure`: Error in operator tensors !!")


def _check_op(psi: DeviceMPS, op, pos: int):
    s = psi.t[pos - 1].inds[1]
    ok = len(op.inds) == 2 and {(ix.id, ix.plev) for ix in op.inds} == {(s.id, 0), (s.id, 1)}
    if not ok:
        raise RuntimeError("`measure`: Error in operator tensors !!")


def _product_on_site(a, b):
    """opdict[pos] = mapprime(prime(a) * b, 2, 1): the operator product a b on one site (host, d x d)."""
    sp = next(ix for ix in a.inds if ix.plev == 1)
    sk = next(ix for ix in a.inds if ix.plev == 0)

    def dense(o):
        perm = [next(k for k, ix in enumerate(o.inds) if ix.plev == p) for p in (1, 0)]
        return np.transpose(o.to_dense(), perm)
    m = dense(a) @ dense(b)
    offs = np.concatenate([[0], np.cumsum(sp.dims)])
    blocks = {}
    for i in range(len(sp.dims)):
        for j in range(len(sk.dims)):
            blk = m[offs[i]:offs[i + 1], offs[j]:offs[j + 1]]
            if np.any(blk != 0):
                blocks[(i, j)] = blk.copy()
    return HostTensor([sp, sk], blocks)


def _finish(val, real: bool):
    if real:
        return float(np.real(val))
    return complex(val)


def measure(psi, op, pos=None, *, sites=None, real: bool = False, op_table=None):
    """measure.jl:143-343.
      measure(psi, opten)                   one single-site operator tensor (s', s)
      measure(psi, "Sz", pos) / (psi, "Sz") operator name at one site / at all sites (`op_table(name, site_index)` builds
                                            the tensor, as ITensors `op`)
      measure(psi, [opten, ...])            product of single-site operator tensors (several on one site are multiplied)
    `real=True` is `measure(Float64, ...)`."""
    psi = as_device_mps(psi, sites)
    if isinstance(op, str):
        if op_table is None:
            raise ValueError("measure with an operator name needs `op_table(name, site_index)`")
        if pos is None:
            return [measure(psi, op_table(op, psi.t[j].inds[1]), real=real) for j in range(len(psi))]
        return measure(psi, op_table(op, psi.t[pos - 1].inds[1]), real=real)
    if not isinstance(op, (list, tuple)):
        op = [op]
    opdict: Dict[int, object] = {}
    for o in op:
        p = _site_pos(psi, o)
        _check_op(psi, o, p)
        opdict[p] = o if p not in opdict else _product_on_site(opdict[p], o)
    order = sorted(opdict)
    nq = len(psi.t[0].inds[1].qns[0])
    total = tuple(sum(_op_flux(opdict[p])[a] for p in order) for a in range(nq))
    if any(total):
        return _finish(0.0, real)
    ctx = psi.t[0].ctx
    # operators with their flux-carrying links (dim 1): link k joins operator k and k + 1
    ops: Dict[int, ITensor] = {}
    run = (0,) * nq
    prev_link = None
    for k, p in enumerate(order):
        o = opdict[p]
        inds, blocks = list(o.inds), {c: np.asarray(b) for c, b in o.blocks.items()}
        if prev_link is not None:
            inds = [prev_link.copy(dir=+1)] + inds
            blocks = {(0,) + c: b[None] for c, b in blocks.items()}
        run = tuple(r + f for r, f in zip(run, _op_flux(o)))
        if k + 1 < len(order):
            prev_link = Index([run], [1], dir=-1, tags="OpLink")
            inds = inds + [prev_link]
            blocks = {c + (0,): b[..., None] for c, b in blocks.items()}
        ops[p] = ITensor.from_host(ctx, HostTensor(inds, blocks), nrow=1)
    minpos, maxpos = order[0], order[-1]
    psi.orthogonalize(minpos)
    if minpos == maxpos:
        ket = psi.t[minpos - 1]
        opket = contract(ops[minpos], ket).noprime()
        return _finish(ket.inner(opket), real)
    A = psi.t[minpos - 1]
    ir = A.inds[2]
    C = contract(contract(A, ops[minpos]), A.prime(1, [A.inds[1], ir]).dag())
    for p in range(minpos + 1, maxpos):
        A = psi.t[p - 1]
        C = contract(C, A)
        if p in ops:
            C = contract(contract(C, ops[p]), A.prime().dag())
        else:
            C = contract(C, A.prime(1, [A.inds[0], A.inds[2]]).dag())
    A = psi.t[maxpos - 1]
    X = contract(contract(C, A), ops[maxpos])                       # (jl', s', r)
    bra = A.prime(1, [A.inds[0], A.inds[1]])
    return _finish(bra.inner(X), real)
