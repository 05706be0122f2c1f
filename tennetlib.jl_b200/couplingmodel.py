"""`CouplingModel` container and its ingestion into the device format -- host-side mirror of
/root/reference/src/base/couplingmodel.jl:14-17 (struct: sites + terms::Vector{IDTensors}).  Building a model from
OpStrings (`_initCouplingModel`, :120-230) is host-side model construction and stays with the caller; what the
device needs is every term tensor in the canonical (wl, s', s, wr) form of `tnl_env_cm_set_term`."""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

from .tensor import HostTensor, Index


class CouplingModel:
    """sites::Vector{Index}, terms[site] = {id: tensor}.  Term tensors carry their site indices (s' with plev 1, s
    with plev 0) and 0, 1 or 2 indices tagged "OpLink"."""

    def __init__(self, sites: Sequence, terms: List[Dict[int, object]]):
        self.sites = list(sites)
        self.terms = terms

    def __len__(self):
        return len(self.sites)

    def __getitem__(self, j):          # 1-based like the reference
        return self.terms[j - 1]


def is_coupling_model(H) -> bool:
    return hasattr(H, "sites") and hasattr(H, "terms")


def _is_oplink(ix) -> bool:
    return "OpLink" in ix.tags


def _same(a, b) -> bool:
    return a.id == b.id and a.plev == b.plev


def canonical_terms(model) -> List[Dict[int, HostTensor]]:
    """Every term tensor as (wl, s', s, wr): wl = the OpLink shared with the term's previous tensor, wr = the one
    shared with its next tensor, a dim-1 charge-0 index where there is none."""
    N = len(model.sites)
    support: Dict[int, List[int]] = {}
    for j in range(N):
        for tid in model.terms[j]:
            support.setdefault(tid, []).append(j)
    out: List[Dict[int, HostTensor]] = [dict() for _ in range(N)]
    for tid, pos in support.items():
        for k, j in enumerate(pos):
            t = model.terms[j][tid]
            links = [ix for ix in t.inds if _is_oplink(ix)]
            sites = [ix for ix in t.inds if not _is_oplink(ix)]
            if len(sites) != 2 or len(links) > 2:
                raise ValueError(f"CouplingModel term {tid} on site {j + 1}: expected (s', s) and at most two OpLinks")
            sp = next(ix for ix in sites if ix.plev == 1)
            sk = next(ix for ix in sites if ix.plev == 0)

            def shared(other):
                if other is None:
                    return None
                for ix in links:
                    if any(_same(ix, jx) for jx in other.inds):
                        return ix
                return None
            wl = shared(model.terms[pos[k - 1]][tid]) if k > 0 else None
            wr = shared(model.terms[pos[k + 1]][tid]) if k + 1 < len(pos) else None
            if len([x for x in (wl, wr) if x is not None]) != len(links):
                raise ValueError(f"CouplingModel term {tid} on site {j + 1}: dangling OpLink")
            nq = len(sk.qns[0])
            order = [ix for ix in (wl, sp, sk, wr) if ix is not None]
            perm = [next(n for n, jx in enumerate(t.inds) if jx is ix) for ix in order]
            triv_l = Index([(0,) * nq], [1], dir=+1, tags="OpLink,trivial")
            triv_r = Index([(0,) * nq], [1], dir=-1, tags="OpLink,trivial")
            inds = [wl if wl is not None else triv_l, sp, sk, wr if wr is not None else triv_r]
            blocks = {}
            for c, b in t.blocks.items():
                cc = tuple(c[p] for p in perm)
                b = np.asarray(b)
                if np.iscomplexobj(b):
                    if np.abs(b.imag).max(initial=0.0) > 0.0:
                        raise NotImplementedError("complex CouplingModel tensors are not supported on the device: only "
                                                  "the state and the environments may be ComplexF64")
                    b = b.real
                bb = np.transpose(np.asarray(b, dtype=np.float64), perm)
                if wl is None:
                    cc, bb = (0,) + cc, bb[None]
                if wr is None:
                    cc, bb = cc + (0,), bb[..., None]
                blocks[cc] = np.ascontiguousarray(bb)
            ht = HostTensor(inds, blocks)
            ht.has_wl, ht.has_wr = wl is not None, wr is not None
            out[j][tid] = ht
    return out
