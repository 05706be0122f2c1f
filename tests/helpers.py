"""Shared test helpers: conversion between the product's host tensors and the oracle's block-sparse tensors."""
import numpy as np

_OFFSET = 10 ** 9          # product and oracle index ids come from independent counters


def to_oracle(t):
    """product HostTensor (or any .inds/.blocks object with product Index) -> oracle BSTensor; index identity
    (id, plev) is preserved so that tensors converted separately still contract with each other."""
    from oracle import blocksparse as ob
    inds = [ob.Index(ix.qns, ix.dims, dir=ix.dir, tags=ix.tags, plev=ix.plev, id=_OFFSET + ix.id) for ix in t.inds]
    cplx = any(np.iscomplexobj(b) for b in t.blocks.values())
    return ob.BSTensor(inds, {c: np.array(b) for c, b in t.blocks.items()}, np.complex128 if cplx else np.float64)


def to_oracle_index(ix):
    from oracle import blocksparse as ob
    return ob.Index(ix.qns, ix.dims, dir=ix.dir, tags=ix.tags, plev=ix.plev, id=_OFFSET + ix.id)


def mpo_dense(H):
    """d^N x d^N matrix of an MPO given as tensors with .to_dense() -> (wl, s', s, wr)."""
    acc = None
    for W in H:
        Wd = W.to_dense()
        acc = Wd[0] if acc is None else np.tensordot(acc, Wd, axes=([-1], [0]))
    acc = acc[..., 0]
    n = acc.ndim // 2
    acc = np.transpose(acc, [2 * k for k in range(n)] + [2 * k + 1 for k in range(n)])
    D = int(np.prod(acc.shape[:n]))
    return acc.reshape(D, D)


def random_qn_mps(sites, links, rng):
    """oracle MPS tensors (random, all allowed blocks) for oracle `sites` and link indices."""
    from oracle import blocksparse as ob
    return [ob.BSTensor.random([links[j].copy(dir=+1), sites[j].copy(dir=+1), links[j + 1].copy(dir=-1)], rng)
            for j in range(len(sites))]
