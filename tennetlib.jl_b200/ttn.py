"""Tree tensor networks on the GPU -- host-side mirror of the reference's TTN interface; every tensor lives in HBM
and every numerical step is a call into libtnl_b200.so (generic algebra + the device Krylov loops over a
sum-of-products operator).  SURVEY.md section 8 rows a12 / a13 / f4.

  /root/reference/src/ttn/ttn.jl:13-40,266-384             TTN, moveisometry_to_next!, isometrize_full!, isometrize!
  /root/reference/src/ttn/linktensors.jl:35-262             LinkTensorsTTN, move_linktensors(_to_next)!, product
  /root/reference/src/ttn/helper_internal_funcs.jl:22-47    _get_links
  /root/reference/src/ttn/linkproj.jl:14-224                LinkProjTTN (excited states)
  /root/reference/src/ttn/environment.jl:14-112             EnvCouplingModelTTN / EnvCouplingModelProjTTN
  /root/reference/src/ttn/state_envs_ttn.jl:13-157          StateEnvsTTN, position!, product
  /root/reference/src/ttn/update_site_ttn.jl:42-109         update_position!, subspace_expand!
  /root/reference/src/ttn/sweep_ttn.jl:36-243               SweepDataTTN, default_sweeppath, fullsweep!
  /root/reference/src/ttn/optimize_ttn.jl:18-218            OptimizeParamsTTN, optimize!
  /root/reference/src/base/helper_internal_funcs.jl:171-247 indexintersection

The effective Hamiltonian of a node is handed to the device once per update as a `tnl_sumop_t` (one term per id: the
link tensors of that id around the node); `eig_solver` / `exp_solver` then run the same device Lanczos loops as for
the MPS (`tnl_sumop_eigsolve`, `tnl_sumop_exponentiate`) -- the host sees one scalar per solve."""
from __future__ import annotations

import ctypes as C
import itertools
from typing import Dict, List, Sequence

import numpy as np

from ._lib import check
from .graph import (Graph, Node, default_graph_sitenodes, find_eccentric_central_node, nextnode_in_path,
                    nodes_from_bfs, shortest_path)
from .itensor import (ITensor, _label, commonind, contract, directsum, factorize, hastags, same_index, uniqueinds)
from .tensor import Context, Index

FLOAT64_THRESHOLD = 1e-14          # Float64_threshold(), src/base/global_variables.jl
_ids = itertools.count(1 << 40)    # gen_rand_id() stand-in for the summed local tensors
_BIG = 1 << 62                     # typemax(Int) stand-in


def _link(a: Node, b: Node):
    return frozenset((a, b))


class TTN:
    """ttn.jl:13-40: sites, graph, tensors::Dict{Int2, ITensor} (device tensors), orthocenter."""

    def __init__(self, sites, graph: Graph, tensors: Dict[Node, ITensor], orthocenter=None):
        self.sites = list(sites)
        self.graph = graph
        self.tensors = dict(tensors)
        self.orthocenter = orthocenter

    @staticmethod
    def from_host(ctx: Context, sites, graph: Graph, tensors: Dict[Node, object], orthocenter=None) -> "TTN":
        return TTN(sites, graph, {n: ITensor.from_host(ctx, t) for n, t in tensors.items()}, orthocenter)

    def to_host(self) -> Dict[Node, object]:
        return {n: t.to_host() for n, t in self.tensors.items()}

    def __getitem__(self, node):
        return self.tensors[node]

    def __setitem__(self, node, t):
        self.tensors[node] = t
        if self.orthocenter != node:
            self.orthocenter = None

    def copy(self):
        return TTN(self.sites, self.graph, dict(self.tensors), self.orthocenter)      # shallow, like Base.copy(::TTN)

    def numsites(self):
        return len(self.sites)

    def findsites(self, t: ITensor) -> List[int]:
        ids = {ix.id for ix in t.inds}
        return [n + 1 for n, s in enumerate(self.sites) if s.id in ids]

    def find_sitenode(self, n: int) -> Node:
        sid = self.sites[n - 1].id
        for node in sorted(self.tensors):
            if any(ix.id == sid for ix in self.tensors[node].inds):
                return node
        raise KeyError(n)

    def maxlinkdim(self) -> int:
        md = 1
        for t in self.tensors.values():
            for ix in t.inds:
                if hastags(ix, "Link"):
                    md = max(md, ix.dim)
        return md

    def normalize(self):
        """normalize!(ttn) (ttn.jl:240-247): scales the orthogonality-centre tensor."""
        t = self.tensors[self.orthocenter]
        n = t.norm()
        self.tensors[self.orthocenter] = t.fresh().scale_(1.0 / n)


# ---------------------------------------------------------------------------------------------- isometry moves
def moveisometry_to_next(ttn: TTN, node1: Node, node2: Node, *, ignore_orthocenter=False, maxdim=None, mindim=1,
                         cutoff=FLOAT64_THRESHOLD, svd_alg="divide_and_conquer", **_):
    """ttn.jl:266-310: SVD when a truncation can happen (cutoff > 0 or the link exceeds maxdim), QR gauge move otherwise;
    the non-isometric factor is absorbed into node2."""
    if ttn.orthocenter != node1 and not ignore_orthocenter:
        raise RuntimeError(f"`moveisometry_to_next!()`: `orthocenter` does not match with the input `node1 = {node1} !!")
    if not ttn.graph.isneighbor(node1, node2):
        raise RuntimeError("`moveisometry_to_next!()`: Input nodes are not neighbors !!")
    A, B = ttn.tensors[node1], ttn.tensors[node2]
    com = commonind(A, B)
    uinds = uniqueinds(A, B)
    big = _BIG if maxdim is None else maxdim
    if cutoff > 0.0 or com.dim > big:
        U, R, spec, _ = factorize(A, uinds, ortho="left", which_decomp="svd", maxdim=maxdim, mindim=mindim, cutoff=cutoff,
                                  tags=com.tags, svd_alg=svd_alg)
    else:
        U, R, spec, _ = factorize(A, uinds, ortho="left", which_decomp="qr", tags=com.tags)
    ttn.tensors[node1] = U
    ttn.tensors[node2] = contract(B, R)
    ttn.orthocenter = node2


def isometrize_full(ttn: TTN, node: Node, normalize=True, **kw):
    """ttn.jl:326-349: gauge moves from the farthest nodes inwards."""
    for n in nodes_from_bfs(ttn.graph, node, reverse=True)[:-1]:
        moveisometry_to_next(ttn, n, nextnode_in_path(ttn.graph, n, node), ignore_orthocenter=True, **kw)
    ttn.orthocenter = node
    if normalize:
        ttn.normalize()


def isometrize(ttn: TTN, node: Node, normalize=True, **kw):
    """ttn.jl:361-384."""
    if ttn.orthocenter is None:
        isometrize_full(ttn, node, normalize=normalize, **kw)
    if node != ttn.orthocenter:
        path = shortest_path(ttn.graph, ttn.orthocenter, node)
        for a, b in zip(path[:-1], path[1:]):
            moveisometry_to_next(ttn, a, b, **kw)
    if normalize:
        ttn.normalize()


# ------------------------------------------------------------------------------------------- link environments
class LinkTensorsTTN(dict):
    """Dict{LinkTypeTTN, IDTensors} (linktensors.jl:1-20): per link the environment tensors of every term id."""


def _get_links(psi: TTN, node: Node, nextnode: Node | None = None):
    """helper_internal_funcs.jl:22-47: the links of `node` (tree neighbours and the virtual links (0, n) of its sites),
    split off the one towards `nextnode`."""
    others = [x for x in psi.graph[node] if x != nextnode] + [(0, n) for n in psi.findsites(psi[node])]
    links = [_link(node, x) for x in others]
    return (_link(node, nextnode), links) if nextnode is not None else links


def _collect(env: LinkTensorsTTN, links) -> Dict[int, List[ITensor]]:
    idtens: Dict[int, List[ITensor]] = {}
    for link in links:
        for tid, t in env.get(link, {}).items():
            idtens.setdefault(tid, []).append(t)
    return idtens


def move_linktensors_to_next(env: LinkTensorsTTN, psi: TTN, node: Node, nextnode: Node):
    """linktensors.jl:63-118: per id `dag(prime(phi; inds touched + next link)) * tensors... * phi`; results that still
    carry an OpLink (order > 2) stay separate, the closed ones are summed into one local tensor (flat axpy)."""
    if not psi.graph.isneighbor(node, nextnode):
        raise RuntimeError("`move_linktensors_to_next!()`: Input nodes are not neighbors !!")
    next_link, prev_links = _get_links(psi, node, nextnode)
    idtens = _collect(env, prev_links)
    if not idtens:
        return
    env[next_link] = {}
    phi = psi[node]
    nextind = commonind(phi, psi[nextnode])
    local = None
    for tid in sorted(idtens):
        to_prime = [commonind(phi, x) for x in idtens[tid]] + [nextind]
        t = phi.prime(1, to_prime).dag()
        for x in idtens[tid]:
            t = contract(t, x)
        t = contract(t, phi)
        if t.rank > 2:
            env[next_link][tid] = t
        elif local is None:
            local = t
        else:
            local.add_(t)
    if local is not None:
        env[next_link][next(_ids)] = local


def move_linktensors(env: LinkTensorsTTN, psi: TTN, source: Node, destination: Node, node_to_skip=None):
    """linktensors.jl:130-147."""
    if source == destination:
        return
    path = shortest_path(psi.graph, source, destination)
    for a, b in zip(path[:-1], path[1:]):
        if a == node_to_skip:
            continue
        move_linktensors_to_next(env, psi, a, b)


def link_tensors_from_model(psi: TTN, M) -> LinkTensorsTTN:
    """LinkTensorsTTN(psi, M::CouplingModel) (linktensors.jl:231-262): the term tensors of site n sit on the virtual
    link (0, n) -- node; the environments are pulled towards the orthogonality centre, farthest nodes first."""
    if psi.orthocenter is None:
        raise RuntimeError("`LinkTensorsTTN()`: TTN does not have a proper orthogonality center !!")
    ctx = next(iter(psi.tensors.values())).ctx
    env = LinkTensorsTTN()
    nodelist = set()
    for n in range(1, psi.numsites() + 1):
        node = psi.find_sitenode(n)
        if M[n]:
            env[_link((0, n), node)] = {tid: (t if isinstance(t, ITensor) else ITensor.from_host(ctx, t))
                                        for tid, t in M[n].items()}
            nodelist.add(node)
    oc = psi.orthocenter
    for n1 in nodes_from_bfs(psi.graph, oc, nodelist, reverse=True)[:-1]:
        move_linktensors_to_next(env, psi, n1, nextnode_in_path(psi.graph, n1, oc))
    return env


def product(env: LinkTensorsTTN, psi: TTN, v: ITensor) -> ITensor:
    """linktensors.jl:183-221 evaluated contraction by contraction (each one a grouped DGEMM on the device).  The
    solvers do not use this: they hand the same terms to the device as one operator (`StateEnvsTTN.sumop`)."""
    idtens = _collect(env, _get_links(psi, psi.orthocenter))
    out = None
    for tid in sorted(idtens):
        Hv = v
        for x in idtens[tid]:
            Hv = contract(Hv, x)
        Hv = Hv.noprime()
        if out is None:
            out = Hv.like(v)
        else:
            out.add_(Hv)
    if out is None or out.rank != v.rank:
        raise RuntimeError("The order of the LinkTensorsTTN-ITensor product P*v is not equal to the order of the ITensor v")
    return out


class LinkProjTTN:
    """linkproj.jl:14-224: overlaps of a fixed TTN `M` with the running state, link by link:
    tensors[link] = dag(prime(M[node]; tags = "Link")) * (tensors on the other links) * psi[node]."""

    def __init__(self, psi: TTN, M: TTN):
        if [s.id for s in psi.sites] != [s.id for s in M.sites]:
            raise RuntimeError("`LinkProjTTN()`: TTNs do not share the site indices !!")
        if psi.orthocenter is None:
            raise RuntimeError("`LinkProjTTN()`: TTN does not have a proper orthogonality center !!")
        self.M = M
        self.tensors: Dict[frozenset, ITensor] = {}
        for node, t in psi.tensors.items():          # both trees must share the dummy QN index (linkproj.jl:201-211)
            qi = [ix for ix in t.inds if ix.tags == "QN"]
            if qi:
                qm = next(ix for tm in M.tensors.values() for ix in tm.inds if ix.tags == "QN")
                if tuple(qm.qns) != tuple(qi[0].qns) or tuple(qm.dims) != tuple(qi[0].dims):
                    raise RuntimeError("`LinkProjTTN()`: TTNs do have same global QN !!")
                psi.tensors[node] = t.replaceinds([qi[0]], [qm])
        oc = psi.orthocenter
        for n1 in nodes_from_bfs(psi.graph, oc, reverse=True)[:-1]:
            self.move_to_next(psi, n1, nextnode_in_path(psi.graph, n1, oc))

    def _mdag(self, node: Node) -> ITensor:
        return self.M[node].prime(1, tags="Link").dag()

    def move_to_next(self, psi: TTN, node: Node, nextnode: Node):
        next_link, prev_links = _get_links(psi, node, nextnode)
        t = self._mdag(node)
        for x in prev_links:
            if x in self.tensors:
                t = contract(t, self.tensors[x])
        self.tensors[next_link] = contract(t, psi[node])

    def move(self, psi: TTN, source: Node, destination: Node, node_to_skip=None):
        if source == destination:
            return
        path = shortest_path(psi.graph, source, destination)
        for a, b in zip(path[:-1], path[1:]):
            if a == node_to_skip:
                continue
            self.move_to_next(psi, a, b)

    def ket(self, psi: TTN) -> ITensor:
        """|m> in the indices of the centre tensor: dag(dag(prime(M[oc])) * link overlaps) (linkproj.jl:150-179)."""
        oc = psi.orthocenter
        t = self._mdag(oc)
        for x in _get_links(psi, oc):
            if x in self.tensors:
                t = contract(t, self.tensors[x])
        return t.dag().noprime()

    def product(self, psi: TTN, v: ITensor) -> ITensor:
        m = self.ket(psi).like(v)
        ov = m.inner(v)
        if np.iscomplexobj(ov) and abs(ov.imag) > 0:
            raise NotImplementedError("complex projector overlaps: use the device operator (sumop)")
        return m.fresh().scale_(float(np.real(ov)))


class StateEnvsTTN:
    """state_envs_ttn.jl:13-66: StateEnvsTTN(psi, M::CouplingModel[, Ms::Vector{TTN}; weight]).  `psi` is copied
    (shallow, as in the reference); all environments are built on the device."""

    def __init__(self, psi: TTN, M, Ms: Sequence[TTN] | None = None, weight: float = -1.0):
        self.psi = psi.copy()
        self.ctx = next(iter(self.psi.tensors.values())).ctx
        self.projs: List[LinkProjTTN] = []
        self.weight = weight
        if Ms:
            if weight <= 0.0:
                raise ValueError(f"`weight` parameter should be > 0.0 (value passed was `weight={weight}`)")
            self.projs = [LinkProjTTN(self.psi, m) for m in Ms]
        self.env = link_tensors_from_model(self.psi, M)
        self.last_solver_info = {}
        self._op = None

    def getpsi(self) -> TTN:
        return self.psi.copy()

    def position(self, node: Node, *, maxdim=None, mindim=1, cutoff=FLOAT64_THRESHOLD, svd_alg="divide_and_conquer",
                 normalize=True, node_to_skip=None):
        """position! (state_envs_ttn.jl:120-140): move the isometry centre, then the environments along the same path."""
        oc = self.psi.orthocenter
        isometrize(self.psi, node, normalize=normalize, maxdim=maxdim, mindim=mindim, cutoff=cutoff, svd_alg=svd_alg)
        move_linktensors(self.env, self.psi, oc, node, node_to_skip=node_to_skip)
        for p in self.projs:
            p.move(self.psi, oc, node, node_to_skip=node_to_skip)

    def product(self, v: ITensor) -> ITensor:
        """sysenv(v) (state_envs_ttn.jl:150-157; EnvCouplingModelProjTTN.product environment.jl:95-101)."""
        op = _SumOp(self, v)
        try:
            return op.apply(v)
        finally:
            op.close()

    __call__ = product


class _SumOp:
    """`tnl_sumop_t` for the effective Hamiltonian at the orthogonality centre."""

    def __init__(self, sysenv: StateEnvsTTN, v: ITensor):
        psi, ctx = sysenv.psi, sysenv.ctx
        self.ctx = ctx
        self.keep = []
        vl = np.ascontiguousarray([_label(ix) for ix in v.inds], dtype=np.int32)
        h = C.c_void_p()
        check(ctx.lib.tnl_sumop_create(ctx.h, v.rank, vl.ctypes.data, C.byref(h)), ctx.h)
        self.h = h
        idtens = _collect(sysenv.env, _get_links(psi, psi.orthocenter))
        for tid in sorted(idtens):
            ops = [x.materialize() for x in idtens[tid]]
            self.keep += ops
            arr = (C.c_void_p * len(ops))(*[x.dt.h for x in ops])
            labels = np.ascontiguousarray([_label(ix) for x in ops for ix in x.inds], dtype=np.int32)
            check(ctx.lib.tnl_sumop_add_term(self.h, len(ops), arr, labels.ctypes.data), ctx.h)
        frm = np.ascontiguousarray([_label(ix.prime()) for ix in v.inds], dtype=np.int32)
        check(ctx.lib.tnl_sumop_set_relabel(self.h, v.rank, frm.ctypes.data, vl.ctypes.data), ctx.h)
        for p in sysenv.projs:
            m = p.ket(psi).like(v)
            self.keep.append(m)
            check(ctx.lib.tnl_sumop_add_projector(self.h, m.dt.h, float(sysenv.weight)), ctx.h)

    def apply(self, v: ITensor) -> ITensor:
        h = C.c_void_p()
        check(self.ctx.lib.tnl_sumop_apply(self.h, v.dt.h, C.byref(h)), self.ctx.h)
        from .tensor import DeviceTensor
        return ITensor(DeviceTensor(self.ctx, h, v.inds), v.inds)

    def apply_flops(self) -> float:
        out = C.c_double()
        check(self.ctx.lib.tnl_sumop_apply_flops(self.h, C.byref(out)), self.ctx.h)
        return out.value

    def close(self):
        if self.h:
            self.ctx.lib.tnl_sumop_destroy(self.h)
            self.h = None


def eig_solver(sysenv: StateEnvsTTN, phi0: ITensor, time_step=None, **kwargs):
    """src/base/solver.jl:23-43 for a tree node: KrylovKit Lanczos `eigsolve` on the device over the node's operator."""
    if time_step is not None:
        raise TypeError("`eig_solver()` is only defined with `time_step=nothing`")
    if kwargs.get("solver_which_eigenvalue", "SR") != "SR" or not kwargs.get("ishermitian", True):
        raise NotImplementedError("device eig_solver supports which=:SR, ishermitian=true")
    phi = phi0.materialize().copy()
    op = _SumOp(sysenv, phi)
    try:
        ev, conv, nops, nit, nres = C.c_double(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_double()
        check(op.ctx.lib.tnl_sumop_eigsolve(op.h, phi.dt.h, float(kwargs.get("solver_tol", 1e-14)),
                                            int(kwargs.get("solver_krylovdim", 5)), int(kwargs.get("solver_maxiter", 2)),
                                            1 if kwargs.get("solver_eager", False) else 0, C.byref(ev), C.byref(conv),
                                            C.byref(nops), C.byref(nit), C.byref(nres)), op.ctx.h)
        sysenv.last_solver_info = dict(converged=conv.value, numops=nops.value, numiter=nit.value, normres=nres.value,
                                       apply_flops=op.apply_flops())
    finally:
        op.close()
    if kwargs.get("solver_check_convergence", False) and conv.value < 1:
        raise RuntimeError("`eig_solver()` not converged !!")
    return ev.value, phi


def exp_solver(sysenv: StateEnvsTTN, phi0: ITensor, time_step, **kwargs):
    """src/base/solver.jl:66-88 for a tree node (`tnl_sumop_exponentiate`)."""
    if time_step is None:
        raise RuntimeError(f"`exp_solver()` is not defined with `time_step={time_step}` !!")
    t = complex(time_step)
    phi = phi0.materialize().copy()
    op = _SumOp(sysenv, phi)
    try:
        conv, nops, nit, err = C.c_int32(), C.c_int32(), C.c_int32(), C.c_double()
        check(op.ctx.lib.tnl_sumop_exponentiate(op.h, phi.dt.h, t.real, t.imag, float(kwargs.get("solver_tol", 1e-12)),
                                                int(kwargs.get("solver_krylovdim", 30)),
                                                int(kwargs.get("solver_maxiter", 100)),
                                                1 if kwargs.get("solver_eager", True) else 0, C.byref(conv), C.byref(nops),
                                                C.byref(nit), C.byref(err)), op.ctx.h)
        sysenv.last_solver_info = dict(converged=conv.value, numops=nops.value, numiter=nit.value, normres=err.value,
                                       apply_flops=op.apply_flops())
    finally:
        op.close()
    return float("nan"), phi


def update_position(sysenv: StateEnvsTTN, solver, node: Node, *, time_step=None, normalize=True, maxdim=None, mindim=1,
                    cutoff=FLOAT64_THRESHOLD, svd_alg="divide_and_conquer", **kw):
    """update_position! (update_site_ttn.jl:42-63)."""
    sysenv.position(node, normalize=normalize, maxdim=maxdim, mindim=mindim, cutoff=cutoff, svd_alg=svd_alg)
    energy, phi = solver(sysenv, sysenv.psi[node], time_step, **kw)
    sysenv.psi[node] = phi
    return energy


# ------------------------------------------------------------------------------------------- subspace expansion
def _fused_sectors(inds: Sequence[Index]) -> Dict[tuple, int]:
    out: Dict[tuple, int] = {}
    nq = len(inds[0].qns[0])
    for combo in itertools.product(*[range(ix.nsect) for ix in inds]):
        q = tuple(sum(ix.dir * ix.qns[k][a] for ix, k in zip(inds, combo)) for a in range(nq))
        out[q] = out.get(q, 0) + int(np.prod([ix.dims[k] for ix, k in zip(inds, combo)]))
    return out


def _cap_dims(dims: List[int], maxdim: int) -> List[int]:
    """proportional reduction of the block dimensions to a total of `maxdim` (helper_internal_funcs.jl:228-244)."""
    tot = sum(dims)
    if tot <= maxdim:
        return dims
    d = [max(1, int(round(maxdim * (x / tot)))) for x in dims]        # round-half-even like Julia
    excess = sum(d) - maxdim
    if excess > 0:
        for i in sorted(range(len(d)), key=lambda i: -d[i])[:excess]:
            if d[i] > 1:
                d[i] -= 1
    return d


def indexintersection(indsB: Sequence[Index], indsA: Sequence[Index], maxdim: int, dir: int, tags: str = "pad") -> Index:
    """`indexintersection(indsB, dag.(indsA); maxdim, dir)` (helper_internal_funcs.jl:171-247): the sectors a new link
    between B(link, indsB...) and A(dag(link), indsA...) can carry."""
    if all(ix.nsect == 1 for ix in list(indsB) + list(indsA)):
        d = min(sum(ix.dim for ix in indsB), sum(ix.dim for ix in indsA), maxdim)
        return Index([indsB[0].qns[0]], [d], dir=dir, tags=tags)
    fb = {tuple(-dir * x for x in q): d for q, d in _fused_sectors(indsB).items()}
    fa = {tuple(dir * x for x in q): d for q, d in _fused_sectors(indsA).items()}
    common = [(q, min(fb[q], fa[q], maxdim)) for q in sorted(fb) if q in fa]
    if not common:
        raise RuntimeError("`indexintersection()`: No common QN blocks present !!")
    dims = _cap_dims([d for _, d in common], maxdim)
    keep = [(q, d) for (q, _), d in zip(common, dims) if d > 0]
    return Index([q for q, _ in keep], [d for _, d in keep], dir=dir, tags=tags)


_seed = itertools.count(0x5EED)


def subspace_expand(psi: TTN, node: Node, nextnode: Node, max_expand_dim: int, noise: float, seed: int | None = None):
    """subspace_expand! (update_site_ttn.jl:75-109): pad the link node -- nextnode on both tensors with random
    directions of relative weight `noise` (device random fill + device direct sum)."""
    A, B = psi.tensors[node].materialize(), psi.tensors[nextnode].materialize()
    ctx = A.ctx
    ind_to_update = commonind(B, A)
    indsA, indsB = uniqueinds(A, B), uniqueinds(B, A)
    ind_padB = indexintersection(indsB, indsA, max_expand_dim, ind_to_update.dir)
    s = next(_seed) if seed is None else seed
    padB = ITensor.random(ctx, [ind_padB] + list(indsB), 2 * s + 1)
    padB.scale_(noise * B.norm() / padB.norm())
    enlargedB, sumB = directsum(B, ind_to_update, padB, ind_padB, tags=ind_to_update.tags)
    psi.tensors[nextnode] = enlargedB
    ind_padA = ind_padB.dag()
    padA = ITensor.random(ctx, [ind_padA] + list(indsA), 2 * s + 2)
    padA.scale_(noise * A.norm() / padA.norm())
    enlargedA, sumA = directsum(A, A.inds[A.find(ind_to_update)], padA, ind_padA)
    psi.tensors[node] = enlargedA.replaceinds([sumA], [sumB.dag()])


# -------------------------------------------------------------------------------------------------------- sweeps
class SweepDataTTN:
    """sweep_ttn.jl:14-26."""

    def __init__(self):
        self.sweepcount = 0
        self.maxchi: List[int] = []
        self.energy: List[float] = []


def default_sweeppath(psi: TTN) -> List[Node]:
    """sweep_ttn.jl:36-52: layer by layer from the top, alternating direction."""
    p = 1
    while p < psi.numsites():
        p <<= 1
    nlayers = (p & -p).bit_length() - 1
    path = []
    for ll in range(nlayers - 1, 0, -1):
        width = p >> ll
        for nn in range(1, width + 1):
            node = (ll, nn if (nlayers - ll) % 2 == 1 else width - nn + 1)
            if ll == 1 and node not in psi.graph.nodes:
                continue
            path.append(node)
    return path


def fullsweep(sysenv: StateEnvsTTN, sweeppath: Sequence[Node], solver, swdata: SweepDataTTN, **kw):
    """fullsweep! (sweep_ttn.jl:94-243): plain sweep (noise == 0) or the subspace-expansion sweep (noise > 0: pad the link
    towards the centre, then `expand_numiter` alternating updates of the two nodes with a shrinking maxdim)."""
    if set(sweeppath) != sysenv.psi.graph.nodes:
        raise RuntimeError("`fullsweep!()`: `sweeppath` must visit every node of the TTN !!")
    kw = dict(kw)
    kw.pop("outputlevel", None)
    time_step = kw.pop("time_step", None)
    maxdim = kw.pop("maxdim", None)
    mindim = kw.pop("mindim", 1)
    cutoff = kw.pop("cutoff", FLOAT64_THRESHOLD)
    svd_alg = kw.pop("svd_alg", "divide_and_conquer")
    normalize = kw.pop("normalize", True)
    noise = kw.pop("noise", 0.0)
    expand_dim = kw.pop("expand_dim", 0 if abs(noise) < 100 * FLOAT64_THRESHOLD else 20)
    max_expand_dim = kw.pop("max_expand_dim", 2 * expand_dim)
    expand_numiter = kw.pop("expand_numiter", 4)
    linkwise_maxdim = kw.pop("linkwise_maxdim", None)
    seed = kw.pop("seed", None)
    if expand_dim != 0 and expand_numiter < 2:
        raise RuntimeError(f"`fullsweep!()`: `expand_numiter={expand_numiter}` cannot be less than 2 for "
                           f"`expand_dim={expand_dim}` !!")
    energy = float("nan")
    swdata.sweepcount += 1
    common = dict(time_step=time_step, normalize=normalize, mindim=mindim, svd_alg=svd_alg)
    there_and_back = list(sweeppath) + list(sweeppath)[::-1]
    if abs(noise) < 100 * FLOAT64_THRESHOLD:
        for node in there_and_back:
            energy = update_position(sysenv, solver, node, maxdim=maxdim, cutoff=-1.0, **common, **kw)
    else:
        central = find_eccentric_central_node(sysenv.psi.graph)
        big = _BIG if maxdim is None else maxdim
        for step, node in enumerate(there_and_back):
            if node == central:
                energy = update_position(sysenv, solver, node, maxdim=maxdim, cutoff=-1.0, **common, **kw)
                continue
            nextnode = nextnode_in_path(sysenv.psi.graph, node, central)
            sysenv.position(node, normalize=normalize, maxdim=maxdim, mindim=mindim, cutoff=-1.0, svd_alg=svd_alg,
                            node_to_skip=nextnode)
            subspace_expand(sysenv.psi, node, nextnode, max_expand_dim, noise,
                            None if seed is None else seed + 7919 * swdata.sweepcount + step)
            sysenv.psi.orthocenter = nextnode
            link = _link(node, nextnode)
            linkmax = linkwise_maxdim[link] if linkwise_maxdim and link in linkwise_maxdim else big
            for it in range(1, expand_numiter + 1):
                newmax = linkmax if it == expand_numiter else linkmax + (max_expand_dim if it == 1 else expand_dim)
                energy = update_position(sysenv, solver, node if it % 2 == 1 else nextnode, maxdim=newmax, cutoff=cutoff,
                                         **common, **kw)
    swdata.maxchi.append(sysenv.psi.maxlinkdim())
    swdata.energy.append(energy)
    return swdata.energy[-1] - swdata.energy[-2] if swdata.sweepcount > 1 else float("nan")


class OptimizeParamsTTN:
    """optimize_ttn.jl:18-99."""

    def __init__(self, *, maxdim, nsweeps, cutoff=FLOAT64_THRESHOLD, noise=0.0, noisedecay=1.0,
                 disable_noise_after=_BIG):
        n = len(nsweeps)
        vec = lambda x, T: [T(v) for v in x] if isinstance(x, (list, tuple)) else [T(x)] * n
        self.maxdim, self.nsweeps = list(maxdim), list(nsweeps)
        self.cutoff, self.noise = vec(cutoff, float), vec(noise, float)
        self.noisedecay, self.disable_noise_after = vec(noisedecay, float), vec(disable_noise_after, int)
        if not (len(self.maxdim) == n == len(self.cutoff) == len(self.noise) == len(self.noisedecay)
                == len(self.disable_noise_after)):
            raise ValueError("`OptimizeParamsTTN()`: Size mismatch in input vectors !!")


def optimize_(sysenv: StateEnvsTTN, params: OptimizeParamsTTN, sweeppath: Sequence[Node], **kw) -> SweepDataTTN:
    """`optimize!` (optimize_ttn.jl:148-218): stage / sweep loop with the noise schedule of `dmrg!`."""
    enerrgoal = kw.pop("energyErrGoal", None)
    swdata = SweepDataTTN()
    for ii in range(len(params.nsweeps)):
        maxdim, cutoff, noise = params.maxdim[ii], params.cutoff[ii], params.noise[ii]
        noisedecay, disable_after = params.noisedecay[ii], params.disable_noise_after[ii]
        for jj in range(1, params.nsweeps[ii] + 1):
            enerr = fullsweep(sysenv, sweeppath, eig_solver, swdata, maxdim=maxdim, cutoff=cutoff, noise=noise, **kw)
            if enerrgoal is not None and abs(enerr) < abs(enerrgoal) and abs(noise) < FLOAT64_THRESHOLD:
                break
            if jj == disable_after:
                noise = 0.0
            noise /= noisedecay
            if noise < 100 * FLOAT64_THRESHOLD:
                noise = 0.0
    return swdata


def optimize(psi0: TTN, H, params: OptimizeParamsTTN, sweeppath: Sequence[Node], Ms=None, weight: float = -1.0, **kw):
    """`optimize(psi0, H, params, sweeppath[, Ms; weight])` -> (energy, psi, swdata)."""
    sysenv = StateEnvsTTN(psi0, H, Ms, weight)
    sw = optimize_(sysenv, params, sweeppath, **kw)
    return sw.energy[-1], sysenv.psi, sw


# ------------------------------------------------------------------------- generators (bench / test workloads)
def dense_index(dim: int, dir: int = +1, tags: str = "") -> Index:
    return Index([(0,)], [dim], dir=dir, tags=tags)


def dense_siteinds(N: int, d: int = 2) -> List[Index]:
    """Site indices without quantum numbers (one charge-0 sector of dimension d)."""
    return [dense_index(d, +1, f"Site,n={j + 1}") for j in range(N)]


def tfi_coupling_model(sites: Sequence[Index], h: float = 1.0, J: float = 1.0):
    """H = -J sum_j Z_j Z_{j+1} - h sum_j X_j as a `CouplingModel` (BASELINE.json configs[4]): one id per bond with a
    dim-1 OpLink between its two tensors, one id per field term (src/base/couplingmodel.jl:120-230 builds the same
    structure from OpStrings)."""
    from .couplingmodel import CouplingModel
    from .tensor import HostTensor
    Z = np.array([[1.0, 0.0], [0.0, -1.0]])
    X = np.array([[0.0, 1.0], [1.0, 0.0]])
    N = len(sites)
    terms: List[Dict[int, object]] = [dict() for _ in range(N)]
    tid = itertools.count(1)
    for j in range(N - 1):
        k = next(tid)
        link = Index([(0,)], [1], dir=-1, tags="OpLink")
        sa, sb = sites[j], sites[j + 1]
        terms[j][k] = HostTensor([sa.prime().copy(dir=+1), sa.copy(dir=-1), link], {(0, 0, 0): (-J * Z)[:, :, None].copy()})
        terms[j + 1][k] = HostTensor([link.copy(dir=+1), sb.prime().copy(dir=+1), sb.copy(dir=-1)], {(0, 0, 0): Z[None].copy()})
    for j in range(N):
        s = sites[j]
        terms[j][next(tid)] = HostTensor([s.prime().copy(dir=+1), s.copy(dir=-1)], {(0, 0): (-h * X).copy()})
    return CouplingModel(sites, terms)


def random_ttn(ctx: Context, sites, graph: Graph, sitenodes: Dict[int, Node], chi: int, seed: int = 0) -> TTN:
    """randomTTN without quantum numbers (ttn_generators.jl:171-246): link dimensions min(product of the other
    dimensions, chi), random tensors on the device, isometrised towards the most central node and normalised."""
    inds: Dict[Node, List[Index]] = {node: [] for node in graph.nodes}
    for b, s in enumerate(sites):
        inds[sitenodes[b + 1]].append(s)
    center = find_eccentric_central_node(graph, list(sitenodes.values()))
    for node in nodes_from_bfs(graph, center, reverse=True)[:-1]:
        nxt = nextnode_in_path(graph, node, center)
        dim = min(int(np.prod([float(ix.dim) for ix in inds[node]])), chi)
        link = dense_index(dim, +1, f"Link,{node}")
        inds[node].append(link.dag())
        inds[nxt].append(link)
    tensors = {}
    for k, node in enumerate(sorted(graph.nodes)):
        t = ITensor.random(ctx, inds[node], seed * 100003 + k + 1)
        tensors[node] = t.scale_(1.0 / t.norm())
    psi = TTN(sites, graph, tensors, None)
    isometrize_full(psi, center, normalize=True, cutoff=0.0, maxdim=chi)
    return psi


def default_random_ttn(ctx: Context, sites, chi: int, seed: int = 0) -> TTN:
    """default_randomTTN (ttn_generators.jl:248-281) for sites without quantum numbers."""
    graph, sitenodes = default_graph_sitenodes(len(sites))
    return random_ttn(ctx, sites, graph, sitenodes, chi, seed)
