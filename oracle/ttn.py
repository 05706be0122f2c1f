"""Tree tensor networks on the CPU (oracle; test infrastructure only) -- SURVEY.md section 8 row a12.

Restates the reference's in-tree TTN path on the oracle's block-sparse tensors, for sites without quantum numbers
(every index one sector of charge 0: a dense tensor with named indices; BASELINE.json configs[4]) and for
QN-conserving sites with the dummy "QN" index at the centre node (the model of test/test_TTN.jl):
  /root/reference/src/base/graph.jl:196-438                 bfs, nodes_from_bfs, shortest_path, central node
  /root/reference/src/ttn/ttn_generators.jl:1-94,96-260      default_graph_sitenodes, randomTTN (both branches)
  /root/reference/src/ttn/ttn.jl:266-384                     moveisometry_to_next!, isometrize_full!, isometrize!
  /root/reference/src/ttn/linktensors.jl:35-221,231-262      LinkTensorsTTN(psi, M), move_linktensors*, product
  /root/reference/src/ttn/helper_internal_funcs.jl:22-47     _get_links
  /root/reference/src/ttn/linkproj.jl:14-224                 LinkProjTTN (excited states: weight * |M><M|)
  /root/reference/src/ttn/state_envs_ttn.jl:13-157           StateEnvsTTN, position!, product
  /root/reference/src/ttn/update_site_ttn.jl:42-109          update_position!, subspace_expand!
  /root/reference/src/ttn/sweep_ttn.jl:36-243                SweepDataTTN, default_sweeppath, fullsweep!
  /root/reference/src/ttn/optimize_ttn.jl:47-218             OptimizeParamsTTN, optimize!
This is the CPU statement the device path of the next round is to be checked against; nothing in the product
package uses it.  Deviations that do not change any result: Julia iterates Sets / Dicts in hash order, here nodes and
ids are visited in sorted order (sums over ids and tie-breaks between equally central nodes are order independent up
to rounding / gauge); `qr` is done with the untruncated SVD (another gauge of the same isometry); the random padding
of `subspace_expand!` draws from a NumPy generator.
"""
from __future__ import annotations

import itertools
from collections import deque
from typing import Dict, List, Sequence, Tuple

import numpy as np

from .blocksparse import BSTensor, Index, commoninds, contract, factorize, inner, uniqueinds
from .couplingmodel import CouplingModel, gen_id
from .dmrg import FLOAT64_THRESHOLD, eig_solver

Node = Tuple[int, int]


# ------------------------------------------------------------------------------------------ graph
class Graph:
    def __init__(self):
        self.adj: Dict[Node, set] = {}

    @property
    def nodes(self):
        return set(self.adj)

    def addedge(self, a: Node, b: Node):
        self.adj.setdefault(a, set()).add(b)
        self.adj.setdefault(b, set()).add(a)

    def __getitem__(self, node):
        return sorted(self.adj[node])

    def isneighbor(self, a, b):
        return b in self.adj.get(a, ())


def bfs(graph: Graph, source: Node, destination: Node | None = None):
    q = deque([source])
    visited = {source}
    parents: Dict[Node, Node] = {}
    dist = {source: 0}
    while q:
        node = q.popleft()
        for nb in graph[node]:
            if nb not in visited:
                q.append(nb)
                visited.add(nb)
                dist[nb] = dist[node] + 1
                parents[nb] = node
                if destination is not None and nb == destination:
                    return dist, parents
    return dist, parents


def nodes_from_bfs(graph: Graph, source: Node, destinations=None, reverse: bool = False) -> List[Node]:
    dist, parents = bfs(graph, source)
    if destinations is None:
        nodes = sorted(dist)
    else:
        res = set()
        for d in destinations:
            if d != source:
                cur = d
                res.add(cur)
                while cur in parents:
                    res.add(parents[cur])
                    cur = parents[cur]
            else:
                res.add(source)
        nodes = sorted(res)
    return sorted(nodes, key=lambda x: dist[x], reverse=reverse)


def shortest_path(graph: Graph, source: Node, destination: Node) -> List[Node]:
    if source == destination:
        return [source]
    _, parents = bfs(graph, source, destination)
    if destination not in parents:
        raise RuntimeError(f"`shortest_path()`: `destination={destination}` is not reachable from `source={source}` !!")
    out = [destination]
    while destination in parents:
        out.append(parents[destination])
        destination = parents[destination]
    return out[::-1]


def nextnode_in_path(graph: Graph, source: Node, destination: Node, n: int = 1) -> Node:
    return shortest_path(graph, source, destination)[n]


def find_eccentric_central_node(graph: Graph, nodes=None) -> Node:
    best, best_d = None, None
    for node in sorted(graph.nodes):
        dist, _ = bfs(graph, node)
        if nodes is not None:
            dist = {k: v for k, v in dist.items() if k in set(nodes)}
        md = max(dist.values())
        if best_d is None or md < best_d:
            best, best_d = node, md
    return best


# --------------------------------------------------------------------------------- default binary tree
def _minimum_power2_greater_than(n: int) -> int:
    p = 1
    while p < n:
        p *= 2
    return p


def _distribute_site_positions(numsites: int, loc=None) -> List[int]:
    """ttn_generators.jl:1-48."""
    if not loc:
        loc = [numsites // 2, numsites // 2 + numsites % 2]
        numsites = _minimum_power2_greater_than(numsites)
    else:
        newloc = []
        power = (len(loc) & -len(loc)).bit_length() - 1          # trailing_zeros
        for m in loc:
            rem = m % 2
            newloc.append(m // 2 + rem if power % 2 == 1 else m // 2)
            newloc.append(m // 2 if power % 2 == 1 else m // 2 + rem)
        loc = newloc
    numsites //= 2
    power = (numsites & -numsites).bit_length() - 1
    if power > 1:
        return _distribute_site_positions(numsites, loc)
    pos = []
    for m in loc:
        if m == 1:
            pos += [1, 0]
        elif m == 2:
            pos += [1, 1]
        else:
            raise RuntimeError("`_distribute_site_positions()`: SOMETHING IS WRONG !!")
    return pos


def default_graph_sitenodes(numsites: int):
    """ttn_generators.jl:69-94: hierarchical binary tree, two sites per bottom node, the two top nodes linked."""
    site_positions = _distribute_site_positions(numsites)
    pow2 = len(site_positions)
    numlayers = (pow2 & -pow2).bit_length() - 1
    sitenodes: Dict[int, Node] = {}
    graph = Graph()
    sitecount = 1
    for ll in range(1, numlayers - 1):
        for nn in range(1, (pow2 >> ll) + 1):
            if ll == 1 and (site_positions[2 * nn - 2] == 0 or site_positions[2 * nn - 1] == 0):
                sitenodes[sitecount] = (ll + 1, (nn + 1) // 2)
                sitecount += 1
                continue
            elif ll == 1:
                sitenodes[sitecount] = (ll, nn)
                sitecount += 1
                sitenodes[sitecount] = (ll, nn)
                sitecount += 1
            graph.addedge((ll, nn), (ll + 1, (nn + 1) // 2))
    graph.addedge((numlayers - 1, 1), (numlayers - 1, 2))
    return graph, sitenodes


# ------------------------------------------------------------------------------------------- TTN
def dense_index(dim: int, dir: int = +1, tags: str = "") -> Index:
    return Index([(0,)], [dim], dir=dir, tags=tags)


def dense_siteinds(N: int, d: int = 2) -> List[Index]:
    return [dense_index(d, +1, f"Site,n={j + 1}") for j in range(N)]


class TTN:
    """ttn.jl:13-40."""

    def __init__(self, sites, graph: Graph, tensors: Dict[Node, BSTensor], orthocenter=None):
        self.sites = list(sites)
        self.graph = graph
        self.tensors = dict(tensors)
        self.orthocenter = orthocenter

    def __getitem__(self, node):
        return self.tensors[node]

    def __setitem__(self, node, t):
        self.tensors[node] = t
        if self.orthocenter != node:
            self.orthocenter = None

    def copy(self):
        return TTN(self.sites, self.graph, {k: v.copy() for k, v in self.tensors.items()}, self.orthocenter)

    def numsites(self):
        return len(self.sites)

    def findsites(self, t: BSTensor) -> List[int]:
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
                if "Link" in ix.tags:
                    md = max(md, ix.dim)
        return md

    def normalize(self):
        t = self.tensors[self.orthocenter]
        self.tensors[self.orthocenter] = t.scale(1.0 / t.norm())


def random_ttn(sites, graph: Graph, sitenodes: Dict[int, Node], chi: int, rng: np.random.Generator) -> TTN:
    """randomTTN, non-QN branch (ttn_generators.jl:171-246): link dims min(prod of the other dims, chi)."""
    inds: Dict[Node, List[Index]] = {node: [] for node in graph.nodes}
    for b, s in enumerate(sites):
        inds[sitenodes[b + 1]].append(s)
    center = find_eccentric_central_node(graph, list(sitenodes.values()))
    for node in nodes_from_bfs(graph, center, reverse=True)[:-1]:
        nxt = nextnode_in_path(graph, node, center)
        dim = min(int(np.prod([ix.dim for ix in inds[node]])), chi)
        link = dense_index(dim, +1, f"Link,{node}")
        inds[node].append(link.copy(dir=-1))
        inds[nxt].append(link)
    tensors = {}
    for node in sorted(graph.nodes):
        t = BSTensor.random(inds[node], rng)
        tensors[node] = t.scale(1.0 / t.norm())
    ttn = TTN(sites, graph, tensors, None)
    isometrize_full(ttn, center, normalize=True, cutoff=0.0, maxdim=chi)
    return ttn


def _sectors_of(ix: Index) -> Dict[tuple, int]:
    out: Dict[tuple, int] = {}
    for q, d in zip(ix.qns, ix.dims):
        out[q] = out.get(q, 0) + d
    return out


def _fuse(a: Dict[tuple, int], b: Dict[tuple, int]) -> Dict[tuple, int]:
    out: Dict[tuple, int] = {}
    for qa, da in a.items():
        for qb, db in b.items():
            q = tuple(x + y for x, y in zip(qa, qb))
            out[q] = min(out.get(q, 0) + da * db, 1 << 40)
    return out


def random_ttn_qn(sites, graph: Graph, sitenodes: Dict[int, Node], chi: int, total_qn, rng: np.random.Generator) -> TTN:
    """randomTTN, QN branch (ttn_generators.jl:171-246): the centre node carries a dummy "QN" index with the global
    charge; every link keeps only the charge sectors that the sites below it can produce AND that the rest of the tree
    can complete to the global charge, with block dimensions reduced proportionally to a total of `chi`.  The
    reference reaches the same fixed point by 100 rounds of `_ttn_ind_cleanup!` / `_ttn_ind_reducedim!`
    (:96-170) with a gradually shrinking cap; here it is one bottom-up and one top-down pass (the start state is
    random either way, only its sector structure matters)."""
    nq = len(sites[0].qns[0])
    total_qn = tuple(total_qn)
    center = find_eccentric_central_node(graph, list(sitenodes.values()))
    order_up = nodes_from_bfs(graph, center, reverse=True)            # leaves first
    parent = {n: nextnode_in_path(graph, n, center) for n in order_up[:-1]}
    children: Dict[Node, List[Node]] = {n: [] for n in graph.nodes}
    for n, pnode in parent.items():
        children[pnode].append(n)
    site_of: Dict[Node, List[Index]] = {n: [] for n in graph.nodes}
    for b, sidx in enumerate(sites):
        site_of[sitenodes[b + 1]].append(sidx)
    zero = {(0,) * nq: 1}
    # bottom-up: charges (as seen by the parent, i.e. sum of dir*qn of everything below) each subtree can produce
    below: Dict[Node, Dict[tuple, int]] = {}
    for n in order_up:
        acc = dict(zero)
        for sidx in site_of[n]:
            acc = _fuse(acc, {tuple(sidx.dir * x for x in q): d for q, d in _sectors_of(sidx).items()})
        for c in sorted(children[n]):
            acc = _fuse(acc, below[c])
        below[n] = acc
    # top-down: sectors of the link above each node that the rest of the tree can complete
    link_sectors: Dict[Node, Dict[tuple, int]] = {}
    above: Dict[Node, Dict[tuple, int]] = {center: {total_qn: 1}}       # charge the node as a whole must carry upwards
    for n in nodes_from_bfs(graph, center):
        parts = [(None, {tuple(sidx.dir * x for x in q): d for q, d in _sectors_of(sidx).items()}) for sidx in site_of[n]]
        parts += [(c, below[c]) for c in sorted(children[n])]
        for k, (c, sect) in enumerate(parts):
            if c is None:
                continue
            others = dict(zero)
            for kk, (_, s2) in enumerate(parts):
                if kk != k:
                    others = _fuse(others, link_sectors.get(parts[kk][0], s2) if parts[kk][0] is not None else s2)
            allowed = {}
            for q, d in sect.items():
                comp = 0
                for up, du in above[n].items():
                    need = tuple(u - x for u, x in zip(up, q))
                    comp += others.get(need, 0) * du
                if comp > 0:
                    allowed[q] = min(d, comp, chi)
            qs = sorted(allowed)
            dims = _cap_dims([allowed[q] for q in qs], chi)
            link_sectors[c] = {q: d for q, d in zip(qs, dims) if d > 0}
        for c in sorted(children[n]):
            # what child c must deliver upwards = its link sectors
            above[c] = dict(link_sectors[c])
    inds: Dict[Node, List[Index]] = {n: list(site_of[n]) for n in graph.nodes}
    for n in order_up[:-1]:
        qs = sorted(link_sectors[n])
        link = Index(qs, [link_sectors[n][q] for q in qs], dir=+1, tags=f"Link,{n}")
        inds[n].append(link.copy(dir=-1))
        inds[parent[n]].append(link)
    inds[center].append(Index([total_qn], [1], dir=-1, tags="QN"))
    tensors = {}
    for n in sorted(graph.nodes):
        t = BSTensor.random(inds[n], rng)
        if not t.blocks:
            raise RuntimeError(f"randomTTN: node {n} has no symmetry-allowed block")
        tensors[n] = t.scale(1.0 / t.norm())
    ttn = TTN(sites, graph, tensors, None)
    isometrize_full(ttn, center, normalize=True, cutoff=0.0, maxdim=chi)
    return ttn


def default_random_ttn(sites, chi: int, rng: np.random.Generator, total_qn=None) -> TTN:
    graph, sitenodes = default_graph_sitenodes(len(sites))
    if sites[0].nsect > 1:
        nq = len(sites[0].qns[0])
        return random_ttn_qn(sites, graph, sitenodes, chi, (0,) * nq if total_qn is None else total_qn, rng)
    return random_ttn(sites, graph, sitenodes, chi, rng)


def moveisometry_to_next(ttn: TTN, node1: Node, node2: Node, *, ignore_orthocenter=False, maxdim=None, mindim=1,
                         cutoff=FLOAT64_THRESHOLD, svd_alg="divide_and_conquer", **_):
    """ttn.jl:266-310: SVD when a truncation can happen (cutoff > 0 or the link exceeds maxdim), QR otherwise."""
    if ttn.orthocenter != node1 and not ignore_orthocenter:
        raise RuntimeError(f"`moveisometry_to_next!()`: `orthocenter` does not match with the input `node1 = {node1} !!")
    if not ttn.graph.isneighbor(node1, node2):
        raise RuntimeError("`moveisometry_to_next!()`: Input nodes are not neighbors !!")
    A, B = ttn.tensors[node1], ttn.tensors[node2]
    com = commoninds(A, B)[0]
    uinds = uniqueinds(A, B)
    big = (1 << 62) if maxdim is None else maxdim
    if cutoff > 0.0 or com.dim > big:
        U, R, spec, u = factorize(A, uinds, ortho="left", which_decomp="svd", maxdim=maxdim, mindim=mindim,
                                  cutoff=cutoff, tags=com.tags)
    else:
        U, R, spec, u = factorize(A, uinds, ortho="left", which_decomp="svd", maxdim=None, mindim=1, cutoff=None,
                                  tags=com.tags)
    ttn.tensors[node1] = U
    ttn.tensors[node2] = contract(B, R)
    ttn.orthocenter = node2


def isometrize_full(ttn: TTN, node: Node, normalize=True, **kw):
    path = nodes_from_bfs(ttn.graph, node, reverse=True)
    for n in path[:-1]:
        nxt = nextnode_in_path(ttn.graph, n, node)
        moveisometry_to_next(ttn, n, nxt, ignore_orthocenter=True, **kw)
    ttn.orthocenter = node
    if normalize:
        ttn.normalize()


def isometrize(ttn: TTN, node: Node, normalize=True, **kw):
    if ttn.orthocenter is None:
        isometrize_full(ttn, node, normalize=normalize, **kw)
    if node != ttn.orthocenter:
        path = shortest_path(ttn.graph, ttn.orthocenter, node)
        for a, b in zip(path[:-1], path[1:]):
            moveisometry_to_next(ttn, a, b, **kw)
    if normalize:
        ttn.normalize()


def ttn_to_dense(ttn: TTN) -> np.ndarray:
    """Full state vector (small N; KAT helper), site 1 = most significant digit."""
    acc = None
    for node in nodes_from_bfs(ttn.graph, ttn.orthocenter or sorted(ttn.graph.nodes)[0]):
        acc = ttn.tensors[node] if acc is None else contract(acc, ttn.tensors[node])
    order = [next(ix for ix in acc.inds if ix.id == s.id) for s in ttn.sites]
    rest = [ix for ix in acc.inds if all(ix.id != s.id for s in ttn.sites)]
    assert all(ix.dim == 1 for ix in rest)
    return acc.permute(order + rest).to_dense().reshape(-1)


# ---------------------------------------------------------------------------------- link environments
def _link(a: Node, b: Node):
    return frozenset((a, b))


class LinkTensorsTTN(dict):
    """Dict{LinkTypeTTN, IDTensors}: per link the environment tensors of every term id (linktensors.jl:1-20)."""


def _get_links(psi: TTN, node: Node, nextnode: Node | None = None):
    nbrs = [x for x in psi.graph[node] if x != nextnode]
    nbrs += [(0, n) for n in psi.findsites(psi[node])]
    links = [_link(node, x) for x in nbrs]
    return (_link(node, nextnode), links) if nextnode is not None else links


def _collect(env: LinkTensorsTTN, links) -> Dict[int, List[BSTensor]]:
    idtens: Dict[int, List[BSTensor]] = {}
    for link in links:
        for tid, t in env.get(link, {}).items():
            idtens.setdefault(tid, []).append(t)
    return idtens


def move_linktensors_to_next(env: LinkTensorsTTN, psi: TTN, node: Node, nextnode: Node):
    """linktensors.jl:63-118: for every id on the other links of `node`, contract dag(prime(phi)) * tensors * phi; results
    that still carry an OpLink (order > 2) stay separate, the others are summed into one local tensor."""
    if not psi.graph.isneighbor(node, nextnode):
        raise RuntimeError("`move_linktensors_to_next!()`: Input nodes are not neighbors !!")
    next_link, prev_links = _get_links(psi, node, nextnode)
    idtens = _collect(env, prev_links)
    if not idtens:
        return
    env[next_link] = {}
    phi = psi[node]
    nextind = commoninds(phi, psi[nextnode])[0]
    local = None
    for tid in sorted(idtens):
        to_prime = [commoninds(phi, x)[0] for x in idtens[tid]] + [nextind]
        t = phi.prime(1, to_prime).dag()
        for x in idtens[tid]:
            t = contract(t, x)
        t = contract(t, phi)
        if t.rank > 2:
            env[next_link][tid] = t
        else:
            local = t if local is None else local.add(t)
    if local is not None:
        env[next_link][gen_id()] = local


def move_linktensors(env: LinkTensorsTTN, psi: TTN, source: Node, destination: Node, node_to_skip=None):
    if source == destination:
        return
    path = shortest_path(psi.graph, source, destination)
    for a, b in zip(path[:-1], path[1:]):
        if a == node_to_skip:
            continue
        move_linktensors_to_next(env, psi, a, b)


def link_tensors_from_model(psi: TTN, M: CouplingModel) -> LinkTensorsTTN:
    """LinkTensorsTTN(psi, M) (linktensors.jl:231-262): site operators sit on the virtual links (0, n) -- node; the
    environments are then pulled towards the orthogonality centre, farthest nodes first."""
    if psi.orthocenter is None:
        raise RuntimeError("`LinkTensorsTTN()`: TTN does not have a proper orthogonality center !!")
    env = LinkTensorsTTN()
    nodelist = set()
    for n in range(1, psi.numsites() + 1):
        node = psi.find_sitenode(n)
        if M[n]:
            env[_link((0, n), node)] = dict(M[n])
            nodelist.add(node)
    oc = psi.orthocenter
    path = nodes_from_bfs(psi.graph, oc, nodelist, reverse=True)
    for n1 in path[:-1]:
        move_linktensors_to_next(env, psi, n1, nextnode_in_path(psi.graph, n1, oc))
    return env


def product(env: LinkTensorsTTN, psi: TTN, v: BSTensor) -> BSTensor:
    """linktensors.jl:183-221: sum over ids of contract(v, link tensors of the id around the orthogonality centre)."""
    idtens = _collect(env, _get_links(psi, psi.orthocenter))
    out = None
    for tid in sorted(idtens):
        Hv = v
        for x in idtens[tid]:
            Hv = contract(Hv, x)
        Hv = Hv.noprime()
        out = Hv if out is None else out.add(Hv)
    if out.rank != v.rank:
        raise RuntimeError("The order of the LinkTensorsTTN-ITensor product P*v is not equal to the order of the ITensor v")
    return out


class LinkProjTTN:
    """Projector environments for a fixed TTN `M` (src/ttn/linkproj.jl:14-224): tensors[link] = overlap of the part
    of the tree behind the link, dag(prime(M[node]; tags = "Link")) * (tensors on the other links) * psi[node]."""

    def __init__(self, psi: TTN, M: TTN):
        assert [s.id for s in psi.sites] == [s.id for s in M.sites]
        if psi.orthocenter is None:
            raise RuntimeError("`LinkProjTTN()`: TTN does not have a proper orthogonality center !!")
        self.M = M
        self.tensors: Dict[frozenset, BSTensor] = {}
        # both trees must share the dummy QN index (linkproj.jl:201-211)
        for node, t in psi.tensors.items():
            qi = [ix for ix in t.inds if ix.tags == "QN"]
            if qi:
                qm = next(ix for tm in M.tensors.values() for ix in tm.inds if ix.tags == "QN")
                if not qm.same_space(qi[0]):
                    raise RuntimeError("`LinkProjTTN()`: TTNs do have same global QN !!")
                psi.tensors[node] = t.replaceinds([qi[0]], [qm])
        oc = psi.orthocenter
        path = nodes_from_bfs(psi.graph, oc, reverse=True)
        for n1 in path[:-1]:
            self.move_to_next(psi, n1, nextnode_in_path(psi.graph, n1, oc))

    def _mdag(self, node: Node) -> BSTensor:
        t = self.M[node]
        return t.prime(1, [ix for ix in t.inds if "Link" in ix.tags]).dag()

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

    def contract_v(self, psi: TTN, v: BSTensor | None) -> BSTensor:
        oc = psi.orthocenter
        t = self._mdag(oc)
        for x in _get_links(psi, oc):
            if x in self.tensors:
                t = contract(t, self.tensors[x])
        return t if v is None else contract(t, v)

    def product(self, psi: TTN, v: BSTensor) -> BSTensor:
        ov = self.contract_v(psi, v)                       # <M|v>, order 0
        m = self.contract_v(psi, None).dag()               # |m> in the indices of the centre tensor
        if ov.rank != 0 or m.rank != v.rank:
            raise RuntimeError("The order of the LinkProjTTN-ITensor product P*v is not equal to the order of the ITensor v")
        return m.scale(ov.scalar()).noprime()


class StateEnvsTTN:
    """state_envs_ttn.jl:13-66: StateEnvsTTN(psi, M::CouplingModel[, Ms::Vector{TTN}; weight])."""

    def __init__(self, psi: TTN, M: CouplingModel, Ms: Sequence[TTN] | None = None, weight: float = -1.0):
        self.psi = psi.copy()
        self.projs: List[LinkProjTTN] = []
        self.weight = weight
        if Ms:
            if weight <= 0.0:
                raise ValueError(f"`weight` parameter should be > 0.0 (value passed was `weight={weight}`)")
            self.projs = [LinkProjTTN(self.psi, m) for m in Ms]
        self.env = link_tensors_from_model(self.psi, M)

    def position(self, node: Node, *, maxdim=None, mindim=1, cutoff=FLOAT64_THRESHOLD, svd_alg="divide_and_conquer",
                 normalize=True, node_to_skip=None):
        oc = self.psi.orthocenter
        isometrize(self.psi, node, normalize=normalize, maxdim=maxdim, mindim=mindim, cutoff=cutoff, svd_alg=svd_alg)
        move_linktensors(self.env, self.psi, oc, node, node_to_skip=node_to_skip)
        for p in self.projs:
            p.move(self.psi, oc, node, node_to_skip=node_to_skip)

    def product(self, v: BSTensor) -> BSTensor:
        Pv = product(self.env, self.psi, v)
        for p in self.projs:                               # EnvCouplingModelProjTTN.product (environment.jl:95-101)
            Pv = Pv.add(p.product(self.psi, v).permute(Pv.inds), self.weight)
        return Pv

    __call__ = product


def update_position_ttn(sysenv: StateEnvsTTN, solver, node: Node, *, time_step=None, normalize=True, maxdim=None,
                        mindim=1, cutoff=FLOAT64_THRESHOLD, svd_alg="divide_and_conquer", **kw):
    """update_site_ttn.jl:42-63."""
    sysenv.position(node, normalize=normalize, maxdim=maxdim, mindim=mindim, cutoff=cutoff, svd_alg=svd_alg)
    phi = sysenv.psi[node]
    energy, phi = solver(sysenv, phi, time_step, **kw)
    sysenv.psi[node] = phi.permute(sysenv.psi[node].inds)
    return energy


def _fused_sectors(inds: Sequence[Index]) -> Dict[tuple, int]:
    """charge -> dimension of the fused space of `inds`, charge = sum_i dir_i * qn_i."""
    out: Dict[tuple, int] = {}
    nq = len(inds[0].qns[0])
    for combo in itertools.product(*[range(ix.nsect) for ix in inds]):
        q = tuple(sum(ix.dir * ix.qns[k][a] for ix, k in zip(inds, combo)) for a in range(nq))
        out[q] = out.get(q, 0) + int(np.prod([ix.dims[k] for ix, k in zip(inds, combo)]))
    return out


def _cap_dims(dims: List[int], maxdim: int) -> List[int]:
    """proportional reduction of block dimensions to a total of `maxdim` (src/base/helper_internal_funcs.jl:228-244)."""
    if sum(dims) <= maxdim:
        return dims
    tot = sum(dims)
    # Julia `round` is round-half-to-even, like Python's
    d = [max(1, int(round(maxdim * (x / tot)))) for x in dims]
    diff = sum(d) - maxdim
    if diff > 0:
        order = sorted(range(len(d)), key=lambda i: -d[i])
        for ii in range(min(diff, len(order))):
            if d[order[ii]] > 1:
                d[order[ii]] -= 1
    return d


def index_intersection_for_link(indsB: Sequence[Index], indsA: Sequence[Index], maxdim: int, dir: int, tags: str) -> Index:
    """`indexintersection(indsB, dag.(indsA); maxdim, dir)` (src/base/helper_internal_funcs.jl:171-247) as used by
    subspace_expand!: the sectors a new link between B(link, indsB...) and A(dag(link), indsA...) can carry.
    Without QNs the reference takes min(SUM of the dims on either side, maxdim) (:180-186); with QNs the common
    charges with the smaller of the two fused block dimensions, each capped at maxdim, then reduced proportionally."""
    if all(ix.nsect == 1 for ix in list(indsB) + list(indsA)):
        d = min(sum(ix.dim for ix in indsB), sum(ix.dim for ix in indsA), maxdim)
        return Index([indsB[0].qns[0]], [d], dir=dir, tags=tags)
    fb = {tuple(-dir * x for x in q): d for q, d in _fused_sectors(indsB).items()}     # flux 0 on B
    fa = {tuple(dir * x for x in q): d for q, d in _fused_sectors(indsA).items()}      # flux 0 on A (link dagged)
    qns, dims = [], []
    for q in sorted(fb):
        if q in fa:
            qns.append(q)
            dims.append(min(fb[q], fa[q], maxdim))
    if not qns:
        raise RuntimeError("`indexintersection()`: No common QN blocks present !!")
    dims = _cap_dims(dims, maxdim)
    keep = [(q, d) for q, d in zip(qns, dims) if d > 0]
    return Index([q for q, _ in keep], [d for _, d in keep], dir=dir, tags=tags)


def _directsum(A: BSTensor, ia: Index, P: BSTensor, ip: Index, tags: str):
    """ITensors `directsum(A => ia, P => ip)`: the new index lists the sectors of `ia`, then those of `ip`; the other
    indices are shared."""
    oa = [ix for ix in A.inds if ix != ia]
    Ap = A.permute(oa + [ia])
    Pp = P.permute([next(jx for jx in P.inds if jx == ix) for ix in oa] + [ip])
    new = Index(list(ia.qns) + list(ip.qns), list(ia.dims) + list(ip.dims), dir=ia.dir, tags=tags)
    out = BSTensor(oa + [new], dtype=np.result_type(A.dtype, P.dtype))
    for c, blk in Ap.blocks.items():
        out.blocks[c] = blk.copy()
    for c, blk in Pp.blocks.items():
        out.blocks[c[:-1] + (c[-1] + ia.nsect,)] = blk.copy()
    return out, new


def subspace_expand(psi: TTN, node: Node, nextnode: Node, max_expand_dim: int, noise: float, rng: np.random.Generator):
    """update_site_ttn.jl:75-109: pad the link between `node` and `nextnode` with random directions of relative size
    `noise` on both tensors."""
    A, B = psi.tensors[node], psi.tensors[nextnode]
    ind_to_update = commoninds(B, A)[0]
    indsA, indsB = uniqueinds(A, B), uniqueinds(B, A)
    ind_padB = index_intersection_for_link(indsB, indsA, max_expand_dim, ind_to_update.dir, "pad")
    padB = BSTensor.random([ind_padB] + list(indsB), rng)
    padB = padB.scale(noise * B.norm() / padB.norm())
    enlargedB, sumB = _directsum(B, ind_to_update, padB, ind_padB, ind_to_update.tags)
    psi.tensors[nextnode] = enlargedB
    ind_padA = ind_padB.copy(dir=-ind_padB.dir)
    padA = BSTensor.random([ind_padA] + list(indsA), rng)
    padA = padA.scale(noise * A.norm() / padA.norm())
    ia = next(ix for ix in A.inds if ix == ind_to_update)
    enlargedA, sumA = _directsum(A, ia, padA, ind_padA, ind_to_update.tags)
    psi.tensors[node] = enlargedA.replaceinds([sumA], [sumB.copy(dir=-sumB.dir)])


def product_ttn(sites: Sequence[Index], states: Sequence[int], graph: Graph | None = None,
                sitenodes: Dict[int, Node] | None = None) -> TTN:
    """Product state on the (default) tree with QN-conserving sites: every link is one sector of dimension 1
    carrying the charge of the sites below it; the centre node holds the dummy "QN" index with the total charge
    (randomTTN attaches the same index, ttn_generators.jl:221-223).  states[n] = sector number of site n+1."""
    if graph is None:
        graph, sitenodes = default_graph_sitenodes(len(sites))
    nq = len(sites[0].qns[0])
    inds: Dict[Node, List[Index]] = {node: [] for node in graph.nodes}
    coords: Dict[Node, List[int]] = {node: [] for node in graph.nodes}
    charge: Dict[Node, tuple] = {node: (0,) * nq for node in graph.nodes}
    for b, s in enumerate(sites):
        node = sitenodes[b + 1]
        inds[node].append(s)
        coords[node].append(states[b])
        charge[node] = tuple(a + s.dir * c for a, c in zip(charge[node], s.qns[states[b]]))
    center = find_eccentric_central_node(graph, list(sitenodes.values()))
    for node in nodes_from_bfs(graph, center, reverse=True)[:-1]:
        nxt = nextnode_in_path(graph, node, center)
        link = Index([charge[node]], [1], dir=+1, tags=f"Link,{node}")
        inds[node].append(link.copy(dir=-1))
        coords[node].append(0)
        inds[nxt].append(link)
        coords[nxt].append(0)
        charge[nxt] = tuple(a + c for a, c in zip(charge[nxt], charge[node]))
    qnindex = Index([charge[center]], [1], dir=-1, tags="QN")
    inds[center].append(qnindex)
    coords[center].append(0)
    tensors = {node: BSTensor(inds[node], {tuple(coords[node]): np.ones([1] * len(inds[node]))}) for node in graph.nodes}
    return TTN(sites, graph, tensors, center)


# ------------------------------------------------------------------------------------------ sweeps
class SweepDataTTN:
    def __init__(self):
        self.sweepcount = 0
        self.maxchi: List[int] = []
        self.energy: List[float] = []


def default_sweeppath(psi: TTN) -> List[Node]:
    """sweep_ttn.jl:36-52."""
    path = []
    nsites = _minimum_power2_greater_than(psi.numsites())
    nlayers = (nsites & -nsites).bit_length() - 1
    for ll in range(nlayers - 1, 0, -1):
        for nn in range(1, (nsites >> ll) + 1):
            nnpos = nn if (nlayers - ll) % 2 == 1 else (nsites >> ll) - nn + 1
            node = (ll, nnpos)
            if ll == 1 and node not in psi.graph.nodes:
                continue
            path.append(node)
    return path


def fullsweep_ttn(sysenv: StateEnvsTTN, sweeppath: Sequence[Node], solver, swdata: SweepDataTTN, rng=None, **kw):
    """sweep_ttn.jl:94-243: plain sweep (noise == 0: isometry moves by QR, truncating only links above maxdim) or the
    subspace-expansion sweep (noise > 0: pad the link towards the centre, then alternate `expand_numiter` updates of the
    two nodes with a shrinking maxdim)."""
    assert set(sweeppath) == sysenv.psi.graph.nodes
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
    if expand_dim != 0 and expand_numiter < 2:
        raise RuntimeError(f"`fullsweep!()`: `expand_numiter={expand_numiter}` cannot be less than 2 for `expand_dim={expand_dim}` !!")
    energy = float("nan")
    swdata.sweepcount += 1
    common = dict(time_step=time_step, normalize=normalize, mindim=mindim, svd_alg=svd_alg)
    if abs(noise) < 100 * FLOAT64_THRESHOLD:
        for node in list(sweeppath) + list(sweeppath)[::-1]:
            energy = update_position_ttn(sysenv, solver, node, maxdim=maxdim, cutoff=-1.0, **common, **kw)
    else:
        rng = rng or np.random.default_rng(0)
        central = find_eccentric_central_node(sysenv.psi.graph)
        big = (1 << 62) if maxdim is None else maxdim
        for ii in list(range(len(sweeppath))) + list(range(len(sweeppath) - 1, -1, -1)):
            node = sweeppath[ii]
            if node == central:
                energy = update_position_ttn(sysenv, solver, node, maxdim=maxdim, cutoff=-1.0, **common, **kw)
                continue
            nextnode = nextnode_in_path(sysenv.psi.graph, node, central)
            sysenv.position(node, normalize=normalize, maxdim=maxdim, mindim=mindim, cutoff=-1.0, svd_alg=svd_alg,
                            node_to_skip=nextnode)
            subspace_expand(sysenv.psi, node, nextnode, max_expand_dim, noise, rng)
            sysenv.psi.orthocenter = nextnode
            link = _link(node, nextnode)
            linkmax = linkwise_maxdim[link] if linkwise_maxdim and link in linkwise_maxdim else big
            for dummy in range(1, expand_numiter + 1):
                newmax = linkmax if dummy == expand_numiter else (linkmax + max_expand_dim if dummy == 1 else linkmax + expand_dim)
                newnode = node if dummy % 2 == 1 else nextnode
                energy = update_position_ttn(sysenv, solver, newnode, maxdim=newmax, cutoff=cutoff, **common, **kw)
    swdata.maxchi.append(sysenv.psi.maxlinkdim())
    swdata.energy.append(energy)
    return swdata.energy[-1] - swdata.energy[-2] if swdata.sweepcount > 1 else float("nan")


class OptimizeParamsTTN:
    """optimize_ttn.jl:18-99."""

    def __init__(self, *, maxdim, nsweeps, cutoff=FLOAT64_THRESHOLD, noise=0.0, noisedecay=1.0,
                 disable_noise_after=(1 << 62)):
        n = len(nsweeps)
        vec = lambda x, T: [T(v) for v in x] if isinstance(x, (list, tuple)) else [T(x)] * n
        self.maxdim, self.nsweeps = list(maxdim), list(nsweeps)
        self.cutoff, self.noise = vec(cutoff, float), vec(noise, float)
        self.noisedecay, self.disable_noise_after = vec(noisedecay, float), vec(disable_noise_after, int)
        if not (len(self.maxdim) == n == len(self.cutoff) == len(self.noise) == len(self.noisedecay)
                == len(self.disable_noise_after)):
            raise ValueError("`OptimizeParamsTTN()`: Size mismatch in input vectors !!")


def optimize_(sysenv: StateEnvsTTN, params: OptimizeParamsTTN, sweeppath: Sequence[Node], rng=None, **kw) -> SweepDataTTN:
    """`optimize!` (optimize_ttn.jl:148-218): stage / sweep loop with the noise schedule of `dmrg!`."""
    enerrgoal = kw.pop("energyErrGoal", None)
    swdata = SweepDataTTN()
    for ii in range(len(params.nsweeps)):
        maxdim, cutoff, noise = params.maxdim[ii], params.cutoff[ii], params.noise[ii]
        noisedecay, disable_after = params.noisedecay[ii], params.disable_noise_after[ii]
        for jj in range(1, params.nsweeps[ii] + 1):
            enerr = fullsweep_ttn(sysenv, sweeppath, eig_solver, swdata, rng=rng, maxdim=maxdim, cutoff=cutoff,
                                  noise=noise, **kw)
            if enerrgoal is not None and abs(enerr) < abs(enerrgoal) and abs(noise) < FLOAT64_THRESHOLD:
                break
            if jj == disable_after:
                noise = 0.0
            noise /= noisedecay
            if noise < 100 * FLOAT64_THRESHOLD:
                noise = 0.0
    return swdata


def optimize(psi0: TTN, H: CouplingModel, params: OptimizeParamsTTN, sweeppath: Sequence[Node], rng=None, Ms=None,
             weight: float = -1.0, **kw):
    sysenv = StateEnvsTTN(psi0, H, Ms, weight)
    sw = optimize_(sysenv, params, sweeppath, rng=rng, **kw)
    return sw.energy[-1], sysenv.psi, sw
