"""Undirected graphs of tree tensor networks -- host-side mirror of /root/reference/src/base/graph.jl:14-438
(`Graph`, `addedge!`, `isneighbor`, `bfs`, `nodes_from_bfs`, `shortest_path`, `nextnode_in_path`,
`find_eccentric_central_node`) and of the default hierarchical binary tree of
/root/reference/src/ttn/ttn_generators.jl:1-94 (`default_graph_sitenodes`).  Pure bookkeeping, no tensors.

The reference iterates Julia `Set`s (hash order); here neighbours are visited in sorted order, which fixes the
tie-breaks between equally short paths / equally central nodes deterministically."""
from __future__ import annotations

from collections import deque
from typing import Dict, Iterable, List, Tuple

Node = Tuple[int, int]


class Graph:
    def __init__(self, edges: Iterable[Tuple[Node, Node]] = ()):
        self.adj: Dict[Node, set] = {}
        for a, b in edges:
            self.addedge(a, b)

    @property
    def nodes(self):
        return set(self.adj)

    def addedge(self, a: Node, b: Node):
        if a == b:
            raise ValueError("`addedge!()`: self loops are not allowed !!")
        self.adj.setdefault(a, set()).add(b)
        self.adj.setdefault(b, set()).add(a)

    def __getitem__(self, node: Node) -> List[Node]:
        return sorted(self.adj[node])

    def isneighbor(self, a: Node, b: Node) -> bool:
        return b in self.adj.get(a, ())

    def edges(self):
        return sorted({tuple(sorted((a, b))) for a in self.adj for b in self.adj[a]})


def bfs(graph: Graph, source: Node, destination: Node | None = None):
    """Breadth-first search: (distance, parent) maps; stops early once `destination` is reached."""
    dist = {source: 0}
    parent: Dict[Node, Node] = {}
    todo = deque([source])
    while todo:
        cur = todo.popleft()
        for nb in graph[cur]:
            if nb in dist:
                continue
            dist[nb] = dist[cur] + 1
            parent[nb] = cur
            if nb == destination:
                return dist, parent
            todo.append(nb)
    return dist, parent


def nodes_from_bfs(graph: Graph, source: Node, destinations=None, reverse: bool = False) -> List[Node]:
    """All nodes (or the nodes on the paths source -> destinations) ordered by distance from `source`."""
    dist, parent = bfs(graph, source)
    if destinations is None:
        sel = set(dist)
    else:
        sel = set()
        for d in destinations:
            sel.add(d)
            while d in parent:
                d = parent[d]
                sel.add(d)
    return sorted(sorted(sel), key=dist.__getitem__, reverse=reverse)


def shortest_path(graph: Graph, source: Node, destination: Node) -> List[Node]:
    if source == destination:
        return [source]
    _, parent = bfs(graph, source, destination)
    if destination not in parent:
        raise RuntimeError(f"`shortest_path()`: `destination={destination}` is not reachable from `source={source}` !!")
    path = [destination]
    while path[-1] in parent:
        path.append(parent[path[-1]])
    return path[::-1]


def nextnode_in_path(graph: Graph, source: Node, destination: Node, n: int = 1) -> Node:
    return shortest_path(graph, source, destination)[n]


def find_eccentric_central_node(graph: Graph, nodes=None) -> Node:
    """Node with the smallest maximal distance to `nodes` (default: all nodes)."""
    only = None if nodes is None else set(nodes)
    best = None
    for cand in sorted(graph.nodes):
        dist, _ = bfs(graph, cand)
        ecc = max(v for k, v in dist.items() if only is None or k in only)
        if best is None or ecc < best[0]:
            best = (ecc, cand)
    return best[1]


def _next_pow2(n: int) -> int:
    p = 1
    while p < n:
        p <<= 1
    return p


def _site_positions(numsites: int) -> List[int]:
    """Occupation (1 / 0) of the 2^k bottom slots: the sites are split as evenly as possible, halving level by level
    (ttn_generators.jl:1-48)."""
    counts = [numsites // 2, numsites // 2 + numsites % 2]
    size = _next_pow2(numsites) // 2
    while (size & -size).bit_length() - 1 > 1:
        odd_level = ((len(counts) & -len(counts)).bit_length() - 1) % 2 == 1
        nxt = []
        for m in counts:
            lo, hi = m // 2, m // 2 + m % 2
            nxt += [hi, lo] if odd_level else [lo, hi]
        counts = nxt
        size //= 2
    slots = []
    for m in counts:
        if m not in (1, 2):
            raise RuntimeError("`_distribute_site_positions()`: SOMETHING IS WRONG !!")
        slots += [1, 1] if m == 2 else [1, 0]
    return slots


def default_graph_sitenodes(numsites: int):
    """(graph, sitenodes): binary tree with two sites per bottom node (a lone site hangs one layer higher), the two
    top nodes joined by an edge; nodes are (layer, position), 1-based (ttn_generators.jl:69-94)."""
    slots = _site_positions(numsites)
    width = len(slots)
    nlayers = (width & -width).bit_length() - 1
    graph = Graph()
    sitenodes: Dict[int, Node] = {}
    site = 1
    for layer in range(1, nlayers - 1):
        for pos in range(1, (width >> layer) + 1):
            up = (layer + 1, (pos + 1) // 2)
            if layer == 1:
                if slots[2 * pos - 2] == 0 or slots[2 * pos - 1] == 0:
                    sitenodes[site] = up
                    site += 1
                    continue
                sitenodes[site] = sitenodes[site + 1] = (layer, pos)
                site += 2
            graph.addedge((layer, pos), up)
    graph.addedge((nlayers - 1, 1), (nlayers - 1, 2))
    return graph, sitenodes
