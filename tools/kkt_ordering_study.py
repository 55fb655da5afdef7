#!/usr/bin/env python
"""Design study (not product code): fill-in and dependency depth of the LDL' factor of the
OSQP KKT matrix of the coupled / decoupled tracking QP under different elimination orderings.

The QP pattern follows the construction order of the reference
(src/coupled_lat_long.jl:233-292, src/decoupled_lat_long.jl:162-211).
Used to choose the static ordering baked into the CUDA ADMM kernel.
"""
import sys
import numpy as np


def coupled_pattern(Ns, Nl):
    N = 1 + Ns + Nl
    T = N - 1
    q = lambda i, t: 6 * t + i
    u = lambda i, t: 6 * N + 2 * t + i
    sg = lambda i, t: 8 * N + 2 * t + i
    sh = lambda t: 8 * N + 2 * T + t
    dd = lambda t: 8 * N + 2 * T + Ns + t
    df = lambda t: 8 * N + 2 * T + Ns + T + t
    n = 8 * N + 2 * T + Ns + 2 * T
    rows = []  # list of (list of cols, stage)
    for t in range(T):
        for i in range(2):
            rows.append(([sg(i, t)], t))
    for t in range(Ns):
        rows.append(([sh(t)], t))
    for t in range(T):
        rows.append(([u(0, t), u(0, t + 1), dd(t)], t))
    for t in range(T):
        rows.append(([u(1, t), u(1, t + 1), df(t)], t))
    for t in range(N):
        rows.append(([q(1, t)], t))
    for t in range(N):
        rows.append(([q(1, t)], t))
    for t in range(N):
        rows.append(([u(1, t)], t))
    for i in range(6):
        rows.append(([q(i, 0)], 0))
    for i in range(2):
        rows.append(([u(i, 0)], 0))
    for t in range(Ns):
        for i in range(6):
            rows.append(([q(j, t) for j in range(6)] + [u(0, t), u(1, t)] + [q(i, t + 1)], t))
    for t in range(Ns):
        rows.append(([u(0, t), u(1, t), sh(t)], t))
    for t in range(Ns, T):
        for i in range(6):
            rows.append(([q(j, t) for j in range(6)] + [u(0, t), u(1, t), u(0, t + 1), u(1, t + 1)] + [q(i, t + 1)], t))
    for t in range(T):
        rows.append(([u(0, t + 1)], t))
        rows.append(([u(0, t + 1)], t))
        rows.append(([u(1, t + 1)], t))
        for k in range(4):
            rows.append(([q(2, t + 1), q(3, t + 1), sg(k // 2, t)], t))
        rows.append(([dd(t)], t))
        rows.append(([dd(t)], t))
    # stage of each variable (for nested dissection)
    vstage = np.zeros(n, int)
    for t in range(N):
        for i in range(6):
            vstage[q(i, t)] = t
        for i in range(2):
            vstage[u(i, t)] = t
    for t in range(T):
        for i in range(2):
            vstage[sg(i, t)] = t + 1
        vstage[dd(t)] = t
        vstage[df(t)] = t
    for t in range(Ns):
        vstage[sh(t)] = t
    return n, rows, vstage


def kkt_adj(n, rows):
    m = len(rows)
    adj = [set() for _ in range(n + m)]
    for r, (cols, _) in enumerate(rows):
        for c in cols:
            adj[n + r].add(c)
            adj[c].add(n + r)
    return adj


def symbolic(adj, perm):
    """Return column structures of L (set per column, in permuted indices) via elimination game."""
    nn = len(adj)
    pos = np.empty(nn, int)
    pos[perm] = np.arange(nn)
    # adjacency in permuted index space, only higher-numbered neighbours
    struct = [set(pos[j] for j in adj[perm[k]] if pos[j] > k) for k in range(nn)]
    parent = -np.ones(nn, int)
    for k in range(nn):
        s = struct[k]
        if s:
            p = min(s)
            parent[k] = p
            struct[p] |= (s - {p})
    return struct, parent


def min_degree(adj):
    nn = len(adj)
    g = [set(a) for a in adj]
    alive = np.ones(nn, bool)
    perm = []
    import heapq
    deg = [len(a) for a in g]
    heap = [(deg[i], i) for i in range(nn)]
    heapq.heapify(heap)
    while heap:
        d, v = heapq.heappop(heap)
        if not alive[v] or d != len(g[v]):
            continue
        alive[v] = False
        perm.append(v)
        nb = list(g[v])
        for a in nb:
            g[a].discard(v)
        for i, a in enumerate(nb):
            for b in nb[i + 1:]:
                if b not in g[a]:
                    g[a].add(b)
                    g[b].add(a)
        for a in nb:
            heapq.heappush(heap, (len(g[a]), a))
    return np.array(perm)


def stats(name, adj, perm):
    struct, parent = symbolic(adj, perm)
    nn = len(adj)
    nnzL = sum(len(s) for s in struct)
    # solve depth: level[k] = 1 + max(level[j] for j<k with L[k,j] != 0)
    level = np.zeros(nn, int)
    for k in range(nn):
        for i in struct[k]:
            level[i] = max(level[i], level[k] + 1)
    flops = sum(len(s) * (len(s) + 1) // 2 for s in struct)
    # etree height
    h = np.zeros(nn, int)
    for k in range(nn):
        if parent[k] >= 0:
            h[parent[k]] = max(h[parent[k]], h[k] + 1)
    widths = np.bincount(level)
    print(f"{name:28s} nnz(L)={nnzL:6d} levels={level.max()+1:4d} etree_h={h.max()+1:4d} "
          f"factor_updates={flops:7d} maxcol={max(len(s) for s in struct):3d} "
          f"level widths min/med/max={widths.min()}/{int(np.median(widths))}/{widths.max()}")
    return struct, level


def nested_dissection_perm(n, rows, vstage, N, leaf_md=True):
    """Order: recursively bisect the stage axis; separators = the q/u variables of the middle stage
    (all coupling between stages goes through q_t,u_t of neighbouring stages via rows attached to stage t)."""
    m = len(rows)
    adj = kkt_adj(n, rows)
    # node stage: variable stage or row stage
    nstage = np.concatenate([vstage, np.array([r[1] for r in rows])])
    return adj, nstage


if __name__ == "__main__":
    Ns, Nl = (10, 20) if len(sys.argv) < 3 else (int(sys.argv[1]), int(sys.argv[2]))
    n, rows, vstage = coupled_pattern(Ns, Nl)
    m = len(rows)
    print("n", n, "m", m, "nnzA", sum(len(r[0]) for r in rows))
    adj = kkt_adj(n, rows)
    stats("natural [x; rows]", adj, np.arange(n + m))
    pmd = min_degree(adj)
    stats("min degree", adj, pmd)


def constrained_min_degree(adj, cls):
    """min-degree restricted to eliminating classes in increasing order."""
    nn = len(adj)
    g = [set(a) for a in adj]
    alive = np.ones(nn, bool)
    perm = []
    import heapq
    for c in sorted(set(cls)):
        heap = [(len(g[i]), i) for i in range(nn) if cls[i] == c and alive[i]]
        heapq.heapify(heap)
        while heap:
            d, v = heapq.heappop(heap)
            if not alive[v] or d != len(g[v]):
                continue
            alive[v] = False
            perm.append(v)
            nb = list(g[v])
            for a in nb:
                g[a].discard(v)
            for i, a in enumerate(nb):
                for b in nb[i + 1:]:
                    if b not in g[a]:
                        g[a].add(b)
                        g[b].add(a)
            for a in nb:
                if cls[a] == c:
                    heapq.heappush(heap, (len(g[a]), a))
    return np.array(perm)


def nd_classes(n, rows, N, leaf):
    m = len(rows)
    cls = np.zeros(n + m, int)
    sep_depth = {}

    def bis(lo, hi, d):
        if hi - lo + 1 <= leaf:
            return
        mid = (lo + hi) // 2
        sep_depth[mid] = d
        bis(lo, mid - 1, d + 1)
        bis(mid + 1, hi, d + 1)
    bis(0, N - 1, 0)
    if sep_depth:
        md = max(sep_depth.values())
        for t, d in sep_depth.items():
            for i in range(6):
                cls[6 * t + i] = md - d + 1
            for i in range(2):
                cls[6 * N + 2 * t + i] = md - d + 1
    return cls


if __name__ == "__main__":
    N = 1 + Ns + Nl
    for leaf in (1, 2, 3, 4, 7):
        cls = nd_classes(n, rows, N, leaf)
        p = constrained_min_degree(adj, cls)
        stats(f"ND leaf={leaf} nsep={int((cls>0).sum()//8)}", adj, p)
