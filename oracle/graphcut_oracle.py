"""TEST INFRASTRUCTURE (never imported by the product package): CPU restatement of the multi-label graph cut behind
`Buffer.sparse_approximation` (autooed/mobo/solver/pareto_discovery/buffer.py:203-284).

The reference calls `pygco.cut_from_graph(edges, unary_cost, pairwise_cost, label_cost)` (buffer.py:249; the fourth positional
argument of that function is `n_iter`, so the reference's `label_cost` is in fact the cycle limit).  pygco — a Cython wrapper
of gco-v3.0 — is a third-party dependency that is neither in /root/reference nor installable here: PARITY UNPINNED against the
library itself.  What is restated is the published algorithm (alpha-expansion, Boykov / Veksler / Zabih 2001, with the binary
moves solved exactly as minimum cuts, Kolmogorov & Zabih 2004), and it is anchored three ways:
  * `expansion`           the algorithm on networkx's `edmonds_karp` (an independent max-flow; the nodes that can reach the sink
                          in its residual network are the same unique minimal sink side the product's routine takes), same
                          label order and stopping rule, no shortcuts;
  * `brute_force`         the global optimum by enumeration (small graphs): a two-label expansion from label 0 must reach it;
  * `is_expansion_optimal` no single expansion move (enumerated) improves the result.
"""
import itertools

import numpy as np


def energy(edges, unary, pairwise, labels):
    labels = np.asarray(labels)
    e = int(unary[np.arange(len(labels)), labels].astype(np.int64).sum())
    if len(edges):
        e += int(pairwise[labels[edges[:, 0]], labels[edges[:, 1]]].astype(np.int64).sum())
    return e


def _move(edges, unary, pairwise, labels, alpha):
    import networkx as nx
    n = len(labels)
    c0 = unary[np.arange(n), labels].astype(np.int64)
    c1 = unary[:, alpha].astype(np.int64).copy()
    g = nx.DiGraph()
    g.add_nodes_from(range(n))
    g.add_node('s')
    g.add_node('t')

    def arc(u, v, c):
        if g.has_edge(u, v):
            g[u][v]['capacity'] += c
        else:
            g.add_edge(u, v, capacity=c)

    for p, q in edges:
        A, B = int(pairwise[labels[p], labels[q]]), int(pairwise[labels[p], alpha])
        C, D = int(pairwise[alpha, labels[q]]), int(pairwise[alpha, alpha])
        w = B + C - A - D
        assert w >= 0, "not submodular"
        c0[p] += A
        c1[p] += C
        c1[q] += D - C
        if w > 0:
            arc(int(p), int(q), w)
    for i in range(n):
        lo = min(c0[i], c1[i])
        if c1[i] > lo:
            arc('s', i, int(c1[i] - lo))
        if c0[i] > lo:
            arc(i, 't', int(c0[i] - lo))
    # the SMALLEST sink side of a minimum cut = the nodes that can still reach t in the residual network of any maximum flow
    # (what nx.minimum_cut returns as its second set); it is empty when keeping every label is optimal
    R = nx.algorithms.flow.edmonds_karp(g, 's', 't')
    sink, stack = {'t'}, ['t']
    while stack:
        v = stack.pop()
        for u in R.predecessors(v):
            if u not in sink and R[u][v]['capacity'] - R[u][v]['flow'] > 0:
                sink.add(u)
                stack.append(u)
    out = labels.copy()
    for i in range(n):
        if i in sink:
            out[i] = alpha
    return out


def expansion(edges, unary, pairwise, n_iter=-1):
    edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    n, L = unary.shape
    labels = np.zeros(n, dtype=np.int64)
    e_cur = energy(edges, unary, pairwise, labels)
    cycle = 0
    while n_iter < 0 or cycle < n_iter:
        e_before = e_cur
        for alpha in range(L):
            if np.all(labels == alpha):
                continue
            labels = _move(edges, unary, pairwise, labels, alpha)
            e_cur = energy(edges, unary, pairwise, labels)
        cycle += 1
        if e_cur >= e_before:
            break
    return labels, e_cur


def brute_force(edges, unary, pairwise):
    n, L = unary.shape
    best = None
    for lab in itertools.product(range(L), repeat=n):
        e = energy(edges, unary, pairwise, np.array(lab))
        if best is None or e < best[0]:
            best = (e, np.array(lab))
    return best


def is_expansion_optimal(edges, unary, pairwise, labels):
    """True when no expansion move (every alpha, every subset of nodes switching to alpha) lowers the energy."""
    n, L = unary.shape
    e0 = energy(edges, unary, pairwise, labels)
    for alpha in range(L):
        for mask in range(1, 1 << n):
            lab = np.array(labels).copy()
            for i in range(n):
                if mask >> i & 1:
                    lab[i] = alpha
            if energy(edges, unary, pairwise, lab) < e0:
                return False
    return True
