// Multi-label graph cut by alpha-expansion — host code (the graphs of ParetoDiscovery's sparse approximation have one node
// per non-empty buffer cell: hundreds to a few thousand nodes, a chain or a triangulated grid of edges; a sequential
// max-flow on them takes microseconds to milliseconds and there is nothing for the GPU in it).
//
// Replaces the third-party `pygco.cut_from_graph(edges, unary_cost, pairwise_cost, n_iter)` call of
// autooed/mobo/solver/pareto_discovery/buffer.py:249 (a Cython wrapper of gco-v3.0, absent from the reference tree and
// from this image).  Energy, as there:
//     E(l) = sum_i unary[i][l_i] + sum_{(i,j) in edges} pairwise[l_i][l_j]          (int32 inputs, int64 accumulation)
// Algorithm (Boykov, Veksler, Zabih 2001; graph construction of Kolmogorov & Zabih 2004 for each binary move):
//   labels start at 0; one cycle visits alpha = 0 .. n_label-1 in order; an expansion move lets every node either keep its
//   label (x = 0) or take alpha (x = 1) and is the exact minimum of that binary problem, found as a minimum s-t cut;
//   cycles repeat until one brings no decrease or n_iter cycles are done (n_iter < 0: until convergence, 0: no move).
//   Every pair term must be submodular for the move:  V(fi,fj) + V(a,a) <= V(fi,a) + V(a,fj)  (true for the reference's
//   V = -C_inf * I); a violating pair is reported as AOED_ERR_UNSUPPORTED instead of being truncated.
//   Of the minimum cuts the one with the SMALLEST sink side is taken (only the nodes that can still reach the sink in the
//   residual graph take alpha): that set is unique — any correct max-flow gives the same result — and it is empty whenever
//   keeping every label is already optimal, so a tie never moves a node and every accepted move lowers the energy strictly.
//   Two exact shortcuts keep the number of max-flows small: a label is not tried again while the labeling has not changed
//   since its last fruitless attempt, and not at all when an upper bound on its gain (term by term) is zero.
// Max-flow: Dinic's algorithm with BFS levels and iterative DFS on a CSR-like arc list.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/aoed_b200.h"

extern "C" int aoed_detail_fail(int code, const char* msg);  // gp.cu: the shared last-error slot of aoed_last_error

namespace {

struct FlowGraph {
    struct Arc {
        int to, next;
        int64_t cap;
    };
    std::vector<int> head, level, it, queue, path;  // (queue / path: scratch of bfs / augment, kept to avoid re-allocation)
    std::vector<Arc> arcs;
    int s, t;

    explicit FlowGraph(int n_inner) : head(n_inner + 2, -1), level(n_inner + 2), it(n_inner + 2), s(n_inner), t(n_inner + 1) {
        queue.reserve(n_inner + 2);
    }
    void reset() {  // same nodes, no arcs: one object serves every move
        std::fill(head.begin(), head.end(), -1);
        arcs.clear();
    }

    void add(int u, int v, int64_t c_uv, int64_t c_vu) {
        arcs.push_back({v, head[u], c_uv});
        head[u] = int(arcs.size()) - 1;
        arcs.push_back({u, head[v], c_vu});
        head[v] = int(arcs.size()) - 1;
    }
    bool bfs() {
        std::fill(level.begin(), level.end(), -1);
        queue.clear();
        queue.push_back(s);
        level[s] = 0;
        for (size_t qi = 0; qi < queue.size(); ++qi) {
            const int u = queue[qi];
            for (int a = head[u]; a >= 0; a = arcs[a].next)
                if (arcs[a].cap > 0 && level[arcs[a].to] < 0) {
                    level[arcs[a].to] = level[u] + 1;
                    queue.push_back(arcs[a].to);
                }
        }
        return level[t] >= 0;
    }
    // one augmenting path along the level graph, iteratively; returns the flow pushed (0: none left)
    int64_t augment() {
        path.clear();  // arcs from s
        int u = s;
        for (;;) {
            if (u == t) {
                int64_t f = INT64_MAX;
                for (int a : path) f = std::min(f, arcs[a].cap);
                for (int a : path) {
                    arcs[a].cap -= f;
                    arcs[a ^ 1].cap += f;
                }
                return f;
            }
            int& a = it[u];
            while (a >= 0 && !(arcs[a].cap > 0 && level[arcs[a].to] == level[u] + 1)) a = arcs[a].next;
            if (a >= 0) {
                path.push_back(a);
                u = arcs[a].to;
            } else {
                level[u] = -1;  // dead end: never enter again in this phase
                if (path.empty()) return 0;
                u = arcs[path.back() ^ 1].to;
                path.pop_back();
            }
        }
    }
    int64_t max_flow() {
        int64_t flow = 0;
        while (bfs()) {
            it = head;
            while (int64_t f = augment()) flow += f;
        }
        return flow;
    }
    // after max_flow(): marks (in level[], 1 / 0) the nodes that can reach t through arcs with residual capacity —
    // the sink side of the minimum cut whose sink side is smallest
    void mark_sink_side() {
        std::fill(level.begin(), level.end(), 0);
        queue.clear();
        queue.push_back(t);
        level[t] = 1;
        for (size_t qi = 0; qi < queue.size(); ++qi) {
            const int v = queue[qi];
            for (int b = head[v]; b >= 0; b = arcs[b].next) {
                const int u = arcs[b].to;  // arc u -> v is the partner of b
                if (!level[u] && arcs[b ^ 1].cap > 0) {
                    level[u] = 1;
                    queue.push_back(u);
                }
            }
        }
    }
    bool sink_side(int u) const { return level[u] != 0; }
};

int64_t energy_of(int n_node, int64_t n_edge, const int32_t* edges, int n_label, const int32_t* unary, const int32_t* pairwise,
                  const int32_t* labels) {
    int64_t e = 0;
    for (int i = 0; i < n_node; ++i) e += unary[int64_t(i) * n_label + labels[i]];
    for (int64_t k = 0; k < n_edge; ++k) e += pairwise[int64_t(labels[edges[2 * k]]) * n_label + labels[edges[2 * k + 1]]];
    return e;
}

}  // namespace

extern "C" int aoed_graph_cut(int n_node, int64_t n_edge, const int32_t* edges, int n_label, const int32_t* unary,
                              const int32_t* pairwise, int n_iter, int32_t* labels, int64_t* energy) {
    auto fail = [](int code, const char* msg) { return aoed_detail_fail(code, msg); };
    if (n_node < 0 || n_edge < 0 || n_label < 1) return fail(AOED_ERR_INVALID, "n_node, n_edge >= 0 and n_label >= 1 required");
    if ((n_node > 0 && (!unary || !labels)) || (n_edge > 0 && !edges) || !pairwise) return fail(AOED_ERR_INVALID, "null argument");
    for (int64_t k = 0; k < n_edge; ++k) {
        const int i = edges[2 * k], j = edges[2 * k + 1];
        if (i < 0 || j < 0 || i >= n_node || j >= n_node || i == j) return fail(AOED_ERR_INVALID, "edge endpoint out of range");
    }
    for (int i = 0; i < n_node; ++i) labels[i] = 0;
    int64_t e_cur = energy_of(n_node, n_edge, edges, n_label, unary, pairwise, labels);
    auto V = [&](int a, int b) -> int64_t { return pairwise[int64_t(a) * n_label + b]; };
    std::vector<int64_t> c0(n_node), c1(n_node);  // unary cost of x = 0 (keep) and x = 1 (take alpha), pair terms folded in
    std::vector<long long> tried_at(n_label, -1);  // labeling version at the label's last fruitless attempt
    long long version = 0;                         // bumped by every accepted move
    FlowGraph g(n_node);
    for (int cycle = 0; n_iter < 0 || cycle < n_iter; ++cycle) {
        const int64_t e_before = e_cur;
        for (int alpha = 0; alpha < n_label; ++alpha) {
            if (tried_at[alpha] == version) continue;  // same labeling as last time: the same (empty) move
            tried_at[alpha] = version;
            // upper bound on what the move can gain, term by term
            int64_t gain = 0;
            bool any_other = false;
            for (int i = 0; i < n_node; ++i) {
                c0[i] = unary[int64_t(i) * n_label + labels[i]];
                c1[i] = unary[int64_t(i) * n_label + alpha];
                gain += std::max<int64_t>(0, c0[i] - c1[i]);
                any_other = any_other || labels[i] != alpha;
            }
            if (!any_other) continue;
            for (int64_t k = 0; k < n_edge && gain == 0; ++k) {
                const int lp = labels[edges[2 * k]], lq = labels[edges[2 * k + 1]];
                gain += std::max<int64_t>(0, V(lp, lq) - std::min(V(alpha, alpha), std::min(V(lp, alpha), V(alpha, lq))));
            }
            if (gain == 0) continue;
            g.reset();
            for (int64_t k = 0; k < n_edge; ++k) {
                const int p = edges[2 * k], q = edges[2 * k + 1];
                // E(xp, xq): A = (0,0), B = (0,1), C = (1,0), D = (1,1)
                //          = A + (C - A) xp + (D - C) xq + (B + C - A - D) (1 - xp) xq
                const int64_t A = V(labels[p], labels[q]), B = V(labels[p], alpha), C = V(alpha, labels[q]), D = V(alpha, alpha);
                const int64_t w = B + C - A - D;
                if (w < 0) return fail(AOED_ERR_UNSUPPORTED, "pairwise cost is not submodular for an expansion move");
                c0[p] += A;  // constant A booked on xp = 0 and xp = 1 alike ...
                c1[p] += C;  // ... so that xp = 1 costs A + (C - A)
                c1[q] += D - C;
                if (w > 0) g.add(p, q, w, 0);  // cut when p keeps (source side) and q takes alpha (sink side)
            }
            for (int i = 0; i < n_node; ++i) {
                const int64_t lo = std::min(c0[i], c1[i]);
                if (c1[i] - lo > 0) g.add(g.s, i, c1[i] - lo, 0);  // cut when i is on the sink side: x = 1
                if (c0[i] - lo > 0) g.add(i, g.t, c0[i] - lo, 0);  // cut when i is on the source side: x = 0
            }
            g.max_flow();
            g.mark_sink_side();
            bool moved = false;
            for (int i = 0; i < n_node; ++i)
                if (g.sink_side(i) && labels[i] != alpha) {
                    labels[i] = alpha;
                    moved = true;
                }
            if (moved) {
                ++version;
                e_cur = energy_of(n_node, n_edge, edges, n_label, unary, pairwise, labels);
            }
        }
        if (e_cur >= e_before) break;
    }
    if (energy) *energy = e_cur;
    return AOED_OK;
}
