"""Performance buffer of the ParetoDiscovery (DGEMO) solver — drop-in for
`autooed/mobo/solver/pareto_discovery/buffer.py` (`BufferBase` :10-183, `Buffer2D` :306-325, `Buffer3D` :328-398).

Same behaviour, array storage: every cell holds its samples as numpy arrays (x, y, distance to the origin, patch id)
sorted by distance, an insert classifies the whole batch at once (cell index, distance) and merges per touched cell
with the reference's own `np.argsort` of [old content ..., new rows in insertion order], so the kept samples and their
order — including ties — are the reference's.  `sample` consumes the global numpy RNG exactly as the reference does.

`sparse_approximation` (:203-284) labels the non-empty cells by an alpha-expansion graph cut.  The reference calls the
third-party `pygco.cut_from_graph(edges, unary, pairwise, label_cost)` (its fourth positional parameter is `n_iter`, so
`label_cost` acts as the cycle limit); pygco is neither in the reference tree nor installable here, so the cut is the
library's own host routine `aoed_graph_cut` (csrc/graphcut.cu: the same energy, the same int32 costs, expansion moves
solved exactly by max-flow, labels starting at 0, alpha in ascending order).  Where gco's tie handling differs the two can
return different labelings of equal energy class — parity against pygco itself is unpinned; when `pygco` is importable
it is used, as in the reference.
"""
import numpy as np


def graph_cut(edges, unary, pairwise, n_iter=-1, return_energy=False):
    """`pygco.cut_from_graph(edges, unary_cost, pairwise_cost, n_iter)` on the library's own alpha-expansion
    (include/aoed_b200.h `aoed_graph_cut`): int32 inputs, labels (n_node,) int32 out."""
    import ctypes

    from .. import _lib
    edges = np.ascontiguousarray(np.asarray(edges, dtype=np.int32).reshape(-1, 2))
    unary = np.ascontiguousarray(unary, dtype=np.int32)
    pairwise = np.ascontiguousarray(pairwise, dtype=np.int32)
    n_node, n_label = unary.shape
    assert pairwise.shape == (n_label, n_label)
    labels = np.zeros(n_node, dtype=np.int32)
    energy = ctypes.c_int64(0)
    _lib.check(_lib.load().aoed_graph_cut(n_node, len(edges), _lib.ptr(edges), n_label, _lib.ptr(unary), _lib.ptr(pairwise),
                                          int(n_iter), _lib.ptr(labels), ctypes.byref(energy)))
    return (labels, energy.value) if return_energy else labels


def generate_weights_batch(n_dim, delta_weight):
    """Uniform weights on the simplex by depth-first enumeration (`utils.py:137-172`), same order as the reference."""
    out = []

    def dfs(i, weight):
        if i == n_dim - 1:
            out.append(weight + [1.0 - np.sum(weight[0:i])])
            return
        w = 0.0
        while w < 1.0 + 0.5 * delta_weight and np.sum(weight[0:i]) + w < 1.0 + 0.5 * delta_weight:
            dfs(i + 1, weight[0:i] + [w])
            w += delta_weight

    dfs(0, [])
    return np.array(out)


class _Cell:
    __slots__ = ("x", "y", "dist", "pid")

    def __init__(self, d, m):
        self.x, self.y = np.empty((0, d)), np.empty((0, m))
        self.dist, self.pid = np.empty(0), np.empty(0, dtype=np.int64)

    def __len__(self):
        return len(self.dist)


class BufferBase:
    def __init__(self, cell_num, cell_size=None, origin=None, origin_constant=1e-2, delta_b=0.2, label_cost=0):
        self.cell_num = cell_num
        self.cell_size = cell_size if cell_size is not None and cell_size > 0 else None
        self.origin = origin
        self.origin_constant = origin_constant
        self.C_inf, self.delta_b, self.label_cost = 10, delta_b, label_cost
        self.cells = None  # created at the first insert (needs n_var)
        self.sample_count = 0

    # -- subclasses ------------------------------------------------------------------------------------------------
    def _find_cell_id(self, F):
        raise NotImplementedError

    def _get_graph_edges(self, valid_cells):
        raise NotImplementedError

    # -- views with the reference's attribute names (lists per cell) -------------------------------------------------
    def _ensure(self, d, m):
        if self.cells is None:
            self.cells = [_Cell(d, m) for _ in range(self.cell_num)]

    @property
    def buffer_x(self):
        return [list(c.x) for c in self.cells]

    @property
    def buffer_y(self):
        return [list(c.y) for c in self.cells]

    @property
    def buffer_dist(self):
        return [list(c.dist) for c in self.cells]

    @property
    def buffer_patch_id(self):
        return [list(c.pid) for c in self.cells]

    # -- insert / origin ---------------------------------------------------------------------------------------------
    def insert(self, X, Y, patch_ids):
        """buffer.py:69-101.  Rows beyond len(patch_ids) are classified but not stored — the reference's `zip` stops at
        its shortest argument (its first insert passes pop_size patch ids for an initial population of another size)."""
        X, Y = np.array(X, dtype=float), np.array(Y, dtype=float)
        self._ensure(X.shape[1], Y.shape[1])
        self.move_origin(np.min(Y, axis=0))
        F = Y - self.origin
        dists = np.linalg.norm(F, axis=1)
        cell_ids = self._find_cell_id(F)
        n = min(len(X), len(patch_ids))
        pid = np.asarray(patch_ids)[:n].astype(np.int64)
        order = np.argsort(cell_ids[:n], kind='stable')  # groups the rows by cell, insertion order kept inside a cell
        bounds = np.flatnonzero(np.diff(cell_ids[:n][order])) + 1
        for rows in np.split(order, bounds):
            if len(rows) == 0:
                continue
            c = self.cells[cell_ids[rows[0]]]
            c.x, c.y = np.vstack([c.x, X[rows]]), np.vstack([c.y, Y[rows]])
            c.dist, c.pid = np.concatenate([c.dist, dists[rows]]), np.concatenate([c.pid, pid[rows]])
        self.sample_count += len(X)
        for cell_id in np.unique(cell_ids):
            self._update_cell(cell_id)

    def move_origin(self, y_min):
        """buffer.py:140-161: when y_min reaches the origin in any coordinate the origin moves and every stored sample is
        re-classified, cell by cell in the old order."""
        if (y_min >= self.origin).all() and not (y_min == self.origin).any():
            return
        self.origin = np.minimum(self.origin, y_min) - self.origin_constant
        old = self.cells
        d, m = old[0].x.shape[1], old[0].y.shape[1]
        self.cells = [_Cell(d, m) for _ in range(self.cell_num)]
        self.sample_count = 0
        for c in old:
            if len(c) > 0:
                self.insert(c.x, c.y, c.pid)

    def _update_cell(self, cell_id):
        """buffer.py:163-181: sort by distance to the origin (np.argsort, default kind, as the reference), keep cell_size."""
        c = self.cells[cell_id]
        if len(c) == 0:
            return
        idx = np.argsort(c.dist)
        if self.cell_size is not None:
            self.sample_count -= max(len(idx) - self.cell_size, 0)
            idx = idx[:self.cell_size]
        c.x, c.y, c.dist, c.pid = c.x[idx], c.y[idx], c.dist[idx], c.pid[idx]

    # -- sampling ----------------------------------------------------------------------------------------------------
    def sample(self, n):
        """buffer.py:103-138: the best sample of n random cells; when n exceeds the non-empty cells, rounds over all non-empty
        cells in random order.  The reference never advances its round counter k (:121-137), so every round takes the BEST
        sample of each cell again — reproduced literally (the stored 2nd, 3rd ... samples only matter to flattened())."""
        sizes = np.array([len(c) for c in self.cells])
        nonempty = [i for i in range(self.cell_num) if sizes[i] > 0]
        if n <= len(nonempty):
            chosen = np.random.choice(nonempty, size=n, replace=False)
            return np.array([self.cells[i].x[0] for i in chosen])
        k, picked = 0, []
        while len(picked) < n:
            rows = [self.cells[i].x[k] for i in range(self.cell_num) if sizes[i] > k]
            if len(rows) == 0:  # fewer samples in the buffer than requested: repeat some
                extra = np.random.choice(np.arange(len(picked)), size=(n - len(picked)))
                picked = list(np.vstack([picked, np.array(picked)[extra]]))
                break
            picked.extend(np.random.permutation(rows))
        return np.array(picked[:n])

    def flattened(self):
        """buffer.py:286-303."""
        xs = [c.x for c in self.cells if len(c) > 0]
        ys = [c.y for c in self.cells if len(c) > 0]
        return np.concatenate(xs), np.concatenate(ys)

    # -- sparse approximation ----------------------------------------------------------------------------------------
    def sparse_approximation(self):
        """buffer.py:203-284 (see the module docstring for the graph cut)."""
        mapping = {}
        for c in self.cells:  # remap patch ids to 0 .. n_label-1 in order of first appearance (:217-226)
            for i, p in enumerate(c.pid):
                c.pid[i] = mapping.setdefault(int(p), len(mapping))
        valid_cells = np.array([i for i, c in enumerate(self.cells) if len(c) > 0])
        n_label = len(mapping)
        if len(valid_cells) > 1:
            unary = self.C_inf * np.ones((len(valid_cells), n_label))
            pairwise = -self.C_inf * np.eye(n_label)
            for i, idx in enumerate(valid_cells):
                c = self.cells[idx]
                unary[i, c.pid] = np.minimum((c.dist - c.dist.min()) / self.delta_b, self.C_inf)
            edges = self._get_graph_edges(valid_cells)
            edges, unary, pairwise = edges.astype(np.int32), unary.astype(np.int32), pairwise.astype(np.int32)
            try:
                from pygco import cut_from_graph
                labels_opt = cut_from_graph(edges, unary, pairwise, np.int32(self.label_cost))
            except ImportError:
                labels_opt = graph_cut(edges, unary, pairwise, int(self.label_cost))
        else:
            labels_opt = [int(self.cells[idx].pid[0]) for idx in valid_cells]  # one node: the unary minimiser
        approx_x, approx_y, labels = [], [], []
        for idx, label in zip(valid_cells, labels_opt):
            c = self.cells[idx]
            hit = np.flatnonzero(c.pid == label)
            j = hit[0] if len(hit) else 0  # cells are sorted by distance: the first hit is the best of that patch (:263-274)
            approx_x.append(c.x[j])
            approx_y.append(c.y[j])
            labels.append(label)
        return labels, np.array(approx_x), np.array(approx_y)


class Buffer2D(BufferBase):
    """Cells = equal angular sectors of the first quadrant around the origin (buffer.py:306-325)."""

    def __init__(self, cell_num, *args, **kwargs):
        if cell_num is None:
            cell_num = 100
        super().__init__(cell_num, *args, **kwargs)
        if self.origin is None:
            self.origin = np.zeros(2)
        self.dtheta = np.pi / 2.0 / self.cell_num

    def _find_cell_id(self, F):
        dist = np.linalg.norm(F, axis=1)
        theta = np.arccos(F[:, 0] / dist)
        return np.minimum((theta / self.dtheta).astype(int), self.cell_num - 1)

    def _get_graph_edges(self, valid_cells):
        return np.array([[i, i + 1] for i in range(len(valid_cells) - 1)])


class Buffer3D(BufferBase):
    """Cells = nearest of `cell_num` unit vectors spread over the positive octant (buffer.py:328-398)."""

    def __init__(self, cell_num, *args, origin=None, **kwargs):
        if cell_num is None:
            cell_num = 1000
        super().__init__(cell_num, *args, origin=origin, **kwargs)
        if self.origin is None:
            self.origin = np.zeros(3)
        edge_cell_num = int(np.sqrt(2 * cell_num + 0.25) + 0.5) - 1
        cell_vecs = generate_weights_batch(n_dim=3, delta_weight=1.0 / (edge_cell_num - 1))
        if len(cell_vecs) < cell_num:
            cell_vecs = np.vstack([cell_vecs, np.random.random((cell_num - len(cell_vecs), 3))])
        self.cell_vecs = cell_vecs / np.linalg.norm(cell_vecs, axis=1)[:, None]

    def _find_cell_id(self, F):
        return np.argmax(F @ self.cell_vecs.T, axis=1)

    def _get_graph_edges(self, valid_cells):
        from scipy.spatial import Delaunay
        if len(valid_cells) == 1:
            raise Exception('only 1 non-empty cell in buffer, cannot do graph cut')
        if len(valid_cells) == 2:
            return np.array([[0, 1]])
        if len(valid_cells) == 3:
            return np.array([[0, 1], [0, 2], [1, 2]])
        vertices = self.cell_vecs[valid_cells]
        same = (vertices == vertices[0]).all(axis=0)
        if same.any():
            order = np.argsort(vertices[:, np.where(np.logical_not(same))[0][0]])
            return np.ascontiguousarray(np.array([order[:-1], order[1:]]).T)
        ind, neigh = Delaunay(vertices).vertex_neighbor_vertices
        edges = [np.sort([i, j]) for i in range(len(vertices)) for j in neigh[ind[i]:ind[i + 1]]]
        return np.unique(edges, axis=0)


def get_buffer(n_obj, *args, **kwargs):
    """buffer.py:401-409."""
    if n_obj == 2:
        return Buffer2D(*args, **kwargs)
    if n_obj == 3:
        return Buffer3D(*args, **kwargs)
    raise NotImplementedError
