"""Seeded CRS test matrices (raw arrays, numpy) shared by the CPU and GPU tests.

Every builder returns (n_rows, n_cols, values, columns, offset_rows) with offsets in the index type,
like `offset_rows: Vec<I>` (sparsemat_crs.rs:14).  Columns inside a row are deliberately unsorted and
may repeat where noted: the reference's mvp sums in storage order (sparsematrix.rs:146-158) and
`to_crs` never sorts (sparsemat_crs.rs:24-36)."""
import numpy as np


def from_row_lengths(rng, lens, n_cols, vdt, idt, band=None):
    lens = np.asarray(lens, np.int64)
    n_rows = lens.size
    offs = np.zeros(n_rows + 1, np.int64)
    np.cumsum(lens, out=offs[1:])
    nnz = int(offs[-1])
    if band is None:
        cols = rng.integers(0, n_cols, nnz)
    else:
        rows = np.repeat(np.arange(n_rows), lens)
        cols = np.clip(rows * n_cols // max(n_rows, 1) + rng.integers(-band, band + 1, nnz), 0, n_cols - 1)
    vals = rng.uniform(-1.0, 1.0, nnz).astype(vdt)
    return n_rows, n_cols, vals, cols.astype(idt), offs.astype(idt)


def ragged(seed, n_rows, n_cols, max_len, vdt, idt, empty_frac=0.2):
    """Row lengths uniform in [0, max_len], a fraction forced empty (incl. first and last row)."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(0, max_len + 1, n_rows)
    lens[rng.random(n_rows) < empty_frac] = 0
    if n_rows:
        lens[0] = 0
        lens[-1] = 0
    return from_row_lengths(rng, lens, n_cols, vdt, idt)


def powerlaw(seed, n_rows, n_cols, max_len, vdt, idt):
    rng = np.random.default_rng(seed)
    u = rng.random(n_rows)
    lens = np.clip(np.floor(8.0 / np.sqrt(u)), 1, max_len).astype(np.int64)
    return from_row_lengths(rng, lens, n_cols, vdt, idt)


def giant_row(seed, n_rows, n_cols, giant_len, vdt, idt, where=None):
    """Short rows plus one row far longer than any staging buffer."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(1, 9, n_rows)
    lens[n_rows // 2 if where is None else where] = giant_len
    return from_row_lengths(rng, lens, n_cols, vdt, idt)


def banded(seed, n, band, per_row, vdt, idt):
    rng = np.random.default_rng(seed)
    lens = np.full(n, per_row, np.int64)
    return from_row_lengths(rng, lens, n, vdt, idt, band=band)


def all_empty(n_rows, n_cols, vdt, idt):
    return n_rows, n_cols, np.zeros(0, vdt), np.zeros(0, idt), np.zeros(n_rows + 1, idt)


def abs_rowsum(values, columns, offsets, x):
    """(|A| |x|)_i in f64: the scale the reordered-reduction tolerance is relative to."""
    prod = np.abs(values.astype(np.float64)) * np.abs(x.astype(np.float64)[columns.astype(np.int64)])
    cs = np.concatenate([[0.0], np.cumsum(prod)])
    o = offsets.astype(np.int64)
    return cs[o[1:]] - cs[o[:-1]]


TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12}   # BASELINE.json north_star


def fem_like(nx, ny, nz, dof, vdt, idt, seed=0):
    """27-point stencil with `dof` unknowns per grid node (a trilinear FEM vector problem): 27 * dof entries per interior
    row, fewer at the boundary; columns ascending inside a row; values random (not symmetric: only the product is tested)."""
    rng = np.random.default_rng(seed)
    n_nodes = nx * ny * nz
    node = np.arange(n_nodes)
    ix, iy, iz = node % nx, (node // nx) % ny, node // (nx * ny)
    parts = []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                ok = (ix + dx >= 0) & (ix + dx < nx) & (iy + dy >= 0) & (iy + dy < ny) & (iz + dz >= 0) & (iz + dz < nz)
                parts.append(np.where(ok, node + dx + dy * nx + dz * nx * ny, -1))
    nbs = np.stack(parts, 1)                                       # [n_nodes, 27], -1 = outside, ascending where valid
    valid = nbs >= 0
    per_node = valid.sum(1)
    flat_nb = nbs[valid]                                           # per node, its valid neighbours in ascending order
    # the columns of ONE row of each node: neighbour * dof + 0..dof-1
    nb_exp = np.repeat(flat_nb, dof) * dof + np.tile(np.arange(dof), flat_nb.size)
    exp_off = np.zeros(n_nodes + 1, np.int64)
    np.cumsum(per_node * dof, out=exp_off[1:])
    n_rows = n_nodes * dof
    lens = np.repeat(per_node * dof, dof)                          # every unknown of a node has the node's column set
    offs = np.zeros(n_rows + 1, np.int64)
    np.cumsum(lens, out=offs[1:])
    node_of_row = np.repeat(node, dof)
    idx = np.arange(int(offs[-1])) - np.repeat(offs[:-1], lens) + np.repeat(exp_off[node_of_row], lens)
    cols = nb_exp[idx]
    vals = rng.uniform(-1.0, 1.0, cols.size).astype(vdt)
    return n_rows, n_rows, vals, cols.astype(idt), offs.astype(idt)
