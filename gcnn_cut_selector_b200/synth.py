"""Synthetic tripartite graph samples in the reference's on-disk sample format.

The reference obtains samples from SCIP (``data_collector.py:135-140``: ``{'data': [state, improvements]}`` with
``state`` the 5-tuple of dicts ``utils.get_state`` returns, ``utils.py:236-238``).  SCIP is not available, so the
benchmark shapes of BASELINE.md section 4 are synthesised here with the same structure: row-major sorted COO edge
indices (``utils.py:102-104`` emits csr -> vstack -> coo), 4 / 14 / 6 node features (``utils.py:60-213``) and one
edge feature.  Structure statistics follow the reference's generators (``instance_generator.py:313-381`` setcov,
``:384-566`` combauc, ``:569-647`` capfac, ``:650-697`` indset); cut counts are a synthetic choice.

Host-side numpy only; nothing here is on the timed path.
"""
from __future__ import annotations

import numpy as np

CONS_F, VAR_F, CUT_F = 4, 14, 6

# name -> (n_cons, n_vars, n_cuts, nnz per cut)
SHAPES = {
    "tiny": (7, 11, 4, 3),
    "mini": (60, 100, 8, 12),
    "setcov": (500, 1000, 64, 100),
    "combauc": (193, 500, 64, 30),
    "capfac": (10201, 10100, 128, 40),
    "indset": (1950, 500, 64, 10),
    "miplib": (100_000, 100_000, 5_000, 100),
    # test shape for the edge kernels' work decomposition: a block of heavy rows (120-390 edges, reduced by a whole CTA)
    # behind rows of degree 0-6, hub columns (the transposed layout has long rows too), cut rows of 150 non-zeros
    "skewed": (300, 400, 16, 150),
}


def _distinct_targets(owner: np.ndarray, n_targets: int, rng: np.random.Generator) -> np.ndarray:
    """For every entry of ``owner`` draw a target in [0, n_targets) such that (owner, target) pairs are distinct."""
    tgt = rng.integers(0, n_targets, size=owner.shape[0])
    for _ in range(64):
        key = owner.astype(np.int64) * n_targets + tgt
        order = np.argsort(key, kind="stable")
        dup = np.zeros(key.shape[0], dtype=bool)
        dup[order[1:]] = key[order[1:]] == key[order[:-1]]
        if not dup.any():
            return tgt
        tgt[dup] = rng.integers(0, n_targets, size=int(dup.sum()))
    raise RuntimeError("could not de-duplicate edges (degree too close to the number of targets)")


def _edges_from_row_degrees(deg: np.ndarray, n_cols: int, rng) -> np.ndarray:
    rows = np.repeat(np.arange(deg.shape[0]), deg)
    cols = _distinct_targets(rows, n_cols, rng)
    order = np.lexsort((cols, rows))  # row-major, columns ascending inside a row (csr -> coo order)
    return np.vstack([rows[order], cols[order]]).astype(np.int64)


def _fit_total(deg: np.ndarray, total: int, lo: int, hi: int, rng) -> np.ndarray:
    deg = np.clip(deg, lo, hi).astype(np.int64)
    while deg.sum() != total:
        diff = int(total - deg.sum())
        step = 1 if diff > 0 else -1
        ok = np.flatnonzero((deg < hi) if step > 0 else (deg > lo))
        pick = rng.choice(ok, size=min(abs(diff), ok.shape[0]), replace=False)
        deg[pick] += step
    return deg


def _structure(shape: str, rng) -> tuple[np.ndarray, int, int]:
    n_cons, n_vars, _, _ = SHAPES[shape]
    if shape == "setcov":
        # instance_generator.py:334-359: 25,000 non-zeros, every column at least two, every row at least one.
        nnz = int(n_cons * n_vars * 0.05)
        col_deg = 2 + rng.multinomial(nnz - 2 * n_vars, np.full(n_vars, 1.0 / n_vars))
        cols = np.repeat(np.arange(n_vars), col_deg)
        rows = _distinct_targets(cols, n_cons, rng)
        empty = np.setdiff1d(np.arange(n_cons), rows)
        for r in empty:  # move one entry of a well-covered row onto each empty row
            counts = np.bincount(rows, minlength=n_cons)
            victim = np.flatnonzero(rows == np.argmax(counts))[0]
            rows[victim] = r
        order = np.lexsort((cols, rows))
        ei = np.vstack([rows[order], cols[order]]).astype(np.int64)
    elif shape == "capfac":
        # instance_generator.py:627-639 row pattern for 100 customers x 100 facilities.
        nc = nf = 100
        x = lambda i, j: i * nf + j
        y = lambda j: nc * nf + j
        rows, cols, r = [], [], 0
        for i in range(nc):  # demand rows, degree 100
            rows += [r] * nf; cols += [x(i, j) for j in range(nf)]; r += 1
        for j in range(nf):  # capacity rows, degree 101
            rows += [r] * (nc + 1); cols += [x(i, j) for i in range(nc)] + [y(j)]; r += 1
        rows += [r] * nf; cols += [y(j) for j in range(nf)]; r += 1  # total capacity row
        for i in range(nc):  # tightening rows, degree 2
            for j in range(nf):
                rows += [r, r]; cols += [x(i, j), y(j)]; r += 1
        rows, cols = np.asarray(rows), np.asarray(cols)
        order = np.lexsort((cols, rows))
        ei = np.vstack([rows[order], cols[order]]).astype(np.int64)
    elif shape == "combauc":
        deg = _fit_total(np.round(rng.lognormal(2.3, 0.8, n_cons)), 2800, 1, 60, rng)
        ei = _edges_from_row_degrees(deg, n_vars, rng)
    elif shape == "indset":
        deg = _fit_total(2 + (rng.random(n_cons) < 0.004) * rng.integers(1, 4, n_cons), 3916, 2, 5, rng)
        ei = _edges_from_row_degrees(deg, n_vars, rng)
    elif shape == "miplib":
        heavy = rng.random(n_cons) < 0.01
        deg = np.where(heavy, np.exp(rng.uniform(np.log(100), np.log(1000), n_cons)), 1 + rng.poisson(5.0, n_cons))
        deg = _fit_total(np.round(deg), 1_000_000, 1, 1000, rng)
        ei = _edges_from_row_degrees(deg, n_vars, rng)
    elif shape == "skewed":
        deg = rng.integers(0, 7, n_cons)          # includes constraints without any edge
        deg[-8:] = rng.integers(120, 391, 8)      # heavy rows next to each other
        rows = np.repeat(np.arange(n_cons), deg)
        hub = rng.random(rows.shape[0]) < np.where(rows >= n_cons - 8, 0.02, 0.6)  # most light entries hit five hub columns
        cols = np.where(hub, rng.integers(0, 5, rows.shape[0]), rng.integers(5, n_vars, rows.shape[0]))
        key = np.unique(rows.astype(np.int64) * n_vars + cols)  # distinct (row, column) pairs, row-major sorted
        ei = np.vstack([key // n_vars, key % n_vars]).astype(np.int64)
    else:  # tiny / mini: a few edges per row, some isolated variables
        deg = rng.integers(1, max(2, min(n_vars, 6)), n_cons)
        ei = _edges_from_row_degrees(deg, n_vars, rng)
    return ei, n_cons, n_vars


def make_sample(shape: str, seed: int, structure=None):
    """One sample ``((cons, cons_edge, var, cut, cut_edge), improvements)`` of the named shape (fp64 / int64 like the
    pickled reference samples; ``load_batch`` casts, utils.py:413-423)."""
    rng = np.random.default_rng(seed)
    _, _, n_cuts, cut_nnz = SHAPES[shape]
    ei, n_cons, n_vars = structure if structure is not None else _structure(shape, rng)

    row_deg = np.maximum(np.bincount(ei[0], minlength=n_cons), 1)
    if shape in ("setcov", "indset"):
        coef = 1.0 / np.sqrt(row_deg[ei[0]])  # unit coefficients divided by the row norm (utils.py:98)
    else:
        coef = rng.standard_normal(ei.shape[1])
    cons = rng.standard_normal((n_cons, CONS_F))
    cons[:, 1] = rng.random(n_cons) < 0.3  # is_tight
    var = rng.standard_normal((n_vars, VAR_F))
    var[:, :4] = np.eye(4)[rng.choice(4, n_vars, p=[0.7, 0.1, 0.0, 0.2])]  # type one-hot (utils.py:118-121)
    var[:, 5:9] = rng.random((n_vars, 4)) < 0.5  # has_lb, has_ub, at_lb, at_ub
    var[:, 9] = rng.random(n_vars) * 0.5  # frac
    cut = rng.standard_normal((n_cuts, CUT_F))
    cut[:, 1:3] = rng.random((n_cuts, 2))  # support, integral support

    cut_nnz = min(cut_nnz, n_vars)
    cut_ei = _edges_from_row_degrees(np.full(n_cuts, cut_nnz), n_vars, rng)
    cut_coef = rng.standard_normal(cut_ei.shape[1]) / np.sqrt(cut_nnz)
    improvements = rng.uniform(0.0, 0.1, n_cuts)

    state = ({"values": cons},
             {"indices": ei, "values": coef.reshape(-1, 1)},
             {"values": var},
             {"values": cut},
             {"indices": cut_ei, "values": cut_coef.reshape(-1, 1)})
    return state, improvements


def make_samples(shape: str, n: int, seed0: int = 0, n_structures: int | None = None):
    """``n`` samples; graph ``g`` uses seed ``seed0 + g``.  ``n_structures`` bounds the number of distinct edge
    structures generated (features stay distinct) to keep host-side generation of large batches fast."""
    structures = None
    if n_structures is not None and n_structures < n:
        structures = [_structure(shape, np.random.default_rng(seed0 + 7919 * (s + 1))) for s in range(n_structures)]
    return [make_sample(shape, seed0 + g, None if structures is None else structures[g % len(structures)])
            for g in range(n)]


def shuffle_edges(sample, seed: int):
    """Same graph with both edge lists in random order (exercises the unsorted path of the CSR build)."""
    (cons, cons_e, var, cut, cut_e), imp = sample
    rng = np.random.default_rng(seed)
    out = []
    for e in (cons_e, cut_e):
        p = rng.permutation(e["indices"].shape[1])
        out.append({"indices": e["indices"][:, p], "values": e["values"][p]})
    return (cons, out[0], var, cut, out[1]), imp
