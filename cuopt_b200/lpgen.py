"""Synthetic LP generators for the BASELINE.json configurations (tests + bench.py; numpy only).

sparse_lp      — configs[1]/[3]: random sparse LP with a PLANTED optimal primal-dual pair, so the optimal
                 objective is known in closed form ("C2"/"C4").
multicommodity — configs[2]: pds-shaped multicommodity-flow LP (node-arc incidence blocks per commodity +
                 joint capacity rows); no pds file exists offline, so the shape is synthesised ("C3").
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class LP:
    """min c'x + offset  s.t.  con_lb <= A x <= con_ub,  var_lb <= x <= var_ub   (CSR A)."""
    offsets: np.ndarray
    indices: np.ndarray
    values: np.ndarray
    c: np.ndarray
    var_lb: np.ndarray
    var_ub: np.ndarray
    con_lb: np.ndarray
    con_ub: np.ndarray
    optimal_objective: float | None = None
    x_star: np.ndarray | None = None
    y_star: np.ndarray | None = None
    name: str = ""

    @property
    def m(self): return len(self.con_lb)
    @property
    def n(self): return len(self.c)
    @property
    def nnz(self): return len(self.values)

    def sense_form(self):
        """(sense bytes, rhs) when every row is E / L / G (true for the generators here)."""
        sense = np.where(self.con_lb == self.con_ub, ord("E"), np.where(np.isinf(self.con_ub), ord("G"), ord("L")))
        rhs = np.where(np.isinf(self.con_ub), self.con_lb, self.con_ub)
        return sense.astype(np.uint8).tobytes(), rhs

    def algorithmic_bytes_per_iteration(self) -> float:
        """B_iter = 24 nnz + 4(m+1) + 4(n+1) + 8(14 n + 7 m)."""
        return 24.0 * self.nnz + 4.0 * (self.m + 1) + 4.0 * (self.n + 1) + 8.0 * (14.0 * self.n + 7.0 * self.m)


def sparse_lp(m: int, n: int, nnz_per_row: int = 8, seed: int = 1234, locality: float = 0.0, bands: int = 8,
              upper_bound: float | None = None) -> LP:
    """Random sparse LP with planted optimum.

    Each row has `nnz_per_row` distinct columns, uniform in [0, n) (a fraction `locality` of them drawn from the
    row's own 1/bands column band instead), values N(0,1).  Rows: 50 % 'E', 25 % 'L', 25 % 'G'; half of the
    inequality rows are inactive at the optimum.  x*: half zero, half U(0,10); reduced costs U(0,1) on the zero half.
    """
    rng = np.random.default_rng(seed)
    k = nnz_per_row
    cols = rng.integers(0, n, size=(m, k), dtype=np.int64)
    if locality > 0.0:
        band = (np.arange(m, dtype=np.int64) * bands // max(m, 1))[:, None]
        width = max(n // bands, 1)
        local = band * width + rng.integers(0, width, size=(m, k), dtype=np.int64)
        cols = np.where(rng.random((m, k)) < locality, np.minimum(local, n - 1), cols)
    cols.sort(axis=1)
    # make the columns of each row distinct (rare repairs)
    for _ in range(64):
        dup = np.zeros((m, k), bool)
        dup[:, 1:] = cols[:, 1:] == cols[:, :-1]
        if not dup.any():
            break
        cols[dup] = rng.integers(0, n, size=int(dup.sum()))
        cols.sort(axis=1)
    vals = rng.standard_normal((m, k))
    offsets = (np.arange(m + 1, dtype=np.int64) * k).astype(np.int32)
    indices = cols.reshape(-1).astype(np.int32)
    values = vals.reshape(-1)

    x_star = np.where(rng.random(n) < 0.5, 0.0, rng.uniform(0.0, 10.0, n))
    if upper_bound is not None:
        x_star = np.minimum(x_star, upper_bound)
    ax = (vals * x_star[cols]).sum(axis=1)
    kind = rng.random(m)
    is_e = kind < 0.5
    is_l = (kind >= 0.5) & (kind < 0.75)
    is_g = kind >= 0.75
    active = rng.random(m) < 0.5
    y_star = rng.standard_normal(m)
    y_star = np.where(is_l, -np.abs(y_star), np.where(is_g, np.abs(y_star), y_star))
    y_star = np.where(~is_e & ~active, 0.0, y_star)
    slack = rng.uniform(0.0, 1.0, m)
    inf = np.inf
    con_lb = np.where(is_e, ax, np.where(is_g, np.where(active, ax, ax - slack), -inf))
    con_ub = np.where(is_e, ax, np.where(is_l, np.where(active, ax, ax + slack), inf))
    r_star = np.where(x_star > 0.0, 0.0, rng.uniform(0.0, 1.0, n))
    if upper_bound is not None:
        r_star = np.where(x_star >= upper_bound, -rng.uniform(0.0, 1.0, n), r_star)
    aty = np.zeros(n)
    np.add.at(aty, cols.reshape(-1), (vals * y_star[:, None]).reshape(-1))
    c = aty + r_star
    var_lb = np.zeros(n)
    var_ub = np.full(n, inf if upper_bound is None else float(upper_bound))
    return LP(offsets, indices, values, c, var_lb, var_ub, con_lb, con_ub, float(c @ x_star), x_star, y_star,
              name=f"sparse_lp(m={m},n={n},k={k},seed={seed},locality={locality})")


def multicommodity(nodes: int = 1200, arcs: int = 3600, commodities: int = 11, seed: int = 1234) -> LP:
    """pds-shaped multicommodity min-cost flow: K node-arc incidence blocks ('E' rows, +-1 entries, 2 nnz/col)
    coupled by one joint capacity row per arc ('L', K nnz/row).  Default sizes give ~17K x 40K; use
    nodes=9000, arcs=27000 for the ~100K x 300K shape of configs[2]."""
    rng = np.random.default_rng(seed)
    tail = rng.integers(0, nodes, arcs)
    head = (tail + 1 + rng.integers(0, nodes - 1, arcs)) % nodes
    # a ring guarantees connectivity, so every supply can be routed
    tail[:nodes] = np.arange(nodes)
    head[:nodes] = (np.arange(nodes) + 1) % nodes
    K = commodities
    n = K * arcs
    m = K * nodes + arcs
    # a feasible flow: each commodity ships `d` units from s to t along the ring
    flow = np.zeros((K, arcs))
    supply = np.zeros((K, nodes))
    for k in range(K):
        s, t = rng.integers(0, nodes, 2)
        d = float(rng.integers(1, 20))
        supply[k, s] += d
        supply[k, t] -= d
        j = s
        while j != t:
            flow[k, j] += d
            j = (j + 1) % nodes
    cap = np.maximum(1.3 * flow.sum(axis=0), 5.0)
    cost = rng.integers(1, 101, size=(K, arcs)).astype(float)
    rows, cols, vals = [], [], []
    for k in range(K):
        col = k * arcs + np.arange(arcs)
        rows += [k * nodes + tail, k * nodes + head, K * nodes + np.arange(arcs)]
        cols += [col, col, col]
        vals += [np.ones(arcs), -np.ones(arcs), np.ones(arcs)]
    rows = np.concatenate(rows); cols = np.concatenate(cols); vals = np.concatenate(vals)
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    offsets = np.zeros(m + 1, np.int64)
    np.add.at(offsets, rows + 1, 1)
    offsets = np.cumsum(offsets).astype(np.int32)
    b = np.concatenate([supply.reshape(-1), cap])
    con_lb = np.concatenate([supply.reshape(-1), np.full(arcs, -np.inf)])
    con_ub = b
    return LP(offsets, cols.astype(np.int32), vals, cost.reshape(-1), np.zeros(n), np.full(n, np.inf), con_lb,
              con_ub, None, None, None, name=f"multicommodity(V={nodes},E={arcs},K={K},seed={seed})")
