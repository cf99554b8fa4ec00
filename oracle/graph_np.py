"""ORACLE (test infrastructure, never the product path): NumPy restatement of the reference's graph
construction, modules/compute_features/graph_features.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may import
this file.  Parity status: PINNED against the reference itself -- tests/golden/make_golden.py imports the
reference's graph_features.py from /root/reference (it is NumPy-only and runs unchanged) and the
committed fixtures tests/golden/graph_*.npz hold its outputs; tests/test_oracle_golden.py checks this
restatement against them bit for bit.

Arithmetic contract restated (each with the reference line it follows):

* squared distance  D[i,j] = fl32(fl32(dx*dx) + fl32(dy*dy)), dx = px[i]-px[j]   (graph_features.py:70-75,
  a batched 1x2 @ 2x1 float32 matmul: two roundings, no FMA -- SURVEY.md appendix C item 1)
* radius gate       D <= eps, diagonal cleared; degree = row sums (int64)            (:11-22, :78)
* kNN               per row the knn+1 smallest of D (self included), or all if knn >= N;
                    both (i,j) and (j,i) set, diagonal cleared                       (:25-44)
                    ties are ordered by (D, column index) here; the reference's argsort is unstable, so
                    ties across the k-th slot are undefined there and excluded from parity data.
* adj_list          np.where(adj): row-major, row = edge_index[0] = source           (:79)
* node features     [vr, rcs, t_norm, degree/10, range_conf, azimuth_conf] in float64 (:117-144, :47-55)
* edge features     [dx/10, dy/10, sqrt(dx'^2+dy'^2)/10, dvx, dvy, sqrt(dvx^2+dvy^2), dt*1e-6]  (:147-164)
"""
from __future__ import annotations

import numpy as np


def squared_distance_matrix(px: np.ndarray, py: np.ndarray) -> np.ndarray:
    """graph_features.py:70-75 -- float32, products rounded separately, then one rounded add."""
    px = np.asarray(px, dtype=np.float32)
    py = np.asarray(py, dtype=np.float32)
    dx = px[:, None] - px[None, :]
    dy = py[:, None] - py[None, :]
    xx = np.multiply(dx, dx, dtype=np.float32)
    yy = np.multiply(dy, dy, dtype=np.float32)
    return np.add(xx, yy, dtype=np.float32)


def radius_gate(d2: np.ndarray, eps) -> np.ndarray:
    """graph_features.py:11-22 (eps is a SQUARED distance, yml:13)."""
    gate = d2 <= eps
    np.fill_diagonal(gate, False)
    return gate


def knn_gate(d2: np.ndarray, knn: int) -> np.ndarray:
    """graph_features.py:25-44 with the (D, index) tie order made explicit."""
    n = d2.shape[0]
    take = n if knn >= n else knn + 1
    order = np.argsort(d2, axis=-1, kind='stable')[:, :take]
    gate = np.zeros((n, n), dtype=np.bool_)
    rows = np.repeat(np.arange(n), take)
    cols = order.reshape(-1)
    gate[rows, cols] = True
    gate[cols, rows] = True
    np.fill_diagonal(gate, False)
    return gate


def adjacency_information(data: dict, eps, knn: int, union_radius: bool = False) -> dict:
    """compute_adjacency_information (:58-84) and, with union_radius=True, the _v2 variant (:87-114)."""
    d2 = squared_distance_matrix(data['meas_px'], data['meas_py'])
    ball = radius_gate(d2, eps)
    adj = knn_gate(d2, knn)
    if union_radius:
        adj = adj | ball
    degree = ball.sum(axis=-1)
    src, dst = np.nonzero(adj)
    return {'adj_matrix': adj, 'distance_mat': d2,
            'adj_list': np.stack((src, dst), axis=0), 'degree': degree}


def time_unit_interval(ts: np.ndarray) -> np.ndarray:
    """normalize_time (:47-55): integer timestamps -> float64 in [0,1], or integer zeros if all equal."""
    lo, hi = ts.min(), ts.max()
    if hi == lo:
        return ts - lo
    return (ts - lo) / (hi - lo)


def node_features(data: dict, degree: np.ndarray, include_region_confidence: bool = False,
                  min_range=None, max_range=None, min_azimuth=None, max_azimuth=None) -> np.ndarray:
    """compute_node_features (:117-144).  Result dtype float64 (np.stack promotion); the caller casts
    to float32 as reference datagen_gnn.py:122 does.  Under NumPy 2 the range/azimuth terms are float64
    when max_range / max_azimuth are np.float64 scalars (SURVEY.md 8(c) "environment drift")."""
    cols = [data['meas_vr'], data['meas_rcs'], time_unit_interval(data['meas_timestamp']), degree / 10]
    if include_region_confidence:
        px, py = data['meas_px'], data['meas_py']
        rng_ = np.sqrt(px ** 2 + py ** 2)
        azi = np.abs(np.arctan2(py, px))
        cols.append((rng_ - max_range) / (min_range - max_range))
        cols.append((azi - max_azimuth) / (min_azimuth - max_azimuth))
    return np.stack(cols, axis=-1)


def edge_features(data: dict, adj_list: np.ndarray) -> np.ndarray:
    """compute_edge_features (:147-164): feature(source) - feature(target), source = adj_list[0]."""
    s, t = adj_list[0], adj_list[1]
    ex = (data['meas_px'][s] - data['meas_px'][t]) / 10
    ey = (data['meas_py'][s] - data['meas_py'][t]) / 10
    el = np.sqrt(ex ** 2 + ey ** 2) / 10
    evx = data['meas_vx'][s] - data['meas_vx'][t]
    evy = data['meas_vy'][s] - data['meas_vy'][t]
    ev = np.sqrt(evx ** 2 + evy ** 2)
    et = (data['meas_timestamp'][s] - data['meas_timestamp'][t]) * 1e-6
    return np.stack((ex, ey, el, evx, evy, ev, et), axis=-1)
