"""Deterministic synthetic RadarScenes-shaped frames (host side, NumPy).

The RadarScenes `radar_data.h5` blobs are not available offline, so every parity test and every
benchmark in this repo runs on frames produced here.  The recipe follows SURVEY.md section 8(d):

* ROI 100 m x 100 m (`x in [0,100)`, `y in [-50,50)`), the grid limits of the reference yml
  (reference configuration_radarscenes_gnn.yml:33-38).
* 30 % of the points are re-drawn as ceil(N/60) Gaussian blobs (sigma ~ U[0.5,2] m) to mimic objects.
* `vr ~ N(0,3^2)`, `rcs ~ N(-5,10^2)`, `vx = vr cos(th)`, `vy = vr sin(th)`, `th = atan2(py,px)`.
* timestamps are int64 microseconds inside a 10-scan window
  (reference modules/data_utils/read_data.py accumulates 10 scans; yml:12).
* labels mirror what reference modules/data_generator/datagen_gnn.py:15-45,126-139 produces:
  node class, node offsets to the blob centre, undirected-edge link labels, clusters = blobs + singletons.

A frame is rejected and re-drawn (next sub-seed) if any row of the squared-distance matrix has a tie
across the k-th neighbour slot, because the reference's unstable argsort leaves that case undefined
(SURVEY.md "Hard parts", appendix C item 2).
"""
from __future__ import annotations

import numpy as np

ROI_MIN_X, ROI_MAX_X = 0.0, 100.0
ROI_MIN_Y, ROI_MAX_Y = -50.0, 50.0
T0_US = 29_727_354_614_437
SCAN_US = 16_500


def _draw(rng: np.random.Generator, n: int):
    px = rng.uniform(ROI_MIN_X, ROI_MAX_X, n)
    py = rng.uniform(ROI_MIN_Y, ROI_MAX_Y, n)
    n_blob = int(np.ceil(n / 60))
    n_in_blob = int(0.3 * n)
    blob_of = np.full(n, -1, dtype=np.int64)
    if n_blob > 0 and n_in_blob > 0:
        cx = rng.uniform(ROI_MIN_X + 3, ROI_MAX_X - 3, n_blob)
        cy = rng.uniform(ROI_MIN_Y + 3, ROI_MAX_Y - 3, n_blob)
        sg = rng.uniform(0.5, 2.0, n_blob)
        members = rng.choice(n, n_in_blob, replace=False)
        which = rng.integers(0, n_blob, n_in_blob)
        px[members] = np.clip(cx[which] + sg[which] * rng.standard_normal(n_in_blob), ROI_MIN_X, ROI_MAX_X - 1e-3)
        py[members] = np.clip(cy[which] + sg[which] * rng.standard_normal(n_in_blob), ROI_MIN_Y, ROI_MAX_Y - 1e-3)
        blob_of[members] = which
    px = px.astype(np.float32)
    py = py.astype(np.float32)
    vr = (3.0 * rng.standard_normal(n)).astype(np.float32)
    rcs = (-5.0 + 10.0 * rng.standard_normal(n)).astype(np.float32)
    th = np.arctan2(py, px)
    vx = (vr * np.cos(th)).astype(np.float32)
    vy = (vr * np.sin(th)).astype(np.float32)
    scan = rng.integers(0, 10, n)
    ts = (T0_US + scan * SCAN_US + rng.integers(0, 1000, n)).astype(np.int64)
    node_class = rng.integers(0, 7, n).astype(np.int64)
    return dict(meas_px=px, meas_py=py, meas_vx=vx, meas_vy=vy, meas_vr=vr, meas_rcs=rcs,
                meas_timestamp=ts), blob_of, node_class


def _has_knn_tie(px: np.ndarray, py: np.ndarray, k: int) -> bool:
    """True if some row has d2[k-th] == d2[(k+1)-th] (self included), in the reference's f32 arithmetic."""
    n = px.shape[0]
    if k + 1 >= n:
        return False
    blk = 2048
    for s in range(0, n, blk):
        dx = px[s:s + blk, None] - px[None, :]
        dy = py[s:s + blk, None] - py[None, :]
        d2 = (dx * dx).astype(np.float32) + (dy * dy).astype(np.float32)
        part = np.partition(d2, (k, k + 1), axis=1)
        if np.any(part[:, k] == part[:, k + 1]):
            return True
    return False


def make_frame(frame_idx: int, n_points: int, knn: int = 10, seed: int = 1234, with_labels: bool = True):
    """Return (data_dict, labels_src) for one frame; labels_src holds blob ids and node classes.

    `data_dict` has the keys the reference's graph functions read
    (reference modules/compute_features/graph_features.py:70,133-138,153-161).
    """
    attempt = 0
    while True:
        rng = np.random.default_rng([seed + frame_idx, attempt])
        data, blob_of, node_class = _draw(rng, n_points)
        if not _has_knn_tie(data['meas_px'], data['meas_py'], knn):
            break
        attempt += 1
        if attempt > 50:
            raise RuntimeError('could not draw a tie-free frame')
    return data, dict(blob_of=blob_of, node_class=node_class)


def make_labels(data: dict, labels_src: dict, adj_list: np.ndarray):
    """Build the label set of reference datagen_gnn.py:126-139 for a synthetic frame.

    adj_list is the (2,E) reference-order edge list; undirected edges are those with row<col, in
    row-major order (reference compute_edge_labels.py:17-19, gnn_blocks.py:295-296).
    Returns dict with numpy arrays: edge_class (E_u,) i64, node_class (N,) i64, node_offsets (N,2) f32,
    cluster_node_idx (list of i64 arrays), cluster_labels (C,) i64.
    """
    blob_of = labels_src['blob_of']
    node_class = labels_src['node_class'].copy()
    n = blob_of.shape[0]
    px, py = data['meas_px'], data['meas_py']
    offsets = np.zeros((n, 2), dtype=np.float32)
    clusters, cluster_labels = [], []
    for b in np.unique(blob_of[blob_of >= 0]):
        idx = np.nonzero(blob_of == b)[0]
        cx, cy = px[idx].mean(dtype=np.float64), py[idx].mean(dtype=np.float64)
        offsets[idx, 0] = (cx - px[idx]).astype(np.float32)
        offsets[idx, 1] = (cy - py[idx]).astype(np.float32)
        node_class[idx] = node_class[idx[0]]          # one class per object
        clusters.append(idx.astype(np.int64))
        cluster_labels.append(node_class[idx[0]])
    for i in np.nonzero(blob_of < 0)[0]:
        clusters.append(np.array([i], dtype=np.int64))
        cluster_labels.append(node_class[i])
    r, c = adj_list[0], adj_list[1]
    und = r < c
    same = (blob_of[r[und]] == blob_of[c[und]]) & (blob_of[r[und]] >= 0)
    return dict(edge_class=same.astype(np.int64), node_class=node_class,
                node_offsets=offsets, cluster_node_idx=clusters,
                cluster_labels=np.asarray(cluster_labels, dtype=np.int64))
