"""ORACLE (test infrastructure, never imported by the product package): CPU restatement of the reference's
`Simple_DBSCAN` (reference modules/inference/clustering.py:8-92), the proposal clustering between the link / offset
heads and the object-class head.

The reference's O(N^2) Python BFS is restated as a threshold graph + connected components; cluster ids are numbered in
order of their smallest member, which is exactly the BFS discovery order of the reference loop.  In the offsets mode `eps`
is compared with the SQUARED distance, in the links mode with the distance, as the reference does (clustering.py:9-41).
Pinned: tests/golden/clusters.npz holds outputs of the reference's own class (tests/golden/make_golden_clusters.py);
tests/test_oracle_golden.py holds this restatement to them.
"""
from __future__ import annotations

import numpy as np
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components


class Simple_DBSCAN:
    def __init__(self, eps, compute_adj_mat_from_links=False):
        self.eps = eps
        self.compute_adj_mat_from_links = compute_adj_mat_from_links
        self.num_clusters = 0
        self.meas_to_cluster_id = None

    @staticmethod
    def _threshold_pairs(xy: np.ndarray, eps) -> tuple:
        xy = np.asarray(xy, dtype=np.float32)
        n = xy.shape[0]
        rows, cols = [], []
        blk = 2048
        for s in range(0, n, blk):
            dx = xy[s:s + blk, None, 0] - xy[None, :, 0]
            dy = xy[s:s + blk, None, 1] - xy[None, :, 1]
            d2 = (dx * dx).astype(np.float32) + (dy * dy).astype(np.float32)
            r, c = np.nonzero(d2 <= eps)
            keep = (r + s) != c
            rows.append(r[keep] + s)
            cols.append(c[keep])
        return np.concatenate(rows), np.concatenate(cols)

    def cluster_nodes(self, meas_xy, pred_edges=None, input_graph_adj_matrix=None, und_pairs=None):
        n = meas_xy.shape[0]
        if self.compute_adj_mat_from_links:
            # keep predicted links whose end points are closer than eps (reference :9-24; note: true distance here)
            if und_pairs is None:
                und_pairs = np.stack(np.nonzero(np.triu(input_graph_adj_matrix, k=1)))
            r, c = und_pairs
            d = np.sqrt((meas_xy[r, 0] - meas_xy[c, 0]) ** 2 + (meas_xy[r, 1] - meas_xy[c, 1]) ** 2)
            keep = (np.asarray(pred_edges) == 1) & ~(d >= self.eps)
            rows, cols = r[keep], c[keep]
        else:
            rows, cols = self._threshold_pairs(meas_xy, self.eps)
        graph = coo_matrix((np.ones(rows.shape[0], dtype=np.int8), (rows, cols)), shape=(n, n))
        ncomp, lab = connected_components(graph, directed=False)
        # renumber components by their first (smallest-index) member = reference BFS discovery order
        first = np.full(ncomp, n, dtype=np.int64)
        np.minimum.at(first, lab, np.arange(n))
        order = np.argsort(first)
        remap = np.empty(ncomp, dtype=np.int64)
        remap[order] = np.arange(ncomp)
        self.meas_to_cluster_id = remap[lab].astype(np.int16 if n < 32768 else np.int64)
        self.num_clusters = int(ncomp)


def proposals_np(cluster_id, px, py, meas_noise_cov, node_logits=None):
    """Restatement of the reference's compute_proposals / compute_cluster_sample_mean_and_cov
    (modules/inference/inference.py:23-47) and of the majority-vote object class (inference/output.py:111-118) for the
    clusters encoded by `cluster_id` (meas_to_cluster_id).  float32 like the reference's arrays; rows are added in member
    order (NumPy's axis-0 reduction of an (n,2) array)."""
    cluster_id = np.asarray(cluster_id)
    xy = np.stack((np.asarray(px, dtype=np.float32), np.asarray(py, dtype=np.float32)), axis=-1)
    noise = np.asarray(meas_noise_cov, dtype=np.float32)
    n_c = int(cluster_id.max()) + 1 if cluster_id.size else 0
    mean, cov, size, vote = [], [], [], []
    pred = np.argmax(node_logits, axis=-1) if node_logits is not None else None
    for c in range(n_c):
        mem = np.nonzero(cluster_id == c)[0]
        v = xy[mem]
        mu = np.sum(v, axis=0) / v.shape[0]
        if v.shape[0] > 1:
            err = np.expand_dims(mu[:2] - v[:, :2], axis=-1)
            sig = np.sum(err @ err.transpose(0, 2, 1), axis=0) / (v.shape[0] - 1) + noise
        else:
            sig = noise
        mean.append(mu); cov.append(sig); size.append(v.shape[0])
        if pred is not None:
            vote.append(int(np.argmax(np.bincount(pred[mem]))))
    return (np.stack(mean).astype(np.float32), np.stack(cov).astype(np.float32), np.array(size, dtype=np.int64),
            np.array(vote, dtype=np.int64) if pred is not None else None)
