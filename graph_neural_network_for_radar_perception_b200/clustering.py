"""Proposal clustering with the reference's `Simple_DBSCAN` interface (reference modules/inference/clustering.py:53-92),
computed on the GPU: `rgnn_cluster_links` / `rgnn_cluster_radius` of librgnn.so (csrc/rgnn_cluster.cu; connected
components by lock-free union-find).  There is no host path: inputs given as NumPy arrays (the reference's calling
convention) are copied to the current CUDA device first.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from ._cabi import check, lib, ptr, stream_ptr


class ClusterResult:
    """Device-side result: cluster_id (N,) int32, cl_ptr (N+1,) int32, cl_members (N,) int32, n_clusters (host int)."""

    def __init__(self, cluster_id, cl_ptr, cl_members, n_clusters):
        self.cluster_id, self.cl_ptr, self.cl_members, self.n_clusters = cluster_id, cl_ptr, cl_members, n_clusters

    def member_lists(self):
        """The reference's `cluster_members_list`: one int64 tensor of node ids per cluster (gnn_detector.py:176-181)."""
        sizes = (self.cl_ptr[1:self.n_clusters + 1] - self.cl_ptr[:self.n_clusters]).cpu().tolist()
        return list(torch.split(self.cl_members[:int(sum(sizes))].to(torch.int64), sizes))


def _outputs(n, dev):
    i32 = dict(dtype=torch.int32, device=dev)
    nbytes = lib().rgnn_cluster_workspace_bytes(n)
    return (torch.empty(n, **i32), torch.zeros(1, **i32), torch.empty(n + 1, **i32), torch.empty(n, **i32),
            torch.empty(nbytes, dtype=torch.uint8, device=dev), nbytes)


def cluster_links(xy: torch.Tensor, und_a: torch.Tensor, und_b: torch.Tensor, link_logits: torch.Tensor, n_und: int,
                  eps: float) -> ClusterResult:
    """Connected components of the predicted links that pass the distance gate (reference clustering.py:9-24)."""
    xy = xy.to(torch.float32).contiguous()
    link_logits = link_logits.to(torch.float32).contiguous()
    n = int(xy.shape[0])
    cid, nc, cl_ptr, cl_members, ws, nbytes = _outputs(n, xy.device)
    check(lib().rgnn_cluster_links(ptr(xy), ptr(und_a), ptr(und_b), ptr(link_logits), n, int(n_und), float(eps), ptr(cid), ptr(nc),
                                   ptr(cl_ptr), ptr(cl_members), ptr(ws), nbytes, stream_ptr()), 'rgnn_cluster_links')
    return ClusterResult(cid, cl_ptr, cl_members, int(nc.item()))


def cluster_radius(xy: torch.Tensor, eps: float, frame_ptr: Optional[list] = None) -> ClusterResult:
    """Connected components of the graph `squared distance <= eps` inside every frame (reference clustering.py:27-41)."""
    xy = xy.to(torch.float32).contiguous()
    n = int(xy.shape[0])
    frame_ptr = [0, n] if frame_ptr is None else list(frame_ptr)
    fp = torch.tensor(frame_ptr, dtype=torch.int32, device=xy.device)
    biggest = max(b - a for a, b in zip(frame_ptr[:-1], frame_ptr[1:]))
    cid, nc, cl_ptr, cl_members, ws, nbytes = _outputs(n, xy.device)
    check(lib().rgnn_cluster_radius(ptr(xy), ptr(fp), len(frame_ptr) - 1, n, int(biggest), float(eps), ptr(cid), ptr(nc), ptr(cl_ptr),
                                    ptr(cl_members), ptr(ws), nbytes, stream_ptr()), 'rgnn_cluster_radius')
    return ClusterResult(cid, cl_ptr, cl_members, int(nc.item()))


class Simple_DBSCAN:
    """Reference interface (clustering.py:53-92): `cluster_nodes(meas_xy, pred_edges, input_graph_adj_matrix)` fills
    `meas_to_cluster_id` (NumPy, like the reference) and `num_clusters`; `result` keeps the device-side lists."""

    def __init__(self, eps, compute_adj_mat_from_links=False):
        self.eps = eps
        self.compute_adj_mat_from_links = compute_adj_mat_from_links
        self.num_clusters = 0
        self.meas_to_cluster_id = None
        self.result: Optional[ClusterResult] = None

    def _finish(self, res: ClusterResult):
        self.result = res
        self.num_clusters = res.n_clusters
        n = int(res.cluster_id.shape[0])
        self.meas_to_cluster_id = res.cluster_id.cpu().numpy().astype(np.int16 if n < 32768 else np.int64)

    def cluster_nodes(self, meas_xy, pred_edges=None, input_graph_adj_matrix=None):
        if not torch.cuda.is_available():
            raise RuntimeError('Simple_DBSCAN runs on the GPU (librgnn.so); there is no CPU path')
        dev = torch.device('cuda', torch.cuda.current_device())
        xy = torch.as_tensor(np.asarray(meas_xy, dtype=np.float32)).to(dev)
        if self.compute_adj_mat_from_links:
            adj = torch.as_tensor(np.asarray(input_graph_adj_matrix)).to(dev)
            r, c = torch.nonzero(torch.triu(adj.to(torch.bool), diagonal=1), as_tuple=True)    # row-major, like np.nonzero
            pred = torch.as_tensor(np.asarray(pred_edges)).to(dev)
            logits = torch.stack((torch.zeros_like(pred, dtype=torch.float32), (pred == 1).to(torch.float32)), dim=1)
            self._finish(cluster_links(xy, r.to(torch.int32), c.to(torch.int32), logits, int(r.shape[0]), self.eps))
        else:
            self._finish(cluster_radius(xy, self.eps))

    def cluster_nodes_device(self, centres: torch.Tensor, gb=None, link_logits: Optional[torch.Tensor] = None):
        """Same, with everything already on the device (Model_Inference.forward): `gb` supplies the undirected pair list."""
        if self.compute_adj_mat_from_links:
            self._finish(cluster_links(centres, gb.und_a, gb.und_b, link_logits, gb.n_und, self.eps))
        else:
            self._finish(cluster_radius(centres, self.eps, gb.frame_node_ptr if gb is not None else None))
        return self.result


def compute_proposals_device(res: ClusterResult, px: torch.Tensor, py: torch.Tensor, meas_noise_cov, node_cls: Optional[torch.Tensor] = None):
    """Reference `compute_proposals` (modules/inference/inference.py:36-47) + the majority-vote object class of
    inference/output.py:111-118 for every cluster of `res`, on the device.  Returns (mean (C,2), cov (C,2,2), size (C,),
    vote (C,) or None) as tensors."""
    import ctypes as C
    dev = px.device
    px, py = px.to(torch.float32).contiguous(), py.to(torch.float32).contiguous()
    c = res.n_clusters
    mean = torch.empty((c, 2), dtype=torch.float32, device=dev)
    cov = torch.empty((c, 2, 2), dtype=torch.float32, device=dev)
    size = torch.empty(c, dtype=torch.int32, device=dev)
    vote = torch.empty(c, dtype=torch.int32, device=dev) if node_cls is not None else None
    noise = (C.c_float * 4)(*[float(v) for v in np.asarray(meas_noise_cov, dtype=np.float32).reshape(-1)])
    ncls = node_cls.to(torch.float32).contiguous() if node_cls is not None else None
    check(lib().rgnn_cluster_proposals(ptr(px), ptr(py), ptr(ncls), int(ncls.shape[1]) if ncls is not None else 0, ptr(res.cl_ptr),
                                       ptr(res.cl_members), c, noise, ptr(mean), ptr(cov), ptr(size), ptr(vote), stream_ptr()),
          'rgnn_cluster_proposals')
    return mean, cov, size, vote


def compute_proposals(cluster_members_list, px, py, meas_noise_cov):
    """Reference signature (inference.py:36-47): list of member tensors + NumPy / tensor coordinates -> lists of means,
    covariances and sizes (NumPy, like the reference)."""
    dev = torch.device('cuda', torch.cuda.current_device())
    sizes = [int(m.shape[0]) for m in cluster_members_list]
    cl_ptr = torch.zeros(len(sizes) + 1, dtype=torch.int32)
    cl_ptr[1:] = torch.tensor(sizes, dtype=torch.int32).cumsum(0)
    members = torch.cat([torch.as_tensor(m).to(dev) for m in cluster_members_list]).to(torch.int32) if sizes else torch.zeros(1, dtype=torch.int32, device=dev)
    res = ClusterResult(None, cl_ptr.to(dev), members, len(sizes))
    mean, cov, size, _ = compute_proposals_device(res, torch.as_tensor(np.asarray(px)).to(dev), torch.as_tensor(np.asarray(py)).to(dev), meas_noise_cov)
    mean, cov = mean.cpu().numpy(), cov.cpu().numpy()
    return [mean[i] for i in range(len(sizes))], [cov[i] for i in range(len(sizes))], sizes
