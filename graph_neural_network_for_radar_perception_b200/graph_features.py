"""Graph construction with the reference's function names and dict-in / dict-out interface
(reference modules/compute_features/graph_features.py:58-164), computed by the CUDA kernels of
csrc/rgnn_graph.cu, plus `build_graph_batch`, the batched device-resident entry the benchmarks and the
training fast path use (many frames per launch, outputs stay in HBM).

Exactness contract (tests/test_graph_gpu.py): adj_list / degree bit-exact for tie-free frames; edge
features and node features [0..4] bit-exact; the azimuth feature within 1 ulp (float32 arctan2).
kNN ties are ordered by (d2, index); the reference leaves them to an unstable argsort.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch

from ._cabi import check, lib, ptr, stream_ptr
from ._engine import GraphBatch, _require_cuda

_KEYS = ('meas_px', 'meas_py', 'meas_vx', 'meas_vy', 'meas_vr', 'meas_rcs', 'meas_timestamp')


def _device(device=None):
    if not torch.cuda.is_available():
        from ._cabi import RgnnError
        raise RgnnError('no CUDA device: graph construction has no CPU path in this implementation')
    return torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())


class BatchedFrames:
    """Device-resident result of build_graph_batch."""
    gb: GraphBatch
    degree: torch.Tensor          # (N,) int32 radius-gate neighbour counts
    edge_row: torch.Tensor        # (E,) int32 edge_index[0] (source), reference order, global node ids
    edge_col: torch.Tensor        # (E,) int32 edge_index[1] (target)
    node_features: Optional[torch.Tensor]   # (N,6) f32
    edge_features: Optional[torch.Tensor]   # (E,7) f32, reference order
    frame_ptr: Sequence[int]

    def edge_index(self) -> torch.Tensor:
        """(2,E) int64 like the reference's adj_list (global node ids)."""
        return torch.stack((self.edge_row, self.edge_col)).to(torch.int64)


def build_graph_batch(points: Dict[str, torch.Tensor], frame_ptr: Sequence[int], eps, knn: int,
                      union_radius: bool = False, min_range=0, max_range=None, min_azimuth=0, max_azimuth=None,
                      with_features: bool = True) -> BatchedFrames:
    """points: device tensors meas_px, meas_py (f32) [+ meas_vx, meas_vy, meas_vr, meas_rcs (f32),
    meas_timestamp (int64) when with_features], all frames concatenated; frame_ptr: host offsets (F+1)."""
    px, py = points['meas_px'].contiguous(), points['meas_py'].contiguous()
    _require_cuda(px, py)
    dev = px.device
    n, nf = int(px.shape[0]), len(frame_ptr) - 1
    fp_host = np.ascontiguousarray(np.asarray(frame_ptr, dtype=np.int32))
    fp_dev = torch.from_numpy(fp_host).to(dev)
    i32 = dict(dtype=torch.int32, device=dev)
    s = stream_ptr()
    degree = torch.empty(n, **i32)
    row_ptr = torch.empty(n + 1, **i32)
    n_edges_dev = torch.zeros(1, **i32)
    nbytes = lib().rgnn_graph_build_workspace_bytes(n, nf, knn)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)

    def run(col, cap):
        check(lib().rgnn_graph_build(ptr(px), ptr(py), ptr(fp_dev), fp_host.ctypes.data, nf, n, float(eps), int(knn),
                                     1 if union_radius else 0, ptr(degree), ptr(row_ptr), ptr(col), cap,
                                     ptr(n_edges_dev), ptr(ws), nbytes, s), 'rgnn_graph_build')

    if union_radius:
        col = torch.empty(1, **i32)
        run(col, 0)                                  # counting pass
        cap = int(n_edges_dev.item())
    else:
        cap = 2 * n * min(knn + 1, max(int(np.diff(fp_host).max()), 1))
    col = torch.empty(max(cap, 1), **i32)
    run(col, cap)
    E = int(n_edges_dev.item())
    if E > cap:
        raise RuntimeError(f'graph_build: edge capacity {cap} < {E}')
    col = col[:E]
    gb = GraphBatch()
    gb.n_nodes, gb.n_edges = n, E
    gb.frame_node_ptr = [int(v) for v in fp_host]
    gb.row_ptr, gb.src = row_ptr, col
    gb.tgt = torch.empty(max(E, 1), **i32)
    gb.perm = torch.empty(max(E, 1), **i32)
    gb.und_a = torch.empty(max(E, 1), **i32)
    gb.und_b = torch.empty(max(E, 1), **i32)
    n_und = torch.zeros(1, **i32)
    nb2 = lib().rgnn_graph_finalize_workspace_bytes(n, E)
    ws2 = torch.empty(nb2, dtype=torch.uint8, device=dev)
    check(lib().rgnn_graph_finalize(ptr(row_ptr), ptr(col), n, E, ptr(gb.tgt), ptr(gb.perm), ptr(gb.und_a),
                                    ptr(gb.und_b), ptr(n_und), ptr(ws2), nb2, s), 'rgnn_graph_finalize')
    # the adjacency built here is symmetric without self loops (sym-kNN / radius U kNN), so exactly half of the
    # directed edges have source < target: no second device-to-host read (tests hold n_und_dev to it)
    gb.n_und = E // 2
    gb.n_und_dev = n_und
    gb.cl_ptr = torch.zeros(1, **i32)
    gb.cl_members = torch.zeros(1, **i32)
    out = BatchedFrames()
    out.gb, out.degree, out.frame_ptr = gb, degree, gb.frame_node_ptr
    # symmetric adjacency: row-major list (row=source, col=target) and target-major CSR share the arrays
    out.edge_row, out.edge_col = gb.tgt[:E], col
    out.node_features = out.edge_features = None
    if with_features:
        node_f = torch.empty((n, 6), dtype=torch.float32, device=dev)
        edge_f = torch.empty((E, 7), dtype=torch.float32, device=dev)
        if max_range is None:
            max_range = np.sqrt(100.0 ** 2 + 50.0 ** 2)
        if max_azimuth is None:
            max_azimuth = np.pi * 0.5
        r64 = isinstance(max_range, np.floating) or isinstance(min_range, np.floating)
        a64 = isinstance(max_azimuth, np.floating) or isinstance(min_azimuth, np.floating)
        ts = points['meas_timestamp'].to(torch.int64).contiguous()
        f = lambda k: points[k].to(torch.float32).contiguous()
        check(lib().rgnn_graph_features(ptr(px), ptr(py), ptr(f('meas_vx')), ptr(f('meas_vy')), ptr(f('meas_vr')),
                                        ptr(f('meas_rcs')), ptr(ts), ptr(degree), ptr(fp_dev), nf, n,
                                        ptr(out.edge_row), ptr(out.edge_col), E, float(min_range), float(max_range),
                                        float(min_azimuth), float(max_azimuth), int(r64), int(a64),
                                        ptr(node_f), ptr(edge_f), s), 'rgnn_graph_features')
        out.node_features, out.edge_features = node_f, edge_f
    return out


def frames_to_device(frames: Sequence[Dict[str, np.ndarray]], device=None):
    """List of reference-style data_dicts (NumPy) -> concatenated device tensors + frame_ptr."""
    dev = _device(device)
    fp = [0]
    for d in frames:
        fp.append(fp[-1] + int(d['meas_px'].shape[0]))
    pts = {}
    for k in _KEYS:
        if k in frames[0]:
            a = np.concatenate([np.asarray(d[k]) for d in frames])
            a = a.astype(np.int64) if k == 'meas_timestamp' else a.astype(np.float32)
            pts[k] = torch.from_numpy(a).to(dev)
    return pts, fp


class _AdjacencyDict(dict):
    """Result dict of compute_adjacency_information.  'adj_list' and 'degree' come from the GPU; the dense
    (N,N) 'adj_matrix' / 'distance_mat' entries of the reference are materialised on the host only if read."""

    def __init__(self, px, py, adj_list, degree):
        super().__init__(adj_list=adj_list, degree=degree)
        self._px, self._py = px, py

    def __missing__(self, key):
        n = self._px.shape[0]
        if key == 'adj_matrix':
            m = np.zeros((n, n), dtype=np.bool_)
            m[self['adj_list'][0], self['adj_list'][1]] = True
        elif key == 'distance_mat':
            dx = self._px[:, None] - self._px[None, :]
            dy = self._py[:, None] - self._py[None, :]
            m = (dx * dx).astype(np.float32) + (dy * dy).astype(np.float32)
        else:
            raise KeyError(key)
        self[key] = m
        return m


def _adjacency(data_dict, eps, knn, union_radius):
    pts, fp = frames_to_device([{k: data_dict[k] for k in ('meas_px', 'meas_py')}])
    bf = build_graph_batch(pts, fp, eps, knn, union_radius=union_radius, with_features=False)
    adj_list = torch.stack((bf.edge_row, bf.edge_col)).to(torch.int64).cpu().numpy()
    return _AdjacencyDict(np.asarray(data_dict['meas_px'], dtype=np.float32),
                          np.asarray(data_dict['meas_py'], dtype=np.float32),
                          adj_list, bf.degree.to(torch.int64).cpu().numpy())


def compute_adjacency_information(data_dict, eps, knn):
    """Symmetrised kNN adjacency + radius degree (reference graph_features.py:58-84)."""
    return _adjacency(data_dict, eps, knn, False)


def compute_adjacency_information_v2(data_dict, eps, knn):
    """radius U kNN adjacency (reference graph_features.py:87-114)."""
    return _adjacency(data_dict, eps, knn, True)


def _features(data_dict, degree, adj_list, min_range, max_range, min_azimuth, max_azimuth, want_nodes, want_edges):
    dev = _device()
    pts, fp = frames_to_device([data_dict], dev)
    n = fp[-1]
    i32 = dict(dtype=torch.int32, device=dev)
    fp_dev = torch.tensor(fp, **i32)
    deg = torch.from_numpy(np.asarray(degree, dtype=np.int32)).to(dev) if degree is not None else torch.zeros(n, **i32)
    node_f = torch.empty((n, 6), dtype=torch.float32, device=dev) if want_nodes else None
    E = 0 if adj_list is None else int(adj_list.shape[1])
    er = ec = edge_f = None
    if want_edges:
        er = torch.from_numpy(np.ascontiguousarray(adj_list[0]).astype(np.int32)).to(dev)
        ec = torch.from_numpy(np.ascontiguousarray(adj_list[1]).astype(np.int32)).to(dev)
        edge_f = torch.empty((E, 7), dtype=torch.float32, device=dev)
    r64 = isinstance(max_range, np.floating) or isinstance(min_range, np.floating)
    a64 = isinstance(max_azimuth, np.floating) or isinstance(min_azimuth, np.floating)
    z = torch.zeros(n, dtype=torch.float32, device=dev)
    g = lambda k: pts.get(k, z)
    check(lib().rgnn_graph_features(ptr(pts['meas_px']), ptr(pts['meas_py']), ptr(g('meas_vx')), ptr(g('meas_vy')),
                                    ptr(g('meas_vr')), ptr(g('meas_rcs')), ptr(pts['meas_timestamp']), ptr(deg),
                                    ptr(fp_dev), 1, n, ptr(er), ptr(ec), E, float(min_range or 0), float(max_range or 1),
                                    float(min_azimuth or 0), float(max_azimuth or 1), int(r64), int(a64),
                                    ptr(node_f), ptr(edge_f), stream_ptr()), 'rgnn_graph_features')
    return node_f, edge_f


def compute_node_features(data_dict, node_degree, include_region_confidence=False, min_range=None, max_range=None,
                          min_azimuth=None, max_azimuth=None):
    """(N,6) [vr, rcs, t_norm, degree/10, range_conf, azimuth_conf] (reference graph_features.py:117-144), or the
    first four columns without region confidence.  Returned as float32 NumPy (the reference returns float64 and
    casts to float32 in datagen_gnn.py:122; the values here ARE those float32 values)."""
    node_f, _ = _features(data_dict, node_degree, None, min_range, max_range, min_azimuth, max_azimuth, True, False)
    out = node_f.cpu().numpy()
    return out if include_region_confidence else out[:, :4]


def compute_edge_features(data_dict, adj_list):
    """(E,7) [dx, dy, dl, dvx, dvy, dv, dt] with feature(source) - feature(target) (reference :147-164)."""
    _, edge_f = _features(data_dict, None, adj_list, None, None, None, None, False, True)
    return edge_f.cpu().numpy()
