"""GPU: graph construction kernels through the C-ABI against the reference fixtures and the NumPy oracle."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from graph_neural_network_for_radar_perception_b200 import graph_features as gf, synth
from oracle import graph_np

GRID_MAX_R = np.sqrt(100.0 ** 2 + 50.0 ** 2)
GRID_MAX_TH = np.pi * 0.5
# azimuth feature = (|atan2f| - pi/2) / (-pi/2): NumPy's float32 arctan2 (SIMD/libm, host dependent) and the
# correctly rounded value used on the GPU may differ by 1-2 ulp of the angle (1.2e-7 each) before the division
AZ_ATOL = 2e-7


def _ulp_diff(a, b):
    a = np.asarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b)


def test_adjacency_and_features_match_reference_fixtures(golden_dir):
    files = sorted(glob.glob(os.path.join(golden_dir, 'graph_*.npz')))
    assert len(files) >= 6
    for f in files:
        g = np.load(f)
        data = {k: g[k] for k in g.files if k.startswith('meas_')}
        fn = gf.compute_adjacency_information_v2 if bool(g['v2']) else gf.compute_adjacency_information
        adj = fn(data, float(g['eps']), int(g['knn']))
        assert adj['adj_list'].dtype == np.int64
        assert np.array_equal(adj['adj_list'], g['adj_list']), f        # bit-exact index contract
        assert np.array_equal(adj['degree'], g['degree']), f
        ef = gf.compute_edge_features(data, adj['adj_list'])
        assert np.array_equal(ef, g['edge_features']), f                # bit-exact (IEEE add/mul/div/sqrt, no FMA)
        nf = gf.compute_node_features(data, adj['degree'], True, 0, GRID_MAX_R, 0, GRID_MAX_TH)
        assert np.array_equal(nf[:, :5], g['node_features'][:, :5]), f
        assert np.abs(nf[:, 5] - g['node_features'][:, 5]).max() <= AZ_ATOL, f
        if int(g['meas_px'].shape[0]) <= 40:
            assert np.array_equal(adj['distance_mat'], g['distance_mat'])
            m = np.zeros_like(adj['adj_matrix']); m[g['adj_list'][0], g['adj_list'][1]] = True
            assert np.array_equal(adj['adj_matrix'], m)


@pytest.mark.parametrize('n,k', [(3, 10), (12, 10), (257, 10), (1000, 10), (3000, 10), (700, 16), (900, 32), (1200, 64)])
def test_batched_build_matches_oracle(n, k):
    frames = [synth.make_frame(100 + i, n + 7 * i, knn=k)[0] for i in range(3)]
    pts, fp = gf.frames_to_device(frames)
    bf = gf.build_graph_batch(pts, fp, 25, k, max_range=np.float64(GRID_MAX_R), max_azimuth=GRID_MAX_TH)
    ei = bf.edge_index().cpu().numpy()
    deg = bf.degree.cpu().numpy()
    nf = bf.node_features.cpu().numpy()
    ef = bf.edge_features.cpu().numpy()
    e0 = 0
    und_a, und_b = bf.gb.und_a.cpu().numpy(), bf.gb.und_b.cpu().numpy()
    u0 = 0
    for i, d in enumerate(frames):
        o = graph_np.adjacency_information(d, 25, k)
        E = o['adj_list'].shape[1]
        sl = slice(e0, e0 + E)
        assert np.array_equal(ei[:, sl] - fp[i], o['adj_list']), (n, k, i)
        assert np.array_equal(deg[fp[i]:fp[i + 1]], o['degree'])
        onf = graph_np.node_features(d, o['degree'], True, 0, np.float64(GRID_MAX_R), 0, GRID_MAX_TH).astype(np.float32)
        oef = graph_np.edge_features(d, o['adj_list']).astype(np.float32)
        assert np.array_equal(ef[sl], oef)
        assert np.array_equal(nf[fp[i]:fp[i + 1], :5], onf[:, :5])
        assert np.abs(nf[fp[i]:fp[i + 1], 5] - onf[:, 5]).max() <= AZ_ATOL
        r, c = o['adj_list']
        m = r < c
        nu = int(m.sum())
        assert np.array_equal(und_a[u0:u0 + nu] - fp[i], r[m]) and np.array_equal(und_b[u0:u0 + nu] - fp[i], c[m])
        u0 += nu
        e0 += E
    assert e0 == bf.gb.n_edges and u0 == bf.gb.n_und
    # reverse-edge permutation: involution, and it maps (s,t) to (t,s)
    perm = bf.gb.perm[:e0].cpu().numpy()
    assert np.array_equal(perm[perm], np.arange(e0))
    assert np.array_equal(ei[0][perm], ei[1]) and np.array_equal(ei[1][perm], ei[0])


def test_union_radius_variant_matches_oracle():
    frames = [synth.make_frame(300 + i, 400 + 50 * i)[0] for i in range(2)]
    pts, fp = gf.frames_to_device(frames)
    bf = gf.build_graph_batch(pts, fp, 25, 10, union_radius=True, with_features=False)
    ei = bf.edge_index().cpu().numpy()
    e0 = 0
    for i, d in enumerate(frames):
        o = graph_np.adjacency_information(d, 25, 10, union_radius=True)
        E = o['adj_list'].shape[1]
        assert np.array_equal(ei[:, e0:e0 + E] - fp[i], o['adj_list'])
        e0 += E
    assert e0 == bf.gb.n_edges


def test_csr_from_arbitrary_edge_index():
    from graph_neural_network_for_radar_perception_b200._engine import GraphBatch
    rng = np.random.default_rng(0)
    n, E = 500, 4000
    ei = rng.integers(0, n, size=(2, E)).astype(np.int64)
    gb = GraphBatch.from_edge_index(torch.from_numpy(ei).cuda(), n)
    row_ptr = gb.row_ptr.cpu().numpy(); src = gb.src.cpu().numpy()[:E]; tgt = gb.tgt.cpu().numpy()[:E]
    perm = gb.perm.cpu().numpy()[:E]
    order = np.lexsort((np.arange(E), ei[1]))           # stable by target
    assert np.array_equal(perm, order)
    assert np.array_equal(src, ei[0][order]) and np.array_equal(tgt, ei[1][order])
    assert np.array_equal(row_ptr, np.concatenate([[0], np.cumsum(np.bincount(ei[1], minlength=n))]))
    m = ei[0] < ei[1]
    assert gb.n_und == int(m.sum())
    assert np.array_equal(gb.und_a.cpu().numpy()[:gb.n_und], ei[0][m])
    assert np.array_equal(gb.und_b.cpu().numpy()[:gb.n_und], ei[1][m])


def test_large_frame_properties():
    """Full-size property checks (no O(N^2) oracle): symmetry, sortedness, degree bounds at N = 20000."""
    d, _ = synth.make_frame(999, 20000)
    pts, fp = gf.frames_to_device([d])
    bf = gf.build_graph_batch(pts, fp, 25, 10, with_features=False)
    ei = bf.edge_index().cpu().numpy()
    E = ei.shape[1]
    assert np.all(ei[0] != ei[1])
    assert np.all(np.diff(ei[0] * 20000 + ei[1]) > 0)                    # strictly row-major sorted, no duplicates
    perm = bf.gb.perm[:E].cpu().numpy()
    assert perm.min() >= 0 and np.array_equal(ei[0][perm], ei[1])          # symmetric
    deg_out = np.bincount(ei[0], minlength=20000)
    assert deg_out.min() >= 10


@pytest.mark.parametrize('kind', ['lattice', 'blobs', 'line', 'identical', 'mixed_sizes'])
def test_grid_knn_is_bit_identical_to_brute_force(kind):
    """The uniform-grid kNN (cell_sort_kernel + knn_grid_kernel) must give exactly the brute-force kernel's graph, including its
    (d2, index) order among EXACT distance ties (integer lattice, coincident points), for points crowded into a few cells (tight
    blobs, a line) and for a batch that mixes tiny and large frames; knn and the radius vary."""
    from graph_neural_network_for_radar_perception_b200._cabi import check, lib
    rng = np.random.default_rng(5)

    def frame(x, y):
        d, _ = synth.make_frame(1, len(x))
        d = dict(d)
        d['meas_px'] = np.asarray(x, dtype=np.float32)
        d['meas_py'] = np.asarray(y, dtype=np.float32)
        return d
    if kind == 'lattice':
        gx, gy = np.meshgrid(np.arange(40), np.arange(30))
        frames = [frame(gx.ravel() * 2.5, gy.ravel() * 2.5 - 37.0), frame(gx.ravel()[:700] * 1.0, gy.ravel()[:700] * 1.0)]
    elif kind == 'blobs':
        c = rng.uniform(0, 100, size=(6, 2))
        p = c[rng.integers(0, 6, 2500)] + rng.normal(0, 0.05, size=(2500, 2))
        frames = [frame(p[:, 0], p[:, 1] - 50), frame(rng.uniform(0, 100, 1500), rng.uniform(-50, 50, 1500))]
    elif kind == 'line':
        t = rng.uniform(0, 100, 1800)
        frames = [frame(t, np.full_like(t, 3.0)), frame(np.full_like(t, 7.0), t - 50)]
    elif kind == 'identical':
        frames = [frame(np.full(600, 12.5), np.full(600, -3.25)), frame(rng.uniform(0, 1e-3, 900), rng.uniform(0, 1e-3, 900))]
    else:
        frames = [frame(rng.uniform(0, 100, n), rng.uniform(-50, 50, n)) for n in (5, 3000, 47, 1, 900)]
    pts, fp = gf.frames_to_device(frames)
    for knn, eps, union in ((10, 25.0, False), (16, 4.0, True), (64, 100.0, False)):
        out = {}
        for mode in (1, 0):
            check(lib().rgnn_set_option(b'knn_grid', mode), 'opt')
            try:
                bf = gf.build_graph_batch(pts, fp, eps, knn, union_radius=union, with_features=False)
                out[mode] = (bf.edge_index().cpu().numpy(), bf.degree.cpu().numpy())
            finally:
                check(lib().rgnn_set_option(b'knn_grid', 1), 'opt')
        assert np.array_equal(out[1][0], out[0][0]), (kind, knn)
        assert np.array_equal(out[1][1], out[0][1]), (kind, knn)
