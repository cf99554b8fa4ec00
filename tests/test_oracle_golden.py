"""CPU: the oracle restatements (oracle/) against the fixtures the REFERENCE produced (tests/golden)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import graph_np, model_torch

GRID_MAX_R = np.sqrt(100.0 ** 2 + 50.0 ** 2)     # reference set_config_gnn.py:44-47
GRID_MAX_TH = np.pi * 0.5


def _graph_cases(golden_dir):
    return sorted(glob.glob(os.path.join(golden_dir, 'graph_*.npz')))


def test_graph_oracle_matches_reference_bit_exact(golden_dir):
    files = _graph_cases(golden_dir)
    assert len(files) >= 6
    for f in files:
        g = np.load(f)
        data = {k: g[k] for k in g.files if k.startswith('meas_')}
        adj = graph_np.adjacency_information(data, float(g['eps']), int(g['knn']), union_radius=bool(g['v2']))
        assert np.array_equal(adj['adj_list'], g['adj_list']), f
        assert np.array_equal(adj['degree'], g['degree']), f
        assert float(adj['distance_mat'].astype(np.float64).sum()) == float(g['d2_checksum']), f
        if 'distance_mat' in g.files:
            assert np.array_equal(adj['distance_mat'], g['distance_mat'])
        nf = graph_np.node_features(data, adj['degree'], True, 0, GRID_MAX_R, 0, GRID_MAX_TH)
        ef = graph_np.edge_features(data, adj['adj_list'])
        assert np.array_equal(nf, g['node_features_f64']), f
        assert np.array_equal(ef, g['edge_features_f64']), f
        assert np.array_equal(nf.astype(np.float32), g['node_features'])
        assert np.array_equal(ef.astype(np.float32), g['edge_features'])
        # structural properties the reference guarantees: symmetric, no self loops, row-major order
        s, t = adj['adj_list']
        assert np.all(s != t)
        assert np.array_equal(np.lexsort((t, s)), np.arange(s.shape[0]))
        assert np.array_equal(adj['adj_matrix'], adj['adj_matrix'].T)


def _clusters(ptr, members):
    return [torch.from_numpy(members[ptr[i]:ptr[i + 1]]) for i in range(len(ptr) - 1)]


@pytest.mark.parametrize('case', ['n48', 'n200'])
def test_model_oracle_matches_reference(golden_dir, ckpt_state_dict, case):
    g = np.load(os.path.join(golden_dir, f'model_{case}.npz'))
    with torch.no_grad():
        out = model_torch.detector_forward(
            ckpt_state_dict, torch.from_numpy(g['node_features']), torch.from_numpy(g['edge_features']),
            torch.from_numpy(g['edge_index']), _clusters(g['cluster_ptr'], g['cluster_members']))
    for o, k in zip(out, ['node_cls', 'node_off', 'link_cls', 'obj_cls']):
        # same ATen ops in the same order as the reference -> expected to be (near) bit-identical
        np.testing.assert_allclose(o.numpy(), g[k], rtol=1e-5, atol=1e-6, err_msg=k)


def test_training_oracle_matches_reference(golden_dir, ckpt_state_dict):
    g = np.load(os.path.join(golden_dir, 'train_2frames.npz'))
    sd = {k: v.clone().requires_grad_(True) for k, v in ckpt_state_dict.items()}
    nf, ef, ei = [], [], []
    labels = {k: [] for k in ('cluster_node_idx', 'cluster_labels', 'edge_class', 'node_class', 'node_offsets')}
    for i in range(2):
        nf.append(torch.from_numpy(g[f'f{i}_node_features']))
        ef.append(torch.from_numpy(g[f'f{i}_edge_features']))
        ei.append(torch.from_numpy(g[f'f{i}_edge_index']))
        labels['cluster_node_idx'].append(_clusters(g[f'f{i}_cluster_ptr'], g[f'f{i}_cluster_members']))
        for k in ('cluster_labels', 'edge_class', 'node_class', 'node_offsets'):
            labels[k].append(torch.from_numpy(g[f'f{i}_{k}']))
    loss, acc, _ = model_torch.training_forward(sd, nf, ef, ei, labels)
    for k, v in loss.items():
        np.testing.assert_allclose(v.detach().numpy(), g[k], rtol=1e-5, err_msg=k)
    for k, v in acc.items():
        np.testing.assert_allclose(v.numpy(), g[k], rtol=1e-6, err_msg=k)
    sum(loss.values()).backward()
    names = [str(n) for n in g['grad_names']]
    assert set(names) == set(sd.keys())
    for n, s_ref, l2_ref in zip(names, g['grad_sums'], g['grad_norms']):
        gr = sd[n].grad.numpy().astype(np.float64)
        assert abs(np.sqrt((gr ** 2).sum()) - l2_ref) <= 1e-4 * l2_ref + 1e-7, n
        key = 'grad::' + n
        if key in g.files:
            np.testing.assert_allclose(sd[n].grad.numpy(), g[key], rtol=1e-4, atol=1e-5 * max(l2_ref, 1e-3), err_msg=n)


def test_init_loss_anchors():
    """with the -ln(99) bias init the edge loss is about 2*0.25*0.99^2*ln(100)."""
    logits = torch.full((100, 2), -float(np.log(99.0)))
    gt = torch.zeros(100, dtype=torch.int64)
    l = model_torch.sigmoid_focal(logits, torch.nn.functional.one_hot(gt, 2).float()).sum(-1).mean() * 2.0
    assert abs(float(l) - 2 * (0.25 * 0.99 ** 2 * np.log(100.0) + 0.75 * 0.01 ** 2 * -np.log(0.99))) < 1e-4


def test_cluster_oracle_matches_reference_dbscan(golden_dir):
    """oracle/clustering_np.py against outputs of the reference's own Simple_DBSCAN (tests/golden/clusters.npz)."""
    from oracle.clustering_np import Simple_DBSCAN
    g = np.load(os.path.join(golden_dir, 'clusters.npz'))
    for c in range(int(g['n_cases'])):
        p = f'c{c}_'
        o = Simple_DBSCAN(float(g[p + 'eps_links']), True)
        o.cluster_nodes(g[p + 'centres'], g[p + 'pred'], g[p + 'adj_matrix'])
        assert o.num_clusters == int(g[p + 'n_links'])
        assert np.array_equal(o.meas_to_cluster_id.astype(np.int64), g[p + 'ids_links'])
        o = Simple_DBSCAN(float(g[p + 'eps_radius']), False)
        o.cluster_nodes(g[p + 'centres'])
        assert o.num_clusters == int(g[p + 'n_radius'])
        assert np.array_equal(o.meas_to_cluster_id.astype(np.int64), g[p + 'ids_radius'])
        # proposals (row f4): bit-exact float32 against the reference's compute_proposals, votes against torch.bincount
        from oracle.clustering_np import proposals_np
        mean, cov, size, vote = proposals_np(g[p + 'ids_links'], g[p + 'centres'][:, 0], g[p + 'centres'][:, 1],
                                             0.5 * np.eye(2, dtype=np.float32), g[p + 'node_logits'])
        assert np.array_equal(mean, g[p + 'prop_mean']) and np.array_equal(cov, g[p + 'prop_cov'])
        assert np.array_equal(size, g[p + 'prop_size']) and np.array_equal(vote, g[p + 'prop_vote'])


def test_accumulation_oracle_matches_reference_bit_exact(golden_dir):
    """the restated window accumulation (oracle/accumulate_np.py) against the outputs of the
    reference's own extract_and_sync_radar_data / compute_ground_truth / select_* (tests/golden/make_golden_accumulate.py)."""
    from oracle import accumulate_np as acc
    from graph_neural_network_for_radar_perception_b200 import synth
    g = np.load(os.path.join(golden_dir, 'accumulate.npz'))
    n_cases = len([k for k in g.files if k.endswith('_args')])
    assert n_cases >= 5
    for c in range(n_cases):
        p = f'c{c}_'
        w, ns, pps, flip = (int(v) for v in g[p + 'args'])
        mounts, rad, odo, win = synth.make_raw_window(w, ns, pps)
        d = acc.accumulate_window(mounts, rad, odo, win, flip_along_x=bool(flip))
        for k in ('meas_px', 'meas_py', 'meas_vx', 'meas_vy', 'meas_vr', 'meas_rcs', 'meas_timestamp', 'stationary_meas_flag'):
            assert d[k].dtype == g[p + 'all_' + k].dtype, (c, k)
            assert np.array_equal(d[k], g[p + 'all_' + k]), (c, k)
        lab = acc.class_labels(d)
        assert np.array_equal(lab, g[p + 'all_class_labels']), c
        d_dyn, lab_dyn, idx = acc.select(d, lab)
        for k in ('meas_px', 'meas_py', 'meas_vx', 'meas_vy', 'meas_vr', 'meas_rcs', 'meas_timestamp', 'meas_sensorid',
                  'meas_label_id', 'stationary_meas_flag', 'meas_trackid'):
            assert np.array_equal(d_dyn[k], g[p + 'dyn_' + k]), (c, k)
        assert np.array_equal(lab_dyn, g[p + 'dyn_class_labels']), c
        assert np.all(np.diff(idx) > 0)
        # the fast (non-emulated) float64 product may differ in the last float64 bit only; after the float32 cast it is
        # the same array on these fixtures
        d2 = acc.accumulate_window(mounts, rad, odo, win, flip_along_x=bool(flip), exact_dgemm=False)
        assert np.array_equal(d2['meas_px'], d['meas_px']) and np.array_equal(d2['meas_py'], d['meas_py'])
