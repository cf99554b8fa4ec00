"""GPU: proposal clustering (csrc/rgnn_cluster.cu through the C-ABI) against the reference's own Simple_DBSCAN outputs
(tests/golden/clusters.npz) and against the oracle restatement on larger / batched inputs.  Integer work: bit-exact."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from graph_neural_network_for_radar_perception_b200 import clustering as cl


def _lists(res):
    return [m.cpu().numpy() for m in res.member_lists()]


def test_reference_interface_matches_reference_outputs(golden_dir):
    g = np.load(os.path.join(golden_dir, 'clusters.npz'))
    for c in range(int(g['n_cases'])):
        p = f'c{c}_'
        o = cl.Simple_DBSCAN(float(g[p + 'eps_links']), True)
        o.cluster_nodes(g[p + 'centres'], g[p + 'pred'], g[p + 'adj_matrix'])
        assert o.num_clusters == int(g[p + 'n_links'])
        assert np.array_equal(o.meas_to_cluster_id.astype(np.int64), g[p + 'ids_links'])
        ids = g[p + 'ids_links']
        for i, m in enumerate(_lists(o.result)):                  # cluster_members_list of gnn_detector.py:176-181
            assert np.array_equal(m, np.nonzero(ids == i)[0])
        # proposals of those clusters (row f4): float32 bit-exact against the reference's compute_proposals
        xy = torch.from_numpy(g[p + 'centres']).cuda()
        mean, cov, size, vote = cl.compute_proposals_device(o.result, xy[:, 0], xy[:, 1], 0.5 * np.eye(2, dtype=np.float32),
                                                            torch.from_numpy(g[p + 'node_logits']).cuda())
        assert np.array_equal(mean.cpu().numpy(), g[p + 'prop_mean']) and np.array_equal(cov.cpu().numpy(), g[p + 'prop_cov'])
        assert np.array_equal(size.cpu().numpy().astype(np.int64), g[p + 'prop_size'])
        assert np.array_equal(vote.cpu().numpy().astype(np.int64), g[p + 'prop_vote'])
        mu_l, sig_l, size_l = cl.compute_proposals(o.result.member_lists(), g[p + 'centres'][:, 0], g[p + 'centres'][:, 1],
                                                   0.5 * np.eye(2, dtype=np.float32))          # reference signature
        assert np.array_equal(np.stack(mu_l), g[p + 'prop_mean']) and size_l == g[p + 'prop_size'].tolist()
        o = cl.Simple_DBSCAN(float(g[p + 'eps_radius']), False)
        o.cluster_nodes(g[p + 'centres'])
        assert o.num_clusters == int(g[p + 'n_radius'])
        assert np.array_equal(o.meas_to_cluster_id.astype(np.int64), g[p + 'ids_radius'])


@pytest.mark.parametrize('n,frames', [(3000, 1), (900, 5)])
def test_large_and_batched_match_oracle(n, frames):
    from oracle.clustering_np import Simple_DBSCAN as Oracle
    rng = np.random.default_rng(n)
    xy_f, ids_want, off = [], [], 0
    for f in range(frames):
        centres = rng.uniform(0, 60, size=(n // 12, 2))
        xy = (centres[rng.integers(0, centres.shape[0], size=n)] + rng.normal(0, 0.4, size=(n, 2))).astype(np.float32)
        o = Oracle(0.5, False)
        o.cluster_nodes(xy)
        ids_want.append(o.meas_to_cluster_id.astype(np.int64) + off)
        off += o.num_clusters
        xy_f.append(xy)
    xy_all = torch.from_numpy(np.concatenate(xy_f)).cuda()
    res = cl.cluster_radius(xy_all, 0.5, [i * n for i in range(frames + 1)])
    assert res.n_clusters == off
    assert np.array_equal(res.cluster_id.cpu().numpy().astype(np.int64), np.concatenate(ids_want))
    want = np.concatenate(ids_want)
    for i, m in enumerate(_lists(res)):
        assert np.array_equal(m, np.nonzero(want == i)[0])
    # links mode on the same points: random pairs, random logits (ties -> class 0)
    npairs = 8 * n * frames
    a = rng.integers(0, n * frames, size=npairs)
    b = rng.integers(0, n * frames, size=npairs)
    keep = a < b
    a, b = a[keep], b[keep]
    logits = rng.normal(size=(a.shape[0], 2)).astype(np.float32)
    logits[::7, 1] = logits[::7, 0]
    pred = (logits[:, 1] > logits[:, 0]).astype(np.int64)
    o = Oracle(2.0, True)
    o.cluster_nodes(np.concatenate(xy_f), pred, und_pairs=np.stack([a, b]))
    res = cl.cluster_links(xy_all, torch.from_numpy(a.astype(np.int32)).cuda(), torch.from_numpy(b.astype(np.int32)).cuda(),
                           torch.from_numpy(logits).cuda(), a.shape[0], 2.0)
    assert res.n_clusters == o.num_clusters
    assert np.array_equal(res.cluster_id.cpu().numpy().astype(np.int64), o.meas_to_cluster_id.astype(np.int64))


def test_detector_forward_without_clusters_matches_two_stage_oracle(ckpt_state_dict):
    """Model_Inference.forward with cluster_node_idx=None (gnn_detector.py:164-187): the clusters found on the device equal the
    oracle's clusters of the same predicted centres / links, and the class head output equals a forward with those clusters."""
    from gpu_util import load_model
    from graph_neural_network_for_radar_perception_b200 import synth
    from graph_neural_network_for_radar_perception_b200.compute_offsets import unnormalize_gt_offsets
    from oracle import graph_np
    from oracle.clustering_np import Simple_DBSCAN as Oracle
    m = load_model(ckpt_state_dict).pred.eval()
    R = np.float64(np.sqrt(100.0 ** 2 + 50.0 ** 2))
    d, src = synth.make_frame(321, 400)
    adj = graph_np.adjacency_information(d, 25, 10)
    nf = torch.from_numpy(graph_np.node_features(d, adj['degree'], True, 0, R, 0, np.pi * 0.5).astype(np.float32)).cuda()
    ef = torch.from_numpy(graph_np.edge_features(d, adj['adj_list']).astype(np.float32)).cuda()
    ei = torch.from_numpy(adj['adj_list']).cuda()
    other = torch.from_numpy(np.stack([d['meas_px'], d['meas_py'], d['meas_vx'], d['meas_vy']], axis=1).astype(np.float32)).cuda()
    for from_links in (True, False):
        m.set_param_for_proposal_extraction(1.4, from_links)
        with torch.no_grad():
            node_cls, node_off, link_cls, obj_cls, members = m(nf, ef, ei, None, None, other)
            reg = unnormalize_gt_offsets(node_off.clone(), m.reg_mu, m.reg_sigma)
            centres = (other[:, :2] + reg).cpu().numpy()
            o = Oracle(1.4, from_links)
            r, c = np.nonzero(np.triu(adj['adj_matrix'], k=1))
            o.cluster_nodes(centres, link_cls.argmax(dim=-1).cpu().numpy(), und_pairs=np.stack([r, c]))
            assert len(members) == o.num_clusters
            for i, mem in enumerate(members):
                assert np.array_equal(mem.cpu().numpy(), np.nonzero(o.meas_to_cluster_id == i)[0])
            again = m(nf, ef, ei, None, [mm for mm in members], other)
            assert torch.equal(again[3], obj_cls)
    m.extract_proposals = False
