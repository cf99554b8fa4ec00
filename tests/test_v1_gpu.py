"""GPU: SURVEY section 8 row a15 (Model_Inference_v1 / node_predictions, reference gnn_detector.py:204-312, gnn_blocks.py:392-439)
and the object-classifier fine-tuning model (gnn_detector.py:481-522) against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden_v1.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from gpu_util import assert_close, clusters_from, scaled_atol


def _v1_state_dict(ck):
    """the mapping of tests/golden/make_golden_v1.py: checkpoint of Model_Training -> state_dict of Model_Inference_v1"""
    sd = {}
    for k, v in ck.items():
        k = k[len('pred.'):]
        if k.startswith('predict_offset.stem'):
            continue
        if k.startswith('predict_offset.pred_offsets'):
            k = 'predict_node.pred_offsets' + k[len('predict_offset.pred_offsets'):]
        sd[k] = v.clone()
    return sd


def test_model_inference_v1_matches_reference_fixture(golden_dir, ckpt_state_dict):
    from graph_neural_network_for_radar_perception_b200 import Model_Inference_v1, config
    g = np.load(os.path.join(golden_dir, 'model_v1_n90.npz'))
    m = Model_Inference_v1(config())
    m.load_state_dict(_v1_state_dict(ckpt_state_dict), strict=True)
    m = m.cuda().eval()
    with torch.no_grad():
        out = m(torch.from_numpy(g['node_features']).cuda(), torch.from_numpy(g['edge_features']).cuda(),
                torch.from_numpy(g['edge_index']).cuda(), None, clusters_from(g['cluster_ptr'], g['cluster_members'], 'cuda'))
        # the dense adjacency matrix of the reference signature gives the same links
        n = g['node_features'].shape[0]
        adj = torch.zeros(n, n, dtype=torch.bool, device='cuda')
        ei = torch.from_numpy(g['edge_index']).cuda()
        adj[ei[0], ei[1]] = True
        out_adj = m(torch.from_numpy(g['node_features']).cuda(), torch.from_numpy(g['edge_features']).cuda(), ei, adj,
                    clusters_from(g['cluster_ptr'], g['cluster_members'], 'cuda'))
    for o, oa, k in zip(out, out_adj, ['node_cls', 'node_off', 'link_cls', 'obj_cls']):
        assert o.shape == g[k].shape, k
        assert_close(o.cpu().numpy(), g[k], 1e-4, scaled_atol(g[k]), k)
        assert torch.equal(o, oa), k


def test_object_classifier_finetuning_matches_reference_fixture(golden_dir, ckpt_state_dict):
    """Model_Object_Classifier_Finetuning: DBSCAN proposals on the device (offsets mode, eps from the yml), majority-vote
    labels, cross entropy on the per-cluster head; only that head receives a gradient after freezing."""
    from graph_neural_network_for_radar_perception_b200 import Model_Object_Classifier_Finetuning, config
    g = np.load(os.path.join(golden_dir, 'finetune_n90.npz'))
    cfg = config()
    assert float(cfg.clustering_eps) == float(g['clustering_eps'])
    m = Model_Object_Classifier_Finetuning(cfg)
    m.load_state_dict(ckpt_state_dict, strict=True)
    m = m.cuda()
    m.pred.freeze_layers_except_object_class_predictor()
    f = lambda i, k: torch.from_numpy(g[f'f{i}_{k}']).cuda()
    args = ([f(i, 'node_features') for i in range(2)], [f(i, 'edge_features') for i in range(2)], [f(i, 'other_features') for i in range(2)],
            [f(i, 'edge_index') for i in range(2)], [None, None], [f(i, 'node_class') for i in range(2)])
    # the clusters the device DBSCAN finds are the reference's (same members, same order)
    for i in range(2):
        with torch.no_grad():
            r = m.pred(node_features=args[0][i], edge_features=args[1][i], other_features=args[2][i], edge_index=args[3][i], adj_matrix=None)
        ptr, mem = g[f'f{i}_member_ptr'], g[f'f{i}_members']
        assert len(r[4]) == len(ptr) - 1
        for j, c in enumerate(r[4]):
            assert np.array_equal(np.asarray(c.cpu() if torch.is_tensor(c) else c), mem[ptr[j]:ptr[j + 1]]), (i, j)
    loss, acc = m(*args)
    assert_close(loss.item(), g['loss'], 1e-4, 1e-6, 'loss')
    assert abs(acc.item() - float(g['accuracy'])) < 1e-6
    loss.backward()
    for n, p in m.named_parameters():
        assert (p.grad is not None) == ('predict_class' in n), n
