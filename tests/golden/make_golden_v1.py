"""Golden fixtures for rows a15 (Model_Inference_v1 / node_predictions) and the object-classifier fine-tuning model, made by
running the UNMODIFIED reference (same torch_geometric stand-in as make_golden.py):

    python tests/golden/make_golden_v1.py

  model_v1_n90.npz   Model_Inference_v1.forward (gnn_detector.py:204-312) on one 90-point frame.  The reference has no
                     checkpoint for this variant, so its state_dict is ASSEMBLED from the checked-in one (same shapes):
                     predict_node.{stem, pred_cls} <- predict_node.*, predict_node.pred_offsets <- predict_offset.pred_offsets;
                     tests/test_v1_gpu.py applies the same mapping.
  finetune_n90.npz   Model_Object_Classifier_Finetuning.forward (gnn_detector.py:481-522) on two frames: loss, accuracy, and
                     the per-frame cluster member lists of the reference's Simple_DBSCAN (offsets mode).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import CKPT, REF, REPO, build_frame, install_shim  # noqa: E402


def v1_state_dict(ck):
    """checkpoint of Model_Training (keys 'pred.<...>') -> state_dict of Model_Inference_v1"""
    sd = {}
    for k, v in ck.items():
        k = k[len('pred.'):]
        if k.startswith('predict_offset.stem'):
            continue                                        # v1 has ONE stem for both node heads
        if k.startswith('predict_offset.pred_offsets'):
            k = 'predict_node.pred_offsets' + k[len('predict_offset.pred_offsets'):]
        sd[k] = v.clone()
    return sd


def main():
    sys.path.insert(0, REPO)
    sys.path.insert(0, REF)
    install_shim()
    os.chdir(REF)
    from modules.compute_features import graph_features as gf
    from modules.set_configurations.set_config_gnn import config
    from modules.neural_net.gnn.gnn_detector import Model_Inference_v1, Model_Object_Classifier_Finetuning
    from graph_neural_network_for_radar_perception_b200 import synth
    torch.set_num_threads(8)
    cfg = config(os.path.join(REF, 'configuration_radarscenes_gnn.yml'))
    eps, knn = cfg.ball_query_eps_square, cfg.k_number_nearest_points
    ck = torch.load(CKPT, map_location='cpu', weights_only=True)

    def tensors(fidx, n):
        data, src = synth.make_frame(fidx, n, knn=knn)
        adj, nf, ef = build_frame(cfg, gf, data, knn, eps)
        lab = synth.make_labels(data, src, adj['adj_list'])
        t = dict(node_features=torch.from_numpy(nf).to(torch.float32), edge_features=torch.from_numpy(ef).to(torch.float32),
                 edge_index=torch.from_numpy(adj['adj_list']).to(torch.int64), adj_matrix=torch.from_numpy(adj['adj_matrix']).to(torch.bool),
                 other_features=torch.from_numpy(np.stack((data['meas_px'], data['meas_py'], data['meas_vx'], data['meas_vy']), -1)).to(torch.float32))
        return data, t, lab

    # ---- Model_Inference_v1 ----
    m = Model_Inference_v1(cfg)
    print(m.load_state_dict(v1_state_dict(ck), strict=True))
    data, t, lab = tensors(30, 90)
    clusters = [torch.from_numpy(c) for c in lab['cluster_node_idx']]
    with torch.no_grad():
        o = m.eval()(t['node_features'], t['edge_features'], t['edge_index'], t['adj_matrix'], clusters)
    np.savez_compressed(os.path.join(HERE, 'model_v1_n90.npz'),
                        node_features=t['node_features'].numpy(), edge_features=t['edge_features'].numpy(), edge_index=t['edge_index'].numpy(),
                        cluster_ptr=np.cumsum([0] + [len(c) for c in lab['cluster_node_idx']]).astype(np.int64),
                        cluster_members=np.concatenate(lab['cluster_node_idx']),
                        node_cls=o[0].numpy(), node_off=o[1].numpy(), link_cls=o[2].numpy(), obj_cls=o[3].numpy())
    print('v1', [tuple(x.shape) for x in o])

    # ---- Model_Object_Classifier_Finetuning ----
    ft = Model_Object_Classifier_Finetuning(cfg)
    print(ft.load_state_dict(ck, strict=True))
    frames = [tensors(31, 90), tensors(32, 60)]
    node_cls = [torch.from_numpy(f[2]['node_class']) for f in frames]
    ft.eval()
    members = []
    with torch.no_grad():
        loss, acc = ft([f[1]['node_features'] for f in frames], [f[1]['edge_features'] for f in frames], [f[1]['other_features'] for f in frames],
                       [f[1]['edge_index'] for f in frames], [f[1]['adj_matrix'] for f in frames], node_cls)
        for f in frames:
            r = ft.pred(node_features=f[1]['node_features'], edge_features=f[1]['edge_features'], other_features=f[1]['other_features'],
                        edge_index=f[1]['edge_index'], adj_matrix=f[1]['adj_matrix'])
            members.append([np.asarray(c) for c in r[4]])
    out = dict(loss=loss.numpy(), accuracy=acc.numpy(), clustering_eps=np.float64(cfg.clustering_eps))
    for i, f in enumerate(frames):
        out[f'f{i}_node_features'] = f[1]['node_features'].numpy()
        out[f'f{i}_edge_features'] = f[1]['edge_features'].numpy()
        out[f'f{i}_other_features'] = f[1]['other_features'].numpy()
        out[f'f{i}_edge_index'] = f[1]['edge_index'].numpy()
        out[f'f{i}_node_class'] = f[2]['node_class']
        out[f'f{i}_member_ptr'] = np.cumsum([0] + [len(c) for c in members[i]]).astype(np.int64)
        out[f'f{i}_members'] = np.concatenate(members[i]).astype(np.int64)
    np.savez_compressed(os.path.join(HERE, 'finetune_n90.npz'), **out)
    print('finetune', float(loss), float(acc), [len(m_) for m_ in members])


if __name__ == '__main__':
    main()
