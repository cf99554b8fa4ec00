"""Generate the golden fixtures in this directory by running the UNMODIFIED reference from /root/reference.

Run once in the build container (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py

What it executes from the reference:
  * modules/compute_features/graph_features.py  (NumPy only, imported unchanged)
  * modules/neural_net/gnn/gnn_detector.py::Model_Training (and everything it imports), with the
    checked-in checkpoint model_weights/gnn/1718175257362/graph_based_detector.pt.
The one absent third-party dependency, torch_geometric (README.md:156 asks for >=2.5.0, no lock file), is
replaced by a stand-in written to a temp dir: MessagePassing.propagate restated as two index_selects, the
module's own message() and an index_add_ onto edge_index[1] (flow='source_to_target', aggr='add').

Outputs (all small, committed):
  graph_<case>.npz   inputs + adj_list/degree/node/edge features of compute_adjacency_information & co.
  model_<case>.npz   per-frame inputs + the four head outputs of Model_Inference
  train_<case>.npz   Model_Training losses / accuracies and gradients (selected tensors in full,
                     every tensor's sum and L2 norm)
  graph_based_detector.pt   copy of the reference checkpoint used (weights are data, not source)
"""
import os
import shutil
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
CKPT = os.path.join(REF, 'model_weights/gnn/1718175257362/graph_based_detector.pt')

SHIM = '''
import torch
class MessagePassing(torch.nn.Module):
    def __init__(self, aggr='add', flow='source_to_target', node_dim=-2):
        super().__init__(); self.aggr = aggr; self.flow = flow
    def propagate(self, edge_index, x, edge_attr=None, size=None):
        j, i = (0, 1) if self.flow == 'source_to_target' else (1, 0)
        x_j = x.index_select(0, edge_index[j]); x_i = x.index_select(0, edge_index[i])
        m = self.message(x_i=x_i, x_j=x_j, edge_attr=edge_attr)
        out = m.new_zeros((x.shape[0], m.shape[1]))
        if self.aggr in ('add', 'sum'): out.index_add_(0, edge_index[i], m)
        else: raise NotImplementedError
        return out
class GATv2Conv(torch.nn.Module):
    def __init__(self, *a, **k): super().__init__()
'''


def install_shim():
    d = tempfile.mkdtemp(prefix='pygshim_')
    pkg = os.path.join(d, 'torch_geometric', 'nn', 'conv')
    os.makedirs(pkg)
    open(os.path.join(d, 'torch_geometric', '__init__.py'), 'w').close()
    open(os.path.join(d, 'torch_geometric', 'nn', '__init__.py'), 'w').close()
    with open(os.path.join(pkg, '__init__.py'), 'w') as f:
        f.write(SHIM)
    sys.path.insert(0, d)
    return d


def build_frame(cfg, gf, data, knn, eps, v2=False):
    fn = gf.compute_adjacency_information_v2 if v2 else gf.compute_adjacency_information
    adj = fn(data, eps, knn)
    nf = gf.compute_node_features(data, adj['degree'], True, cfg.grid_min_r, cfg.grid_max_r,
                                  cfg.grid_min_th, cfg.grid_max_th)
    ef = gf.compute_edge_features(data, adj['adj_list'])
    return adj, nf, ef


def main():
    sys.path.insert(0, REPO)
    sys.path.insert(0, REF)
    shim_dir = install_shim()
    os.chdir(REF)
    from modules.compute_features import graph_features as gf
    from modules.set_configurations.set_config_gnn import config
    from modules.neural_net.gnn.gnn_detector import Model_Training
    from graph_neural_network_for_radar_perception_b200 import synth

    torch.manual_seed(0)
    torch.set_num_threads(8)
    cfg = config(os.path.join(REF, 'configuration_radarscenes_gnn.yml'))
    eps, knn = cfg.ball_query_eps_square, cfg.k_number_nearest_points

    # ---------------- graph fixtures ----------------
    graph_cases = [('n2', 0, 2, knn, False), ('n11', 1, 11, knn, False), ('n40', 2, 40, knn, False),
                   ('n300', 3, 300, knn, False), ('n300_v2', 3, 300, knn, True), ('n150_k4', 4, 150, 4, False),
                   ('n1000', 5, 1000, knn, False)]
    for name, fidx, n, k, v2 in graph_cases:
        data, _ = synth.make_frame(fidx, n, knn=k)
        adj, nf, ef = build_frame(cfg, gf, data, k, eps, v2)
        out = dict(data)
        out.update(eps=np.float64(eps), knn=np.int64(k), v2=np.bool_(v2),
                   adj_list=adj['adj_list'].astype(np.int64), degree=adj['degree'].astype(np.int64),
                   node_features=nf.astype(np.float32), edge_features=ef.astype(np.float32),
                   node_features_f64=nf, edge_features_f64=ef,
                   d2_checksum=np.float64(adj['distance_mat'].astype(np.float64).sum()))
        if n <= 40:
            out['distance_mat'] = adj['distance_mat']
        np.savez_compressed(os.path.join(HERE, f'graph_{name}.npz'), **out)
        print('graph', name, 'E =', adj['adj_list'].shape[1])

    # ---------------- model fixtures ----------------
    shutil.copyfile(CKPT, os.path.join(HERE, 'graph_based_detector.pt'))
    model = Model_Training(cfg, 'cpu')
    print(model.load_state_dict(torch.load(CKPT, map_location='cpu', weights_only=True)))

    def tensors(fidx, n):
        data, src = synth.make_frame(fidx, n, knn=knn)
        adj, nf, ef = build_frame(cfg, gf, data, knn, eps)
        lab = synth.make_labels(data, src, adj['adj_list'])
        t = dict(node_features=torch.from_numpy(nf).to(torch.float32),
                 edge_features=torch.from_numpy(ef).to(torch.float32),
                 edge_index=torch.from_numpy(adj['adj_list']).to(torch.int64),
                 adj_matrix=torch.from_numpy(adj['adj_matrix']).to(torch.bool))
        return data, t, lab

    for name, fidx, n in [('n48', 10, 48), ('n200', 11, 200)]:
        data, t, lab = tensors(fidx, n)
        clusters = [torch.from_numpy(c) for c in lab['cluster_node_idx']]
        with torch.no_grad():
            o = model.pred.eval()(t['node_features'], t['edge_features'], t['edge_index'], t['adj_matrix'], clusters)
        cl_ptr = np.cumsum([0] + [len(c) for c in lab['cluster_node_idx']]).astype(np.int64)
        np.savez_compressed(
            os.path.join(HERE, f'model_{name}.npz'),
            node_features=t['node_features'].numpy(), edge_features=t['edge_features'].numpy(),
            edge_index=t['edge_index'].numpy(), cluster_ptr=cl_ptr,
            cluster_members=np.concatenate(lab['cluster_node_idx']),
            node_cls=o[0].numpy(), node_off=o[1].numpy(), link_cls=o[2].numpy(), obj_cls=o[3].numpy())
        print('model', name, [tuple(x.shape) for x in o])

    # ---------------- training fixture (2 frames, losses + gradients) ----------------
    model.train()
    frames = [tensors(20, 40), tensors(21, 72)]
    labels = {'cluster_node_idx': [[torch.from_numpy(c) for c in f[2]['cluster_node_idx']] for f in frames],
              'cluster_labels': [torch.from_numpy(f[2]['cluster_labels']) for f in frames],
              'edge_class': [torch.from_numpy(f[2]['edge_class']) for f in frames],
              'node_class': [torch.from_numpy(f[2]['node_class']) for f in frames],
              'node_offsets': [torch.from_numpy(f[2]['node_offsets'].copy()) for f in frames]}
    model.zero_grad()
    loss, acc = model([f[1]['node_features'] for f in frames], [f[1]['edge_features'] for f in frames],
                      [f[1]['edge_index'] for f in frames], [f[1]['adj_matrix'] for f in frames], labels)
    total = sum(loss.values())
    total.backward()
    out = {}
    for i, f in enumerate(frames):
        out[f'f{i}_node_features'] = f[1]['node_features'].numpy()
        out[f'f{i}_edge_features'] = f[1]['edge_features'].numpy()
        out[f'f{i}_edge_index'] = f[1]['edge_index'].numpy()
        out[f'f{i}_cluster_ptr'] = np.cumsum([0] + [len(c) for c in f[2]['cluster_node_idx']]).astype(np.int64)
        out[f'f{i}_cluster_members'] = np.concatenate(f[2]['cluster_node_idx'])
        out[f'f{i}_cluster_labels'] = f[2]['cluster_labels']
        out[f'f{i}_edge_class'] = f[2]['edge_class']
        out[f'f{i}_node_class'] = f[2]['node_class']
        out[f'f{i}_node_offsets'] = f[2]['node_offsets']          # un-normalised (pre in-place edit)
    for k_, v in loss.items():
        out[k_] = v.detach().numpy()
    for k_, v in acc.items():
        out[k_] = v.detach().numpy()
    names, sums, norms = [], [], []
    full = ('encode_node_feat.encoder.0', 'encode_edge_feat.encoder.3', 'conv_blk.0.', 'conv_blk.6.',
            'predict_link.pred_cls', 'predict_class.pred_cls', 'predict_node.stem.0', 'predict_offset.pred_offsets',
            'predict_link.compute_edge', '.block.1.')
    for k_, p in model.named_parameters():
        g = p.grad.detach().numpy()
        names.append(k_)
        sums.append(g.astype(np.float64).sum())
        norms.append(np.sqrt((g.astype(np.float64) ** 2).sum()))
        if any(s in k_ for s in full):
            out['grad::' + k_] = g
    out['grad_names'] = np.array(names)
    out['grad_sums'] = np.array(sums)
    out['grad_norms'] = np.array(norms)
    np.savez_compressed(os.path.join(HERE, 'train_2frames.npz'), **out)
    print('train', {k_: float(v) for k_, v in loss.items()}, {k_: float(v) for k_, v in acc.items()})
    shutil.rmtree(shim_dir)


if __name__ == '__main__':
    main()
