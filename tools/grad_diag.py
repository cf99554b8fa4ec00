"""Diagnostic: per-parameter gradient error of the CUDA training step against the reference fixture."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from gpu_util import load_model, clusters_from
g = np.load(os.path.join(ROOT, 'tests/golden/train_2frames.npz'))
ck = torch.load(os.path.join(ROOT, 'tests/golden/graph_based_detector.pt'), map_location='cpu', weights_only=True)
m = load_model(ck).train()
dev = 'cuda'
nf, ef, ei = [], [], []
labels = {k: [] for k in ('cluster_node_idx', 'cluster_labels', 'edge_class', 'node_class', 'node_offsets')}
for i in range(2):
    nf.append(torch.from_numpy(g[f'f{i}_node_features']).to(dev)); ef.append(torch.from_numpy(g[f'f{i}_edge_features']).to(dev))
    ei.append(torch.from_numpy(g[f'f{i}_edge_index']).to(dev))
    labels['cluster_node_idx'].append(clusters_from(g[f'f{i}_cluster_ptr'], g[f'f{i}_cluster_members'], dev))
    for k in ('cluster_labels', 'edge_class', 'node_class', 'node_offsets'):
        labels[k].append(torch.from_numpy(g[f'f{i}_{k}']).to(dev))
loss, acc = m(nf, ef, ei, [None, None], labels)
sum(loss.values()).backward()
rows = []
for n, p in m.named_parameters():
    key = 'grad::' + n
    if key in g.files and g[key].size > 1:
        ref = g[key].astype(np.float64); got = p.grad.cpu().numpy().astype(np.float64)
        rows.append((np.abs(got - ref).max() / np.abs(ref).max(), n))
rows.sort(reverse=True)
for e, n in rows[:12]:
    print(f'{e:.2e} {n}')
print('n compared', len(rows))
