"""Diagnostic: forward error of the CUDA detector against the reference fixtures (per output: max abs error, max |ref|)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from gpu_util import load_model, clusters_from
ck = torch.load(os.path.join(ROOT, 'tests/golden/graph_based_detector.pt'), map_location='cpu', weights_only=True)
m = load_model(ck).pred.eval()
for case in ('n48', 'n200'):
    g = np.load(os.path.join(ROOT, f'tests/golden/model_{case}.npz'))
    with torch.no_grad():
        out = m(torch.from_numpy(g['node_features']).cuda(), torch.from_numpy(g['edge_features']).cuda(),
                torch.from_numpy(g['edge_index']).cuda(), None, clusters_from(g['cluster_ptr'], g['cluster_members'], 'cuda'))
    for o, k in zip(out, ['node_cls', 'node_off', 'link_cls', 'obj_cls']):
        a = o.cpu().numpy().astype(np.float64); b = g[k].astype(np.float64)
        err = np.abs(a - b)
        print(f'{case} {k:9s} max|ref| {np.abs(b).max():8.3f}  max abs err {err.max():.2e}  max err/(1e-4|ref|+2e-5) {(err / (1e-4 * np.abs(b) + 2e-5)).max():.2f}'
              f'  rms err {np.sqrt((err ** 2).mean()):.2e}')
