#!/usr/bin/env python
"""BASELINE.json configs[4]: edge-count stress sweep -- points per frame x kNN degree, reference channel plan.
Builds the graph and runs the detector forward on the GPU for every point of the sweep, checks the edge list of the first
frame against the oracle where that finishes in seconds (N <= 5000), and prints one JSON line per point.

    python tools/sweep.py [--frames-budget 600000] > profiles/sweep_rNN.jsonl
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from graph_neural_network_for_radar_perception_b200 import config, Model_Training, synth  # noqa: E402
from graph_neural_network_for_radar_perception_b200 import graph_features as gf  # noqa: E402

CKPT = os.path.join(ROOT, 'tests', 'golden', 'graph_based_detector.pt')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--points-budget', type=int, default=400000, help='points per batch (frames = budget / N)')
    ap.add_argument('--reps', type=int, default=3)
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    torch.cuda.set_device(dev)
    model = Model_Training(config(), dev)
    model.load_state_dict(torch.load(CKPT, map_location='cpu', weights_only=True))
    det = model.to(dev).pred.eval()
    R = np.float64(np.sqrt(100.0 ** 2 + 50.0 ** 2))
    for n in (500, 1000, 2000, 5000, 10000, 20000):
        for k in (8, 16, 32, 64):
            n_frames = max(1, min(64, args.points_budget // n))
            base = [synth.make_frame(7000 + i, n, knn=k)[0] for i in range(min(2, n_frames))]
            frames = [base[i % len(base)] for i in range(n_frames)]
            pts, fp = gf.frames_to_device(frames, dev)

            def step():
                bf = gf.build_graph_batch(pts, fp, 25, k, max_range=R, max_azimuth=np.pi * 0.5)
                cl = torch.arange(bf.gb.n_nodes, device=dev, dtype=torch.int32)
                bf.gb.cl_ptr = torch.arange(bf.gb.n_nodes + 1, device=dev, dtype=torch.int32)      # every node its own cluster
                bf.gb.cl_members, bf.gb.n_clusters = cl, bf.gb.n_nodes
                with torch.no_grad():
                    out = det.forward_batch(bf.gb, bf.node_features, bf.edge_features, training=False)
                return bf, out
            bf, out = step()
            torch.cuda.synchronize()
            ok = None
            if n <= 5000:
                from oracle import graph_np
                adj = graph_np.adjacency_information(base[0], 25, k)
                e0 = int(bf.gb.row_ptr[fp[1]].item())
                ok = bool(np.array_equal(bf.edge_index().cpu().numpy()[:, :e0], adj['adj_list']))
            e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0_.record()
            for _ in range(args.reps):
                step()
            e1_.record()
            torch.cuda.synchronize()
            ms = e0_.elapsed_time(e1_) / args.reps
            finite = all(bool(torch.isfinite(o).all()) for o in out)
            print(json.dumps({'points_per_frame': n, 'knn': k, 'frames': n_frames, 'edges': bf.gb.n_edges,
                              'edges_per_node': round(bf.gb.n_edges / bf.gb.n_nodes, 2), 'ms_graph_plus_forward': round(ms, 3),
                              'frames_per_s': round(n_frames / (ms * 1e-3), 1), 'edges_per_s': round(bf.gb.n_edges / (ms * 1e-3)),
                              'edge_index_matches_oracle': ok, 'outputs_finite': finite}), flush=True)
            del bf, out
            torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
