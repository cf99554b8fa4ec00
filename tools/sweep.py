#!/usr/bin/env python
"""BASELINE.json configs[4]: edge-count stress sweep -- points per frame x kNN degree x hidden width.
Builds the graph and runs the detector forward on the GPU for every point of the sweep; checks the edge list of the first
frame against the oracle where that finishes in seconds (N <= 5000) and, for N <= 2000, the FOUR OUTPUTS of the first frame
against the float32 oracle (max |got - want| / (1e-4 |want| + 1e-5 max|want|), <= 1 passes); prints one JSON line per point.
Hidden width 64 = the reference channel plan with the reference checkpoint; 32 / 128 / 256 = random-init weights (seed 1234)
with graph_convolution_stem_channels, the encoder tails and the head stems scaled and msg_mlp_hidden_dim = 2 x hidden.

    python tools/sweep.py [--points-budget 400000] [--hidden 32,64,128,256] > profiles/sweep_rNN.jsonl
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
    ap.add_argument('--hidden', default='64,32,128,256')
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    torch.cuda.set_device(dev)
    R = np.float64(np.sqrt(100.0 ** 2 + 50.0 ** 2))
    for hidden in [int(h) for h in args.hidden.split(',')]:
        cfg = config()
        if hidden != 64:
            cfg.node_feat_enc_stem_channels = [256, 128, hidden]
            cfg.edge_feat_enc_stem_channels = [256, 128, 128, hidden]
            cfg.graph_convolution_stem_channels = [hidden] * 7
            cfg.msg_mlp_hidden_dim = 2 * hidden
            cfg.link_pred_stem_channels = [hidden] * 3
            cfg.node_pred_stem_channels = [hidden] * 3
        torch.manual_seed(1234)
        model = Model_Training(cfg, dev)
        if hidden == 64:
            model.load_state_dict(torch.load(CKPT, map_location='cpu', weights_only=True))
        det = model.to(dev).pred.eval()
        sd_cpu = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        sizes = (500, 1000, 2000, 5000, 10000, 20000) if hidden == 64 else (500, 2000, 10000)
        degrees = (8, 16, 32, 64) if hidden == 64 else (8, 32)
        sweep_one(args, dev, det, sd_cpu, hidden, sizes, degrees, R)


def sweep_one(args, dev, det, sd_cpu, hidden, sizes, degrees, R):
    for n in sizes:
        for k in degrees:
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
            out_ratio = None
            if n <= 2000:       # the four outputs of frame 0 against the float32 oracle
                from oracle import graph_np, model_torch as mt
                adj = graph_np.adjacency_information(base[0], 25, k)
                nf = torch.from_numpy(graph_np.node_features(base[0], adj['degree'], True, 0, R, 0, np.pi * 0.5).astype(np.float32))
                ef = torch.from_numpy(graph_np.edge_features(base[0], adj['adj_list']).astype(np.float32))
                ei = torch.from_numpy(adj['adj_list'])
                with torch.no_grad():
                    want = mt.detector_forward(sd_cpu, nf, ef, ei, [torch.tensor([i]) for i in range(n)])
                n_und0 = int((ei[0] < ei[1]).sum())
                got0 = (out[0][:n], out[1][:n], out[2][:n_und0], out[3][:n])
                out_ratio = 0.0
                for g_, w_ in zip(got0, want):
                    w_ = w_.double()
                    bound = 1e-4 * w_.abs() + 1e-5 * max(float(w_.abs().max()), 1.0)
                    out_ratio = max(out_ratio, float(((g_.double().cpu() - w_).abs() / bound).max()))
            e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0_.record()
            for _ in range(args.reps):
                step()
            e1_.record()
            torch.cuda.synchronize()
            ms = e0_.elapsed_time(e1_) / args.reps
            finite = all(bool(torch.isfinite(o).all()) for o in out)
            print(json.dumps({'hidden': hidden, 'points_per_frame': n, 'knn': k, 'frames': n_frames, 'edges': bf.gb.n_edges,
                              'edges_per_node': round(bf.gb.n_edges / bf.gb.n_nodes, 2), 'ms_graph_plus_forward': round(ms, 3),
                              'frames_per_s': round(n_frames / (ms * 1e-3), 1), 'edges_per_s': round(bf.gb.n_edges / (ms * 1e-3)),
                              'edge_index_matches_oracle': ok, 'outputs_vs_oracle_ratio': None if out_ratio is None else round(out_ratio, 3),
                              'outputs_finite': finite}), flush=True)
            del bf, out
            torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
