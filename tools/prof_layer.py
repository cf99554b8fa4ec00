#!/usr/bin/env python
"""Profiling driver: one message-passing layer (residual_graph_conv_block forward, optionally the detector
forward / training step) on a synthetic batch, a fixed number of times, nothing else.  Meant to be run under
`ncu` on the GPU box (see profiles/README.md); it prints CUDA-event timings when run plain.

    python tools/prof_layer.py [--frames 32] [--points 3000] [--reps 3] [--what layer|forward|train]
"""
from __future__ import annotations

import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from graph_neural_network_for_radar_perception_b200 import config, Model_Training, synth  # noqa: E402
from graph_neural_network_for_radar_perception_b200 import graph_features as gf  # noqa: E402
from graph_neural_network_for_radar_perception_b200._cabi import check, lib, ptr, stream_ptr  # noqa: E402
from graph_neural_network_for_radar_perception_b200._engine import detector_table  # noqa: E402

CKPT = os.path.join(ROOT, 'tests', 'golden', 'graph_based_detector.pt')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=32)
    ap.add_argument('--points', type=int, default=3000)
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--what', default='layer')
    ap.add_argument('--opt', action='append', default=[], help='name=value for rgnn_set_option')
    args = ap.parse_args()
    for o in args.opt:
        k, v = o.split('=')
        check(lib().rgnn_set_option(k.encode(), int(v)), 'set_option')
    dev = torch.device('cuda:0')
    torch.cuda.set_device(dev)
    model = Model_Training(config(), dev)
    model.load_state_dict(torch.load(CKPT, map_location='cpu', weights_only=True))
    model = model.to(dev)
    det = model.pred.eval()
    base = [synth.make_frame(i, args.points, knn=10) for i in range(min(8, args.frames))]
    frames = [base[i % len(base)][0] for i in range(args.frames)]
    pts, fp = gf.frames_to_device(frames, dev)
    bf = gf.build_graph_batch(pts, fp, 25, 10, max_range=np.float64(np.sqrt(100.0 ** 2 + 50.0 ** 2)),
                              max_azimuth=np.pi * 0.5)
    gb = bf.gb
    N, E = gb.n_nodes, gb.n_edges
    print(f'nodes {N} edges {E} undirected {gb.n_und}', flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.what == 'layer':
        table = detector_table(det)
        table.refill(None)
        check(lib().rgnn_pack_detector(C.byref(table.det), stream_ptr()), 'pack')
        x = torch.randn(N, 64, device=dev)
        e = torch.randn(E, 64, device=dev)
        out, agg = torch.empty_like(x), torch.empty_like(x)
        proj = torch.empty(N, 256, device=dev)
        g = gb.c_struct()
        conv = table.det.conv[0]
        s = stream_ptr()
        fn = lambda: check(lib().rgnn_conv_block_fwd(C.byref(conv), C.byref(g), ptr(x), ptr(e), ptr(out), ptr(agg),
                                                     ptr(proj), s), 'conv')
    elif args.what == 'edges':      # the message kernel alone (its launch also zero-fills agg)
        table = detector_table(det)
        table.refill(None)
        check(lib().rgnn_pack_detector(C.byref(table.det), stream_ptr()), 'pack')
        e = torch.randn(E, 64, device=dev)
        agg = torch.empty(N, 64, device=dev)
        proj = torch.randn(N, 256, device=dev)
        g = gb.c_struct()
        conv = table.det.conv[0]
        s = stream_ptr()
        fn = lambda: check(lib().rgnn_conv_edges_fwd(C.byref(conv), C.byref(g), ptr(e), ptr(proj), ptr(agg), s), 'edges')
    elif args.what == 'edges16':    # the fp16-split message kernel alone on pre-split rows (its launch also zero-fills agg)
        table = detector_table(det)
        table.refill(None)
        check(lib().rgnn_pack_detector(C.byref(table.det), stream_ptr()), 'pack')
        e = torch.randn(E, 64, device=dev)
        es = torch.empty(lib().rgnn_split_edge_embedding_words(E), dtype=torch.int32, device=dev)
        check(lib().rgnn_split_edge_embedding(ptr(e), E, ptr(es), stream_ptr()), 'split')
        agg = torch.empty(N, 64, device=dev)
        proj = torch.randn(N, 256, device=dev)
        g = gb.c_struct()
        conv = table.det.conv[0]
        s = stream_ptr()
        fn = lambda: check(lib().rgnn_conv_edges_f16_fwd(C.byref(conv), C.byref(g), ptr(es), ptr(proj), ptr(agg), s), 'edges16')
    elif args.what == 'forward':
        gb.set_clusters([[torch.arange(i, min(i + 4, n)) for i in range(0, n, 4)] for n in np.diff(fp)], fp[:-1], dev)

        def fn():
            with torch.no_grad():
                det.forward_batch(gb, bf.node_features, bf.edge_features, training=False)
    elif args.what == 'train':
        from graph_neural_network_for_radar_perception_b200.training import DataParallelTrainer
        model.train()
        trainer = DataParallelTrainer(model)
        ei = bf.edge_index().cpu().numpy()
        row_ptr = gb.row_ptr.cpu().numpy()
        labs, cl_lists = [], []
        for i in range(args.frames):
            d, src = base[i % len(base)]
            ea, eb = int(row_ptr[fp[i]]), int(row_ptr[fp[i + 1]])
            lab = synth.make_labels(d, src, ei[:, ea:eb] - fp[i])
            labs.append(lab)
            cl_lists.append([torch.from_numpy(c) for c in lab['cluster_node_idx']])
        gb.set_clusters(cl_lists, fp[:-1], dev)
        labels = {k: torch.cat([torch.from_numpy(l[k]) for l in labs]).to(dev)
                  for k in ('cluster_labels', 'edge_class', 'node_class', 'node_offsets')}
        fn = lambda: trainer.step(gb, bf.node_features, bf.edge_features, labels)
    else:
        raise SystemExit('unknown --what')
    fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f'{args.what}: {e0.elapsed_time(e1) / args.reps:.3f} ms per rep', flush=True)


if __name__ == '__main__':
    main()
