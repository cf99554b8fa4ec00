#!/usr/bin/env python
"""Row f2 measurement: proposal clustering of a whole bench batch (256 frames x 3000 points) on the GPU
(rgnn_cluster_links / rgnn_cluster_radius) next to the oracle restatement of the reference's Simple_DBSCAN on the host
cores for a 4-frame sample.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from graph_neural_network_for_radar_perception_b200 import clustering as cl, synth  # noqa: E402
from graph_neural_network_for_radar_perception_b200 import graph_features as gf  # noqa: E402


def main():
    n_frames, n = 256, 3000
    dev = torch.device('cuda:0')
    base = [synth.make_frame(i, n, knn=10)[0] for i in range(8)]
    frames = [base[i % 8] for i in range(n_frames)]
    pts, fp = gf.frames_to_device(frames, dev)
    bf = gf.build_graph_batch(pts, fp, 25, 10, max_range=np.float64(np.sqrt(100.0 ** 2 + 50.0 ** 2)), max_azimuth=np.pi * 0.5)
    gb = bf.gb
    g = torch.Generator(device=dev).manual_seed(0)
    xy = torch.stack((pts['meas_px'], pts['meas_py']), dim=1).float() + 0.3 * torch.randn(gb.n_nodes, 2, device=dev, generator=g)
    logits = torch.randn(gb.n_und, 2, device=dev, generator=g)

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, r
    ms_links, r1 = timed(lambda: cl.cluster_links(xy, gb.und_a, gb.und_b, logits, gb.n_und, 1.4))
    ms_radius, r2 = timed(lambda: cl.cluster_radius(xy, 1.4, fp))
    # host: oracle on 4 frames (links mode needs the pair list of those frames)
    from oracle.clustering_np import Simple_DBSCAN as Oracle
    und_a, und_b = gb.und_a[:gb.n_und].cpu().numpy(), gb.und_b[:gb.n_und].cpu().numpy()
    xy_h, lg = xy.cpu().numpy(), logits.cpu().numpy()
    t0 = time.perf_counter()
    for f in range(4):
        sel = (und_a >= fp[f]) & (und_a < fp[f + 1])
        o = Oracle(1.4, True)
        o.cluster_nodes(xy_h[fp[f]:fp[f + 1]], (lg[sel, 1] > lg[sel, 0]).astype(np.int64), und_pairs=np.stack([und_a[sel] - fp[f], und_b[sel] - fp[f]]))
    t_links = (time.perf_counter() - t0) / 4
    t0 = time.perf_counter()
    for f in range(4):
        o = Oracle(1.4, False)
        o.cluster_nodes(xy_h[fp[f]:fp[f + 1]])
    t_radius = (time.perf_counter() - t0) / 4
    print(json.dumps({'frames': n_frames, 'nodes': gb.n_nodes, 'undirected_pairs': gb.n_und,
                      'gpu_ms_links_mode': round(ms_links, 3), 'gpu_ms_radius_mode': round(ms_radius, 3),
                      'clusters_links': r1.n_clusters, 'clusters_radius': r2.n_clusters,
                      'gpu_frames_per_s_links': round(n_frames / (ms_links * 1e-3)), 'gpu_frames_per_s_radius': round(n_frames / (ms_radius * 1e-3)),
                      'cpu_oracle_ms_per_frame_links': round(t_links * 1e3, 2), 'cpu_oracle_ms_per_frame_radius': round(t_radius * 1e3, 2),
                      'note': 'GPU times include the device->host read of the cluster count; the reference itself runs an O(N^2) Python BFS per frame'}))


if __name__ == '__main__':
    main()
