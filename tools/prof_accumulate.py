#!/usr/bin/env python
"""Row f3 measurement: the window accumulation pre-pass for a whole bench batch (256 windows x 10 scans, ~17.6 k raw
detections per window so that ~3 000 dynamic detections survive the region-of-interest filter) on the GPU
(rgnn_accumulate_windows, raw detections resident in HBM) next to the oracle restatement of the reference's
per-scan NumPy loop on the host for a 4-window sample.  Prints one JSON line (HBM bytes: 39 B read twice per raw
detection + 45 B written per kept one)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from graph_neural_network_for_radar_perception_b200 import accumulate as acc, synth  # noqa: E402


def main():
    n_windows, pps = 256, 1760
    dev = torch.device('cuda:0')
    base = [synth.make_raw_window(i, 10, pps) for i in range(8)]
    windows = [base[i % 8] for i in range(n_windows)]
    raw, scan_of, params, raw_ptr = acc._pack_windows(windows)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    f = {k: up(raw[k].astype(np.float32)) for k in acc._RAW_F32}
    ts, lab = up(raw['timestamp'].astype(np.int64)), up(raw['label_id'].astype(np.uint8))
    trk, so, pr = up((raw['track_id'] != b'').astype(np.uint8)), up(scan_of), up(params)
    fn = lambda: acc.accumulate_windows_device(f, ts, lab, trk, so, pr, raw_ptr)
    res = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        res = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    n_raw, n_kept = int(raw.shape[0]), res.frame_ptr[-1]
    # host: oracle on 4 windows (per-scan NumPy, as the reference does it)
    from oracle import accumulate_np as o
    t0 = time.perf_counter()
    for w in base[:4]:
        d = o.accumulate_window(*w, exact_dgemm=False)
        o.select(d, o.class_labels(d))
    t_cpu = (time.perf_counter() - t0) / 4
    hbm = 2 * 39.0 * n_raw + 45.0 * n_kept
    print(json.dumps({'windows': n_windows, 'raw_detections': n_raw, 'kept_detections': n_kept,
                      'gpu_ms_incl_frame_ptr_readback': round(ms, 3), 'windows_per_s_gpu': round(n_windows / (ms * 1e-3)),
                      'algorithmic_GBps': round(hbm / (ms * 1e-3) / 1e9, 1),
                      'cpu_oracle_ms_per_window': round(t_cpu * 1e3, 2), 'windows_per_s_cpu_1core': round(1.0 / t_cpu, 1)}))


if __name__ == '__main__':
    main()
