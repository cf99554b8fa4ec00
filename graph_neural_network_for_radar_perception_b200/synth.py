"""Deterministic synthetic RadarScenes-shaped frames (host side, NumPy).

The RadarScenes `radar_data.h5` blobs are not available offline, so every parity test and every
benchmark in this repo runs on frames produced here.  The recipe follows SURVEY.md section 8(d):

* ROI 100 m x 100 m (`x in [0,100)`, `y in [-50,50)`), the grid limits of the reference yml
  (reference configuration_radarscenes_gnn.yml:33-38).
* 30 % of the points are re-drawn as ceil(N/60) Gaussian blobs (sigma ~ U[0.5,2] m) to mimic objects.
* `vr ~ N(0,3^2)`, `rcs ~ N(-5,10^2)`, `vx = vr cos(th)`, `vy = vr sin(th)`, `th = atan2(py,px)`.
* timestamps are int64 microseconds inside a 10-scan window
  (reference modules/data_utils/read_data.py accumulates 10 scans; yml:12).
* labels mirror what reference modules/data_generator/datagen_gnn.py:15-45,126-139 produces:
  node class, node offsets to the blob centre, undirected-edge link labels, clusters = blobs + singletons.

A frame is rejected and re-drawn (next sub-seed) if any row of the squared-distance matrix has a tie
across the k-th neighbour slot, because the reference's unstable argsort leaves that case undefined
(SURVEY.md "Hard parts", appendix C item 2).
"""
from __future__ import annotations

import numpy as np

ROI_MIN_X, ROI_MAX_X = 0.0, 100.0
ROI_MIN_Y, ROI_MAX_Y = -50.0, 50.0
T0_US = 29_727_354_614_437
SCAN_US = 16_500


def _draw(rng: np.random.Generator, n: int):
    px = rng.uniform(ROI_MIN_X, ROI_MAX_X, n)
    py = rng.uniform(ROI_MIN_Y, ROI_MAX_Y, n)
    n_blob = int(np.ceil(n / 60))
    n_in_blob = int(0.3 * n)
    blob_of = np.full(n, -1, dtype=np.int64)
    if n_blob > 0 and n_in_blob > 0:
        cx = rng.uniform(ROI_MIN_X + 3, ROI_MAX_X - 3, n_blob)
        cy = rng.uniform(ROI_MIN_Y + 3, ROI_MAX_Y - 3, n_blob)
        sg = rng.uniform(0.5, 2.0, n_blob)
        members = rng.choice(n, n_in_blob, replace=False)
        which = rng.integers(0, n_blob, n_in_blob)
        px[members] = np.clip(cx[which] + sg[which] * rng.standard_normal(n_in_blob), ROI_MIN_X, ROI_MAX_X - 1e-3)
        py[members] = np.clip(cy[which] + sg[which] * rng.standard_normal(n_in_blob), ROI_MIN_Y, ROI_MAX_Y - 1e-3)
        blob_of[members] = which
    px = px.astype(np.float32)
    py = py.astype(np.float32)
    vr = (3.0 * rng.standard_normal(n)).astype(np.float32)
    rcs = (-5.0 + 10.0 * rng.standard_normal(n)).astype(np.float32)
    th = np.arctan2(py, px)
    vx = (vr * np.cos(th)).astype(np.float32)
    vy = (vr * np.sin(th)).astype(np.float32)
    scan = rng.integers(0, 10, n)
    ts = (T0_US + scan * SCAN_US + rng.integers(0, 1000, n)).astype(np.int64)
    node_class = rng.integers(0, 7, n).astype(np.int64)
    return dict(meas_px=px, meas_py=py, meas_vx=vx, meas_vy=vy, meas_vr=vr, meas_rcs=rcs,
                meas_timestamp=ts), blob_of, node_class


def _has_knn_tie(px: np.ndarray, py: np.ndarray, k: int) -> bool:
    """True if some row has d2[k-th] == d2[(k+1)-th] (self included), in the reference's f32 arithmetic."""
    n = px.shape[0]
    if k + 1 >= n:
        return False
    blk = 2048
    for s in range(0, n, blk):
        dx = px[s:s + blk, None] - px[None, :]
        dy = py[s:s + blk, None] - py[None, :]
        d2 = (dx * dx).astype(np.float32) + (dy * dy).astype(np.float32)
        part = np.partition(d2, (k, k + 1), axis=1)
        if np.any(part[:, k] == part[:, k + 1]):
            return True
    return False


def make_frame(frame_idx: int, n_points: int, knn: int = 10, seed: int = 1234, with_labels: bool = True):
    """Return (data_dict, labels_src) for one frame; labels_src holds blob ids and node classes.

    `data_dict` has the keys the reference's graph functions read
    (reference modules/compute_features/graph_features.py:70,133-138,153-161).
    """
    attempt = 0
    while True:
        rng = np.random.default_rng([seed + frame_idx, attempt])
        data, blob_of, node_class = _draw(rng, n_points)
        if not _has_knn_tie(data['meas_px'], data['meas_py'], knn):
            break
        attempt += 1
        if attempt > 50:
            raise RuntimeError('could not draw a tie-free frame')
    return data, dict(blob_of=blob_of, node_class=node_class)


def make_labels(data: dict, labels_src: dict, adj_list: np.ndarray):
    """Build the label set of reference datagen_gnn.py:126-139 for a synthetic frame.

    adj_list is the (2,E) reference-order edge list; undirected edges are those with row<col, in
    row-major order (reference compute_edge_labels.py:17-19, gnn_blocks.py:295-296).
    Returns dict with numpy arrays: edge_class (E_u,) i64, node_class (N,) i64, node_offsets (N,2) f32,
    cluster_node_idx (list of i64 arrays), cluster_labels (C,) i64.
    """
    blob_of = labels_src['blob_of']
    node_class = labels_src['node_class'].copy()
    n = blob_of.shape[0]
    px, py = data['meas_px'], data['meas_py']
    offsets = np.zeros((n, 2), dtype=np.float32)
    clusters, cluster_labels = [], []
    for b in np.unique(blob_of[blob_of >= 0]):
        idx = np.nonzero(blob_of == b)[0]
        cx, cy = px[idx].mean(dtype=np.float64), py[idx].mean(dtype=np.float64)
        offsets[idx, 0] = (cx - px[idx]).astype(np.float32)
        offsets[idx, 1] = (cy - py[idx]).astype(np.float32)
        node_class[idx] = node_class[idx[0]]          # one class per object
        clusters.append(idx.astype(np.int64))
        cluster_labels.append(node_class[idx[0]])
    for i in np.nonzero(blob_of < 0)[0]:
        clusters.append(np.array([i], dtype=np.int64))
        cluster_labels.append(node_class[i])
    r, c = adj_list[0], adj_list[1]
    und = r < c
    same = (blob_of[r[und]] == blob_of[c[und]]) & (blob_of[r[und]] >= 0)
    return dict(edge_class=same.astype(np.int64), node_class=node_class,
                node_offsets=offsets, cluster_node_idx=clusters,
                cluster_labels=np.asarray(cluster_labels, dtype=np.int64))


# ---------------------------------------------------------------------------------------------------------------
# Raw sliding-window input: what the reference's extract_and_sync_radar_data
# (modules/data_utils/read_data.py:227-303) reads -- RadarScenes-shaped structured arrays `radar_data`, `odometry`,
# the sensors.json mount dict and one window of create_dataset_sliding_window (read_data.py:202-224).
# ---------------------------------------------------------------------------------------------------------------
RADAR_DTYPE = np.dtype([('timestamp', '<i8'), ('sensor_id', 'u1'), ('range_sc', '<f4'), ('azimuth_sc', '<f4'),
                        ('rcs', '<f4'), ('vr', '<f4'), ('vr_compensated', '<f4'), ('x_cc', '<f4'), ('y_cc', '<f4'),
                        ('x_seq', '<f4'), ('y_seq', '<f4'), ('uuid', 'S32'), ('track_id', 'S32'), ('label_id', 'u1')])
ODOMETRY_DTYPE = np.dtype([('timestamp', '<i8'), ('x_seq', '<f4'), ('y_seq', '<f4'), ('yaw_seq', '<f4'),
                           ('vx', '<f4'), ('yaw_rate', '<f4')])
RADAR_MOUNTS = {'radar_1': {'id': 1, 'x': 3.663, 'y': -0.873, 'yaw': -1.48418552},
                'radar_2': {'id': 2, 'x': 3.86, 'y': -0.7, 'yaw': -0.436185662},
                'radar_3': {'id': 3, 'x': 3.86, 'y': 0.7, 'yaw': 0.436},
                'radar_4': {'id': 4, 'x': 3.663, 'y': 0.873, 'yaw': 1.484}}


def make_raw_window(window_idx: int, n_scans: int = 10, points_per_scan: int = 300, seed: int = 1234,
                    dynamic_fraction: float = 0.35):
    """One synthetic sliding window: (radar_mount_data, radar_data_all_scenes, odometry_data_all_scenes,
    windowed_data), the four arguments of the reference's extract_and_sync_radar_data.

    The ego vehicle drives a gentle arc (vx ~ 8..15 m/s, |yaw rate| < 0.15 rad/s); scans cycle through the four
    radars every ~16.5 ms; stationary detections carry the range rate the ego motion predicts plus N(0, 0.3^2)
    noise, moving ones (tracked objects of random old-label class, or clutter without a track) an arbitrary one;
    detections are spread over x_cc in [-20, 120) and y_cc in [-70, 70) so that the 100 m x 100 m region-of-interest
    filter has something to drop."""
    rng = np.random.default_rng([seed, 7919, window_idx])
    vx0, w0 = rng.uniform(8.0, 15.0), rng.uniform(-0.15, 0.15)
    odo = np.zeros(n_scans, dtype=ODOMETRY_DTYPE)
    t = T0_US + np.arange(n_scans, dtype=np.int64) * SCAN_US + rng.integers(0, 400, n_scans)
    dt = (t - t[0]) * 1e-6
    yaw = 0.3 + w0 * dt
    odo['timestamp'] = t
    odo['yaw_seq'] = yaw
    odo['x_seq'] = 120.0 + vx0 * dt * np.cos(yaw)
    odo['y_seq'] = -40.0 + vx0 * dt * np.sin(yaw)
    odo['vx'] = vx0 + 0.05 * rng.standard_normal(n_scans)
    odo['yaw_rate'] = w0 + 0.005 * rng.standard_normal(n_scans)
    counts = rng.integers(max(points_per_scan // 2, 1), points_per_scan * 3 // 2 + 1, n_scans)
    bounds = np.concatenate([[0], np.cumsum(counts)])
    rad = np.zeros(int(bounds[-1]), dtype=RADAR_DTYPE)
    radar_ids = [int(1 + (window_idx + s) % 4) for s in range(n_scans)]
    n_tracks = max(2, points_per_scan // 40)
    track_names = np.array([('trk%03d_%05d' % (k, window_idx)).encode() for k in range(n_tracks)], dtype='S32')
    track_class = rng.integers(0, 11, n_tracks)
    for s in range(n_scans):
        a, b = int(bounds[s]), int(bounds[s + 1])
        n = b - a
        m = RADAR_MOUNTS['radar_%d' % radar_ids[s]]
        x = rng.uniform(-20.0, 120.0, n).astype(np.float32)
        y = rng.uniform(-70.0, 70.0, n).astype(np.float32)
        # sensor-frame azimuth of the detection
        xs, ys = x - m['x'], y - m['y']
        az = (np.arctan2(ys, xs) - m['yaw']).astype(np.float32)
        vxs = float(odo['vx'][s]) - float(odo['yaw_rate'][s]) * m['y']
        vys = float(odo['yaw_rate'][s]) * m['x']
        c, sn = np.cos(-m['yaw']), np.sin(-m['yaw'])
        vxs, vys = vxs * c - vys * sn, vxs * sn + vys * c
        pred = -(vxs * np.cos(az.astype(np.float64)) + vys * np.sin(az.astype(np.float64)))
        moving = rng.random(n) < dynamic_fraction
        vr = np.where(moving, pred + rng.uniform(-12.0, 12.0, n), pred + 0.3 * rng.standard_normal(n))
        has_track = moving & (rng.random(n) < 0.7)
        which = rng.integers(0, n_tracks, n)
        r = rad[a:b]
        r['timestamp'] = t[s] + rng.integers(0, 200, n)
        r['sensor_id'] = radar_ids[s]
        r['range_sc'] = np.hypot(xs, ys)
        r['azimuth_sc'] = az
        r['rcs'] = -5.0 + 10.0 * rng.standard_normal(n)
        r['vr'] = vr
        r['vr_compensated'] = vr - pred
        r['x_cc'], r['y_cc'] = x, y
        r['track_id'] = np.where(has_track, track_names[which], np.array(b'', dtype='S32'))
        r['label_id'] = np.where(has_track, track_class[which], 11)
    windowed = {'current_timestamps': [int(v) for v in t], 'radar_id': radar_ids,
                'odometry_timestamp': [int(v) for v in t], 'odometry_index': list(range(n_scans)),
                'radar_data_indices': [[int(bounds[s]), int(bounds[s + 1])] for s in range(n_scans)]}
    return RADAR_MOUNTS, rad, odo, windowed
