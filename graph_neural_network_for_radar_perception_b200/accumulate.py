"""Sliding-window accumulation pre-pass on the device: from the raw detections of the
scans of a window to the `meas_*` arrays the graph construction reads, for many windows per launch.

Reference interface mirrored here (same argument meaning, NumPy in / NumPy out for the single-window functions):

    get_data_for_datagen(...)            reference modules/data_utils/read_data.py:489-537 (minus the file reading:
                                         radar_mount_data / radar_data / odometry arrays are passed in, as
                                         extract_and_sync_radar_data :227-303 receives them)
    select_meas_within_the_grid(...)     reference modules/compute_features/grid_features.py:162-173
    select_moving_data(...)              reference modules/compute_features/graph_features.py:167-182

and the batched device-resident entry `accumulate_windows`, whose output plugs straight into
`graph_features.build_graph_batch` (points dict + frame_ptr).  Per-scan scalars (relative pose, sensor velocity) are
NumPy float64 host math written exactly as the reference writes it; everything per detection runs in
csrc/rgnn_accumulate.cu through `rgnn_accumulate_windows`.  There is no CPU path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from ._cabi import check, lib, ptr, stream_ptr
from .graph_features import _device

LABELS_TO_ID = {'CAR': 0, 'PEDESTRIAN': 1, 'PEDESTRIAN_GROUP': 2, 'TWO_WHEELER': 3, 'LARGE_VEHICLE': 4, 'NONE': 5,
                'FALSE': 6, 'STATIC': 7}            # reference modules/data_utils/labels.py:60-70
_RAW_F32 = ('x_cc', 'y_cc', 'vr', 'vr_compensated', 'azimuth_sc', 'rcs')
_OUT_KEYS = ('meas_px', 'meas_py', 'meas_vx', 'meas_vy', 'meas_vr', 'meas_rcs')


def _se2(px, py, theta):
    """reference modules/data_utils/meas_sync.py:24-33"""
    T = np.eye(3)
    T[:2, :2] = np.array([[np.cos(theta), -np.sin(theta)], [np.sin(theta), np.cos(theta)]])
    T[:2, 2:] = np.array([[px], [py]])
    return T


def scan_parameters(radar_mount_data: Dict, odometry_data_all_scenes: np.ndarray, windowed_data: Dict) -> np.ndarray:
    """(n_scans, 9) float64 rows [R00 R01 R10 R11 tx ty mount_yaw vx_sensor vy_sensor] of one window.
    [R|t] = inv(T_curr) @ T_prev with T_curr the pose of the window's last scan (meas_sync.py:55,88-96); the sensor
    velocity is generate_meas_sensor_frame with vy_ego = 0 (meas_selection.py:11-37,184)."""
    odo = [odometry_data_all_scenes[i] for i in windowed_data['odometry_index']]
    poses = [_se2(o['x_seq'], o['y_seq'], o['yaw_seq']) for o in odo]
    inv_curr = np.linalg.inv(poses[-1])
    out = np.zeros((len(odo), 9), dtype=np.float64)
    for s, (radar_id, o, T_prev) in enumerate(zip(windowed_data['radar_id'], odo, poses)):
        T = inv_curr @ T_prev
        mount = radar_mount_data['radar_' + str(radar_id)]
        theta = -mount['yaw']
        vx_sensor = o['vx'] - o['yaw_rate'] * mount['y']
        vy_sensor = 0.0 + o['yaw_rate'] * mount['x']
        out[s, 0:4] = T[:2, :2].reshape(4)
        out[s, 4:6] = T[:2, 2]
        out[s, 6] = mount['yaw']
        out[s, 7] = vx_sensor * np.cos(theta) - vy_sensor * np.sin(theta)
        out[s, 8] = vx_sensor * np.sin(theta) + vy_sensor * np.cos(theta)
    return out


class AccumulatedWindows:
    """Device-resident result of accumulate_windows."""
    points: Dict[str, torch.Tensor]      # meas_px .. meas_rcs (f32), meas_timestamp (i64): input of build_graph_batch
    frame_ptr: List[int]                 # host offsets of the windows in `points` (n_windows + 1)
    class_labels: torch.Tensor           # (n,) f32 new-label ids (labels.py:60-70)
    src_index: torch.Tensor              # (n,) i32 index of each kept detection inside its window
    stationary_flag: torch.Tensor        # (n_raw,) u8 per raw detection
    raw_ptr: List[int]                   # host offsets of the windows in the raw detections


def _pack_windows(windows: Sequence[Tuple[Dict, np.ndarray, np.ndarray, Dict]]):
    """Concatenate the scans of every window (in window order) into flat host arrays + the scan parameter table."""
    parts, scan_of, params, raw_ptr = [], [], [], [0]
    n_scans = 0
    for mounts, radar, odo, win in windows:
        params.append(scan_parameters(mounts, odo, win))
        n = 0
        for s, (a, b) in enumerate(win['radar_data_indices']):
            parts.append(radar[a:b])
            scan_of.append(np.full(b - a, n_scans + s, dtype=np.int32))
            n += b - a
        n_scans += len(win['radar_data_indices'])
        raw_ptr.append(raw_ptr[-1] + n)
    return np.concatenate(parts), np.concatenate(scan_of), np.concatenate(params), raw_ptr


def accumulate_windows(windows: Sequence[Tuple[Dict, np.ndarray, np.ndarray, Dict]],
                       flip_along_x: Optional[Sequence[bool]] = None, select: bool = True,
                       min_x=0, max_x=100, min_y=-50, max_y=50, device=None) -> AccumulatedWindows:
    """windows: sequence of (radar_mount_data, radar_data_all_scenes, odometry_data_all_scenes, windowed_data), the
    arguments of the reference's extract_and_sync_radar_data.  One H2D copy per field, one launch set for all windows."""
    dev = _device(device)
    if len(windows) == 0:
        z = lambda dt: torch.zeros(0, dtype=dt, device=dev)
        return accumulate_windows_device({k: z(torch.float32) for k in _RAW_F32}, z(torch.int64), z(torch.uint8), z(torch.uint8),
                                         z(torch.int32), torch.zeros((0, 9), dtype=torch.float64, device=dev), [0], None,
                                         select, min_x, max_x, min_y, max_y)
    raw, scan_of, params, raw_ptr = _pack_windows(windows)
    n, nw = int(raw.shape[0]), len(windows)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=True)
    f = {k: up(raw[k].astype(np.float32, copy=False)) for k in _RAW_F32}
    ts = up(raw['timestamp'].astype(np.int64, copy=False))
    lab = up(raw['label_id'].astype(np.uint8, copy=False))
    trk = up((raw['track_id'] != b'').astype(np.uint8))
    return accumulate_windows_device(f, ts, lab, trk, up(scan_of), up(params), raw_ptr, flip_along_x, select,
                                     min_x, max_x, min_y, max_y)


def accumulate_windows_device(fields: Dict[str, torch.Tensor], timestamp: torch.Tensor, label_id: torch.Tensor,
                              has_track: torch.Tensor, point_scan: torch.Tensor, scan_params: torch.Tensor,
                              raw_ptr: Sequence[int], flip_along_x: Optional[Sequence[bool]] = None, select: bool = True,
                              min_x=0, max_x=100, min_y=-50, max_y=50) -> AccumulatedWindows:
    """Same as accumulate_windows with the raw detections already resident in HBM (benchmarks, streaming callers)."""
    dev = timestamp.device
    if dev.type != 'cuda':
        from ._cabi import RgnnError
        raise RgnnError('accumulate_windows: tensors must live on a CUDA device (there is no CPU path)')
    n, nw = int(timestamp.shape[0]), len(raw_ptr) - 1
    wp = torch.tensor(list(raw_ptr), dtype=torch.int32, device=dev)
    flip = None
    if flip_along_x is not None:
        flip = torch.tensor([1 if v else 0 for v in flip_along_x], dtype=torch.uint8, device=dev)
    cap = max(n, 1)
    out = {k: torch.empty(cap, dtype=torch.float32, device=dev) for k in _OUT_KEYS}
    ts_out = torch.empty(cap, dtype=torch.int64, device=dev)
    labels = torch.empty(cap, dtype=torch.float32, device=dev)
    src = torch.empty(cap, dtype=torch.int32, device=dev)
    stat = torch.empty(cap, dtype=torch.uint8, device=dev)
    out_ptr = torch.empty(nw + 1, dtype=torch.int32, device=dev)
    nbytes = lib().rgnn_accumulate_workspace_bytes(nw)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    check(lib().rgnn_accumulate_windows(
        ptr(fields['x_cc']), ptr(fields['y_cc']), ptr(fields['vr']), ptr(fields['vr_compensated']),
        ptr(fields['azimuth_sc']), ptr(fields['rcs']), ptr(timestamp), ptr(label_id), ptr(has_track), ptr(point_scan),
        ptr(scan_params), ptr(wp), ptr(flip), nw, n, 1 if select else 0, float(min_x), float(max_x), float(min_y),
        float(max_y), *(ptr(out[k]) for k in _OUT_KEYS), ptr(ts_out), ptr(labels), ptr(src), ptr(stat), ptr(out_ptr),
        ptr(ws), nbytes, stream_ptr()), 'rgnn_accumulate_windows')
    fp = [int(v) for v in out_ptr.cpu().numpy()]          # the one read-back: frame_ptr sizes the graph construction
    m = fp[-1]
    res = AccumulatedWindows()
    res.points = {k: v[:m] for k, v in out.items()}
    res.points['meas_timestamp'] = ts_out[:m]
    res.frame_ptr, res.class_labels, res.src_index = fp, labels[:m], src[:m]
    res.stationary_flag, res.raw_ptr = stat[:n], [int(v) for v in raw_ptr]
    return res


def get_data_for_datagen(radar_mount_data: Dict, radar_data_all_scenes: np.ndarray, odometry_data_all_scenes: np.ndarray,
                         windowed_data: Dict, reject_outlier: bool = False, flip_along_x: bool = False) -> Dict[str, np.ndarray]:
    """The reference's data_dict of one window (read_data.py:489-537), computed on the device: every detection of the
    window, unselected, plus 'class_labels' (generate_gt_labels, compute_node_labels.py:71-86)."""
    if reject_outlier:
        raise NotImplementedError('reject_outlier_by_ransac: the reference shuffles with the global NumPy RNG '
                                  '(meas_selection.py:124-131) and its configuration disables it (constants.py:7)')
    res = accumulate_windows([(radar_mount_data, radar_data_all_scenes, odometry_data_all_scenes, windowed_data)],
                             flip_along_x=[flip_along_x], select=False)
    raw = np.concatenate([radar_data_all_scenes[a:b] for a, b in windowed_data['radar_data_indices']])
    d = {k: v.cpu().numpy() for k, v in res.points.items()}
    d.update(meas_trackid=raw['track_id'], meas_sensorid=raw['sensor_id'],
             stationary_meas_flag=res.stationary_flag.cpu().numpy().astype(np.bool_), meas_label_id=raw['label_id'],
             class_labels=res.class_labels.cpu().numpy())
    return d


def select_meas_within_the_grid(meas_dict: Dict[str, np.ndarray], label_dict: Dict[str, np.ndarray],
                                min_x=0, max_x=100, min_y=-50, max_y=50):
    """reference grid_features.py:162-173 (host boolean mask on dicts the caller already holds on the host; the fused
    device path is accumulate_windows(select=True))."""
    px, py = meas_dict['meas_px'], meas_dict['meas_py']
    flag = (px >= min_x) & (px < max_x) & (py >= min_y) & (py < max_y)
    return {k: v[flag] for k, v in meas_dict.items()}, {k: v[flag] for k, v in label_dict.items()}


def select_moving_data(data_dict: Dict[str, np.ndarray], gt_dict: Dict[str, np.ndarray], new_labels_to_id_dict=LABELS_TO_ID):
    """reference graph_features.py:167-182"""
    flag = gt_dict['class_labels'] != new_labels_to_id_dict['STATIC']
    return {k: v[flag] for k, v in data_dict.items()}, {k: v[flag] for k, v in gt_dict.items()}
