"""ORACLE (test infrastructure, never imported by the product package): CPU restatement of the reference's
sliding-window accumulation pre-pass, the step that produces the `meas_*` arrays
the graph construction reads:

    extract_and_sync_radar_data      reference modules/data_utils/read_data.py:227-303
      identify_stationary_measurements / gate_stationary_meas_flag     modules/data_utils/meas_selection.py:24-70,169-200
      vr_cartesian_vf, construct_SE2_group_element, ego_compensate_*   modules/data_utils/meas_sync.py:15-102
    convert_list_ndarry_to_ndarray + float32 casts (+ flip)            read_data.py:306-327, 489-537
    generate_gt_labels                                                 modules/compute_groundtruth/compute_node_labels.py:71-86
    grid_properties.select_meas_within_the_grid                        modules/compute_features/grid_features.py:162-173
    select_moving_data                                                 modules/compute_features/graph_features.py:167-182

It is written per window over concatenated arrays (a per-scan parameter table + one vectorised pass over the
points) instead of the reference's per-scan Python lists; the arithmetic of every element is the reference's:

* positions: `R @ [px; py] + t` in float64 with `T = inv(T_curr) @ T_prev` (NumPy `linalg.inv` + matmul on 3x3
  float64), cast to float32 afterwards.  The 2x2 @ 2xN product goes through OpenBLAS dgemm, whose micro-kernel
  evaluates `fma(R01, py, R00 * px)` (probed in this container, tests/golden/make_golden_accumulate.py); the
  restatement reproduces that with an exact two-product sum, so float64 results are bit-identical here.
* velocities: float32 throughout (`vr * cos(azimuth + mount_yaw)`; the Python-float mount yaw is a weak scalar under
  NEP 50, so the sum and the cos/sin stay float32); ego compensation leaves them unrotated (meas_sync.py:67-68).
* stationary gate: per-scan sensor velocity in float64 from the odometry scalars, `cos/sin` of the float32 azimuth
  promoted to float64 by the float64 scalar factors, `|vr_pred - vr| <= 1.5`.

Pinned: tests/golden/accumulate.npz holds the outputs of the reference's own functions on synthetic windows
(tests/golden/make_golden_accumulate.py imports the unmodified files); tests/test_oracle_golden.py holds this
restatement to them bit-exactly.
"""
from __future__ import annotations

from fractions import Fraction
from typing import Dict, Tuple

import numpy as np

GAMMA_STATIONARY = 1.5                    # reference modules/data_utils/constants.py:15
LABEL_FALSE, LABEL_STATIC = 6, 7          # reference modules/data_utils/labels.py:60-70
# old label id -> new label id (labels.py:18-31, 44-58, 90-100)
OLD_TO_NEW_LABEL = np.array([0, 4, 4, 4, 4, 3, 3, 1, 2, 5, 5, 7], dtype=np.int32)


def se2(px, py, theta) -> np.ndarray:
    """meas_sync.py:24-33."""
    T = np.eye(3)
    T[:2, :2] = np.array([[np.cos(theta), -np.sin(theta)], [np.sin(theta), np.cos(theta)]])
    T[:2, 2:] = np.array([[px], [py]])
    return T


def scan_table(radar_mount_data: Dict, odometry: np.ndarray, windowed_data: Dict) -> Dict[str, np.ndarray]:
    """Per-scan parameters of one window, all float64 unless noted:
    R (S,2,2), t (S,2): T = inv(T_curr) @ T_prev (meas_sync.py:55, 88-96; T_curr = the LAST scan of the window);
    mount_yaw (S,) float64 Python value; vxs, vys (S,): ego velocity at the sensor in the sensor frame
    (meas_selection.py:24-37 with vy_ego = 0, :184)."""
    ids = windowed_data['radar_id']
    odo = [odometry[i] for i in windowed_data['odometry_index']]
    Ts = [se2(o['x_seq'], o['y_seq'], o['yaw_seq']) for o in odo]
    Tc_inv = np.linalg.inv(Ts[-1])
    S = len(ids)
    out = dict(R=np.zeros((S, 2, 2)), t=np.zeros((S, 2)), mount_yaw=np.zeros(S), vxs=np.zeros(S), vys=np.zeros(S))
    for s, (rid, o, T_prev) in enumerate(zip(ids, odo, Ts)):
        T = Tc_inv @ T_prev
        out['R'][s], out['t'][s] = T[:2, :2], T[:2, 2]
        m = radar_mount_data['radar_' + str(rid)]
        tx, ty, theta = m['x'], m['y'], m['yaw']
        vx_ego, vy_ego, w = o['vx'], 0.0, o['yaw_rate']
        vx_s = vx_ego - w * ty
        vy_s = vy_ego + w * tx
        th = -theta
        out['vxs'][s] = vx_s * np.cos(th) - vy_s * np.sin(th)          # meas_selection.py:11-20
        out['vys'][s] = vx_s * np.sin(th) + vy_s * np.cos(th)
        out['mount_yaw'][s] = theta
    return out


def _dgemm_2term(a, x, b, y):
    """fma(b, y, a * x) elementwise in float64: what the OpenBLAS dgemm micro-kernel computes for a K = 2 product."""
    p = a * x                                       # rounded first product
    out = np.empty_like(p)
    fb = Fraction(float(b))
    for i in range(p.shape[0]):
        out[i] = float(fb * Fraction(float(y[i])) + Fraction(float(p[i])))
    return out


def accumulate_window(radar_mount_data: Dict, radar_data: np.ndarray, odometry: np.ndarray, windowed_data: Dict,
                      flip_along_x: bool = False, exact_dgemm: bool = True) -> Dict[str, np.ndarray]:
    """The reference's get_data_for_datagen (read_data.py:489-537) minus file reading: extract_and_sync_radar_data,
    concatenation, float32 casts, optional flip.  `exact_dgemm=False` replaces the exact fma emulation (a Python
    loop) by `a*x + b*y`, which differs by at most 1 ulp in float64 and practically never after the float32 cast;
    only large benchmark samples use it."""
    tab = scan_table(radar_mount_data, odometry, windowed_data)
    parts = [radar_data[a:b] for a, b in windowed_data['radar_data_indices']]
    px, py, vx, vy, stat = [], [], [], [], []
    for s, r in enumerate(parts):
        az, vr_raw = r['azimuth_sc'], r['vr']
        # stationary gate (meas_selection.py:39-70): float32 cos/sin, float64 products
        vr_pred = -(np.float64(tab['vxs'][s]) * np.cos(az) + np.float64(tab['vys'][s]) * np.sin(az))
        stat.append(np.abs(vr_pred - vr_raw) <= GAMMA_STATIONARY)
        # velocity of the compensated range rate in the vehicle frame (meas_sync.py:15-21): float32
        angle = az + float(tab['mount_yaw'][s])
        vx.append(r['vr_compensated'] * np.cos(angle))
        vy.append(r['vr_compensated'] * np.sin(angle))
        # ego compensation of the positions (meas_sync.py:44-70)
        R, t = tab['R'][s], tab['t'][s]
        x64, y64 = r['x_cc'].astype(np.float64), r['y_cc'].astype(np.float64)
        if exact_dgemm:
            px.append(_dgemm_2term(R[0, 0], x64, R[0, 1], y64) + t[0])
            py.append(_dgemm_2term(R[1, 0], x64, R[1, 1], y64) + t[1])
        else:
            px.append(R[0, 0] * x64 + R[0, 1] * y64 + t[0])
            py.append(R[1, 0] * x64 + R[1, 1] * y64 + t[1])
    cat = np.concatenate
    px, py, vx, vy = cat(px), cat(py), cat(vx), cat(vy)
    allr = cat(parts)
    if flip_along_x:
        py, vy = -py, -vy
    return {'meas_px': px.astype(np.float32), 'meas_py': py.astype(np.float32),
            'meas_vx': vx.astype(np.float32), 'meas_vy': vy.astype(np.float32),
            'meas_vr': allr['vr_compensated'].astype(np.float32), 'meas_rcs': allr['rcs'].astype(np.float32),
            'meas_timestamp': allr['timestamp'], 'meas_trackid': allr['track_id'], 'meas_sensorid': allr['sensor_id'],
            'stationary_meas_flag': cat(stat), 'meas_label_id': allr['label_id']}


def class_labels(data_dict: Dict[str, np.ndarray]) -> np.ndarray:
    """generate_gt_labels (compute_node_labels.py:71-86): float32 class id per measurement."""
    trk, stat = data_dict['meas_trackid'], data_dict['stationary_meas_flag']
    lab = np.zeros(trk.shape[0], dtype=np.float32)
    has = trk != b''
    lab[has] = OLD_TO_NEW_LABEL[data_dict['meas_label_id']][has]
    lab[~has & ~stat] = LABEL_FALSE
    lab[~has & stat] = LABEL_STATIC
    return lab


def select(data_dict: Dict[str, np.ndarray], labels: np.ndarray, min_x=0, max_x=100, min_y=-50, max_y=50
           ) -> Tuple[Dict[str, np.ndarray], np.ndarray, np.ndarray]:
    """select_meas_within_the_grid (grid_features.py:162-173) followed by select_moving_data
    (graph_features.py:167-182).  Returns (data_dict_dyn, class_labels_dyn, kept indices into the window)."""
    px, py = data_dict['meas_px'], data_dict['meas_py']
    keep = (px >= min_x) & (px < max_x) & (py >= min_y) & (py < max_y) & (labels != LABEL_STATIC)
    idx = np.nonzero(keep)[0]
    return {k: v[idx] for k, v in data_dict.items()}, labels[idx], idx
