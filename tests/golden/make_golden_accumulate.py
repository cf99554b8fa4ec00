"""Golden fixtures for the sliding-window accumulation pre-pass: runs the UNMODIFIED
reference functions from /root/reference in the build container on synthetic RadarScenes-shaped windows
(synth.make_raw_window) and stores their outputs.

    python tests/golden/make_golden_accumulate.py      ->  tests/golden/accumulate.npz

Reference entry points exercised: modules/data_utils/read_data.py::extract_and_sync_radar_data,
convert_list_ndarry_to_ndarray (+ the float32 casts / flip of get_data_for_datagen :524-537),
modules/compute_groundtruth/compute_node_labels.py::compute_ground_truth,
modules/compute_features/grid_features.py::grid_properties.select_meas_within_the_grid,
modules/compute_features/graph_features.py::select_moving_data.  read_data.py imports h5py (absent here, only used
by the file readers), which is stubbed by an empty module for the import.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, REPO)
sys.path.insert(0, REF)
sys.modules.setdefault('h5py', types.ModuleType('h5py'))

CASES = [  # (window_idx, n_scans, points_per_scan, flip_along_x)
    (0, 10, 300, False), (1, 10, 120, True), (2, 3, 40, False), (3, 1, 25, False), (4, 10, 900, False)]


def reference_window(rd, cnl, gf, grid, labels, mounts, rad, odo, win, flip):
    lists = rd.extract_and_sync_radar_data(mounts, rad, odo, win, False)
    (px, py, vx, vy, vr, rcs, ts, trk, sid, stat, lab) = rd.convert_list_ndarry_to_ndarray(*lists)
    if flip:
        py, vy = -py, -vy
    d = {'meas_px': px.astype(np.float32), 'meas_py': py.astype(np.float32), 'meas_vx': vx.astype(np.float32),
         'meas_vy': vy.astype(np.float32), 'meas_vr': vr.astype(np.float32), 'meas_rcs': rcs.astype(np.float32),
         'meas_timestamp': ts, 'meas_trackid': trk, 'meas_sensorid': sid, 'stationary_meas_flag': stat,
         'meas_label_id': lab}
    full = {k: v.copy() for k, v in d.items()}
    gt = cnl.compute_ground_truth(d, labels.compute_new_labels_to_id_dict(), labels.compute_old_to_new_label_id_map())
    full_labels = gt['class_labels'].copy()
    d, gt = grid.select_meas_within_the_grid(d, gt)
    d_dyn, gt_dyn = gf.select_moving_data(d, gt, labels.compute_new_labels_to_id_dict())
    return full, full_labels, px, py, d_dyn, gt_dyn


def main():
    from modules.data_utils import read_data as rd, labels
    from modules.compute_groundtruth import compute_node_labels as cnl
    from modules.compute_features import graph_features as gf
    from modules.compute_features.grid_features import grid_properties
    from graph_neural_network_for_radar_perception_b200 import synth
    grid = grid_properties(0, 100, -50, 50, 0.1, 1.0, 0.1, 1.0, 1.0, 1.0)     # ROI of configuration_radarscenes_gnn.yml:33-38
    out = {}
    for c, (w, ns, pps, flip) in enumerate(CASES):
        mounts, rad, odo, win = synth.make_raw_window(w, ns, pps)
        full, full_labels, px64, py64, d_dyn, gt_dyn = reference_window(rd, cnl, gf, grid, labels, mounts, rad, odo, win, flip)
        p = f'c{c}_'
        out[p + 'args'] = np.array([w, ns, pps, int(flip)], dtype=np.int64)
        out[p + 'px64'], out[p + 'py64'] = px64, py64
        for k in ('meas_px', 'meas_py', 'meas_vx', 'meas_vy', 'meas_vr', 'meas_rcs', 'meas_timestamp', 'stationary_meas_flag'):
            out[p + 'all_' + k] = full[k]
        out[p + 'all_class_labels'] = full_labels
        for k in ('meas_px', 'meas_py', 'meas_vx', 'meas_vy', 'meas_vr', 'meas_rcs', 'meas_timestamp', 'meas_sensorid',
                  'meas_label_id', 'stationary_meas_flag'):
            out[p + 'dyn_' + k] = d_dyn[k]
        out[p + 'dyn_meas_trackid'] = d_dyn['meas_trackid']
        out[p + 'dyn_class_labels'] = gt_dyn['class_labels']
        out[p + 'dyn_offsetx'], out[p + 'dyn_offsety'] = gt_dyn['offsetx'], gt_dyn['offsety']
        print(f'case {c}: window {w}, {ns} scans, {full["meas_px"].shape[0]} detections -> {d_dyn["meas_px"].shape[0]} '
              f'dynamic in the ROI; stationary {full["stationary_meas_flag"].mean():.2f}')
    np.savez_compressed(os.path.join(HERE, 'accumulate.npz'), **out)


if __name__ == '__main__':
    main()
