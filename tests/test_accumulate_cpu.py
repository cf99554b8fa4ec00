"""CPU: host logic of the window accumulation pre-pass.  The per-scan parameter table the
host mirror hands to rgnn_accumulate_windows must equal the oracle's (which is pinned against the reference's own outputs
in tests/test_oracle_golden.py), window packing must keep scan order and offsets, and the host-side selection helpers must
reproduce the reference's fixtures.  No compute calls (no GPU here)."""
import os

import numpy as np
import pytest

from graph_neural_network_for_radar_perception_b200 import accumulate as acc, synth
from oracle import accumulate_np as ref


@pytest.mark.parametrize('w,ns,pps', [(0, 10, 300), (5, 1, 7), (9, 4, 60)])
def test_scan_parameters_equal_the_oracle_table(w, ns, pps):
    mounts, rad, odo, win = synth.make_raw_window(w, ns, pps)
    p = acc.scan_parameters(mounts, odo, win)
    t = ref.scan_table(mounts, odo, win)
    assert p.shape == (ns, 9) and p.dtype == np.float64
    assert np.array_equal(p[:, 0:4].reshape(ns, 2, 2), t['R'])
    assert np.array_equal(p[:, 4:6], t['t'])
    assert np.array_equal(p[:, 6], t['mount_yaw'])
    assert np.array_equal(p[:, 7], t['vxs']) and np.array_equal(p[:, 8], t['vys'])
    # the last scan of a window is its own reference frame: identity pose up to rounding of inv(T) @ T
    assert np.allclose(p[-1, 0:4], [1, 0, 0, 1], atol=1e-12) and np.allclose(p[-1, 4:6], 0, atol=1e-9)


def test_pack_windows_keeps_scan_order_and_offsets():
    windows = [synth.make_raw_window(20 + i, ns, pps) for i, (ns, pps) in enumerate([(10, 50), (1, 3), (6, 200)])]
    raw, scan_of, params, raw_ptr = acc._pack_windows(windows)
    assert raw_ptr[0] == 0 and raw_ptr[-1] == raw.shape[0] == scan_of.shape[0]
    assert params.shape == (17, 9)
    assert np.all(np.diff(scan_of) >= 0) and scan_of[0] == 0 and scan_of[-1] == 16
    s0 = 0
    for i, (_, rad, _, win) in enumerate(windows):
        a, b = raw_ptr[i], raw_ptr[i + 1]
        want = np.concatenate([rad[x:y] for x, y in win['radar_data_indices']])
        assert np.array_equal(raw[a:b], want)
        assert scan_of[a] == s0 and scan_of[b - 1] == s0 + len(win['radar_id']) - 1
        s0 += len(win['radar_id'])


def test_host_selection_helpers_match_reference_fixtures(golden_dir):
    g = np.load(os.path.join(golden_dir, 'accumulate.npz'))
    for c in range(len([k for k in g.files if k.endswith('_args')])):
        p = f'c{c}_'
        d = {k: g[p + 'all_' + k] for k in ('meas_px', 'meas_py', 'meas_vx', 'meas_vy', 'meas_vr', 'meas_rcs', 'meas_timestamp')}
        gt = {'class_labels': g[p + 'all_class_labels']}
        d, gt = acc.select_meas_within_the_grid(d, gt)
        d, gt = acc.select_moving_data(d, gt)
        for k in d:
            assert np.array_equal(d[k], g[p + 'dyn_' + k]), (c, k)
        assert np.array_equal(gt['class_labels'], g[p + 'dyn_class_labels'])


def test_no_cuda_device_raises_instead_of_falling_back():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA device present')
    from graph_neural_network_for_radar_perception_b200._cabi import RgnnError
    with pytest.raises(RgnnError):
        acc.accumulate_windows([synth.make_raw_window(0, 2, 10)])


def test_window_shards_of_two_ranks_concatenate_to_the_single_process_pack():
    """Inference shards windows across ranks with no collective: the packs of the per-rank shards
    (training.shard_range) are exactly the slices of the single-process pack, so per-rank results concatenate to the batch."""
    from graph_neural_network_for_radar_perception_b200.training import shard_range
    windows = [synth.make_raw_window(40 + i, 3 + i % 4, 30 + 7 * i) for i in range(7)]
    raw, scan_of, params, raw_ptr = acc._pack_windows(windows)
    got_raw, got_params, n0 = [], [], 0
    for rank in range(2):
        rg = shard_range(len(windows), rank, 2)
        a, b = rg.start, rg.stop
        r, s, p, rp = acc._pack_windows(windows[a:b])
        assert rp == [v - raw_ptr[a] for v in raw_ptr[a:b + 1]]
        got_raw.append(r)
        got_params.append(p)
        n0 += b - a
    assert n0 == len(windows)
    assert np.array_equal(np.concatenate(got_raw), raw)
    assert np.array_equal(np.concatenate(got_params), params)
