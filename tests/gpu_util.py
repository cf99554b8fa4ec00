import numpy as np
import torch

from graph_neural_network_for_radar_perception_b200 import config, Model_Training


def clusters_from(ptr, members, device=None):
    out = [torch.from_numpy(members[ptr[i]:ptr[i + 1]]) for i in range(len(ptr) - 1)]
    return [c.to(device) for c in out] if device is not None else out


def load_model(sd, device='cuda'):
    m = Model_Training(config(), device)
    m.load_state_dict(sd, strict=True)
    return m.to(device)


def assert_close(a, b, rtol, atol, name=''):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    if not np.all(err <= tol):
        i = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError(f'{name}: max violation at {i}: got {a[i]} want {b[i]} (err {err[i]:.3e}, tol {tol[i]:.3e}); '
                             f'max abs err {err.max():.3e}')


def scaled_atol(ref, frac=1e-5, floor=2e-6):
    """Absolute tolerance as a fraction of the reference tensor's largest magnitude (see tests/test_model_gpu.py)."""
    return max(float(np.abs(np.asarray(ref)).max()) * frac, floor)
