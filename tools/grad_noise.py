"""Developer aid (GPU): where does the CUDA gradient stand against the reference's own float32 noise?
    python tools/grad_noise.py [--sizes 90,7,161] [--seed0 700] [--rand]
Prints, for several library option settings, the yardstick ratio of tests/gpu_util.GradientYardstick and the ten worst tensors."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from gpu_util import GradientYardstick, batch_labels, load_model, synth_batch  # noqa: E402
from graph_neural_network_for_radar_perception_b200 import _cabi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sizes', default='90,7,161')
    ap.add_argument('--seed0', type=int, default=700)
    ap.add_argument('--rand', action='store_true')
    ap.add_argument('--members', type=int, default=5)
    ap.add_argument('--flip-window', type=float, default=0.0)
    args = ap.parse_args()
    sizes = tuple(int(v) for v in args.sizes.split(','))
    if args.rand:
        from graph_neural_network_for_radar_perception_b200 import config, Model_Training
        torch.manual_seed(1234)
        sd0 = {k: v.detach().clone() for k, v in Model_Training(config(), 'cpu').state_dict().items()}
    else:
        sd0 = torch.load(os.path.join(ROOT, 'tests', 'golden', 'graph_based_detector.pt'), map_location='cpu', weights_only=True)
    frames = synth_batch(sizes, seed0=args.seed0)
    ys = GradientYardstick(sd0, frames, members=args.members, flip_window=args.flip_window or None)
    print('activations inside the kink window:', ys.n_flips)
    loo = [ys.ratio(ys.runs[j], ys.runs[:j] + ys.runs[j + 1:])[0] for j in range(len(ys.runs))]
    print('held-out float32 oracle runs:', [round(x, 2) for x in loo])
    lib = _cabi.lib()
    settings = [dict(), dict(tensor_cores_bwd=0), dict(tensor_cores=0, tensor_cores_bwd=0), dict(wgrad_tma=0)]
    for st in settings:
        for k in ('tensor_cores', 'tensor_cores_bwd', 'wgrad_tma'):
            lib.rgnn_set_option(k.encode(), int(st.get(k, 1)))
        m = load_model(sd0).train()
        loss, _ = m([f['nf'].cuda() for f in frames], [f['ef'].cuda() for f in frames], [f['ei'].cuda() for f in frames],
                    [None] * len(frames), batch_labels(frames, 'cuda'))
        sum(loss.values()).backward()
        got = {n: p.grad.detach().cpu().numpy() for n, p in m.named_parameters()}
        rows = []
        for n in ys.names:
            ex = ys.exact[n]
            err = np.abs(got[n].astype(np.float64) - ex)
            tol = np.maximum(1e-4 * np.abs(ex), max(ys._noise(ys.runs, n), 1e-30))
            if ys.flip_tol is not None:
                tol = tol + ys.flip_tol[n]
            rows.append(((err / tol).max(), n, err.max() / max(np.abs(ex).max(), 1e-30), ys._noise(ys.runs, n) / max(np.abs(ex).max(), 1e-30)))
        rows.sort(reverse=True)
        print(f'--- options {st or "default"}: worst ratio {rows[0][0]:.1f}, median ratio {np.median([r[0] for r in rows]):.2f}, '
              f'tensors over 2: {sum(r[0] > 2 for r in rows)}')
        for r in rows[:10]:
            print(f'   ratio {r[0]:9.1f}  err/max {r[2]:.2e}  noise/max {r[3]:.2e}  {r[1]}')
        loss_err = {k: abs(float(loss[k]) - float(ys.loss32[k])) / abs(float(ys.loss32[k])) for k in loss}
        print('   loss rel err vs float32 oracle:', {k: f'{v:.1e}' for k, v in loss_err.items()})


if __name__ == '__main__':
    main()
