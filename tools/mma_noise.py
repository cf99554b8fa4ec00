"""Developer aid (GPU): rounding noise of one Linear on the tensor-core path against float64."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graph_neural_network_for_radar_perception_b200 import _cabi
from graph_neural_network_for_radar_perception_b200._engine import apply_stack
lib = _cabi.lib()
torch.manual_seed(0)
for K, N in ((64, 64), (128, 64), (128, 128), (64, 128)):
    lin = torch.nn.Linear(K, N).cuda()
    x = torch.randn(20000, K, device='cuda')
    x = torch.where(x > 0, x, 0.01 * x) * 1.3
    ref = (x.double() @ lin.weight.double().T + lin.bias.double())
    out = {}
    for name, opts in (('tc 3xtf32', dict(tensor_cores=1, tf32_passes=3)), ('tc 1xtf32', dict(tensor_cores=1, tf32_passes=1)), ('ffma', dict(tensor_cores=0, tf32_passes=3))):
        for k, v in opts.items():
            lib.rgnn_set_option(k.encode(), v)
        with torch.no_grad():
            y = apply_stack(x, [lin])
        err = (y.double() - ref)
        out[name] = (float(err.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()), float(err.abs().max() / ref.abs().max()),
                     float((err * ref.sign()).mean() / ref.abs().mean()))
    y32 = x @ lin.weight.T + lin.bias
    err = y32.double() - ref
    out['torch fp32'] = (float(err.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()), float(err.abs().max() / ref.abs().max()), float((err * ref.sign()).mean() / ref.abs().mean()))
    print(f'K={K} N={N}: ' + '; '.join(f'{k}: rms {v[0]:.2e} max {v[1]:.2e} signed-bias {v[2]:+.2e}' for k, v in out.items()))
