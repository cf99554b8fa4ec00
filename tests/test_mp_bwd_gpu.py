"""GPU: backward of the message function of one residual_graph_conv_block through the C-ABI (rgnn_conv_msg_bwd) against
float64 torch autograd over the reference's formula (gnn_blocks.py:106-113: cat(x_i, x_j, e) -> ffn 192->128 -> ffn 128->64
-> sum at the target), for the fused fp16-split kernel (csrc/rgnn_mp_bwd_f16.cu: dgrad + both weight gradients in one
launch) and for the 3xTF32 kernel + weight-gradient GEMMs it replaces (rgnn_set_option("f16_bwd", 0)).

Tolerance, per tensor: |got - want| <= 1e-4 |want| + 2e-5 max|want|.  A LeakyReLU whose pre-activation float32 cannot
tell from zero takes either branch (tests/test_train_gpu.py), which moves single rows of d(emb) by a finite amount: up to
2e-4 of the ROWS may exceed the bound (the float64 reference itself flips those when its inputs move by one float32 ulp);
the weight gradients, sums over all edges, must meet it everywhere at 5x the bound."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from graph_neural_network_for_radar_perception_b200._cabi import check, lib, ptr, stream_ptr
from graph_neural_network_for_radar_perception_b200._engine import GraphBatch, detector_table, flat_grads


def _leaky(z):
    return torch.where(z > 0, z, 0.01 * z)


def _norm(z, s, m):
    mu = z.mean(dim=1, keepdim=True)
    sd = z.std(dim=1, keepdim=True)
    return s * ((z - mu) / (sd + 1e-5)) + m


def _reference(sd, l, P, e, src, tgt, dagg):
    """float64 autograd: returns d(P), d(e), and the gradients of msg.0 (edge columns, bias, scalars) and msg.1."""
    p = f'pred.pass_messages.conv_blk.{l}.'
    W1 = sd[p + 'msg.0.block.0.weight'].double()
    We = W1[:, 128:].clone().requires_grad_(True)
    s1 = sd[p + 'msg.0.block.1.std'].double().clone().requires_grad_(True)
    m1 = sd[p + 'msg.0.block.1.mu'].double().clone().requires_grad_(True)
    W2 = sd[p + 'msg.1.block.0.weight'].double().clone().requires_grad_(True)
    b2 = sd[p + 'msg.1.block.0.bias'].double().clone().requires_grad_(True)
    s2 = sd[p + 'msg.1.block.1.std'].double().clone().requires_grad_(True)
    m2 = sd[p + 'msg.1.block.1.mu'].double().clone().requires_grad_(True)
    Pd = P.double().clone().requires_grad_(True)
    ed = e.double().clone().requires_grad_(True)
    z1 = ed @ We.t() + Pd[tgt, :128] + Pd[src, 128:]
    z1.retain_grad()
    y1 = _leaky(_norm(z1, s1, m1))
    msg = _leaky(_norm(y1 @ W2.t() + b2, s2, m2))
    (msg * dagg.double()[tgt]).sum().backward()
    return {'dP': Pd.grad, 'de': ed.grad, 'We': We.grad, 'b1': z1.grad.sum(0), 's1': s1.grad, 'm1': m1.grad, 'W2': W2.grad,
            'b2': b2.grad, 's2': s2.grad, 'm2': m2.grad}


def _run(det, table, views, l, gb, P, e, dagg):
    for v in views:
        if v is not None:
            v.zero_()
    table.refill(views)
    s = stream_ptr()
    table.ensure_packed(s)
    g = gb.c_struct()
    conv = table.det.conv[l]
    L = lib()
    nb = L.rgnn_conv_msg_bwd_workspace_bytes(C.byref(conv), C.byref(g))
    ws = torch.empty(nb, dtype=torch.uint8, device='cuda')
    dP = torch.full((gb.n_nodes, 256), float('nan'), device='cuda')
    de = torch.full((gb.n_edges, 64), float('nan'), device='cuda')
    check(L.rgnn_conv_msg_bwd(C.byref(conv), C.byref(g), ptr(e), ptr(P), ptr(dagg), ptr(dP), ptr(de), ptr(ws), nb, s), 'conv_msg_bwd')
    torch.cuda.synchronize()
    blk = det.pass_messages.conv_blk[l]
    params = table.tab.tensors

    def grad_of(t):
        return views[[i for i, q in enumerate(params) if q is t][0]].detach().double().cpu()
    m0, m1 = blk.msg[0].block, blk.msg[1].block
    out = {'dP': dP.double().cpu(), 'de': de.double().cpu(), 'We': grad_of(m0[0].weight)[:, 128:], 'b1': grad_of(m0[0].bias),
           's1': grad_of(m0[1].std), 'm1': grad_of(m0[1].mu), 'W2': grad_of(m1[0].weight), 'b2': grad_of(m1[0].bias),
           's2': grad_of(m1[1].std), 'm2': grad_of(m1[1].mu)}
    table.refill(None)
    return out


def _graph(sizes, seed, k=10):
    from graph_neural_network_for_radar_perception_b200 import synth
    from oracle import graph_np
    eis = []
    for i, n in enumerate(sizes):
        d, _ = synth.make_frame(seed + i, n, knn=k)
        eis.append(torch.from_numpy(graph_np.adjacency_information(d, 25, k)['adj_list']).cuda())
    return GraphBatch.from_frames(eis, sizes)


@pytest.mark.parametrize('sizes,layer,gscale', [
    ((260, 97, 2), 0, 1e-3),          # ragged tiles, last tile partial
    ((1500,), 3, 3e-7),               # tiny gradients: the per-launch power-of-two scale has to carry them
    ((700, 640), 6, 40.0),            # large gradients
])
def test_message_backward_matches_float64_autograd(ckpt_state_dict, sizes, layer, gscale):
    from gpu_util import load_model
    model = load_model(ckpt_state_dict).train()
    det = model.pred
    table = detector_table(det)
    params = table.tab.tensors
    flat, views = flat_grads(params, [True] * len(params))
    gb = _graph(sizes, 4100 + layer)
    N, E = gb.n_nodes, gb.n_edges
    gen = torch.Generator(device='cuda').manual_seed(17 + layer)
    x = torch.randn(N, 64, device='cuda', generator=gen)
    e = torch.randn(E, 64, device='cuda', generator=gen)
    sd = {k: v.cuda() for k, v in ckpt_state_dict.items()}
    W1 = sd[f'pred.pass_messages.conv_blk.{layer}.msg.0.block.0.weight']
    b1 = sd[f'pred.pass_messages.conv_blk.{layer}.msg.0.block.0.bias']
    P = torch.cat((x @ W1[:, :64].t() + b1, x @ W1[:, 64:128].t()), dim=1).contiguous()
    # gradients with a wide dynamic range over the nodes, as in a real step (tools/grad_sensitivity.py: 4 decades)
    dagg = torch.randn(N, 64, device='cuda', generator=gen) * gscale * torch.pow(10.0, -3.0 * torch.rand(N, 1, device='cuda', generator=gen))
    # the kernels run on the TARGET-major order of the graph: hand the reference the same edge rows
    src, tgt = gb.src.long(), gb.tgt.long()
    want = _reference({k: v.cpu() for k, v in sd.items()}, layer, P.cpu(), e.cpu(), src.cpu(), tgt.cpu(), dagg.cpu())
    res = {}
    for mode in (1, 0):
        check(lib().rgnn_set_option(b'f16_bwd', mode), 'opt')
        try:
            res[mode] = _run(det, table, views, layer, gb, P, e, dagg)
        finally:
            check(lib().rgnn_set_option(b'f16_bwd', 1), 'opt')
    for mode, name in ((1, 'fused fp16-split kernel'), (0, '3xTF32 kernel + wgrad GEMMs')):
        got = res[mode]
        for k, w in want.items():
            w = w.detach()
            g_ = got[k].reshape(w.shape)
            assert bool(torch.isfinite(g_).all()), (name, k)
            bound = 1e-4 * w.abs() + 2e-5 * w.abs().max()
            ratio = ((g_ - w).abs() / bound)
            if k in ('dP', 'de'):
                bad_rows = (ratio.max(dim=1)[0] > 1.0).double().mean().item()
                assert bad_rows <= 2e-4, f'{name}: {k}: {bad_rows:.2e} of the rows outside the bound (worst ratio {ratio.max().item():.1f})'
            else:
                assert ratio.max().item() <= 5.0, f'{name}: {k}: worst ratio {ratio.max().item():.2f}'
