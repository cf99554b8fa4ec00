"""GPU: the tensor-core weight-gradient GEMM (rgnn_wgrad, csrc/rgnn_wgrad_tc.cu) against a float64 matmul, and the
tensor-core backward of the message function against the FFMA backward tile programs on the same inputs.

Tolerance: 3xTF32 products with fp32 accumulation behave like an fp32 GEMM: |got - want| <= 1e-5 * sum_r |A||B| (the
bound of a length-`rows` fp32 dot product is far looser); written per element below."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from graph_neural_network_for_radar_perception_b200._cabi import check, lib, ptr, stream_ptr


@pytest.mark.parametrize('rows,wa,wb,lda,ldb', [
    (4096, 128, 64, 128, 64),        # dW of msg.1 / the edge part of msg.0
    (1000, 128, 256, 128, 256),      # widest operand pair, rows not a multiple of the chunk
    (333, 64, 64, 64, 64),           # A narrower than the UMMA M (zero padded)
    (5000, 128, 7, 128, 7),          # raw feature operand: unaligned rows, width padded to 16
    (17, 128, 64, 128, 64),          # fewer rows than one chunk
    (70000, 128, 128, 160, 128),     # several CTAs (split-K), leading dimension > width
])
def test_wgrad_matches_float64(rows, wa, wb, lda, ldb):
    g = torch.Generator(device='cuda').manual_seed(rows + wa)
    A = torch.randn(rows, lda, device='cuda', generator=g)
    B = torch.randn(rows, ldb, device='cuda', generator=g) * 3.0
    dst0 = torch.randn(wa, wb, device='cuda', generator=g)
    dst = dst0.clone()
    ca = torch.zeros(wa, device='cuda')
    cb = torch.zeros(wb, device='cuda')
    check(lib().rgnn_wgrad(ptr(A), lda, wa, ptr(B), ldb, wb, rows, ptr(dst), ptr(ca), ptr(cb), stream_ptr()), 'wgrad')
    torch.cuda.synchronize()
    Ad, Bd = A[:, :wa].double(), B[:, :wb].double()
    want = dst0.double() + Ad.t() @ Bd
    bound = 1e-5 * (Ad.abs().t() @ Bd.abs()) + 1e-6
    err = (dst.double() - want).abs()
    assert bool((err <= bound).all()), float((err / bound).max())
    assert torch.allclose(ca.double(), Ad.sum(0), rtol=1e-5, atol=1e-4 * float(Ad.abs().sum(0).max()))
    assert torch.allclose(cb.double(), Bd.sum(0), rtol=1e-5, atol=1e-4 * float(Bd.abs().sum(0).max()))


def _grads(model, batch):
    nf, ef, ei, labels = batch
    model.zero_grad(set_to_none=True)
    loss, _ = model(nf, ef, ei, [None] * len(nf), labels)
    sum(loss.values()).backward()
    torch.cuda.synchronize()
    return {n: p.grad.detach().double().cpu().numpy() for n, p in model.named_parameters()}


def test_tensor_core_backward_matches_ffma_backward(ckpt_state_dict):
    """Same forward, same losses; only the backward of the message function differs (tcgen05 + scratch + wgrad GEMM
    vs recompute tile programs on the CUDA cores).  Both are fp32-equivalent, so they agree far inside the tolerance
    either has against the reference (test_train_gpu.py)."""
    from gpu_util import load_model
    from graph_neural_network_for_radar_perception_b200 import synth
    from oracle import graph_np
    m = load_model(ckpt_state_dict).train()
    R = np.float64(np.sqrt(100.0 ** 2 + 50.0 ** 2))
    nf, ef, ei = [], [], []
    labels = {k: [] for k in ('cluster_node_idx', 'cluster_labels', 'edge_class', 'node_class', 'node_offsets')}
    for i, n in enumerate((700, 431, 97)):
        d, src = synth.make_frame(70 + i, n)
        adj = graph_np.adjacency_information(d, 25, 10)
        lab = synth.make_labels(d, src, adj['adj_list'])
        nf.append(torch.from_numpy(graph_np.node_features(d, adj['degree'], True, 0, R, 0, np.pi * 0.5).astype(np.float32)).cuda())
        ef.append(torch.from_numpy(graph_np.edge_features(d, adj['adj_list']).astype(np.float32)).cuda())
        ei.append(torch.from_numpy(adj['adj_list']).cuda())
        labels['cluster_node_idx'].append([torch.from_numpy(c).cuda() for c in lab['cluster_node_idx']])
        for k in ('cluster_labels', 'edge_class', 'node_class', 'node_offsets'):
            labels[k].append(torch.from_numpy(lab[k]).cuda())
    off0 = [t.clone() for t in labels['node_offsets']]

    def batch():
        lb = dict(labels)
        lb['node_offsets'] = [t.clone() for t in off0]       # normalised in place by the training forward
        return nf, ef, ei, lb
    assert lib().rgnn_get_option(b'tensor_cores_bwd') == 1
    g_tc = _grads(m, batch())
    try:
        check(lib().rgnn_set_option(b'tensor_cores_bwd', 0), 'set_option')
        g_ff = _grads(m, batch())
    finally:
        check(lib().rgnn_set_option(b'tensor_cores_bwd', 1), 'set_option')
    rel = []
    for n in g_tc:
        scale = max(np.abs(g_ff[n]).max(), 1e-9)
        rel.append(np.abs(g_tc[n] - g_ff[n]).max() / scale)
        if g_ff[n].size > 1:
            assert rel[-1] < 5e-3, (n, rel[-1])      # kink allowance, see test_train_gpu.py
    assert np.median(rel) < 2e-5, (np.median(rel), max(rel))
