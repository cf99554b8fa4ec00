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


def test_tensor_core_backward_and_ffma_backward_both_meet_the_yardstick(ckpt_state_dict):
    """Same forward, same losses; only the backward of the message function differs (tcgen05 + scratch + wgrad GEMM vs
    recompute tile programs on the CUDA cores).  The two recompute the edge activations with different (both
    fp32-equivalent) arithmetic, so they may land on different sides of an undecidable LeakyReLU kink: they are not
    compared with each other through a blanket allowance but each is held to the reference's float64 gradient with the
    yardstick of tests/test_train_gpu.py (float32 noise of the reference + the exact jumps of the listed kinks)."""
    from gpu_util import GradientYardstick, batch_labels, load_model, synth_batch
    from test_train_gpu import FLIP_WINDOW
    frames = synth_batch((260, 97), seed0=70)
    ys = GradientYardstick(ckpt_state_dict, frames, members=4, flip_window=FLIP_WINDOW)
    m = load_model(ckpt_state_dict).train()

    def grads():
        m.zero_grad(set_to_none=True)
        loss, _ = m([f['nf'].cuda() for f in frames], [f['ef'].cuda() for f in frames], [f['ei'].cuda() for f in frames],
                    [None] * len(frames), batch_labels(frames, 'cuda'))
        sum(loss.values()).backward()
        return {n: p.grad.detach().double().cpu().numpy() for n, p in m.named_parameters()}
    assert lib().rgnn_get_option(b'tensor_cores_bwd') == 1
    ys.check(grads(), what='tensor-core backward')
    try:
        check(lib().rgnn_set_option(b'tensor_cores_bwd', 0), 'set_option')
        ys.check(grads(), what='FFMA backward')
    finally:
        check(lib().rgnn_set_option(b'tensor_cores_bwd', 1), 'set_option')
