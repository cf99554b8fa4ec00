"""GPU: training step (forward + multi-task loss + backward kernels) through the C-ABI against the reference's
own losses / gradients (tests/golden/train_2frames.npz) and against torch autograd over the oracle.

Losses and accuracies: rtol 1e-4 (+1e-6) against the reference's float32 values.
Gradients: north_star asks rtol 1e-4.  The yardstick for everything rtol 1e-4 cannot decide is the reference's OWN
float32 rounding noise, per tensor, measured in the test (gpu_util.GradientYardstick): the exact gradient is the oracle in
float64; five float32 evaluations of the oracle (one plain, four with inputs and parameters moved by one float32 ulp)
give each tensor's noise floor; an element passes when |got - exact| <= max(1e-4 |exact|, noise(tensor)) times a factor
that is calibrated by holding each float32 run out in turn -- the CUDA gradient may score at most 2x the worst held-out
float32 run.  LeakyReLU kinks are handled explicitly instead of by a blanket floor (FLIP_WINDOW): the oracle lists the
activations whose pre-activation is closer to zero than rounding can resolve and measures, in float64, the exact jump of
every gradient tensor when one of them takes the other branch; only those jumps widen the tolerance.  There is no
constant floor: where the reference is quiet (random-init weights) the bound is rtol 1e-4 plus a few 1e-6 of the tensor
maximum, and a 0.3 % error in one tensor fails (tests/test_yardstick_cpu.py); where the reference's own gradient moves by
per cent under a one-ulp perturbation (trained checkpoint, 2 x 3000 points: every gradient is a heavily cancelling sum)
the bound follows it."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from gpu_util import GradientYardstick, assert_close, batch_labels, clusters_from, load_model, synth_batch

# |pre-activation| below which the branch of a LeakyReLU is treated as undecidable.  Activations are O(1) after
# channel_normalization; the tensor cores' round-toward-zero accumulation shifts a GEMM output by up to -5.4e-7 relative
# (tools/mma_noise.py) and a few GEMMs stack up between two normalisations.
FLIP_WINDOW = 2e-6


def model_grads(m):
    return {n: p.grad.detach().cpu().numpy() for n, p in m.named_parameters() if p.grad is not None}


def _golden_batch(g, dev):
    nf, ef, ei = [], [], []
    labels = {k: [] for k in ('cluster_node_idx', 'cluster_labels', 'edge_class', 'node_class', 'node_offsets')}
    for i in range(2):
        nf.append(torch.from_numpy(g[f'f{i}_node_features']).to(dev))
        ef.append(torch.from_numpy(g[f'f{i}_edge_features']).to(dev))
        ei.append(torch.from_numpy(g[f'f{i}_edge_index']).to(dev))
        labels['cluster_node_idx'].append(clusters_from(g[f'f{i}_cluster_ptr'], g[f'f{i}_cluster_members'], dev))
        for k in ('cluster_labels', 'edge_class', 'node_class', 'node_offsets'):
            labels[k].append(torch.from_numpy(g[f'f{i}_{k}']).to(dev))
    return nf, ef, ei, labels


def test_training_step_matches_reference_fixture(golden_dir, ckpt_state_dict):
    g = np.load(os.path.join(golden_dir, 'train_2frames.npz'))
    m = load_model(ckpt_state_dict).train()
    nf, ef, ei, labels = _golden_batch(g, 'cuda')
    off_before = [t.clone() for t in labels['node_offsets']]
    loss, acc = m(nf, ef, ei, [None, None], labels)
    for k in ('loss_node_cls', 'loss_node_reg', 'loss_edge_cls', 'loss_obj_cls'):
        assert_close(loss[k].item(), g[k], 1e-4, 1e-6, k)
    for k in ('segment_accuracy', 'edge_accuracy', 'object_accuracy'):
        assert abs(acc[k].item() - float(g[k])) < 1e-6, k
    for a, b in zip(off_before, labels['node_offsets']):
        assert torch.equal(a, b)                       # caller's label tensors are left untouched
    sum(loss.values()).backward()
    names = [str(n) for n in g['grad_names']]
    params = dict(m.named_parameters())
    assert set(names) == set(params)
    # element-wise: against the float64 oracle on the fixture's inputs, with the float32 oracle's noise as the yardstick (the
    # oracle is pinned to this very fixture's reference gradients by tests/test_oracle_golden.py)
    frames = [dict(nf=torch.from_numpy(g[f'f{i}_node_features']), ef=torch.from_numpy(g[f'f{i}_edge_features']),
                   ei=torch.from_numpy(g[f'f{i}_edge_index']),
                   lab={'cluster_node_idx': [c.numpy() for c in clusters_from(g[f'f{i}_cluster_ptr'], g[f'f{i}_cluster_members'])],
                        **{k: g[f'f{i}_{k}'] for k in ('cluster_labels', 'edge_class', 'node_class', 'node_offsets')}})
              for i in range(2)]
    ys = GradientYardstick(ckpt_state_dict, frames, flip_window=FLIP_WINDOW)
    got = model_grads(m)
    assert set(got) == set(names)
    ys.check(got, what='2-frame reference fixture')
    # and the reference's own stored float32 gradients sit inside the same yardstick (fixture <-> oracle <-> CUDA)
    for n in names:
        key = 'grad::' + n
        if key in g.files:
            noise = max(ys._noise(ys.runs, n), 1e-30)
            assert np.all(np.abs(g[key] - ys.exact[n]) <= np.maximum(1e-4 * np.abs(ys.exact[n]), 4 * noise)), n


def _train_case(sd0, sizes, seed0, members=5, knn=10, flip_window=FLIP_WINDOW):
    frames = synth_batch(sizes, seed0=seed0, knn=knn)
    ys = GradientYardstick(sd0, frames, members=members, flip_window=flip_window)
    m = load_model(sd0).train()
    loss, acc = m([f['nf'].cuda() for f in frames], [f['ef'].cuda() for f in frames], [f['ei'].cuda() for f in frames],
                  [None] * len(frames), batch_labels(frames, 'cuda'))
    for k in loss:
        assert_close(loss[k].item(), ys.loss32[k].item(), 1e-4, 1e-6, k)
    for k in acc:
        assert abs(acc[k].item() - ys.acc32[k].item()) < 2.0 / max(sum(sizes), 1) + 1e-6, k    # one arg-max tie at most
    sum(loss.values()).backward()
    return ys, model_grads(m)


def test_training_step_matches_oracle_autograd_all_parameters(ckpt_state_dict):
    """Every one of the 184 parameter tensors, on a 3-frame batch with ragged sizes (one frame smaller than k)."""
    ys, got = _train_case(ckpt_state_dict, (90, 7, 161), 700)
    assert set(got) == set(ys.names)
    ys.check(got, what='ragged 3-frame batch, trained checkpoint')


def test_training_step_with_the_round1_backward_kernels(ckpt_state_dict):
    """The kernels the fused fp16-split backward replaced stay selectable (rgnn_set_option f16_bwd / f16_node_bwd = 0: 3xTF32
    message backward + weight-gradient GEMMs, row-MLP interpreter for the node-sized halves) and are held to the same yardstick."""
    from graph_neural_network_for_radar_perception_b200._cabi import check, lib
    try:
        check(lib().rgnn_set_option(b'f16_bwd', 0), 'opt')
        check(lib().rgnn_set_option(b'f16_node_bwd', 0), 'opt')
        ys, got = _train_case(ckpt_state_dict, (90, 7, 161), 700)
    finally:
        check(lib().rgnn_set_option(b'f16_bwd', 1), 'opt')
        check(lib().rgnn_set_option(b'f16_node_bwd', 1), 'opt')
    ys.check(got, what='ragged 3-frame batch, round-1 backward kernels')


def test_training_step_random_init_is_tight(ckpt_state_dict):
    """Random-init weights (seed 1234): gradients are not cancelling sums, the reference's noise floor is ~1e-6 of each
    tensor's maximum, so this case holds the CUDA backward to rtol 1e-4 almost everywhere and catches a systematic
    error of a few 1e-4 in any single tensor."""
    from graph_neural_network_for_radar_perception_b200 import config, Model_Training
    torch.manual_seed(1234)
    sd0 = {k: v.detach().clone() for k, v in Model_Training(config(), 'cpu').state_dict().items()}
    ys, got = _train_case(sd0, (150, 260), 720)
    r = ys.check(got, what='2 frames, random-init weights')
    worst_noise = max(ys._noise(ys.runs, n) / max(np.abs(ys.exact[n]).max(), 1e-30) for n in ys.names)
    assert worst_noise < 5e-3, worst_noise          # the yardstick really is tight here


@pytest.mark.timeout(900)
def test_training_step_at_baseline_shape_two_frames_of_3000_points(ckpt_state_dict):
    """BASELINE.json configs[2] shape: 3000 points per frame (E ~ 68 k directed edges, 530 edge tiles + the split-K
    weight-gradient chunks + scratch reuse across the 7 layers), trained checkpoint."""
    # ~1.5e8 activations: the kinks are too many to enumerate (and the reference's own one-ulp noise is already 4e-5 of
    # the tensor maximum in the median, 2.6e-2 in the worst tensor, tools/grad_noise.py), so the yardstick is the noise alone
    ys, got = _train_case(ckpt_state_dict, (3000, 3000), 740, members=5, flip_window=None)
    ys.check(got, what='2 frames x 3000 points, trained checkpoint')


def test_ffn_stack_backward_standalone(ckpt_state_dict):
    """graph_feature_encoding used on its own with autograd (block-level API of gnn_blocks.py)."""
    from gpu_util import one_ulp
    from oracle import model_torch as mt
    torch.manual_seed(3)
    m = load_model(ckpt_state_dict).pred
    x = torch.randn(77, 7)
    w = torch.randn(77, 64)
    (m.encode_edge_feat(x.cuda()) * w.cuda()).sum().backward()

    def grad_fn(dtype, seed):
        gen = torch.Generator().manual_seed(seed or 0)
        pert = (lambda t: t) if seed is None else (lambda t: one_ulp(t, gen))
        sd = {k: pert(v.clone()).to(dtype).requires_grad_(True) for k, v in ckpt_state_dict.items() if 'encode_edge_feat' in k}
        (mt.ffn_stack(sd, 'pred.encode_edge_feat.encoder', pert(x).to(dtype)) * w.to(dtype)).sum().backward()
        return {k[len('pred.encode_edge_feat.'):]: v.grad.double().numpy() for k, v in sd.items()}
    ys = GradientYardstick(grad_fn=grad_fn)
    ys.check({n: p.grad.cpu().numpy() for n, p in m.encode_edge_feat.named_parameters()}, what='edge encoder stand-alone')


def test_frozen_layers_get_no_gradient(ckpt_state_dict, golden_dir):
    g = np.load(os.path.join(golden_dir, 'train_2frames.npz'))
    m = load_model(ckpt_state_dict).train()
    m.pred.freeze_layers_except_object_class_predictor()
    nf, ef, ei, labels = _golden_batch(g, 'cuda')
    loss, _ = m(nf, ef, ei, [None, None], labels)
    sum(loss.values()).backward()
    frames = [dict(nf=torch.from_numpy(g[f'f{i}_node_features']), ef=torch.from_numpy(g[f'f{i}_edge_features']),
                   ei=torch.from_numpy(g[f'f{i}_edge_index']),
                   lab={'cluster_node_idx': [c.numpy() for c in clusters_from(g[f'f{i}_cluster_ptr'], g[f'f{i}_cluster_members'])],
                        **{k: g[f'f{i}_{k}'] for k in ('cluster_labels', 'edge_class', 'node_class', 'node_offsets')}})
              for i in range(2)]
    ys = GradientYardstick(ckpt_state_dict, frames, members=4, flip_window=FLIP_WINDOW)
    live = [n for n, p in m.named_parameters() if 'predict_class' in n]
    for n, p in m.named_parameters():
        assert (p.grad is not None) == (n in live), n
    ys.check(model_grads(m), names=live, what='object-class head only')


def test_sgd_step_matches_torch():
    import ctypes as C
    from graph_neural_network_for_radar_perception_b200._cabi import check, lib, ptr, stream_ptr
    torch.manual_seed(0)
    p = torch.randn(10007, device='cuda'); g1 = torch.randn_like(p); g2 = torch.randn_like(p)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.SGD([ref], lr=0.005, momentum=0.9, weight_decay=1e-4)
    buf = torch.zeros_like(p)
    for i, g in enumerate((g1, g2)):
        ref.grad = g.clone(); opt.step()
        check(lib().rgnn_sgd_step(ptr(p), ptr(g), ptr(buf), p.numel(), 0.005, 0.9, 1e-4, 1.0, int(i == 0), stream_ptr()), 'sgd')
    torch.testing.assert_close(p, ref.detach(), rtol=1e-6, atol=1e-7)


def test_trainer_step_equals_reference_sgd_step(golden_dir, ckpt_state_dict):
    """DataParallelTrainer.step (world size 1): two optimisation steps == torch.optim.SGD on the oracle's gradients."""
    from graph_neural_network_for_radar_perception_b200.training import DataParallelTrainer
    from oracle import model_torch as mt
    g = np.load(os.path.join(golden_dir, 'train_2frames.npz'))
    m = load_model(ckpt_state_dict).train()
    trainer = DataParallelTrainer(m, lr=0.005, momentum=0.9, weight_decay=1e-4)
    nf, ef, ei, labels = _golden_batch(g, 'cuda')
    gb, nfp, efp = m.pack_batch(nf, ef, ei, labels['cluster_node_idx'])
    # reference: the oracle + torch.optim.SGD on the CPU
    sd = {k: torch.nn.Parameter(v.clone()) for k, v in ckpt_state_dict.items()}
    opt = torch.optim.SGD(list(sd.values()), lr=0.005, momentum=0.9, weight_decay=1e-4)
    nf_c, ef_c, ei_c, lab_c = _golden_batch(g, 'cpu')
    for step in range(2):
        loss, _ = trainer.step(gb, nfp, efp, labels)
        opt.zero_grad()
        loss_o, _, _ = mt.training_forward(sd, nf_c, ef_c, ei_c, lab_c)
        sum(loss_o.values()).backward()
        opt.step()
        for k in loss:
            assert_close(loss[k].item(), loss_o[k].item(), 2e-4, 1e-6, f'step {step} {k}')
    for n, p in m.named_parameters():
        ref = sd[n].detach().numpy()
        if ref.size > 1:
            # lr * (gradient tolerance, see the module docstring) * (1 + momentum) over two steps
            assert_close(p.detach().cpu().numpy(), ref, 1e-5, 1e-4, n)


def _width_config(hidden):
    """BASELINE.json configs[4] / SURVEY.md 8(d) C5: `hidden` scales graph_convolution_stem_channels, the encoder tails and the
    head stems; msg_mlp_hidden_dim = 2 x hidden (reference set_config_gnn.py:48-72)."""
    from graph_neural_network_for_radar_perception_b200 import config
    cfg = config()
    cfg.node_feat_enc_stem_channels = [256, 128, hidden]
    cfg.edge_feat_enc_stem_channels = [256, 128, 128, hidden]
    cfg.graph_convolution_stem_channels = [hidden] * 3
    cfg.msg_mlp_hidden_dim = 2 * hidden
    cfg.link_pred_stem_channels = [hidden] * 3
    cfg.node_pred_stem_channels = [hidden] * 3
    return cfg


@pytest.mark.parametrize('hidden', [32, 128, 256])
def test_hidden_width_sweep_forward_matches_oracle(hidden):
    """Hidden widths off the reference plan run the FFMA tile programs (csrc/rgnn_chain.cu: 32-row tiles and regions of up to
    512 columns for hidden 256, whose msg.0 is 768 -> 512 with the node half hoisted) and, where a stack fits it, the generic
    tensor-core row-MLP stages.  Random-init weights (seed 1234), one 150-point frame, all four outputs held to the oracle at
    the forward tolerance of tests/test_model_gpu.py."""
    from graph_neural_network_for_radar_perception_b200 import Model_Training, synth
    from oracle import graph_np, model_torch as mt
    torch.manual_seed(1234)
    m = Model_Training(_width_config(hidden), 'cuda').to('cuda')
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    R = np.float64(np.sqrt(100.0 ** 2 + 50.0 ** 2))
    d, src = synth.make_frame(900, 150)
    adj = graph_np.adjacency_information(d, 25, 10)
    lab = synth.make_labels(d, src, adj['adj_list'])
    nf = torch.from_numpy(graph_np.node_features(d, adj['degree'], True, 0, R, 0, np.pi * 0.5).astype(np.float32))
    ef = torch.from_numpy(graph_np.edge_features(d, adj['adj_list']).astype(np.float32))
    ei = torch.from_numpy(adj['adj_list'])
    clusters = [torch.from_numpy(c) for c in lab['cluster_node_idx']]
    want = mt.detector_forward(sd, nf, ef, ei, clusters)
    with torch.no_grad():
        got = m.pred.eval()(nf.cuda(), ef.cuda(), ei.cuda(), None, [c.cuda() for c in clusters])
    for g, w, name in zip(got, want, ('node_cls', 'node_off', 'link_cls', 'obj_cls')):
        w = w.detach().numpy()
        assert_close(g.cpu().numpy(), w, 1e-4, 1e-5 * max(np.abs(w).max(), 1.0), f'hidden {hidden}: {name}')


@pytest.mark.parametrize('hidden', [32, 128])
def test_hidden_width_training_step_matches_oracle_autograd(hidden):
    """Training off the reference plan: hidden width 32 (every gradient through the recompute tile programs on the CUDA cores; the
    natural-layout weight operands of their dgrad steps are 32 / 96 columns wide, zero-filled to the 64-column blocks of the tile GEMM),
    random-init weights, against the same yardstick as the reference plan."""
    from graph_neural_network_for_radar_perception_b200 import Model_Training
    torch.manual_seed(1234)
    cfg = _width_config(hidden)
    sd0 = {k: v.detach().clone() for k, v in Model_Training(cfg, 'cpu').state_dict().items()}
    frames = synth_batch((150, 90), seed0=760)
    ys = GradientYardstick(sd0, frames, members=5, flip_window=FLIP_WINDOW)
    m = Model_Training(cfg, 'cuda')
    m.load_state_dict(sd0, strict=True)
    m = m.to('cuda').train()
    loss, acc = m([f['nf'].cuda() for f in frames], [f['ef'].cuda() for f in frames], [f['ei'].cuda() for f in frames],
                  [None] * len(frames), batch_labels(frames, 'cuda'))
    for k in loss:
        assert_close(loss[k].item(), ys.loss32[k].item(), 1e-4, 1e-6, k)
    sum(loss.values()).backward()
    got = model_grads(m)
    assert set(got) == set(ys.names)
    ys.check(got, what=f'hidden width {hidden}, random-init weights')


@pytest.mark.parametrize('hidden', [256])
def test_hidden_width_beyond_the_training_envelope_fails_loudly(hidden):
    """Outside the training envelope: the backward holds the hoisted projection gradient (2 x msg hidden
    channels) as one row of at most 256 channels; a wider plan must raise, not compute something else (at hidden 128 it used to run
    and return wrong msg.0 gradients, found with the yardstick of the test above)."""
    from graph_neural_network_for_radar_perception_b200 import Model_Training
    from graph_neural_network_for_radar_perception_b200._cabi import RgnnError
    torch.manual_seed(1234)
    m = Model_Training(_width_config(hidden), 'cuda').to('cuda').train()
    frames = synth_batch((60,), seed0=770)
    loss, _ = m([f['nf'].cuda() for f in frames], [f['ef'].cuda() for f in frames], [f['ei'].cuda() for f in frames],
                [None], batch_labels(frames, 'cuda'))
    with pytest.raises(RgnnError):
        sum(loss.values()).backward()


def _tiny_loader(n_batches, seed0):
    """Batches in the reference's collate format (datagen_gnn.py:143-190): dict of per-frame lists + labels dict."""
    from graph_neural_network_for_radar_perception_b200 import synth
    from oracle import graph_np
    R = np.float64(np.sqrt(100.0 ** 2 + 50.0 ** 2))
    out = []
    for b in range(n_batches):
        gfeat = {k: [] for k in ('node_features_dyn', 'edge_features_dyn', 'edge_index_dyn', 'adj_matrix_dyn')}
        labels = {k: [] for k in ('cluster_node_idx', 'cluster_labels', 'edge_class', 'node_class', 'node_offsets')}
        for f in range(2):
            d, src = synth.make_frame(seed0 + 10 * b + f, 80 + 20 * f)
            adj = graph_np.adjacency_information(d, 25, 10)
            lab = synth.make_labels(d, src, adj['adj_list'])
            gfeat['node_features_dyn'].append(torch.from_numpy(graph_np.node_features(d, adj['degree'], True, 0, R, 0, np.pi * 0.5).astype(np.float32)).cuda())
            gfeat['edge_features_dyn'].append(torch.from_numpy(graph_np.edge_features(d, adj['adj_list']).astype(np.float32)).cuda())
            gfeat['edge_index_dyn'].append(torch.from_numpy(adj['adj_list']).cuda())
            gfeat['adj_matrix_dyn'].append(None)
            labels['cluster_node_idx'].append([torch.from_numpy(c).cuda() for c in lab['cluster_node_idx']])
            for k in ('cluster_labels', 'edge_class', 'node_class', 'node_offsets'):
                labels[k].append(torch.from_numpy(lab[k]).cuda())
        out.append((gfeat, labels))
    return out


def test_nan_batch_is_skipped_on_the_device_and_optimizer_state_round_trips(ckpt_state_dict):
    """Row f1: the reference's skip_batch (training.py:40-45) as a device flag of the fused SGD kernel; trainer.state_dict()."""
    from graph_neural_network_for_radar_perception_b200.training import DataParallelTrainer
    m = load_model(ckpt_state_dict).train()
    t = DataParallelTrainer(m)
    (gf0, lab0), (gf1, lab1) = _tiny_loader(2, 300)

    def step(trainer, model, gfeat, labels):
        lb = {k: ([x.clone() for x in v] if k == 'node_offsets' else v) for k, v in labels.items()}
        gb, nf, ef = model.pack_batch(gfeat['node_features_dyn'], gfeat['edge_features_dyn'], gfeat['edge_index_dyn'], lb['cluster_node_idx'])
        return trainer.step(gb, nf, ef, lb)
    step(t, m, gf0, lab0)
    before = t.buffers.flat_param.clone()
    mom_before = t.buffers.momentum.clone()
    bad = dict(lab1)
    bad['node_offsets'] = [x.clone() for x in lab1['node_offsets']]
    bad['node_offsets'][0][3, 0] = float('nan')
    loss, _ = step(t, m, gf1, bad)
    assert bool(torch.isnan(sum(loss.values())))
    assert torch.equal(t.buffers.flat_param, before) and torch.equal(t.buffers.momentum, mom_before)     # nothing was written
    # checkpoint: a fresh trainer restored from (model, optimizer) state continues identically up to the summation order of
    # the split-K weight-gradient reductions (red.global.add across CTAs; the reference's GPU scatter-add is unordered too)
    sd_model = {k: v.clone() for k, v in m.state_dict().items()}
    sd_opt = t.state_dict()
    m2 = load_model(sd_model).train()
    t2 = DataParallelTrainer(m2)
    t2.load_state_dict(sd_opt)
    step(t, m, gf1, lab1)
    step(t2, m2, gf1, lab1)
    assert torch.allclose(t.buffers.flat_param, t2.buffers.flat_param, rtol=1e-5, atol=1e-7)
    assert torch.allclose(t.buffers.momentum, t2.buffers.momentum, rtol=1e-4, atol=1e-7)
    assert not torch.equal(t.buffers.flat_param, before)


def test_train_model_loop_runs_without_per_step_host_reads(ckpt_state_dict):
    from graph_neural_network_for_radar_perception_b200.training import DataParallelTrainer, train_model
    m = load_model(ckpt_state_dict).train()
    t = DataParallelTrainer(m, lr=1e-3)
    batches = _tiny_loader(3, 500)

    def endless():
        while True:
            for b in batches:
                yield (b[0], {k: ([x.clone() for x in v] if k == 'node_offsets' else v) for k, v in b[1].items()})
    val = [(b[0], {k: ([x.clone() for x in v] if k == 'node_offsets' else v) for k, v in b[1].items()}) for b in batches[:1]]
    saved = []
    tracker = train_model(m, t, lr_milestones=[4], dataloader_train=endless(), dataloader_val=val, tb_writer=None, max_iters=6,
                          log_period=100, val_period=5, save_fn=lambda det, trn, it: saved.append(it))
    hist = tracker.loss_history
    assert len(hist) == 6 and all(np.isfinite(hist)) and hist[-1] < hist[0]       # it optimises
    assert saved == [0, 5] and t.steps == 6
