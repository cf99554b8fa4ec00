"""GPU: detector forward through the C-ABI against the reference fixtures (tests/golden) and the torch oracle.
Tolerance: north_star asks fp32 outputs within rtol 1e-4.  Elements near zero need an absolute floor; it is stated
relative to each output tensor's scale: |got - want| <= 1e-4 |want| + 1e-5 max|want| (logits reach ~10, so ~1e-4;
offsets ~0.7, so ~7e-6).  Measured on B200 (tools/fwd_diag.py): 3xTF32 tensor-core path rms error 3e-6 / max 3.5e-5
on logits of magnitude 10 (the tensor core accumulates in fp32 with truncation, ~4x the FFMA path's 8e-7 / 5.7e-6)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from gpu_util import assert_close, clusters_from, load_model, scaled_atol

RTOL = 1e-4


@pytest.mark.parametrize('case', ['n48', 'n200'])
def test_inference_forward_matches_reference_fixture(golden_dir, ckpt_state_dict, case):
    g = np.load(os.path.join(golden_dir, f'model_{case}.npz'))
    m = load_model(ckpt_state_dict).pred.eval()
    dev = 'cuda'
    with torch.no_grad():
        out = m(torch.from_numpy(g['node_features']).to(dev), torch.from_numpy(g['edge_features']).to(dev),
                torch.from_numpy(g['edge_index']).to(dev), None,
                clusters_from(g['cluster_ptr'], g['cluster_members'], dev))
    for o, k in zip(out, ['node_cls', 'node_off', 'link_cls', 'obj_cls']):
        assert o.shape == g[k].shape, k
        assert_close(o.cpu().numpy(), g[k], RTOL, scaled_atol(g[k]), k)


def test_blocks_standalone_match_oracle(ckpt_state_dict):
    """graph_feature_encoding / heads / conv block used directly, as the reference's gnn_blocks API allows."""
    from oracle import model_torch as mt
    torch.manual_seed(1)
    m = load_model(ckpt_state_dict).pred.eval()
    sd = ckpt_state_dict
    x6 = torch.randn(333, 6)
    e7 = torch.randn(1000, 7)
    with torch.no_grad():
        a = m.encode_node_feat(x6.cuda()).cpu()
        b = mt.ffn_stack(sd, 'pred.encode_node_feat.encoder', x6)
        assert_close(a.numpy(), b.numpy(), RTOL, scaled_atol(b), 'node encoder')
        a = m.encode_edge_feat(e7.cuda()).cpu()
        b = mt.ffn_stack(sd, 'pred.encode_edge_feat.encoder', e7)
        assert_close(a.numpy(), b.numpy(), RTOL, scaled_atol(b), 'edge encoder')
        x = torch.randn(333, 64)
        a = m.predict_node(x.cuda()).cpu()
        b = mt.task_head(sd, 'pred.predict_node.pred_cls', mt.ffn_stack(sd, 'pred.predict_node.stem', x))
        assert_close(a.numpy(), b.numpy(), RTOL, scaled_atol(b), 'node head')
        a = m.predict_node.stem[0](x.cuda()).cpu()
        b = mt.ffn(sd, 'pred.predict_node.stem.0.block', x)
        assert_close(a.numpy(), b.numpy(), RTOL, scaled_atol(b), 'single ffn_block')
        # conv block on a random (non symmetric, unsorted) edge list
        ei = torch.randint(0, 333, (2, 2000))
        e = torch.randn(2000, 64)
        a = m.pass_messages.conv_blk[0](x.cuda(), e.cuda(), ei.cuda()).cpu()
        b = mt.conv_block(sd, 'pred.pass_messages.conv_blk.0', x, e, ei)
        assert_close(a.numpy(), b.numpy(), RTOL, scaled_atol(b, 2e-5), 'conv block')


def test_batched_packing_is_bit_identical(ckpt_state_dict):
    """rgnn_pack_detector with the batched pack queue (one table-driven launch per kind and 48 images, csrc/rgnn_pack.cuh) writes the
    same bytes as one launch per image: entries of a batch run concurrently, so this also guards the one image that is written
    twice (layer 1 of the edge encoder, re-packed in K blocks)."""
    from graph_neural_network_for_radar_perception_b200._cabi import check, lib, stream_ptr
    from graph_neural_network_for_radar_perception_b200._engine import detector_table
    m = load_model(ckpt_state_dict).pred.eval()
    table = detector_table(m)
    images = []
    try:
        for batched in (0, 1, 1, 0):
            check(lib().rgnn_set_option(b'pack_batch', batched), 'opt')
            table.packed.fill_(123.0)
            table.invalidate_packed()
            table.ensure_packed(stream_ptr())
            torch.cuda.synchronize()
            images.append(table.packed.view(torch.int32).clone())
    finally:
        check(lib().rgnn_set_option(b'pack_batch', 1), 'opt')
    assert (images[0] != 123.0).any()
    for im in images[1:]:
        assert torch.equal(im, images[0])


def test_standalone_object_head_segment_max_matches_torch(ckpt_state_dict):
    """object_classification used on its own (gnn_blocks.py:347-389): the per-cluster max (rgnn_segment_max_fwd / _bwd) against
    torch.max(x[idx], dim=0) per cluster as the reference writes it, values and the gradient routed to the arg-max rows."""
    torch.manual_seed(3)
    m = load_model(ckpt_state_dict).pred.train()
    head = m.predict_class
    n = 500
    x = torch.randn(n, 64)
    perm = torch.randperm(n)
    cuts = [0, 1, 2, 40, 41, 300, 500]
    clusters = [perm[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    xg = x.cuda().requires_grad_(True)
    out = head(xg, [c.cuda() for c in clusters])
    w = torch.randn_like(out)
    (out * w).sum().backward()
    from graph_neural_network_for_radar_perception_b200._engine import apply_stack
    xr = x.cuda().requires_grad_(True)
    g = apply_stack(xr, list(head.stem))
    pooled = torch.stack([torch.max(g[c.cuda()], dim=0)[0] for c in clusters])
    ref = apply_stack(pooled, list(head.pred_cls.head))
    (ref * w).sum().backward()
    assert torch.equal(out, ref)
    assert_close(xg.grad.cpu().numpy(), xr.grad.cpu().numpy(), 1e-5, 1e-7, 'dL/dx through the segment max')


def test_standalone_conv_block_backward_matches_oracle_autograd(ckpt_state_dict):
    """residual_graph_conv_block used on its own under torch autograd (rgnn_conv_block_bwd): gradients w.r.t. x, the edge features
    (caller's edge order) and every parameter of the block against float64 autograd of the oracle, on a random unsorted edge list."""
    from oracle import model_torch as mt
    torch.manual_seed(11)
    m = load_model(ckpt_state_dict).pred.train()
    blk = m.pass_messages.conv_blk[2]
    stem = 'pred.pass_messages.conv_blk.2'
    n, ne = 700, 6000
    x, e = torch.randn(n, 64), torch.randn(ne, 64)
    ei = torch.randint(0, n, (2, ne))
    w = torch.randn(n, 64) / 8          # loss = sum(out * w)
    xg, eg = x.cuda().requires_grad_(True), e.cuda().requires_grad_(True)
    for p_ in blk.parameters():
        p_.grad = None
    out = blk(xg, eg, ei.cuda())
    (out * w.cuda()).sum().backward()
    sd64 = {k: v.double().requires_grad_(k.startswith(stem)) for k, v in ckpt_state_dict.items()}
    x64, e64 = x.double().requires_grad_(True), e.double().requires_grad_(True)
    (mt.conv_block(sd64, stem, x64, e64, ei) * w.double()).sum().backward()

    def close(got, want, what):
        want = want.numpy()
        assert_close(got.detach().cpu().numpy(), want, 2e-4, 2e-5 * max(np.abs(want).max(), 1e-12), what)
    close(xg.grad, x64.grad, 'dL/dx')
    close(eg.grad, e64.grad, 'dL/de')
    names = dict(blk.named_parameters())
    assert len(names) == 12
    for k, p_ in names.items():
        close(p_.grad, sd64[f'{stem}.{k}'].grad, k)


@pytest.mark.parametrize('n,e,layer', [(5000, 100_000, 0), (1500, 151_555, 3), (40_000, 420_000, 6)])
def test_conv_block_many_tiles_per_cta_matches_oracle(ckpt_state_dict, n, e, layer):
    """The persistent message kernel in its steady state: 780 - 3 280 tiles of 128 edges over 148 CTAs (5 - 22 tiles per CTA,
    ragged last tile), so that every hand-over of its pipelined tile loop (GEMM1 of the next tile under epilogue 2, the
    staging / segmented-sum helper warps, segments cut by tile boundaries, nodes without incoming edges) runs many
    times; skewed target degrees (a few nodes own hundreds of edges, i.e. segments longer than a tile)."""
    from oracle import model_torch as mt
    g = torch.Generator().manual_seed(7 * n + layer)
    m = load_model(ckpt_state_dict).pred.eval()
    x = torch.randn(n, 64, generator=g)
    emb = torch.randn(e, 64, generator=g)
    src = torch.randint(0, n, (e,), generator=g)
    tgt = torch.randint(0, n, (e,), generator=g)
    heavy = torch.randint(0, n, (4,), generator=g)
    tgt[: e // 50] = heavy[torch.randint(0, 4, (e // 50,), generator=g)]        # 2 % of the edges land on four nodes
    tgt[tgt == 1] = 2                                                          # node 1 receives nothing
    ei = torch.stack((src, tgt))
    with torch.no_grad():
        a = m.pass_messages.conv_blk[layer](x.cuda(), emb.cuda(), ei.cuda()).cpu()
        b = mt.conv_block(ckpt_state_dict, f'pred.pass_messages.conv_blk.{layer}', x, emb, ei)
    assert_close(a.numpy(), b.numpy(), RTOL, scaled_atol(b, 2e-5), f'conv block {layer}')


@pytest.mark.parametrize('n_frames,n_pts,knn', [(3, 150, 10), (2, 1000, 10), (4, 3000, 10), (1, 1200, 64)])
def test_batched_forward_matches_oracle_per_frame(ckpt_state_dict, n_frames, n_pts, knn):
    """Block-diagonal batch == the reference's per-frame loop (checked against the oracle frame by frame).  (4, 3000) is the
    BASELINE.json configs[1] frame shape (4 x ~3000 points, E ~ 137 k: every CTA of every persistent kernel runs several
    tiles); (1, 1200, k=64) the dense end of the configs[4] sweep (degree up to ~100, segments longer than a tile row
    group)."""
    from graph_neural_network_for_radar_perception_b200 import graph_features as gf, synth
    from oracle import graph_np, model_torch as mt
    m = load_model(ckpt_state_dict).pred.eval()
    frames, labs = [], []
    for i in range(n_frames):
        d, src = synth.make_frame(500 + i, n_pts + 13 * i, knn=knn)
        frames.append(d)
        labs.append(src)
    pts, fp = gf.frames_to_device(frames)
    R = np.float64(np.sqrt(100.0 ** 2 + 50.0 ** 2))
    bf = gf.build_graph_batch(pts, fp, 25, knn, max_range=R, max_azimuth=np.pi * 0.5)
    assert int(bf.gb.n_und_dev.item()) == bf.gb.n_und          # symmetric adjacency: E / 2 links, no read-back needed
    cl_lists, o_out = [], []
    for i, d in enumerate(frames):
        adj = graph_np.adjacency_information(d, 25, knn)
        lab = synth.make_labels(d, labs[i], adj['adj_list'])
        cl = [torch.from_numpy(c) for c in lab['cluster_node_idx']]
        cl_lists.append(cl)
        nf = torch.from_numpy(graph_np.node_features(d, adj['degree'], True, 0, R, 0, np.pi * 0.5).astype(np.float32))
        ef = torch.from_numpy(graph_np.edge_features(d, adj['adj_list']).astype(np.float32))
        with torch.no_grad():
            o_out.append(mt.detector_forward(ckpt_state_dict, nf, ef, torch.from_numpy(adj['adj_list']), cl))
    bf.gb.set_clusters(cl_lists, fp[:-1], 'cuda')
    with torch.no_grad():
        out = m.forward_batch(bf.gb, bf.node_features, bf.edge_features)
    for k in range(4):
        want = torch.cat([o[k] for o in o_out]).numpy()
        assert out[k].shape == want.shape
        assert_close(out[k].cpu().numpy(), want, RTOL, scaled_atol(want), f'output {k}')


def test_reduced_precision_mode_stated_tolerance(ckpt_state_dict):
    """BASELINE.json north_star: 'bf16 variants within a stated tolerance'.  The reduced-precision mode of this library is
    rgnn_set_option("f16_passes", 1): the fp16-split kernels (edge encoder, 7 message kernels, stems / heads) run ONE pass
    with plain fp16 operands and fp32 accumulation (~2.5e-4 relative per GEMM, tools/mma_noise.py) instead of the three
    passes of the fp32-parity mode.  Stated tolerance: every output within 5e-3 of its tensor's largest magnitude against
    the float32 oracle, on the two reference fixtures' model and a 3000-point frame; arg-max decisions of the node / link /
    object heads agree on >= 99.5 % of the rows."""
    from graph_neural_network_for_radar_perception_b200 import graph_features as gf, synth
    from graph_neural_network_for_radar_perception_b200._cabi import check, lib
    from oracle import graph_np, model_torch as mt
    m = load_model(ckpt_state_dict).pred.eval()
    R = np.float64(np.sqrt(100.0 ** 2 + 50.0 ** 2))
    check(lib().rgnn_set_option(b'f16_passes', 1), 'opt')
    try:
        for seed, n in ((31, 200), (32, 3000)):
            d, src = synth.make_frame(seed, n)
            adj = graph_np.adjacency_information(d, 25, 10)
            lab = synth.make_labels(d, src, adj['adj_list'])
            cl = [torch.from_numpy(c) for c in lab['cluster_node_idx']]
            nf = torch.from_numpy(graph_np.node_features(d, adj['degree'], True, 0, R, 0, np.pi * 0.5).astype(np.float32))
            ef = torch.from_numpy(graph_np.edge_features(d, adj['adj_list']).astype(np.float32))
            ei = torch.from_numpy(adj['adj_list'])
            with torch.no_grad():
                want = mt.detector_forward(ckpt_state_dict, nf, ef, ei, cl)
                got = m(nf.cuda(), ef.cuda(), ei.cuda(), None, [c.cuda() for c in cl])
            for g, w, name in zip(got, want, ('node_cls', 'node_off', 'link_cls', 'obj_cls')):
                g, w = g.cpu().numpy(), w.numpy()
                assert np.abs(g - w).max() <= 5e-3 * np.abs(w).max(), (name, n, np.abs(g - w).max() / np.abs(w).max())
                if name != 'node_off':
                    assert (g.argmax(-1) == w.argmax(-1)).mean() >= 0.995, (name, n)
    finally:
        check(lib().rgnn_set_option(b'f16_passes', 3), 'opt')
    # and the fp32-parity mode is back
    assert lib().rgnn_get_option(b'f16_passes') == 3


def test_rows_prefetch_is_bit_identical(ckpt_state_dict):
    """rgnn_set_option("rows_prefetch", 0 / 1): the next-tile L2 prefetch of the row-owning kernels (node update, 64-wide chains, edge
    encoder) is a hint, the four outputs must not change in a single bit.  32 x 3000 points = 96 k nodes / 1.17 M edges: every CTA of
    the node kernels runs several tiles per group (the prefetch and the two-tiles-ahead index registers are live), and the last tiles are
    ragged."""
    from graph_neural_network_for_radar_perception_b200 import graph_features as gf, synth
    from graph_neural_network_for_radar_perception_b200._cabi import check, lib
    m = load_model(ckpt_state_dict).pred.eval()
    base = [synth.make_frame(700 + i, 3000 - 7 * i, knn=10)[0] for i in range(4)]
    frames = [base[i % 4] for i in range(32)]
    pts, fp = gf.frames_to_device(frames)
    R = np.float64(np.sqrt(100.0 ** 2 + 50.0 ** 2))
    bf = gf.build_graph_batch(pts, fp, 25, 10, max_range=R, max_azimuth=np.pi * 0.5)
    bf.gb.set_clusters([[torch.arange(i, min(i + 5, n)) for i in range(0, n, 5)] for n in np.diff(fp)], fp[:-1], 'cuda')
    outs = {}
    try:
        for v in (0, 1):
            check(lib().rgnn_set_option(b'rows_prefetch', v), 'opt')
            with torch.no_grad():
                outs[v] = [o.clone() for o in m.forward_batch(bf.gb, bf.node_features, bf.edge_features)]
    finally:
        check(lib().rgnn_set_option(b'rows_prefetch', 1), 'opt')
    assert lib().rgnn_get_option(b'rows_prefetch') == 1
    for a, b in zip(outs[0], outs[1]):
        assert a.shape == b.shape and a.numel() > 0 and torch.isfinite(a).all()
        assert torch.equal(a, b)


def test_conv_layer_entry_points_match_oracle(ckpt_state_dict):
    """The layer-level C-ABI entries bench.py times (rgnn_split_edge_embedding, rgnn_conv_layer_f16_fwd): one residual conv
    block on a random symmetric-free edge list == oracle conv_block, plus the next block's hoisted projection."""
    import ctypes as C
    from graph_neural_network_for_radar_perception_b200._cabi import check, lib, ptr, stream_ptr
    from graph_neural_network_for_radar_perception_b200._engine import GraphBatch, detector_table
    from oracle import model_torch as mt
    torch.manual_seed(5)
    m = load_model(ckpt_state_dict).pred.eval()
    n, e = 3000, 40_000
    x, emb = torch.randn(n, 64), torch.randn(e, 64)
    ei = torch.stack((torch.randint(0, n, (e,)), torch.randint(0, n, (e,))))
    gb = GraphBatch.from_edge_index(ei.cuda(), n)
    table = detector_table(m)
    table.refill(None)
    table.ensure_packed(stream_ptr())
    g = gb.c_struct()
    # the kernels work in target-major order: row k of the embedding belongs to edge perm[k] of the caller's list
    emb_tm = emb.cuda()[gb.perm.long()].contiguous()
    es = torch.empty(lib().rgnn_split_edge_embedding_words(e), dtype=torch.int32, device='cuda')
    check(lib().rgnn_split_edge_embedding(ptr(emb_tm), e, ptr(es), stream_ptr()), 'split')
    # projection of layer 3 from x (through the stand-alone block API), then the layer entry
    proj = torch.empty(n, 256, device='cuda')
    out, agg, proj_next = torch.empty(n, 64, device='cuda'), torch.empty(n, 64, device='cuda'), torch.empty(n, 256, device='cuda')
    xc = x.cuda()
    sd = ckpt_state_dict
    W0 = sd['pred.pass_messages.conv_blk.3.msg.0.block.0.weight']
    b0 = sd['pred.pass_messages.conv_blk.3.msg.0.block.0.bias']
    proj.copy_(torch.cat((x @ W0[:, :64].T + b0, x @ W0[:, 64:128].T), dim=1).cuda())
    check(lib().rgnn_conv_layer_f16_fwd(C.byref(table.det.conv[3]), C.byref(table.det.conv[4]), C.byref(g), ptr(xc), ptr(es), ptr(proj),
                                        ptr(out), ptr(agg), ptr(proj_next), stream_ptr()), 'layer')
    want = mt.conv_block(sd, 'pred.pass_messages.conv_blk.3', x, emb, ei)
    assert_close(out.cpu().numpy(), want.numpy(), RTOL, scaled_atol(want, 2e-5), 'layer output')
    W4 = sd['pred.pass_messages.conv_blk.4.msg.0.block.0.weight']
    b4 = sd['pred.pass_messages.conv_blk.4.msg.0.block.0.bias']
    want_p = torch.cat((want @ W4[:, :64].T + b4, want @ W4[:, 64:128].T), dim=1)
    assert_close(proj_next.cpu().numpy(), want_p.numpy(), RTOL, scaled_atol(want_p, 2e-5), 'next projection')
