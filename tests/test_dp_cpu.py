"""CPU, world size 2, gloo: the host-side logic of data-parallel training (training.py) -- frame sharding, the
global-count all-reduce that reproduces the reference's whole-batch loss normalisation (gnn/loss.py:58,62,66,70),
and the flat gradient all-reduce.  The per-rank model arithmetic is played by the oracle here (tests may use it);
the CUDA step itself is covered by tests/test_train_gpu.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graph_neural_network_for_radar_perception_b200 import training as tr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def test_shard_range_partitions_exactly():
    for n in (1, 2, 7, 64, 65):
        for w in (1, 2, 3, 8):
            parts = [list(tr.shard_range(n, r, w)) for r in range(w)]
            assert sorted(sum(parts, [])) == list(range(n))
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1


def test_shard_by_edges_balances_and_is_deterministic():
    rng = np.random.default_rng(0)
    edges = [int(v) for v in rng.integers(1000, 40000, size=64)]
    a = tr.shard_by_edges(edges, 8)
    assert a == tr.shard_by_edges(edges, 8)
    assert sorted(sum(a, [])) == list(range(64))
    loads = [sum(edges[i] for i in part) for part in a]
    assert max(loads) - min(loads) <= max(edges)


def test_multistep_lr_matches_torch():
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=0.005)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[5, 8], gamma=0.1)
    for it in range(12):
        assert abs(tr.multistep_lr(0.005, it, [5, 8]) - opt.param_groups[0]['lr']) < 1e-12
        opt.step()
        sched.step()


def test_flat_buffers_alias_parameters():
    m = torch.nn.Sequential(torch.nn.Linear(3, 4), torch.nn.Linear(4, 2))
    before = [p.detach().clone() for p in m.parameters()]
    fb = tr.FlatBuffers(m)
    assert fb.numel >= sum(p.numel() for p in m.parameters()) and fb.numel % tr.FlatBuffers.ALIGN == 0
    for p, b in zip(m.parameters(), before):
        assert torch.equal(p.detach(), b)
    fb.flat_param.mul_(2.0)                      # the parameters ARE the flat buffer
    for p, b in zip(m.parameters(), before):
        assert torch.equal(p.detach(), 2 * b)
    m(torch.ones(5, 3)).sum().backward()         # autograd accumulates into the flat gradient buffer
    assert float(fb.flat_grad.abs().sum()) > 0
    for p, off in zip(m.parameters(), fb.offsets):
        assert p.grad.data_ptr() == fb.flat_grad[off:].data_ptr() and off % tr.FlatBuffers.ALIGN == 0
    fb.zero_grad()
    assert float(fb.flat_grad.abs().sum()) == 0


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _frame_inputs(g, i):
    nf = torch.from_numpy(g[f'f{i}_node_features'])
    ef = torch.from_numpy(g[f'f{i}_edge_features'])
    ei = torch.from_numpy(g[f'f{i}_edge_index'])
    ptr, mem = g[f'f{i}_cluster_ptr'], g[f'f{i}_cluster_members']
    lab = {'cluster_node_idx': [[torch.from_numpy(mem[ptr[j]:ptr[j + 1]]) for j in range(len(ptr) - 1)]]}
    for k in ('cluster_labels', 'edge_class', 'node_class', 'node_offsets'):
        lab[k] = [torch.from_numpy(g[f'f{i}_{k}'])]
    return nf, ef, ei, lab


def _dp_worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, ROOT)
        from oracle import model_torch as mt
        g = np.load(os.path.join(GOLDEN, 'train_2frames.npz'))
        ck = torch.load(os.path.join(GOLDEN, 'graph_based_detector.pt'), map_location='cpu', weights_only=True)
        frames = list(tr.shard_range(2, rank, world))
        assert frames == [rank]
        nf, ef, ei, lab = _frame_inputs(g, frames[0])
        sd = {k: v.clone().requires_grad_(True) for k, v in ck.items()}
        loss, _, outs = mt.training_forward(sd, [nf], [ef], [ei], lab)
        local = (outs[0].shape[0], outs[2].shape[0], outs[3].shape[0])
        glob = tr.allreduce_counts(local, 'cpu')
        # this rank's share of the whole-batch loss: local sums divided by the GLOBAL counts
        share = {'loss_node_cls': loss['loss_node_cls'] * local[0] / glob[0],
                 'loss_node_reg': loss['loss_node_reg'] * local[0] / glob[0],
                 'loss_edge_cls': loss['loss_edge_cls'] * local[1] / glob[1],
                 'loss_obj_cls': loss['loss_obj_cls'] * local[2] / glob[2]}
        sum(share.values()).backward()
        names = sorted(sd)
        flat = torch.cat([sd[k].grad.reshape(-1) for k in names])
        tr.allreduce_flat_(flat)
        tot = torch.stack([share[k].detach() for k in sorted(share)])
        dist.all_reduce(tot)
        if rank == 0:
            torch.save({'counts': glob, 'loss': tot, 'flat': flat, 'names': names}, out_path)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_losses_and_gradients_add_up_to_the_single_process_batch(tmp_path):
    """Frames sharded over 2 gloo ranks; global counts + SUM all-reduce reproduce the reference's 2-frame batch."""
    out = str(tmp_path / 'dp.pt')
    mp.spawn(_dp_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out, weights_only=False)
    g = np.load(os.path.join(GOLDEN, 'train_2frames.npz'))
    n_nodes = sum(g[f'f{i}_node_features'].shape[0] for i in range(2))
    assert res['counts'][0] == n_nodes
    want = [float(g[k]) for k in sorted(('loss_node_cls', 'loss_node_reg', 'loss_edge_cls', 'loss_obj_cls'))]
    got = res['loss'].tolist()
    for a, b in zip(got, want):
        assert abs(a - b) <= 1e-5 * abs(b) + 1e-7, (got, want)
    off = 0
    ck = torch.load(os.path.join(GOLDEN, 'graph_based_detector.pt'), map_location='cpu', weights_only=True)
    checked = 0
    for n in res['names']:
        k = ck[n].numel()
        key = 'grad::' + n
        if key in g.files:
            ref = g[key].reshape(-1)
            got_g = res['flat'][off:off + k].numpy()
            assert np.abs(got_g - ref).max() <= 1e-4 * np.abs(ref).max() + 1e-8, n
            checked += 1
        off += k
    assert checked >= 8


def test_device_side_trackers_match_reference_semantics():
    """LossTracker / AccuracyTracker (reference gnn/training.py:336-450): means over the recorded steps, steps with
    total_loss <= 0 are not recorded (training.py:89) -- here a device-side weight instead of a host-side `if`."""
    lt, at = tr.LossTracker(), tr.AccuracyTracker()
    steps = [(2.0, (0.5, 0.5, 0.5, 0.5)), (0.0, (0.0, 0.0, 0.0, 0.0)), (4.0, (1.0, 1.0, 1.0, 1.0))]
    for total, parts in steps:
        losses = {k: torch.tensor(v) for k, v in zip(('loss_node_cls', 'loss_node_reg', 'loss_edge_cls', 'loss_obj_cls'), parts)}
        lt.append_training_loss_for_tb(torch.tensor(total), losses)
        at.append_training_acc_for_tb({'segment_accuracy': torch.tensor(0.5), 'edge_accuracy': torch.tensor(1.0),
                                       'object_accuracy': torch.tensor(total / 4)})
    avg = lt.compute_avg_training_loss()
    assert avg[0] == pytest.approx(3.0) and avg[1] == pytest.approx(0.75)        # the zero-loss step is excluded
    assert lt.loss_history == [2.0, 0.0, 4.0]
    assert at.compute_avg_training_acc() == pytest.approx((0.5, 1.0, 0.5))
    lt.reset_training_loss_for_tb()
    assert np.isnan(lt.compute_avg_training_loss()[0])
    assert tr.multistep_lr(0.005, 10, [5, 20]) == pytest.approx(0.0005)
