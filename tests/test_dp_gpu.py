"""GPU, 2 ranks over NCCL (skipped on a single-GPU box): the data-parallel step (training.DataParallelTrainer) -- counts
all-reduced on the stream, ONE gradient all-reduce carrying the NaN flag and the loss shares, fused SGD -- leaves every rank
with the parameters a single process gets from the same global batch, and reports the GLOBAL losses on every rank."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
SIZES = (180, 90, 140, 60)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _run_steps(rank, world, frames_idx, n_steps, device):
    """n_steps optimisation steps on the frames `frames_idx` of the 4-frame global batch; returns (flat params, losses)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from gpu_util import batch_labels, load_model, synth_batch
    from graph_neural_network_for_radar_perception_b200.training import DataParallelTrainer
    ck = torch.load(os.path.join(GOLDEN, 'graph_based_detector.pt'), map_location='cpu', weights_only=True)
    frames = synth_batch(SIZES, seed0=900)
    mine = [frames[i] for i in frames_idx]
    m = load_model(ck, device).train()
    trainer = DataParallelTrainer(m, lr=0.005, momentum=0.9, weight_decay=1e-4)
    lab = batch_labels(mine, device)
    gb, nf, ef = m.pack_batch([f['nf'].to(device) for f in mine], [f['ef'].to(device) for f in mine],
                              [f['ei'].to(device) for f in mine], lab['cluster_node_idx'])
    losses = []
    for _ in range(n_steps):
        lb = dict(lab)
        lb['node_offsets'] = [t.clone() for t in lab['node_offsets']]
        loss, _ = trainer.step(gb, nf, ef, lb)
        losses.append([float(loss[k]) for k in ('loss_node_cls', 'loss_node_reg', 'loss_edge_cls', 'loss_obj_cls')])
    return trainer.buffers.flat_param.detach().cpu().clone(), losses


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        from graph_neural_network_for_radar_perception_b200.training import shard_range
        idx = list(shard_range(len(SIZES), rank, world))
        params, losses = _run_steps(rank, world, idx, 2, torch.device('cuda', rank))
        torch.save({'params': params, 'losses': losses}, os.path.join(out_dir, f'rank{rank}.pt'))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_step_equals_single_process_step(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (gpurun --gpus 2)')
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / 'rank0.pt', weights_only=False)
    r1 = torch.load(tmp_path / 'rank1.pt', weights_only=False)
    single_params, single_losses = _run_steps(0, 1, list(range(len(SIZES))), 2, torch.device('cuda', 0))
    # identical on both ranks (same all-reduced gradients, same update)
    assert torch.equal(r0['params'], r1['params'])
    assert r0['losses'] == r1['losses']
    # the losses every rank reports are those of the GLOBAL batch
    for a, b in zip(r0['losses'], single_losses):
        np.testing.assert_allclose(a, b, rtol=2e-5, atol=1e-6)
    # and two steps leave the parameters where the single process leaves them (summation order of the gradients differs:
    # per-rank partial sums + all-reduce vs one sum; lr = 0.005 scales any gradient noise down)
    d = (r0['params'] - single_params).abs()
    assert float(d.max()) <= 1e-6 + 1e-5 * float(single_params.abs().max()), float(d.max())


def test_tensors_on_a_non_current_device_fail_loudly(ckpt_state_dict):
    """The C library launches on the current device and its current stream: a model on cuda:1 called while cuda:0 is current must
    raise, and work inside torch.cuda.device(1)."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from gpu_util import load_model, synth_batch
    from graph_neural_network_for_radar_perception_b200._cabi import RgnnError
    f = synth_batch((120,), seed0=31)[0]
    d1 = torch.device('cuda', 1)
    m = load_model(ckpt_state_dict, d1).pred.eval()
    args = (f['nf'].to(d1), f['ef'].to(d1), f['ei'].to(d1), None, [c.to(d1) for c in f['clusters']] if 'clusters' in f else [torch.arange(4, device=d1)])
    torch.cuda.set_device(0)
    with torch.no_grad():
        with pytest.raises(RgnnError):
            m(*args)
        with torch.cuda.device(1):
            out = m(*args)
        m0 = load_model(ckpt_state_dict, 'cuda:0').pred.eval()
        ref = m0(*[a.to('cuda:0') if isinstance(a, torch.Tensor) else ([c.to('cuda:0') for c in a] if a is not None else None) for a in args])
    for a, b in zip(out, ref):
        assert torch.equal(a.cpu(), b.cpu())

