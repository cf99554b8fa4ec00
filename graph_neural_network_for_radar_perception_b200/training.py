"""Data-parallel training step for the radar GNN detector (the B200 replacement of the inner loop of the
reference's `train_model`, modules/neural_net/gnn/training.py:66-85, with torch.optim.SGD as configured in
modules/set_configurations/set_param_for_training_gnn.py:46).

Frames are independent graphs, so a global batch is sharded by frame: one process per GPU, each packs its frames
into one block-diagonal graph and runs the fused forward/backward kernels.  Two collectives per step, both SUM, both
enqueued on the stream -- the host never reads a device value inside a step:

  1. three counts (nodes, undirected links, clusters) as a float64 device tensor.  The reference divides every loss
     term by the count of the WHOLE batch (gnn/loss.py:58,62,66,70); each rank therefore scales its own sums by the
     global counts, which makes the per-rank losses (and gradients) add up to exactly the single-process result.  The
     counts must be known before the backward starts (a shared-trunk gradient mixes all four terms), so they cannot ride
     behind the gradients; the loss kernel reads them from device memory (rgnn_losses_fwdbwd, counts_dev).
  2. one flat fp32 buffer: the gradients (463 144 floats = 1.85 MB for the reference configuration) followed by a 5-float
     tail written by the loss kernel: the NaN flag of the reference's skip_batch and this rank's four loss shares, so the
     same all-reduce yields the global skip decision and the GLOBAL losses for logging.

The update itself is one fused kernel (rgnn_sgd_step) on the flat parameter buffer.  The communication helpers are
backend agnostic (they are covered on CPU with gloo, world size 2, in tests/test_dp_cpu.py); the compute is CUDA only.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


# -------------------------------------------------------------------------------------------------
# host-side logic (no CUDA needed)
# -------------------------------------------------------------------------------------------------
def shard_range(n_items: int, rank: int, world_size: int) -> range:
    """Contiguous block of frames owned by `rank`; block sizes differ by at most one."""
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def shard_by_edges(edge_counts: Sequence[int], world_size: int) -> List[List[int]]:
    """Greedy longest-processing-time assignment of frames to ranks, balancing directed-edge counts (the work of a
    frame is proportional to its edges).  Deterministic: ties go to the lowest rank."""
    order = sorted(range(len(edge_counts)), key=lambda i: (-edge_counts[i], i))
    load = [0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += edge_counts[i]
    return [sorted(v) for v in out]


def allreduce_counts(local_counts: Sequence[int], device, group=None) -> Tuple[int, int, int]:
    """SUM of (nodes, undirected links, clusters) over the ranks of `group` (identity without torch.distributed)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return tuple(int(c) for c in local_counts)
    t = torch.tensor([int(c) for c in local_counts], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return tuple(int(v) for v in t.tolist())


def allreduce_counts_device_(counts: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM of a 3-element count tensor that stays on its device (no host read; identity at world size 1)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def allreduce_flat_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of one contiguous buffer (no bucketing: the whole model is 1.85 MB)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


class FlatBuffers:
    """Re-homes the parameters of a module in ONE contiguous fp32 buffer (parameters become views of it) and creates
    a same-shaped flat gradient buffer whose views are installed as `.grad`, plus the SGD momentum buffer.
    Every tensor starts on a 256-byte boundary (the kernels read weights with 16-byte vector loads); the padding
    stays zero in all three buffers, so it is inert in the all-reduce and in the update."""

    ALIGN = 64   # floats

    def __init__(self, module: torch.nn.Module):
        params = [p for p in module.parameters() if p.requires_grad]
        if not params:
            raise ValueError('module has no trainable parameters')
        dev = params[0].device
        self.offsets, total = [], 0
        for p in params:
            self.offsets.append(total)
            total += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.params = params
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=dev)
        # behind the gradients: the tail the loss kernel writes (include/rgnn.h, dp_tail): [0] "skip this batch" flag (NaN
        # loss on any rank), [1..4] this rank's loss shares; all-reduced with the gradients
        self.flat_grad = torch.zeros(total + self.ALIGN, dtype=torch.float32, device=dev)
        self.tail = self.flat_grad[total:total + 5]
        self.skip_flag = self.flat_grad[total:total + 1]
        self.momentum = torch.zeros(total, dtype=torch.float32, device=dev)
        for p, off in zip(params, self.offsets):
            n = p.numel()
            self.flat_param[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + n].view(p.shape)
            p.grad = self.flat_grad[off:off + n].view(p.shape)
        self.grad_views = [p.grad for p in params]
        self._grad_ptrs = [g.data_ptr() for g in self.grad_views]
        self.numel = total

    def grad_sink(self):
        """parameter storage -> gradient view: the detector-level backward accumulates straight into the flat gradient buffer
        (no per-parameter temporary, no 184 AccumulateGrad kernels per step)."""
        return {p.data_ptr(): g for p, g in zip(self.params, self.grad_views)}

    def zero_grad(self):
        self.flat_grad.zero_()
        for p, g, gp in zip(self.params, self.grad_views, self._grad_ptrs):   # autograd may have replaced .grad; point it back
            if p.grad is None or p.grad.data_ptr() != gp:
                p.grad = g


def multistep_lr(base_lr: float, steps_done: int, milestones: Sequence[int], gamma: float = 0.1) -> float:
    """torch.optim.lr_scheduler.MultiStepLR as the reference configures it (set_param_for_training_gnn.py:50-56).
    `steps_done` counts scheduler steps since the loop started: the reference creates the scheduler afresh on resume and
    gives it milestones RELATIVE to the start iteration (`int(0.5 * max_train_iter - init_start)`), so pass
    `iteration - iter_start_offset`, not the absolute iteration."""
    return base_lr * (gamma ** sum(1 for m in milestones if steps_done >= m))


# -------------------------------------------------------------------------------------------------
# the training step (CUDA)
# -------------------------------------------------------------------------------------------------
class DataParallelTrainer:
    """One optimisation step = forward + four losses + backward (CUDA kernels) + gradient all-reduce + fused SGD.

    model: Model_Training on a CUDA device.  lr / momentum / weight_decay default to the reference's
    configuration_radarscenes_gnn.yml values (0.005 / 0.9 / 1e-4)."""

    def __init__(self, model, lr: float = 0.005, momentum: float = 0.9, weight_decay: float = 1e-4, group=None):
        self.model, self.group = model, group
        self.lr, self.mu, self.wd = float(lr), float(momentum), float(weight_decay)
        self.buffers = FlatBuffers(model)
        # the detector's backward writes parameter gradients directly into the flat buffer (see DetectorFn.backward)
        object.__setattr__(model.pred, '_rgnn_grad_sink', self.buffers.grad_sink())
        self.steps = 0
        self._counts = None

    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def step(self, gb, node_features: torch.Tensor, edge_features: torch.Tensor, labels: Dict[str, object],
             lr: Optional[float] = None):
        """gb / node_features / edge_features: this rank's packed frames (Model_Training.pack_batch or
        graph_features.build_graph_batch); labels: dict of concatenated tensors or per-frame lists.
        Returns (loss dict, accuracy dict).  The losses are those of the GLOBAL batch (identical on every rank: the four
        shares ride in the gradient all-reduce); the accuracies are this rank's.  Nothing in here reads device memory from
        the host."""
        from ._cabi import check, lib, ptr, stream_ptr
        dev = node_features.device
        b = self.buffers
        b.zero_grad()                       # also clears the tail (NaN flag, loss shares)
        counts = None
        if self.world_size > 1:
            if self._counts is None:
                self._counts = torch.zeros(3, dtype=torch.float64, device=dev)
            # scalar fills are kernel arguments: no pageable host-to-device copy, no stream drain
            for i, v in enumerate((gb.n_nodes, gb.n_und, gb.n_clusters)):
                self._counts[i].fill_(float(v))
            counts = allreduce_counts_device_(self._counts, self.group)
        loss, acc = self.model.forward_packed(gb, node_features, edge_features, labels, global_counts=counts, dp_tail=b.tail)
        total = loss['loss_node_cls'] + loss['loss_node_reg'] + loss['loss_edge_cls'] + loss['loss_obj_cls']
        total.backward()
        # reference skip_batch (gnn/training.py:40-45,79-84): a NaN loss on ANY rank skips the update on every rank; the flag
        # (written by the loss kernel) rides behind the gradients through the all-reduce
        allreduce_flat_(b.flat_grad, self.group)
        check(lib().rgnn_sgd_step_guarded(ptr(b.flat_param), ptr(b.flat_grad), ptr(b.momentum), b.numel,
                                          self.lr if lr is None else float(lr), self.mu, self.wd, 1.0,
                                          1 if self.steps == 0 else 0, ptr(b.skip_flag), stream_ptr()), 'rgnn_sgd_step_guarded')
        self.steps += 1
        table = getattr(self.model.pred, '_rgnn_table', None)
        if table is not None:
            table.invalidate_packed()       # the SGD kernel wrote the parameters behind ATen's back
        if self.world_size > 1:
            g = b.tail[1:5].clone()         # global losses (the buffer is cleared by the next step)
            loss = {k: g[i] for i, k in enumerate(_LOSS_KEYS)}
        return loss, acc

    # ---- optimizer-state checkpointing (the reference saves only detector.state_dict(), training.py:16-18) ----
    def state_dict(self) -> Dict[str, object]:
        return {'momentum': self.buffers.momentum.detach().cpu().clone(), 'steps': self.steps, 'lr': self.lr,
                'momentum_coef': self.mu, 'weight_decay': self.wd, 'numel': self.buffers.numel}

    def load_state_dict(self, sd: Dict[str, object]) -> None:
        if int(sd['numel']) != self.buffers.numel:
            raise ValueError(f"optimizer state for {sd['numel']} flat parameters, model has {self.buffers.numel}")
        self.buffers.momentum.copy_(sd['momentum'].to(self.buffers.momentum.device))
        self.steps, self.lr = int(sd['steps']), float(sd['lr'])
        self.mu, self.wd = float(sd['momentum_coef']), float(sd['weight_decay'])


# -------------------------------------------------------------------------------------------------
# loss / accuracy tracking without per-step host syncs, and the reference's training loop on top of the trainer
# -------------------------------------------------------------------------------------------------
class _DeviceSeries:
    """Append-only list of 0-dim device tensors; the mean over the entries whose weight is non-zero is taken in ONE
    device-to-host read (the reference calls .item() on every value of every step, training.py:353-365)."""

    def __init__(self):
        self.values: List[torch.Tensor] = []
        self.weights: List[torch.Tensor] = []

    def append(self, value: torch.Tensor, weight: Optional[torch.Tensor] = None) -> None:
        v = value.detach().to(torch.float64).reshape(())
        self.values.append(v)
        self.weights.append(torch.ones_like(v) if weight is None else weight.detach().to(torch.float64).reshape(()))

    def mean(self) -> float:
        if not self.values:
            return float('nan')
        v, w = torch.stack(self.values), torch.stack(self.weights)
        ws = w.sum()
        m = torch.where(ws > 0, torch.where(w > 0, v, torch.zeros_like(v)).sum() / ws.clamp_min(1e-30), torch.full_like(ws, float('nan')))
        return float(m.item())          # the one device-to-host read

    def reset(self) -> None:
        self.values, self.weights = [], []

    def tolist(self) -> List[float]:
        return torch.stack(self.values).cpu().tolist() if self.values else []


_LOSS_KEYS = ('loss_node_cls', 'loss_node_reg', 'loss_edge_cls', 'loss_obj_cls')


class LossTracker:
    """Reference interface (gnn/training.py:336-400); entries are kept on the device.  The reference only records steps
    with `total_loss > 0` (training.py:89): here that test is a device-side weight."""

    def __init__(self):
        self._history = _DeviceSeries()
        self._train = {k: _DeviceSeries() for k in ('total',) + _LOSS_KEYS}
        self._val = {k: _DeviceSeries() for k in ('total',) + _LOSS_KEYS}

    @property
    def loss_history(self) -> List[float]:
        return self._history.tolist()

    @staticmethod
    def _append(series, total_loss, losses):
        w = (total_loss.detach() > 0).to(torch.float64)
        series['total'].append(total_loss, w)
        for k in _LOSS_KEYS:
            series[k].append(losses[k], w)

    def append_training_loss_for_tb(self, total_loss, losses):
        self._append(self._train, total_loss, losses)
        self._history.append(total_loss)

    def append_validation_loss_for_tb(self, total_loss, losses):
        self._append(self._val, total_loss, losses)

    def reset_training_loss_for_tb(self):
        for s in self._train.values():
            s.reset()

    def reset_validation_loss_for_tb(self):
        for s in self._val.values():
            s.reset()

    def compute_avg_training_loss(self):
        return tuple(self._train[k].mean() for k in ('total',) + _LOSS_KEYS)

    def compute_avg_val_loss(self):
        return tuple(self._val[k].mean() for k in ('total',) + _LOSS_KEYS)


_ACC_KEYS = ('segment_accuracy', 'edge_accuracy', 'object_accuracy')


class AccuracyTracker:
    """Reference interface (gnn/training.py:403-450), device-side like LossTracker."""

    def __init__(self):
        self._train = {k: _DeviceSeries() for k in _ACC_KEYS}
        self._val = {k: _DeviceSeries() for k in _ACC_KEYS}

    def append_training_acc_for_tb(self, accuracy):
        for k in _ACC_KEYS:
            self._train[k].append(accuracy[k])

    def append_validation_acc_for_tb(self, accuracy):
        for k in _ACC_KEYS:
            self._val[k].append(accuracy[k])

    def reset_training_acc_for_tb(self):
        for s in self._train.values():
            s.reset()

    def reset_validation_acc_for_tb(self):
        for s in self._val.values():
            s.reset()

    def compute_avg_training_acc(self):
        return tuple(self._train[k].mean() for k in _ACC_KEYS)

    def compute_avg_val_acc(self):
        return tuple(self._val[k].mean() for k in _ACC_KEYS)


def train_model(detector, trainer: DataParallelTrainer, lr_milestones: Sequence[int], dataloader_train, dataloader_val, tb_writer,
                max_iters: int, log_period: int, val_period: int, iter_start_offset: int = 0, save_fn=None, base_lr: Optional[float] = None):
    """The reference's `train_model` (gnn/training.py:48-186) on top of DataParallelTrainer: same iteration structure, logging
    and validation cadence, but no host synchronisation inside a training step (no loss.item(), no isnan() on the host);
    under data parallelism the logged training losses are those of the global batch.  LR milestones count from
    `iter_start_offset`, as the reference's freshly created scheduler does on resume.
    dataloader_*: iterables of (graph_features, labels) in the reference's collate format (datagen_gnn.py:143-190:
    dict with 'node_features_dyn', 'edge_features_dyn', 'edge_index_dyn', 'adj_matrix_dyn').  tb_writer may be None;
    save_fn(detector, trainer, iteration) is called where the reference saves weights."""
    loss_tracker, acc_tracker = LossTracker(), AccuracyTracker()
    base_lr = trainer.lr if base_lr is None else base_lr
    it_train = iter(dataloader_train)
    rank0 = not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0
    for it in range(iter_start_offset, max_iters):
        graph_features, labels = next(it_train)
        if graph_features is not None:
            detector.train()
            gb, nf, ef = detector.pack_batch(graph_features['node_features_dyn'], graph_features['edge_features_dyn'],
                                             graph_features['edge_index_dyn'], labels['cluster_node_idx'])
            loss, accuracy = trainer.step(gb, nf, ef, labels, lr=multistep_lr(base_lr, it - iter_start_offset, lr_milestones))
            total = sum(loss.values())
            loss_tracker.append_training_loss_for_tb(total, loss)
            acc_tracker.append_training_acc_for_tb(accuracy)
            if it % log_period == 0 and rank0:
                print(f'[Iter {it}][loss: {float(total.detach()):.5f}]' + ''.join(f'[{k}: {float(v.detach()):.5f}]' for k, v in loss.items()))
        if (it % val_period == 0) or (it == max_iters - 1):
            if save_fn is not None and rank0:
                save_fn(detector, trainer, it)
            detector.eval()
            with torch.no_grad():
                for graph_features, labels in dataloader_val:
                    if graph_features is None:
                        continue
                    loss, accuracy = detector(node_features=graph_features['node_features_dyn'],
                                              edge_features=graph_features['edge_features_dyn'],
                                              edge_index=graph_features['edge_index_dyn'],
                                              adj_matrix=graph_features['adj_matrix_dyn'], labels=labels)
                    loss_tracker.append_validation_loss_for_tb(sum(loss.values()), loss)
                    acc_tracker.append_validation_acc_for_tb(accuracy)
            tr, va = loss_tracker.compute_avg_training_loss(), loss_tracker.compute_avg_val_loss()
            tra, vaa = acc_tracker.compute_avg_training_acc(), acc_tracker.compute_avg_val_acc()
            if tb_writer is not None and rank0:
                for name, i in (('Total_Loss', 0), ('Loss_Node_Segmentation', 1), ('Loss_Node_Offset', 2),
                                ('Loss_Edge_Classification', 3), ('Loss_Object_Classification', 4)):
                    tb_writer.add_scalars(name, {'train': tr[i], 'val': va[i]}, it)
                for name, i in (('Acc_Node_Segmentation', 0), ('Acc_Edge_Classification', 1), ('Acc_Object_Classification', 2)):
                    tb_writer.add_scalars(name, {'train': tra[i], 'val': vaa[i]}, it)
            loss_tracker.reset_training_loss_for_tb(); loss_tracker.reset_validation_loss_for_tb()
            acc_tracker.reset_training_acc_for_tb(); acc_tracker.reset_validation_acc_for_tb()
    return loss_tracker
