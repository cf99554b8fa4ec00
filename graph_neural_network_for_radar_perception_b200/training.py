"""Data-parallel training step for the radar GNN detector (the B200 replacement of the inner loop of the
reference's `train_model`, modules/neural_net/gnn/training.py:66-85, with torch.optim.SGD as configured in
modules/set_configurations/set_param_for_training_gnn.py:46).

Frames are independent graphs, so a global batch is sharded by frame: one process per GPU, each packs its frames
into one block-diagonal graph and runs the fused forward/backward kernels.  Two collectives per step, both SUM:

  1. three int64 counts (nodes, undirected links, clusters).  The reference divides every loss term by the count of
     the WHOLE batch (gnn/loss.py:58,62,66,70); each rank therefore scales its own sums by the global counts, which
     makes the per-rank losses (and gradients) add up to exactly the single-process result.  Averaging per-rank
     mean-losses would be wrong whenever ranks hold different numbers of nodes / links / clusters.
  2. one flat fp32 gradient buffer (463 144 floats = 1.85 MB for the reference configuration) over NCCL/NVLink.

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
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.momentum = torch.zeros(total, dtype=torch.float32, device=dev)
        for p, off in zip(params, self.offsets):
            n = p.numel()
            self.flat_param[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + n].view(p.shape)
            p.grad = self.flat_grad[off:off + n].view(p.shape)
        self.numel = total

    def zero_grad(self):
        self.flat_grad.zero_()
        for p, off in zip(self.params, self.offsets):   # autograd may have replaced .grad; point it back at the flat buffer
            g = self.flat_grad[off:off + p.numel()].view(p.shape)
            if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                p.grad = g


def multistep_lr(base_lr: float, iteration: int, milestones: Sequence[int], gamma: float = 0.1) -> float:
    """torch.optim.lr_scheduler.MultiStepLR as the reference configures it (set_param_for_training_gnn.py:50-56)."""
    return base_lr * (gamma ** sum(1 for m in milestones if iteration >= m))


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
        self.steps = 0

    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def step(self, gb, node_features: torch.Tensor, edge_features: torch.Tensor, labels: Dict[str, object],
             lr: Optional[float] = None):
        """gb / node_features / edge_features: this rank's packed frames (Model_Training.pack_batch or
        graph_features.build_graph_batch); labels: dict of concatenated tensors or per-frame lists.
        Returns (loss dict, accuracy dict) of THIS rank's share; summing the losses over ranks gives the global loss."""
        from ._cabi import check, lib, ptr, stream_ptr
        dev = node_features.device
        n_obj = gb.n_clusters
        counts = allreduce_counts((gb.n_nodes, gb.n_und, n_obj), dev, self.group)
        self.model.global_counts = counts
        self.buffers.zero_grad()
        loss, acc = self.model.forward_packed(gb, node_features, edge_features, labels)
        total = loss['loss_node_cls'] + loss['loss_node_reg'] + loss['loss_edge_cls'] + loss['loss_obj_cls']
        total.backward()
        allreduce_flat_(self.buffers.flat_grad, self.group)
        b = self.buffers
        check(lib().rgnn_sgd_step(ptr(b.flat_param), ptr(b.flat_grad), ptr(b.momentum), b.numel,
                                  self.lr if lr is None else float(lr), self.mu, self.wd, 1.0,
                                  1 if self.steps == 0 else 0, stream_ptr()), 'rgnn_sgd_step')
        self.steps += 1
        return loss, acc
