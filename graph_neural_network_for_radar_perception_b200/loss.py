"""Multi-task loss with the reference's interface (reference modules/neural_net/gnn/loss.py:10-76,
modules/neural_net/lossfunc.py:19-55): sigmoid focal loss on links, class-weighted cross entropy on nodes,
0.5*MSE on offsets, cross entropy on objects -- each `sum / count`, times its loss weight.

One CUDA kernel (rgnn_losses_fwdbwd) produces the four scalars, the gradients w.r.t. all logits and the
three arg-max hit counts in a single pass over the predictions.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch
from torch import nn

from . import _cabi
from ._cabi import check, lib, ptr, stream_ptr
from ._engine import _f32c, _require_cuda


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, counts, dp_tail, node_cls, node_off, link_cls, obj_cls, node_gt, off_gt, link_gt, obj_gt):
        _require_cuda(node_cls, node_off, link_cls, obj_cls, node_gt, off_gt, link_gt, obj_gt)
        node_cls, node_off, link_cls, obj_cls, off_gt = map(_f32c, (node_cls, node_off, link_cls, obj_cls, off_gt))
        node_gt, link_gt, obj_gt = (t.to(torch.int64).contiguous() for t in (node_gt, link_gt, obj_gt))
        dev = node_cls.device
        n, eu, nc = node_cls.shape[0], link_cls.shape[0], obj_cls.shape[0]
        counts_dev = None
        if isinstance(counts, torch.Tensor):          # device-resident global counts (data parallel): never read by the host
            if not counts.is_cuda or counts.dtype != torch.float64 or counts.numel() != 3:
                raise _cabi.RgnnError('global_counts tensor must be 3 float64 values on the CUDA device')
            counts_dev, counts = counts.contiguous(), None
        if dp_tail is not None and (not dp_tail.is_cuda or dp_tail.dtype != torch.float32 or dp_tail.numel() < 5
                                    or not dp_tail.is_contiguous()):
            raise _cabi.RgnnError('dp_tail must be >= 5 contiguous float32 values on the CUDA device')
        cn, cu, cc = counts if counts is not None else (n, eu, nc)
        grads = [torch.empty_like(t) for t in (node_cls, node_off, link_cls, obj_cls)]
        losses = torch.empty(4, dtype=torch.float64, device=dev)
        correct = torch.empty(3, dtype=torch.int32, device=dev)
        check(lib().rgnn_losses_fwdbwd(C.byref(cfg), ptr(node_cls), ptr(node_off), ptr(link_cls), ptr(obj_cls),
                                       ptr(node_gt), ptr(off_gt), ptr(link_gt), ptr(obj_gt), n, eu, nc,
                                       float(cn), float(cu), float(cc), ptr(grads[0]), ptr(grads[1]), ptr(grads[2]),
                                       ptr(grads[3]), ptr(losses), ptr(correct), ptr(counts_dev), ptr(dp_tail),
                                       stream_ptr()), 'rgnn_losses_fwdbwd')
        ctx.save_for_backward(*grads)
        ctx.mark_non_differentiable(correct)
        out = losses.to(torch.float32)
        return out[0], out[1], out[2], out[3], correct

    @staticmethod
    def backward(ctx, g0, g1, g2, g3, _gc):
        d_node_cls, d_node_off, d_link, d_obj = ctx.saved_tensors
        return (None, None, None, d_node_cls * g0, d_node_off * g1, d_link * g2, d_obj * g3, None, None, None, None)


class Loss_Graph(nn.Module):
    def __init__(self, net_config, device: str):
        super().__init__()
        self.node_cls_loss_weight = net_config.node_cls_loss_weight
        self.edge_cls_loss_weight = net_config.edge_cls_loss_weight
        self.node_reg_loss_weight = net_config.node_reg_loss_weight
        self.obj_cls_loss_weight = net_config.obj_cls_loss_weight
        self.new_labels_to_id_dict = net_config.new_labels_to_id_dict_dyn
        self.num_classes_edge = net_config.num_edge_classes
        self.num_classes = net_config.num_classes
        self.class_weights = torch.tensor(net_config.class_weights_dyn, dtype=torch.float32, device=device)
        self.device = device
        cfg = _cabi.rgnn_loss_cfg()
        for i, w in enumerate(net_config.class_weights_dyn):
            cfg.class_weights[i] = float(w)
        cfg.n_classes, cfg.n_edge_classes = self.num_classes, self.num_classes_edge
        cfg.w_node_cls, cfg.w_node_reg = float(self.node_cls_loss_weight), float(self.node_reg_loss_weight)
        cfg.w_edge_cls, cfg.w_obj_cls = float(self.edge_cls_loss_weight), float(self.obj_cls_loss_weight)
        cfg.focal_alpha, cfg.focal_gamma = 0.25, 2.0          # lossfunc.py:48 defaults
        self._cfg = cfg
        self.last_correct = None

    def compute_valid_object_mask(self, gt_class_logits):
        return gt_class_logits != self.new_labels_to_id_dict['FALSE']

    def forward(self, pred, gt, global_counts=None, dp_tail: Optional[torch.Tensor] = None):
        """pred / gt: 4-tuples (node_class_logits, node_reg_deltas, edge_class_logits, obj_class_logits); gt classes
        are index tensors, gt offsets already normalised.  global_counts = (N, E_u, C) over all data-parallel
        ranks -- host numbers or a 3-element float64 device tensor -- makes each rank's share sum to the reference's
        global-batch loss (loss.py:58,62,66,70); None = this batch's own counts, as the reference divides.
        dp_tail: 5 float32 device values receiving the NaN flag and this rank's four loss shares (include/rgnn.h)."""
        l_node, l_reg, l_edge, l_obj, correct = _LossFn.apply(
            self._cfg, global_counts, dp_tail, pred[0], pred[1], pred[2], pred[3], gt[0], gt[1], gt[2], gt[3])
        self.last_correct = correct
        return {'loss_node_cls': l_node, 'loss_node_reg': l_reg, 'loss_edge_cls': l_edge, 'loss_obj_cls': l_obj}


class Loss_Object_Class(nn.Module):
    """Object-class cross entropy only (reference loss.py:79-89), used by the fine-tuning model."""

    def __init__(self, net_config):
        super().__init__()
        self.num_classes = net_config.num_classes

    def forward(self, pred_obj_class_logits, gt_obj_class_logits):
        # tiny (C x 7); plain torch op on the device
        return torch.nn.functional.cross_entropy(pred_obj_class_logits, gt_obj_class_logits, reduction='sum') \
            / pred_obj_class_logits.shape[0]
