"""ORACLE (test infrastructure, never the product path): plain-PyTorch fp32 CPU restatement of the
reference's radar GNN detector forward, losses and (through torch autograd) backward.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may import
this file.  It is written functionally over a flat `state_dict` with the reference's key layout
(SURVEY.md appendix B), so it can run the checked-in checkpoint without any nn.Module of the reference.

Parity status: PINNED against the reference itself -- tests/golden/make_golden.py imports the
reference's own gnn_detector.py / gnn_blocks.py / loss.py from /root/reference (with a 20-line stand-in
for the absent third-party `torch_geometric.nn.conv.MessagePassing`, restated from PyG >= 2.5 semantics:
x_j = x[edge_index[0]], x_i = x[edge_index[1]], out = zeros.index_add_(0, edge_index[1], message)) and
stores its outputs, losses and gradients in tests/golden/model_*.npz; tests/test_oracle_golden.py holds
this restatement to them.

Reference lines restated:
  ffn_block / channel_normalization / LeakyReLU   modules/neural_net/common.py:185-220,256-267
  graph_feature_encoding                          modules/neural_net/gnn/gnn_blocks.py:19-42
  residual_graph_conv_block (+PyG propagate)      gnn_blocks.py:45-113
  graph_convolution                               gnn_blocks.py:116-164
  FFN_TaskSpecificHead and the four heads         gnn_blocks.py:167-389
  Model_Inference.forward (cluster list given)    modules/neural_net/gnn/gnn_detector.py:141-162
  Model_Training.forward                          gnn_detector.py:428-478
  Loss_Graph                                      modules/neural_net/gnn/loss.py:37-76
  CE / MSE / sigmoid focal                        modules/neural_net/lossfunc.py:19-55
  normalize_gt_offsets                            modules/compute_groundtruth/compute_offsets.py:6-11
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

EPS = 1e-5            # modules/neural_net/constants.py:9
LEAKY_SLOPE = 0.01    # constants.py:10

# Test knob (tests/gpu_util.GradientYardstick; 0 everywhere else): every Linear output is multiplied by (1 - LINEAR_OUTPUT_SHRINK) -- what
# accumulating a dot product with round-toward-zero instead of round-to-nearest does on average (the tensor cores truncate
# their fp32 accumulator: measured signed bias of -1.5e-7 .. -5.4e-7 per GEMM on B200, tools/mma_noise.py).
LINEAR_OUTPUT_SHRINK = 0.0


class ActivationProbe:
    """Test instrumentation (tests/gpu_util.GradientYardstick).  The loss gradient of a LeakyReLU(0.01) stack is
    DISCONTINUOUS where a pre-activation crosses zero: two evaluations that agree to float32 rounding can sit on different
    sides of a kink and their gradients differ by a finite jump.  The probe (a) lists the activations whose pre-activation
    lies within `window` of zero, and (b) re-evaluates with chosen activations taken on the OTHER branch, which gives the
    exact size of each jump.  Activations are addressed as (ffn call index, row, channel); the call order of the functions
    below is deterministic."""

    def __init__(self, window=None, flips=None):
        self.window, self.flips = window, flips or {}
        self.calls, self.found = 0, []


_PROBE = None


def _leaky(z: torch.Tensor) -> torch.Tensor:
    y = F.leaky_relu(z, LEAKY_SLOPE)
    p = _PROBE
    if p is not None:
        idx = p.calls
        p.calls += 1
        if p.window is not None:
            for r, c in (z.detach().abs() < p.window).nonzero().tolist():
                p.found.append((idx, r, c))
        if idx in p.flips:
            mask = torch.zeros(z.shape, dtype=torch.bool)
            for r, c in p.flips[idx]:
                mask[r, c] = True
            y = torch.where(mask, torch.where(z > 0, LEAKY_SLOPE * z, z), y)
    return y


def channel_norm(z: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor) -> torch.Tensor:
    """common.py:215-220: per-row mean, UNBIASED std, eps added to the std, scalar affine."""
    mu = z.mean(dim=1, keepdim=True)
    sd = z.std(dim=1, keepdim=True)
    return scale * ((z - mu) / (sd + EPS)) + shift


def ffn(sd: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor) -> torch.Tensor:
    """One ffn_block whose parameters live under `prefix` ('....block'): Linear -> [norm] -> LeakyReLU."""
    z = F.linear(x, sd[prefix + '.0.weight'], sd[prefix + '.0.bias'])
    if LINEAR_OUTPUT_SHRINK:
        z = z * (1.0 - LINEAR_OUTPUT_SHRINK)
    if (prefix + '.1.mu') in sd:
        z = channel_norm(z, sd[prefix + '.1.std'], sd[prefix + '.1.mu'])
    return _leaky(z)


def _count(sd, stem: str) -> int:
    n = 0
    while f'{stem}.{n}.block.0.weight' in sd:
        n += 1
    return n


def ffn_stack(sd, stem: str, x: torch.Tensor) -> torch.Tensor:
    for i in range(_count(sd, stem)):
        x = ffn(sd, f'{stem}.{i}.block', x)
    return x


def task_head(sd, stem: str, x: torch.Tensor) -> torch.Tensor:
    """FFN_TaskSpecificHead (gnn_blocks.py:167-197): ffn_block then a bare Linear."""
    x = ffn(sd, f'{stem}.head.0.block', x)
    return F.linear(x, sd[f'{stem}.head.1.weight'], sd[f'{stem}.head.1.bias'])


def conv_block(sd, stem: str, x: torch.Tensor, e: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
    """residual_graph_conv_block with 64->64 identity residual (gnn_blocks.py:96-113)."""
    src, dst = edge_index[0], edge_index[1]
    m = torch.cat((x.index_select(0, dst), x.index_select(0, src), e), dim=-1)
    m = ffn_stack(sd, f'{stem}.msg', m)
    agg = torch.zeros((x.shape[0], m.shape[1]), dtype=m.dtype).index_add_(0, dst, m)
    return x + ffn_stack(sd, f'{stem}.upd', torch.cat((x, agg), dim=-1))


def detector_forward(sd, node_feat, edge_feat, edge_index, clusters: Sequence[torch.Tensor], p: str = 'pred.'):
    """Model_Inference.forward with cluster_node_idx given (gnn_detector.py:151-162).
    Undirected links are edge_index[:, src<dst], equal to nonzero(triu(adj,1)) in row-major order
    (gnn_blocks.py:295-296)."""
    x = ffn_stack(sd, p + 'encode_node_feat.encoder', node_feat)
    e = ffn_stack(sd, p + 'encode_edge_feat.encoder', edge_feat)
    layer = 0
    while f'{p}pass_messages.conv_blk.{layer}.msg.0.block.0.weight' in sd:
        x = conv_block(sd, f'{p}pass_messages.conv_blk.{layer}', x, e, edge_index)
        layer += 1
    node_cls = task_head(sd, p + 'predict_node.pred_cls', ffn_stack(sd, p + 'predict_node.stem', x))
    node_off = task_head(sd, p + 'predict_offset.pred_offsets', ffn_stack(sd, p + 'predict_offset.stem', x))
    h = ffn_stack(sd, p + 'predict_link.compute_edge.stem', x)
    und = edge_index[0] < edge_index[1]
    r, c = edge_index[0][und], edge_index[1][und]
    link = task_head(sd, p + 'predict_link.pred_cls', ffn_stack(sd, p + 'predict_link.stem', h[r] + h[c]))
    g = ffn_stack(sd, p + 'predict_class.stem', x)
    pooled = torch.cat([g[idx].max(dim=0, keepdim=True)[0] for idx in clusters], dim=0)
    obj = task_head(sd, p + 'predict_class.pred_cls', pooled)
    return node_cls, node_off, link, obj


def sigmoid_focal(logits, targets, alpha=0.25, gamma=2.0):
    """torchvision.ops.sigmoid_focal_loss restated (lossfunc.py:55), reduction 'none'."""
    prob = torch.sigmoid(logits)
    ce = F.binary_cross_entropy_with_logits(logits, targets, reduction='none')
    p_t = prob * targets + (1 - prob) * (1 - targets)
    loss = ce * ((1 - p_t) ** gamma)
    a_t = alpha * targets + (1 - alpha) * (1 - targets)
    return a_t * loss


def graph_losses(node_cls, node_off, link, obj, node_cls_gt, node_off_gt_normalised, link_gt, obj_gt,
                 class_weights=(1., 1., 1., 1., 1., 1., .5),
                 w_node_cls=1.0, w_node_reg=5.0, w_edge_cls=2.0, w_obj_cls=1.0):
    """Loss_Graph.forward (loss.py:37-76); each term is sum/count over the whole (global) batch."""
    cw = torch.tensor(class_weights, dtype=torch.float32)
    edge = sigmoid_focal(link, F.one_hot(link_gt, link.shape[1]).float()).sum(-1)
    edge = edge.sum() / edge.shape[0]
    node = -(F.log_softmax(node_cls, dim=-1) * F.one_hot(node_cls_gt, node_cls.shape[1]).float() * cw).sum(-1)
    node = node.sum() / node.shape[0]
    reg = 0.5 * ((node_off - node_off_gt_normalised) ** 2).sum(-1)
    reg = reg.sum() / reg.shape[0]
    ob = -(F.log_softmax(obj, dim=-1) * F.one_hot(obj_gt, obj.shape[1]).float()).sum(-1)
    ob = ob.sum() / ob.shape[0]
    return {'loss_node_cls': node * w_node_cls, 'loss_node_reg': reg * w_node_reg,
            'loss_edge_cls': edge * w_edge_cls, 'loss_obj_cls': ob * w_obj_cls}


def accuracy(logits, gt):
    """compute_accuracy (gnn_detector.py:23-28)."""
    return (logits.argmax(dim=-1) == gt).sum() / gt.shape[0]


def training_forward(sd, node_features: List[torch.Tensor], edge_features: List[torch.Tensor],
                     edge_index: List[torch.Tensor], labels: Dict[str, list],
                     offset_mu=(0., 0.), offset_sigma=(8., 4.)):
    """Model_Training.forward (gnn_detector.py:428-478): per-frame loop, concat, losses, accuracies.
    Unlike the reference this does not normalise labels['node_offsets'] in place (it works on a copy)."""
    outs = [detector_forward(sd, nf, ef, ei, cl)
            for nf, ef, ei, cl in zip(node_features, edge_features, edge_index, labels['cluster_node_idx'])]
    node_cls, node_off, link, obj = (torch.cat([o[i] for o in outs], dim=0) for i in range(4))
    off_gt = torch.cat(labels['node_offsets'], dim=0).clone()
    off_gt[:, 0] = (off_gt[:, 0] - offset_mu[0]) / offset_sigma[0]
    off_gt[:, 1] = (off_gt[:, 1] - offset_mu[1]) / offset_sigma[1]
    node_gt = torch.cat(labels['node_class'], dim=0)
    link_gt = torch.cat(labels['edge_class'], dim=0)
    obj_gt = torch.cat(labels['cluster_labels'], dim=0)
    loss = graph_losses(node_cls, node_off, link, obj, node_gt, off_gt, link_gt, obj_gt)
    acc = {'segment_accuracy': accuracy(node_cls, node_gt), 'edge_accuracy': accuracy(link, link_gt),
           'object_accuracy': accuracy(obj, obj_gt)}
    return loss, acc, (node_cls, node_off, link, obj)
