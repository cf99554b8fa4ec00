"""GNN building blocks with the reference's class names, constructor arguments and state_dict layout
(reference modules/neural_net/gnn/gnn_blocks.py), executing on librgnn.so tile programs.

Every class here is a parameter container plus a thin forward that launches fused CUDA programs:
a whole nn.Sequential of ffn_blocks is one kernel; the message-passing block is two (edge tile program
with gather + edge MLP + segmented sum, node tile program with update MLP + residual).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
from torch import nn

from . import _cabi
from ._cabi import check, lib, ptr, stream_ptr
from ._engine import GraphBatch, ParamTable, apply_stack, flat_grads, stack_refs, _check_params, _f32c, _require_cuda
from .common import ffn_block
from .constants import (_CLS_BIAS_INIT_, _HEAD_WEIGHT_MEAN_INIT_, _HEAD_WEIGHT_STD_INIT_, _REG_BIAS_INIT_)


def _ffn_sequence(in_channels: int, widths: List[int], activation: str, norm_layer, num_groups,
                  first_without_norm: bool = False) -> nn.Sequential:
    blocks, c = [], in_channels
    for i, w in enumerate(widths):
        if first_without_norm and i == 0:
            blocks.append(ffn_block(in_channels=c, out_channels=w, activation=activation))
        else:
            blocks.append(ffn_block(in_channels=c, out_channels=w, activation=activation,
                                    norm_layer=norm_layer, num_groups=num_groups))
        c = w
    return nn.Sequential(*blocks)


class graph_feature_encoding(nn.Module):
    """Raw feature -> embedding MLP; the first block has no norm (reference gnn_blocks.py:19-42)."""

    def __init__(self, in_channels: int, stem_channels: List[int], activation: str, norm_layer: str, num_groups: int):
        super().__init__()
        self.encoder = _ffn_sequence(in_channels, stem_channels, activation, norm_layer, num_groups,
                                     first_without_norm=True)

    def forward(self, x: torch.Tensor):
        return apply_stack(x, list(self.encoder))


def _build_conv(blk, device, grads=None):
    """rgnn_conv struct of one block: parameter pointers, freshly packed weight images, optional gradient views."""
    tab = ParamTable()
    msg, upd = stack_refs(blk.msg), stack_refs(blk.upd)
    cn = upd[-1].out_features
    dims = (cn, msg[0].in_features - 2 * cn, msg[0].out_features)
    conv = _cabi.rgnn_conv()
    conv.msg.n, conv.upd.n = len(msg), len(upd)
    slots_m = [tab.add(r, dims if i == 0 else None) for i, r in enumerate(msg)]
    slots_u = [tab.add(r) for r in upd]
    _check_params(tab.tensors)
    packed = torch.empty(tab.packed_floats(), dtype=torch.float32, device=device)
    off = 0
    for i, s in enumerate(slots_m):
        off += tab.fill(s, conv.msg.layer[i], packed, off, grads)
    for i, s in enumerate(slots_u):
        off += tab.fill(s, conv.upd.layer[i], packed, off, grads)
    check(lib().rgnn_pack_conv(C.byref(conv), stream_ptr()), 'rgnn_pack_conv')
    return tab, conv, packed, dims


class _ConvBlockFn(torch.autograd.Function):
    """Stand-alone residual_graph_conv_block: forward on the fused kernels, backward through rgnn_conv_block_bwd (generic tile
    programs; a training run of the whole detector goes through the detector-level Function and its fused backward kernels).
    `params` are the block's parameters in ParamTable order (msg stack, then upd stack: weight, bias, mu, std per layer)."""

    @staticmethod
    def forward(ctx, blk, gb: GraphBatch, x, e_tm, *params):
        tab, conv, packed, dims = _build_conv(blk, x.device)
        st = stream_ptr()
        out = torch.empty_like(x)
        agg = torch.empty_like(x)
        proj = torch.empty((x.shape[0], 2 * dims[2]), dtype=torch.float32, device=x.device)
        g = gb.c_struct()
        check(lib().rgnn_conv_block_fwd(C.byref(conv), C.byref(g), ptr(x), ptr(e_tm), ptr(out), ptr(agg), ptr(proj), st),
              'rgnn_conv_block_fwd')
        ctx.blk, ctx.gb = blk, gb
        ctx.save_for_backward(x, e_tm, agg, proj, *params)
        return out

    @staticmethod
    def backward(ctx, g_out):
        x, e_tm, agg, proj = ctx.saved_tensors[:4]
        params = ctx.saved_tensors[4:]
        needs = list(ctx.needs_input_grad[4:])
        flat, views = flat_grads(params, needs)
        tab, conv, packed, dims = _build_conv(ctx.blk, x.device, views)
        if [t.data_ptr() for t in tab.tensors] != [p.data_ptr() for p in params]:
            raise _cabi.RgnnError('residual_graph_conv_block: parameters moved between forward and backward')
        g = ctx.gb.c_struct()
        dx = torch.empty_like(x)
        de = torch.empty_like(e_tm)
        nbytes = lib().rgnn_conv_block_bwd_workspace_bytes(C.byref(conv), C.byref(g))
        ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=x.device)
        check(lib().rgnn_conv_block_bwd(C.byref(conv), C.byref(g), ptr(x), ptr(e_tm), ptr(agg), ptr(proj), ptr(_f32c(g_out)),
                                        ptr(dx), ptr(de), ptr(ws), nbytes, stream_ptr()), 'rgnn_conv_block_bwd')
        return (None, None, dx, de) + tuple(views)


def _conv_params(blk):
    """The block's parameters in ParamTable order (what _build_conv registers), so that the gradient views line up."""
    tab = ParamTable()
    msg, upd = stack_refs(blk.msg), stack_refs(blk.upd)
    cn = upd[-1].out_features
    dims = (cn, msg[0].in_features - 2 * cn, msg[0].out_features)
    for i, r in enumerate(msg):
        tab.add(r, dims if i == 0 else None)
    for r in upd:
        tab.add(r)
    return list(tab.tensors)


class residual_graph_conv_block(nn.Module):
    """x' = x + upd(cat(x, sum_{s->t} msg(cat(x_t, x_s, e_st))))   (reference gnn_blocks.py:45-113; PyG
    MessagePassing with aggr='add', flow='source_to_target').  Only the identity residual (in == out) and
    'add' aggregation of the reference configuration are implemented."""

    def __init__(self, in_node_channels: int, in_edge_channels: int, mlp_stem_channels_msg: List[int],
                 mlp_stem_channels_upd: List[int], aggregation: str, activation: str, norm_layer: str,
                 num_groups: int, in_extra_feature_dim: Optional[int] = None):
        super().__init__()
        if aggregation not in ('add', 'sum'):
            raise NotImplementedError(f"aggregation '{aggregation}': only 'add' (the reference yml) is implemented")
        if in_extra_feature_dim is not None:
            raise NotImplementedError('extra (augmented) node features are not used by the reference model')
        self.aggr = aggregation
        self.in_extra_feature_dim = None
        self.msg = _ffn_sequence(2 * in_node_channels + in_edge_channels, mlp_stem_channels_msg, activation,
                                 norm_layer, num_groups)
        self.upd = _ffn_sequence(in_node_channels + mlp_stem_channels_msg[-1], mlp_stem_channels_upd, activation,
                                 norm_layer, num_groups)
        self.match_channels = in_node_channels != mlp_stem_channels_upd[-1]
        self.residual_connection = None
        if self.match_channels:
            raise NotImplementedError('channel-matching residual (in != out) is not implemented: the reference '
                                      'configuration keeps 64 channels through all conv blocks')

    def forward(self, node_features: torch.Tensor, edge_features: torch.Tensor, edge_index: torch.Tensor,
                extra_features: Optional[torch.Tensor] = None):
        _require_cuda(node_features, edge_features, edge_index)
        gb = GraphBatch.from_edge_index(edge_index, node_features.shape[0])
        e_tm = _f32c(edge_features).index_select(0, gb.perm[:gb.n_edges].long()) if gb.n_edges else _f32c(edge_features)
        params = _conv_params(self)
        return _ConvBlockFn.apply(self, gb, _f32c(node_features), e_tm, *params)


class graph_convolution(nn.Module):
    """Stack of residual_graph_conv_blocks sharing one edge embedding (reference gnn_blocks.py:116-164)."""

    def __init__(self, in_node_channels: int, in_edge_channels: int, stem_channels: List[int], msg_mlp_hidden_dim: int,
                 activation: str, aggregation: str, norm_layer: str, num_groups: int,
                 append_extra_features: Optional[List[bool]] = None, in_extra_feature_dim: Optional[int] = None):
        super().__init__()
        if append_extra_features is not None and any(append_extra_features) and in_extra_feature_dim is not None:
            raise NotImplementedError('extra (augmented) node features are not used by the reference model')
        self.conv_blk = nn.ModuleList()
        c = in_node_channels
        for w in stem_channels:
            self.conv_blk.append(residual_graph_conv_block(
                in_node_channels=c, in_edge_channels=in_edge_channels, mlp_stem_channels_msg=[msg_mlp_hidden_dim, w],
                mlp_stem_channels_upd=[w], aggregation=aggregation, activation=activation, norm_layer=norm_layer,
                num_groups=num_groups))
            c = w

    def forward(self, node_features, edge_features, edge_index, extra_features=None):
        x = node_features
        for blk in self.conv_blk:
            x = blk(x, edge_features, edge_index, extra_features)
        return x


class FFN_TaskSpecificHead(nn.Module):
    """ffn_block(C->C) then a bare Linear(C->out), N(mu,sigma) / constant-bias init (reference :167-197)."""

    def __init__(self, in_channels: int, out_channels: int, activation: str, norm_layer: str, num_groups: int,
                 init_weight_mu: float, init_weight_sigma: float, init_bias: float):
        super().__init__()
        last = nn.Linear(in_features=in_channels, out_features=out_channels, bias=True)
        nn.init.normal_(last.weight, mean=init_weight_mu, std=init_weight_sigma)
        nn.init.constant_(last.bias, init_bias)
        self.head = nn.Sequential(
            ffn_block(in_channels=in_channels, out_channels=in_channels, activation=activation,
                      norm_layer=norm_layer, num_groups=num_groups), last)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return apply_stack(x, list(self.head))


def _cls_head(c, n_out, activation, norm_layer, num_groups):
    return FFN_TaskSpecificHead(c, n_out, activation, norm_layer, num_groups, _HEAD_WEIGHT_MEAN_INIT_,
                                _HEAD_WEIGHT_STD_INIT_, _CLS_BIAS_INIT_)


class node_segmentation(nn.Module):
    """Per-node class logits (reference :200-234).  stem + head run as one kernel."""

    def __init__(self, in_channels: int, stem_channels: List[int], num_classes: int, activation: str,
                 norm_layer: str, num_groups: int):
        super().__init__()
        self.stem = _ffn_sequence(in_channels, stem_channels, activation, norm_layer, num_groups)
        self.pred_cls = _cls_head(stem_channels[-1], num_classes, activation, norm_layer, num_groups)

    def forward(self, x: torch.Tensor):
        return apply_stack(x, list(self.stem) + list(self.pred_cls.head))


class node_offset_predictions(nn.Module):
    """Per-node (dx,dy) offsets to the object centre (reference :237-271)."""

    def __init__(self, in_channels: int, stem_channels: List[int], reg_offset_dim: int, activation: str,
                 norm_layer: str, num_groups: int):
        super().__init__()
        self.stem = _ffn_sequence(in_channels, stem_channels, activation, norm_layer, num_groups)
        self.pred_offsets = FFN_TaskSpecificHead(stem_channels[-1], reg_offset_dim, activation, norm_layer, num_groups,
                                                 _HEAD_WEIGHT_MEAN_INIT_, _HEAD_WEIGHT_STD_INIT_, _REG_BIAS_INIT_)

    def forward(self, x: torch.Tensor):
        return apply_stack(x, list(self.stem) + list(self.pred_offsets.head))


def _undirected_pairs(adj_matrix: torch.Tensor):
    """nonzero(triu(adj,1)) in row-major order (reference :295-296) -- API-compatibility path on a dense
    matrix; the detector-level forward derives the same list from edge_index instead.  A (2, E) int64 `edge_index`
    (reference convention, source-sorted) is accepted in place of the dense matrix: its columns with source < target ARE
    that list, in the same order."""
    if adj_matrix.dtype == torch.int64 and adj_matrix.dim() == 2 and adj_matrix.shape[0] == 2:
        und = adj_matrix[0] < adj_matrix[1]
        return adj_matrix[0][und], adj_matrix[1][und]
    r, c = torch.nonzero(torch.triu(adj_matrix, diagonal=1), as_tuple=True)
    return r, c


class edge_formation(nn.Module):
    """h = stem(x); one feature row per undirected link: h[r] + h[c] (reference :274-298)."""

    def __init__(self, in_channels: int, num_blocks: int, activation: str, norm_layer: str, num_groups: int):
        super().__init__()
        self.stem = _ffn_sequence(in_channels, [in_channels] * num_blocks, activation, norm_layer, num_groups)

    def forward(self, x: torch.Tensor, adj_matrix: torch.Tensor):
        h = apply_stack(x, list(self.stem))
        r, c = _undirected_pairs(adj_matrix)
        return h.index_select(0, r) + h.index_select(0, c)


class link_predictions(nn.Module):
    """Link logits per undirected edge (reference :301-344)."""

    def __init__(self, in_channels: int, num_blks_for_edges: int, stem_channels: List[int], num_classes: int,
                 activation: str, norm_layer: str, num_groups: int):
        super().__init__()
        self.compute_edge = edge_formation(in_channels, num_blks_for_edges, activation, norm_layer, num_groups)
        self.stem = _ffn_sequence(in_channels, stem_channels, activation, norm_layer, num_groups)
        self.pred_cls = _cls_head(stem_channels[-1], num_classes, activation, norm_layer, num_groups)

    def forward(self, x: torch.Tensor, adj_matrix: torch.Tensor):
        u = self.compute_edge(x, adj_matrix)
        return apply_stack(u, list(self.stem) + list(self.pred_cls.head))


class _SegMaxFn(torch.autograd.Function):
    """pooled[c] = max over the member rows of cluster c (rgnn_segment_max_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, g, cl_ptr, cl_members, n_clusters):
        _require_cuda(g)
        g = _f32c(g)
        pooled = torch.empty((n_clusters, g.shape[1]), dtype=torch.float32, device=g.device)
        arg = torch.empty((max(n_clusters, 1), g.shape[1]), dtype=torch.int32, device=g.device)
        check(lib().rgnn_segment_max_fwd(ptr(g), g.shape[1], ptr(cl_ptr), ptr(cl_members), n_clusters, ptr(pooled), ptr(arg),
                                         stream_ptr()), 'rgnn_segment_max_fwd')
        ctx.save_for_backward(arg)
        ctx.shape = (g.shape[0], g.shape[1], n_clusters)
        return pooled

    @staticmethod
    def backward(ctx, d_pooled):
        (arg,) = ctx.saved_tensors
        n, w, c = ctx.shape
        dx = torch.empty((n, w), dtype=torch.float32, device=d_pooled.device)
        check(lib().rgnn_segment_max_bwd(ptr(_f32c(d_pooled)), ptr(arg), c, w, n, ptr(dx), stream_ptr()), 'rgnn_segment_max_bwd')
        return dx, None, None, None


class object_classification(nn.Module):
    """Per-cluster class logits: stem on nodes, max-pool over cluster members, head (reference :347-389)."""

    def __init__(self, in_channels: int, stem_channels: List[int], num_classes: int, activation: str,
                 norm_layer: str, num_groups: int):
        super().__init__()
        self.stem = _ffn_sequence(in_channels, stem_channels, activation, norm_layer, num_groups)
        self.pred_cls = _cls_head(stem_channels[-1], num_classes, activation, norm_layer, num_groups)

    def forward(self, x: torch.Tensor, cluster_node_idx: List[torch.Tensor]):
        g = apply_stack(x, list(self.stem))
        # stand-alone API path (the detector-level forward fuses the segment-max into the head kernel)
        gb = GraphBatch()
        gb.set_clusters([list(cluster_node_idx)], [0], g.device)
        pooled = _SegMaxFn.apply(g, gb.cl_ptr, gb.cl_members, gb.n_clusters)
        return apply_stack(pooled, list(self.pred_cls.head))


class node_predictions(nn.Module):
    """Merged segmentation + offset head of Model_Inference_v1 (reference :392-439)."""

    def __init__(self, in_channels: int, stem_channels: List[int], num_classes: int, reg_offset_dim: int,
                 activation: str, norm_layer: str, num_groups: int):
        super().__init__()
        self.stem = _ffn_sequence(in_channels, stem_channels, activation, norm_layer, num_groups)
        self.pred_cls = _cls_head(stem_channels[-1], num_classes, activation, norm_layer, num_groups)
        self.pred_offsets = FFN_TaskSpecificHead(stem_channels[-1], reg_offset_dim, activation, norm_layer, num_groups,
                                                 _HEAD_WEIGHT_MEAN_INIT_, _HEAD_WEIGHT_STD_INIT_, _REG_BIAS_INIT_)

    def forward(self, x: torch.Tensor):
        s = apply_stack(x, list(self.stem))
        return apply_stack(s, list(self.pred_cls.head)), apply_stack(s, list(self.pred_offsets.head))
