"""Host-side glue between the PyTorch-facing modules and librgnn.so: parameter tables, packed-weight
buffers, the batched graph container and the autograd Functions.  PyTorch is used for device memory,
streams and autograd bookkeeping only; all arithmetic happens in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _cabi
from ._cabi import check, lib, ptr, stream_ptr


def _require_cuda(*tensors):
    """No CPU path, and the C library launches on the CURRENT device with the current stream of that device: tensors on another
    GPU would be read by kernels of the wrong device, so that fails here (one process per GPU, or torch.cuda.device(i))."""
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise _cabi.RgnnError('this implementation has no CPU path: tensors must live on a CUDA device '
                                  f'(got {t.device})')
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise _cabi.RgnnError(f'tensor on {t.device} but the current CUDA device is cuda:{cur}: call inside '
                                  f'`with torch.cuda.device({t.device.index}):` (kernels run on the current device and stream)')


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    return t.contiguous()


# -------------------------------------------------------------------------------------------------
# parameter tables
# -------------------------------------------------------------------------------------------------
class LinearRef:
    """The tensors of one reference ffn_block (or bare nn.Linear) in the order they are handed to autograd."""

    def __init__(self, weight, bias, scale=None, shift=None, activation=True):
        self.weight, self.bias, self.scale, self.shift, self.activation = weight, bias, scale, shift, activation

    @property
    def out_features(self):
        return self.weight.shape[0]

    @property
    def in_features(self):
        return self.weight.shape[1]

    def tensors(self):
        return [t for t in (self.weight, self.bias, self.scale, self.shift) if t is not None]


def linear_ref_of(block) -> LinearRef:
    """Accepts our ffn_block or an nn.Linear."""
    if isinstance(block, torch.nn.Linear):
        return LinearRef(block.weight, block.bias, None, None, False)
    lin = block.block[0]
    norm = block.block[1] if len(block.block) == 3 else None
    if norm is not None and not getattr(norm, 'is_channel_normalization', False):
        raise NotImplementedError(
            'only channel_normalization is implemented (layer/group normalisation use cross-row statistics and '
            'do not batch per frame; the reference configuration does not use them)')
    if getattr(block.block[-1], 'kind', 'leakyrelu') != 'leakyrelu':
        raise NotImplementedError('only the LeakyReLU activation of the reference configuration is implemented')
    return LinearRef(lin.weight, lin.bias, norm.std if norm is not None else None,
                     norm.mu if norm is not None else None, True)


def stack_refs(blocks: Sequence) -> List[LinearRef]:
    return [linear_ref_of(b) for b in blocks]


class ParamTable:
    """Flattens a list of LinearRef groups into: the tensor list given to autograd, packed-weight storage,
    optional flat gradient storage, and filled C structs."""

    def __init__(self):
        self.tensors: List[torch.Tensor] = []
        self._slots = []      # (LinearRef, (iw, ib, is, ih), packed_floats, is_conv_msg0 dims or None)

    def add(self, ref: LinearRef, conv_dims=None):
        idx = []
        for t in (ref.weight, ref.bias, ref.scale, ref.shift):
            if t is None:
                idx.append(-1)
            else:
                idx.append(len(self.tensors))
                self.tensors.append(t)
        if conv_dims is not None:
            n = lib().rgnn_packed_conv_msg0_floats(*conv_dims)
        else:
            n = lib().rgnn_packed_weight_floats(ref.in_features, ref.out_features)
        self._slots.append((ref, tuple(idx), int(n)))
        return len(self._slots) - 1

    def packed_floats(self) -> int:
        return sum((s[2] + 63) // 64 * 64 for s in self._slots)

    def fill(self, slot: int, dst: _cabi.rgnn_linear, packed: torch.Tensor, packed_off: int,
             grads: Optional[List[Optional[torch.Tensor]]]):
        ref, idx, n = self._slots[slot]
        dst.weight = ptr(ref.weight)
        dst.weight_t = packed.data_ptr() + 4 * packed_off
        dst.bias = ptr(ref.bias)
        dst.norm_scale = ptr(ref.scale)
        dst.norm_shift = ptr(ref.shift)
        g = [None] * 4
        if grads is not None:
            g = [grads[i] if i >= 0 else None for i in idx]
        dst.grad_weight, dst.grad_bias, dst.grad_norm_scale, dst.grad_norm_shift = (ptr(t) for t in g)
        dst.in_features = ref.in_features
        dst.out_features = ref.out_features
        dst.activation = 1 if ref.activation else 0
        return (n + 63) // 64 * 64


def _check_params(tensors):
    for t in tensors:
        if not t.is_cuda:
            raise _cabi.RgnnError('module parameters are on the CPU; move the module to a CUDA device '
                                  '(this implementation has no CPU path)')
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise _cabi.RgnnError('parameters must be contiguous float32')


def flat_grads(tensors, needs):
    """One zero-filled flat buffer with a view per parameter that needs a gradient."""
    total = sum(t.numel() for t, n in zip(tensors, needs) if n)
    flat = torch.zeros(max(total, 1), dtype=torch.float32, device=tensors[0].device)
    views, off = [], 0
    for t, n in zip(tensors, needs):
        if n:
            views.append(flat[off:off + t.numel()].view(t.shape))
            off += t.numel()
        else:
            views.append(None)
    return flat, views


# -------------------------------------------------------------------------------------------------
# ffn stack (graph_feature_encoding / stems / heads used stand-alone)
# -------------------------------------------------------------------------------------------------
def _build_stack(refs: List[LinearRef], grads=None):
    if len(refs) > _cabi.RGNN_MAX_STACK:
        raise _cabi.RgnnError(f'stack of {len(refs)} blocks exceeds RGNN_MAX_STACK')
    tab = ParamTable()
    slots = [tab.add(r) for r in refs]
    _check_params(tab.tensors)
    packed = torch.empty(tab.packed_floats(), dtype=torch.float32, device=tab.tensors[0].device)
    st = _cabi.rgnn_stack()
    st.n = len(refs)
    off = 0
    for i, s in enumerate(slots):
        off += tab.fill(s, st.layer[i], packed, off, grads)
    return tab, st, packed


class StackFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, refs, *params):
        _require_cuda(x)
        x = _f32c(x)
        tab, st, packed = _build_stack(refs)
        s = stream_ptr()
        check(lib().rgnn_pack_stack(C.byref(st), s), 'rgnn_pack_stack')
        y = torch.empty((x.shape[0], refs[-1].out_features), dtype=torch.float32, device=x.device)
        check(lib().rgnn_ffn_stack_fwd(C.byref(st), ptr(x), x.shape[0], ptr(y), s), 'rgnn_ffn_stack_fwd')
        ctx.refs = refs
        ctx.save_for_backward(x, *params)
        return y

    @staticmethod
    def backward(ctx, gy):
        x = ctx.saved_tensors[0]
        params = ctx.saved_tensors[1:]
        refs = ctx.refs
        needs = list(ctx.needs_input_grad[2:])
        flat, views = flat_grads(params, needs)
        tab, st, packed = _build_stack(refs, views)
        s = stream_ptr()
        check(lib().rgnn_pack_stack(C.byref(st), s), 'rgnn_pack_stack')
        gy = _f32c(gy)
        if ctx.needs_input_grad[0] and refs[0].in_features % 64 != 0:
            raise NotImplementedError('gradient w.r.t. the input of an ffn stack needs in_features % 64 == 0 '
                                      '(raw 6/7-wide feature inputs never require it in the reference)')
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        nbytes = lib().rgnn_ffn_stack_bwd_workspace_bytes(C.byref(st))
        ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=x.device)
        check(lib().rgnn_ffn_stack_bwd(C.byref(st), ptr(x), ptr(gy), x.shape[0], ptr(gx), ptr(ws), nbytes, s),
              'rgnn_ffn_stack_bwd')
        return (gx, None) + tuple(views)


def apply_stack(x: torch.Tensor, blocks: Sequence) -> torch.Tensor:
    refs = stack_refs(blocks)
    params = [t for r in refs for t in r.tensors()]
    return StackFn.apply(x, refs, *params)


# -------------------------------------------------------------------------------------------------
# batched graph container
# -------------------------------------------------------------------------------------------------
class GraphBatch:
    """Block-diagonal batch of per-frame graphs in the layout of `rgnn_graph` (include/rgnn.h)."""

    def __init__(self):
        self.n_nodes = self.n_edges = self.n_und = self.n_clusters = 0
        self.row_ptr = self.src = self.tgt = self.perm = None
        self.und_a = self.und_b = self.cl_ptr = self.cl_members = None
        self.frame_node_ptr = None   # host list, len F+1
        self.frame_und_ptr = None    # host list or None
        self.frame_cluster_ptr = None

    def c_struct(self) -> _cabi.rgnn_graph:
        g = _cabi.rgnn_graph()
        g.n_nodes, g.n_edges, g.n_und, g.n_clusters = self.n_nodes, self.n_edges, self.n_und, self.n_clusters
        g.row_ptr, g.src, g.tgt, g.perm = ptr(self.row_ptr), ptr(self.src), ptr(self.tgt), ptr(self.perm)
        g.und_a, g.und_b = ptr(self.und_a), ptr(self.und_b)
        g.cl_ptr, g.cl_members = ptr(self.cl_ptr), ptr(self.cl_members)
        return g

    def set_clusters(self, cluster_lists: Sequence[Sequence[torch.Tensor]], node_offsets: Sequence[int], device):
        """cluster_lists[f] = list of LongTensors with frame-local node ids (reference labels['cluster_node_idx']).
        One concatenation of all member tensors + one offset add (the offsets per member are expanded on the host from the
        list lengths, which are host values anyway) instead of one tiny launch per cluster."""
        import numpy as np
        lens, parts, offs, fptr = [], [], [], [0]
        for clusters, off in zip(cluster_lists, node_offsets):
            for c in clusters:
                lens.append(int(c.shape[0]))
                parts.append(c)
                offs.append(int(off))
            fptr.append(len(lens))
        self.n_clusters = len(lens)
        self.frame_cluster_ptr = fptr
        if self.n_clusters == 0:
            self.cl_ptr = torch.zeros(1, dtype=torch.int32, device=device)
            self.cl_members = torch.zeros(1, dtype=torch.int32, device=device)
            return
        lens_np = np.asarray(lens, dtype=np.int64)
        cl_ptr = np.zeros(self.n_clusters + 1, dtype=np.int32)
        np.cumsum(lens_np, out=cl_ptr[1:])
        self.cl_ptr = torch.from_numpy(cl_ptr).to(device, non_blocking=True)
        dev = torch.device(device)
        if any(p.device != dev for p in parts):
            parts = [p.to(dev) for p in parts]
        members = parts[0] if len(parts) == 1 else torch.cat(parts)
        if any(offs):
            off_dev = torch.from_numpy(np.repeat(np.asarray(offs, dtype=np.int64), lens_np)).to(dev, non_blocking=True)
            members = members + off_dev
        self.cl_members = members.to(torch.int32)

    def set_clusters_packed(self, cl_ptr: torch.Tensor, cl_members: torch.Tensor, frame_cluster_ptr=None):
        """Already packed clusters: cl_ptr (C+1,) / cl_members (global node ids), int32 on the device."""
        self.cl_ptr, self.cl_members = cl_ptr, cl_members
        self.n_clusters = int(cl_ptr.shape[0]) - 1
        self.frame_cluster_ptr = frame_cluster_ptr

    @staticmethod
    def from_edge_index(edge_index: torch.Tensor, n_nodes: int) -> 'GraphBatch':
        """General path: any (2,E) int64 edge list with global node ids (reference edge_index convention:
        row 0 = source, row 1 = target; gnn_blocks.py:57 flow='source_to_target').  The number of undirected links
        (src < dst) sizes the link head's output, so it is the ONE value this path reads back from the device;
        graph_features.build_graph_batch, whose adjacency is symmetric by construction, does not need it."""
        _require_cuda(edge_index)
        if edge_index.dtype != torch.int64:
            edge_index = edge_index.to(torch.int64)
        edge_index = edge_index.contiguous()
        dev = edge_index.device
        E = int(edge_index.shape[1])
        gb = GraphBatch()
        gb.n_nodes, gb.n_edges = int(n_nodes), E
        i32 = dict(dtype=torch.int32, device=dev)
        gb.row_ptr = torch.empty(n_nodes + 1, **i32)
        gb.src = torch.empty(max(E, 1), **i32)
        gb.tgt = torch.empty(max(E, 1), **i32)
        gb.perm = torch.empty(max(E, 1), **i32)
        gb.und_a = torch.empty(max(E, 1), **i32)
        gb.und_b = torch.empty(max(E, 1), **i32)
        n_und = torch.zeros(1, **i32)
        nbytes = lib().rgnn_csr_from_edge_index_workspace_bytes(n_nodes, E)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        check(lib().rgnn_csr_from_edge_index(ptr(edge_index[0]), ptr(edge_index[1]), n_nodes, E, ptr(gb.row_ptr),
                                             ptr(gb.src), ptr(gb.tgt), ptr(gb.perm), ptr(gb.und_a), ptr(gb.und_b),
                                             ptr(n_und), ptr(ws), nbytes, stream_ptr()), 'rgnn_csr_from_edge_index')
        gb.n_und = int(n_und.item())
        gb.frame_node_ptr = [0, n_nodes]
        gb.cl_ptr = torch.zeros(1, **i32)
        gb.cl_members = torch.zeros(1, **i32)
        return gb

    @staticmethod
    def from_frames(edge_index_list: Sequence[torch.Tensor], n_nodes_list: Sequence[int]) -> 'GraphBatch':
        offs = [0]
        for n in n_nodes_list:
            offs.append(offs[-1] + int(n))
        parts = [ei if o == 0 else ei + o for ei, o in zip(edge_index_list, offs)]
        ei = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
        gb = GraphBatch.from_edge_index(ei, offs[-1])
        gb.frame_node_ptr = offs
        return gb


# -------------------------------------------------------------------------------------------------
# detector
# -------------------------------------------------------------------------------------------------
class DetectorTable:
    """rgnn_detector struct + packed weights for a Model_Inference module."""

    def __init__(self, model, grads=None):
        tab = ParamTable()
        det = _cabi.rgnn_detector()
        plan = []   # (stack struct, [slot...])

        def add_stack(cstack, refs, conv_msg0=None):
            if len(refs) > _cabi.RGNN_MAX_STACK:
                raise _cabi.RgnnError('stack too deep')
            cstack.n = len(refs)
            slots = []
            for i, r in enumerate(refs):
                slots.append(tab.add(r, conv_msg0 if i == 0 else None))
            plan.append((cstack, slots))

        add_stack(det.node_enc, stack_refs(model.encode_node_feat.encoder))
        add_stack(det.edge_enc, stack_refs(model.encode_edge_feat.encoder))
        blks = model.pass_messages.conv_blk
        if len(blks) > _cabi.RGNN_MAX_CONV:
            raise _cabi.RgnnError('too many conv blocks')
        det.n_conv = len(blks)
        for l, blk in enumerate(blks):
            if blk.match_channels or blk.in_extra_feature_dim is not None:
                raise NotImplementedError('only the identity-residual conv block without extra features is implemented')
            msg = stack_refs(blk.msg)
            upd = stack_refs(blk.upd)
            cn = upd[-1].out_features
            dims = (cn, msg[0].in_features - 2 * cn, msg[0].out_features)
            add_stack(det.conv[l].msg, msg, dims)
            add_stack(det.conv[l].upd, upd)

        def head(stem, head_mod):
            return stack_refs(stem) + [linear_ref_of(head_mod.head[0]), linear_ref_of(head_mod.head[1])]

        add_stack(det.head_node, head(model.predict_node.stem, model.predict_node.pred_cls))
        add_stack(det.head_offset, head(model.predict_offset.stem, model.predict_offset.pred_offsets))
        add_stack(det.link_node, stack_refs(model.predict_link.compute_edge.stem))
        add_stack(det.head_link, head(model.predict_link.stem, model.predict_link.pred_cls))
        add_stack(det.class_node, stack_refs(model.predict_class.stem))
        add_stack(det.head_class, [linear_ref_of(model.predict_class.pred_cls.head[0]),
                                   linear_ref_of(model.predict_class.pred_cls.head[1])])
        _check_params(tab.tensors)
        self.tab, self.det, self.plan = tab, det, plan
        self.packed = torch.empty(tab.packed_floats(), dtype=torch.float32, device=tab.tensors[0].device)
        self.packed_versions = None     # parameter versions the packed images were built from (None: stale)
        self.refill(grads)

    def ensure_packed(self, stream):
        """(Re)build the packed / split weight images only when a parameter changed since the last call: ATen bumps
        `_version` on every in-place update (optimizers, load_state_dict, .copy_); code that writes parameter memory
        behind ATen's back (DataParallelTrainer's fused SGD kernel) calls invalidate_packed()."""
        versions = tuple(t._version for t in self.tab.tensors)
        if self.packed_versions != versions:
            check(lib().rgnn_pack_detector(C.byref(self.det), stream), 'rgnn_pack_detector')
            self.packed_versions = versions

    def invalidate_packed(self):
        self.packed_versions = None

    def refill(self, grads=None):
        """Write the parameter (and gradient) pointers into the C structs.  The gradient-free fill is idempotent as long as no
        parameter storage moved -- detector_table() rebuilds the whole table when one did -- so repeated inference calls skip it."""
        if grads is None and getattr(self, '_filled_plain', False):
            return
        off = 0
        for cstack, slots in self.plan:
            for i, s in enumerate(slots):
                off += self.tab.fill(s, cstack.layer[i], self.packed, off, grads)
        self._filled_plain = grads is None

    def key(self):
        return tuple(t.data_ptr() for t in self.tab.tensors)


def _param_slots(model):
    """(owning module, name, Parameter) of every parameter: re-validated per call with one dict lookup each instead of a walk over
    the module tree (model.parameters() costs ~0.4 ms per call for the 184 tensors of the detector: a third of a single-frame call)."""
    out = []
    for mod in model.modules():
        for name, p in mod._parameters.items():
            if p is not None:
                out.append((mod, name, p))
    return out


def detector_table(model) -> DetectorTable:
    """Cached per module; rebuilt when any parameter storage moved (e.g. after .to(device)) or a Parameter object was replaced
    (load_state_dict(assign=True), manual re-assignment).  Replacing a whole SUB-MODULE of a model that has already run is not
    detected: delete `model._rgnn_table` after such surgery."""
    cached = getattr(model, '_rgnn_table', None)
    if cached is not None:
        key = tuple(p.data_ptr() for p in cached.tab.tensors)
        if key == cached._key and all(mod._parameters.get(name) is p for mod, name, p in cached._slots):
            return cached
    t = DetectorTable(model)
    t._key = t.key()
    t._slots = _param_slots(model)
    object.__setattr__(model, '_rgnn_table', t)
    return t


class DetectorFn(torch.autograd.Function):
    """Model_Inference.forward on a GraphBatch: four head outputs; backward recomputes edge activations."""

    @staticmethod
    def forward(ctx, model, gb: GraphBatch, node_features, edge_features, training, *params):
        _require_cuda(node_features, edge_features)
        nf, ef = _f32c(node_features), _f32c(edge_features)
        table = detector_table(model)
        table.refill(None)
        s = stream_ptr()
        g = gb.c_struct()
        table.ensure_packed(s)
        nbytes = lib().rgnn_detector_workspace_bytes(C.byref(table.det), C.byref(g), 1 if training else 0)
        if nbytes == 0:
            raise _cabi.RgnnError('rgnn_detector_workspace_bytes: ' + lib().rgnn_last_error().decode())
        ws = torch.empty(nbytes, dtype=torch.uint8, device=nf.device)
        dev = nf.device
        n_cls = table.det.head_node.layer[table.det.head_node.n - 1].out_features
        n_off = table.det.head_offset.layer[table.det.head_offset.n - 1].out_features
        n_lnk = table.det.head_link.layer[table.det.head_link.n - 1].out_features
        n_obj = table.det.head_class.layer[table.det.head_class.n - 1].out_features
        node_cls = torch.empty((gb.n_nodes, n_cls), dtype=torch.float32, device=dev)
        node_off = torch.empty((gb.n_nodes, n_off), dtype=torch.float32, device=dev)
        link_cls = torch.empty((gb.n_und, n_lnk), dtype=torch.float32, device=dev)
        obj_cls = torch.empty((gb.n_clusters, n_obj), dtype=torch.float32, device=dev)
        check(lib().rgnn_detector_fwd(C.byref(table.det), C.byref(g), ptr(nf), ptr(ef), ptr(node_cls), ptr(node_off),
                                      ptr(link_cls), ptr(obj_cls), ptr(ws), nbytes, 1 if training else 0, s),
              'rgnn_detector_fwd')
        if training:
            ctx.model, ctx.gb, ctx.ws, ctx.nbytes = model, gb, ws, nbytes
            ctx.save_for_backward(nf, ef, *params)
        return node_cls, node_off, link_cls, obj_cls

    @staticmethod
    def backward(ctx, g_node_cls, g_node_off, g_link, g_obj):
        nf, ef = ctx.saved_tensors[0], ctx.saved_tensors[1]
        params = ctx.saved_tensors[2:]
        needs = list(ctx.needs_input_grad[5:])
        table = detector_table(ctx.model)
        # DataParallelTrainer installs a gradient sink: views of its flat gradient buffer, keyed by parameter storage.  The kernels
        # accumulate straight into them and autograd gets no per-parameter gradients to add (184 tiny kernels per step otherwise).
        sink = getattr(ctx.model, '_rgnn_grad_sink', None)
        # only while every needed parameter's .grad still IS its view of that buffer (a zero_grad(set_to_none=True) or a foreign
        # optimizer in between falls back to ordinary autograd accumulation)
        live = table.tab.tensors
        direct = sink is not None and len(live) == len(params) and all(
            (not n) or (p.data_ptr() in sink and q.grad is not None and q.grad.data_ptr() == sink[p.data_ptr()].data_ptr())
            for p, q, n in zip(params, live, needs))
        if direct:
            views = [sink[p.data_ptr()] if n else None for p, n in zip(params, needs)]
        else:
            flat, views = flat_grads(params, needs)
        table.refill(views)
        s = stream_ptr()
        g = ctx.gb.c_struct()
        dev = nf.device

        def grad_or_zero(gt, shape):
            if gt is None:
                return torch.zeros(shape, dtype=torch.float32, device=dev)
            return _f32c(gt)
        gb = ctx.gb
        det = table.det
        g_node_cls = grad_or_zero(g_node_cls, (gb.n_nodes, det.head_node.layer[det.head_node.n - 1].out_features))
        g_node_off = grad_or_zero(g_node_off, (gb.n_nodes, det.head_offset.layer[det.head_offset.n - 1].out_features))
        g_link = grad_or_zero(g_link, (gb.n_und, det.head_link.layer[det.head_link.n - 1].out_features))
        g_obj = grad_or_zero(g_obj, (gb.n_clusters, det.head_class.layer[det.head_class.n - 1].out_features))
        check(lib().rgnn_detector_bwd(C.byref(det), C.byref(g), ptr(nf), ptr(ef), ptr(g_node_cls), ptr(g_node_off),
                                      ptr(g_link), ptr(g_obj), ptr(ctx.ws), ctx.nbytes, s), 'rgnn_detector_bwd')
        table.refill(None)
        ctx.ws = None
        if direct:
            return (None, None, None, None, None) + (None,) * len(params)
        return (None, None, None, None, None) + tuple(views)


def run_detector_then_cluster(model, gb: GraphBatch, node_features, edge_features, cluster_fn):
    """Inference with proposal extraction in ONE pass over the graph (no gradient): the detector forward with no clusters, then
    `cluster_fn(node_off, link_cls)` sets gb.cl_ptr / cl_members / n_clusters, then ONLY the per-cluster class head runs on the
    per-node stem output that is still in the forward's workspace (rgnn_detector_obj_head).  The reference evaluates the encoders
    and the heads once and the class head afterwards as well (gnn_detector.py:164-187)."""
    _require_cuda(node_features, edge_features)
    nf, ef = _f32c(node_features), _f32c(edge_features)
    table = detector_table(model)
    table.refill(None)
    s = stream_ptr()
    table.ensure_packed(s)
    g = gb.c_struct()
    nbytes = lib().rgnn_detector_workspace_bytes(C.byref(table.det), C.byref(g), 0)
    if nbytes == 0:
        raise _cabi.RgnnError('rgnn_detector_workspace_bytes: ' + lib().rgnn_last_error().decode())
    dev = nf.device
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    det = table.det
    n_cls = det.head_node.layer[det.head_node.n - 1].out_features
    n_off = det.head_offset.layer[det.head_offset.n - 1].out_features
    n_lnk = det.head_link.layer[det.head_link.n - 1].out_features
    n_obj = det.head_class.layer[det.head_class.n - 1].out_features
    node_cls = torch.empty((gb.n_nodes, n_cls), dtype=torch.float32, device=dev)
    node_off = torch.empty((gb.n_nodes, n_off), dtype=torch.float32, device=dev)
    link_cls = torch.empty((gb.n_und, n_lnk), dtype=torch.float32, device=dev)
    none = torch.empty((max(gb.n_clusters, 0), n_obj), dtype=torch.float32, device=dev)
    check(lib().rgnn_detector_fwd(C.byref(det), C.byref(g), ptr(nf), ptr(ef), ptr(node_cls), ptr(node_off), ptr(link_cls),
                                  ptr(none), ptr(ws), nbytes, 0, s), 'rgnn_detector_fwd')
    cluster_fn(node_off, link_cls)              # fills the cluster fields of gb
    g2 = gb.c_struct()
    if lib().rgnn_detector_workspace_bytes(C.byref(det), C.byref(g2), 0) > nbytes:
        raise _cabi.RgnnError('detector workspace grew with the clusters')      # the plan does not depend on them
    obj_cls = torch.empty((gb.n_clusters, n_obj), dtype=torch.float32, device=dev)
    check(lib().rgnn_detector_obj_head(C.byref(det), C.byref(g2), ptr(obj_cls), ptr(ws), nbytes, 0, s), 'rgnn_detector_obj_head')
    return node_cls, node_off, link_cls, obj_cls


def run_detector(model, gb: GraphBatch, node_features, edge_features, training: Optional[bool] = None):
    params = detector_table(model).tab.tensors      # ParamTable order == order of the returned gradients
    if training is None:
        training = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    return DetectorFn.apply(model, gb, node_features, edge_features, bool(training), *params)
