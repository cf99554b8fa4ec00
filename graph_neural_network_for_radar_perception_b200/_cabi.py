"""ctypes binding of include/rgnn.h (librgnn.so).  This is the only bridge between the PyTorch-facing
modules and the CUDA kernels; there is no other compute path and no CPU fallback -- if the library is
missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

RGNN_MAX_STACK = 8
RGNN_MAX_CONV = 16

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'csrc', 'librgnn.so')

c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
c_double_p = C.POINTER(C.c_double)


class RgnnError(RuntimeError):
    pass


class rgnn_linear(C.Structure):
    _fields_ = [('weight', C.c_void_p), ('weight_t', C.c_void_p), ('bias', C.c_void_p),
                ('norm_scale', C.c_void_p), ('norm_shift', C.c_void_p),
                ('grad_weight', C.c_void_p), ('grad_bias', C.c_void_p),
                ('grad_norm_scale', C.c_void_p), ('grad_norm_shift', C.c_void_p),
                ('in_features', C.c_int), ('out_features', C.c_int), ('activation', C.c_int), ('reserved', C.c_int)]


class rgnn_stack(C.Structure):
    _fields_ = [('n', C.c_int), ('reserved', C.c_int), ('layer', rgnn_linear * RGNN_MAX_STACK)]


class rgnn_conv(C.Structure):
    _fields_ = [('msg', rgnn_stack), ('upd', rgnn_stack)]


class rgnn_detector(C.Structure):
    _fields_ = [('node_enc', rgnn_stack), ('edge_enc', rgnn_stack), ('n_conv', C.c_int), ('reserved', C.c_int),
                ('conv', rgnn_conv * RGNN_MAX_CONV),
                ('head_node', rgnn_stack), ('head_offset', rgnn_stack), ('link_node', rgnn_stack),
                ('head_link', rgnn_stack), ('class_node', rgnn_stack), ('head_class', rgnn_stack)]


class rgnn_graph(C.Structure):
    _fields_ = [('n_nodes', C.c_int), ('n_edges', C.c_int), ('n_und', C.c_int), ('n_clusters', C.c_int),
                ('row_ptr', C.c_void_p), ('src', C.c_void_p), ('tgt', C.c_void_p), ('perm', C.c_void_p),
                ('und_a', C.c_void_p), ('und_b', C.c_void_p), ('cl_ptr', C.c_void_p), ('cl_members', C.c_void_p)]


class rgnn_loss_cfg(C.Structure):
    _fields_ = [('class_weights', C.c_float * 16), ('n_classes', C.c_int), ('n_edge_classes', C.c_int),
                ('w_node_cls', C.c_float), ('w_node_reg', C.c_float), ('w_edge_cls', C.c_float),
                ('w_obj_cls', C.c_float), ('focal_alpha', C.c_float), ('focal_gamma', C.c_float)]


_V = C.c_void_p
_I = C.c_int
_SZ = C.c_size_t
# name -> (restype, argtypes); mirrors include/rgnn.h one to one
SIGNATURES = {
    'rgnn_version': (_I, []),
    'rgnn_last_error': (C.c_char_p, []),
    'rgnn_set_option': (_I, [C.c_char_p, _I]),
    'rgnn_get_option': (_I, [C.c_char_p]),
    'rgnn_graph_build_workspace_bytes': (_SZ, [_I, _I, _I]),
    'rgnn_graph_build': (_I, [_V, _V, _V, _V, _I, _I, C.c_float, _I, _I, _V, _V, _V, C.c_int32, _V, _V, _SZ, _V]),
    'rgnn_graph_finalize_workspace_bytes': (_SZ, [_I, _I]),
    'rgnn_graph_finalize': (_I, [_V, _V, _I, _I, _V, _V, _V, _V, _V, _V, _SZ, _V]),
    'rgnn_csr_from_edge_index_workspace_bytes': (_SZ, [_I, _I]),
    'rgnn_csr_from_edge_index': (_I, [_V, _V, _I, _I, _V, _V, _V, _V, _V, _V, _V, _V, _SZ, _V]),
    'rgnn_graph_features': (_I, [_V, _V, _V, _V, _V, _V, _V, _V, _V, _I, _I, _V, _V, _I,
                                 C.c_double, C.c_double, C.c_double, C.c_double, _I, _I, _V, _V, _V]),
    'rgnn_packed_weight_floats': (_SZ, [_I, _I]),
    'rgnn_packed_conv_msg0_floats': (_SZ, [_I, _I, _I]),
    'rgnn_pack_linear': (_I, [_V, _I, _I, _V, _V]),
    'rgnn_pack_stack': (_I, [C.POINTER(rgnn_stack), _V]),
    'rgnn_pack_conv': (_I, [C.POINTER(rgnn_conv), _V]),
    'rgnn_pack_detector': (_I, [C.POINTER(rgnn_detector), _V]),
    'rgnn_ffn_stack_fwd': (_I, [C.POINTER(rgnn_stack), _V, _I, _V, _V]),
    'rgnn_ffn_stack_bwd_workspace_bytes': (_SZ, [C.POINTER(rgnn_stack)]),
    'rgnn_ffn_stack_bwd': (_I, [C.POINTER(rgnn_stack), _V, _V, _I, _V, _V, _SZ, _V]),
    'rgnn_conv_block_fwd': (_I, [C.POINTER(rgnn_conv), C.POINTER(rgnn_graph), _V, _V, _V, _V, _V, _V]),
    'rgnn_conv_block_bwd_workspace_bytes': (_SZ, [C.POINTER(rgnn_conv), C.POINTER(rgnn_graph)]),
    'rgnn_conv_block_bwd': (_I, [C.POINTER(rgnn_conv), C.POINTER(rgnn_graph), _V, _V, _V, _V, _V, _V, _V, _V, _SZ, _V]),
    'rgnn_conv_edges_fwd': (_I, [C.POINTER(rgnn_conv), C.POINTER(rgnn_graph), _V, _V, _V, _V]),
    'rgnn_split_edge_embedding_words': (_SZ, [_I]),
    'rgnn_split_edge_embedding': (_I, [_V, _I, _V, _V]),
    'rgnn_conv_edges_f16_fwd': (_I, [C.POINTER(rgnn_conv), C.POINTER(rgnn_graph), _V, _V, _V, _V]),
    'rgnn_edge_encoder_f16_fwd': (_I, [C.POINTER(rgnn_stack), _V, _V, _I, _V, _V, _V]),
    'rgnn_conv_layer_f16_fwd': (_I, [C.POINTER(rgnn_conv), C.POINTER(rgnn_conv), C.POINTER(rgnn_graph), _V, _V, _V, _V, _V, _V, _V]),
    'rgnn_conv_msg_bwd_workspace_bytes': (_SZ, [C.POINTER(rgnn_conv), C.POINTER(rgnn_graph)]),
    'rgnn_conv_msg_bwd': (_I, [C.POINTER(rgnn_conv), C.POINTER(rgnn_graph), _V, _V, _V, _V, _V, _V, _SZ, _V]),
    'rgnn_conv_msg_f16_bwd': (_I, [C.POINTER(rgnn_conv), C.POINTER(rgnn_graph), _V, _V, _V, _V, _V, _V, _SZ, _V]),
    'rgnn_detector_workspace_bytes': (_SZ, [C.POINTER(rgnn_detector), C.POINTER(rgnn_graph), _I]),
    'rgnn_detector_fwd': (_I, [C.POINTER(rgnn_detector), C.POINTER(rgnn_graph), _V, _V, _V, _V, _V, _V, _V, _SZ, _I, _V]),
    'rgnn_cluster_workspace_bytes': (_SZ, [_I]),
    'rgnn_cluster_links': (_I, [_V, _V, _V, _V, _I, _I, C.c_float, _V, _V, _V, _V, _V, _SZ, _V]),
    'rgnn_cluster_radius': (_I, [_V, _V, _I, _I, _I, C.c_float, _V, _V, _V, _V, _V, _SZ, _V]),
    'rgnn_cluster_proposals': (_I, [_V, _V, _V, _I, _V, _V, _I, _V, _V, _V, _V, _V, _V]),
    'rgnn_accumulate_workspace_bytes': (_SZ, [_I]),
    'rgnn_accumulate_windows': (_I, [_V] * 13 + [_I, _I, _I, C.c_float, C.c_float, C.c_float, C.c_float] + [_V] * 11 + [_V, _SZ, _V]),
    'rgnn_segment_max_fwd': (_I, [_V, _I, _V, _V, _I, _V, _V, _V]),
    'rgnn_segment_max_bwd': (_I, [_V, _V, _I, _I, _I, _V, _V]),
    'rgnn_detector_obj_head': (_I, [C.POINTER(rgnn_detector), C.POINTER(rgnn_graph), _V, _V, _SZ, _I, _V]),
    'rgnn_detector_bwd': (_I, [C.POINTER(rgnn_detector), C.POINTER(rgnn_graph), _V, _V, _V, _V, _V, _V, _V, _SZ, _V]),
    'rgnn_wgrad': (_I, [_V, _I, _I, _V, _I, _I, C.c_longlong, _V, _V, _V, _V]),
    'rgnn_losses_fwdbwd': (_I, [C.POINTER(rgnn_loss_cfg), _V, _V, _V, _V, _V, _V, _V, _V, _I, _I, _I,
                                C.c_double, C.c_double, C.c_double, _V, _V, _V, _V, _V, _V, _V, _V, _V]),
    'rgnn_sgd_step': (_I, [_V, _V, _V, _SZ, C.c_float, C.c_float, C.c_float, C.c_float, _I, _V]),
    'rgnn_sgd_step_guarded': (_I, [_V, _V, _V, _SZ, C.c_float, C.c_float, C.c_float, C.c_float, _I, _V, _V]),
}

_lib = None


def lib() -> C.CDLL:
    """Load librgnn.so (once).  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RgnnError(f'{LIB_PATH} not found: build it with '
                            '`python -m graph_neural_network_for_radar_perception_b200.build` '
                            '(the CUDA extension is the only compute path)')
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)          # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        # optional process-wide numeric options (include/rgnn.h: rgnn_set_option), e.g. RGNN_OPT_TENSOR_CORES=0
        for key, val in os.environ.items():
            if key.startswith('RGNN_OPT_'):
                if l.rgnn_set_option(key[len('RGNN_OPT_'):].lower().encode(), int(val)) != 0:
                    raise RgnnError(f'{key}={val}: ' + l.rgnn_last_error().decode())
        _lib = l
    return _lib


def check(rc: int, what: str = '') -> None:
    if rc != 0:
        msg = lib().rgnn_last_error()
        raise RgnnError(f'{what} failed (code {rc}): {msg.decode() if msg else ""}')


def ptr(t) -> int:
    """Device pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
