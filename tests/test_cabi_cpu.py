"""CPU: the C-ABI library builds/loads and exports every symbol include/rgnn.h declares; host-side module
structure matches the reference checkpoint.  No compute calls (no GPU here)."""
import os
import re

import pytest
import torch

from graph_neural_network_for_radar_perception_b200 import _cabi, config, Model_Training

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, 'include', 'rgnn.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(rgnn_[a-z0-9_]+)\s*\(', txt)))


def test_library_exports_every_declared_symbol():
    lib = _cabi.lib()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/rgnn.h but not exported by librgnn.so'
    assert set(names) == set(_cabi.SIGNATURES), set(names) ^ set(_cabi.SIGNATURES)
    assert lib.rgnn_version() >= 100


def test_pure_host_size_queries():
    lib = _cabi.lib()
    assert lib.rgnn_packed_weight_floats(7, 256) == 8 * 256 + 2 * 8 * 256 + 16 * 256       # + fp16 image pair, K zero padded to 16
    assert lib.rgnn_packed_weight_floats(64, 7) == 64 * 64 + 2 * 64 * 32 + 2 * 8 * 64 + 64 * 16      # + transposed image (backward) + fp16 image pair
    assert lib.rgnn_packed_conv_msg0_floats(64, 64, 128) == 64 * 256 + 64 * 128 + 256 * 64 + 256 + 2 * (2 * 64 * 128 + 2 * 128 * 64) + 2 * 64 * 256 + 2 * (2 * 128 * 64) \
        + (2 * 64 * 128 + 2 * 128 * 64) // 2 \
        + 2 * 64 * 128         # fp16 hi / lo images of W_e and W_2 (rgnn_mp_f16.cu) and of the two projection halves
    assert lib.rgnn_packed_conv_msg0_floats(32, 32, 64) == 32 * 128 + 32 * 64 + 128 * 64 + 128
    assert lib.rgnn_graph_build_workspace_bytes(1000, 1, 10) > 1000 * 11 * 4


def test_struct_sizes_match_header():
    import ctypes as C
    assert C.sizeof(_cabi.rgnn_linear) == 9 * 8 + 4 * 4
    assert C.sizeof(_cabi.rgnn_stack) == 8 + _cabi.RGNN_MAX_STACK * C.sizeof(_cabi.rgnn_linear)
    assert C.sizeof(_cabi.rgnn_graph) == 16 + 8 * 8


def test_state_dict_layout_is_the_reference_checkpoint(ckpt_state_dict):
    m = Model_Training(config(), 'cpu')
    res = m.load_state_dict(ckpt_state_dict, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert sum(p.numel() for p in m.parameters()) == 463144
    own = m.state_dict()
    assert list(own.keys()) == list(ckpt_state_dict.keys())     # same order as well
    for k, v in ckpt_state_dict.items():
        assert own[k].shape == v.shape, k


def test_head_init_matches_reference_constants():
    import math
    torch.manual_seed(0)
    m = Model_Training(config(), 'cpu').pred
    assert torch.allclose(m.predict_node.pred_cls.head[1].bias, torch.full((7,), -math.log(99)))
    assert torch.all(m.predict_offset.pred_offsets.head[1].bias == 0)
    assert float(m.predict_link.pred_cls.head[1].weight.std()) < 0.02


def test_cpu_tensors_are_rejected_loudly():
    m = Model_Training(config(), 'cpu').pred
    x = torch.zeros(4, 6)
    with pytest.raises(Exception):
        m.encode_node_feat(x)


def test_unsupported_configurations_raise():
    from graph_neural_network_for_radar_perception_b200.common import ffn_block
    with pytest.raises(NotImplementedError):
        ffn_block(8, 8, 'relu')
    with pytest.raises(NotImplementedError):
        ffn_block(8, 8, 'leakyrelu', norm_layer='layer_normalization')


def test_options_are_host_state_with_loud_rejection():
    """rgnn_set_option / rgnn_get_option (include/rgnn.h) are pure host state: every documented name reads back its default, a
    value outside the documented set and an unknown name are rejected with an error message, and nothing changes."""
    from graph_neural_network_for_radar_perception_b200._cabi import lib
    L = lib()
    defaults = {b'tf32_passes': 3, b'tensor_cores': 1, b'tensor_cores_bwd': 1, b'f16_fwd': 1, b'f16_bwd': 1, b'f16_node_bwd': 1,
                b'f16_edge_enc': 1, b'f16_passes': 3, b'f16_chain': 1, b'rows_prefetch': 1, b'f16_stagers': 6, b'pack_batch': 1}
    for name, want in defaults.items():
        assert L.rgnn_get_option(name) == want, name
    assert L.rgnn_get_option(b'no_such_option') == -1
    assert L.rgnn_set_option(b'no_such_option', 1) != 0
    assert b'no_such_option' in L.rgnn_last_error()
    for name, bad in ((b'f16_passes', 2), (b'rows_prefetch', 2), (b'f16_stagers', 1), (b'tf32_passes', 0)):
        assert L.rgnn_set_option(name, bad) != 0, name
        assert L.rgnn_get_option(name) == defaults[name], name
    try:
        assert L.rgnn_set_option(b'rows_prefetch', 0) == 0 and L.rgnn_get_option(b'rows_prefetch') == 0
        assert L.rgnn_set_option(b'f16_stagers', 4) == 0 and L.rgnn_get_option(b'f16_stagers') == 4
    finally:
        assert L.rgnn_set_option(b'rows_prefetch', 1) == 0
        assert L.rgnn_set_option(b'f16_stagers', 6) == 0
