// Host-side helpers that assemble tile programs from the C-ABI parameter structs.
#pragma once
#include "rgnn_common.cuh"

namespace rgnn {

constexpr int TR_FWD = 64;   // rows per tile, forward programs
constexpr int TR_BWD = 32;   // rows per tile, backward programs (they keep every layer's activations in smem)

struct ProgBuilder {
    Program p;
    bool ok = true;
    int n_regions = 0;
    int n_sigma = 0;
    ProgBuilder(int n_rows, int tr);
    int region(int width);                // new shared-memory region [tr][width(+pad)]; returns its id
    int sigma_slot();
    Step* add(int op, int ra = 0, int rb = 0);
    void load_rows(int ra, const float* src, int ld, int w, int dcol = 0, int padto = 0, const int* ridx = nullptr,
                   int scol = 0);
    // c_exist: columns of Wt that exist in memory (0: all Cpad; natural-layout operands of a backward are not padded)
    void gemm(int ra, int rb, const float* Wt, int ldw, int K, int k_valid, int C, int Cpad, const float* bias, int c_exist = 0);
    void norm_act(int ra, const rgnn_linear& L, int slot = -1);
    void linear(int ra, int rb, const rgnn_linear& L, int slot = -1);   // OP_LINEAR (+ OP_NORM_ACT on rb)
    void store_rows(int ra, float* dst, int ld, int w, int dcol = 0, bool accumulate = false, int scol = 0);
};

bool check_linear(const rgnn_linear& L);
int launch_program(const Program& p, cudaStream_t stream);

// geometry of a residual_graph_conv_block as this library supports it
struct ConvDims {
    int cn;   // node channels (in == out, identity residual)
    int ce;   // edge channels
    int h;    // hidden width of msg.0
};
bool conv_dims(const rgnn_conv& c, ConvDims* d);
// packed layout of msg.0: [Wt_P (cn_pad x round_up(2h,64))] [Wt_c (ce_pad x round_up(h,64))] [WP_nat (2h x round_up(cn,64))]
//                          [bias2h (round_up(2h,64))] [tensor-core operands, rgnn_mp_tc.cu]
//   Wt_P  : k-major node projection  [x W_target^T | x W_source^T]      (forward, per node)
//   Wt_c  : k-major edge part W_edge^T                                  (forward, per edge)
//   WP_nat: rows = projection column c, cols = node channel k           (backward: dX = dP * WP_nat)
inline size_t conv_msg0_proj_floats(const ConvDims& d) { return (size_t)round_up(d.cn, 8) * round_up(2 * d.h, 64); }
inline size_t conv_msg0_edge_floats(const ConvDims& d) { return (size_t)round_up(d.ce, 8) * round_up(d.h, 64); }
inline size_t conv_msg0_projnat_floats(const ConvDims& d) { return (size_t)2 * d.h * round_up(d.cn, 64); }
//   bias2h: [msg.0.bias (h) | zeros]  -- the Linear bias rides on the target half of the node projection P, so the
//           per-edge kernels add nothing but P_t[target] + P_s[source]
inline size_t conv_msg0_bias_floats(const ConvDims& d) { return (size_t)round_up(2 * d.h, 64); }
inline size_t conv_msg0_tc_offset(const ConvDims& d) {
    return conv_msg0_proj_floats(d) + conv_msg0_edge_floats(d) + conv_msg0_projnat_floats(d) + conv_msg0_bias_floats(d);
}

}  // namespace rgnn
