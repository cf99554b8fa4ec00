// Host-side helpers that assemble tile programs from the C-ABI parameter structs.
#pragma once
#include "rgnn_common.cuh"

namespace rgnn {

struct ProgBuilder {
    Program p;
    bool ok = true;
    explicit ProgBuilder(int n_rows) {
        memset(&p, 0, sizeof(p));
        p.n_rows = n_rows;
    }
    bool add(int op, int i0 = 0, int i1 = 0, int i2 = 0, int i3 = 0, const void* p0 = nullptr,
             const void* p1 = nullptr, const void* p2 = nullptr, const void* p3 = nullptr);
    void linear(const rgnn_linear& L);                       // OP_LINEAR (+ OP_NORM_ACT)
    void stack(const rgnn_stack& s, int first = 0, int last = -1);
    void load_rows(const float* src, int ld, int w, int dcol = 0, int padto = 0, const int* ridx = nullptr);
    void store_rows(float* dst, int ld, int w, int dcol = 0);
};

bool check_linear(const rgnn_linear& L);
int launch_fwd(const Program& p, cudaStream_t stream);
int launch_bwd(const Program& p, cudaStream_t stream);

// geometry of a residual_graph_conv_block as this library supports it
struct ConvDims {
    int cn;   // node channels (in == out, identity residual)
    int ce;   // edge channels
    int h;    // hidden width of msg.0
};
bool conv_dims(const rgnn_conv& c, ConvDims* d);
// packed layout of msg.0: [Wt_P (cn_pad x round_up(2h,64))] [Wt_c (ce_pad x round_up(h,64))]
inline size_t conv_msg0_proj_floats(const ConvDims& d) { return (size_t)round_up(d.cn, 8) * round_up(2 * d.h, 64); }
inline size_t conv_msg0_edge_floats(const ConvDims& d) { return (size_t)round_up(d.ce, 8) * round_up(d.h, 64); }

}  // namespace rgnn
