// Internal declarations shared by the forward and backward host code of the detector.
#pragma once
#include <functional>

#include "rgnn_programs.h"

namespace rgnn {

// what the forward of a training step leaves behind for tc_stack_bwd: layer outputs (post-activation), norm sigmas, the
// assembled input rows (pair sums); nullptr entries are not saved
struct TcSave {
    float* x_in;
    float* y[RGNN_MAX_STACK];
    float* sd[RGNN_MAX_STACK];
};

// Where everything lives inside the caller-provided detector workspace.
struct DetPlan {
    ConvDims d;
    bool training;
    int link_w, cls_w;
    float* x[RGNN_MAX_CONV + 1];   // node embeddings entering conv l (x[L] feeds the heads)
    float* P[RGNN_MAX_CONV];       // per-node projections of msg.0: [x W1_target^T | x W1_source^T]
    float* agg[RGNN_MAX_CONV];     // aggregated messages
    float* emb;                    // edge embedding, target-major order
    uint32_t* emb_hl;              // the same rows pre-split for the f16 message kernel: [hi 64 fp16 | lo 64 fp16] x 16 (nullptr: not used)
    float* hlink;                  // predict_link.compute_edge.stem output per node
    float* gcls;                   // predict_class.stem output per node
    // ---- backward only ----
    float* dx;       // (N, cn) running gradient w.r.t. the node embedding
    float* dP;       // (N, 2h)
    float* dagg;     // (N, cn)
    float* demb;     // (E, ce) accumulated over all conv layers
    float* dh;       // (N, link_w)
    float* dg;       // (N, cls_w)
    float* escr;     // per-edge scratch of the tensor-core backward: y1 (E,h) | dz1 (E,h) | dz2 (E,cn)
    int* sidx;       // workspace of the source-major edge index (build_src_index)
    // training with the tensor-core chain backward: saved layer outputs / sigmas of the edge encoder and the link head
    bool enc_tc_bwd, link_tc_bwd;
    TcSave enc_save, link_save;
    // node-sized stacks with the chain backward: 0 node encoder, 1 head_node, 2 head_offset, 3 link_node, 4 class_node
    bool node_tc_bwd[5];
    TcSave node_save[5];
    bool conv_tc_bwd;                  // node update of every conv block
    float* u_save[RGNN_MAX_CONV];      // update output before the residual
    float* usd_save[RGNN_MAX_CONV];    // its sigma
    float* cscr;     // scratch of tc_stack_bwd (dz per layer), shared by the two chains
    size_t bytes;
};

typedef std::function<float*(size_t)> TakeFn;
int plan_detector(const rgnn_detector& net, const rgnn_graph& g, int training, void* base, DetPlan* pl);
void plan_detector_bwd(const rgnn_detector& net, const rgnn_graph& g, const TakeFn& take, DetPlan* pl);

int stack_in(const rgnn_stack& s);
int stack_out(const rgnn_stack& s);
int run_stack_fwd(const rgnn_stack& s, const float* x, int n_rows, float* y, cudaStream_t stream, const TcSave* save = nullptr);
void add_message_layers(ProgBuilder& b, const rgnn_conv& c, const ConvDims& d, const rgnn_graph& g, const float* P,
                        int r_in, int r_mid, int r_out, int slot0, int slot1);

// tensor-core (tcgen05) message kernel, rgnn_mp_tc.cu
bool mp_tc_supported(const ConvDims& d);
size_t mp_tc_pack_floats(const ConvDims& d);      // extra floats at the end of msg.0's packed buffer
int mp_tc_pack(const rgnn_conv& c, const ConvDims& d, float* dst, cudaStream_t stream);
int run_conv_edges_tc(const rgnn_conv& c, const ConvDims& d, const rgnn_graph& g, const float* emb, const float* P,
                      const float* wpack, float* agg, cudaStream_t stream);

// fp16-split role-pipelined message kernel, rgnn_mp_f16.cu (forward; the packed images sit behind the projection images)
bool mp_f16_supported(const ConvDims& d);
int mp_f16_passes();
size_t mp_f16_pack_floats(const ConvDims& d);
size_t mp_f16_emb_words(int n_edges);
int mp_f16_pack(const rgnn_conv& c, const ConvDims& d, float* dst, cudaStream_t stream);
int mp_f16_split_emb(const float* emb, int n_edges, uint32_t* out, cudaStream_t stream);
int run_conv_edges_f16(const rgnn_conv& c, const ConvDims& d, const rgnn_graph& g, const uint32_t* emb_hl, const float* P,
                       const float* wpack, float* agg, cudaStream_t stream);
int mp_f16_set_option(const char* name, int value);
int mp_f16_get_option(const char* name);
inline size_t conv_msg0_f16_offset(const ConvDims& d);

// fixed-shape 64-wide chains with fp16-split operands, rgnn_chain_f16.cu (inference forward of the stems / heads)
size_t f16_image_floats(int in_features, int out_features);
int f16_pack_linear(const rgnn_linear& L, float* dst, cudaStream_t stream);
bool chain64_supported(const rgnn_stack& s);
int run_chain64(const rgnn_stack& s, const float* x, int ldx, const int* ia, const int* ib, int n_rows, float* y, cudaStream_t stream,
                const TcSave* save = nullptr);
size_t conv_proj_f16_floats(const ConvDims& d);
int conv_proj_f16_pack(const rgnn_conv& c, const ConvDims& d, float* dst, cudaStream_t stream);
bool conv_nodes_f16_supported(const rgnn_conv& c, const ConvDims& d);
int run_conv_nodes_f16(const rgnn_conv& c, const ConvDims& d, int n_nodes, const float* x, const float* agg, float* out,
                       const rgnn_conv* next, const float* next_proj_images, float* P_next, cudaStream_t stream,
                       float* u_save = nullptr, float* sd_save = nullptr);
inline size_t conv_msg0_proj16_offset(const ConvDims& d);
int chain_f16_set_option(const char* name, int value);
int chain_f16_get_option(const char* name);

// fixed-shape edge encoder with fp16-split operands, rgnn_edge_enc_f16.cu
bool edge_enc_f16_supported(const rgnn_stack& s);
int edge_enc_f16_pack(const rgnn_stack& s, cudaStream_t stream);
int run_edge_enc_f16(const rgnn_stack& s, const float* feat, const int* perm, int n_rows, uint32_t* emb_hl, float* emb, cudaStream_t stream,
                     const TcSave* save = nullptr);
int edge_enc_f16_set_option(const char* name, int value);
int edge_enc_f16_get_option(const char* name);

// tensor-core backward of the message function (rgnn_mp_bwd_tc.cu) and the generic weight-gradient GEMM (rgnn_wgrad_tc.cu)
bool mp_bwd_tc_supported(const ConvDims& d);
size_t mp_bwd_tc_scratch_floats(const ConvDims& d, int n_edges);
int run_conv_edges_bwd_tc(const rgnn_conv& c, const ConvDims& d, const rgnn_graph& g, const float* emb, const float* P,
                          const float* dagg, float* dP, float* demb, bool first_demb, float* scratch, const int* sptr,
                          const int* slist, cudaStream_t stream);
// fused fp16-split backward of the message function (rgnn_mp_bwd_f16.cu): dgrad + both weight gradients in one kernel
bool mp_bwd_f16_supported(const ConvDims& d);
int run_conv_edges_bwd_f16(const rgnn_conv& c, const ConvDims& d, const rgnn_graph& g, const uint32_t* emb_hl, const float* P,
                           const float* dagg, float* dP, float* demb, bool first_demb, float* scratch, const int* sptr,
                           const int* slist, cudaStream_t stream);
int mp_bwd_f16_set_option(const char* name, int value);
int mp_bwd_f16_get_option(const char* name);
// fixed-shape fp16-split node-level backward with fused weight gradients (rgnn_node_bwd_f16.cu)
bool node_bwd_f16_supported(const rgnn_conv& c, const ConvDims& d);
int run_proj_bwd_f16(const rgnn_conv& c, const ConvDims& d, const float* dP, const float* x, int n_nodes, float* dx, float* scalar,
                     cudaStream_t stream);
int run_upd_bwd_f16(const rgnn_conv& c, const ConvDims& d, int n_nodes, const float* x, const float* agg, const float* u, const float* sd,
                    float* dx, float* dagg, float* scalar, cudaStream_t stream);
bool chain64_bwd_f16_supported(const rgnn_stack& s);
int run_chain64_bwd_f16(const rgnn_stack& s, const TcSave& save, const float* x_rows, const float* y_out, const float* g_top, int n_rows,
                        float* dx, int dx_mode, const int* ia, const int* ib, float* scalar, cudaStream_t stream);
int node_bwd_f16_set_option(const char* name, int value);
int node_bwd_f16_get_option(const char* name);
// source-major index of the target-major edge list (edge positions grouped by source node), built once per backward call
size_t src_index_ints(int n_nodes, int n_edges);
int build_src_index(const rgnn_graph& g, int* ws, const int** sptr_out, const int** slist_out, cudaStream_t stream);
// device-wide exclusive scan of int32 (rgnn_graph.cu)
size_t scan_ws_ints(int n);
int exclusive_scan(const int* in, int n, int* out, int* ws, cudaStream_t stream);
int launch_wgrad_tc(const float* A, int lda, int wa, const float* B, int ldb, int wb, long long rows, float* dst, long long sm,
                    long long sn, float* colsum_a, float* colsum_b, cudaStream_t stream, const int* b_ridx = nullptr);

// tensor-core row-MLP programs, rgnn_model_tc.cu
bool tc_stack_supported(const rgnn_stack& s);
bool tc_proj_supported(const ConvDims& d);
size_t tc_proj_pack_floats(const ConvDims& d);
int tc_pack_linear(const rgnn_linear& L, cudaStream_t stream);
int tc_pack_projection(const rgnn_conv& c, const ConvDims& d, cudaStream_t stream);
inline size_t conv_msg0_f16_offset(const ConvDims& d) { return conv_msg0_tc_offset(d) + mp_tc_pack_floats(d) + tc_proj_pack_floats(d); }
inline size_t conv_msg0_proj16_offset(const ConvDims& d) { return conv_msg0_f16_offset(d) + mp_f16_pack_floats(d); }
int tc_run_stack(const rgnn_stack& s, const float* x, const int* ridx, int n_rows, float* y, cudaStream_t stream,
                 const TcSave* save = nullptr);
bool tc_stack_bwd_supported(const rgnn_stack& s);
size_t tc_stack_bwd_scratch_floats(const rgnn_stack& s, int n_rows);
// dx_mode: 0 = overwrite, 1 = accumulate, 2 = atomic pair scatter onto rows ia[], ib[] of dx (dx may be nullptr: not wanted)
int tc_stack_bwd(const rgnn_stack& s, const TcSave& save, const float* x_rows, const int* x_ridx, const float* y_out,
                 const float* g_top, int n_rows, float* scratch, float* dx, int dx_mode, const int* ia, const int* ib,
                 cudaStream_t stream);
int tc_run_node_encoder(const rgnn_stack& enc, const rgnn_conv& first, const ConvDims& d, const float* node_features, int n_nodes,
                        float* x0, float* P0, cudaStream_t stream, const TcSave* save = nullptr);
int tc_run_conv_nodes(const rgnn_conv& c, const ConvDims& d, int n_nodes, const float* x, const float* agg, float* out,
                      const rgnn_conv* next, float* P_next, cudaStream_t stream, float* u_save = nullptr, float* sd_save = nullptr);
int tc_proj_bwd(const rgnn_conv& c, const ConvDims& d, const float* dP, const float* x, int n_nodes, float* dx, cudaStream_t stream);
bool tc_conv_nodes_bwd_supported(const rgnn_conv& c, const ConvDims& d);
int tc_conv_nodes_bwd(const rgnn_conv& c, const ConvDims& d, int n_nodes, const float* x, const float* agg, const float* u,
                      const float* sd, float* dx, float* dagg, float* dz_scratch, cudaStream_t stream);
int tc_run_pairsum_stack(const rgnn_stack& s, const float* h, int ld, const int* ia, const int* ib, int n_rows, float* y,
                         cudaStream_t stream, const TcSave* save = nullptr);
int tc_run_segmax_stack(const rgnn_stack& s, const float* g, int ld, const int* ptr, const int* members, int n_rows, float* y,
                        cudaStream_t stream);
size_t tc_linear_pack_floats(int in_features, int out_features);

}  // namespace rgnn
