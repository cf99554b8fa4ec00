/*
 * rgnn.h -- C-ABI of librgnn.so, the B200 (sm_100a) implementation of the radar GNN hot path of
 * UditBhaskar19/GRAPH_NEURAL_NETWORK_FOR_RADAR_PERCEPTION.
 *
 * The reference is pure Python/PyTorch (no FFI of its own), so the "binding" a maintainer adds is a ctypes
 * stub; INTEGRATION.md shows it.  Every entry point below names the reference interface it replaces.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host; the caller allocates every output
 *     and every workspace (sizes from the *_workspace_bytes functions); the library owns no memory.
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous w.r.t. the host and performs
 *     no hidden synchronisation or allocation.
 *   - return value 0 = ok; otherwise an RGNN_ERR_* code, with text from rgnn_last_error() (thread local).
 *   - floating point is fp32; node / edge indices are int32 inside the library (the PyTorch-facing layer
 *     converts the reference's int64 tensors).
 *   - there is no CPU fallback: a missing CUDA device is an error.
 */
#ifndef RGNN_H_
#define RGNN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RGNN_OK 0
#define RGNN_ERR_INVALID 1
#define RGNN_ERR_CUDA 2
#define RGNN_ERR_WORKSPACE 3

#define RGNN_MAX_STACK 8   /* ffn_blocks in one nn.Sequential of the reference */
#define RGNN_MAX_CONV 16   /* residual_graph_conv_block count */

int rgnn_version(void);
const char* rgnn_last_error(void);

/* Process-wide numeric options (the reference has none: it always computes in fp32).
 *   "tf32_passes"  3 (default) = 3xTF32 split-operand tensor-core products, fp32-equivalent (parity mode);
 *                  1           = single TF32 pass (faster, ~1e-3 relative error per layer; never the default)
 *   "tensor_cores" 1 (default) = tcgen05 kernels where the channel plan allows; 0 = FFMA tile programs only.
 *   "tensor_cores_bwd" 1 (default) = tcgen05 backward of the message function; 0 = FFMA backward tile programs.
 *   "f16_fwd" / "f16_bwd" / "f16_node_bwd" / "f16_edge_enc"  1 (default) = the fixed-shape fp16-split kernels of round 2; 0 = the round-1 kernels
 *   "f16_passes"   3 (default) = hi/lo split products, fp32-equivalent; 1 = plain fp16 operands, fp32 accumulate (reduced precision:
 *                  5e-3 of each output's largest magnitude, tests/test_model_gpu.py; never the default)
 *   "f16_chain"    1 (default) = fixed-shape fp16-split kernels for the 64-wide stems / heads / node update; 0 = the row-MLP interpreter
 *   "rows_prefetch" 1 (default) = row-owning kernels (node update, 64-wide chains, edge encoder) request the next tile's input rows into
 *                  L2 one tile ahead (prefetch.global.L2; results unchanged); 0 = off (A/B aid)
 *   "f16_stagers"  2 .. 6 (default 6) staging warps of the forward message kernel (developer aid: profiles/README.md)
 *   "debug"        bit mask of developer aids: 8 = per-role cycle counters / event timeline of the message kernels (synchronises)
 * rgnn_get_option returns -1 for an unknown name. */
int rgnn_set_option(const char* name, int value);
int rgnn_get_option(const char* name);

/* ------------------------------------------------------------------------------------------------
 * One reference `ffn_block` (modules/neural_net/common.py:185-205): nn.Linear, optional
 * channel_normalization (:208-220; scalar affine `std`,`mu`), optional LeakyReLU(0.01) (:256-267).
 * A bare nn.Linear (FFN_TaskSpecificHead's last layer, gnn_blocks.py:185-188) has norm_* NULL, activation 0.
 * grad_* are written (accumulated, +=) by the *_bwd entry points and may be NULL.
 * ---------------------------------------------------------------------------------------------- */
typedef struct rgnn_linear {
    const float* weight;      /* (out_features, in_features) row-major: the PyTorch state_dict layout   */
    const float* weight_t;    /* packed copy from rgnn_pack_linear: (round_up(in,8), round_up(out,64))   */
    const float* bias;        /* (out_features) or NULL                                                 */
    const float* norm_scale;  /* channel_normalization.std, shape (1,), or NULL when there is no norm   */
    const float* norm_shift;  /* channel_normalization.mu,  shape (1,)                                  */
    float* grad_weight;       /* same layouts as the parameters; NULL = not wanted                      */
    float* grad_bias;
    float* grad_norm_scale;
    float* grad_norm_shift;
    int in_features;
    int out_features;
    int activation;           /* 1 = LeakyReLU(0.01), 0 = none */
    int reserved;
} rgnn_linear;

typedef struct rgnn_stack {   /* an nn.Sequential of ffn_blocks (graph_feature_encoding.encoder, *.stem ...) */
    int n;
    int reserved;
    rgnn_linear layer[RGNN_MAX_STACK];
} rgnn_stack;

/* residual_graph_conv_block (gnn_blocks.py:45-113): msg = [ffn(2*Cn+Ce -> H), ffn(H -> Cn)], upd = [ffn(2*Cn -> Cn)] */
typedef struct rgnn_conv {
    rgnn_stack msg;
    rgnn_stack upd;
} rgnn_conv;

/* Model_Inference (modules/neural_net/gnn/gnn_detector.py:31-201), parameters only */
typedef struct rgnn_detector {
    rgnn_stack node_enc;       /* encode_node_feat.encoder                      gnn_blocks.py:19-42   */
    rgnn_stack edge_enc;       /* encode_edge_feat.encoder                                            */
    int n_conv;
    int reserved;
    rgnn_conv conv[RGNN_MAX_CONV]; /* pass_messages.conv_blk                    gnn_blocks.py:116-164 */
    rgnn_stack head_node;      /* predict_node: stem + pred_cls.head (last = bare Linear)  :200-234   */
    rgnn_stack head_offset;    /* predict_offset: stem + pred_offsets.head                 :237-271   */
    rgnn_stack link_node;      /* predict_link.compute_edge.stem (node-wise part)          :274-298   */
    rgnn_stack head_link;      /* predict_link.stem + pred_cls.head (per undirected edge)  :301-344   */
    rgnn_stack class_node;     /* predict_class.stem (node-wise part)                      :347-389   */
    rgnn_stack head_class;     /* predict_class.pred_cls.head (per cluster, after max-pool)           */
} rgnn_detector;

/* A batch of per-frame graphs packed block-diagonally (node ids are global within the batch).
 * The edges are held target-major (CSR over edge_index[1]) so that the sum-aggregation of
 * MessagePassing.propagate (gnn_blocks.py:106; PyG scatter-add onto edge_index[1]) is a segmented sum. */
typedef struct rgnn_graph {
    int n_nodes, n_edges, n_und, n_clusters;
    const int32_t* row_ptr;    /* (n_nodes+1) CSR over targets                                         */
    const int32_t* src;        /* (n_edges)   source node (edge_index[0]) of the k-th target-major edge */
    const int32_t* tgt;        /* (n_edges)   target node (edge_index[1])                              */
    const int32_t* perm;       /* (n_edges)   row of the caller's edge_features / edge_index for edge k */
    const int32_t* und_a;      /* (n_und) undirected links r<c in row-major order = nonzero(triu(adj,1)) */
    const int32_t* und_b;      /*          (gnn_blocks.py:295-296)                                      */
    const int32_t* cl_ptr;     /* (n_clusters+1) cluster_node_idx lists flattened (gnn_blocks.py:384-386) */
    const int32_t* cl_members; /* (cl_ptr[n_clusters]) */
} rgnn_graph;

/* ------------------------------------------------------------------------------------------------
 * (1) graph construction -- replaces modules/compute_features/graph_features.py
 * ---------------------------------------------------------------------------------------------- */

/* compute_adjacency_information (graph_features.py:58-84) and _v2 (:87-114) for a batch of frames.
 * px,py: (n_points) f32, frames concatenated; frame_ptr_host: (n_frames+1) host array of point offsets.
 * Outputs (all caller-allocated):
 *   degree      (n_points) int32   radius-gate neighbour count, D <= eps2, diagonal excluded (:11-22,:78)
 *   row_ptr     (n_points+1) int32 CSR over edge_index[0] in reference order (== CSR over targets, the
 *                                  adjacency is symmetric)
 *   col         (edge_capacity) int32  edge_index[1], GLOBAL node ids, ascending within a row  (:79)
 *   n_edges_out (1) int32          number of directed edges written (device scalar)
 * kNN ties are ordered by (d2, index) (the reference's argsort is unstable there).
 * Returns RGNN_ERR_WORKSPACE if edge_capacity is too small (checked on device; n_edges_out holds the need). */
size_t rgnn_graph_build_workspace_bytes(int n_points, int n_frames, int knn);
int rgnn_graph_build(const float* px, const float* py, const int32_t* frame_ptr_dev, const int32_t* frame_ptr_host,
                     int n_frames, int n_points, float eps2, int knn, int union_radius,
                     int32_t* degree, int32_t* row_ptr, int32_t* col, int32_t edge_capacity, int32_t* n_edges_out,
                     void* workspace, size_t workspace_bytes, void* stream);

/* Expand a symmetric row-major CSR (from rgnn_graph_build) into everything rgnn_graph needs plus the
 * reference-facing edge_index: src_of_row (n_edges) = edge_index[0] (global ids), perm = reverse-edge
 * permutation, und_a/und_b = pairs with row<col in order, n_und_out device scalar. */
size_t rgnn_graph_finalize_workspace_bytes(int n_points, int n_edges);
int rgnn_graph_finalize(const int32_t* row_ptr, const int32_t* col, int n_points, int n_edges,
                        int32_t* row_of_edge, int32_t* perm, int32_t* und_a, int32_t* und_b, int32_t* n_und_out,
                        void* workspace, size_t workspace_bytes, void* stream);

/* General path: build the target-major CSR from a caller-supplied edge_index (any order, int64 like the
 * reference API, node ids already global).  und_* are the edges with edge_index[0] < edge_index[1] in the
 * caller's order (identical to nonzero(triu(adj,1)) when edge_index is in reference order). */
size_t rgnn_csr_from_edge_index_workspace_bytes(int n_nodes, int n_edges);
int rgnn_csr_from_edge_index(const int64_t* edge_src, const int64_t* edge_dst, int n_nodes, int n_edges,
                             int32_t* row_ptr, int32_t* src, int32_t* tgt, int32_t* perm,
                             int32_t* und_a, int32_t* und_b, int32_t* n_und_out,
                             void* workspace, size_t workspace_bytes, void* stream);

/* compute_node_features (graph_features.py:117-144, include_region_confidence=True) and
 * compute_edge_features (:147-164) for a batch; outputs are the float32 tensors of datagen_gnn.py:121-122.
 * t_min/t_max per frame come from frame_ptr (device). edge rows follow (edge_row, edge_col) = edge_index in
 * reference order.  node_features (n_points,6), edge_features (n_edges,7); either may be NULL to skip it.
 * range_in_f64 / azimuth_in_f64 say whether NumPy evaluates (r - max)/(min - max) in float64 (max given as an
 * np.float64 scalar, as datagen_gnn.py:77 does for the range) or in float32 (Python float, NEP 50 weak scalar,
 * as for the azimuth np.pi*0.5 of datagen_gnn.py:75). */
int rgnn_graph_features(const float* px, const float* py, const float* vx, const float* vy, const float* vr,
                        const float* rcs, const int64_t* timestamp_us, const int32_t* degree,
                        const int32_t* frame_ptr_dev, int n_frames, int n_points,
                        const int32_t* edge_row, const int32_t* edge_col, int n_edges,
                        double min_range, double max_range, double min_azimuth, double max_azimuth,
                        int range_in_f64, int azimuth_in_f64,
                        float* node_features, float* edge_features, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (2)-(4) model forward -- replaces gnn_blocks.py / gnn_detector.py forward passes
 * ---------------------------------------------------------------------------------------------- */

/* Packed k-major copy of one Linear weight for the tile GEMM: dst (round_up(in,8), round_up(out,64)), zero padded. */
size_t rgnn_packed_weight_floats(int in_features, int out_features);
int rgnn_pack_linear(const float* weight, int in_features, int out_features, float* weight_t, void* stream);
/* Pack every Linear of a stack / conv block / detector (weight -> weight_t) in one launch.
 * The first Linear of a conv block's msg stack (in = 2*Cn+Ce) is packed as two operands back to back:
 * the node projection [W_target^T | W_source^T] (round_up(Cn,8) x round_up(2H,64)) followed by the edge part
 * W_edge^T (round_up(Ce,8) x round_up(H,64)), the natural-layout projection for the backward and -- for the channel
 * plan of the reference configuration (Cn=Ce=64, H=128) -- the hi/lo split tensor-core operands of msg.0's edge
 * part and of msg.1; its weight_t buffer needs rgnn_packed_conv_msg0_floats floats. */
size_t rgnn_packed_conv_msg0_floats(int node_channels, int edge_channels, int hidden);
int rgnn_pack_stack(const rgnn_stack* stack, void* stream);
int rgnn_pack_conv(const rgnn_conv* blk, void* stream);
int rgnn_pack_detector(const rgnn_detector* net, void* stream);

/* nn.Sequential of ffn_blocks applied row-wise (graph_feature_encoding.forward gnn_blocks.py:41-42,
 * the *.stem stacks and FFN_TaskSpecificHead.forward :196-197).  x (n_rows, in), y (n_rows, out). */
int rgnn_ffn_stack_fwd(const rgnn_stack* stack, const float* x, int n_rows, float* y, void* stream);
int rgnn_ffn_stack_bwd(const rgnn_stack* stack, const float* x, const float* grad_y, int n_rows,
                       float* grad_x /* NULL = not wanted */, void* workspace, size_t workspace_bytes, void* stream);
size_t rgnn_ffn_stack_bwd_workspace_bytes(const rgnn_stack* stack);

/* residual_graph_conv_block.forward (gnn_blocks.py:96-113): x (N,Cn), edge embedding e (E,Ce) in TARGET-MAJOR
 * order (row k belongs to edge k of g), out (N,Cn).  scratch: agg (N,Cn) and proj (N,2H) caller-allocated. */
int rgnn_conv_block_fwd(const rgnn_conv* blk, const rgnn_graph* g, const float* x, const float* e,
                        float* out, float* agg, float* proj, void* stream);

/* Backward of ONE stand-alone block (torch autograd over gnn_blocks.py:45-113): x, e (target-major), and agg / proj exactly as
 * rgnn_conv_block_fwd left them; d_out (n_nodes, cn) -> dx (n_nodes, cn), de (n_edges, ce, target-major); parameter gradients are
 * ACCUMULATED into blk's grad_* pointers.  Generic tile programs (every channel plan of the training envelope); a training run of the
 * detector goes through rgnn_detector_bwd, which owns the fused tensor-core kernels. */
size_t rgnn_conv_block_bwd_workspace_bytes(const rgnn_conv* blk, const rgnn_graph* g);
int rgnn_conv_block_bwd(const rgnn_conv* blk, const rgnn_graph* g, const float* x, const float* e, const float* agg, const float* proj,
                        const float* d_out, float* dx, float* de, void* workspace, size_t workspace_bytes, void* stream);

/* The message + aggregation half of the block alone (MessagePassing.propagate, gnn_blocks.py:106-113):
 * agg[t] = sum over edges s->t of msg(x_t, x_s, e); proj = the hoisted node projections written by
 * rgnn_conv_block_fwd (or by the previous block).  This is the dominant kernel of the forward; bench.py times it. */
int rgnn_conv_edges_fwd(const rgnn_conv* blk, const rgnn_graph* g, const float* e, const float* proj, float* agg, void* stream);
/* The fp16-split message kernels (csrc/rgnn_mp_f16.cu, rgnn_mp_bwd_f16.cu) read the edge embedding PRE-SPLIT and TILED: per
 * edge 64 fp16 hi + 64 fp16 lo values of 16 x the embedding (hi = fp16(16 e), lo = fp16(16 e - hi); 256 bytes per edge like the
 * fp32 row), stored per tile of 128 edges as one 32 KB block [hi | lo][chunk of 8 channels][edge][8 fp16] (the chunk-major operand
 * image of the tile: coalesced for row-owning threads, one bulk copy per image).  The embedding is produced once per forward and
 * read by every conv block, so it is split once.  The buffer holds WHOLE tiles: rgnn_split_edge_embedding_words(n_edges) 32-bit words.
 * rgnn_split_edge_embedding: e (n_edges, 64) fp32, target-major -> e_split.
 * rgnn_conv_edges_f16_fwd: rgnn_conv_edges_fwd on the split rows (64 / 64 / 128 channel plan; other plans return an error). */
size_t rgnn_split_edge_embedding_words(int n_edges);
int rgnn_split_edge_embedding(const float* e, int n_edges, void* e_split, void* stream);
int rgnn_conv_edges_f16_fwd(const rgnn_conv* blk, const rgnn_graph* g, const void* e_split, const float* proj, float* agg,
                            void* stream);
/* graph_feature_encoding of the edges (reference plan <= 7 -> 256 -> 128 -> 128 -> 64) as ONE fixed-shape fp16-split kernel
 * (csrc/rgnn_edge_enc_f16.cu): row k of the output is the embedding of feature row perm[k] (perm nullable), written in the
 * pre-split format above (e_split) and optionally also as fp32 (e_f32, nullable). */
int rgnn_edge_encoder_f16_fwd(const rgnn_stack* enc, const float* edge_features, const int32_t* perm, int n_edges, void* e_split,
                              float* e_f32, void* stream);
/* One whole layer as the detector forward runs it: message function + aggregation on the split rows, then the node update
 * out = x + upd(cat(x, agg)) and -- when `next` is given -- the hoisted projection of the next block into proj_next. */
int rgnn_conv_layer_f16_fwd(const rgnn_conv* blk, const rgnn_conv* next, const rgnn_graph* g, const float* x, const void* e_split,
                            const float* proj, float* out, float* agg, float* proj_next, void* stream);
/* Backward of the message function + aggregation of one block (autograd of gnn_blocks.py:106-113): given d(agg) it
 * recomputes the edge activations and produces d(proj) (N, 2h), d(e) (E, ce, overwritten) and -- through the grad_* fields
 * of blk->msg -- the weight / bias / norm-scalar gradients of msg.0 (edge columns) and msg.1 (accumulated).
 * e is the fp32 edge embedding (target-major).  Used by rgnn_detector_bwd; exported for measurement and tests. */
size_t rgnn_conv_msg_bwd_workspace_bytes(const rgnn_conv* blk, const rgnn_graph* g);
int rgnn_conv_msg_bwd(const rgnn_conv* blk, const rgnn_graph* g, const float* e, const float* proj, const float* dagg,
                      float* dproj, float* de, void* workspace, size_t workspace_bytes, void* stream);
/* The same on the pre-split rows of rgnn_split_edge_embedding / rgnn_edge_encoder_f16_fwd (what the detector backward passes):
 * ONE fused kernel (csrc/rgnn_mp_bwd_f16.cu: recompute, data gradients and both weight gradients on fp16-split operands)
 * plus the gather of d(proj).  Reference plan only (64 / 64 / 128); same workspace as rgnn_conv_msg_bwd. */
int rgnn_conv_msg_f16_bwd(const rgnn_conv* blk, const rgnn_graph* g, const void* e_split, const float* proj, const float* dagg,
                          float* dproj, float* de, void* workspace, size_t workspace_bytes, void* stream);

/* Model_Inference.forward with cluster_node_idx given (gnn_detector.py:141-162).
 * edge_features rows are in the caller's order (g->perm maps them).  Outputs: node_cls (N,7), node_off (N,2),
 * link_cls (n_und,2), obj_cls (n_clusters,7).  With training != 0 the node-level activations needed by
 * rgnn_detector_bwd stay in the workspace (edge activations are never stored; they are recomputed). */
size_t rgnn_detector_workspace_bytes(const rgnn_detector* net, const rgnn_graph* g, int training);
int rgnn_detector_fwd(const rgnn_detector* net, const rgnn_graph* g, const float* node_features,
                      const float* edge_features, float* node_cls, float* node_off, float* link_cls, float* obj_cls,
                      void* workspace, size_t workspace_bytes, int training, void* stream);

/* Proposal extraction (gnn_detector.py:164-187; Simple_DBSCAN, modules/inference/clustering.py:8-92) on the device.
 * The reference grows clusters by breadth-first search over a dense adjacency: that labels the connected components of the
 * graph, numbered by their smallest member, members in ascending order.  Two ways to get the graph, as in the reference:
 *   rgnn_cluster_links:  undirected pair k = (und_a[k], und_b[k]) is kept iff its predicted class is 1 (link_logits
 *                        (n_und,2), arg-max with ties to class 0) and NOT sqrt(dx^2+dy^2) >= eps between the predicted
 *                        centres xy (n_nodes,2)                                                (clustering.py:9-24)
 *   rgnn_cluster_radius: all pairs of one frame with dx^2+dy^2 <= eps (eps compares with the SQUARED distance, :27-41);
 *                        frame_ptr_dev (n_frames+1), max_frame_nodes = largest frame (grid sizing)
 * Outputs: cluster_id (n_nodes) = meas_to_cluster_id, n_clusters_out (device scalar), cl_ptr (n_nodes+1; entries past
 * n_clusters repeat n_nodes) and cl_members (n_nodes) = the cluster_node_idx lists flattened, ready for rgnn_graph. */
size_t rgnn_cluster_workspace_bytes(int n_nodes);
int rgnn_cluster_links(const float* xy, const int32_t* und_a, const int32_t* und_b, const float* link_logits, int n_nodes,
                       int n_und, float eps, int32_t* cluster_id, int32_t* n_clusters_out, int32_t* cl_ptr,
                       int32_t* cl_members, void* workspace, size_t workspace_bytes, void* stream);
int rgnn_cluster_radius(const float* xy, const int32_t* frame_ptr_dev, int n_frames, int n_nodes, int max_frame_nodes,
                        float eps, int32_t* cluster_id, int32_t* n_clusters_out, int32_t* cl_ptr, int32_t* cl_members,
                        void* workspace, size_t workspace_bytes, void* stream);
/* Proposals from the clusters (modules/inference/inference.py:23-47, output.py:111-118): per cluster the sample mean of
 * the member positions (px, py), covariance = sum (mean - x)(mean - x)^T / (n - 1) + noise_cov (noise_cov alone for a
 * single member), the size, and -- when node_cls (n_nodes, n_classes) is given -- the majority vote over the members'
 * arg-max class (ties to the lowest class, like torch.bincount + argmax).  noise_cov_host: 4 floats on the HOST
 * (Model_Inference.meas_noise_cov).  Outputs: mean (C,2), cov (C,2,2), size (C) int32, vote (C) int32 or NULL. */
int rgnn_cluster_proposals(const float* px, const float* py, const float* node_cls, int n_classes, const int32_t* cl_ptr,
                           const int32_t* cl_members, int n_clusters, const float* noise_cov_host, float* mean, float* cov,
                           int32_t* size, int32_t* vote, void* stream);

/* Sliding-window accumulation pre-pass: replaces the per-scan Python loop of
 * extract_and_sync_radar_data (modules/data_utils/read_data.py:227-303: stationary gate meas_selection.py:39-70,169-200;
 * vr_cartesian_vf and ego compensation meas_sync.py:15-21,44-102), the float32 casts / flip of get_data_for_datagen
 * (read_data.py:524-537), generate_gt_labels (compute_groundtruth/compute_node_labels.py:71-86),
 * grid_properties.select_meas_within_the_grid (compute_features/grid_features.py:162-173) and select_moving_data
 * (compute_features/graph_features.py:167-182) for MANY windows per call.
 * Inputs (device, detections of all windows concatenated, scans in window order): the RadarScenes radar_data fields
 * x_cc, y_cc, vr, vr_compensated, azimuth_sc, rcs (float32), timestamp (int64), label_id (uint8), has_track (uint8,
 * track_id != b''), point_scan (int32 row of scan_params for each detection); scan_params (n_scans_total, 9) float64 =
 * [R00 R01 R10 R11 tx ty mount_yaw vx_sensor vy_sensor] with [R|t] = inv(T_curr) T_prev of the window (meas_sync.py:55)
 * and the ego velocity at the sensor in the sensor frame (meas_selection.py:24-37), computed by the host in float64;
 * window_ptr_dev (n_windows+1) detection offsets; flip_dev (n_windows) uint8 or NULL (flip_along_x augmentation).
 * select != 0 keeps the detections inside [min_x,max_x) x [min_y,max_y) whose class label is not STATIC (7), select == 0
 * keeps all.  Outputs (device, capacity n_points, windows back to back, detections in their original order):
 * meas_px .. meas_rcs (float32), meas_timestamp (int64), class_labels (float32, labels.py:60-70 ids), src_index (int32,
 * index of the detection inside its window), stationary_flag (n_points uint8, per INPUT detection, or NULL), out_ptr_dev
 * (n_windows+1) offsets of the windows in the outputs = frame_ptr of rgnn_graph_build.  Positions, flags, labels and
 * selection are exact; velocities within 2 ulp (float32 cos/sin). */
size_t rgnn_accumulate_workspace_bytes(int n_windows);
int rgnn_accumulate_windows(const float* x_cc, const float* y_cc, const float* vr, const float* vr_compensated,
                            const float* azimuth_sc, const float* rcs, const int64_t* timestamp, const uint8_t* label_id,
                            const uint8_t* has_track, const int32_t* point_scan, const double* scan_params,
                            const int32_t* window_ptr_dev, const uint8_t* flip_dev, int n_windows, int n_points, int select,
                            float min_x, float max_x, float min_y, float max_y, float* meas_px, float* meas_py,
                            float* meas_vx, float* meas_vy, float* meas_vr, float* meas_rcs, int64_t* meas_timestamp,
                            float* class_labels, int32_t* src_index, uint8_t* stationary_flag, int32_t* out_ptr_dev,
                            void* workspace, size_t workspace_bytes, void* stream);

/* Per-cluster max over member rows for object_classification used on its own (gnn_blocks.py:384-386: torch.max(x[idx], dim=0) per
 * cluster).  x (n_rows, width); pooled (n_clusters, width); argrow (n_clusters, width) receives the row that supplied each maximum
 * (first in member order on ties; nullable).  The backward zero-fills dx (n_rows, width) and routes d_pooled to those rows, like
 * autograd of torch.max(dim).  The detector-level forward fuses this reduction into the class-head kernel. */
int rgnn_segment_max_fwd(const float* x, int width, const int32_t* cl_ptr, const int32_t* cl_members, int n_clusters,
                         float* pooled, int32_t* argrow, void* stream);
int rgnn_segment_max_bwd(const float* d_pooled, const int32_t* argrow, int n_clusters, int width, int n_rows, float* dx, void* stream);

/* predict_class on clusters found after the forward pass: the per-node stem output of predict_class is still in the
 * workspace of the preceding rgnn_detector_fwd call on the same graph (same `training` flag); g carries the new
 * cl_ptr / cl_members / n_clusters.  obj_cls (n_clusters, n_classes). */
int rgnn_detector_obj_head(const rgnn_detector* net, const rgnn_graph* g, float* obj_cls, void* workspace,
                           size_t workspace_bytes, int training, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (5) backward and training -- replaces torch autograd over the above (gnn/training.py:81) and
 *     Loss_Graph (gnn/loss.py:37-76)
 * ---------------------------------------------------------------------------------------------- */
int rgnn_detector_bwd(const rgnn_detector* net, const rgnn_graph* g, const float* node_features,
                      const float* edge_features, const float* grad_node_cls, const float* grad_node_off,
                      const float* grad_link_cls, const float* grad_obj_cls,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Weight gradient of one nn.Linear on the tensor cores (what autograd's addmm backward computes, common.py:185-205):
 * dst (wa x wb, row-major) += A^T B with A (rows x wa, leading dimension lda, wa <= 128) and B (rows x wb, ldb,
 * wb <= 256) row-major in device memory; colsum_a (wa) += column sums of A, colsum_b (wb) += column sums of B (bias
 * gradients); either may be NULL.  3xTF32 split products, fp32 accumulation (option "tf32_passes").
 * rgnn_detector_bwd uses it for the message-function weights; option "tensor_cores_bwd" = 0 selects the FFMA
 * tile programs for the whole backward instead. */
int rgnn_wgrad(const float* A, int lda, int wa, const float* B, int ldb, int wb, long long rows, float* dst,
               float* colsum_a, float* colsum_b, void* stream);

/* Loss_Graph.forward + its gradient w.r.t. the logits, and compute_accuracy (gnn_detector.py:23-28).
 * counts_* are the GLOBAL batch sizes each sum is divided by (loss.py:58,62,66,70) -- under data parallelism
 * they are the all-reduced counts, which is what makes per-rank sums add up to the reference's loss.
 * losses_out (4) f64: node_cls, node_reg, edge_cls, obj_cls (already weighted, this rank's share);
 * correct_out (3) int32: argmax hits for node / edge / object.  node_off_gt is already normalised.
 * counts_dev (nullable, 3 x f64 on the device): when given it overrides count_* -- the data-parallel driver all-reduces
 * the counts on the stream and the host never reads them.  dp_tail (nullable, 5 x f32, zeroed by the caller): [0] is set
 * to 1 when a loss term is NaN (the reference's skip_batch test, gnn/training.py:40-45), [1..4] += this rank's four
 * weighted loss shares; the driver keeps it behind the flat gradient buffer so that it rides in the one all-reduce. */
typedef struct rgnn_loss_cfg {
    float class_weights[16];
    int n_classes, n_edge_classes;
    float w_node_cls, w_node_reg, w_edge_cls, w_obj_cls;
    float focal_alpha, focal_gamma;
} rgnn_loss_cfg;
int rgnn_losses_fwdbwd(const rgnn_loss_cfg* cfg, const float* node_cls, const float* node_off, const float* link_cls,
                       const float* obj_cls, const int64_t* node_cls_gt, const float* node_off_gt,
                       const int64_t* link_gt, const int64_t* obj_gt, int n_nodes, int n_und, int n_clusters,
                       double count_nodes, double count_und, double count_clusters,
                       float* grad_node_cls, float* grad_node_off, float* grad_link_cls, float* grad_obj_cls,
                       double* losses_out, int32_t* correct_out, const double* counts_dev, float* dp_tail, void* stream);

/* torch.optim.SGD(momentum, weight_decay) step on a flat parameter buffer
 * (modules/set_configurations/set_param_for_training_gnn.py:46): g += wd*p; buf = mu*buf + g; p -= lr*buf.
 * grad_scale multiplies the (all-reduced) gradient first. */
int rgnn_sgd_step(float* params, const float* grads, float* momentum_buf, size_t n, float lr, float momentum,
                  float weight_decay, float grad_scale, int first_step, void* stream);
/* Same, guarded by a device flag: when *skip_flag is non-zero or NaN nothing is written.  Replaces the host-side
 * `skip_batch(total_loss)` / `torch.isnan(loss)` of the reference loop (gnn/training.py:40-45,79-84) without a
 * device-to-host read per step; under data parallelism the flag travels inside the gradient all-reduce. */
int rgnn_sgd_step_guarded(float* params, const float* grads, float* momentum_buf, size_t n, float lr, float momentum,
                          float weight_decay, float grad_scale, int first_step, const float* skip_flag, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RGNN_H_ */
