// Host side of the model entry points: weight packing and the forward tile programs of the detector.
#include "rgnn_model.h"
#include "rgnn_pack.cuh"

namespace rgnn {

// ---------------------------------------------------------------------------------------------
// weight packing: k-major, zero padded copies for the tile GEMM
// ---------------------------------------------------------------------------------------------
struct PackEntry {
    const float* W;   // (C, ldW) row-major
    float* dst;       // (Kpad, ldd)
    int ldW, koff, K, C, Kpad, ldd, c0, cw, transpose;
};
constexpr int PACK_BATCH = 48;
struct PackTable {
    int n;
    PackEntry e[PACK_BATCH];
};

__global__ void pack_kernel(const __grid_constant__ PackTable tab) {
    const PackEntry& e = tab.e[blockIdx.y];
    const int tot = e.Kpad * e.cw;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += gridDim.x * blockDim.x) {
        const int k = i / e.cw, c = i - k * e.cw;
        float v = 0.f;
        if (k < e.K && c < e.C) v = e.transpose ? e.W[(size_t)c * e.ldW + e.koff + k] : e.W[(size_t)k * e.ldW + e.koff + c];
        e.dst[(size_t)k * e.ldd + e.c0 + c] = v;
    }
}

struct Packer {
    PackTable tab;
    cudaStream_t stream;
    int rc = RGNN_OK;
    explicit Packer(cudaStream_t s) : stream(s) { tab.n = 0; }
    void flush() {
        if (tab.n == 0 || rc != RGNN_OK) { tab.n = 0; return; }
        pack_kernel<<<dim3(8, tab.n), 256, 0, stream>>>(tab);
        if (cudaGetLastError() != cudaSuccess) { set_error("pack_kernel launch failed"); rc = RGNN_ERR_CUDA; }
        tab.n = 0;
    }
    void push(const PackEntry& e) {
        tab.e[tab.n] = e;
        tab.e[tab.n].transpose = 1;
        if (++tab.n == PACK_BATCH) flush();
    }
    // plain (non transposed) copy: dst[r][c] = W[r][koff + c], r < K (rows), c < C (cols), zero padded to (Kpad, cw)
    void push_nat(const PackEntry& e) {
        tab.e[tab.n] = e;
        tab.e[tab.n].transpose = 0;
        if (++tab.n == PACK_BATCH) flush();
    }
    void linear(const rgnn_linear& L) {
        if (L.weight == nullptr || L.weight_t == nullptr) { set_error("pack: null weight pointer"); rc = RGNN_ERR_INVALID; return; }
        const int Kpad = round_up(L.in_features, 8), Cpad = round_up(L.out_features, 64);
        push({L.weight, const_cast<float*>(L.weight_t), L.in_features, 0, L.in_features, L.out_features, Kpad, Cpad, 0, Cpad});
        const int r = tc_pack_linear(L, stream);     // chunked hi/lo operands of the tensor-core programs
        if (r != RGNN_OK) rc = r;
    }
    void stack(const rgnn_stack& s) { for (int i = 0; i < s.n; ++i) linear(s.layer[i]); }
    void conv(const rgnn_conv& c) {
        ConvDims d;
        if (!conv_dims(c, &d)) { rc = RGNN_ERR_INVALID; return; }
        const rgnn_linear& L = c.msg.layer[0];
        if (L.weight == nullptr || L.weight_t == nullptr) { set_error("pack: null weight pointer"); rc = RGNN_ERR_INVALID; return; }
        float* wp = const_cast<float*>(L.weight_t);
        float* wc = wp + conv_msg0_proj_floats(d);
        const int ldW = L.in_features, cnp = round_up(d.cn, 8), cep = round_up(d.ce, 8);
        const int ldp = round_up(2 * d.h, 64), ldc = round_up(d.h, 64);
        // columns [0,h): x_target block, [h,2h): x_source block (reference gnn_blocks.py:113 concat order)
        push({L.weight, wp, ldW, 0, d.cn, d.h, cnp, ldp, 0, d.h});
        push({L.weight, wp, ldW, d.cn, d.cn, d.h, cnp, ldp, d.h, ldp - d.h});
        push({L.weight, wc, ldW, 2 * d.cn, d.ce, d.h, cep, ldc, 0, ldc});
        // WP_nat[c][k] (2h x round_up(cn,64)) for the backward: rows [0,h) = W[:, 0:cn], rows [h,2h) = W[:, cn:2cn]
        float* wn = wc + conv_msg0_edge_floats(d);
        const int ldn = round_up(d.cn, 64);
        push_nat({L.weight, wn, ldW, 0, d.h, d.cn, d.h, ldn, 0, ldn});
        push_nat({L.weight, wn + (size_t)d.h * ldn, ldW, d.cn, d.h, d.cn, d.h, ldn, 0, ldn});
        // bias2h = [bias | 0]
        float* wb = wn + conv_msg0_projnat_floats(d);
        const int ldb = (int)conv_msg0_bias_floats(d);
        if (L.bias != nullptr) push_nat({L.bias, wb, d.h, 0, 1, d.h, 1, ldb, 0, ldb});
        else if (cudaMemsetAsync(wb, 0, ldb * sizeof(float), stream) != cudaSuccess) rc = RGNN_ERR_CUDA;
        for (int i = 1; i < c.msg.n; ++i) linear(c.msg.layer[i]);
        stack(c.upd);
        // hi/lo split operands of the tcgen05 message kernel (after the three FFMA-path blocks)
        if (c.msg.n == 2 && c.msg.layer[1].weight != nullptr) {
            int r = mp_tc_pack(c, d, wp + conv_msg0_tc_offset(d), stream);
            if (r == RGNN_OK) r = tc_pack_projection(c, d, stream);
            if (r == RGNN_OK) r = mp_f16_pack(c, d, wp + conv_msg0_f16_offset(d), stream);
            if (r == RGNN_OK) r = conv_proj_f16_pack(c, d, wp + conv_msg0_proj16_offset(d), stream);
            if (r != RGNN_OK) rc = r;
        }
    }
};

bool conv_dims(const rgnn_conv& c, ConvDims* d) {
    if (c.msg.n < 1 || c.upd.n < 1) { set_error("conv block needs msg and upd stacks"); return false; }
    const rgnn_linear& m0 = c.msg.layer[0];
    const rgnn_linear& ml = c.msg.layer[c.msg.n - 1];
    const rgnn_linear& u0 = c.upd.layer[0];
    const rgnn_linear& ul = c.upd.layer[c.upd.n - 1];
    d->cn = ul.out_features;
    d->h = m0.out_features;
    d->ce = m0.in_features - 2 * d->cn;
    if (d->ce <= 0 || ml.out_features != d->cn || u0.in_features != 2 * d->cn) {
        set_error("conv block: unsupported channel plan (msg in %d, msg out %d, upd in %d, upd out %d); only the "
                  "identity-residual form in==out is implemented", m0.in_features, ml.out_features, u0.in_features, ul.out_features);
        return false;
    }
    if ((d->cn % 8) || (d->ce % 8) || (d->h % 32) || d->h > 512 || d->cn > 256 || d->ce > 256) {
        set_error("conv block: widths cn=%d ce=%d h=%d outside the supported range", d->cn, d->ce, d->h);
        return false;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// forward programs (two ping-pong regions of the maximum width)
// ---------------------------------------------------------------------------------------------
int stack_in(const rgnn_stack& s) { return s.layer[0].in_features; }
int stack_out(const rgnn_stack& s) { return s.layer[s.n - 1].out_features; }

// widest activation row of a stack / a conv block (sizes the two ping-pong regions)
static int stack_maxw(const rgnn_stack& s) {
    int w = 64;
    for (int i = 0; i < s.n; ++i) {
        w = w > s.layer[i].in_features ? w : s.layer[i].in_features;
        w = w > s.layer[i].out_features ? w : s.layer[i].out_features;
    }
    return w;
}
static int conv_maxw(const ConvDims& d) { return d.h > 2 * d.cn ? (d.h > d.ce ? d.h : d.ce) : (2 * d.cn > d.ce ? 2 * d.cn : d.ce); }

// Two ping-pong regions.  The reference plan (every row <= 256 wide) runs 64-row tiles; wider channel plans (hidden width 128 / 256
// of BASELINE.json's sweep: rows of up to 512 values) run 32-row tiles with regions of the width they need.
struct FwdBuilder : ProgBuilder {
    int cur, nxt, width;
    explicit FwdBuilder(int n_rows, int maxw = 256) : ProgBuilder(n_rows, maxw > 256 ? 32 : TR_FWD) {
        width = maxw > 256 ? round_up(maxw, 64) : 256;
        cur = region(width);
        nxt = region(width);
    }
    void swap() { int t = cur; cur = nxt; nxt = t; }
    void lin(const rgnn_linear& L) { linear(cur, nxt, L); swap(); }
    void stack(const rgnn_stack& s, int first = 0) { for (int i = first; i < s.n; ++i) lin(s.layer[i]); }
    // hoisted node half of msg.0: P = [x W_t^T + b | x W_s^T], 2h columns, produced and stored in pieces of at most `width` columns
    void proj(const rgnn_conv& c, const ConvDims& d, float* P) {
        const int Cp = round_up(2 * d.h, 64), Kp = round_up(d.cn, 8);
        const float* bias2h = c.msg.layer[0].weight_t + conv_msg0_proj_floats(d) + conv_msg0_edge_floats(d) + conv_msg0_projnat_floats(d);
        const int src = cur;
        for (int c0 = 0; c0 < Cp; c0 += width) {
            const int cw = Cp - c0 < width ? Cp - c0 : width;
            const int valid = 2 * d.h - c0 < cw ? 2 * d.h - c0 : cw;
            gemm(src, nxt, c.msg.layer[0].weight_t + c0, Cp, Kp, Kp, valid, cw, bias2h + c0);
            store_rows(nxt, P, 2 * d.h, valid, c0);
        }
    }
};

int run_stack_fwd(const rgnn_stack& s, const float* x, int n_rows, float* y, cudaStream_t stream, const TcSave* save) {
    RGNN_REQUIRE(s.n >= 1 && s.n <= RGNN_MAX_STACK, "stack with %d layers", s.n);
    // inference (nothing to save for a backward): the fixed-shape fp16-split chain kernel
    // (a training step passes `save`: the same kernel then leaves the layer outputs and sigmas behind for chain64_bwd_f16_kernel)
    if (chain64_supported(s) && (save == nullptr || chain64_bwd_f16_supported(s)))
        return run_chain64(s, x, stack_in(s), nullptr, nullptr, n_rows, y, stream, save);
    if (tc_stack_supported(s)) return tc_run_stack(s, x, nullptr, n_rows, y, stream, save);
    FwdBuilder b(n_rows, stack_maxw(s));
    b.load_rows(b.cur, x, stack_in(s), stack_in(s), 0, round_up(stack_in(s), 8));
    b.stack(s);
    b.store_rows(b.cur, y, stack_out(s), stack_out(s));
    if (!b.ok) return RGNN_ERR_INVALID;
    return launch_program(b.p, stream);
}

int run_proj(const rgnn_conv& c, const float* x, int n_nodes, float* P, cudaStream_t stream) {
    ConvDims d;
    if (!conv_dims(c, &d)) return RGNN_ERR_INVALID;
    RGNN_REQUIRE(c.msg.layer[0].weight_t != nullptr, "conv msg.0 not packed");
    FwdBuilder b(n_nodes, conv_maxw(d));
    b.load_rows(b.cur, x, d.cn, d.cn, 0, round_up(d.cn, 8));
    b.proj(c, d, P);
    if (!b.ok) return RGNN_ERR_INVALID;
    return launch_program(b.p, stream);
}

// Recomputable part of the message function, shared by forward and backward builders:
// z1 = e W_edge^T + P_t[tgt] + P_s[src] (P_t carries the bias); y1 = act(norm(z1)); y2 = msg[1..](y1)
void add_message_layers(ProgBuilder& b, const rgnn_conv& c, const ConvDims& d, const rgnn_graph& g, const float* P,
                        int r_in, int r_mid, int r_out, int slot0, int slot1) {
    const rgnn_linear& m0 = c.msg.layer[0];
    const int Kp = round_up(d.ce, 8), Cp = round_up(d.h, 64);
    b.gemm(r_in, r_mid, m0.weight_t + conv_msg0_proj_floats(d), Cp, Kp, Kp, d.h, Cp, nullptr);   // bias is inside P_t
    Step* s = b.add(OP_ADD_GATHER2, r_mid);
    s->p0 = P; s->p1 = g.tgt; s->p2 = g.src;
    s->i0 = 2 * d.h; s->i1 = d.h; s->i2 = d.h;
    b.norm_act(r_mid, m0, slot0);
    b.linear(r_mid, r_out, c.msg.layer[1], slot1);
}

// message + aggregation: agg[t] = sum_{e: s->t} msg(x_t, x_s, e)
int run_conv_edges(const rgnn_conv& c, const rgnn_graph& g, const float* emb, const float* P, float* agg,
                   cudaStream_t stream, const uint32_t* emb_hl = nullptr) {
    ConvDims d;
    if (!conv_dims(c, &d)) return RGNN_ERR_INVALID;
    RGNN_REQUIRE(c.msg.n == 2, "conv block: msg stack must have 2 ffn_blocks (has %d)", c.msg.n);
    RGNN_CHECK_CUDA(cudaMemsetAsync(agg, 0, (size_t)g.n_nodes * d.cn * sizeof(float), stream));
    RGNN_REQUIRE(c.msg.layer[0].weight_t != nullptr, "conv msg.0 not packed");
    if (g.n_edges > 0 && mp_f16_supported(d)) {
        const float* wpack = c.msg.layer[0].weight_t + conv_msg0_f16_offset(d);
        if (emb_hl != nullptr) return run_conv_edges_f16(c, d, g, emb_hl, P, wpack, agg, stream);
        // stand-alone call with fp32 edge embeddings (block-level API): split them into a stream-ordered temporary
        uint32_t* tmp = nullptr;
        RGNN_CHECK_CUDA(cudaMallocAsync(&tmp, mp_f16_emb_words(g.n_edges) * sizeof(uint32_t), stream));
        int rc = mp_f16_split_emb(emb, g.n_edges, tmp, stream);
        if (rc == RGNN_OK) rc = run_conv_edges_f16(c, d, g, tmp, P, wpack, agg, stream);
        cudaFreeAsync(tmp, stream);
        return rc;
    }
    if (g.n_edges > 0 && mp_tc_supported(d)) {
        const float* wpack = c.msg.layer[0].weight_t + conv_msg0_tc_offset(d);
        return run_conv_edges_tc(c, d, g, emb, P, wpack, agg, stream);
    }
    FwdBuilder b(g.n_edges, conv_maxw(d));
    b.load_rows(b.cur, emb, d.ce, d.ce, 0, round_up(d.ce, 8));
    add_message_layers(b, c, d, g, P, b.cur, b.nxt, b.cur, -1, -1);
    Step* s = b.add(OP_SEGSUM, b.cur);
    s->p0 = agg; s->p1 = g.tgt; s->p2 = g.row_ptr;
    s->i0 = d.cn; s->i1 = d.cn;
    if (!b.ok) return RGNN_ERR_INVALID;
    return launch_program(b.p, stream);
}

// node update: out = x + upd(cat(x, agg)); optionally the next layer's projections
int run_conv_nodes(const rgnn_conv& c, int n_nodes, const float* x, const float* agg, float* out,
                   const rgnn_conv* next, float* P_next, cudaStream_t stream, float* u_save = nullptr, float* sd_save = nullptr) {
    ConvDims d;
    if (!conv_dims(c, &d)) return RGNN_ERR_INVALID;
    // fixed-shape fp16-split kernel; a training step has it leave the update's output and sigma behind when the fp16-split node
    // backward will read them (the interpreter's backward wants the same two arrays, so either forward serves either backward)
    if (conv_nodes_f16_supported(c, d) && ((u_save == nullptr && sd_save == nullptr) || node_bwd_f16_supported(c, d))) {
        if (next != nullptr) {
            ConvDims dn;
            if (!conv_dims(*next, &dn)) return RGNN_ERR_INVALID;
            RGNN_REQUIRE(dn.cn == d.cn && dn.h == d.h && dn.ce == d.ce, "conv blocks with different channel plans");
        }
        return run_conv_nodes_f16(c, d, n_nodes, x, agg, out, next,
                                  next != nullptr ? next->msg.layer[0].weight_t + conv_msg0_proj16_offset(d) : nullptr, P_next, stream,
                                  u_save, sd_save);
    }
    if (tc_stack_supported(c.upd) && c.upd.n == 1 && (next == nullptr || tc_proj_supported(d))) {
        if (next != nullptr) {
            ConvDims dn;
            if (!conv_dims(*next, &dn)) return RGNN_ERR_INVALID;
            RGNN_REQUIRE(dn.cn == d.cn && dn.h == d.h, "conv blocks with different channel plans");
        }
        return tc_run_conv_nodes(c, d, n_nodes, x, agg, out, next, P_next, stream, u_save, sd_save);
    }
    FwdBuilder b(n_nodes, conv_maxw(d));
    b.load_rows(b.cur, x, d.cn, d.cn, 0);
    b.load_rows(b.cur, agg, d.cn, d.cn, d.cn);
    b.stack(c.upd);
    Step* s = b.add(OP_ADD_ROWS, b.cur);
    s->p0 = x; s->i0 = d.cn; s->i1 = d.cn;
    b.store_rows(b.cur, out, d.cn, d.cn);
    if (next != nullptr) {
        ConvDims dn;
        if (!conv_dims(*next, &dn)) return RGNN_ERR_INVALID;
        RGNN_REQUIRE(dn.cn == d.cn, "conv blocks with different node widths");
        b.proj(*next, dn, P_next);
    }
    if (!b.ok) return RGNN_ERR_INVALID;
    return launch_program(b.p, stream);
}

// ---------------------------------------------------------------------------------------------
// detector workspace
// ---------------------------------------------------------------------------------------------
int plan_detector(const rgnn_detector& net, const rgnn_graph& g, int training, void* base, DetPlan* pl) {
    RGNN_REQUIRE(net.n_conv >= 1 && net.n_conv <= RGNN_MAX_CONV, "n_conv = %d", net.n_conv);
    ConvDims d;
    if (!conv_dims(net.conv[0], &d)) return RGNN_ERR_INVALID;
    for (int l = 1; l < net.n_conv; ++l) {
        ConvDims dl;
        if (!conv_dims(net.conv[l], &dl)) return RGNN_ERR_INVALID;
        RGNN_REQUIRE(dl.cn == d.cn && dl.ce == d.ce && dl.h == d.h, "conv blocks must share one channel plan");
    }
    RGNN_REQUIRE(stack_out(net.node_enc) == d.cn && stack_out(net.edge_enc) == d.ce, "encoder widths do not match conv");
    pl->d = d;
    pl->training = training != 0;
    pl->link_w = stack_out(net.link_node);
    pl->cls_w = stack_out(net.class_node);
    RGNN_REQUIRE(stack_in(net.head_link) == pl->link_w && stack_in(net.head_class) == pl->cls_w &&
                 (pl->link_w % 4) == 0 && (pl->cls_w % 4) == 0, "head widths inconsistent");
    char* p = static_cast<char*>(base);
    size_t off = 0;
    auto take = [&](size_t n_floats) {
        float* r = reinterpret_cast<float*>(p + off);
        off += align256(n_floats * sizeof(float));
        return r;
    };
    const size_t N = (size_t)g.n_nodes, E = (size_t)g.n_edges;
    const int L = net.n_conv;
    if (training) {
        for (int l = 0; l <= L; ++l) pl->x[l] = take(N * d.cn);
        for (int l = 0; l < L; ++l) pl->P[l] = take(N * 2 * d.h);
        for (int l = 0; l < L; ++l) pl->agg[l] = take(N * d.cn);
    } else {
        float* xa = take(N * d.cn);
        float* xb = take(N * d.cn);
        float* P = take(N * 2 * d.h);
        float* agg = take(N * d.cn);
        for (int l = 0; l <= L; ++l) pl->x[l] = (l & 1) ? xb : xa;
        for (int l = 0; l < L; ++l) { pl->P[l] = P; pl->agg[l] = agg; }
    }
    pl->emb = take(E * d.ce);
    pl->emb_hl = (E > 0 && mp_f16_supported(d)) ? reinterpret_cast<uint32_t*>(take(mp_f16_emb_words(g.n_edges))) : nullptr;
    pl->hlink = take(N * pl->link_w);
    pl->gcls = take(N * pl->cls_w);
    pl->enc_tc_bwd = pl->link_tc_bwd = pl->conv_tc_bwd = false;
    for (int i = 0; i < 5; ++i) { pl->node_tc_bwd[i] = false; memset(&pl->node_save[i], 0, sizeof(TcSave)); }
    for (int l = 0; l < RGNN_MAX_CONV; ++l) pl->u_save[l] = pl->usd_save[l] = nullptr;
    memset(&pl->enc_save, 0, sizeof(TcSave));
    memset(&pl->link_save, 0, sizeof(TcSave));
    pl->cscr = nullptr;
    if (training) plan_detector_bwd(net, g, take, pl);
    pl->bytes = off;
    return RGNN_OK;
}

// object_classification head on the per-node stem output pl.gcls: per-cluster max (gnn_blocks.py:384-386) -> FFN_TaskSpecificHead
static int run_obj_head(const rgnn_detector& net, const rgnn_graph& g, const DetPlan& pl, float* obj_cls, cudaStream_t stream) {
    int rc;
    if (g.n_clusters > 0 && tc_stack_supported(net.head_class)) {
        if ((rc = tc_run_segmax_stack(net.head_class, pl.gcls, pl.cls_w, g.cl_ptr, g.cl_members, g.n_clusters, obj_cls, stream))) return rc;
    } else if (g.n_clusters > 0) {
        FwdBuilder b(g.n_clusters, stack_maxw(net.head_class));
        Step* s = b.add(OP_LOAD_SEGMAX, b.cur);
        s->p0 = pl.gcls; s->p1 = g.cl_ptr; s->p2 = g.cl_members; s->i0 = pl.cls_w; s->i1 = pl.cls_w;
        b.stack(net.head_class);
        b.store_rows(b.cur, obj_cls, stack_out(net.head_class), stack_out(net.head_class));
        if (!b.ok) return RGNN_ERR_INVALID;
        if ((rc = launch_program(b.p, stream))) return rc;
    }
    return RGNN_OK;
}

int detector_fwd(const rgnn_detector& net, const rgnn_graph& g, const float* node_features,
                 const float* edge_features, float* node_cls, float* node_off, float* link_cls, float* obj_cls,
                 const DetPlan& pl, cudaStream_t stream) {
    const ConvDims& d = pl.d;
    const int N = g.n_nodes, E = g.n_edges, L = net.n_conv;
    int rc;
    if (tc_stack_supported(net.node_enc) && tc_proj_supported(d)) {
        if ((rc = tc_run_node_encoder(net.node_enc, net.conv[0], d, node_features, N, pl.x[0], pl.P[0], stream,
                                      pl.node_tc_bwd[0] ? &pl.node_save[0] : nullptr)))
            return rc;
    } else {   // node encoder (+ first layer's projections)
        FwdBuilder b(N, stack_maxw(net.node_enc) > conv_maxw(d) ? stack_maxw(net.node_enc) : conv_maxw(d));
        const int in = stack_in(net.node_enc);
        b.load_rows(b.cur, node_features, in, in, 0, round_up(in, 8));
        b.stack(net.node_enc);
        b.store_rows(b.cur, pl.x[0], d.cn, d.cn);
        b.proj(net.conv[0], d, pl.P[0]);
        if (!b.ok) return RGNN_ERR_INVALID;
        if ((rc = launch_program(b.p, stream))) return rc;
    }
    bool emb_split_done = false;
    if (E > 0 && pl.emb_hl != nullptr && edge_enc_f16_supported(net.edge_enc)) {
        // the fixed-shape encoder writes the pre-split rows of the message kernel directly; the fp32 embedding is written as well
        // only when a backward will read it (a training step whose backward recomputes the encoder on the CUDA cores)
        // (a training step with the chain backward also gets the layer outputs and sigmas that backward reads: enc_save)
        if ((rc = run_edge_enc_f16(net.edge_enc, edge_features, g.perm, E, pl.emb_hl, pl.training ? pl.emb : nullptr, stream,
                                   pl.enc_tc_bwd ? &pl.enc_save : nullptr)))
            return rc;
        emb_split_done = true;
    } else if (E > 0 && tc_stack_supported(net.edge_enc)) {
        if ((rc = tc_run_stack(net.edge_enc, edge_features, g.perm, E, pl.emb, stream, pl.enc_tc_bwd ? &pl.enc_save : nullptr))) return rc;
    } else if (E > 0) {   // edge encoder, rows gathered into target-major order
        FwdBuilder b(E, stack_maxw(net.edge_enc));
        const int in = stack_in(net.edge_enc);
        b.load_rows(b.cur, edge_features, in, in, 0, round_up(in, 8), g.perm);
        b.stack(net.edge_enc);
        b.store_rows(b.cur, pl.emb, d.ce, d.ce);
        if (!b.ok) return RGNN_ERR_INVALID;
        if ((rc = launch_program(b.p, stream))) return rc;
    }
    if (pl.emb_hl != nullptr && !emb_split_done && (rc = mp_f16_split_emb(pl.emb, E, pl.emb_hl, stream))) return rc;
    for (int l = 0; l < L; ++l) {
        if ((rc = run_conv_edges(net.conv[l], g, pl.emb, pl.P[l], pl.agg[l], stream, pl.emb_hl))) return rc;
        if ((rc = run_conv_nodes(net.conv[l], N, pl.x[l], pl.agg[l], pl.x[l + 1], l + 1 < L ? &net.conv[l + 1] : nullptr,
                                 l + 1 < L ? pl.P[l + 1] : nullptr, stream, pl.conv_tc_bwd ? pl.u_save[l] : nullptr,
                                 pl.conv_tc_bwd ? pl.usd_save[l] : nullptr)))
            return rc;
    }
    const float* xL = pl.x[L];
    auto sv = [&](int i) { return pl.node_tc_bwd[i] ? &pl.node_save[i] : nullptr; };
    if ((rc = run_stack_fwd(net.head_node, xL, N, node_cls, stream, sv(1)))) return rc;
    if ((rc = run_stack_fwd(net.head_offset, xL, N, node_off, stream, sv(2)))) return rc;
    if ((rc = run_stack_fwd(net.link_node, xL, N, pl.hlink, stream, sv(3)))) return rc;
    if (g.n_und > 0 && chain64_supported(net.head_link) && (!pl.link_tc_bwd || chain64_bwd_f16_supported(net.head_link))) {
        if ((rc = run_chain64(net.head_link, pl.hlink, pl.link_w, g.und_a, g.und_b, g.n_und, link_cls, stream,
                              pl.link_tc_bwd ? &pl.link_save : nullptr)))
            return rc;
    } else if (g.n_und > 0 && tc_stack_supported(net.head_link)) {
        if ((rc = tc_run_pairsum_stack(net.head_link, pl.hlink, pl.link_w, g.und_a, g.und_b, g.n_und, link_cls, stream,
                                       pl.link_tc_bwd ? &pl.link_save : nullptr)))
            return rc;
    } else if (g.n_und > 0) {
        FwdBuilder b(g.n_und, stack_maxw(net.head_link));
        Step* s = b.add(OP_LOAD_PAIRSUM, b.cur);
        s->p0 = pl.hlink; s->p1 = g.und_a; s->p2 = g.und_b; s->i0 = pl.link_w; s->i1 = pl.link_w;
        b.stack(net.head_link);
        b.store_rows(b.cur, link_cls, stack_out(net.head_link), stack_out(net.head_link));
        if (!b.ok) return RGNN_ERR_INVALID;
        if ((rc = launch_program(b.p, stream))) return rc;
    }
    if ((rc = run_stack_fwd(net.class_node, xL, N, pl.gcls, stream, sv(4)))) return rc;
    if ((rc = run_obj_head(net, g, pl, obj_cls, stream))) return rc;
    return RGNN_OK;
}

}  // namespace rgnn

// ---------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------
using namespace rgnn;

extern "C" size_t rgnn_packed_weight_floats(int in_features, int out_features) {
    return (size_t)round_up(in_features, 8) * round_up(out_features, 64) + tc_linear_pack_floats(in_features, out_features);
}

extern "C" int rgnn_pack_linear(const float* weight, int in_features, int out_features, float* weight_t, void* stream) {
    rgnn_linear L;
    memset(&L, 0, sizeof(L));
    L.weight = weight; L.weight_t = weight_t; L.in_features = in_features; L.out_features = out_features;
    Packer pk(static_cast<cudaStream_t>(stream));
    pk.linear(L);
    pk.flush();
    return pk.rc;
}

extern "C" int rgnn_pack_stack(const rgnn_stack* s, void* stream) {
    Packer pk(static_cast<cudaStream_t>(stream));
    pk.stack(*s);
    pk.flush();
    return pk.rc;
}

extern "C" int rgnn_pack_conv(const rgnn_conv* c, void* stream) {
    Packer pk(static_cast<cudaStream_t>(stream));
    pk.conv(*c);
    pk.flush();
    return pk.rc;
}

extern "C" int rgnn_pack_detector(const rgnn_detector* net, void* stream) {
    Packer pk(static_cast<cudaStream_t>(stream));
    if (g_pack_batch) packq_begin(static_cast<cudaStream_t>(stream));        // ~214 per-image launches -> a handful (rgnn_pack.cuh)
    pk.stack(net->node_enc);
    pk.stack(net->edge_enc);
    // edge_enc_f16_pack REWRITES layer 1's fp16 image in four K blocks: what the generic packer queued for it must be launched first
    if (pk.rc == RGNN_OK) pk.rc = packq_flush();
    if (pk.rc == RGNN_OK) pk.rc = edge_enc_f16_pack(net->edge_enc, static_cast<cudaStream_t>(stream));
    for (int l = 0; l < net->n_conv; ++l) pk.conv(net->conv[l]);
    pk.stack(net->head_node);
    pk.stack(net->head_offset);
    pk.stack(net->link_node);
    pk.stack(net->head_link);
    pk.stack(net->class_node);
    pk.stack(net->head_class);
    pk.flush();
    {
        const int rq = packq_end();
        if (pk.rc == RGNN_OK) pk.rc = rq;
    }
    return pk.rc;
}

extern "C" int rgnn_ffn_stack_fwd(const rgnn_stack* stack, const float* x, int n_rows, float* y, void* stream) {
    return run_stack_fwd(*stack, x, n_rows, y, static_cast<cudaStream_t>(stream));
}

extern "C" int rgnn_conv_block_fwd(const rgnn_conv* blk, const rgnn_graph* g, const float* x, const float* e,
                                   float* out, float* agg, float* proj, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc;
    if ((rc = run_proj(*blk, x, g->n_nodes, proj, s))) return rc;
    if ((rc = run_conv_edges(*blk, *g, e, proj, agg, s))) return rc;
    return run_conv_nodes(*blk, g->n_nodes, x, agg, out, nullptr, nullptr, s);
}

extern "C" int rgnn_conv_edges_fwd(const rgnn_conv* blk, const rgnn_graph* g, const float* e, const float* proj, float* agg, void* stream) {
    return run_conv_edges(*blk, *g, e, proj, agg, static_cast<cudaStream_t>(stream));
}

extern "C" size_t rgnn_split_edge_embedding_words(int n_edges) { return mp_f16_emb_words(n_edges); }

extern "C" int rgnn_split_edge_embedding(const float* e, int n_edges, void* e_split, void* stream) {
    return mp_f16_split_emb(e, n_edges, static_cast<uint32_t*>(e_split), static_cast<cudaStream_t>(stream));
}

extern "C" int rgnn_conv_edges_f16_fwd(const rgnn_conv* blk, const rgnn_graph* g, const void* e_split, const float* proj,
                                       float* agg, void* stream) {
    ConvDims d;
    if (!conv_dims(*blk, &d)) return RGNN_ERR_INVALID;
    RGNN_REQUIRE(mp_f16_supported(d), "conv_edges_f16: channel plan %d / %d / %d is not instantiated (64 / 64 / 128)", d.cn, d.ce, d.h);
    return run_conv_edges(*blk, *g, nullptr, proj, agg, static_cast<cudaStream_t>(stream), static_cast<const uint32_t*>(e_split));
}

extern "C" int rgnn_edge_encoder_f16_fwd(const rgnn_stack* enc, const float* edge_features, const int32_t* perm, int n_edges,
                                         void* e_split, float* e_f32, void* stream) {
    RGNN_REQUIRE(edge_enc_f16_supported(*enc), "edge_encoder_f16: only the reference plan (<= 7 -> 256 -> 128 -> 128 -> 64) is instantiated");
    return run_edge_enc_f16(*enc, edge_features, perm, n_edges, static_cast<uint32_t*>(e_split), e_f32, static_cast<cudaStream_t>(stream));
}

extern "C" int rgnn_conv_layer_f16_fwd(const rgnn_conv* blk, const rgnn_conv* next, const rgnn_graph* g, const float* x,
                                       const void* e_split, const float* proj, float* out, float* agg, float* proj_next, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ConvDims d;
    if (!conv_dims(*blk, &d)) return RGNN_ERR_INVALID;
    RGNN_REQUIRE(mp_f16_supported(d), "conv_layer_f16: channel plan %d / %d / %d is not instantiated (64 / 64 / 128)", d.cn, d.ce, d.h);
    int rc = run_conv_edges(*blk, *g, nullptr, proj, agg, s, static_cast<const uint32_t*>(e_split));
    if (rc) return rc;
    return run_conv_nodes(*blk, g->n_nodes, x, agg, out, next, next != nullptr ? proj_next : nullptr, s);
}

extern "C" size_t rgnn_conv_msg_bwd_workspace_bytes(const rgnn_conv* blk, const rgnn_graph* g) {
    ConvDims d;
    if (!conv_dims(*blk, &d)) return 0;
    return align256(mp_bwd_tc_scratch_floats(d, g->n_edges) * sizeof(float)) + align256(src_index_ints(g->n_nodes, g->n_edges) * sizeof(int));
}

extern "C" int rgnn_conv_msg_bwd(const rgnn_conv* blk, const rgnn_graph* g, const float* e, const float* proj, const float* dagg,
                                 float* dproj, float* de, void* workspace, size_t workspace_bytes, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ConvDims d;
    if (!conv_dims(*blk, &d)) return RGNN_ERR_INVALID;
    RGNN_REQUIRE(mp_bwd_tc_supported(d), "conv_msg_bwd: channel plan %d / %d / %d is not instantiated (64 / 64 / 128)", d.cn, d.ce, d.h);
    RGNN_REQUIRE(workspace_bytes >= rgnn_conv_msg_bwd_workspace_bytes(blk, g), "conv_msg_bwd: workspace too small");
    float* scratch = static_cast<float*>(workspace);
    int* sidx = reinterpret_cast<int*>(static_cast<char*>(workspace) + align256(mp_bwd_tc_scratch_floats(d, g->n_edges) * sizeof(float)));
    const int* sptr = nullptr;
    const int* slist = nullptr;
    int rc = build_src_index(*g, sidx, &sptr, &slist, s);
    if (rc) return rc;
    if (g->n_edges > 0 && mp_bwd_f16_supported(d)) {       // fused fp16-split kernel on pre-split rows (stream-ordered temporary)
        uint32_t* tmp = nullptr;
        RGNN_CHECK_CUDA(cudaMallocAsync(&tmp, mp_f16_emb_words(g->n_edges) * sizeof(uint32_t), s));
        rc = mp_f16_split_emb(e, g->n_edges, tmp, s);
        if (rc == RGNN_OK) rc = run_conv_edges_bwd_f16(*blk, d, *g, tmp, proj, dagg, dproj, de, true, scratch, sptr, slist, s);
        cudaFreeAsync(tmp, s);
        return rc;
    }
    return run_conv_edges_bwd_tc(*blk, d, *g, e, proj, dagg, dproj, de, true, scratch, sptr, slist, s);
}

extern "C" int rgnn_conv_msg_f16_bwd(const rgnn_conv* blk, const rgnn_graph* g, const void* e_split, const float* proj, const float* dagg,
                                     float* dproj, float* de, void* workspace, size_t workspace_bytes, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ConvDims d;
    if (!conv_dims(*blk, &d)) return RGNN_ERR_INVALID;
    RGNN_REQUIRE(mp_bwd_f16_supported(d), "conv_msg_f16_bwd: channel plan %d / %d / %d is not instantiated (64 / 64 / 128)", d.cn, d.ce, d.h);
    RGNN_REQUIRE(workspace_bytes >= rgnn_conv_msg_bwd_workspace_bytes(blk, g), "conv_msg_f16_bwd: workspace too small");
    float* scratch = static_cast<float*>(workspace);
    int* sidx = reinterpret_cast<int*>(static_cast<char*>(workspace) + align256(mp_bwd_tc_scratch_floats(d, g->n_edges) * sizeof(float)));
    const int* sptr = nullptr;
    const int* slist = nullptr;
    int rc = build_src_index(*g, sidx, &sptr, &slist, s);
    if (rc) return rc;
    if (g->n_edges == 0) return cudaMemsetAsync(dproj, 0, (size_t)g->n_nodes * 2 * d.h * sizeof(float), s) == cudaSuccess ? RGNN_OK : RGNN_ERR_CUDA;
    return run_conv_edges_bwd_f16(*blk, d, *g, static_cast<const uint32_t*>(e_split), proj, dagg, dproj, de, true, scratch, sptr, slist, s);
}

extern "C" size_t rgnn_detector_workspace_bytes(const rgnn_detector* net, const rgnn_graph* g, int training) {
    DetPlan pl;
    if (plan_detector(*net, *g, training, nullptr, &pl) != RGNN_OK) return 0;
    return pl.bytes;
}

extern "C" int rgnn_detector_fwd(const rgnn_detector* net, const rgnn_graph* g, const float* node_features,
                                 const float* edge_features, float* node_cls, float* node_off, float* link_cls,
                                 float* obj_cls, void* workspace, size_t workspace_bytes, int training, void* stream) {
    DetPlan pl;
    int rc = plan_detector(*net, *g, training, workspace, &pl);
    if (rc) return rc;
    if (pl.bytes > workspace_bytes) {
        set_error("detector workspace too small: need %zu bytes, got %zu", pl.bytes, workspace_bytes);
        return RGNN_ERR_WORKSPACE;
    }
    return detector_fwd(*net, *g, node_features, edge_features, node_cls, node_off, link_cls, obj_cls, pl,
                        static_cast<cudaStream_t>(stream));
}

extern "C" int rgnn_detector_obj_head(const rgnn_detector* net, const rgnn_graph* g, float* obj_cls, void* workspace,
                                      size_t workspace_bytes, int training, void* stream) {
    DetPlan pl;
    int rc = plan_detector(*net, *g, training, workspace, &pl);
    if (rc) return rc;
    if (pl.bytes > workspace_bytes) {
        set_error("detector workspace too small: need %zu bytes, got %zu", pl.bytes, workspace_bytes);
        return RGNN_ERR_WORKSPACE;
    }
    return run_obj_head(*net, *g, pl, obj_cls, static_cast<cudaStream_t>(stream));
}

extern "C" size_t rgnn_packed_conv_msg0_floats(int node_channels, int edge_channels, int hidden) {
    ConvDims d{node_channels, edge_channels, hidden};
    return conv_msg0_proj16_offset(d) + conv_proj_f16_floats(d);
}
