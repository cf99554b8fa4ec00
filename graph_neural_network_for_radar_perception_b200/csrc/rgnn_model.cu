// Host side of the model entry points: weight packing and the forward tile programs of the detector.
#include "rgnn_model.h"

namespace rgnn {

// ---------------------------------------------------------------------------------------------
// weight packing: k-major, zero padded copies for the tile GEMM
// ---------------------------------------------------------------------------------------------
struct PackEntry {
    const float* W;   // (C, ldW) row-major
    float* dst;       // (Kpad, ldd)
    int ldW, koff, K, C, Kpad, ldd, c0, cw;
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
        if (k < e.K && c < e.C) v = e.W[(size_t)c * e.ldW + e.koff + k];
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
        tab.e[tab.n++] = e;
        if (tab.n == PACK_BATCH) flush();
    }
    void linear(const rgnn_linear& L) {
        if (L.weight == nullptr || L.weight_t == nullptr) { set_error("pack: null weight pointer"); rc = RGNN_ERR_INVALID; return; }
        const int Kpad = round_up(L.in_features, 8), Cpad = round_up(L.out_features, 64);
        push({L.weight, const_cast<float*>(L.weight_t), L.in_features, 0, L.in_features, L.out_features, Kpad, Cpad, 0, Cpad});
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
        for (int i = 1; i < c.msg.n; ++i) linear(c.msg.layer[i]);
        stack(c.upd);
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
    if ((d->cn % 8) || (d->ce % 8) || (d->h % 32) || 2 * d->h > 256 || 2 * d->cn > 256) {
        set_error("conv block: widths cn=%d ce=%d h=%d outside the supported range", d->cn, d->ce, d->h);
        return false;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// forward programs
// ---------------------------------------------------------------------------------------------
static int stack_in(const rgnn_stack& s) { return s.layer[0].in_features; }
static int stack_out(const rgnn_stack& s) { return s.layer[s.n - 1].out_features; }

int run_stack_fwd(const rgnn_stack& s, const float* x, int n_rows, float* y, cudaStream_t stream) {
    RGNN_REQUIRE(s.n >= 1 && s.n <= RGNN_MAX_STACK, "stack with %d layers", s.n);
    ProgBuilder b(n_rows);
    b.load_rows(x, stack_in(s), stack_in(s), 0, round_up(stack_in(s), 8));
    b.stack(s);
    b.store_rows(y, stack_out(s), stack_out(s));
    if (!b.ok) return RGNN_ERR_INVALID;
    return launch_fwd(b.p, stream);
}

static void add_proj(ProgBuilder& b, const rgnn_conv& c, const ConvDims& d, float* P) {
    b.add(OP_LINEAR, round_up(d.cn, 8), 2 * d.h, round_up(2 * d.h, 64), 0, c.msg.layer[0].weight_t, nullptr);
    b.store_rows(P, 2 * d.h, 2 * d.h);
}

int run_proj(const rgnn_conv& c, const float* x, int n_nodes, float* P, cudaStream_t stream) {
    ConvDims d;
    if (!conv_dims(c, &d)) return RGNN_ERR_INVALID;
    ProgBuilder b(n_nodes);
    b.load_rows(x, d.cn, d.cn);
    add_proj(b, c, d, P);
    if (!b.ok) return RGNN_ERR_INVALID;
    return launch_fwd(b.p, stream);
}

// message + aggregation: agg[t] = sum_{e: s->t} msg(x_t, x_s, e)
int run_conv_edges(const rgnn_conv& c, const rgnn_graph& g, const float* emb, const float* P, float* agg,
                   cudaStream_t stream) {
    ConvDims d;
    if (!conv_dims(c, &d)) return RGNN_ERR_INVALID;
    RGNN_CHECK_CUDA(cudaMemsetAsync(agg, 0, (size_t)g.n_nodes * d.cn * sizeof(float), stream));
    const rgnn_linear& m0 = c.msg.layer[0];
    RGNN_REQUIRE(m0.weight_t != nullptr, "conv msg.0 not packed");
    ProgBuilder b(g.n_edges);
    b.load_rows(emb, d.ce, d.ce);
    b.add(OP_LINEAR, round_up(d.ce, 8), d.h, round_up(d.h, 64), 0, m0.weight_t + conv_msg0_proj_floats(d), m0.bias);
    b.add(OP_ADD_GATHER2, 2 * d.h, d.h, d.h, 0, P, g.tgt, g.src);
    if (m0.norm_scale != nullptr || m0.activation)
        b.add(OP_NORM_ACT, d.h, m0.activation, 0, 0, m0.norm_scale, m0.norm_shift);
    b.stack(c.msg, 1);
    b.add(OP_SEGSUM, d.cn, d.cn, 0, 0, agg, g.tgt, g.row_ptr);
    if (!b.ok) return RGNN_ERR_INVALID;
    return launch_fwd(b.p, stream);
}

// node update: out = x + upd(cat(x, agg)); optionally the next layer's projections
int run_conv_nodes(const rgnn_conv& c, int n_nodes, const float* x, const float* agg, float* out,
                   const rgnn_conv* next, float* P_next, cudaStream_t stream) {
    ConvDims d;
    if (!conv_dims(c, &d)) return RGNN_ERR_INVALID;
    ProgBuilder b(n_nodes);
    b.load_rows(x, d.cn, d.cn, 0);
    b.load_rows(agg, d.cn, d.cn, d.cn);
    b.stack(c.upd);
    b.add(OP_ADD_ROWS, d.cn, d.cn, 0, 0, x);
    b.store_rows(out, d.cn, d.cn);
    if (next != nullptr) {
        ConvDims dn;
        if (!conv_dims(*next, &dn)) return RGNN_ERR_INVALID;
        RGNN_REQUIRE(dn.cn == d.cn, "conv blocks with different node widths");
        add_proj(b, *next, dn, P_next);
    }
    if (!b.ok) return RGNN_ERR_INVALID;
    return launch_fwd(b.p, stream);
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
    pl->hlink = take(N * pl->link_w);
    pl->gcls = take(N * pl->cls_w);
    if (training) plan_detector_bwd(net, g, take, pl);
    pl->bytes = off;
    return RGNN_OK;
}

int detector_fwd(const rgnn_detector& net, const rgnn_graph& g, const float* node_features,
                 const float* edge_features, float* node_cls, float* node_off, float* link_cls, float* obj_cls,
                 const DetPlan& pl, cudaStream_t stream) {
    const ConvDims& d = pl.d;
    const int N = g.n_nodes, E = g.n_edges, L = net.n_conv;
    int rc;
    {   // node encoder (+ first layer's projections)
        ProgBuilder b(N);
        const int in = stack_in(net.node_enc);
        b.load_rows(node_features, in, in, 0, round_up(in, 8));
        b.stack(net.node_enc);
        b.store_rows(pl.x[0], d.cn, d.cn);
        add_proj(b, net.conv[0], d, pl.P[0]);
        if (!b.ok) return RGNN_ERR_INVALID;
        if ((rc = launch_fwd(b.p, stream))) return rc;
    }
    {   // edge encoder, rows gathered into target-major order
        ProgBuilder b(E);
        const int in = stack_in(net.edge_enc);
        b.load_rows(edge_features, in, in, 0, round_up(in, 8), g.perm);
        b.stack(net.edge_enc);
        b.store_rows(pl.emb, d.ce, d.ce);
        if (!b.ok) return RGNN_ERR_INVALID;
        if ((rc = launch_fwd(b.p, stream))) return rc;
    }
    for (int l = 0; l < L; ++l) {
        if ((rc = run_conv_edges(net.conv[l], g, pl.emb, pl.P[l], pl.agg[l], stream))) return rc;
        if ((rc = run_conv_nodes(net.conv[l], N, pl.x[l], pl.agg[l], pl.x[l + 1], l + 1 < L ? &net.conv[l + 1] : nullptr,
                                 l + 1 < L ? pl.P[l + 1] : nullptr, stream)))
            return rc;
    }
    const float* xL = pl.x[L];
    if ((rc = run_stack_fwd(net.head_node, xL, N, node_cls, stream))) return rc;
    if ((rc = run_stack_fwd(net.head_offset, xL, N, node_off, stream))) return rc;
    if ((rc = run_stack_fwd(net.link_node, xL, N, pl.hlink, stream))) return rc;
    if (g.n_und > 0) {
        ProgBuilder b(g.n_und);
        b.add(OP_LOAD_PAIRSUM, pl.link_w, pl.link_w, 0, 0, pl.hlink, g.und_a, g.und_b);
        b.stack(net.head_link);
        b.store_rows(link_cls, stack_out(net.head_link), stack_out(net.head_link));
        if (!b.ok) return RGNN_ERR_INVALID;
        if ((rc = launch_fwd(b.p, stream))) return rc;
    }
    if ((rc = run_stack_fwd(net.class_node, xL, N, pl.gcls, stream))) return rc;
    if (g.n_clusters > 0) {
        ProgBuilder b(g.n_clusters);
        b.add(OP_LOAD_SEGMAX, pl.cls_w, pl.cls_w, 0, 0, pl.gcls, g.cl_ptr, g.cl_members);
        b.stack(net.head_class);
        b.store_rows(obj_cls, stack_out(net.head_class), stack_out(net.head_class));
        if (!b.ok) return RGNN_ERR_INVALID;
        if ((rc = launch_fwd(b.p, stream))) return rc;
    }
    return RGNN_OK;
}

}  // namespace rgnn

// ---------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------
using namespace rgnn;

extern "C" size_t rgnn_packed_weight_floats(int in_features, int out_features) {
    return (size_t)round_up(in_features, 8) * round_up(out_features, 64);
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
    pk.stack(net->node_enc);
    pk.stack(net->edge_enc);
    for (int l = 0; l < net->n_conv; ++l) pk.conv(net->conv[l]);
    pk.stack(net->head_node);
    pk.stack(net->head_offset);
    pk.stack(net->link_node);
    pk.stack(net->head_link);
    pk.stack(net->class_node);
    pk.stack(net->head_class);
    pk.flush();
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

extern "C" size_t rgnn_packed_conv_msg0_floats(int node_channels, int edge_channels, int hidden) {
    ConvDims d{node_channels, edge_channels, hidden};
    return conv_msg0_proj_floats(d) + conv_msg0_edge_floats(d);
}
