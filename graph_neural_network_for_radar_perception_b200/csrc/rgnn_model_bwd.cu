// Backward of the detector as tile programs.  Every program first recomputes the forward activations of its
// tile into shared memory (nothing per-edge is ever read back from HBM), then walks the layers in reverse:
//   OP_ACTNORM_BWD (LeakyReLU + channel norm)  ->  OP_WGRAD (dW, db by RED into the gradient buffers)
//   ->  OP_LINEAR with the natural-layout weight as operand (dX = dZ W).
// Replaces torch autograd over gnn_blocks.py / gnn_detector.py (reference gnn/training.py:81).
#include <vector>

#include "rgnn_model.h"

namespace rgnn {

struct BwdBuilder : ProgBuilder {
    int wcur, wnxt;
    // tr = 64 doubles the register blocking of the tile GEMM (4 rows per thread instead of 2); it is used where
    // the activations of 64 rows still fit in shared memory (narrow stacks, the message function)
    explicit BwdBuilder(int n_rows, int tr = TR_BWD) : ProgBuilder(n_rows, tr), wcur(-1), wnxt(-1) {}
    void work_regions(int width = 256) {
        wcur = region(width);
        wnxt = region(width);
    }
    void swap() { int t = wcur; wcur = wnxt; wnxt = t; }

    // forward recompute of layers [first, n): outputs go to fresh regions, one sigma slot per normalised layer
    void chain_fwd(const rgnn_stack& s, int first, int r_in, std::vector<int>& regs, std::vector<int>& slots) {
        regs.assign(s.n, -1);
        slots.assign(s.n, -1);
        int prev = r_in;
        for (int i = first; i < s.n; ++i) {
            const rgnn_linear& L = s.layer[i];
            regs[i] = region(round_up(L.out_features, 64));
            if (L.norm_scale != nullptr) slots[i] = sigma_slot();
            linear(prev, regs[i], L, slots[i]);
            prev = regs[i];
        }
    }

    void actnorm_bwd(const rgnn_linear& L, int r_y, int slot) {
        if (L.norm_scale == nullptr && !L.activation) return;
        if (L.out_features > 256) {
            set_error("backward: layers wider than 256 channels (%d) are outside the training envelope", L.out_features);
            ok = false;
            return;
        }
        Step* st = add(OP_ACTNORM_BWD, wcur, r_y);
        st->i0 = L.out_features; st->i1 = L.activation; st->i2 = slot < 0 ? 0 : slot;
        st->p0 = L.norm_scale; st->p1 = L.norm_shift; st->p2 = L.grad_norm_scale; st->p3 = L.grad_norm_shift;
    }

    void wgrad(int r_dz, int r_x, int C, int K, float* dW, int ldW, int wcol, int zcol, float* db) {
        if (dW == nullptr && db == nullptr) return;
        Step* st = add(OP_WGRAD, r_dz, r_x);
        st->i0 = C; st->i1 = K; st->i2 = ldW; st->i3 = wcol; st->i4 = zcol;
        st->p0 = dW; st->p1 = db;
    }

    // dX = dZ * W with W in its natural (out, in) layout: rows = reduction index, in_features columns (not padded)
    bool dgrad(const float* W, int ldw, int n_out_rows, int n_in_cols) {
        if (n_in_cols % 8 != 0) {
            set_error("backward: input width %d of a differentiated Linear must be a multiple of 8", n_in_cols);
            ok = false;
            return false;
        }
        gemm(wcur, wnxt, W, ldw, round_up(n_out_rows, 8), n_out_rows, n_in_cols, round_up(n_in_cols, 64), nullptr, n_in_cols);
        swap();
        return true;
    }

    // reverse pass over layers [first, n); dY is in wcur on entry, dX (if wanted) in wcur on exit
    void chain_bwd(const rgnn_stack& s, int first, int r_in, const std::vector<int>& regs, const std::vector<int>& slots,
                   bool need_dx_first) {
        for (int i = s.n - 1; i >= first; --i) {
            const rgnn_linear& L = s.layer[i];
            actnorm_bwd(L, regs[i], slots[i]);
            const int r_x = (i == first) ? r_in : regs[i - 1];
            wgrad(wcur, r_x, L.out_features, L.in_features, L.grad_weight, L.in_features, 0, 0, L.grad_bias);
            if (i > first || need_dx_first) dgrad(L.weight, L.in_features, L.out_features, L.in_features);
        }
    }

    void zero_region(int r, int w) { load_rows(r, nullptr, 4, 0, 0, w); }
    void add_region(int ra, int rb, int w, int ca = 0, int cb = 0) {
        Step* st = add(OP_ADD_REGION, ra, rb);
        st->i1 = w; st->i2 = ca; st->i3 = cb;
    }
};

// generic ffn-stack backward; x rows optionally gathered through ridx
static int stack_bwd(const rgnn_stack& s, const float* x, const int* ridx, const float* grad_y, int n_rows,
                     float* grad_x, bool accumulate_gx, cudaStream_t stream) {
    RGNN_REQUIRE(s.n >= 1 && s.n <= RGNN_MAX_STACK, "stack with %d layers", s.n);
    BwdBuilder b(n_rows);
    const int in = stack_in(s), out = stack_out(s);
    int wmax = round_up(in, 64);
    for (int i = 0; i < s.n; ++i) wmax = wmax > round_up(s.layer[i].out_features, 64) ? wmax : round_up(s.layer[i].out_features, 64);
    const int r_in = b.region(round_up(in, 8));
    b.load_rows(r_in, x, in, in, 0, round_up(in, 8), ridx);
    std::vector<int> regs, slots;
    b.chain_fwd(s, 0, r_in, regs, slots);
    b.work_regions(wmax);
    b.load_rows(b.wcur, grad_y, out, out, 0, round_up(out, 8));
    b.chain_bwd(s, 0, r_in, regs, slots, grad_x != nullptr);
    if (grad_x != nullptr) b.store_rows(b.wcur, grad_x, in, in, 0, accumulate_gx);
    if (!b.ok) return RGNN_ERR_INVALID;
    return launch_program(b.p, stream);
}

// dL/dP (N, 2h) -> weight gradient of the projection part of msg.0 and the contribution to dL/dx, added into wcur
static void proj_bwd(BwdBuilder& b, const rgnn_conv& c, const ConvDims& d, const float* dP, int r_x, int r_tmp) {
    const rgnn_linear& m0 = c.msg.layer[0];
    const float* wp_nat = m0.weight_t + conv_msg0_proj_floats(d) + conv_msg0_edge_floats(d);
    const int ldn = round_up(d.cn, 64);
    // msg.0.weight (h, 2cn+ce): columns [0,cn) act on x_target, [cn,2cn) on x_source (gnn_blocks.py:113)
    if (2 * d.h <= 256) {       // the whole dP row fits a work region
        b.load_rows(b.wnxt, dP, 2 * d.h, 2 * d.h);
        b.wgrad(b.wnxt, r_x, d.h, d.cn, m0.grad_weight, m0.in_features, 0, 0, nullptr);
        b.wgrad(b.wnxt, r_x, d.h, d.cn, m0.grad_weight, m0.in_features, d.cn, d.h, nullptr);
        b.gemm(b.wnxt, r_tmp, wp_nat, ldn, 2 * d.h, 2 * d.h, d.cn, ldn, nullptr);
        b.add_region(b.wcur, r_tmp, d.cn);
        return;
    }
    // wider message layers (h <= 256): the target and the source half of dP pass through the 256-column work region one after the
    // other (a 2 h = 512 column row used to overrun it: wrong, run-to-run different msg.0 gradients at hidden width 128)
    for (int half = 0; half < 2; ++half) {
        b.load_rows(b.wnxt, dP + (size_t)half * d.h, 2 * d.h, d.h);
        b.wgrad(b.wnxt, r_x, d.h, d.cn, m0.grad_weight, m0.in_features, half * d.cn, 0, nullptr);
        b.gemm(b.wnxt, r_tmp, wp_nat + (size_t)half * d.h * ldn, ldn, d.h, d.h, d.cn, ldn, nullptr);
        b.add_region(b.wcur, r_tmp, d.cn);
    }
}

static int conv_nodes_bwd(const rgnn_conv& c, const ConvDims& d, int n_nodes, const float* x, const float* agg,
                          const rgnn_conv* next, const float* x_next, const float* dP_next, float* dx, float* dagg,
                          cudaStream_t stream) {
    BwdBuilder b(n_nodes);
    const int r_cat = b.region(2 * d.cn);
    b.load_rows(r_cat, x, d.cn, d.cn, 0);
    b.load_rows(r_cat, agg, d.cn, d.cn, d.cn);
    std::vector<int> regs, slots;
    b.chain_fwd(c.upd, 0, r_cat, regs, slots);
    const int r_keep = b.region(d.cn);
    int r_xn = -1, r_tmp = -1;
    if (next != nullptr) {
        r_xn = b.region(d.cn);
        r_tmp = b.region(round_up(d.cn, 64));
    }
    b.work_regions();
    b.load_rows(b.wcur, dx, d.cn, d.cn, 0, round_up(d.cn, 8));
    if (next != nullptr) {
        b.load_rows(r_xn, x_next, d.cn, d.cn, 0, round_up(d.cn, 8));
        proj_bwd(b, *next, d, dP_next, r_xn, r_tmp);
    }
    b.zero_region(r_keep, d.cn);
    b.add_region(r_keep, b.wcur, d.cn);           // dL/dx_{l+1}: needed again for the identity residual
    b.chain_bwd(c.upd, 0, r_cat, regs, slots, true);
    b.add_region(b.wcur, r_keep, d.cn);
    b.store_rows(b.wcur, dx, d.cn, d.cn, 0, false, 0);
    b.store_rows(b.wcur, dagg, d.cn, d.cn, 0, false, d.cn);
    if (!b.ok) return RGNN_ERR_INVALID;
    return launch_program(b.p, stream);
}

static int conv_edges_bwd(const rgnn_conv& c, const ConvDims& d, const rgnn_graph& g, const float* emb, const float* P,
                          const float* dagg, float* dP, float* demb, bool first_demb, cudaStream_t stream) {
    RGNN_REQUIRE(c.msg.n == 2, "conv block: msg stack must have 2 ffn_blocks");
    RGNN_CHECK_CUDA(cudaMemsetAsync(dP, 0, (size_t)g.n_nodes * 2 * d.h * sizeof(float), stream));
    if (g.n_edges == 0) return RGNN_OK;
    const rgnn_linear& m0 = c.msg.layer[0];
    const rgnn_linear& m1 = c.msg.layer[1];
    const int wmax = round_up(d.h > d.ce ? (d.h > d.cn ? d.h : d.cn) : (d.ce > d.cn ? d.ce : d.cn), 64);
    BwdBuilder b(g.n_edges);        // 32-row tiles: two CTAs per SM fit and overlap each other's memory phases (measured 3 % faster than 64-row tiles, one CTA)
    const int r_e = b.region(round_up(d.ce, 8));
    const int r_1 = b.region(round_up(d.h, 64));
    const int r_2 = b.region(round_up(d.cn, 64));
    const int s0 = m0.norm_scale ? b.sigma_slot() : -1;
    const int s1 = m1.norm_scale ? b.sigma_slot() : -1;
    b.load_rows(r_e, emb, d.ce, d.ce, 0, round_up(d.ce, 8));
    add_message_layers(b, c, d, g, P, r_e, r_1, r_2, s0, s1);
    b.work_regions(wmax);
    b.load_rows(b.wcur, dagg, d.cn, d.cn, 0, round_up(d.cn, 8), g.tgt);      // d(message) = d(agg)[target]
    b.actnorm_bwd(m1, r_2, s1);
    b.wgrad(b.wcur, r_1, m1.out_features, m1.in_features, m1.grad_weight, m1.in_features, 0, 0, m1.grad_bias);
    b.dgrad(m1.weight, m1.in_features, m1.out_features, m1.in_features);
    b.actnorm_bwd(m0, r_1, s0);                                               // wcur = dz1 (E_tile, h)
    b.wgrad(b.wcur, r_e, d.h, d.ce, m0.grad_weight, m0.in_features, 2 * d.cn, 0, m0.grad_bias);
    const int r_dz1 = b.wcur;
    b.gemm(b.wcur, b.wnxt, m0.weight + 2 * d.cn, m0.in_features, round_up(d.h, 8), d.h, d.ce, round_up(d.ce, 64), nullptr, d.ce);
    b.store_rows(b.wnxt, demb, d.ce, d.ce, 0, !first_demb);
    Step* s = b.add(OP_SEGSUM, r_dz1);
    s->p0 = dP; s->p1 = g.tgt; s->p2 = g.row_ptr;
    s->i0 = 2 * d.h; s->i1 = d.h; s->i2 = 0; s->i4 = 0;
    s = b.add(OP_SCATTER_ADD, r_dz1);
    s->p0 = dP; s->p1 = g.src;
    s->i0 = 2 * d.h; s->i1 = d.h; s->i2 = d.h; s->i4 = 0;
    if (!b.ok) return RGNN_ERR_INVALID;
    return launch_program(b.p, stream);
}

void plan_detector_bwd(const rgnn_detector& net, const rgnn_graph& g, const TakeFn& take, DetPlan* pl) {
    const size_t N = (size_t)g.n_nodes, E = (size_t)g.n_edges;
    pl->dx = take(N * pl->d.cn);
    pl->dP = take(N * 2 * pl->d.h);
    pl->dagg = take(N * pl->d.cn);
    pl->demb = take((E > 0 ? E : 1) * pl->d.ce);
    pl->dh = take(N * pl->link_w);
    pl->dg = take(N * pl->cls_w);
    // tensor-core backward of the message function: per-edge scratch (y1 | dz1 | dz2) and the source-major edge index
    pl->escr = nullptr;
    pl->sidx = nullptr;
    if (mp_bwd_tc_supported(pl->d)) {
        pl->escr = take(mp_bwd_tc_scratch_floats(pl->d, g.n_edges));
        pl->sidx = reinterpret_cast<int*>(take(src_index_ints(g.n_nodes, g.n_edges)));
    }
    // tensor-core chain backward of the edge encoder and the link head: the forward saves layer outputs and sigmas
    size_t cscr = 0;
    auto plan_save = [&](const rgnn_stack& s, size_t rows, bool want_x_in, TcSave* sv) {
        if (want_x_in) sv->x_in = take(rows * round_up(stack_in(s), 8));
        for (int i = 0; i < s.n; ++i) {
            const rgnn_linear& L = s.layer[i];
            if (i + 1 < s.n && !(i == 0 && L.in_features <= 8)) sv->y[i] = take(rows * L.out_features);
            if (L.norm_scale != nullptr) sv->sd[i] = take(rows);
        }
        const size_t need = tc_stack_bwd_scratch_floats(s, (int)rows);
        cscr = cscr > need ? cscr : need;
    };
    pl->enc_tc_bwd = E > 0 && tc_stack_bwd_supported(net.edge_enc);
    if (pl->enc_tc_bwd) plan_save(net.edge_enc, E, false, &pl->enc_save);
    pl->link_tc_bwd = g.n_und > 0 && tc_stack_bwd_supported(net.head_link);
    if (pl->link_tc_bwd) plan_save(net.head_link, (size_t)g.n_und, true, &pl->link_save);
    // node-sized stacks and the node update of the conv blocks
    const rgnn_stack* node_stacks[5] = {&net.node_enc, &net.head_node, &net.head_offset, &net.link_node, &net.class_node};
    for (int i = 0; i < 5; ++i) {
        pl->node_tc_bwd[i] = N > 0 && tc_stack_bwd_supported(*node_stacks[i]) && (i != 0 || tc_proj_supported(pl->d));
        if (pl->node_tc_bwd[i]) plan_save(*node_stacks[i], N, false, &pl->node_save[i]);
    }
    pl->conv_tc_bwd = N > 0;
    for (int l = 0; l < net.n_conv; ++l) pl->conv_tc_bwd = pl->conv_tc_bwd && tc_conv_nodes_bwd_supported(net.conv[l], pl->d);
    if (pl->conv_tc_bwd) {
        for (int l = 0; l < net.n_conv; ++l) { pl->u_save[l] = take(N * pl->d.cn); pl->usd_save[l] = take(N); }
        cscr = cscr > N * pl->d.cn ? cscr : N * pl->d.cn;
    }
    if (cscr > 0) pl->cscr = take(cscr);
}

static int detector_bwd(const rgnn_detector& net, const rgnn_graph& g, const float* node_features,
                        const float* edge_features, const float* g_node_cls, const float* g_node_off,
                        const float* g_link, const float* g_obj, const DetPlan& pl, cudaStream_t stream) {
    const ConvDims& d = pl.d;
    const int N = g.n_nodes, E = g.n_edges, L = net.n_conv;
    const float* xL = pl.x[L];
    int rc;
    // Training envelope: the backward tile programs hold rows of at most 256 channels (dz1 of msg.0, each half of the
    // hoisted projection gradient dP); wider plans (hidden width 256 of the sweep: msg_mlp_hidden_dim 512) run the forward only
    RGNN_REQUIRE(d.h <= 256 && d.cn <= 128, "backward: msg_mlp_hidden_dim %d / node width %d are outside the training envelope (<= 256 / <= 128)", d.h, d.cn);
    // ---- heads: accumulate dL/dx_L in pl.dx ----
    // node-sized stack: tensor-core chain backward when the forward saved its activations, else the recompute tile program
    auto node_bwd = [&](int i, const rgnn_stack& s, const float* y_out, const float* g_top, bool accumulate) -> int {
        if (pl.node_tc_bwd[i] && pl.cscr != nullptr && chain64_bwd_f16_supported(s))       // fixed-shape fp16-split chain with fused weight gradients
            return run_chain64_bwd_f16(s, pl.node_save[i], xL, y_out, g_top, N, pl.dx, accumulate ? 1 : 0, nullptr, nullptr, pl.cscr, stream);
        if (pl.node_tc_bwd[i])
            return tc_stack_bwd(s, pl.node_save[i], xL, nullptr, y_out, g_top, N, pl.cscr, pl.dx, accumulate ? 1 : 0, nullptr, nullptr, stream);
        return stack_bwd(s, xL, nullptr, g_top, N, pl.dx, accumulate, stream);
    };
    if ((rc = node_bwd(1, net.head_node, nullptr, g_node_cls, false))) return rc;
    if ((rc = node_bwd(2, net.head_offset, nullptr, g_node_off, true))) return rc;
    RGNN_CHECK_CUDA(cudaMemsetAsync(pl.dh, 0, (size_t)N * pl.link_w * sizeof(float), stream));
    if (g.n_und > 0 && pl.link_tc_bwd && pl.cscr != nullptr && chain64_bwd_f16_supported(net.head_link)) {
        if ((rc = run_chain64_bwd_f16(net.head_link, pl.link_save, nullptr, nullptr, g_link, g.n_und, pl.dh, 2, g.und_a, g.und_b, pl.cscr, stream)))
            return rc;
    } else if (g.n_und > 0 && pl.link_tc_bwd) {
        if ((rc = tc_stack_bwd(net.head_link, pl.link_save, nullptr, nullptr, nullptr, g_link, g.n_und, pl.cscr, pl.dh, 2, g.und_a,
                               g.und_b, stream)))
            return rc;
    } else if (g.n_und > 0) {
        int lw = pl.link_w;
        for (int i = 0; i < net.head_link.n; ++i) lw = lw > net.head_link.layer[i].out_features ? lw : net.head_link.layer[i].out_features;
        lw = round_up(lw, 64);
        BwdBuilder b(g.n_und, lw <= 64 ? 64 : TR_BWD);
        const int r_in = b.region(pl.link_w);
        Step* s = b.add(OP_LOAD_PAIRSUM, r_in);
        s->p0 = pl.hlink; s->p1 = g.und_a; s->p2 = g.und_b; s->i0 = pl.link_w; s->i1 = pl.link_w;
        std::vector<int> regs, slots;
        b.chain_fwd(net.head_link, 0, r_in, regs, slots);
        b.work_regions(lw);
        const int out = stack_out(net.head_link);
        b.load_rows(b.wcur, g_link, out, out, 0, round_up(out, 8));
        b.chain_bwd(net.head_link, 0, r_in, regs, slots, true);
        s = b.add(OP_PAIR_SCATTER, b.wcur);
        s->p0 = pl.dh; s->p1 = g.und_a; s->p2 = g.und_b; s->i0 = pl.link_w; s->i1 = pl.link_w;
        if (!b.ok) return RGNN_ERR_INVALID;
        if ((rc = launch_program(b.p, stream))) return rc;
    }
    if ((rc = node_bwd(3, net.link_node, pl.hlink, pl.dh, true))) return rc;
    RGNN_CHECK_CUDA(cudaMemsetAsync(pl.dg, 0, (size_t)N * pl.cls_w * sizeof(float), stream));
    if (g.n_clusters > 0) {
        BwdBuilder b(g.n_clusters);
        const int r_in = b.region(pl.cls_w);
        Step* s = b.add(OP_LOAD_SEGMAX, r_in);
        s->p0 = pl.gcls; s->p1 = g.cl_ptr; s->p2 = g.cl_members; s->i0 = pl.cls_w; s->i1 = pl.cls_w;
        std::vector<int> regs, slots;
        b.chain_fwd(net.head_class, 0, r_in, regs, slots);
        b.work_regions();
        const int out = stack_out(net.head_class);
        b.load_rows(b.wcur, g_obj, out, out, 0, round_up(out, 8));
        b.chain_bwd(net.head_class, 0, r_in, regs, slots, true);
        s = b.add(OP_SEGMAX_BWD, b.wcur);
        s->p0 = pl.dg; s->p1 = g.cl_ptr; s->p2 = g.cl_members; s->p3 = pl.gcls; s->i0 = pl.cls_w; s->i1 = pl.cls_w;
        if (!b.ok) return RGNN_ERR_INVALID;
        if ((rc = launch_program(b.p, stream))) return rc;
    }
    if ((rc = node_bwd(4, net.class_node, pl.gcls, pl.dg, true))) return rc;

    // ---- message-passing layers, last to first ----
    const bool tc_edges = pl.escr != nullptr && E > 0;
    const int* sptr = nullptr;
    const int* slist = nullptr;
    if (tc_edges && (rc = build_src_index(g, pl.sidx, &sptr, &slist, stream))) return rc;
    for (int l = L - 1; l >= 0; --l) {
        const bool has_next = l + 1 < L;
        if (pl.conv_tc_bwd && pl.cscr != nullptr && node_bwd_f16_supported(net.conv[l], d)) {
            // fixed-shape fp16-split kernels with the weight gradients fused in (rgnn_node_bwd_f16.cu); cscr[0] holds max |gradient|
            if (has_next && (rc = run_proj_bwd_f16(net.conv[l + 1], d, pl.dP, pl.x[l + 1], N, pl.dx, pl.cscr, stream))) return rc;
            rc = run_upd_bwd_f16(net.conv[l], d, N, pl.x[l], pl.agg[l], pl.u_save[l], pl.usd_save[l], pl.dx, pl.dagg, pl.cscr, stream);
        } else if (pl.conv_tc_bwd) {
            if (has_next && (rc = tc_proj_bwd(net.conv[l + 1], d, pl.dP, pl.x[l + 1], N, pl.dx, stream))) return rc;
            rc = tc_conv_nodes_bwd(net.conv[l], d, N, pl.x[l], pl.agg[l], pl.u_save[l], pl.usd_save[l], pl.dx, pl.dagg, pl.cscr, stream);
        } else {
            rc = conv_nodes_bwd(net.conv[l], d, N, pl.x[l], pl.agg[l], has_next ? &net.conv[l + 1] : nullptr,
                                has_next ? pl.x[l + 1] : nullptr, has_next ? pl.dP : nullptr, pl.dx, pl.dagg, stream);
        }
        if (rc) return rc;
        if (tc_edges && net.conv[l].msg.n == 2 && pl.emb_hl != nullptr && mp_bwd_f16_supported(d))
            rc = run_conv_edges_bwd_f16(net.conv[l], d, g, pl.emb_hl, pl.P[l], pl.dagg, pl.dP, pl.demb, l == L - 1, pl.escr, sptr, slist, stream);
        else if (tc_edges && net.conv[l].msg.n == 2)
            rc = run_conv_edges_bwd_tc(net.conv[l], d, g, pl.emb, pl.P[l], pl.dagg, pl.dP, pl.demb, l == L - 1, pl.escr, sptr, slist, stream);
        else
            rc = conv_edges_bwd(net.conv[l], d, g, pl.emb, pl.P[l], pl.dagg, pl.dP, pl.demb, l == L - 1, stream);
        if (rc) return rc;
    }
    // ---- node encoder (receives dL/dx_0 and the projection gradient of layer 0) ----
    if (pl.node_tc_bwd[0]) {
        if (pl.conv_tc_bwd && pl.cscr != nullptr && node_bwd_f16_supported(net.conv[0], d)) {
            if ((rc = run_proj_bwd_f16(net.conv[0], d, pl.dP, pl.x[0], N, pl.dx, pl.cscr, stream))) return rc;
        } else if ((rc = tc_proj_bwd(net.conv[0], d, pl.dP, pl.x[0], N, pl.dx, stream))) return rc;
        if ((rc = tc_stack_bwd(net.node_enc, pl.node_save[0], node_features, nullptr, pl.x[0], pl.dx, N, pl.cscr, nullptr, 0, nullptr,
                               nullptr, stream)))
            return rc;
    } else {
        BwdBuilder b(N);
        const rgnn_stack& s = net.node_enc;
        const int in = stack_in(s);
        const int r_in = b.region(round_up(in, 8));
        b.load_rows(r_in, node_features, in, in, 0, round_up(in, 8));
        std::vector<int> regs, slots;
        b.chain_fwd(s, 0, r_in, regs, slots);
        const int r_tmp = b.region(round_up(d.cn, 64));
        b.work_regions();
        b.load_rows(b.wcur, pl.dx, d.cn, d.cn, 0, round_up(d.cn, 8));
        proj_bwd(b, net.conv[0], d, pl.dP, regs[s.n - 1], r_tmp);
        b.chain_bwd(s, 0, r_in, regs, slots, false);
        if (!b.ok) return RGNN_ERR_INVALID;
        if ((rc = launch_program(b.p, stream))) return rc;
    }
    // ---- edge encoder ----
    if (E > 0 && pl.enc_tc_bwd) {
        if ((rc = tc_stack_bwd(net.edge_enc, pl.enc_save, edge_features, g.perm, pl.emb, pl.demb, E, pl.cscr, nullptr, 0, nullptr,
                               nullptr, stream)))
            return rc;
    } else if (E > 0) {
        if ((rc = stack_bwd(net.edge_enc, edge_features, g.perm, pl.demb, E, nullptr, false, stream))) return rc;
    }
    return RGNN_OK;
}

}  // namespace rgnn

using namespace rgnn;

extern "C" size_t rgnn_ffn_stack_bwd_workspace_bytes(const rgnn_stack*) { return 0; }

extern "C" int rgnn_ffn_stack_bwd(const rgnn_stack* stack, const float* x, const float* grad_y, int n_rows, float* grad_x,
                                  void*, size_t, void* stream) {
    return stack_bwd(*stack, x, nullptr, grad_y, n_rows, grad_x, false, static_cast<cudaStream_t>(stream));
}

extern "C" int rgnn_detector_bwd(const rgnn_detector* net, const rgnn_graph* g, const float* node_features,
                                 const float* edge_features, const float* grad_node_cls, const float* grad_node_off,
                                 const float* grad_link_cls, const float* grad_obj_cls, void* workspace,
                                 size_t workspace_bytes, void* stream) {
    DetPlan pl;
    int rc = plan_detector(*net, *g, 1, workspace, &pl);
    if (rc) return rc;
    if (pl.bytes > workspace_bytes) {
        set_error("detector workspace too small for backward: need %zu bytes, got %zu", pl.bytes, workspace_bytes);
        return RGNN_ERR_WORKSPACE;
    }
    return detector_bwd(*net, *g, node_features, edge_features, grad_node_cls, grad_node_off, grad_link_cls,
                        grad_obj_cls, pl, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------
// Backward of ONE stand-alone residual_graph_conv_block (gnn_blocks.py:45-113 under torch autograd): the block-level modules of
// the boundary are differentiable on their own (Model_Inference_v1-style compositions, unit tests of a single block).  Generic tile
// programs on the CUDA cores for every channel plan of the training envelope; a training run goes through rgnn_detector_bwd, which
// owns the fused tensor-core kernels.
//   x (N, cn), e (E, ce) target-major, agg (N, cn) and proj (N, 2h) as rgnn_conv_block_fwd left them, d_out (N, cn)
//   -> dx (N, cn), de (E, ce) target-major; parameter gradients are ACCUMULATED into blk's grad_* pointers
// workspace: dagg (N, cn) | dP (N, 2h)
// ---------------------------------------------------------------------------------------------
extern "C" size_t rgnn_conv_block_bwd_workspace_bytes(const rgnn_conv* blk, const rgnn_graph* g) {
    using namespace rgnn;
    ConvDims d;
    if (!conv_dims(*blk, &d)) return 0;
    return align256((size_t)g->n_nodes * d.cn * sizeof(float)) + align256((size_t)g->n_nodes * 2 * d.h * sizeof(float));
}

extern "C" int rgnn_conv_block_bwd(const rgnn_conv* blk, const rgnn_graph* g, const float* x, const float* e, const float* agg,
                                   const float* proj, const float* d_out, float* dx, float* de, void* workspace,
                                   size_t workspace_bytes, void* stream_) {
    using namespace rgnn;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    ConvDims d;
    if (!conv_dims(*blk, &d)) return RGNN_ERR_INVALID;
    RGNN_REQUIRE(d.h <= 256 && d.cn <= 128, "conv_block_bwd: msg_mlp_hidden_dim %d / node width %d are outside the training envelope", d.h, d.cn);
    RGNN_REQUIRE(workspace_bytes >= rgnn_conv_block_bwd_workspace_bytes(blk, g), "conv_block_bwd: workspace too small");
    const int N = g->n_nodes;
    float* dagg = static_cast<float*>(workspace);
    float* dP = reinterpret_cast<float*>(static_cast<char*>(workspace) + align256((size_t)N * d.cn * sizeof(float)));
    int rc;
    // node update: dx <- d_out (identity residual) + upd backward w.r.t. x; dagg
    RGNN_CHECK_CUDA(cudaMemcpyAsync(dx, d_out, (size_t)N * d.cn * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    if ((rc = conv_nodes_bwd(*blk, d, N, x, agg, nullptr, nullptr, nullptr, dx, dagg, stream))) return rc;
    // message function: dP (hoisted node half), de, weight gradients of msg.0's edge part / bias and of msg.1
    if ((rc = conv_edges_bwd(*blk, d, *g, e, proj, dagg, dP, de, true, stream))) return rc;
    // hoisted projection: dx += dP W_proj, dW_msg0[:, 0:2cn] += dP^T x
    BwdBuilder b(N);
    const int r_x = b.region(d.cn);
    b.load_rows(r_x, x, d.cn, d.cn, 0, round_up(d.cn, 8));
    const int r_tmp = b.region(round_up(d.cn, 64));
    b.work_regions();
    b.load_rows(b.wcur, dx, d.cn, d.cn, 0, round_up(d.cn, 8));
    proj_bwd(b, *blk, d, dP, r_x, r_tmp);
    b.store_rows(b.wcur, dx, d.cn, d.cn, 0, false, 0);
    if (!b.ok) return RGNN_ERR_INVALID;
    return launch_program(b.p, stream);
}

