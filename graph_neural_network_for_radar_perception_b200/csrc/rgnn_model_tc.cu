// Host side of the tensor-core row-MLP programs: turns the C-ABI parameter structs (rgnn_stack, rgnn_conv) into
// TcPrograms (rgnn_rowmlp_tc.cuh) -- stage lists with TMEM columns assigned -- for the encoders, the node update of
// the conv block (+ the hoisted projection of the next block) and the four heads of the detector.
#include "rgnn_model.h"
#include "rgnn_rowmlp_tc.cuh"

namespace rgnn {

extern int g_use_tensor_cores_flag();

// TMEM columns: three 128-column regions rotate as (A_hi, A_lo, D); a stage's result goes back in place (hi over D,
// lo over the consumed A_lo), then D becomes A_hi and the old A_hi the next D.  [384, 400) holds narrow raw inputs,
// [496, 512) the row-statistics exchange.
constexpr int R0 = 0, R1 = 128, R2 = 256, R_IN = 384;

static const float* tc_weights(const rgnn_linear& L) {
    return L.weight_t + (size_t)round_up(L.in_features, 8) * round_up(L.out_features, 64);
}

// Transposed images for the backward (d input = dz W): the layer W^T has K' = out_features and N' = in_features; inputs
// wider than 128 are packed as independent blocks of 128 output columns (each its own chunk stream).
static bool tc_has_transposed(int in_features, int out_features) {
    return in_features >= 16 && in_features % 8 == 0 && in_features <= 256 && out_features <= 256 && tc_kp(out_features) <= 128;
}
static int tc_t_blocks(int in_features) { return (in_features + 127) / 128; }
static int tc_t_block_cols(int in_features, int blk) { return in_features - 128 * blk < 128 ? in_features - 128 * blk : 128; }
static size_t tc_t_block_floats(int in_features, int out_features, int blk) {
    return tc_pack_floats(out_features, tc_t_block_cols(in_features, blk));
}
static size_t tc_t_pack_floats(int in_features, int out_features) {
    if (!tc_has_transposed(in_features, out_features)) return 0;
    size_t n = 0;
    for (int b = 0; b < tc_t_blocks(in_features); ++b) n += tc_t_block_floats(in_features, out_features, b);
    return n;
}
static const float* tc_weights_t(const rgnn_linear& L, int blk) {
    const float* p = tc_weights(L) + tc_pack_floats(L.in_features, L.out_features);
    for (int b = 0; b < blk; ++b) p += tc_t_block_floats(L.in_features, L.out_features, b);
    return p;
}

struct TcBuilder {
    TcProgram p;
    bool ok = true;
    int hi = R0, lo = R1, d = R2;
    int width = 0;          // padded width of the activation currently sitting in (hi, lo)

    explicit TcBuilder(int n_rows) {
        memset(&p, 0, sizeof(p));
        p.n_rows = n_rows;
    }
    TcStage* new_stage() {
        if (p.n_stages >= TC_MAX_STAGES) { ok = false; set_error("tensor-core program too long"); return &p.st[0]; }
        TcStage* s = &p.st[p.n_stages++];
        memset(s, 0, sizeof(*s));
        s->epi.y_hi = s->epi.y_lo = -1;
        return s;
    }
    void input(int mode, const float* p0, int ld0, int w0, const float* p1, int ld1, int w1, const int* i0, const int* i1) {
        TcInput& in = p.in;
        in.mode = mode; in.p0 = p0; in.ld0 = ld0; in.w0 = w0; in.p1 = p1; in.ld1 = ld1; in.w1 = w1; in.i0 = i0; in.i1 = i1;
        in.k_pad = round_up(w0 + w1, 8);
        in.a_hi = hi; in.a_lo = lo;
        width = in.k_pad;
    }
    static void fill_epi(TcEpi& e, int d, int n_true, const float* bias, const float* scale, const float* shift, int act) {
        e.d = d; e.n_true = n_true; e.n_cols = tc_np(n_true);
        e.bias = bias; e.scale = scale; e.shift = shift; e.act = act;
    }
    // One Linear (+ norm + act) consuming the current activation.  `part` selects output rows [n_off, n_off+n_part) of a
    // wider layer (used for the 256-wide projection, evaluated as two 128-column stages that share the A operand).
    TcEpi* layer(const float* wtc, int in_features, int out_features, const float* bias, const float* scale, const float* shift,
                 int act, bool feed, int n_off = 0, int n_part = -1) {
        const int Kp = tc_kp(in_features), Np = tc_np(out_features);
        const int kc = tc_chunk_k(Kp, Np), nch = Kp / kc;
        const int N = n_part < 0 ? Np : n_part;
        TcStage* s = new_stage();
        if (Kp != width || nch > TC_MAX_MMA || N > 128 || (scale != nullptr && (n_part >= 0 || (Np != 32 && Np != 64 && Np != 128) || Np != out_features))) {
            ok = false;
            set_error("layer %dx%d not expressible as a tensor-core stage", out_features, in_features);
            return &s->epi;
        }
        s->n_mma = nch;
        for (int c = 0; c < nch; ++c) {
            TcMma& m = s->mma[c];
            m.a_hi = hi + c * kc; m.a_lo = lo + c * kc; m.d = d;
            m.N = N; m.K = kc; m.ldn = Np; m.n_off = n_off; m.acc = c > 0;
            m.w = wtc + (size_t)c * 2 * kc * Np;
        }
        fill_epi(s->epi, d, n_part < 0 ? out_features : n_part, bias ? bias + n_off : nullptr, scale, shift, act);
        if (n_part >= 0) s->epi.n_cols = n_part;
        if (feed) {
            s->epi.y_hi = d; s->epi.y_lo = lo;
            const int t = hi; hi = d; d = t;        // (hi, lo, d) -> (d, lo, hi)
            width = tc_kp(out_features);
            if (width != Np && width != out_features) { ok = false; set_error("layer width %d cannot feed a tensor-core stage", out_features); }
        }
        return &s->epi;
    }
    TcEpi* layer(const rgnn_linear& L, bool feed) {
        return layer(tc_weights(L), L.in_features, L.out_features, L.bias, L.norm_scale, L.norm_shift, L.activation, feed);
    }
    // First two layers of an encoder: L0 = (<= 7 inputs -> 256, no norm), L1 = (256 -> <= 128).  L0 is so narrow that the
    // workers evaluate it on the CUDA cores straight into the A operand (input mode LIN0), 128 columns at a time; L1
    // consumes the two halves as K chunks 0-1 and 2-3 of its accumulator.
    bool encoder_head_lin0(const rgnn_linear& L0, const rgnn_linear& L1) {
        const int Np1 = tc_np(L1.out_features);
        if (L0.in_features > 7 || L0.out_features != 256 || L0.norm_scale != nullptr || L1.in_features != 256 || Np1 > 128 ||
            Np1 != L1.out_features || tc_chunk_k(256, Np1) != 64 || p.in.mode != TC_IN_ROWS || p.in.p1 != nullptr) {
            ok = false;
            set_error("encoder head %dx%d / %dx%d not expressible on tensor cores", L0.out_features, L0.in_features, L1.out_features, L1.in_features);
            return false;
        }
        p.in.mode = TC_IN_LIN0;
        p.in.p1 = L0.weight;
        p.in.lin_b = L0.bias;
        p.in.lin_act = L0.activation;
        p.in.k_pad = 128;
        p.in.a_hi = R0; p.in.a_lo = R1;
        const float* w1 = tc_weights(L1);
        for (int half = 0; half < 2; ++half) {
            TcStage* s = new_stage();
            s->n_mma = 2;
            for (int c = 0; c < 2; ++c) {
                TcMma& m = s->mma[c];
                m.a_hi = R0 + 64 * c; m.a_lo = R1 + 64 * c; m.d = R2;
                m.N = Np1; m.K = 64; m.ldn = Np1; m.n_off = 0; m.acc = (half | c) != 0;
                m.w = w1 + (size_t)(2 * half + c) * 2 * 64 * Np1;
            }
            if (half == 0) {
                s->epi.refill = 1;
            } else {
                fill_epi(s->epi, R2, L1.out_features, L1.bias, L1.norm_scale, L1.norm_shift, L1.activation);
                s->epi.y_hi = R2; s->epi.y_lo = R1;
            }
        }
        hi = R2; lo = R1; d = R0;
        width = Np1;
        return true;
    }
    // (kept for reference / other shapes) L0 evaluated on the tensor cores in four 64-column blocks (double buffered in
    // R1 / R2) that L1 consumes as K chunks, accumulating in R0.
    bool encoder_head(const rgnn_linear& L0, const rgnn_linear& L1) {
        const int Kp0 = tc_kp(L0.in_features), N0 = L0.out_features, Np1 = tc_np(L1.out_features);
        if (Kp0 != 8 || N0 != 256 || L0.norm_scale != nullptr || L1.in_features != 256 || Np1 > 128 || Np1 != L1.out_features ||
            tc_chunk_k(256, Np1) != 64 || width != 8) {
            ok = false;
            set_error("encoder head %dx%d / %dx%d not expressible on tensor cores", L0.out_features, L0.in_features, L1.out_features, L1.in_features);
            return false;
        }
        const float* w0 = tc_weights(L0);
        const float* w1 = tc_weights(L1);
        const int a_hi = p.in.a_hi, a_lo = p.in.a_lo;
        const int yb[2] = {R1, R2};
        for (int blk = 0; blk <= 4; ++blk) {
            TcStage* s = new_stage();
            int j = 0;
            if (blk > 0) {          // consume block blk-1 as K chunk blk-1 of L1
                TcMma& m = s->mma[j++];
                m.a_hi = yb[(blk - 1) & 1]; m.a_lo = yb[(blk - 1) & 1] + 64; m.d = R0;
                m.N = Np1; m.K = 64; m.ldn = Np1; m.n_off = 0; m.acc = blk > 1;
                m.w = w1 + (size_t)(blk - 1) * 2 * 64 * Np1;
            }
            if (blk < 4) {          // produce block blk of L0
                TcMma& m = s->mma[j++];
                m.a_hi = a_hi; m.a_lo = a_lo; m.d = yb[blk & 1];
                m.N = 64; m.K = 8; m.ldn = 256; m.n_off = 64 * blk; m.acc = 0;
                m.w = w0;
                fill_epi(s->epi, yb[blk & 1], 64, L0.bias ? L0.bias + 64 * blk : nullptr, nullptr, nullptr, L0.activation);
                s->epi.y_hi = yb[blk & 1]; s->epi.y_lo = yb[blk & 1] + 64;
            } else {                // full epilogue of L1
                fill_epi(s->epi, R0, L1.out_features, L1.bias, L1.norm_scale, L1.norm_shift, L1.activation);
                s->epi.y_hi = R0; s->epi.y_lo = R1;
            }
            s->n_mma = j;
        }
        hi = R0; lo = R1; d = R2;
        width = Np1;
        return true;
    }
    TcEpi& last_epi() { return p.st[p.n_stages - 1].epi; }
    int run(cudaStream_t stream) {
        if (!ok) return RGNN_ERR_INVALID;
        // must mirror the kernel's `staged_in` test: staged input rows occupy ring slot 2
        const TcInput& in = p.in;
        const bool pair = in.mode == TC_IN_PAIRSUM;
        const int srow = pair ? 2 * in.k_pad : in.k_pad;
        const bool staged = (in.mode == TC_IN_ROWS && in.i0 == nullptr && (in.k_pad % 32) == 0 && in.w0 + in.w1 == in.k_pad &&
                             ((in.ld0 | in.w0 | in.ld1 | in.w1) & 3) == 0) ||
                            (pair && (in.k_pad % 32) == 0 && (in.ld0 & 3) == 0 && in.w0 == in.k_pad && srow <= 128);
        p.n_slots = staged ? 2 : 3;
        return launch_rowmlp_tc(p, stream);
    }
};

static bool stack_layers_ok(const rgnn_stack& s, int first) {
    for (int i = first; i < s.n; ++i) {
        const rgnn_linear& L = s.layer[i];
        if (L.weight_t == nullptr || L.in_features > 128 || L.in_features % 8 != 0 || L.out_features > 128) return false;
        if (L.norm_scale != nullptr && L.out_features != 32 && L.out_features != 64 && L.out_features != 128) return false;
        if (i + 1 < s.n && L.out_features % 32 != 0) return false;
    }
    return true;
}

// Can the whole nn.Sequential run as one tensor-core program?  Two shapes are covered: every layer <= 128 wide with the
// input a multiple of 8, or an encoder whose first block is (<= 8 -> 256, no norm) followed by (256 -> <= 128).
bool tc_stack_supported(const rgnn_stack& s) {
    if (!g_use_tensor_cores_flag() || s.n < 1) return false;
    const rgnn_linear& L0 = s.layer[0];
    if (s.n >= 2 && L0.in_features <= 8 && L0.out_features == 256 && L0.norm_scale == nullptr && L0.weight_t != nullptr) {
        const rgnn_linear& L1 = s.layer[1];
        if (L1.in_features != 256 || L1.weight_t == nullptr || (L1.out_features != 64 && L1.out_features != 128)) return false;
        if (L1.norm_scale == nullptr) return false;
        int stages = 5;
        for (int i = 2; i < s.n; ++i) ++stages;
        return stages + 2 <= TC_MAX_STAGES && stack_layers_ok(s, 2);
    }
    return s.n + 2 <= TC_MAX_STAGES && stack_layers_ok(s, 0);
}

static void set_store(TcEpi& e, float* dst, int ld, int w) {
    e.store = dst; e.store_ld = ld; e.store_w = w;
}

// training step: what the forward leaves behind for the tensor-core backward (tc_stack_bwd)
static void apply_save(TcEpi& e, const rgnn_linear& L, const TcSave* save, int i, bool is_last) {
    if (save == nullptr) return;
    if (L.norm_scale != nullptr) e.sd_store = save->sd[i];
    if (!is_last && save->y[i] != nullptr) set_store(e, save->y[i], L.out_features, L.out_features);
}

static void add_stack(TcBuilder& b, const rgnn_stack& s, bool feed_last, const TcSave* save = nullptr) {
    int first = 0;
    if (s.layer[0].in_features <= 8 && s.layer[0].out_features == 256 && s.n >= 2) {
        if (s.layer[0].in_features <= 7) {
            b.encoder_head_lin0(s.layer[0], s.layer[1]);
        } else {
            b.p.in.a_hi = R_IN;             // the narrow raw input lives outside the rotating regions
            b.p.in.a_lo = R_IN + 8;
            b.encoder_head(s.layer[0], s.layer[1]);
        }
        apply_save(b.last_epi(), s.layer[1], save, 1, s.n == 2);
        first = 2;
    }
    for (int i = first; i < s.n; ++i) {
        TcEpi* e = b.layer(s.layer[i], feed_last || i + 1 < s.n);
        apply_save(*e, s.layer[i], save, i, i + 1 == s.n);
    }
    if (save != nullptr && save->x_in != nullptr) b.p.in.in_store = save->x_in;
}

// hoisted node part of the next conv block's msg.0: P = [x W_t^T + b | x W_s^T], two 128-column stages
static void add_projection(TcBuilder& b, const rgnn_conv& c, const ConvDims& d, float* P) {
    const float* base = c.msg.layer[0].weight_t;
    const float* bias2h = base + conv_msg0_proj_floats(d) + conv_msg0_edge_floats(d) + conv_msg0_projnat_floats(d);
    const float* wtc = base + conv_msg0_tc_offset(d) + mp_tc_pack_floats(d);
    for (int half = 0; half < 2; ++half) {
        TcEpi* e = b.layer(wtc, d.cn, 2 * d.h, bias2h, nullptr, nullptr, 0, false, half * d.h, d.h);
        set_store(*e, P + half * d.h, 2 * d.h, d.h);
    }
}

bool tc_proj_supported(const ConvDims& d) { return g_use_tensor_cores_flag() && d.cn == 64 && d.h == 128; }
// forward image of the hoisted projection (cn -> 2h) followed by the two transposed images of its backward
// (dx += dP_t W_t + dP_s W_s: K' = h, N' = cn each)
static size_t tc_proj_fwd_floats(const ConvDims& d) { return tc_pack_floats(d.cn, 2 * d.h); }
size_t tc_proj_pack_floats(const ConvDims& d) {
    return (d.cn == 64 && d.h == 128) ? tc_proj_fwd_floats(d) + 2 * tc_pack_floats(d.h, d.cn) : 0;
}

static size_t tc_linear_tc_floats(int in_features, int out_features) {
    return (in_features > 256 || out_features > 256) ? 0 : tc_pack_floats(in_features, out_features) + tc_t_pack_floats(in_features, out_features);
}
// tf32 chunk streams (forward + transposed) followed by the fp16 image pair of the f16 kernels
size_t tc_linear_pack_floats(int in_features, int out_features) {
    return tc_linear_tc_floats(in_features, out_features) + f16_image_floats(in_features, out_features);
}
const float* f16_weights(const rgnn_linear& L) { return tc_weights(L) + tc_linear_tc_floats(L.in_features, L.out_features); }

int tc_pack_linear(const rgnn_linear& L, cudaStream_t stream) {
    if (L.in_features > 256 || L.out_features > 256) return RGNN_OK;
    {
        const int rc16 = f16_pack_linear(L, const_cast<float*>(f16_weights(L)), stream);
        if (rc16) return rc16;
    }
    const int Kp = tc_kp(L.in_features), Np = tc_np(L.out_features);
    float* dst = const_cast<float*>(tc_weights(L));
    int rc = pack_tc(L.weight, L.in_features, 0, L.out_features, 0, Np, 0, L.in_features, Kp, tc_chunk_k(Kp, Np), true, dst, stream);
    if (rc || !tc_has_transposed(L.in_features, L.out_features)) return rc;
    for (int b = 0; b < tc_t_blocks(L.in_features); ++b) {     // W^T: rows (n) = input channels of block b, columns (k) = output channels
        const int Nt = tc_t_block_cols(L.in_features, b), Npt = tc_np(Nt), Kpt = tc_kp(L.out_features);
        rc = pack_tc(L.weight, L.in_features, 128 * b, Nt, 0, Npt, 0, L.out_features, Kpt, tc_chunk_k(Kpt, Npt), true,
                     const_cast<float*>(tc_weights_t(L, b)), stream, true);
        if (rc) return rc;
    }
    return RGNN_OK;
}

int tc_pack_projection(const rgnn_conv& c, const ConvDims& d, cudaStream_t stream) {
    if (tc_proj_pack_floats(d) == 0) return RGNN_OK;
    const rgnn_linear& m0 = c.msg.layer[0];
    float* dst = const_cast<float*>(m0.weight_t) + conv_msg0_tc_offset(d) + mp_tc_pack_floats(d);
    const int Kp = tc_kp(d.cn), Np = 2 * d.h, kc = tc_chunk_k(Kp, Np);
    int rc = pack_tc(m0.weight, m0.in_features, 0, d.h, 0, Np, 0, d.cn, Kp, kc, false, dst, stream);          // target half
    if (rc) return rc;
    rc = pack_tc(m0.weight, m0.in_features, 0, d.h, d.h, Np, d.cn, d.cn, Kp, kc, false, dst, stream);      // source half
    if (rc) return rc;
    // backward: element (n = node channel, k = projection column) = msg.0.weight[k][half * cn + n]
    const int Kt = tc_kp(d.h), Nt = tc_np(d.cn);
    for (int half = 0; half < 2; ++half) {
        rc = pack_tc(m0.weight, m0.in_features, half * d.cn, d.cn, 0, Nt, 0, d.h, Kt, tc_chunk_k(Kt, Nt), true,
                     dst + tc_proj_fwd_floats(d) + (size_t)half * tc_pack_floats(d.h, d.cn), stream, true);
        if (rc) return rc;
    }
    return RGNN_OK;
}

// dx (N, cn) += dP_t W_t + dP_s W_s: the gradient of the hoisted projection P = [x W_t^T + b | x W_s^T] w.r.t. x, and the
// weight gradient of the node columns of msg.0:  dW[:, 0:cn] += dP_t^T x,  dW[:, cn:2cn] += dP_s^T x
int tc_proj_bwd(const rgnn_conv& c, const ConvDims& d, const float* dP, const float* x, int n_nodes, float* dx, cudaStream_t stream) {
    const rgnn_linear& m0 = c.msg.layer[0];
    const float* wt = m0.weight_t + conv_msg0_tc_offset(d) + mp_tc_pack_floats(d) + tc_proj_fwd_floats(d);
    for (int half = 0; half < 2; ++half) {
        TcBuilder b(n_nodes);
        b.input(TC_IN_ROWS, dP + half * d.h, 2 * d.h, d.h, nullptr, 0, 0, nullptr, nullptr);
        TcEpi* e = b.layer(wt + (size_t)half * tc_pack_floats(d.h, d.cn), d.h, d.cn, nullptr, nullptr, nullptr, 0, false);
        set_store(*e, dx, d.cn, d.cn);
        e->store_mode = 1;
        int rc = b.run(stream);
        if (rc) return rc;
        if (m0.grad_weight != nullptr) {
            rc = launch_wgrad_tc(dP + half * d.h, 2 * d.h, d.h, x, d.cn, d.cn, n_nodes, m0.grad_weight + half * d.cn, m0.in_features, 1,
                                 nullptr, nullptr, stream);
            if (rc) return rc;
        }
    }
    return RGNN_OK;
}

// backward of the node update out = x + upd(cat(x, agg)) (one ffn_block): dx holds d out on entry and d x on exit
// (residual + the x half of the update's input gradient), dagg receives the agg half.  u = the update's output before the
// residual and its sigma were saved by the forward.
int tc_conv_nodes_bwd(const rgnn_conv& c, const ConvDims& d, int n_nodes, const float* x, const float* agg, const float* u,
                      const float* sd, float* dx, float* dagg, float* dz_scratch, cudaStream_t stream) {
    const rgnn_linear& L = c.upd.layer[0];
    TcBuilder b(n_nodes);
    b.input(TC_IN_BWD, dx, d.cn, d.cn, nullptr, 0, 0, nullptr, nullptr);
    TcBwd& bw = b.p.in.bwd;
    memset(&bw, 0, sizeof(bw));
    bw.y = u; bw.y_ld = d.cn; bw.sd = sd; bw.scale = L.norm_scale; bw.shift = L.norm_shift;
    bw.g_scale = L.grad_norm_scale; bw.g_shift = L.grad_norm_shift; bw.act = L.activation;
    b.p.in.bwd_store = dz_scratch;
    // d cat = dz W_upd (cn -> 2cn): two stages of cn columns sharing the A operand
    TcEpi* e0 = b.layer(tc_weights_t(L, 0), d.cn, 2 * d.cn, nullptr, nullptr, nullptr, 0, false, 0, d.cn);
    set_store(*e0, dx, d.cn, d.cn);
    e0->store_mode = 1;                       // identity residual: d x = d out + ...
    TcEpi* e1 = b.layer(tc_weights_t(L, 0), d.cn, 2 * d.cn, nullptr, nullptr, nullptr, 0, false, d.cn, d.cn);
    set_store(*e1, dagg, d.cn, d.cn);
    int rc = b.run(stream);
    if (rc) return rc;
    if (L.grad_weight != nullptr || L.grad_bias != nullptr) {
        rc = launch_wgrad_tc(dz_scratch, d.cn, d.cn, x, d.cn, d.cn, n_nodes, L.grad_weight, L.in_features, 1, L.grad_bias, nullptr, stream);
        if (rc) return rc;
        rc = launch_wgrad_tc(dz_scratch, d.cn, d.cn, agg, d.cn, d.cn, n_nodes, L.grad_weight ? L.grad_weight + d.cn : nullptr,
                             L.in_features, 1, nullptr, nullptr, stream);
    }
    return rc;
}

bool tc_conv_nodes_bwd_supported(const rgnn_conv& c, const ConvDims& d) {
    const rgnn_linear& L = c.upd.layer[0];
    return rgnn_get_option("tensor_cores_bwd") != 0 && tc_proj_supported(d) && c.upd.n == 1 && tc_stack_supported(c.upd) &&
           L.norm_scale != nullptr && L.in_features == 2 * d.cn && d.cn == 64 && tc_has_transposed(L.in_features, L.out_features);
}

// ---- the programs of the detector forward ----------------------------------------------------------------

int tc_run_stack(const rgnn_stack& s, const float* x, const int* ridx, int n_rows, float* y, cudaStream_t stream, const TcSave* save) {
    TcBuilder b(n_rows);
    b.input(TC_IN_ROWS, x, stack_in(s), stack_in(s), nullptr, 0, 0, ridx, nullptr);
    add_stack(b, s, false, save);
    set_store(b.last_epi(), y, stack_out(s), stack_out(s));
    return b.run(stream);
}

int tc_run_node_encoder(const rgnn_stack& enc, const rgnn_conv& first, const ConvDims& d, const float* node_features, int n_nodes,
                        float* x0, float* P0, cudaStream_t stream, const TcSave* save) {
    TcBuilder b(n_nodes);
    b.input(TC_IN_ROWS, node_features, stack_in(enc), stack_in(enc), nullptr, 0, 0, nullptr, nullptr);
    add_stack(b, enc, true, save);
    set_store(b.last_epi(), x0, d.cn, d.cn);
    add_projection(b, first, d, P0);
    return b.run(stream);
}

// out = x + upd(cat(x, agg)); optionally the next block's projection of `out`
int tc_run_conv_nodes(const rgnn_conv& c, const ConvDims& d, int n_nodes, const float* x, const float* agg, float* out,
                      const rgnn_conv* next, float* P_next, cudaStream_t stream, float* u_save, float* sd_save) {
    TcBuilder b(n_nodes);
    b.input(TC_IN_ROWS, x, d.cn, d.cn, agg, d.cn, d.cn, nullptr, nullptr);
    add_stack(b, c.upd, next != nullptr);
    TcEpi& e = b.last_epi();
    e.resid = x; e.resid_ld = d.cn;
    e.pre_store = u_save; e.sd_store = sd_save;
    set_store(e, out, d.cn, d.cn);
    if (next != nullptr) add_projection(b, *next, d, P_next);
    return b.run(stream);
}

int tc_run_pairsum_stack(const rgnn_stack& s, const float* h, int ld, const int* ia, const int* ib, int n_rows, float* y,
                         cudaStream_t stream, const TcSave* save) {
    TcBuilder b(n_rows);
    b.input(TC_IN_PAIRSUM, h, ld, stack_in(s), nullptr, 0, 0, ia, ib);
    add_stack(b, s, false, save);
    set_store(b.last_epi(), y, stack_out(s), stack_out(s));
    return b.run(stream);
}

int tc_run_segmax_stack(const rgnn_stack& s, const float* g, int ld, const int* ptr, const int* members, int n_rows, float* y,
                        cudaStream_t stream) {
    TcBuilder b(n_rows);
    b.input(TC_IN_SEGMAX, g, ld, stack_in(s), nullptr, 0, 0, ptr, members);
    add_stack(b, s, false);
    set_store(b.last_epi(), y, stack_out(s), stack_out(s));
    return b.run(stream);
}

// ---- backward of a row-MLP chain on the tensor cores ------------------------------------------------------------
// The forward of the training step saved every layer's output rows and norm sigmas (TcSave), so the backward is the
// dgrad chain alone: dz_l = norm'(act'(g_l)) in the epilogue / input transform, g_{l-1} = dz_l W_l on the tensor cores
// (transposed weight images), every dz_l written to a scratch buffer; the weight gradients dW_l = dz_l^T x_l and the
// bias gradients are contracted afterwards by rgnn_wgrad_tc.cu.  Shapes covered: every layer <= 128 wide, optionally
// behind the encoder head (<= 7 -> 256 no norm, 256 -> 128), whose activation mask is recomputed from the raw features.

static bool lin0_head(const rgnn_stack& s) {
    return s.n >= 2 && s.layer[0].in_features <= 7 && s.layer[0].out_features == 256 && s.layer[0].norm_scale == nullptr &&
           s.layer[1].in_features == 256 && s.layer[1].out_features == 128;
}

bool tc_stack_bwd_supported(const rgnn_stack& s) {
    if (!tc_stack_supported(s) || rgnn_get_option("tensor_cores_bwd") == 0) return false;
    const int first = lin0_head(s) ? 1 : 0;
    if (!lin0_head(s) && s.layer[0].in_features <= 8) return false;
    for (int i = first; i < s.n; ++i) {
        const rgnn_linear& L = s.layer[i];
        if (!tc_has_transposed(L.in_features, L.out_features)) return false;
        if (i > first || !lin0_head(s)) {
            if (L.in_features != 32 && L.in_features != 64 && L.in_features != 128) return false;
        }
        const bool has_fn = L.norm_scale != nullptr || L.activation;
        if (has_fn && L.out_features != 32 && L.out_features != 64 && L.out_features != 128) return false;
        if (has_fn && L.norm_scale == nullptr) return false;          // act without norm only in the encoder head
    }
    return s.n - first + 3 <= TC_MAX_STAGES;
}

// scratch: dz_l (rows x out_l) for every layer, plus the recomputed first activation a0 (rows x 256) behind an encoder head
size_t tc_stack_bwd_scratch_floats(const rgnn_stack& s, int n_rows) {
    size_t per_row = 0;
    for (int i = 0; i < s.n; ++i) per_row += round_up(s.layer[i].out_features, 8);
    if (lin0_head(s)) per_row += 256;
    return (size_t)(n_rows > 0 ? n_rows : 1) * per_row;
}

int tc_stack_bwd(const rgnn_stack& s, const TcSave& save, const float* x_rows, const int* x_ridx, const float* y_out,
                 const float* g_top, int n_rows, float* scratch, float* dx, int dx_mode, const int* ia, const int* ib,
                 cudaStream_t stream) {
    if (n_rows <= 0) return RGNN_OK;
    const bool head = lin0_head(s);
    const int n = s.n, first = head ? 1 : 0;
    // scratch layout
    float* dz[RGNN_MAX_STACK];
    float* p = scratch;
    for (int i = 0; i < n; ++i) { dz[i] = p; p += (size_t)n_rows * round_up(s.layer[i].out_features, 8); }
    float* a0 = head ? p : nullptr;
    auto fill_bwd = [&](TcBwd& b, int i, const float* y, int y_ld) {
        const rgnn_linear& L = s.layer[i];
        memset(&b, 0, sizeof(b));
        b.y = y; b.y_ld = y_ld; b.sd = L.norm_scale ? save.sd[i] : nullptr;
        b.scale = L.norm_scale; b.shift = L.norm_shift; b.g_scale = L.grad_norm_scale; b.g_shift = L.grad_norm_shift;
        b.act = L.activation;
    };
    TcBuilder b(n_rows);
    const rgnn_linear& Lt = s.layer[n - 1];
    const float* dz_top;
    if (Lt.norm_scale != nullptr || Lt.activation) {
        b.input(TC_IN_BWD, g_top, Lt.out_features, Lt.out_features, nullptr, 0, 0, nullptr, nullptr);
        fill_bwd(b.p.in.bwd, n - 1, y_out, Lt.out_features);
        b.p.in.bwd_store = dz[n - 1];
        dz_top = dz[n - 1];
    } else {
        b.input(TC_IN_ROWS, g_top, Lt.out_features, Lt.out_features, nullptr, 0, 0, nullptr, nullptr);
        dz_top = g_top;
    }
    int dz_top_ld = (Lt.norm_scale != nullptr || Lt.activation) ? b.p.in.k_pad : Lt.out_features;
    for (int l = n - 1; l >= first; --l) {
        const rgnn_linear& L = s.layer[l];
        if (l == 1 && head) {
            // d a0 = dz_1 W_1 (256 wide, two blocks of 128 columns sharing the A operand); dz_0 = act'(d a0) with the mask of
            // a0 = act(W_0 f + b_0) recomputed from the raw features; a0 itself is the operand of dW_1
            for (int half = 0; half < 2; ++half) {
                TcEpi* e = b.layer(tc_weights_t(L, half), L.out_features, 128, nullptr, nullptr, nullptr, 0, false);
                e->is_bwd = 1;
                memset(&e->bwd, 0, sizeof(e->bwd));
                e->bwd.lin0 = 1; e->bwd.lin0_off = 128 * half; e->bwd.act = s.layer[0].activation;
                e->bwd.y_store = a0; e->bwd.y_ld = 256;
                set_store(*e, dz[0] + 128 * half, 256, 128);
            }
            b.p.lin0.f = x_rows; b.p.lin0.ridx = x_ridx; b.p.lin0.W = s.layer[0].weight; b.p.lin0.b = s.layer[0].bias;
            b.p.lin0.ld = s.layer[0].in_features; b.p.lin0.w = s.layer[0].in_features; b.p.lin0.act = s.layer[0].activation;
        } else if (l > 0) {
            const rgnn_linear& Lp = s.layer[l - 1];
            TcEpi* e = b.layer(tc_weights_t(L, 0), L.out_features, L.in_features, nullptr, nullptr, nullptr, 0, true);
            e->is_bwd = 1;
            fill_bwd(e->bwd, l - 1, save.y[l - 1], Lp.out_features);
            set_store(*e, dz[l - 1], round_up(Lp.out_features, 8), Lp.out_features);
        } else if (dx != nullptr) {
            TcEpi* e = b.layer(tc_weights_t(L, 0), L.out_features, L.in_features, nullptr, nullptr, nullptr, 0, false);
            set_store(*e, dx, L.in_features, L.in_features);
            e->store_mode = dx_mode; e->ia = ia; e->ib = ib;
        }
    }
    int rc = b.run(stream);
    if (rc) return rc;
    // weight and bias gradients: dW_l (out x in) += dz_l^T x_l, db_l += column sums of dz_l
    for (int l = n - 1; l >= 0; --l) {
        const rgnn_linear& L = s.layer[l];
        const float* dzl = l == n - 1 ? dz_top : dz[l];
        const int dz_ld = l == n - 1 ? dz_top_ld : round_up(L.out_features, 8);
        const float* xl;
        int x_ld;
        const int* xidx = nullptr;
        if (l == 0 && (head || x_rows != nullptr)) { xl = x_rows; x_ld = L.in_features; xidx = x_ridx; }
        else if (l == 0) { xl = save.x_in; x_ld = round_up(L.in_features, 8); }
        else if (l == 1 && head) { xl = a0; x_ld = 256; }
        else { xl = save.y[l - 1]; x_ld = s.layer[l - 1].out_features; }
        if (L.grad_weight == nullptr && L.grad_bias == nullptr) continue;
        if (L.out_features <= 64 && L.in_features <= 128 && L.in_features > L.out_features && xidx == nullptr) {
            // the narrower operand goes second (<= 64 columns: the two-CTA-per-SM variant): D[m = input][n = output] -> dW[n][m]
            rc = launch_wgrad_tc(xl, x_ld, L.in_features, dzl, dz_ld, L.out_features, n_rows, L.grad_weight, 1, L.in_features,
                                 nullptr, L.grad_bias, stream);
            if (rc) return rc;
            continue;
        }
        for (int m0 = 0; m0 < L.out_features; m0 += 128) {
            const int wa = L.out_features - m0 < 128 ? L.out_features - m0 : 128;
            rc = launch_wgrad_tc(dzl + m0, dz_ld, wa, xl, x_ld, L.in_features, n_rows,
                                 L.grad_weight ? L.grad_weight + (size_t)m0 * L.in_features : nullptr, L.in_features, 1,
                                 L.grad_bias ? L.grad_bias + m0 : nullptr, nullptr, stream, xidx);
            if (rc) return rc;
        }
    }
    return RGNN_OK;
}

}  // namespace rgnn
