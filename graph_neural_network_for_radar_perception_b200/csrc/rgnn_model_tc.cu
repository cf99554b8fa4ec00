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

static void add_stack(TcBuilder& b, const rgnn_stack& s, bool feed_last) {
    int first = 0;
    if (s.layer[0].in_features <= 8 && s.layer[0].out_features == 256 && s.n >= 2) {
        if (s.layer[0].in_features <= 7) {
            b.encoder_head_lin0(s.layer[0], s.layer[1]);
        } else {
            b.p.in.a_hi = R_IN;             // the narrow raw input lives outside the rotating regions
            b.p.in.a_lo = R_IN + 8;
            b.encoder_head(s.layer[0], s.layer[1]);
        }
        first = 2;
    }
    for (int i = first; i < s.n; ++i) b.layer(s.layer[i], feed_last || i + 1 < s.n);
}

static void set_store(TcEpi& e, float* dst, int ld, int w) {
    e.store = dst; e.store_ld = ld; e.store_w = w;
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
size_t tc_proj_pack_floats(const ConvDims& d) { return (d.cn == 64 && d.h == 128) ? tc_pack_floats(d.cn, 2 * d.h) : 0; }

size_t tc_linear_pack_floats(int in_features, int out_features) {
    return (in_features > 256 || out_features > 256) ? 0 : tc_pack_floats(in_features, out_features);
}

int tc_pack_linear(const rgnn_linear& L, cudaStream_t stream) {
    if (L.in_features > 256 || L.out_features > 256) return RGNN_OK;
    const int Kp = tc_kp(L.in_features), Np = tc_np(L.out_features);
    float* dst = const_cast<float*>(tc_weights(L));
    return pack_tc(L.weight, L.in_features, 0, L.out_features, 0, Np, 0, L.in_features, Kp, tc_chunk_k(Kp, Np), true, dst, stream);
}

int tc_pack_projection(const rgnn_conv& c, const ConvDims& d, cudaStream_t stream) {
    if (tc_proj_pack_floats(d) == 0) return RGNN_OK;
    const rgnn_linear& m0 = c.msg.layer[0];
    float* dst = const_cast<float*>(m0.weight_t) + conv_msg0_tc_offset(d) + mp_tc_pack_floats(d);
    const int Kp = tc_kp(d.cn), Np = 2 * d.h, kc = tc_chunk_k(Kp, Np);
    int rc = pack_tc(m0.weight, m0.in_features, 0, d.h, 0, Np, 0, d.cn, Kp, kc, false, dst, stream);          // target half
    if (rc) return rc;
    return pack_tc(m0.weight, m0.in_features, 0, d.h, d.h, Np, d.cn, d.cn, Kp, kc, false, dst, stream);      // source half
}

// ---- the programs of the detector forward ----------------------------------------------------------------

int tc_run_stack(const rgnn_stack& s, const float* x, const int* ridx, int n_rows, float* y, cudaStream_t stream) {
    TcBuilder b(n_rows);
    b.input(TC_IN_ROWS, x, stack_in(s), stack_in(s), nullptr, 0, 0, ridx, nullptr);
    add_stack(b, s, false);
    set_store(b.last_epi(), y, stack_out(s), stack_out(s));
    return b.run(stream);
}

int tc_run_node_encoder(const rgnn_stack& enc, const rgnn_conv& first, const ConvDims& d, const float* node_features, int n_nodes,
                        float* x0, float* P0, cudaStream_t stream) {
    TcBuilder b(n_nodes);
    b.input(TC_IN_ROWS, node_features, stack_in(enc), stack_in(enc), nullptr, 0, 0, nullptr, nullptr);
    add_stack(b, enc, true);
    set_store(b.last_epi(), x0, d.cn, d.cn);
    add_projection(b, first, d, P0);
    return b.run(stream);
}

// out = x + upd(cat(x, agg)); optionally the next block's projection of `out`
int tc_run_conv_nodes(const rgnn_conv& c, const ConvDims& d, int n_nodes, const float* x, const float* agg, float* out,
                      const rgnn_conv* next, float* P_next, cudaStream_t stream) {
    TcBuilder b(n_nodes);
    b.input(TC_IN_ROWS, x, d.cn, d.cn, agg, d.cn, d.cn, nullptr, nullptr);
    add_stack(b, c.upd, next != nullptr);
    TcEpi& e = b.last_epi();
    e.resid = x; e.resid_ld = d.cn;
    set_store(e, out, d.cn, d.cn);
    if (next != nullptr) add_projection(b, *next, d, P_next);
    return b.run(stream);
}

int tc_run_pairsum_stack(const rgnn_stack& s, const float* h, int ld, const int* ia, const int* ib, int n_rows, float* y,
                         cudaStream_t stream) {
    TcBuilder b(n_rows);
    b.input(TC_IN_PAIRSUM, h, ld, stack_in(s), nullptr, 0, 0, ia, ib);
    add_stack(b, s, false);
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

}  // namespace rgnn
