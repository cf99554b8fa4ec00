// Interpreter kernel for row-MLP chains on the tensor cores (program format: rgnn_rowmlp_tc.cuh).
// Roles inside the CTA (1 CTA / SM, persistent over tiles of 128 rows):
//   256 worker threads   thread = (row, half of the columns): input rows -> TMEM, all epilogues, stores
//   MMA warp  (lane 0)   issues every tcgen05.mma (A operand from TMEM, weights from the shared-memory ring)
//   load warp (lane 0)   streams the weight chunks of every stage, L2 -> shared memory, with cp.async.bulk
// Synchronisation: mbarriers full[slot] (bulk copy landed) / empty[slot] (tcgen05.commit: the MMAs that read the slot
// are done) / d_ready (tcgen05.commit: a stage's accumulator is complete), and a named barrier on which the workers
// arrive when the next A operand is in TMEM.
#include <type_traits>

#include "rgnn_pack.cuh"
#include "rgnn_rowmlp_tc.cuh"
#include "rgnn_tc_rows.cuh"
#include "rgnn_tile.cuh"

namespace rgnn {

constexpr int RM_NQ = 2;                    // threads per row
constexpr int RM_NW = 128 * RM_NQ;          // worker threads
constexpr int RM_NT = RM_NW + 128;          // + one warpgroup holding the MMA and the load warp
constexpr int RM_STG_FLOATS = 128 * 128;     // input staging: 128 rows x up to 128 columns (64 KB)
constexpr int RM_CST_LD = 256 + 4;          // per stage: bias[256] (zero padded), scale, shift
constexpr int RM_OFF_STG = 2 * TC_SLOT_FLOATS;                     // aliases ring slot 2
constexpr int RM_OFF_CST = TC_SLOTS * TC_SLOT_FLOATS;
static_assert(RM_STG_FLOATS <= TC_SLOT_FLOATS, "input staging must fit in one ring slot");
constexpr int RM_OFF_W0 = RM_OFF_CST + TC_MAX_STAGES * RM_CST_LD;   // LIN0: [8][256] = W0^T rows 0..6, bias in row 7
constexpr int RM_OFF_BAR = RM_OFF_W0 + 8 * 256;
constexpr int RM_OFF_DACC = RM_OFF_BAR + 16 + 256;                // barriers, TMEM slot, 256 pair-sum node ids
constexpr int RM_DACC_N = 2 * (TC_MAX_STAGES + 1);                // doubles: (d scale, d shift) per stage, last pair = input transform
constexpr size_t RM_SMEM = (size_t)(RM_OFF_DACC + 2 * RM_DACC_N) * 4;
static_assert((RM_OFF_DACC % 2) == 0, "double alignment");
static_assert(RM_SMEM <= 227 * 1024 && (RM_OFF_BAR % 2) == 0, "shared memory budget / mbarrier alignment");

namespace tc {
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
}  // namespace tc

// ---------------------------------------------------------------------------------------------
// input rows -> A operand in TMEM (hi | lo)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_input_block(const TcInput& in, int row_g, bool valid, int c, float (&v)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (!valid) return;
    if (in.mode == TC_IN_ROWS) {
        const size_t r = in.i0 ? (size_t)__ldg(in.i0 + row_g) : (size_t)row_g;
        if (c + 8 <= in.w0 && ((in.ld0 | in.w0) & 3) == 0) {
            const float4* p = reinterpret_cast<const float4*>(in.p0 + r * in.ld0 + c);
            const float4 a = __ldg(p), b = __ldg(p + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else if (c >= in.w0 && c + 8 <= in.w0 + in.w1 && ((in.ld1 | in.w0 | in.w1) & 3) == 0) {
            const float4* p = reinterpret_cast<const float4*>(in.p1 + r * in.ld1 + (c - in.w0));
            const float4 a = __ldg(p), b = __ldg(p + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int cc = c + j;
                if (cc < in.w0) v[j] = __ldg(in.p0 + r * in.ld0 + cc);
                else if (cc - in.w0 < in.w1) v[j] = __ldg(in.p1 + r * in.ld1 + (cc - in.w0));
            }
        }
    } else if (in.mode == TC_IN_PAIRSUM) {
        const size_t a = (size_t)__ldg(in.i0 + row_g), b = (size_t)__ldg(in.i1 + row_g);
        const float4* pa = reinterpret_cast<const float4*>(in.p0 + a * in.ld0 + c);
        const float4* pb = reinterpret_cast<const float4*>(in.p0 + b * in.ld0 + c);
        const float4 a0 = __ldg(pa), a1 = __ldg(pa + 1), b0 = __ldg(pb), b1 = __ldg(pb + 1);
        v[0] = a0.x + b0.x; v[1] = a0.y + b0.y; v[2] = a0.z + b0.z; v[3] = a0.w + b0.w;
        v[4] = a1.x + b1.x; v[5] = a1.y + b1.y; v[6] = a1.z + b1.z; v[7] = a1.w + b1.w;
    } else {   // TC_IN_SEGMAX: max over the member rows of cluster row_g (reference gnn_blocks.py:384-386)
        const int m0 = __ldg(in.i0 + row_g), m1 = __ldg(in.i0 + row_g + 1);
        if (m1 > m0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = -INFINITY;
            for (int m = m0; m < m1; ++m) {
                const float4* p = reinterpret_cast<const float4*>(in.p0 + (size_t)__ldg(in.i1 + m) * in.ld0 + c);
                const float4 a = __ldg(p), b = __ldg(p + 1);
                v[0] = fmaxf(v[0], a.x); v[1] = fmaxf(v[1], a.y); v[2] = fmaxf(v[2], a.z); v[3] = fmaxf(v[3], a.w);
                v[4] = fmaxf(v[4], b.x); v[5] = fmaxf(v[5], b.y); v[6] = fmaxf(v[6], b.z); v[7] = fmaxf(v[7], b.w);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// epilogues
// ---------------------------------------------------------------------------------------------
// result pairs z[CP] (columns col0 .. col0 + 2 CP of the row): residual, store, next A operand
template <int CP>
__device__ __forceinline__ void finish_columns(float2 (&z)[CP], const TcEpi& e, int row_g, bool valid, int col0, uint32_t t_row,
                                               const float2 (&res)[CP] /* prefetched residual */, bool has_res) {
    if (e.pre_store != nullptr && valid) {      // n_true is a multiple of 16 here (norm layers)
        float* o = e.pre_store + (size_t)row_g * e.n_true + col0;
#pragma unroll
        for (int c = 0; c < CP; c += 4)
            if (col0 + 2 * c < e.n_true) stg256(o + 2 * c, z[c], z[c + 1], z[c + 2], z[c + 3]);
    }
    if (has_res) {
#pragma unroll
        for (int c = 0; c < CP; ++c) z[c] = __fadd2_rn(z[c], res[c]);
    }
    if (e.store != nullptr && valid && e.store_mode == 2) {      // d(pair sum): both end points receive the row
        float* oa = e.store + (size_t)__ldg(e.ia + row_g) * e.store_ld + col0;
        float* ob = e.store + (size_t)__ldg(e.ib + row_g) * e.store_ld + col0;
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            if (col0 + 2 * c < e.store_w) { atomicAdd(oa + 2 * c, z[c].x); atomicAdd(ob + 2 * c, z[c].x); }
            if (col0 + 2 * c + 1 < e.store_w) { atomicAdd(oa + 2 * c + 1, z[c].y); atomicAdd(ob + 2 * c + 1, z[c].y); }
        }
    } else if (e.store != nullptr && valid && e.store_mode == 1) {
        float* o = e.store + (size_t)row_g * e.store_ld + col0;
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            if (col0 + 2 * c < e.store_w) o[2 * c] += z[c].x;
            if (col0 + 2 * c + 1 < e.store_w) o[2 * c + 1] += z[c].y;
        }
    } else if (e.store != nullptr && valid) {
        float* o = e.store + (size_t)row_g * e.store_ld + col0;
        if (CP % 4 == 0 && ((e.store_ld | e.store_w) & 7) == 0 && (reinterpret_cast<uintptr_t>(e.store) & 31) == 0) {
#pragma unroll
            for (int c = 0; c + 3 < CP; c += 4)      // full 32-byte sectors per lane
                if (col0 + 2 * c < e.store_w) stg256(o + 2 * c, z[c], z[c + 1], z[c + 2], z[c + 3]);
        } else if (((e.store_ld | e.store_w) & 3) == 0) {
#pragma unroll
            for (int c = 0; c < CP; c += 2)
                if (col0 + 2 * c < e.store_w) *reinterpret_cast<float4*>(o + 2 * c) = make_float4(z[c].x, z[c].y, z[c + 1].x, z[c + 1].y);
        } else {
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                if (col0 + 2 * c < e.store_w) o[2 * c] = z[c].x;
                if (col0 + 2 * c + 1 < e.store_w) o[2 * c + 1] = z[c].y;
            }
        }
    }
    if (e.y_hi >= 0) {
#pragma unroll
        for (int c = 0; c < CP; c += 8) {
            float2 hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) tc::split_tf32(z[c + j], hi[j], lo[j]);
            tc::tmem_st16(t_row + e.y_hi + col0 + 2 * c, hi);
            tc::tmem_st16(t_row + e.y_lo + col0 + 2 * c, lo);
        }
    }
}

// accumulator columns + bias; `cb` = this stage's constants in shared memory (bias zero padded to 256, scale, shift)
template <int CP>
__device__ __forceinline__ void load_acc_bias(float2 (&z)[CP], const TcEpi& e, int col0, uint32_t t_row, const float* __restrict__ cb) {
#pragma unroll
    for (int c = 0; c < CP; c += 8) tc::tmem_ld16(t_row + e.d + col0 + 2 * c, z + c);
    tc::tmem_wait_ld();
#pragma unroll
    for (int c = 0; c < CP; ++c) z[c] = __fadd2_rn(z[c], *reinterpret_cast<const float2*>(cb + col0 + 2 * c));
}

// whole (per-thread share of the) row in registers: needed when the layer normalises over the row
template <int CPT>
__device__ __forceinline__ void epilogue_norm(const TcEpi& e, int row_g, bool valid, int q, uint32_t t_row, int bar_id,
                                              const float* __restrict__ cb) {
    float2 z[CPT / 2];
    const int col0 = q * CPT;
    // the residual row (identity residual of the conv block) is requested before anything else so that its L2 round
    // trip overlaps the normalisation
    float2 res[CPT / 2];
    const bool has_res = e.resid != nullptr;
#pragma unroll
    for (int c = 0; c < CPT / 2; ++c)
        res[c] = (has_res && valid) ? __ldg(reinterpret_cast<const float2*>(e.resid + (size_t)row_g * e.resid_ld + col0) + c) : make_float2(0.f, 0.f);
    load_acc_bias<CPT / 2>(z, e, col0, t_row, cb);
    float sd = 0.f;
    row_norm_act<CPT / 2, RM_NQ>(z, e.n_true, true, cb[256], cb[257], e.act != 0, t_row + TC_XS_COL, q, bar_id, &sd);
    if (e.sd_store != nullptr && valid && q == 0) e.sd_store[row_g] = sd;
    finish_columns<CPT / 2>(z, e, row_g, valid, col0, t_row, res, has_res);
}

// ---------------------------------------------------------------------------------------------
// backward of norm + activation on a gradient row (TcBwd): g (CPT columns of this thread, in registers) -> dz in place.
// y: the layer's forward output for the same columns.  Scalar-gradient partials are returned for the caller to pool.
// ---------------------------------------------------------------------------------------------
template <int CPT>
__device__ __forceinline__ void bwd_transform(float2 (&g)[CPT / 2], const float2 (&y)[CPT / 2], const TcBwd& b, int n_true, bool has_y,
                                              float scale, float shift, float sd, uint32_t t_xs, int q, int bar_id, float& ps,
                                              float& pm) {
    const bool norm = b.scale != nullptr, act = b.act != 0;
    ps = 0.f; pm = 0.f;
    if (!has_y) return;                    // plain Linear: dz = g
    if (!norm) {
        if (act) {
#pragma unroll
            for (int c = 0; c < CPT / 2; ++c) {
                if (!(y[c].x > 0.f)) g[c].x *= LEAKY;
                if (!(y[c].y > 0.f)) g[c].y *= LEAKY;
            }
        }
        return;
    }
    const float inv_scale = scale != 0.f ? 1.f / scale : 0.f;
    const float2 is2 = make_float2(inv_scale, inv_scale), nsh2 = make_float2(-shift * inv_scale, -shift * inv_scale);
    const float2 sc2 = make_float2(scale, scale);
    float2 ps2 = make_float2(0.f, 0.f), pm2 = ps2, sum2 = ps2, dot2 = ps2;
    float2 nv[CPT / 2];
#pragma unroll
    for (int c = 0; c < CPT / 2; ++c) {       // packed f32x2 throughout: the epilogue is instruction-bound (8 worker warps per SM)
        act_bwd_pair(g[c], nv[c], y[c], act, is2, nsh2);
        ps2 = __ffma2_rn(g[c], nv[c], ps2);
        pm2 = __fadd2_rn(pm2, g[c]);
        g[c] = __fmul2_rn(g[c], sc2);
        sum2 = __fadd2_rn(sum2, g[c]);
        dot2 = __ffma2_rn(g[c], nv[c], dot2);
    }
    ps = ps2.x + ps2.y;
    pm = pm2.x + pm2.y;
    float sum_dn = sum2.x + sum2.y, dot = dot2.x + dot2.y;
    row_allreduce2(sum_dn, dot, t_xs, q, bar_id);
    const float inv_den = 1.f / (sd + NORM_EPS);
    const float mean_dn = sum_dn / (float)n_true;
    const float coef = sd > 0.f ? dot / ((float)(n_true - 1) * sd) : 0.f;
    const float2 nm2 = make_float2(-mean_dn, -mean_dn), id2 = make_float2(inv_den, inv_den), nc2 = make_float2(-coef, -coef);
#pragma unroll
    for (int c = 0; c < CPT / 2; ++c) g[c] = __ffma2_rn(nv[c], nc2, __fmul2_rn(__fadd2_rn(g[c], nm2), id2));
}

// forward output columns of this thread's row (32-byte loads: y_ld and col0 are multiples of 8)
template <int CPT>
__device__ __forceinline__ void load_y(float2 (&y)[CPT / 2], const TcBwd& b, int row_g, bool valid, int col0) {
#pragma unroll
    for (int c = 0; c < CPT / 2; ++c) y[c] = make_float2(0.f, 0.f);
    if (b.y == nullptr || !valid) return;
    const float* p = b.y + (size_t)row_g * b.y_ld + col0;
#pragma unroll
    for (int c8 = 0; c8 < CPT / 8; ++c8) ldg256(p + 8 * c8, y[4 * c8], y[4 * c8 + 1], y[4 * c8 + 2], y[4 * c8 + 3]);
}

// pool the per-thread partials of the scalar gradients of one stage: warp sum, then one shared-memory atomic per warp
__device__ __forceinline__ void pool_scalar_grads(float ps, float pm, bool valid, double* dacc, int lane) {
    if (!valid) { ps = 0.f; pm = 0.f; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ps += __shfl_xor_sync(0xffffffffu, ps, o);
        pm += __shfl_xor_sync(0xffffffffu, pm, o);
    }
    if (lane == 0) {
        atomicAdd(dacc, (double)ps);
        atomicAdd(dacc + 1, (double)pm);
    }
}

// backward epilogue of a stage: D = d(layer output) -> dz = norm'(act'(D)) -> store / next A operand
struct BwdScalars { float scale, shift, sd; };
__device__ __forceinline__ BwdScalars load_bwd_scalars(const TcBwd& b, int row_g, bool valid) {
    BwdScalars r;
    const bool norm = b.scale != nullptr;
    r.scale = norm ? __ldg(b.scale) : 1.f;
    r.shift = norm ? __ldg(b.shift) : 0.f;
    r.sd = (norm && valid) ? __ldg(b.sd + row_g) : 0.f;
    return r;
}

template <int CPT>
__device__ __forceinline__ void epilogue_bwd(const TcEpi& e, int row_g, bool valid, int q, uint32_t t_row, int bar_id,
                                             const float2 (&y)[CPT / 2], const BwdScalars& bs, double* dacc, int lane) {
    float2 z[CPT / 2];
    const int col0 = q * CPT;
#pragma unroll
    for (int c = 0; c < CPT / 2; c += 8) tc::tmem_ld16(t_row + e.d + col0 + 2 * c, z + c);
    tc::tmem_wait_ld();
    const TcBwd& b = e.bwd;
    const bool norm = b.scale != nullptr;
    const float scale = bs.scale, shift = bs.shift, sd = bs.sd;
    float ps, pm;
    bwd_transform<CPT>(z, y, b, e.n_true, b.y != nullptr || b.lin0 != 0, scale, shift, sd, t_row + TC_XS_COL + 4, q, bar_id, ps, pm);
    if (norm && b.g_scale != nullptr) pool_scalar_grads(ps, pm, valid, dacc, lane);
    float2 none[CPT / 2];
#pragma unroll
    for (int c = 0; c < CPT / 2; ++c) none[c] = make_float2(0.f, 0.f);
    finish_columns<CPT / 2>(z, e, row_g, valid, col0, t_row, none, false);
}

// no normalisation: 16 columns at a time
__device__ __forceinline__ void epilogue_plain(const TcEpi& e, int row_g, bool valid, int q, uint32_t t_row, const float* __restrict__ cb) {
    const int cpt = e.n_cols / RM_NQ;
    for (int b = 0; b < cpt; b += 16) {
        float2 z[8];
        const int col0 = q * cpt + b;
        load_acc_bias<8>(z, e, col0, t_row, cb);
        if (e.act) {
            const float2 sl = make_float2(LEAKY, LEAKY);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float2 t = __fmul2_rn(z[c], sl);
                z[c].x = fmaxf(z[c].x, t.x);
                z[c].y = fmaxf(z[c].y, t.y);
            }
        }
        float2 none[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) none[c] = make_float2(0.f, 0.f);
        finish_columns<8>(z, e, row_g, valid, col0, t_row, none, false);
    }
}

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
template <bool PROFILE>
__global__ void __launch_bounds__(RM_NT, 1) rowmlp_tc_kernel(const __grid_constant__ TcProgram pg, long long* prof) {
    extern __shared__ __align__(1024) float smem[];
    float* ring = smem;
    float* stg = smem + RM_OFF_STG;
    float* cst = smem + RM_OFF_CST;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RM_OFF_BAR);   // full[] empty[] d_ready
    uint64_t* full = bars;
    uint64_t* empty = bars + TC_SLOTS;
    uint64_t* d_ready = bars + 2 * TC_SLOTS;
    uint32_t* slot_ptr = reinterpret_cast<uint32_t*>(bars + 2 * TC_SLOTS + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < TC_SLOTS; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
        tc::mbar_init(d_ready, 1);
        tc::mbar_init_fence();
    }
    if (warp == 0) tc::tmem_alloc(slot_ptr, 512);
    for (int i = tid; i < pg.n_stages * RM_CST_LD; i += RM_NT) {     // per-stage constants -> shared memory, once
        const int s = i / RM_CST_LD, c = i - s * RM_CST_LD;
        const TcEpi& e = pg.st[s].epi;
        float v = 0.f;
        if (c < 256) v = (e.bias != nullptr && c < e.n_true) ? __ldg(e.bias + c) : 0.f;
        else if (c == 256) v = e.scale != nullptr ? __ldg(e.scale) : 1.f;
        else if (c == 257) v = e.shift != nullptr ? __ldg(e.shift) : 0.f;
        cst[i] = v;
    }
    float* w0s = smem + RM_OFF_W0;
    double* dacc = reinterpret_cast<double*>(smem + RM_OFF_DACC);
    for (int i = tid; i < RM_DACC_N; i += RM_NT) dacc[i] = 0.;
    // first encoder Linear (<= 7 inputs): evaluated on the CUDA cores, either as the input (LIN0 mode) or to recompute the
    // activation mask of that layer in a backward program (pg.lin0)
    const bool lin0_in = pg.in.mode == TC_IN_LIN0;
    const float* l0_f = lin0_in ? pg.in.p0 : pg.lin0.f;
    const int* l0_idx = lin0_in ? pg.in.i0 : pg.lin0.ridx;
    const float* l0_W = lin0_in ? pg.in.p1 : pg.lin0.W;
    const float* l0_b = lin0_in ? pg.in.lin_b : pg.lin0.b;
    const int l0_ld = lin0_in ? pg.in.ld0 : pg.lin0.ld, l0_w = lin0_in ? pg.in.w0 : pg.lin0.w;
    const int l0_act = lin0_in ? pg.in.lin_act : pg.lin0.act;
    if (l0_W != nullptr && (lin0_in || pg.lin0.f != nullptr)) {
        for (int i = tid; i < 8 * 256; i += RM_NT) {
            const int k = i >> 8, c = i & 255;
            float v = 0.f;
            if (k < l0_w) v = __ldg(l0_W + (size_t)c * l0_w + k);
            else if (k == 7 && l0_b != nullptr) v = __ldg(l0_b + c);
            w0s[i] = v;
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *slot_ptr;
    const int n_tiles = (pg.n_rows + 127) / 128;

    if (tid >= RM_NW) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        const int w = warp - RM_NW / 32;
        if (w == 0) {
            // =========================== MMA issue warp ===========================
            uint32_t cnt = 0;
            long long mt[3] = {0, 0, 0}, mlast = 0;   // PROFILE: wait for the A operand | wait for a weight chunk | issue
            if (PROFILE) mlast = clock64();
            auto mtick = [&](int i) { if (PROFILE) { const long long now = clock64(); mt[i] += now - mlast; mlast = now; } };
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int s = 0; s < pg.n_stages; ++s) {
                    const TcStage& st = pg.st[s];
                    group_sync(BAR_Y_READY, RM_NW + 32);      // this stage's A operand is in TMEM
                    tc::tc_fence_after();
                    mtick(0);
                    if (lane == 0) {
                        for (int j = 0; j < st.n_mma; ++j, ++cnt) {
                            const TcMma& m = st.mma[j];
                            const uint32_t slot = cnt % (uint32_t)pg.n_slots, par = (cnt / (uint32_t)pg.n_slots) & 1u;
                            tc::mbar_wait(&full[slot], par);
                            tc::tc_fence_after();
                            mtick(1);
                            // One lane issues every MMA, so this loop is a chain of dependent scalar instructions: with the
                            // descriptor recomputed from the program fields inside a rolled loop it ran ~115 cycles per
                            // MMA (an MMA executes in 16 - 32), i.e. the ISSUE paced every stage (measured: 22 k of 46 k
                            // cycles per edge-encoder tile).  Fields in registers, four independent issues per trip.
                            const uint32_t sb = tc::smem_u32(ring + slot * TC_SLOT_FLOATS);
                            const uint32_t lbo = (uint32_t)m.ldn * 16u;
                            const uint32_t idesc = tc::idesc_tf32(128, m.N);
                            const uint32_t dcol = tmem + (uint32_t)m.d, a_hi = tmem + (uint32_t)m.a_hi, a_lo = tmem + (uint32_t)m.a_lo;
                            const uint32_t lo_off = (uint32_t)(m.K * m.ldn * 4), n_off = (uint32_t)m.n_off * 16u;
                            const int k8 = m.K >> 3;
                            const uint64_t step = (uint64_t)((2 * lbo) >> 4);     // descriptor address units per k-step
                            uint32_t acc = m.acc != 0 ? 1u : 0u;
#pragma unroll
                            for (int p = 0; p < 3; ++p) {      // 3xTF32, small terms first: lo*hi, hi*lo, hi*hi
                                uint32_t ac = p == 0 ? a_lo : a_hi;
                                uint64_t bd = tc::smem_desc(sb + (p == 1 ? lo_off : 0u) + n_off, lbo, 128);
                                int ks = 0;
                                for (; ks + 4 <= k8; ks += 4) {
                                    tc::mma_tf32_ts(dcol, ac, bd, idesc, acc != 0);
                                    tc::mma_tf32_ts(dcol, ac + 8, bd + step, idesc, true);
                                    tc::mma_tf32_ts(dcol, ac + 16, bd + 2 * step, idesc, true);
                                    tc::mma_tf32_ts(dcol, ac + 24, bd + 3 * step, idesc, true);
                                    acc = 1u; ac += 32; bd += 4 * step;
                                }
                                for (; ks < k8; ++ks) {
                                    tc::mma_tf32_ts(dcol, ac, bd, idesc, acc != 0);
                                    acc = 1u; ac += 8; bd += step;
                                }
                            }
                            tc::mma_commit(&empty[slot]);     // slot reusable once these MMAs have read it
                            mtick(2);
                        }
                        tc::mma_commit(d_ready);              // accumulator of the stage complete
                    }
                    __syncwarp();
                }
            }
            if (PROFILE && lane == 0 && prof != nullptr)
                for (int i = 0; i < 3; ++i) prof[blockIdx.x * (2 * TC_MAX_STAGES + 5) + 2 * TC_MAX_STAGES + 2 + i] = mt[i];
        } else if (w == 1) {
            // =========================== weight load warp ===========================
            if (lane == 0) {
                uint32_t cnt = 0;
                for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                    for (int s = 0; s < pg.n_stages; ++s) {
                        const TcStage& st = pg.st[s];
                        for (int j = 0; j < st.n_mma; ++j, ++cnt) {
                            const TcMma& m = st.mma[j];
                            const uint32_t slot = cnt % (uint32_t)pg.n_slots, par = (cnt / (uint32_t)pg.n_slots) & 1u;
                            tc::mbar_wait(&empty[slot], par ^ 1u);
                            const uint32_t bytes = (uint32_t)(m.K * m.ldn * 8);
                            tc::mbar_expect_tx(&full[slot], bytes);
                            tc::bulk_g2s(ring + slot * TC_SLOT_FLOATS, m.w, bytes, &full[slot]);
                        }
                    }
                }
            }
        }
    } else {
        // =========================== worker warps ===========================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");   // all the auxiliary warpgroup released: 128 x (168 - 40) = 256 x (232 - 168)
        const int row = tid & 127, q = tid >> 7;
        const int bar_id = 1 + (row >> 5);
        const uint32_t t_row = tmem + ((uint32_t)(row & ~31) << 16);
        uint32_t dphase = 0;
        long long pt[2 * TC_MAX_STAGES + 2], tlast = 0;
        if (PROFILE) {
            for (int i = 0; i < 2 * TC_MAX_STAGES + 2; ++i) pt[i] = 0;
            tlast = clock64();
        }
        auto tick = [&](int i) {
            if (PROFILE && tid == 0) { const long long now = clock64(); pt[i] += now - tlast; tlast = now; }
        };
        // Input staging: rows that can be fetched as whole 16-byte-aligned pieces (plain / concatenated rows without
        // a row index, pair sums) are copied by cp.async into shared memory ONE TILE AHEAD, with every warp
        // instruction covering whole rows (coalesced); the row-owning threads then pick their columns out of the
        // XOR-swizzled staging tile.  Per-thread sector gathers of 256-byte rows cost ~9k cycles per tile here.
        const TcInput& in = pg.in;
        const bool pair = in.mode == TC_IN_PAIRSUM;
        const int srow = pair ? 2 * in.k_pad : in.k_pad;          // staged floats per row
        const bool staged_in = pg.n_slots == 2;                  // decided by the host (TcBuilder::run): staging aliases ring slot 2
        // pair sums: the two node ids of every row of a tile, fetched one call ahead (thread t < 128: first node of
        // row t, else second node of row t - 128) and published to shared memory when the tile is staged -- a global
        // index load inside the copy loop would put an L2 round trip in front of every 16-byte copy
        int* idx_s = reinterpret_cast<int*>(smem + RM_OFF_BAR + 16);
        int my_idx = 0;
        auto load_pair_idx = [&](int tile) {
            const int rg = tile * 128 + (tid & 127);
            my_idx = (pair && rg < pg.n_rows) ? __ldg((tid < 128 ? in.i0 : in.i1) + rg) : 0;
        };
        auto stage_input = [&](int tile) {
            const int c4n = srow >> 2;                            // 16-byte chunks per staged row
            const int rows_here = min(128, pg.n_rows - tile * 128);
            if (pair) {
                idx_s[tid] = my_idx;
                group_sync(BAR_WORKERS, RM_NW);
            }
            for (int i = tid; i < 128 * c4n; i += RM_NW) {
                const int r = i / c4n, c4 = i - r * c4n;
                if (r >= rows_here) continue;
                const int rg = tile * 128 + r, c = 4 * c4;
                const float* src;
                if (pair) {
                    const int node = idx_s[(c < in.k_pad ? 0 : 128) + r];
                    src = in.p0 + (size_t)node * in.ld0 + (c < in.k_pad ? c : c - in.k_pad);
                } else {
                    src = c < in.w0 ? in.p0 + (size_t)rg * in.ld0 + c : in.p1 + (size_t)rg * in.ld1 + (c - in.w0);
                }
                cp_async16(stg + r * srow + ((c4 ^ (r & 7)) << 2), src);
            }
            cp_async_commit();
            if (pair) {
                group_sync(BAR_WORKERS, RM_NW);                   // idx_s may be overwritten by the next call
                if (tile + (int)gridDim.x < n_tiles) load_pair_idx(tile + gridDim.x);
            }
        };
        if (staged_in && (int)blockIdx.x < n_tiles) {
            load_pair_idx(blockIdx.x);
            stage_input(blockIdx.x);
        }
        // LIN0: raw feature row of this thread's row (f[7] = 1 multiplies the bias row), fetched one tile ahead
        const bool lin0 = in.mode == TC_IN_LIN0;
        const bool lin0_aux = !lin0 && pg.lin0.f != nullptr;
        float fcur[8], fnext[8];
        auto load_feat = [&](int tile, float (&f)[8]) {
            const int rg = tile * 128 + row;
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = 0.f;
            if (rg < pg.n_rows) {
                const size_t r = l0_idx ? (size_t)__ldg(l0_idx + rg) : (size_t)rg;
#pragma unroll
                for (int k = 0; k < 7; ++k)
                    if (k < l0_w) f[k] = __ldg(l0_f + r * l0_ld + k);
                f[7] = 1.f;
            }
        };
        // y[64] = act(W0 f + b0) for columns [128 half + 64 q, +64), then hi/lo -> the A operand
        auto lin0_compute = [&](const float (&f)[8], int half, float2 (&y)[32]) {
            const float* w = w0s + half * 128 + q * 64;
#pragma unroll
            for (int c = 0; c < 32; ++c) y[c] = make_float2(0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float2 fk = make_float2(f[k], f[k]);
#pragma unroll
                for (int c = 0; c < 32; ++c) y[c] = __ffma2_rn(*reinterpret_cast<const float2*>(w + k * 256 + 2 * c), fk, y[c]);
            }
            if (l0_act) {
                const float2 sl = make_float2(LEAKY, LEAKY);
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const float2 t = __fmul2_rn(y[c], sl);
                    y[c].x = fmaxf(y[c].x, t.x);
                    y[c].y = fmaxf(y[c].y, t.y);
                }
            }
        };
        auto lin0_store = [&](float2 (&y)[32]) {
#pragma unroll
            for (int c = 0; c < 32; c += 8) {
                float2 hi[8], lo[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) tc::split_tf32(y[c + j], hi[j], lo[j]);
                tc::tmem_st16(t_row + in.a_hi + q * 64 + 2 * c, hi);
                tc::tmem_st16(t_row + in.a_lo + q * 64 + 2 * c, lo);
            }
        };
        if ((lin0 || lin0_aux) && (int)blockIdx.x < n_tiles) load_feat(blockIdx.x, fcur);

        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int row_g = tile * 128 + row;
            const bool valid = row_g < pg.n_rows;
            // ---- input rows -> TMEM (hi | lo) ----
            {
                const bool split = (in.k_pad % (8 * RM_NQ)) == 0;
                const int cw = split ? in.k_pad / RM_NQ : in.k_pad;
                const int c0 = split ? q * cw : 0;
                if (lin0_aux && tile + (int)gridDim.x < n_tiles) load_feat(tile + gridDim.x, fnext);
                if (lin0) {
                    float2 y[32];
                    lin0_compute(fcur, 0, y);
                    lin0_store(y);
                    if (tile + (int)gridDim.x < n_tiles) load_feat(tile + gridDim.x, fnext);
                } else if (in.mode == TC_IN_BWD) {
                    // gradient rows -> dz = norm'(act'(g)) -> A operand (and the scratch buffer of the weight gradient)
                    const TcBwd& b = in.bwd;
                    const bool norm = b.scale != nullptr;
                    const float scale = norm ? __ldg(b.scale) : 1.f, shift = norm ? __ldg(b.shift) : 0.f;
                    const float sd = (norm && valid) ? __ldg(b.sd + row_g) : 0.f;
                    auto run = [&](auto tag) {
                        constexpr int CPT = decltype(tag)::value;
                        float2 g[CPT / 2], y[CPT / 2];
                        const int col0 = q * CPT;
                        const size_t r = in.i0 ? (size_t)__ldg(in.i0 + (valid ? row_g : 0)) : (size_t)row_g;
#pragma unroll
                        for (int c8 = 0; c8 < CPT / 8; ++c8) {
                            g[4 * c8] = g[4 * c8 + 1] = g[4 * c8 + 2] = g[4 * c8 + 3] = make_float2(0.f, 0.f);
                            if (valid) ldg256(in.p0 + r * in.ld0 + col0 + 8 * c8, g[4 * c8], g[4 * c8 + 1], g[4 * c8 + 2], g[4 * c8 + 3]);
                        }
                        load_y<CPT>(y, b, row_g, valid, col0);
                        float ps, pm;
                        bwd_transform<CPT>(g, y, b, in.w0, b.y != nullptr, scale, shift, sd, t_row + TC_XS_COL + 8, q, bar_id, ps, pm);
                        if (norm && b.g_scale != nullptr) pool_scalar_grads(ps, pm, valid, dacc + 2 * TC_MAX_STAGES, lane);
                        if (in.bwd_store != nullptr && valid) {
                            float* o = in.bwd_store + (size_t)row_g * in.k_pad + col0;
#pragma unroll
                            for (int c8 = 0; c8 < CPT / 8; ++c8) stg256(o + 8 * c8, g[4 * c8], g[4 * c8 + 1], g[4 * c8 + 2], g[4 * c8 + 3]);
                        }
#pragma unroll
                        for (int c = 0; c < CPT / 2; c += 8) {
                            float2 hi[8], lo[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) tc::split_tf32(g[c + j], hi[j], lo[j]);
                            tc::tmem_st16(t_row + in.a_hi + col0 + 2 * c, hi);
                            tc::tmem_st16(t_row + in.a_lo + col0 + 2 * c, lo);
                        }
                    };
                    if (in.k_pad == 128) run(std::integral_constant<int, 64>{});
                    else if (in.k_pad == 64) run(std::integral_constant<int, 32>{});
                    else run(std::integral_constant<int, 16>{});
                } else if (staged_in) {
                    cp_async_wait<0>();
                    group_sync(BAR_WORKERS, RM_NW);               // every thread's copies have landed
                    for (int c = c0; c < c0 + cw; c += 8) {
                        float v[8], hi[8], lo[8];
                        const float4 a0 = *reinterpret_cast<const float4*>(stg + row * srow + (((c >> 2) ^ (row & 7)) << 2));
                        const float4 a1 = *reinterpret_cast<const float4*>(stg + row * srow + ((((c >> 2) + 1) ^ (row & 7)) << 2));
                        v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
                        if (pair) {
                            const int cb = (c + in.k_pad) >> 2;
                            const float4 b0 = *reinterpret_cast<const float4*>(stg + row * srow + ((cb ^ (row & 7)) << 2));
                            const float4 b1 = *reinterpret_cast<const float4*>(stg + row * srow + (((cb + 1) ^ (row & 7)) << 2));
                            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (!valid) v[j] = 0.f;
                            tc::split_tf32(v[j], hi[j], lo[j]);
                        }
                        if (in.in_store != nullptr && valid)
                            stg256(in.in_store + (size_t)row_g * in.k_pad + c, make_float2(v[0], v[1]), make_float2(v[2], v[3]),
                                   make_float2(v[4], v[5]), make_float2(v[6], v[7]));
                        tc::tmem_st8(t_row + in.a_hi + c, hi);
                        tc::tmem_st8(t_row + in.a_lo + c, lo);
                    }
                    group_sync(BAR_WORKERS, RM_NW);               // staging consumed: the next tile's rows may land
                    if (tile + (int)gridDim.x < n_tiles) stage_input(tile + gridDim.x);
                } else if (split || q == 0) {
                    for (int c = c0; c < c0 + cw; c += 8) {
                        float v[8], hi[8], lo[8];
                        load_input_block(in, row_g, valid, c, v);
                        if (in.in_store != nullptr && valid)
                            stg256(in.in_store + (size_t)row_g * in.k_pad + c, make_float2(v[0], v[1]), make_float2(v[2], v[3]),
                                   make_float2(v[4], v[5]), make_float2(v[6], v[7]));
#pragma unroll
                        for (int j = 0; j < 8; ++j) tc::split_tf32(v[j], hi[j], lo[j]);
                        tc::tmem_st8(t_row + in.a_hi + c, hi);
                        tc::tmem_st8(t_row + in.a_lo + c, lo);
                    }
                }
                tc::tmem_wait_st();
                tc::tc_fence_before();
                bar_arrive(BAR_Y_READY, RM_NW + 32);
            }
            tick(0);
            for (int s = 0; s < pg.n_stages; ++s) {
                const TcEpi& e = pg.st[s].epi;
                if (e.refill) {     // second LIN0 half: computed while this stage's MMAs run, stored once they have read the first half
                    float2 y[32];
                    lin0_compute(fcur, 1, y);
                    tc::mbar_wait(d_ready, dphase);
                    dphase ^= 1u;
                    tc::tc_fence_after();
                    lin0_store(y);
                    tc::tmem_wait_st();
                    tc::tc_fence_before();
                    bar_arrive(BAR_Y_READY, RM_NW + 32);
                    tick(2 + 2 * s);
                    continue;
                }
                const float* cb = cst + s * RM_CST_LD;
                if (e.is_bwd) {
                    // the saved forward rows are requested BEFORE waiting for the accumulator: their L2 / HBM round trip
                    // overlaps the MMAs of this stage
                    const int cpt = e.n_cols / RM_NQ;
                    const BwdScalars bs = load_bwd_scalars(e.bwd, row_g, valid);
                    auto wait_d = [&]() {
                        tc::mbar_wait(d_ready, dphase);
                        dphase ^= 1u;
                        tc::tc_fence_after();
                        tick(2 + 2 * s);
                    };
                    if (e.bwd.lin0) {           // mask of the first encoder layer: recomputed from the raw features
                        float2 y[32];
                        lin0_compute(fcur, e.bwd.lin0_off >> 7, y);
                        if (e.bwd.y_store != nullptr && valid) {
                            float* o = e.bwd.y_store + (size_t)row_g * e.bwd.y_ld + e.bwd.lin0_off + q * 64;
#pragma unroll
                            for (int c8 = 0; c8 < 8; ++c8) stg256(o + 8 * c8, y[4 * c8], y[4 * c8 + 1], y[4 * c8 + 2], y[4 * c8 + 3]);
                        }
                        wait_d();
                        epilogue_bwd<64>(e, row_g, valid, q, t_row, bar_id, y, bs, dacc + 2 * s, lane);
                    } else if (cpt == 64) {
                        float2 y[32];
                        load_y<64>(y, e.bwd, row_g, valid, q * 64);
                        wait_d();
                        epilogue_bwd<64>(e, row_g, valid, q, t_row, bar_id, y, bs, dacc + 2 * s, lane);
                    } else if (cpt == 32) {
                        float2 y[16];
                        load_y<32>(y, e.bwd, row_g, valid, q * 32);
                        wait_d();
                        epilogue_bwd<32>(e, row_g, valid, q, t_row, bar_id, y, bs, dacc + 2 * s, lane);
                    } else {
                        float2 y[8];
                        load_y<16>(y, e.bwd, row_g, valid, q * 16);
                        wait_d();
                        epilogue_bwd<16>(e, row_g, valid, q, t_row, bar_id, y, bs, dacc + 2 * s, lane);
                    }
                    tc::tmem_wait_st();
                    tc::tc_fence_before();
                    if (s + 1 < pg.n_stages) bar_arrive(BAR_Y_READY, RM_NW + 32);
                    tick(3 + 2 * s);
                    continue;
                }
                tc::mbar_wait(d_ready, dphase);
                dphase ^= 1u;
                tc::tc_fence_after();
                tick(2 + 2 * s);
                if (e.scale != nullptr) {
                    const int cpt = e.n_cols / RM_NQ;
                    if (cpt == 64) epilogue_norm<64>(e, row_g, valid, q, t_row, bar_id, cb);
                    else if (cpt == 32) epilogue_norm<32>(e, row_g, valid, q, t_row, bar_id, cb);
                    else epilogue_norm<16>(e, row_g, valid, q, t_row, bar_id, cb);
                } else {
                    epilogue_plain(e, row_g, valid, q, t_row, cb);
                }
                tc::tmem_wait_st();
                tc::tc_fence_before();
                if (s + 1 < pg.n_stages) bar_arrive(BAR_Y_READY, RM_NW + 32);
                tick(3 + 2 * s);
            }
            group_sync(BAR_WORKERS, RM_NW);     // every worker is done with this tile's TMEM columns
            if (lin0 || lin0_aux) {
#pragma unroll
                for (int k = 0; k < 8; ++k) fcur[k] = fnext[k];
            }
            tick(1);
        }
        if (PROFILE && tid == 0 && prof != nullptr)
            for (int i = 0; i < 2 * TC_MAX_STAGES + 2; ++i) prof[blockIdx.x * (2 * TC_MAX_STAGES + 5) + i] = pt[i];
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
    // gradients of the channel_normalization scalars: one atomic per CTA, stage and scalar
    if ((int)blockIdx.x < n_tiles && tid <= pg.n_stages) {
        const TcBwd* b = nullptr;
        int slot_i = 0;
        if (tid < pg.n_stages) {
            if (pg.st[tid].epi.is_bwd) { b = &pg.st[tid].epi.bwd; slot_i = tid; }
        } else if (pg.in.mode == TC_IN_BWD) {
            b = &pg.in.bwd; slot_i = TC_MAX_STAGES;
        }
        if (b != nullptr && b->scale != nullptr && b->g_scale != nullptr) {
            atomicAdd(b->g_scale, (float)dacc[2 * slot_i]);
            atomicAdd(b->g_shift, (float)dacc[2 * slot_i + 1]);
        }
    }
}

int g_rowmlp_profile = 0;

int launch_rowmlp_tc(const TcProgram& pg, cudaStream_t stream) {
    if (pg.n_rows <= 0) return RGNN_OK;
    static PerDeviceOnce once;
    if (once.needed()) {
        RGNN_CHECK_CUDA(cudaFuncSetAttribute(rowmlp_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RM_SMEM));
        RGNN_CHECK_CUDA(cudaFuncSetAttribute(rowmlp_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RM_SMEM));
        once.mark();
    }
    const int n_tiles = (pg.n_rows + 127) / 128;
    const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
    if (g_rowmlp_profile) {   // developer aid (rgnn_set_option("debug", 8)): per-stage cycles of worker thread 0; synchronises
        constexpr int NP = 2 * TC_MAX_STAGES + 5;
        long long* prof = nullptr;
        RGNN_CHECK_CUDA(cudaMalloc(&prof, sizeof(long long) * NP * grid));
        rowmlp_tc_kernel<true><<<grid, RM_NT, RM_SMEM, stream>>>(pg, prof);
        RGNN_CHECK_CUDA(cudaStreamSynchronize(stream));
        long long* h = new long long[NP * grid];
        RGNN_CHECK_CUDA(cudaMemcpy(h, prof, sizeof(long long) * NP * grid, cudaMemcpyDeviceToHost));
        double tot[NP] = {0};
        for (int b = 0; b < grid; ++b) for (int i = 0; i < NP; ++i) tot[i] += (double)h[b * NP + i];
        fprintf(stderr, "[rowmlp_tc profile rows=%d stages=%d] cycles/tile: input=%.0f end=%.0f |", pg.n_rows, pg.n_stages, tot[0] / n_tiles, tot[1] / n_tiles);
        for (int s = 0; s < pg.n_stages; ++s) fprintf(stderr, " s%d wait=%.0f epi=%.0f", s, tot[2 + 2 * s] / n_tiles, tot[3 + 2 * s] / n_tiles);
        fprintf(stderr, " | mma warp: wait A=%.0f wait W=%.0f issue=%.0f\n", tot[NP - 3] / n_tiles, tot[NP - 2] / n_tiles, tot[NP - 1] / n_tiles);
        delete[] h;
        cudaFree(prof);
        return RGNN_OK;
    }
    rowmlp_tc_kernel<false><<<grid, RM_NT, RM_SMEM, stream>>>(pg, nullptr);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

// ---------------------------------------------------------------------------------------------
// weight packing for the chunk stream: rows [n0, n0+Nt) x columns [k0 + c*kc, ...) of W (row stride ldW) ->
// per K chunk:  hi (kc/4, Np, 4)  |  lo (kc/4, Np, 4),   zero padded to Np rows / Kp columns
// ---------------------------------------------------------------------------------------------
__global__ void pack_tc_kernel(const PackTcArgs a) { pack_tc_body(a, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x); }

// transpose: the packed layer is W^T (element (n, k) = W[k0 + k][n0 + n])
int pack_tc(const float* W, int ldW, int n0, int Nt, int nd0, int Np, int k0, int Kt, int Kp, int kc, bool pad_rows, float* dst,
            cudaStream_t stream, bool transpose) {
    const int n_loop = pad_rows ? Np - nd0 : Nt;
    const int blocks = (Kp * n_loop + 255) / 256;
    const PackTcArgs a{W, dst, ldW, n0, Nt, nd0, Np, k0, Kt, Kp, kc, n_loop, transpose ? 1 : 0};
    if (packq_push(a)) return RGNN_OK;          // inside rgnn_pack_detector: one table-driven launch for many images (rgnn_pack.cuh)
    pack_tc_kernel<<<blocks > 64 ? 64 : blocks, 256, 0, stream>>>(a);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

}  // namespace rgnn
