// Edge encoder (reference graph_feature_encoding, gnn_blocks.py:19-42, with the reference channel plan 7 -> 256 -> 128 -> 128 -> 64)
// as one fixed-shape kernel on the tensor cores with fp16-split operands (rgnn_f16.cuh).  Per edge 28 bytes are read and one
// 256-byte row is written -- directly in the pre-split format the message kernel (rgnn_mp_f16.cu) consumes -- and none of the
// 576 intermediate activations leaves the SM.
//
//   f (7 raw features of edge perm[k], x 16, zero padded to K = 16)                                A0  shared memory (SS MMA)
//   L0  256 outputs, no norm: evaluated in four 64-column chunks  D0[c & 1] = A0 W0[64c .. 64c+64)^T   -> E0: + b0, LeakyReLU
//   L1  D1 (128) += y0[chunk c] W1[:, 64c .. 64c+64)^T   (the chunks are the K blocks of L1)          -> E1: + b1, norm, act
//   L2  D2 (128) = y1 W2^T   (D2 reuses the two D0 buffers)                                          -> E2
//   L3  D3 (64)  = y2 W3^T   (D3 reuses D1's columns)                                                -> E3 -> emb row [hi 64 | lo 64]
//
// One CTA per SM, persistent; thread = TMEM lane = edge row; TWO tiles of 128 edges in flight, each owned by one group of 4
// epilogue warps (256 TMEM columns per group: D0[2] 2 x 64 | D1 128).  An epilogue rewrites its accumulator in place into the
// next A operand (per 32 fp32 columns: 16 packed hi | 16 packed lo columns), so E0 of chunk c + 1 runs while the MMAs of
// chunk c accumulate.  W0, W2, W3 (112 KB as fp16 hi | lo) are resident in shared memory; W1 (128 KB) does not fit beside
// them and is streamed L2 -> shared memory in four 32 KB K-blocks per tile PAIR (cp.async.bulk, two-slot mbarrier ring):
// the single MMA-issue lane follows a static schedule in which both groups consume a K-block back to back.
#include "rgnn_f16.cuh"
#include "rgnn_model.h"
#include "rgnn_tc_rows.cuh"

namespace rgnn {

struct EdgeEncArgs {
    const float* feat;          // (E, 7) raw edge features, reference order
    const int* perm;            // target-major row k <-> feature row perm[k] (nullptr: identity)
    int n_rows, n_feat;
    const uint32_t* w0;         // [hi | lo] images: K = 16 (7 valid) x N = 256
    const uint32_t* w1;         // four K blocks, each [hi | lo] of K = 64 x N = 128
    const uint32_t* w2;         // K = 128 x N = 128
    const uint32_t* w3;         // K = 128 x N = 64
    const float* b0; const float* b1; const float* b2; const float* b3;
    const float* s1; const float* m1; const float* s2; const float* m2; const float* s3; const float* m3;
    uint32_t* emb_hl;           // pre-split rows of 16 x the embedding, TILED (rgnn_f16.cuh: emb_tile_word)
    float* emb;                 // optional fp32 copy (E, 64) or nullptr
    float* save_y1; float* save_y2;                         // training: outputs of layers 1 / 2 (E, 128) and the three sigmas (E), nullable
    float* save_sd1; float* save_sd2; float* save_sd3;
    int passes;
    int prefetch;               // feature rows of the next tile are requested into L2 one tile ahead (their perm entry two tiles ahead)
};

namespace een {
constexpr int TM = 128, C0 = 256, C1 = 128, C2 = 128, C3 = 64, KF = 16;
constexpr int NTHREADS = 384;           // 2 x 4 epilogue warps, MMA warp, load warp (+ 2 idle)
constexpr int W0_WORDS = KF * C0;       // hi + lo: 2 x 16 x 256 x 2 B = 16 KB = 4096 words
constexpr int W1_BLK_WORDS = 64 * C1;   // one K block, hi + lo: 32 KB
constexpr int W2_WORDS = C1 * C2;       // 64 KB
constexpr int W3_WORDS = C2 * C3;       // 32 KB
constexpr int A0_WORDS = TM * KF;       // per group, hi + lo: 8 KB
constexpr int OFF_W0 = 0;
constexpr int OFF_W2 = OFF_W0 + W0_WORDS;
constexpr int OFF_W3 = OFF_W2 + W2_WORDS;
constexpr int OFF_RING = OFF_W3 + W3_WORDS;             // 2 slots
constexpr int OFF_A0 = OFF_RING + 2 * W1_BLK_WORDS;     // 2 groups
constexpr int OFF_CST = OFF_A0 + 2 * A0_WORDS;          // b0[256] b1[128] b2[128] b3[64] s1 m1 s2 m2 s3 m3
constexpr int CST_B0 = 0, CST_B1 = 256, CST_B2 = 384, CST_B3 = 512, CST_S = 576;
constexpr int OFF_BAR = OFF_CST + 584;                  // a_full[2][2] d0_full[2][2] d_full[2] full[2] empty[2]
constexpr int N_BARS = 14;
constexpr int OFF_SLOT = OFF_BAR + 2 * N_BARS;
constexpr int WORDS = OFF_SLOT + 2;
constexpr size_t SMEM = (size_t)WORDS * 4;
static_assert((OFF_BAR % 2) == 0 && (OFF_A0 % 4) == 0 && (OFF_RING % 4) == 0, "alignment");
static_assert(SMEM <= 227 * 1024, "shared memory budget");
enum { B_A_FULL = 0, B_D0_FULL = 4, B_D_FULL = 8, B_FULL = 10, B_EMPTY = 12 };
}  // namespace een

namespace tc {
__device__ __forceinline__ void mbar_expect_tx_e(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_e(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
}  // namespace tc

// norm + act + split of a 128- or 64-column accumulator that sits in TMEM at `dreg`, in place (A format), thread = row
template <int C>
__device__ __forceinline__ void enc_norm_epilogue(uint32_t dreg, const float* __restrict__ bias, bool has_norm, float gain, float shift, int np,
                                                  float* __restrict__ save_row = nullptr, float* __restrict__ save_sd = nullptr) {
    static_assert(C == 128 || C == 64, "chunks of 32 columns");
    const float2 us = make_float2(f16::D_UNSCALE, f16::D_UNSCALE);
    float k = f16::A_SCALE, sh = 0.f, mean = 0.f;
    if (has_norm) {
        RowStats st;
        st.init();
#pragma unroll 1
        for (int c = 0; c < C; c += 32) {
            float2 v[16];
            tc::tmem_ld16(dreg + c, v);
            tc::tmem_ld16(dreg + c + 16, v + 8);
            tc::tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __ffma2_rn(v[i], us, *reinterpret_cast<const float2*>(bias + c + 2 * i));
            st.add_chunk(v);
        }
        const float sd = st.sigma(C);
        if (save_sd != nullptr) *save_sd = sd;          // a training step keeps sigma and (below) the layer output for the backward
        k = f16::A_SCALE * gain * __frcp_rn(sd + NORM_EPS);
        sh = f16::A_SCALE * shift;
        mean = st.mean;
    }
    const float2 k2 = make_float2(k, k), sh2 = make_float2(sh, sh), sl = make_float2(LEAKY, LEAKY), nm = make_float2(-mean, -mean);
#pragma unroll 1
    for (int c = 0; c < C; c += 32) {
        float2 v[16];
        tc::tmem_ld16(dreg + c, v);
        tc::tmem_ld16(dreg + c + 16, v + 8);
        tc::tmem_wait_ld();
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float2 y = __ffma2_rn(v[i], us, *reinterpret_cast<const float2*>(bias + c + 2 * i));
            y = __ffma2_rn(__fadd2_rn(y, nm), k2, sh2);
            const float2 t = __fmul2_rn(y, sl);
            y.x = fmaxf(y.x, t.x);
            y.y = fmaxf(y.y, t.y);
            f16::split(y, hi[i], lo[i]);
            v[i] = y;
        }
        f16::tmem_st16u(dreg + c, hi);
        if (np != 1) f16::tmem_st16u(dreg + c + 16, lo);
        if (save_row != nullptr) {
            const float2 un = make_float2(1.f / f16::A_SCALE, 1.f / f16::A_SCALE);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                stg256(save_row + c + 8 * i, __fmul2_rn(v[4 * i], un), __fmul2_rn(v[4 * i + 1], un), __fmul2_rn(v[4 * i + 2], un), __fmul2_rn(v[4 * i + 3], un));
        }
    }
}

__global__ void __launch_bounds__(een::NTHREADS, 1) edge_enc_f16_kernel(const __grid_constant__ EdgeEncArgs a) {
    using namespace een;
    extern __shared__ __align__(1024) uint32_t smem_u[];
    float* smem_f = reinterpret_cast<float*>(smem_u);
    float* cst = smem_f + OFF_CST;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_u + OFF_BAR);
    uint32_t* slot = smem_u + OFF_SLOT;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = warp >> 2, w4 = warp & 3, row = tid & 127;
    const int G = (int)gridDim.x;
    const int n_tiles = (a.n_rows + TM - 1) / TM;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / G + 1 : 0;
    const int n_pairs = (my_tiles + 1) / 2;         // both groups run every pair (a missing tile has no valid rows)
    const int np = a.passes == 1 ? 1 : 3;

    // ---- one-time setup ----
    {
        auto copy = [&](uint32_t* dst, const uint32_t* src, int words) {
            const uint4* s4 = reinterpret_cast<const uint4*>(src);
            uint4* d4 = reinterpret_cast<uint4*>(dst);
            for (int i = tid; i < words / 4; i += NTHREADS) d4[i] = __ldg(s4 + i);
        };
        copy(smem_u + OFF_W0, a.w0, W0_WORDS);
        copy(smem_u + OFF_W2, a.w2, W2_WORDS);
        copy(smem_u + OFF_W3, a.w3, W3_WORDS);
        for (int i = tid; i < 2 * A0_WORDS; i += NTHREADS) smem_u[OFF_A0 + i] = 0u;      // K columns 8 .. 15 stay zero
        for (int i = tid; i < 584; i += NTHREADS) {
            float v = 0.f;
            if (i < CST_B1) v = a.b0 ? __ldg(a.b0 + i) : 0.f;
            else if (i < CST_B2) v = a.b1 ? __ldg(a.b1 + (i - CST_B1)) : 0.f;
            else if (i < CST_B3) v = a.b2 ? __ldg(a.b2 + (i - CST_B2)) : 0.f;
            else if (i < CST_S) v = a.b3 ? __ldg(a.b3 + (i - CST_B3)) : 0.f;
            else if (i == CST_S + 0) v = a.s1 ? __ldg(a.s1) : 1.f;
            else if (i == CST_S + 1) v = a.m1 ? __ldg(a.m1) : 0.f;
            else if (i == CST_S + 2) v = a.s2 ? __ldg(a.s2) : 1.f;
            else if (i == CST_S + 3) v = a.m2 ? __ldg(a.m2) : 0.f;
            else if (i == CST_S + 4) v = a.s3 ? __ldg(a.s3) : 1.f;
            else if (i == CST_S + 5) v = a.m3 ? __ldg(a.m3) : 0.f;
            cst[i] = v;
        }
    }
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&bars[B_A_FULL + 2 * i], 4);          // two barriers per group, used alternately: a group can be two
            tc::mbar_init(&bars[B_A_FULL + 2 * i + 1], 4);      // operands ahead of the MMA warp, never three
            tc::mbar_init(&bars[B_D0_FULL + 2 * i], 1);
            tc::mbar_init(&bars[B_D0_FULL + 2 * i + 1], 1);
            tc::mbar_init(&bars[B_D_FULL + i], 1);
            tc::mbar_init(&bars[B_FULL + i], 1);
            tc::mbar_init(&bars[B_EMPTY + i], 1);
        }
        tc::mbar_init_fence();
    }
    if (warp == 0) tc::tmem_alloc(slot, 512);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *slot;

    if (g < 2) {
        // =========================== epilogue groups ===========================
        const uint32_t t_base = tmem + ((uint32_t)w4 << 21) + (uint32_t)g * 256;
        const uint32_t d0r[2] = {t_base, t_base + 64}, d1r = t_base + 128;
        uint32_t* a0 = smem_u + OFF_A0 + g * A0_WORDS;          // [hi: 2 chunks x 128 rows x 16 B | lo: same]
        uint32_t n_d0[2] = {0, 0}, n_d = 0, n_arr = 0;         // completed phases of this group's barriers / operands handed over
        const bool n1 = a.s1 != nullptr, n2 = a.s2 != nullptr, n3 = a.s3 != nullptr;
        // The feature row of an edge is reached through perm: two dependent trips to HBM at the head of every tile of this group's
        // serial chain.  With `prefetch` the perm entries run two tiles ahead in registers and the feature row of the next tile is
        // requested into L2 while this tile is evaluated, so the head of a tile costs one L2 round trip.
        const bool pf = a.prefetch != 0 && a.perm != nullptr;
        auto perm_of = [&](int jj) -> int {
            const long long rq = ((long long)blockIdx.x + (long long)jj * G) * TM + row;
            return (jj < my_tiles && rq < a.n_rows) ? __ldg(a.perm + rq) : -1;
        };
        int p1 = -1, p2 = -1;               // perm entries of this group's next tile / the one after
        if (pf) { p1 = perm_of(g); p2 = perm_of(2 + g); }
        for (int t = 0; t < n_pairs; ++t) {
            const int j = 2 * t + g;
            const int tile = (int)blockIdx.x + j * G;
            const int r = tile * TM + row;
            const bool valid = j < my_tiles && r < a.n_rows;
            // ---- raw features -> A0 (x 16, fp16 hi | lo) ----
            {
                float f[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) f[k] = 0.f;
                const int p0 = p1;
                if (pf) {
                    p1 = p2;
                    if (p1 >= 0) {
                        const float* q = a.feat + (size_t)p1 * a.n_feat;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(q + a.n_feat - 1));      // a 28-byte row can straddle two lines
                    }
                    p2 = perm_of(j + 4);
                }
                if (valid) {
                    const size_t fr = pf ? (size_t)p0 : (a.perm != nullptr ? (size_t)__ldg(a.perm + r) : (size_t)r);
#pragma unroll
                    for (int k = 0; k < 7; ++k)
                        if (k < a.n_feat) f[k] = __ldg(a.feat + fr * a.n_feat + k) * f16::A_SCALE;
                }
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) f16::split(make_float2(f[2 * k], f[2 * k + 1]), hi[k], lo[k]);
                *reinterpret_cast<uint4*>(a0 + row * 4) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(a0 + A0_WORDS / 2 + row * 4) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                tc::fence_async_smem();         // generic-proxy stores -> visible to the MMA's operand reads
                tc::tc_fence_before();          // (and this tile's operand follows the previous tile's last TMEM reads)
                { warp_arrive(&bars[B_A_FULL + 2 * g + (n_arr & 1u)], lane); ++n_arr; }
            }
            // ---- L0 chunks: + b0, LeakyReLU (no norm), in place ----
            const float2 us = make_float2(f16::D_UNSCALE, f16::D_UNSCALE), sl = make_float2(LEAKY, LEAKY), s16 = make_float2(f16::A_SCALE, f16::A_SCALE);
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const int bf = c & 1;
                tc::mbar_wait(&bars[B_D0_FULL + 2 * g + bf], n_d0[bf] & 1u);
                ++n_d0[bf];
                tc::tc_fence_after();
                const float* b0 = cst + CST_B0 + 64 * c;
#pragma unroll
                for (int h = 0; h < 64; h += 32) {
                    float2 v[16];
                    tc::tmem_ld16(d0r[bf] + h, v);
                    tc::tmem_ld16(d0r[bf] + h + 16, v + 8);
                    tc::tmem_wait_ld();
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float2 y = __ffma2_rn(v[i], us, *reinterpret_cast<const float2*>(b0 + h + 2 * i));
                        const float2 tt = __fmul2_rn(y, sl);
                        y.x = fmaxf(y.x, tt.x);
                        y.y = fmaxf(y.y, tt.y);
                        f16::split(__fmul2_rn(y, s16), hi[i], lo[i]);
                    }
                    f16::tmem_st16u(d0r[bf] + h, hi);
                    if (np != 1) f16::tmem_st16u(d0r[bf] + h + 16, lo);
                }
                tc::tmem_wait_st();
                tc::tc_fence_before();
                { warp_arrive(&bars[B_A_FULL + 2 * g + (n_arr & 1u)], lane); ++n_arr; }
            }
            // ---- L1 ----
            tc::mbar_wait(&bars[B_D_FULL + g], n_d & 1u);
            ++n_d;
            tc::tc_fence_after();
            enc_norm_epilogue<C1>(d1r, cst + CST_B1, n1, cst[CST_S], cst[CST_S + 1], np,
                                  (valid && a.save_y1 != nullptr) ? a.save_y1 + (size_t)r * C1 : nullptr,
                                  (valid && a.save_sd1 != nullptr) ? a.save_sd1 + r : nullptr);
            tc::tmem_wait_st();
            tc::tc_fence_before();
            { warp_arrive(&bars[B_A_FULL + 2 * g + (n_arr & 1u)], lane); ++n_arr; }
            // ---- L2 (accumulator = the two D0 buffers) ----
            tc::mbar_wait(&bars[B_D_FULL + g], n_d & 1u);
            ++n_d;
            tc::tc_fence_after();
            enc_norm_epilogue<C2>(d0r[0], cst + CST_B2, n2, cst[CST_S + 2], cst[CST_S + 3], np,
                                  (valid && a.save_y2 != nullptr) ? a.save_y2 + (size_t)r * C2 : nullptr,
                                  (valid && a.save_sd2 != nullptr) ? a.save_sd2 + r : nullptr);
            tc::tmem_wait_st();
            tc::tc_fence_before();
            { warp_arrive(&bars[B_A_FULL + 2 * g + (n_arr & 1u)], lane); ++n_arr; }
            // ---- L3 -> emb row ----
            tc::mbar_wait(&bars[B_D_FULL + g], n_d & 1u);
            ++n_d;
            tc::tc_fence_after();
            {
                float2 va[16], vb[16];
                tc::tmem_ld16(d1r, va);
                tc::tmem_ld16(d1r + 16, va + 8);
                tc::tmem_ld16(d1r + 32, vb);
                tc::tmem_ld16(d1r + 48, vb + 8);
                tc::tmem_wait_ld();
                const float* b3 = cst + CST_B3;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    va[i] = __ffma2_rn(va[i], us, *reinterpret_cast<const float2*>(b3 + 2 * i));
                    vb[i] = __ffma2_rn(vb[i], us, *reinterpret_cast<const float2*>(b3 + 32 + 2 * i));
                }
                float k = 1.f, sh = 0.f, mean = 0.f;
                if (n3) {
                    RowStats st;
                    st.init();
                    st.add_chunk(va);
                    st.add_chunk(vb);
                    const float sd = st.sigma(C3);
                    if (valid && a.save_sd3 != nullptr) a.save_sd3[r] = sd;
                    k = cst[CST_S + 4] * __frcp_rn(sd + NORM_EPS);
                    sh = cst[CST_S + 5];
                    mean = st.mean;
                }
                const float2 k2 = make_float2(k, k), sh2 = make_float2(sh, sh), nm = make_float2(-mean, -mean);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    va[i] = __ffma2_rn(__fadd2_rn(va[i], nm), k2, sh2);
                    vb[i] = __ffma2_rn(__fadd2_rn(vb[i], nm), k2, sh2);
                    const float2 ta = __fmul2_rn(va[i], sl), tb = __fmul2_rn(vb[i], sl);
                    va[i].x = fmaxf(va[i].x, ta.x); va[i].y = fmaxf(va[i].y, ta.y);
                    vb[i].x = fmaxf(vb[i].x, tb.x); vb[i].y = fmaxf(vb[i].y, tb.y);
                }
                if (valid) {
                    if (a.emb != nullptr) {
                        float* o = a.emb + (size_t)r * C3;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            stg256(o + 8 * i, va[4 * i], va[4 * i + 1], va[4 * i + 2], va[4 * i + 3]);
                            stg256(o + 32 + 8 * i, vb[4 * i], vb[4 * i + 1], vb[4 * i + 2], vb[4 * i + 3]);
                        }
                    }
                    // tiled rows (rgnn_f16.cuh: emb_tile_word): 16 bytes per chunk of 8 channels, the warp's 32 rows write 512 contiguous bytes
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) f16::split(__fmul2_rn(va[i], s16), hi[i], lo[i]);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        *reinterpret_cast<uint4*>(a.emb_hl + emb_tile_word(r, 0, k)) = make_uint4(hi[4 * k], hi[4 * k + 1], hi[4 * k + 2], hi[4 * k + 3]);
                        *reinterpret_cast<uint4*>(a.emb_hl + emb_tile_word(r, 1, k)) = make_uint4(lo[4 * k], lo[4 * k + 1], lo[4 * k + 2], lo[4 * k + 3]);
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) f16::split(__fmul2_rn(vb[i], s16), hi[i], lo[i]);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        *reinterpret_cast<uint4*>(a.emb_hl + emb_tile_word(r, 0, 4 + k)) = make_uint4(hi[4 * k], hi[4 * k + 1], hi[4 * k + 2], hi[4 * k + 3]);
                        *reinterpret_cast<uint4*>(a.emb_hl + emb_tile_word(r, 1, 4 + k)) = make_uint4(lo[4 * k], lo[4 * k + 1], lo[4 * k + 2], lo[4 * k + 3]);
                    }
                }
            }
        }
    } else if (warp == 8) {
        // =========================== MMA issue warp (static schedule over the tile pair) ===========================
        if (lane == 0) {
            constexpr uint32_t ID64 = f16::idesc(TM, 64), ID128 = f16::idesc(TM, 128);
            const uint32_t sW0 = tc::smem_u32(smem_u + OFF_W0), sW2 = tc::smem_u32(smem_u + OFF_W2), sW3 = tc::smem_u32(smem_u + OFF_W3);
            const uint32_t sRing = tc::smem_u32(smem_u + OFF_RING);
            const uint32_t sA0[2] = {tc::smem_u32(smem_u + OFF_A0), tc::smem_u32(smem_u + OFF_A0 + A0_WORDS)};
            uint32_t n_a[2] = {0, 0}, n_full[2] = {0, 0};
            uint32_t blk = 0;           // running K-block counter of the W1 ring
            auto wait_a = [&](int x) {
                tc::mbar_wait(&bars[B_A_FULL + 2 * x + (n_a[x] & 1u)], (n_a[x] >> 1) & 1u);
                ++n_a[x];
                tc::tc_fence_after();
            };
            // L0 chunk c of group x: D0[c & 1] = A0 (smem) W0[64 c ..]^T, K = 16: one MMA per pass
            auto g0 = [&](int x, int c) {
                const uint32_t dcol = tmem + (uint32_t)x * 256 + (uint32_t)(c & 1) * 64;
                bool acc = false;
                for (int p = 0; p < np; ++p) {
                    const int pa = (np == 1) ? 0 : (p == 0 ? 1 : 0);
                    const int pb = (np == 1) ? 0 : (p == 1 ? 1 : 0);
                    const uint64_t ad = tc::smem_desc(sA0[x] + pa * (A0_WORDS * 2), TM * 16, 128);
                    const uint64_t bd = tc::smem_desc(sW0 + pb * (W0_WORDS * 2) + (uint32_t)c * 64 * 16, C0 * 16, 128);
                    f16::mma_ss(dcol, ad, bd, ID64, acc);
                    acc = true;
                }
                tc::mma_commit(&bars[B_D0_FULL + 2 * x + (c & 1)]);
            };
            for (int t = 0; t < n_pairs; ++t) {
                for (int x = 0; x < 2; ++x) {           // first two L0 chunks of both groups
                    wait_a(x);
                    g0(x, 0);
                    g0(x, 1);
                }
                for (int c = 0; c < 4; ++c, ++blk) {
                    const uint32_t sl = blk & 1u;
                    tc::mbar_wait(&bars[B_FULL + sl], n_full[sl] & 1u);      // K block c of W1 has landed
                    ++n_full[sl];
                    for (int x = 0; x < 2; ++x) {
                        wait_a(x);                       // y0 chunk c of group x is in D0[c & 1]
                        f16::gemm_ts<64, C1>(tmem + (uint32_t)x * 256 + 128, tmem + (uint32_t)x * 256 + (uint32_t)(c & 1) * 64,
                                            sRing + sl * (W1_BLK_WORDS * 4), ID128, c > 0, np);
                        if (c + 2 < 4) g0(x, c + 2);     // the buffer is free again: next chunk of L0
                        if (c == 3) tc::mma_commit(&bars[B_D_FULL + x]);
                    }
                    tc::mma_commit(&bars[B_EMPTY + sl]);                      // both groups' MMAs on this block are done -> refill
                }
                for (int x = 0; x < 2; ++x) {           // L2: A = y1 (D1 columns), D2 = the two D0 buffers
                    wait_a(x);
                    f16::gemm_ts<C1, C2>(tmem + (uint32_t)x * 256, tmem + (uint32_t)x * 256 + 128, sW2, ID128, false, np);
                    tc::mma_commit(&bars[B_D_FULL + x]);
                }
                for (int x = 0; x < 2; ++x) {           // L3: A = y2, D3 = first 64 columns of D1
                    wait_a(x);
                    f16::gemm_ts<C2, C3>(tmem + (uint32_t)x * 256 + 128, tmem + (uint32_t)x * 256, sW3, ID64, false, np);
                    tc::mma_commit(&bars[B_D_FULL + x]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 9) {
        // =========================== W1 load warp: four K blocks per tile pair through a two-slot ring ===========================
        if (lane == 0) {
            uint32_t n_empty[2] = {0, 0};
            uint32_t blk = 0;
            for (int t = 0; t < n_pairs; ++t) {
                for (int c = 0; c < 4; ++c, ++blk) {
                    const uint32_t sl = blk & 1u;
                    if (blk >= 2) {
                        tc::mbar_wait(&bars[B_EMPTY + sl], n_empty[sl] & 1u);
                        ++n_empty[sl];
                    }
                    tc::mbar_expect_tx_e(&bars[B_FULL + sl], W1_BLK_WORDS * 4);
                    tc::bulk_g2s_e(smem_u + OFF_RING + sl * W1_BLK_WORDS, a.w1 + (size_t)c * W1_BLK_WORDS, W1_BLK_WORDS * 4, &bars[B_FULL + sl]);
                }
            }
        }
        __syncwarp();
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int pack_f16_image(const float* W, int ldw, int K, int N, int n_valid, int k_valid, uint32_t* dst, cudaStream_t stream);
const float* f16_weights(const rgnn_linear& L);

static int g_enc_f16 = 1;
int edge_enc_f16_set_option(const char* name, int value) {
    if (strcmp(name, "f16_edge_enc") == 0 && (value == 0 || value == 1)) { g_enc_f16 = value; return 1; }
    return 0;
}
int edge_enc_f16_get_option(const char* name) { return strcmp(name, "f16_edge_enc") == 0 ? g_enc_f16 : -2; }

// reference channel plan of graph_feature_encoding for the edges: <= 7 -> 256 (no norm) -> 128 -> 128 -> 64
bool edge_enc_f16_supported(const rgnn_stack& s) {
    if (!g_enc_f16 || s.n != 4) return false;
    const rgnn_linear& L0 = s.layer[0];
    if (L0.in_features > 7 || L0.out_features != 256 || L0.norm_scale != nullptr || !L0.activation) return false;
    const int in[3] = {256, 128, 128}, out[3] = {128, 128, 64};
    for (int i = 0; i < 3; ++i) {
        const rgnn_linear& L = s.layer[i + 1];
        if (L.in_features != in[i] || L.out_features != out[i] || L.norm_scale == nullptr || !L.activation || L.weight_t == nullptr) return false;
    }
    return L0.weight_t != nullptr;
}

// layer 1's image is stored as four K blocks (each [hi | lo]): it OVERWRITES the single K = 256 image the generic packer
// (f16_pack_linear) left in the same place, so this runs after the stack has been packed

int edge_enc_f16_pack(const rgnn_stack& s, cudaStream_t stream) {
    if (!edge_enc_f16_supported(s)) return RGNN_OK;
    const rgnn_linear& L1 = s.layer[1];
    int rc = RGNN_OK;
    uint32_t* w1 = reinterpret_cast<uint32_t*>(const_cast<float*>(f16_weights(L1)));
    for (int c = 0; c < 4; ++c) {       // K block c: columns [64 c, 64 c + 64) of W1
        rc = pack_f16_image(L1.weight + 64 * c, L1.in_features, 64, een::C1, een::C1, 64, w1 + (size_t)c * een::W1_BLK_WORDS, stream);
        if (rc) return rc;
    }
    return RGNN_OK;
}

int run_edge_enc_f16(const rgnn_stack& s, const float* feat, const int* perm, int n_rows, uint32_t* emb_hl, float* emb, cudaStream_t stream,
                     const TcSave* save) {
    if (n_rows <= 0) return RGNN_OK;
    EdgeEncArgs a;
    memset(&a, 0, sizeof(a));
    a.feat = feat; a.perm = perm; a.n_rows = n_rows; a.n_feat = s.layer[0].in_features;
    a.w0 = reinterpret_cast<const uint32_t*>(f16_weights(s.layer[0]));
    a.w1 = reinterpret_cast<const uint32_t*>(f16_weights(s.layer[1]));
    a.w2 = reinterpret_cast<const uint32_t*>(f16_weights(s.layer[2]));
    a.w3 = reinterpret_cast<const uint32_t*>(f16_weights(s.layer[3]));
    a.b0 = s.layer[0].bias; a.b1 = s.layer[1].bias; a.b2 = s.layer[2].bias; a.b3 = s.layer[3].bias;
    a.s1 = s.layer[1].norm_scale; a.m1 = s.layer[1].norm_shift;
    a.s2 = s.layer[2].norm_scale; a.m2 = s.layer[2].norm_shift;
    a.s3 = s.layer[3].norm_scale; a.m3 = s.layer[3].norm_shift;
    a.emb_hl = emb_hl; a.emb = emb;
    if (save != nullptr) {
        a.save_y1 = save->y[1]; a.save_y2 = save->y[2];
        a.save_sd1 = save->sd[1]; a.save_sd2 = save->sd[2]; a.save_sd3 = save->sd[3];
    }
    a.passes = mp_f16_passes();
    a.prefetch = chain_f16_get_option("rows_prefetch");
    static PerDeviceOnce once;
    if (once.needed()) {
        RGNN_CHECK_CUDA(cudaFuncSetAttribute(edge_enc_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)een::SMEM));
        once.mark();
    }
    const int n_tiles = (n_rows + een::TM - 1) / een::TM;
    const int want = (n_tiles + 1) / 2;                 // a CTA works on tile pairs
    const int grid = want < sm_count() ? want : sm_count();
    edge_enc_f16_kernel<<<grid, een::NTHREADS, een::SMEM, stream>>>(a);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

}  // namespace rgnn
