// Message function + sum aggregation of one residual_graph_conv_block (reference gnn_blocks.py:96-113) on the 5th-generation
// tensor cores with fp16-split operands (rgnn_f16.cuh) and a ROLE PIPELINE over double-buffered tensor memory:
//
//   agg[t] = sum_{e: s->t} ffn2( ffn1( cat(x_t, x_s, emb_e) ) ),      z1 = emb_e W_e^T + P_t[tgt] + P_s[src]   (P hoisted per node)
//
// A CTA (one per SM, persistent, 896 threads) owns tiles of 128 target-major edges; thread = TMEM lane = edge row in every
// role, so no row statistic ever crosses a thread.  Roles work on DIFFERENT tiles at the same time:
//
//   F    (4 warps)  tile j   : emb rows (pre-split fp16 hi | lo, 256 B / edge) -> TMEM A operand;  4096 (P_t[tgt] + P_s[src]) -> D1[b].
//                              Every tile passes through this role serially, so nothing it needs may cost a memory round trip
//                              inside the tile: the emb row and the first half of the P_t row of tile j + 1 are requested while
//                              tile j's pre-load runs (128 registers of loads in flight across the loop edge: 200 registers)
//   MMA  (2 lanes in 2 warps): G1(j): D1[b] += emb W_e^T   |   G2(j): D2[b] = y1 W_2^T ; each lane BLOCKS on its own barrier sequence
//   E1   (2 x 4 warps, both groups on every tile) : D1[b] -> mean / unbiased std / affine / LeakyReLU -> y1 as fp16 hi | lo, in
//                              place over D1[b]; each group rewrites two of the four 32-column chunks
//   E2   (2 x 4 warps, group g owns the tiles of parity g) : D2[b] -> + b2 -> norm -> act -> message tile in shared memory ->
//                              segmented sum over equal consecutive targets -> agg (plain stores; atomics only for the <= 2 segments
//                              a tile boundary cuts, so the result is deterministic).  The one message tile is handed from group
//                              to group through the mbarrier B_STAGE
//   stagers (6 warps) tile j+2: the 128 P_s[src] rows (512 B each) -> shared memory, whole rows per warp instruction (cp.async)
//
// with b = j & 1.  Tensor memory (512 columns): D1[2] 2 x 128 | D2[2] 2 x 64 | emb[2] 2 x (32 hi + 32 lo).  The node half of
// msg.0's hoisted projection is PRE-LOADED into the accumulator (the MMAs of G1 accumulate onto it).
// Scales: the accumulators hold 4096 x the true values (rgnn_f16.cuh); E1's normalisation is scale invariant (sigma and
// eps scale along), E2 multiplies by 2^-12 inside the FFMA that adds the bias.
// Hand-over (mbarriers, phase = (j >> 1) & 1):
//   ps_full[b]  stager -> F      ps_free[b]  F -> stager
//   a_full[b]   F -> MMA         d1_full[b]  tcgen05.commit(G1) -> E1
//   y1_full[b]  E1 -> MMA        d2_full[b]  tcgen05.commit(G2) -> E2, and -> F (D1[b] / emb[b] of tile j may be reused by j+2)
//   d2_free[b]  E2 -> MMA (D2[b] has been read)      stage (one): E2 group -> the other E2 group (completion k = tile k is summed)
// Register budgets (setmaxnreg; 65536 / 896 -> 72 at launch): E1 64, E2 64, F 200, MMA + stagers 24.
// What the measurements of round 2 say about this kernel (profiles/README.md): it is bound by the SM's issue
// slots and load/store path, not by the tensor pipe (25 % active) or HBM; every role that was relieved (a second E2 group, more
// stagers, blocking MMA lanes) only paid once the F role stopped waiting for its own loads.
#include "rgnn_f16.cuh"
#include "rgnn_model.h"
#include "rgnn_pack.cuh"
#include "rgnn_tc_rows.cuh"
#include "rgnn_tile.cuh"

namespace rgnn {

struct MpF16Args {
    const uint32_t* emb;    // pre-split edge rows, values x 16, target-major, TILED (rgnn_f16.cuh: emb_tile_word)
    const float* P;         // (N, 2H) fp32: [x W_t^T + b1 | x W_s^T]
    const int* tgt;
    const int* src;
    const uint32_t* wpack;  // [W_e hi | W_e lo | W_2 hi | W_2 lo], fp16 x 256, chunk-major (rgnn_f16.cuh)
    float* agg;             // (N, CN), zero-initialised by the caller
    int n_edges;
    int passes;             // 3 = fp32 parity, 1 = plain fp16 operands
    int act1, act2;
    const float* s1;        // channel_normalization gain / shift of msg.0 (device scalars) or nullptr
    const float* m1;
    const float* b2;        // msg.1 bias (CN) or nullptr
    const float* s2;
    const float* m2;
    long long* prof;        // PROFILE builds
    int nstager;            // staging warps in use (2 .. NSTAGER)
};

namespace mpf {
constexpr int CE = 64, H = 128, CN = 64, TM = 128;
// warp groups (4 warps each): E1 x 2 (each owns half of the columns of every tile), E2 x 2 (+ segmented sum), F, aux x 2 (2 MMA issue warps + 6 stagers)
constexpr int NTHREADS = 896;
constexpr int WG_E2 = 2, WG_F = 4, WG_AUX = 5;         // warpgroups 0, 1: E1;      // E2 group g (warpgroup WG_E2 + g) owns the tiles of parity g
constexpr int NSTAGER = 6;
constexpr int W1_WORDS = CE * H / 2, W2_WORDS = H * CN / 2;         // 32-bit words per (hi or lo) image
constexpr int OFF_W = 0;                                            // words: W_e hi | lo | W_2 hi | lo
constexpr int OFF_PS = OFF_W + 2 * W1_WORDS + 2 * W2_WORDS;          // [2][TM][H] floats
constexpr int OFF_STAGE = OFF_PS + 2 * TM * H;                      // [TM][CN] floats, 16-byte chunks XOR-swizzled
constexpr int SEG = TM + 4;
constexpr int OFF_SEG = OFF_STAGE + TM * CN;                        // [2 groups][SEG] int2
constexpr int OFF_MASK = OFF_SEG + 2 * 2 * SEG;                     // [2 groups]([4] ballots | nseg | cut | cut_first | cut_last)
constexpr int OFF_BAR = OFF_MASK + 2 * 8;                           // 15 mbarriers (+ 1 pad)
constexpr int OFF_BIAS = OFF_BAR + 2 * 16;                          // msg.1 bias (CN floats; broadcast reads)
constexpr int OFF_SLOT = OFF_BIAS + CN;
constexpr int WORDS = OFF_SLOT + 2;
constexpr size_t SMEM = (size_t)WORDS * 4;
static_assert((OFF_BAR % 2) == 0 && (OFF_SEG % 2) == 0, "mbarrier / int2 alignment");
static_assert(SMEM <= 227 * 1024, "shared memory budget");
// tensor memory columns
constexpr uint32_t COL_D1 = 0, COL_D2 = 256, COL_EMB = 384;
// register budget per role (setmaxnreg; launch = 65536 / 896 -> 72).  setmaxnreg.inc only draws what .dec released.
constexpr int REG_LAUNCH = 72, REG_E1 = 64, REG_E2 = 64, REG_F = 200, REG_AUX = 24;
static_assert(256 * REG_E1 + 256 * REG_E2 + 128 * REG_F + 256 * REG_AUX <= NTHREADS * REG_LAUNCH, "setmaxnreg pool");
// named barriers
constexpr int BAR_E2 = 12 /* and 14 for group 1 */, BAR_E1 = 13;
enum { B_PS_FULL = 0, B_PS_FREE = 2, B_A_FULL = 4, B_D1_FULL = 6, B_Y1_FULL = 8, B_D2_FULL = 10, B_D2_FREE = 12, B_STAGE = 14 };
}  // namespace mpf

#define MPF_EV(k) do { if (PROFILE && a.prof != nullptr && blockIdx.x == 0 && j >= 16 && j < 32) a.prof[(size_t)gridDim.x * 28 * 4 + (j - 16) * 16 + (k)] = clock64(); } while (0)
template <bool PROFILE>
__global__ void __launch_bounds__(mpf::NTHREADS, 1) mp_edge_f16_kernel(const __grid_constant__ MpF16Args a) {
    using namespace mpf;
    extern __shared__ __align__(1024) uint32_t smem_u[];
    float* smem_f = reinterpret_cast<float*>(smem_u);
    uint32_t* wsm = smem_u + OFF_W;
    float* ps = smem_f + OFF_PS;
    float* stage = smem_f + OFF_STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_u + OFF_BAR);
    uint32_t* slot = smem_u + OFF_SLOT;
    float* bias_s = smem_f + OFF_BIAS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wg = warp >> 2, w4 = warp & 3;
    const int row = tid & 127;
    const int G = (int)gridDim.x;
    const int n_tiles = (a.n_edges + TM - 1) / TM;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / G + 1 : 0;
    const int np = a.passes == 1 ? 1 : 3;

    // ---- one-time setup: weight images -> shared memory, barriers, tensor memory ----
    {
        const uint4* g = reinterpret_cast<const uint4*>(a.wpack);
        uint4* s = reinterpret_cast<uint4*>(wsm);
        constexpr int N4 = (2 * W1_WORDS + 2 * W2_WORDS) / 4;
        for (int i = tid; i < N4; i += NTHREADS) s[i] = __ldg(g + i);
    }
    if (tid < CN) bias_s[tid] = a.b2 != nullptr ? __ldg(a.b2 + tid) : 0.f;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&bars[B_PS_FULL + i], 32 * a.nstager);    // every stager lane: cp.async.mbarrier.arrive.noinc when its copies land
            tc::mbar_init(&bars[B_PS_FREE + i], 4);
            tc::mbar_init(&bars[B_A_FULL + i], 4);
            tc::mbar_init(&bars[B_D1_FULL + i], 1);
            tc::mbar_init(&bars[B_Y1_FULL + i], 8);      // both E1 groups
            tc::mbar_init(&bars[B_D2_FULL + i], 1);
            tc::mbar_init(&bars[B_D2_FREE + i], 4);
        }
        tc::mbar_init(&bars[B_STAGE], 4);           // the segmented sum of a tile is complete: the message tile may be rewritten
        tc::mbar_init_fence();
    }
    if (warp == 0) tc::tmem_alloc(slot, 512);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *slot;
    const uint32_t t_row = tmem + ((uint32_t)w4 << 21);       // this warp's 32 TMEM lanes (lane field = bits [16, 32))

    if (wg < WG_E2) {
        // =========================== E1 (two groups): epilogue of GEMM1 ===========================
        // Both groups take every tile: each gathers the row statistics over all 128 columns (cheap: TMEM reads are ~65 cycles
        // per 128 columns and nothing crosses a thread) and then normalises / splits only ITS two 32-column chunks, so the
        // latency of this stage, which sits on the F -> GEMM1 -> E1 -> GEMM2 loop of a TMEM buffer, is halved.
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REG_E1));
        const int g = wg;
        const float eps_s = NORM_EPS * f16::D_SCALE;
        const bool norm1 = a.s1 != nullptr;
        const float s1v = norm1 ? __ldg(a.s1) : 1.f, m1v = norm1 ? __ldg(a.m1) : 0.f;
        long long pt[3] = {0, 0, 0}, tl = 0;
        if (PROFILE) tl = clock64();
        for (int j = 0; j < my_tiles; ++j) {
            const int b = j & 1;
            const uint32_t ph = (uint32_t)(j >> 1) & 1u;
            const uint32_t d1 = t_row + COL_D1 + (uint32_t)b * H;
            tc::mbar_wait(&bars[B_D1_FULL + b], ph);
            tc::tc_fence_after();
            if (PROFILE && lane == 0) { const long long n = clock64(); pt[0] += n - tl; tl = n; }
            if (tid == 0) MPF_EV(5);
            float k, sh, mean;
            if (norm1) {
                // sweep 1: statistics on the 4096-fold values (power-of-two scaling is exact; sigma and eps scale along).  Each group
                // takes ITS two 32-column chunks (the ones it rewrites in sweep 2); the two partial (mean, M2) pairs meet in four
                // columns of the emb operand's region -- dead since GEMM1 completed, rewritten by the F role only after GEMM2 of this
                // tile -- and are merged with the exact pairwise update.  The barrier is the one the in-place stores need anyway.
                RowStats st;
                st.init();
                {
                    float2 v[16];
                    tc::tmem_ld16(d1 + 32 * g, v);
                    tc::tmem_ld16(d1 + 32 * g + 16, v + 8);
                    tc::tmem_wait_ld();
                    st.add_chunk(v);
                    tc::tmem_ld16(d1 + 32 * g + 64, v);
                    tc::tmem_ld16(d1 + 32 * g + 80, v + 8);
                    tc::tmem_wait_ld();
                    st.add_chunk(v);
                }
                const uint32_t xc = t_row + COL_EMB + (uint32_t)b * 64;
                asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(xc + 2 * g), "f"(st.mean), "f"(st.m2) : "memory");
                tc::tmem_wait_st();
                if (PROFILE && lane == 0) { const long long n = clock64(); pt[1] += n - tl; tl = n; }
                tc::tc_fence_before();
                group_sync(BAR_E1, 256);           // the other group has read its columns: in-place stores may begin; partials are visible
                tc::tc_fence_after();
                float m0, q0s, m1, q1s;
                tc::tmem_ld4(xc, m0, q0s, m1, q1s);
                tc::tmem_wait_ld();
                const float dm = m1 - m0;
                mean = 0.5f * (m0 + m1);
                const float m2 = (q0s + q1s) + dm * dm * (float)(H / 4);      // n_a n_b / n = 64 * 64 / 128
                const float sd = __fsqrt_rn(m2 * (1.f / (float)(H - 1)));
                k = f16::A_SCALE * s1v * __frcp_rn(sd + eps_s);
                sh = f16::A_SCALE * m1v;
            } else {
                k = f16::A_SCALE * f16::D_UNSCALE;
                sh = 0.f;
                mean = 0.f;
                if (PROFILE && lane == 0) { const long long n = clock64(); pt[1] += n - tl; tl = n; }
                tc::tc_fence_before();
                group_sync(BAR_E1, 256);
                tc::tc_fence_after();
            }
            if (tid == 0) MPF_EV(6);
            const float2 k2 = make_float2(k, k), sh2 = make_float2(sh, sh), sl = make_float2(LEAKY, LEAKY), nm = make_float2(-mean, -mean);
            const bool act = a.act1 != 0;
            // sweep 2 (this group's chunks g and g + 2): y1 x 16 = act(k (z - mean) + sh) -> fp16 hi | lo, IN PLACE: the 32 fp32
            // columns [c, c + 32) become 16 packed hi columns [c, c + 16) and 16 packed lo columns [c + 16, c + 32) (the MMA
            // warp addresses the K steps of GEMM2 accordingly)
            long long q0 = 0, q1 = 0, q2 = 0;
#pragma unroll 1
            for (int cb = 32 * g; cb < H; cb += 64) {
                // one 32-column chunk in two 16-column halves (register budget): the first half's lo pairs wait in registers until
                // the second half's fp32 values have been read from the columns they overwrite
                long long ta = 0;
                if (PROFILE) ta = clock64();
                uint32_t hi0[8], lo0[8], hi1[8], lo1[8];
                auto half = [&](int c, uint32_t (&hi)[8], uint32_t (&lo)[8]) {
                    float2 v[8];
                    tc::tmem_ld16(d1 + c, v);
                    tc::tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float2 y = __ffma2_rn(__fadd2_rn(v[i], nm), k2, sh2);
                        if (act) {
                            const float2 t = __fmul2_rn(y, sl);
                            y.x = fmaxf(y.x, t.x);
                            y.y = fmaxf(y.y, t.y);
                        }
                        f16::split(y, hi[i], lo[i]);
                    }
                };
                half(cb, hi0, lo0);
                half(cb + 16, hi1, lo1);
                if (PROFILE) { const long long tb = clock64(); q1 += tb - ta; ta = tb; }
                f16::tmem_st8u(d1 + cb, hi0);
                f16::tmem_st8u(d1 + cb + 8, hi1);
                if (np != 1) {
                    f16::tmem_st8u(d1 + cb + 16, lo0);
                    f16::tmem_st8u(d1 + cb + 24, lo1);
                }
                if (PROFILE) { tc::tmem_wait_st(); const long long tb = clock64(); q2 += tb - ta; }
            }
            if (PROFILE && lane == 0 && a.prof != nullptr && warp == 0) {
                atomicAdd((unsigned long long*)&a.prof[(blockIdx.x * 28 + 21) * 4 + 0], (unsigned long long)q0);
                atomicAdd((unsigned long long*)&a.prof[(blockIdx.x * 28 + 21) * 4 + 1], (unsigned long long)q1);
                atomicAdd((unsigned long long*)&a.prof[(blockIdx.x * 28 + 21) * 4 + 2], (unsigned long long)q2);
            }
            tc::tmem_wait_st();
            tc::tc_fence_before();
            warp_arrive(&bars[B_Y1_FULL + b], lane);
            if (tid == 0) MPF_EV(7);
            if (PROFILE && lane == 0) { const long long n = clock64(); pt[2] += n - tl; tl = n; }
        }
        if (PROFILE && lane == 0 && a.prof != nullptr)
            for (int i = 0; i < 3; ++i) a.prof[(blockIdx.x * 28 + warp) * 4 + i] = pt[i];
    } else if (wg < WG_F) {
        // =========================== E2 (two groups): epilogue of GEMM2, message tile, segmented sum ===========================
        // Group g takes the tiles of parity g (= TMEM buffer g): this is the role with the most work per tile (epilogue + segmented
        // sum), so two groups alternate.  The epilogue arithmetic of tile j + 1 overlaps the segmented sum of tile j; the ONE message
        // tile in shared memory is handed from group to group through the mbarrier B_STAGE (completion k = the sum of tile k is done).
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REG_E2));
        const int g = wg - WG_E2;
        const int bar_e2 = BAR_E2 + 2 * g;
        int2* seg_s = reinterpret_cast<int2*>(smem_u + OFF_SEG) + g * SEG;
        unsigned* mask_s = smem_u + OFF_MASK + 8 * g;
        int* nseg_s = reinterpret_cast<int*>(mask_s + 4);
        int* cut_s = nseg_s + 1;
        int* cut_first_s = nseg_s + 2;
        int* cut_last_s = nseg_s + 3;
        long long pt[3] = {0, 0, 0}, tl = 0;
        if (PROFILE) tl = clock64();
        const bool norm2 = a.s2 != nullptr;
        const float s2v = norm2 ? __ldg(a.s2) : 1.f, m2sv = norm2 ? __ldg(a.m2) : 0.f;
        int t_my = -1, t_first = -2, t_last = -3;
        auto load_idx = [&](int j) {
            const int tile = (int)blockIdx.x + j * G;
            const int e = tile * TM + row;
            const bool v = j < my_tiles && e < a.n_edges;
            t_my = v ? __ldg(a.tgt + e) : -1;
            t_first = (lane == 0 && v && e > 0) ? __ldg(a.tgt + e - 1) : -2;          // target of the edge before this warp's rows
            t_last = (row == TM - 1 && v && e + 1 < a.n_edges) ? __ldg(a.tgt + e + 1) : -3;
        };
        load_idx(g);
        for (int j = g; j < my_tiles; j += 2) {
            const int b = g;
            const uint32_t ph = (uint32_t)(j >> 1) & 1u;
            const int tile = (int)blockIdx.x + j * G;
            const int nvalid = min(TM, a.n_edges - tile * TM);
            // ---- segment heads of this tile (needs only the indices): ballots -> table, while GEMM2 runs ----
            const int t = t_my;
            {
                const bool v = row < nvalid;
                int prev = __shfl_up_sync(0xffffffffu, t, 1);
                if (lane == 0) prev = t_first;
                const unsigned m = __ballot_sync(0xffffffffu, v && (row == 0 || t != prev));
                if (lane == 0) mask_s[w4] = m;
                if (row == 0) cut_first_s[0] = (t_first == t) ? 1 : 0;
                if (row == nvalid - 1) cut_last_s[0] = (row == TM - 1 && t_last == t) ? 2 : 0;
            }
            group_sync(bar_e2, 128);           // also: this group's previous segmented sum is complete (its table is reusable)
            {
                const unsigned m = mask_s[w4];
                int base = 0;
                for (int w = 0; w < w4; ++w) base += __popc(mask_s[w]);
                if ((m >> lane) & 1u) seg_s[base + __popc(m & ((1u << lane) - 1u))] = make_int2(row, t);
                if (row == TM - 1) {
                    const int n = base + __popc(m);
                    seg_s[n] = make_int2(nvalid, -1);
                    nseg_s[0] = n;
                    cut_s[0] = cut_first_s[0] | cut_last_s[0];
                }
            }
            load_idx(j + 2);
            // ---- D2[b] -> registers ----
            tc::mbar_wait(&bars[B_D2_FULL + b], ph);
            tc::tc_fence_after();
            if (PROFILE && lane == 0) { const long long n = clock64(); pt[0] += n - tl; tl = n; }
            if (w4 == 0 && lane == 0) MPF_EV(10);
            // two sweeps over D2[b] in 32-column chunks (a TMEM read costs ~65 cycles; holding all 64 columns would not fit the
            // register budget of two E2 groups): statistics, then normalise / activate / store.  Chunk 1 is still in registers
            // when the second sweep starts, so only chunk 0 is read (and biased) twice.
            const uint32_t d2 = t_row + COL_D2 + (uint32_t)b * CN;
            const float2 us = make_float2(f16::D_UNSCALE, f16::D_UNSCALE);
            float2 v[16];
            auto load_chunk = [&](int h) {
                tc::tmem_ld16(d2 + 32 * h, v);
                tc::tmem_ld16(d2 + 32 * h + 16, v + 8);
                tc::tmem_wait_ld();
#pragma unroll
                for (int c = 0; c < 16; ++c) v[c] = __ffma2_rn(v[c], us, *reinterpret_cast<const float2*>(bias_s + 32 * h + 2 * c));
            };
            float k = 1.f, sh = 0.f, mean = 0.f;
            if (norm2) {
                RowStats st;
                st.init();
                load_chunk(0);
                st.add_chunk(v);
                load_chunk(1);
                st.add_chunk(v);
                k = s2v * __frcp_rn(st.sigma(CN) + NORM_EPS);
                sh = m2sv;
                mean = st.mean;
            } else {
                load_chunk(1);
            }
            const float2 k2 = make_float2(k, k), sh2 = make_float2(sh, sh), sl = make_float2(LEAKY, LEAKY), nm = make_float2(-mean, -mean);
            const bool act = a.act2 != 0;
            auto finish = [&]() {
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    v[c] = __ffma2_rn(__fadd2_rn(v[c], nm), k2, sh2);
                    if (act) {
                        const float2 t2 = __fmul2_rn(v[c], sl);
                        v[c].x = fmaxf(v[c].x, t2.x); v[c].y = fmaxf(v[c].y, t2.y);
                    }
                }
            };
            auto put = [&](int h) {
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4)
                    *reinterpret_cast<float4*>(stage + row * CN + (((c4 + 8 * h) ^ (row & 7)) << 2)) =
                        make_float4(v[2 * c4].x, v[2 * c4].y, v[2 * c4 + 1].x, v[2 * c4 + 1].y);
            };
            finish();
            if (j > 0) tc::mbar_wait(&bars[B_STAGE], (uint32_t)(j - 1) & 1u);      // the other group has summed tile j - 1
            if (w4 == 0 && lane == 0) MPF_EV(11);
            put(1);
            load_chunk(0);
            tc::tc_fence_before();
            warp_arrive(&bars[B_D2_FREE + b], lane);          // GEMM2 of tile j + 2 may overwrite D2[b]
            finish();
            put(0);
            if (PROFILE && lane == 0) { const long long n = clock64(); pt[1] += n - tl; tl = n; }
            if (w4 == 0 && lane == 0) MPF_EV(12);
            group_sync(bar_e2, 128);           // message tile and segment table complete
            // ---- segmented sum over equal consecutive targets: 16 threads (one float4 of columns each) per segment.  Interior
            // segments are whole CSR rows (edges are target-major): plain stores in source-ascending order like the reference's
            // index_add_; only the first / last segment of a tile can be cut by its boundary and uses atomicAdd (<= 2 partials
            // each, so the result is deterministic)
            {
                constexpr int RB = 8, NGRP = 8;
                const int s0 = row >> 4, c4 = row & 15;
                const int nseg = nseg_s[0];
                const int cut = cut_s[0];          // bit 0: first segment continues from the previous tile; bit 1: last one continues
                for (int sI = s0; sI < nseg; sI += NGRP) {
                    const int2 a0 = seg_s[sI], a1 = seg_s[sI + 1];
                    const int rs = a0.x, re = a1.x;
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int r0 = rs; r0 < re; r0 += RB) {
                        float4 v[RB];
#pragma unroll
                        for (int i = 0; i < RB; ++i) {
                            const int r = r0 + i;
                            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (r < re) v[i] = *reinterpret_cast<const float4*>(stage + r * CN + ((c4 ^ (r & 7)) << 2));
                        }
#pragma unroll
                        for (int i = 0; i < RB; ++i) { acc.x += v[i].x; acc.y += v[i].y; acc.z += v[i].z; acc.w += v[i].w; }
                    }
                    const bool whole = !((sI == 0 && (cut & 1)) || (sI == nseg - 1 && (cut & 2)));
                    float* o = a.agg + (size_t)a0.y * CN + 4 * c4;
                    if (whole) {
                        *reinterpret_cast<float4*>(o) = acc;
                    } else {
                        atomicAdd(o, acc.x); atomicAdd(o + 1, acc.y); atomicAdd(o + 2, acc.z); atomicAdd(o + 3, acc.w);
                    }
                }
            }
            warp_arrive(&bars[B_STAGE], lane);
            if (w4 == 0 && lane == 0) MPF_EV(13);
            if (PROFILE && lane == 0) { const long long n = clock64(); pt[2] += n - tl; tl = n; }
        }
        if (PROFILE && lane == 0 && a.prof != nullptr)
            for (int i = 0; i < 3; ++i) a.prof[(blockIdx.x * 28 + warp) * 4 + i] = pt[i];
    } else if (wg == WG_F) {
        // =========================== F: A operand and accumulator pre-load ===========================
        // This role waits most of the time for a TMEM buffer to come back (GEMM2 of tile j - 2), so everything it needs from
        // global memory for tile j is requested BEFORE that wait and sits in registers across it: three of the four
        // 32-column chunks of P_t[target] and the hi half of the emb row (the L2 round trip of a chunk is ~1 000 cycles here).
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REG_F));
        // This role is the one every tile passes through serially, so nothing it needs may cost a memory round trip inside the
        // tile: the emb row (hi | lo, 64 registers) of tile j + 1 is requested while tile j's accumulator pre-load runs, and the
        // P_t[target] row of tile j + 1 is brought to L1 one tile ahead (consecutive rows share their target: ~12 rows of 512 B
        // per tile; the streaming emb loads do not allocate in L1), so the pre-load reads it in 16-column pieces with L1 latency.
        long long pt[3] = {0, 0, 0}, tl = 0;
        if (PROFILE) tl = clock64();
        auto tgt_of = [&](int jj) {
            const int e = ((int)blockIdx.x + jj * G) * TM + row;
            return (jj < my_tiles && e < a.n_edges) ? __ldg(a.tgt + e) : -1;
        };
        auto prefetch_pt = [&](int t) {      // -> L2: the row is read into registers at the end of the previous tile
            if (t >= 0) {
                const float* q = a.P + (size_t)t * (2 * H);
#pragma unroll
                for (int i = 0; i < 4; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + 32 * i));
            }
        };
        uint32_t eh[32], el[32];
        // Rows past the end of the edge list (last tile) are not zeroed anywhere in this role: every row of the tile is independent
        // through both GEMMs and both epilogues, and E2 never sums rows >= nvalid.  The loads stay in bounds: the emb buffer holds
        // whole tiles, and an invalid row reads P row 0.
        auto load_emb = [&](int jj) {
            // (unconditional: past the CTA's last tile the last tile is simply read again; a predicated load would keep the
            // arrays in local memory)
            const long long e = ((long long)blockIdx.x + (long long)(jj < my_tiles ? jj : my_tiles - 1) * G) * TM + row;
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
                const uint4 vh = ldg128u_na(a.emb + emb_tile_word(e, 0, c8));
                eh[4 * c8] = vh.x; eh[4 * c8 + 1] = vh.y; eh[4 * c8 + 2] = vh.z; eh[4 * c8 + 3] = vh.w;
            }
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
                const uint4 vl = ldg128u_na(a.emb + emb_tile_word(e, 1, c8));
                el[4 * c8] = vl.x; el[4 * c8 + 1] = vl.y; el[4 * c8 + 2] = vl.z; el[4 * c8 + 3] = vl.w;
            }
        };
        int t_cur = tgt_of(0), t_nx = tgt_of(1);
        load_emb(0);
        float2 p0[16], p1[16];
        auto load_pt = [&](float2 (&v)[16], const float* Pt, int c) {
#pragma unroll
            for (int i = 0; i < 4; ++i) ldg256(Pt + c + 8 * i, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        };
        {
            const float* Pt0 = a.P + (size_t)(t_cur >= 0 ? t_cur : 0) * (2 * H);
            load_pt(p0, Pt0, 0);
            load_pt(p1, Pt0, 32);
        }
        const float2 sc = make_float2(f16::D_SCALE, f16::D_SCALE);
        for (int j = 0; j < my_tiles; ++j) {
            const int b = j & 1;
            const uint32_t ph = (uint32_t)(j >> 1) & 1u;
            const int tile = (int)blockIdx.x + j * G;
            const int e = tile * TM + row;
            const bool valid = e < a.n_edges;
            const int t = t_cur;
            const float* Pt = a.P + (size_t)(valid && t >= 0 ? t : 0) * (2 * H);
            t_cur = t_nx;
            prefetch_pt(t_cur);                 // tile j + 1's P_t row -> L2
            t_nx = tgt_of(j + 2);
            {   // the emb rows of tile j + 2 -> L2 (they stream from DRAM); tile j + 1's are requested below
                const long long e2 = ((long long)blockIdx.x + (long long)(j + 2) * G) * TM + row;
                if (j + 2 < my_tiles && e2 < a.n_edges) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(a.emb + e2 * 64));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(a.emb + e2 * 64 + 32));
                }
            }
            if (j >= 2) {       // D1[b] / emb[b] were last used by tile j - 2: free once its GEMM2 has completed
                tc::mbar_wait(&bars[B_D2_FULL + b], ph ^ 1u);
                tc::tc_fence_after();
            }
            if (PROFILE && lane == 0) { const long long n = clock64(); pt[0] += n - tl; tl = n; }
            if (tid == WG_F * 128) MPF_EV(0);
            const uint32_t c_emb = t_row + COL_EMB + (uint32_t)b * 64;
            f16::tmem_st16u(c_emb, eh);
            f16::tmem_st16u(c_emb + 16, eh + 16);
            if (np != 1) {
                f16::tmem_st16u(c_emb + 32, el);
                f16::tmem_st16u(c_emb + 48, el + 16);
            }
            load_emb(j + 1);
            // accumulator pre-load: 4096 (P_t[target] + P_s[source]); P_s from the staged rows
            const float* Ps = ps + (size_t)b * TM * H + row * H;
            const uint32_t d1 = t_row + COL_D1 + (uint32_t)b * H;
            tc::mbar_wait(&bars[B_PS_FULL + b], ph);
            if (PROFILE && lane == 0) { const long long n = clock64(); pt[1] += n - tl; tl = n; }
            if (tid == WG_F * 128) MPF_EV(1);
            auto emit = [&](float2 (&v)[16], int c) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 s4 = *reinterpret_cast<const float4*>(Ps + ((((c >> 2) + i) ^ (row & 7)) << 2));
                    v[2 * i] = __ffma2_rn(v[2 * i], sc, __fmul2_rn(make_float2(s4.x, s4.y), sc));
                    v[2 * i + 1] = __ffma2_rn(v[2 * i + 1], sc, __fmul2_rn(make_float2(s4.z, s4.w), sc));
                }
                tc::tmem_st16(d1 + c, v);
                tc::tmem_st16(d1 + c + 16, v + 8);
            };
            emit(p0, 0);
            load_pt(p0, Pt, 64);
            emit(p1, 32);
            load_pt(p1, Pt, 96);
            emit(p0, 64);
            emit(p1, 96);
            tc::tmem_wait_st();
            tc::tc_fence_before();
            warp_arrive(&bars[B_PS_FREE + b], lane);          // staged rows consumed
            warp_arrive(&bars[B_A_FULL + b], lane);           // -> GEMM1 of this tile
            {   // the first half of tile j + 1's P_t row (L2: prefetched at the top of this tile)
                const float* Pn = a.P + (size_t)(t_cur >= 0 ? t_cur : 0) * (2 * H);
                load_pt(p0, Pn, 0);
                load_pt(p1, Pn, 32);
            }
            if (tid == WG_F * 128) MPF_EV(2);
            if (PROFILE && lane == 0) { const long long n = clock64(); pt[2] += n - tl; tl = n; }
        }
        if (PROFILE && lane == 0 && a.prof != nullptr)
            for (int i = 0; i < 3; ++i) a.prof[(blockIdx.x * 28 + warp) * 4 + i] = pt[i];
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REG_AUX));
        if (warp == 4 * WG_AUX) {
            // =========================== MMA issue warp 1: GEMM1 ===========================
            // Two issue warps, one per GEMM, each BLOCKING on its own barrier sequence (mbarrier.try_wait suspends the warp in
            // hardware).  One lane polling both sequences executed a third of the kernel's instructions and took its issue slots
            // from the row-owning warps of scheduler 0, which every role's barriers then waited for.
            if (lane == 0) {
                constexpr uint32_t IDESC1 = f16::idesc(TM, H);
                constexpr uint32_t LBO_W1 = H * 16, SBO = 128;
                const uint32_t sW1 = tc::smem_u32(wsm);
                long long idle = 0;
                for (int j = 0; j < my_tiles; ++j) {
                    const int b = j & 1;
                    long long t0 = 0;
                    if (PROFILE) t0 = clock64();
                    tc::mbar_wait(&bars[B_A_FULL + b], (uint32_t)(j >> 1) & 1u);      // F waited for GEMM2(j - 2) before it filled D1[b] / emb[b]
                    if (PROFILE) idle += clock64() - t0;
                    MPF_EV(3);
                    tc::tc_fence_after();
                    const uint32_t dcol = tmem + COL_D1 + (uint32_t)b * H;
                    for (int p = 0; p < np; ++p) {      // small terms first: lo*hi, hi*lo, then hi*hi
                        const int pa = (np == 1) ? 0 : (p == 0 ? 1 : 0);
                        const int pb = (np == 1) ? 0 : (p == 1 ? 1 : 0);
                        const uint32_t acol = tmem + COL_EMB + (uint32_t)b * 64 + (pa ? 32u : 0u);
                        const uint64_t bd0 = tc::smem_desc(sW1 + pb * (W1_WORDS * 4), LBO_W1, SBO);
#pragma unroll
                        for (int ks = 0; ks < CE / 16; ++ks)
                            f16::mma_ts(dcol, acol + ks * 8, bd0 + (uint64_t)((ks * 2 * LBO_W1) >> 4), IDESC1, true);
                    }
                    tc::mma_commit(&bars[B_D1_FULL + b]);
                    MPF_EV(4);
                }
                if (PROFILE && a.prof != nullptr) a.prof[(blockIdx.x * 28 + warp) * 4] = idle;
            }
            __syncwarp();
        } else if (warp == 4 * WG_AUX + 1) {
            // =========================== MMA issue warp 2: GEMM2 ===========================
            if (lane == 0) {
                constexpr uint32_t IDESC2 = f16::idesc(TM, CN);
                constexpr uint32_t LBO_W2 = CN * 16, SBO = 128;
                const uint32_t sW2 = tc::smem_u32(wsm + 2 * W1_WORDS);
                long long idle = 0;
                for (int j = 0; j < my_tiles; ++j) {
                    const int b = j & 1;
                    const uint32_t ph = (uint32_t)(j >> 1) & 1u;
                    long long t0 = 0;
                    if (PROFILE) t0 = clock64();
                    if (j >= 2) tc::mbar_wait(&bars[B_D2_FREE + b], ph ^ 1u);          // E2 has read D2[b] of tile j - 2
                    tc::mbar_wait(&bars[B_Y1_FULL + b], ph);
                    if (PROFILE) idle += clock64() - t0;
                    MPF_EV(8);
                    tc::tc_fence_after();
                    const uint32_t dcol = tmem + COL_D2 + (uint32_t)b * CN;
                    bool acc = false;
                    for (int p = 0; p < np; ++p) {
                        const int pa = (np == 1) ? 0 : (p == 0 ? 1 : 0);
                        const int pb = (np == 1) ? 0 : (p == 1 ? 1 : 0);
                        // y1 as E1 leaves it: per 32-column chunk [hi: K steps 2i, 2i+1 | lo: the same]
                        const uint32_t acol = tmem + COL_D1 + (uint32_t)b * H + (pa ? 16u : 0u);
                        const uint64_t bd0 = tc::smem_desc(sW2 + pb * (W2_WORDS * 4), LBO_W2, SBO);
#pragma unroll
                        for (int ks = 0; ks < H / 16; ++ks) {
                            f16::mma_ts(dcol, acol + (ks >> 1) * 32 + (ks & 1) * 8, bd0 + (uint64_t)((ks * 2 * LBO_W2) >> 4), IDESC2, acc);
                            acc = true;
                        }
                    }
                    tc::mma_commit(&bars[B_D2_FULL + b]);
                    MPF_EV(9);
                }
                if (PROFILE && a.prof != nullptr) a.prof[(blockIdx.x * 28 + 20) * 4 + 1] = idle;
            }
            __syncwarp();
        } else {
            // =========================== staging warps: P_s[source] rows, two tiles ahead ===========================
            // warp s copies rows s, s + 3, s + 6, ... of every tile: whole 512-byte rows per warp instruction (cp.async, 16 B per
            // lane; ONE warp needs ~7 300 cycles per tile for the 128 rows, four need ~2 000: tools/micro/bench_tmem.cu).  Row r
            // is stored with its 16-byte chunks XOR-swizzled by (r & 7) so that the row-owning F threads read it conflict free.
            const int sI = warp - 4 * WG_AUX - 2;
            const int NS = a.nstager;
            const int RPS = (TM + NS - 1) / NS;       // rows per stager
            int id0 = -1, id1 = -1;
            auto load_ids = [&](int j) {
                const int e0 = ((int)blockIdx.x + j * G) * TM;
                const int r0 = sI + NS * lane, r1 = sI + NS * (lane + 32);
                id0 = (j < my_tiles && r0 < TM && e0 + r0 < a.n_edges) ? __ldg(a.src + e0 + r0) : -1;
                id1 = (j < my_tiles && r1 < TM && e0 + r1 < a.n_edges) ? __ldg(a.src + e0 + r1) : -1;
            };
            const int n_my = sI < NS ? my_tiles : 0;        // surplus stagers idle
            load_ids(0);
            long long pt[2] = {0, 0}, tl = 0;
            if (PROFILE) tl = clock64();
            for (int j = 0; j < n_my; ++j) {
                const int b = j & 1;
                if (j >= 2) tc::mbar_wait(&bars[B_PS_FREE + b], ((uint32_t)(j >> 1) & 1u) ^ 1u);    // F has consumed tile j - 2
                if (PROFILE && lane == 0) { const long long n = clock64(); pt[0] += n - tl; tl = n; }
                float* dst = ps + (size_t)b * TM * H;
                const int my0 = id0, my1 = id1;
                load_ids(j + 1);
#pragma unroll 4
                for (int i = 0; i < RPS; ++i) {
                    const int r = sI + NS * i;
                    const int sn = __shfl_sync(0xffffffffu, i < 32 ? my0 : my1, i & 31);
                    if (sn >= 0) cp_async16(dst + r * H + ((lane ^ (r & 7)) << 2), a.P + (size_t)sn * (2 * H) + H + 4 * lane);
                }
                // the barrier completes when the copies of all stager lanes have landed; nobody waits here
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc::smem_u32(&bars[B_PS_FULL + b])) : "memory");
                if (PROFILE && lane == 0) { const long long n = clock64(); pt[1] += n - tl; tl = n; }
            }
            cp_async_wait<0>();
            if (PROFILE && lane == 0 && a.prof != nullptr && sI == 0)
                for (int i = 0; i < 2; ++i) a.prof[(blockIdx.x * 28 + 22) * 4 + i] = pt[i];
        }
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// fp32 edge embedding (E, 64) -> pre-split fp16 rows [hi 64 | lo 64], values x 16  (until the edge encoder writes them itself)
// ---------------------------------------------------------------------------------------------
__global__ void emb_split_f16_kernel(const float* __restrict__ emb, uint32_t* __restrict__ out, long long n_rows) {
    const long long n_tiles = (n_rows + 127) >> 7;
    const long long total = n_tiles * 1024;       // one thread per (edge, chunk of 8 values); the edge index runs fastest: coalesced image writes
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = (i >> 10) * 128 + (i & 127);
        const int c8 = (int)((i >> 7) & 7);
        uint32_t hi[4] = {0u, 0u, 0u, 0u}, lo[4] = {0u, 0u, 0u, 0u};
        if (r < n_rows) {
            float2 v[4];
            ldg256(emb + r * 64 + c8 * 8, v[0], v[1], v[2], v[3]);
#pragma unroll
            for (int k = 0; k < 4; ++k) f16::split(make_float2(v[k].x * f16::A_SCALE, v[k].y * f16::A_SCALE), hi[k], lo[k]);
        }
        *reinterpret_cast<uint4*>(out + emb_tile_word(r, 0, c8)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(out + emb_tile_word(r, 1, c8)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// weights: element (n, k) = W[off + n * sn + k * sk], x 256 -> fp16 hi / lo chunk-major images [K/8][N][8] (rgnn_pack.cuh)
__global__ void pack_split_f16_kernel(const PackF16Args a) { pack_f16_body(a, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x); }

static int launch_pack_f16(const PackF16Args& a, int blocks, cudaStream_t stream) {
    if (packq_push(a)) return RGNN_OK;          // inside rgnn_pack_detector: batched
    pack_split_f16_kernel<<<blocks, 256, 0, stream>>>(a);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

static int g_f16_fwd = 1;
static int g_f16_passes = 3;
static int g_f16_profile = 0;
static int g_f16_stagers = mpf::NSTAGER;

bool mp_f16_supported(const ConvDims& d) { return g_f16_fwd && d.cn == 64 && d.ce == 64 && d.h == 128; }
int mp_f16_passes() { return g_f16_passes; }
size_t mp_f16_pack_floats(const ConvDims& d) { return (d.cn == 64 && d.ce == 64 && d.h == 128) ? (size_t)(2 * mpf::W1_WORDS + 2 * mpf::W2_WORDS) : 0; }
size_t mp_f16_emb_words(int n_edges) { return (size_t)((n_edges > 0 ? n_edges : 1) + 127) / 128 * EMB_TILE_WORDS; }     // whole tiles

int mp_f16_pack(const rgnn_conv& c, const ConvDims& d, float* dst, cudaStream_t stream) {
    if (mp_f16_pack_floats(d) == 0) return RGNN_OK;
    const rgnn_linear& m0 = c.msg.layer[0];
    const rgnn_linear& m1 = c.msg.layer[1];
    __half* w = reinterpret_cast<__half*>(dst);
    const int W1 = d.ce * d.h, W2 = d.h * d.cn;
    int rc = launch_pack_f16({m0.weight, w, w + W1, 2 * d.cn, m0.in_features, 1, d.ce, d.h, d.h, d.ce}, 16, stream);
    if (rc) return rc;
    return launch_pack_f16({m1.weight, w + 2 * W1, w + 2 * W1 + W2, 0, m1.in_features, 1, d.h, d.cn, d.cn, d.h}, 16, stream);
}

// (K x N) image pair of one Linear for the f16 kernels: element (n, k) = W[n * ldw + k] (n < n_valid, k < k_valid, else 0)
int pack_f16_image(const float* W, int ldw, int K, int N, int n_valid, int k_valid, uint32_t* dst, cudaStream_t stream) {
    __half* w = reinterpret_cast<__half*>(dst);
    const int blocks = (K * N + 255) / 256;
    return launch_pack_f16({W, w, w + (size_t)K * N, 0, ldw, 1, K, N, n_valid, k_valid}, blocks > 32 ? 32 : blocks, stream);
}

int mp_f16_split_emb(const float* emb, int n_edges, uint32_t* out, cudaStream_t stream) {
    if (n_edges <= 0) return RGNN_OK;
    const long long total = (((long long)n_edges + 127) / 128) * 1024;
    long long blocks = (total + 255) / 256;
    if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
    emb_split_f16_kernel<<<(unsigned)blocks, 256, 0, stream>>>(emb, out, n_edges);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

int run_conv_edges_f16(const rgnn_conv& c, const ConvDims& d, const rgnn_graph& g, const uint32_t* emb_hl, const float* P,
                       const float* wpack, float* agg, cudaStream_t stream) {
    (void)d;
    const rgnn_linear& m0 = c.msg.layer[0];
    const rgnn_linear& m1 = c.msg.layer[1];
    MpF16Args a;
    a.emb = emb_hl; a.P = P; a.tgt = g.tgt; a.src = g.src;
    a.wpack = reinterpret_cast<const uint32_t*>(wpack);
    a.agg = agg; a.n_edges = g.n_edges;
    a.passes = g_f16_passes;
    a.act1 = m0.activation; a.act2 = m1.activation;
    a.s1 = m0.norm_scale; a.m1 = m0.norm_shift;
    a.b2 = m1.bias; a.s2 = m1.norm_scale; a.m2 = m1.norm_shift;
    a.prof = nullptr;
    a.nstager = g_f16_stagers;
    static PerDeviceOnce once;
    if (once.needed()) {
        RGNN_CHECK_CUDA(cudaFuncSetAttribute(mp_edge_f16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mpf::SMEM));
        RGNN_CHECK_CUDA(cudaFuncSetAttribute(mp_edge_f16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mpf::SMEM));
        once.mark();
    }
    const int n_tiles = (g.n_edges + mpf::TM - 1) / mpf::TM;
    const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
    if (g_f16_profile) {    // developer aid (rgnn_set_option("debug", 8)): per-role cycle counters; synchronises
        long long* prof = nullptr;
        const size_t n = (size_t)grid * 28 * 4 + 16 * 16;
        RGNN_CHECK_CUDA(cudaMalloc(&prof, sizeof(long long) * n));
        RGNN_CHECK_CUDA(cudaMemsetAsync(prof, 0, sizeof(long long) * n, stream));
        a.prof = prof;
        mp_edge_f16_kernel<true><<<grid, mpf::NTHREADS, mpf::SMEM, stream>>>(a);
        RGNN_CHECK_CUDA(cudaStreamSynchronize(stream));
        long long* h = new long long[n];
        RGNN_CHECK_CUDA(cudaMemcpy(h, prof, sizeof(long long) * n, cudaMemcpyDeviceToHost));
        static const char* role[24] = {"E1a", "", "", "", "E1b", "", "", "", "E2a", "", "", "", "E2b", "", "", "", "F", "", "", "", "MMA wait: GEMM1 | GEMM2", "E1a sweep2: ld+wait | math | st+wait", "stager 0: wait | issue", ""};
        static const int slots[8] = {0, 4, 8, 12, 16, 20, 21, 22};
        for (int si = 0; si < 8; ++si) {
            const int w = slots[si];
            double tot[4] = {0, 0, 0, 0};
            for (int b = 0; b < grid; ++b) for (int i = 0; i < 4; ++i) tot[i] += (double)h[((size_t)b * 28 + w) * 4 + i];
            fprintf(stderr, "[mp_edge_f16 profile] %s warp %d cycles per tile: p0=%.0f p1=%.0f p2=%.0f\n", role[w], w, tot[0] / n_tiles,
                    tot[1] / n_tiles, tot[2] / n_tiles);
        }
        {
            const long long* ev = h + (size_t)grid * 28 * 4;
            const long long t0 = ev[0];
            fprintf(stderr, "[mp_edge_f16 events] block 0, tiles 16..31 (cycles since F(16) got its buffer): F buf | F ps | F done | G1 see | G1 iss | E1 see | E1 stat | E1 done | G2 see | G2 iss | E2 see | E2 stage | E2 epi | E2 sum\n");
            for (int t = 0; t < 16; ++t) {
                fprintf(stderr, "  tile %2d:", 16 + t);
                for (int k = 0; k < 14; ++k) fprintf(stderr, " %6lld", ev[t * 16 + k] ? ev[t * 16 + k] - t0 : -1);
                fprintf(stderr, "\n");
            }
        }
        delete[] h;
        cudaFree(prof);
        return RGNN_OK;
    }
    mp_edge_f16_kernel<false><<<grid, mpf::NTHREADS, mpf::SMEM, stream>>>(a);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

int mp_f16_set_option(const char* name, int value) {
    if (strcmp(name, "f16_fwd") == 0 && (value == 0 || value == 1)) { g_f16_fwd = value; return 1; }
    if (strcmp(name, "f16_passes") == 0 && (value == 1 || value == 3)) { g_f16_passes = value; return 1; }
    if (strcmp(name, "f16_stagers") == 0 && value >= 2 && value <= mpf::NSTAGER) { g_f16_stagers = value; return 1; }
    if (strcmp(name, "debug") == 0) { g_f16_profile = (value & 8) != 0; return 0; }      // shared with the other kernels
    return 0;
}
int mp_f16_get_option(const char* name) {
    if (strcmp(name, "f16_fwd") == 0) return g_f16_fwd;
    if (strcmp(name, "f16_passes") == 0) return g_f16_passes;
    if (strcmp(name, "f16_stagers") == 0) return g_f16_stagers;
    return -2;
}

}  // namespace rgnn
