// Backward of the message function of one residual_graph_conv_block, FUSED: recompute + data gradients + both weight
// gradients in ONE kernel on the 5th-generation tensor cores with fp16-split operands (rgnn_f16.cuh).
// Replaces torch autograd over gnn_blocks.py:106-113 (reference gnn/training.py:81).
//
// Per tile of 128 target-major edges (thread = TMEM lane = edge row, two threads per row: one half of the columns each):
//   G1  D1 = 4096 (P_t + P_s) + emb W_e^T        TS, accumulator pre-loaded by the F role    E1: norm -> act -> y1
//   G2  D2 = y1 W_2^T                            TS                                          E2: + b2, statistics, mask;
//                                                                                                d(msg) = dagg[tgt] -> act' -> norm' -> dz2
//   G3  D3 = dz2 W_2     = d(y1)                 TS, B = the FORWARD image of W_2 read MN-major
//   dW2 += y1^T dz2                              SS, both operands MN-major: the reduction runs over the EDGES of the tile
//                                                                                            E3: act' -> norm' -> dz1
//   G4  D4 = dz1 W_e     = d(emb)                TS, B = the FORWARD image of W_e read MN-major
//   dWe += dz1^T [emb | 1]                       SS, MN-major; the ones column delivers the bias gradient sum_e dz1
//                                                                                            E4: demb (+)= D4
// What makes the fusion fit: with 16-bit operands ONE "chunk-major" image uint4 op[K/8][rows] is at the same time the
// canonical K-major operand (rows x K) and the canonical MN-major operand (K x rows) of tcgen05.mma (instruction
// descriptor bits 15 / 16; verified bit-exact by tools/micro/test_mnmajor.cu).  So (a) the two forward weight images serve
// W and W^T (64 KB instead of the 256 KB of 3xTF32 images the tf32 kernel streams per tile), and (b) the activation images
// the epilogues write for the weight gradients -- y1 (dz1 replaces it in place), dz2, emb: 136 KB -- are read with the
// edge index as the K dimension.  The weight-gradient accumulators live in tensor memory for the whole kernel
// (144 columns) and are flushed once per CTA.  Nothing per edge leaves the SM but demb (read-modify-write) and dz1 (the
// per-node projection gradient dP is a gather over dz1 by target and by source: dproj_gather_kernel).
//
// Gradient range: fp16 has 5 exponent bits, so the gradient operands (dz2, dz1) carry ONE power-of-two scale S per launch,
// chosen from max |dagg| (a 10-microsecond reduction launched in front) so that S max|dagg| lies in (2^7, 2^8]: 2^8 of headroom
// below the fp16 maximum for the amplification by s / sigma, and every element above 2^-19 of the largest keeps the full
// 22-bit hi + lo representation (smaller ones keep an absolute error of 2^-25 / S: they do not matter in a sum).  All scales
// are powers of two and are taken out again exactly.
//
// Tensor memory (512 columns): R1 [0,128) D1 -> y1 hi|lo -> D3 -> dz1 hi|lo;  X0 [128,192) / X1 [400,464) emb hi|lo of the
// even / odd tiles, then dz2 hi|lo;  R3 [192,256) D2 -> D4;  [256,320) dW2 accumulator;  [320,400) dWe accumulator (N = 80).
// Roles (512 threads): 8 worker warps (epilogues E1..E4), 4 F warps (next tile's emb operand and accumulator pre-load,
// one tile ahead), 1 MMA-issue lane.  One tile is in flight per CTA: both the tensor memory (400 columns) and the shared
// memory (208 KB) hold exactly one tile's operands next to the accumulators and the weight images.
#include <vector>

#include "rgnn_f16.cuh"
#include "rgnn_model.h"
#include "rgnn_tc_rows.cuh"
#include "rgnn_tile.cuh"

namespace rgnn {

struct MpBwdF16Args {
    const uint32_t* emb;    // pre-split edge rows, values x 16, target-major, TILED (rgnn_f16.cuh: emb_tile_word)
    const float* P;         // (N, 2H) fp32: [x W_t^T + b1 | x W_s^T]
    const float* dagg;      // (N, CN) gradient w.r.t. the aggregated messages
    const int* tgt;
    const int* src;
    const uint32_t* wpack;  // [W_e hi | W_e lo | W_2 hi | W_2 lo], fp16 x 256, chunk-major (the forward kernel's images)
    const float* s1; const float* m1;      // channel_normalization scalars of msg.0 / msg.1 (nullptr = no norm)
    const float* b2; const float* s2; const float* m2;
    const float* gmax;      // device scalar: max |dagg|
    float* dz1_out;         // (E, H) fp32 (input of dproj_gather_kernel)
    float* demb;            // (E, CE) accumulated over the layers
    float* gW2;             // (CN, H) += dz2^T y1
    float* gb2;             // (CN)    += sum_e dz2
    float* gWe;             // rows of ldWe floats: (H, CE) block of msg.0's weight gradient, += dz1^T emb
    int ldWe;
    float* gb1;             // (H)     += sum_e dz1
    float* g_s1; float* g_m1; float* g_s2; float* g_m2;     // gradients of the norm scalars (+=), nullable
    int n_edges;
    int act1, act2;
    int demb_accumulate;
    int passes;
    long long* prof;        // PROFILE builds: [grid][24] cycle counters: worker thread 0 [0,12), F thread 0 [12,18), MMA lane [18,24)
};

namespace mbf {
constexpr int CE = 64, H = 128, CN = 64, TM = 128;
constexpr int NTHREADS = 512, NW = 256;                      // workers: (row, half of the columns)
constexpr int W1_WORDS = CE * H / 2, W2_WORDS = H * CN / 2;  // 32-bit words per (hi or lo) weight image
constexpr int NE = CE + 16;                                  // emb image columns incl. the ones column (N of the dWe MMA)
constexpr int OFF_W = 0;                                     // words: W_e hi | lo | W_2 hi | lo
constexpr int Y_WORDS = TM * H / 2;                          // one (hi or lo) image of y1 / dz1: 32 KB
constexpr int OFF_Y = OFF_W + 2 * W1_WORDS + 2 * W2_WORDS;
constexpr int Z_WORDS = TM * CN / 2;                         // dz2: 16 KB
constexpr int OFF_Z = OFF_Y + 2 * Y_WORDS;
constexpr int E_WORDS = TM * NE / 2;                         // emb (+ ones): 20 KB
constexpr int STAGE_PITCH = CE + 4;                          // words per row of the d(emb) staging tile that shares the dz2 image's region (+ 2 KB)
static_assert(TM * STAGE_PITCH <= 2 * Z_WORDS + 512, "staging tile");
constexpr int OFF_E = OFF_Z + 2 * Z_WORDS + 512;
constexpr int OFF_XCH = OFF_E + 2 * E_WORDS;                 // [2 phases][TM][2 q] float2
constexpr int OFF_BIAS = OFF_XCH + 2 * TM * 2 * 2;
constexpr int OFF_RED = OFF_BIAS + CN;                       // 4 x 8 doubles
constexpr int OFF_BAR = OFF_RED + 64;                        // 10 mbarriers
constexpr int OFF_SLOT = OFF_BAR + 2 * 16;
constexpr int WORDS = OFF_SLOT + 2;
constexpr size_t SMEM = (size_t)WORDS * 4;
static_assert((OFF_BAR % 2) == 0 && (OFF_RED % 2) == 0 && (OFF_XCH % 2) == 0 && (OFF_Y % 4) == 0 && (OFF_Z % 4) == 0 && (OFF_E % 4) == 0,
              "alignment");
static_assert(SMEM <= 227 * 1024, "shared memory budget");
constexpr uint32_t COL_R1 = 0, COL_X0 = 128, COL_R3 = 192, COL_ACC2 = 256, COL_ACCE = 320, COL_X1 = 400;
enum { B_A_FULL = 0, B_D1_FULL, B_Y1_FULL, B_D2_FULL, B_Z2_FULL, B_D3_FULL, B_W2_DONE, B_Z1_FULL, B_D4_FULL, B_WE_DONE, B_D4_READ, B_STAGE_FREE, B_YIMG_READ };
constexpr int REG_W = 152, REG_F = 168, REG_AUX = 40;
static_assert(NW * REG_W + 128 * REG_F + 128 * REG_AUX <= 65536, "setmaxnreg pool");
__host__ __device__ constexpr uint32_t idesc_mn(int M, int N, int a_mn, int b_mn) {
    return f16::idesc(M, N) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}
}  // namespace mbf

// sum over the 32 rows of a warp of 32 per-thread columns v[0..31]: recursive halving, lane l ends with column l
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {
            const float keep = up ? v[i + o] : v[i];
            const float send = up ? v[i] : v[i + o];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return v[0];
}

template <bool PROFILE>
__global__ void __launch_bounds__(mbf::NTHREADS, 1) mp_edge_bwd_f16_kernel(const __grid_constant__ MpBwdF16Args a) {
    using namespace mbf;
    extern __shared__ __align__(1024) uint32_t smem_u[];
    float* smem_f = reinterpret_cast<float*>(smem_u);
    uint32_t* wsm = smem_u + OFF_W;
    uint4* yimg = reinterpret_cast<uint4*>(smem_u + OFF_Y);          // [hi | lo][H/8][TM] 16-byte chunks
    uint4* zimg = reinterpret_cast<uint4*>(smem_u + OFF_Z);          // [hi | lo][CN/8][TM]
    uint4* eimg = reinterpret_cast<uint4*>(smem_u + OFF_E);          // [hi | lo][NE/8][TM]
    float* stage = smem_f + OFF_Z;                                   // [TM][STAGE_PITCH] fp32, alias of the dz2 image (E4)
    float2* xch = reinterpret_cast<float2*>(smem_u + OFF_XCH);
    float* bias_s = smem_f + OFF_BIAS;
    double* red = reinterpret_cast<double*>(smem_u + OFF_RED);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_u + OFF_BAR);
    uint32_t* slot = smem_u + OFF_SLOT;
    constexpr int YI = Y_WORDS / 4, ZI = Z_WORDS / 4, EI = E_WORDS / 4;   // uint4 per image

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = tid & 127;
    const int G = (int)gridDim.x;
    const int n_tiles = (a.n_edges + TM - 1) / TM;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / G + 1 : 0;
    const int np = a.passes == 1 ? 1 : 3;

    // ---- one-time setup: weight images, the constant ones column of the emb image, barriers, tensor memory ----
    {
        const uint4* g = reinterpret_cast<const uint4*>(a.wpack);
        uint4* s = reinterpret_cast<uint4*>(wsm);
        constexpr int N4 = (2 * W1_WORDS + 2 * W2_WORDS) / 4;
        for (int i = tid; i < N4; i += NTHREADS) s[i] = __ldg(g + i);
    }
    for (int i = tid; i < 2 * 2 * TM; i += NTHREADS) {      // chunks CE/8 and CE/8 + 1 of both images: column CE = 16.0 (hi), rest 0
        const int img = i / (2 * TM), r = i % (2 * TM);      // r = chunk * TM + edge
        const bool one = img == 0 && r < TM;
        eimg[img * EI + (CE / 8) * TM + r] = make_uint4(one ? 0x4C00u : 0u, 0u, 0u, 0u);      // fp16 16.0 = 0x4C00 in channel CE
    }
    if (tid < CN) bias_s[tid] = a.b2 != nullptr ? __ldg(a.b2 + tid) : 0.f;
    if (tid == 0) {
        tc::mbar_init(&bars[B_A_FULL], 4);
        tc::mbar_init(&bars[B_D1_FULL], 1);
        tc::mbar_init(&bars[B_Y1_FULL], 8);
        tc::mbar_init(&bars[B_D2_FULL], 1);
        tc::mbar_init(&bars[B_Z2_FULL], 8);
        tc::mbar_init(&bars[B_D3_FULL], 1);
        tc::mbar_init(&bars[B_W2_DONE], 1);
        tc::mbar_init(&bars[B_Z1_FULL], 8);
        tc::mbar_init(&bars[B_D4_FULL], 1);
        tc::mbar_init(&bars[B_WE_DONE], 1);
        tc::mbar_init(&bars[B_D4_READ], 4);
        tc::mbar_init(&bars[B_STAGE_FREE], 4);
        tc::mbar_init(&bars[B_YIMG_READ], 3);
        tc::mbar_init_fence();
    }
    if (warp == 0) tc::tmem_alloc(slot, 512);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *slot;
    const uint32_t t_row = tmem + ((uint32_t)(warp & 3) << 21);       // this warp's 32 TMEM lanes

    // gradient scale S = 2^(8 - e) with max|dagg| = m 2^e, m in [0.5, 1)
    float S = 1.f, inv_S = 1.f;
    {
        const float gm = __ldg(a.gmax);
        if (gm > 0.f && gm < 3.0e38f) {
            int e = (int)((__float_as_uint(gm) >> 23) & 0xFFu) - 126;
            int k = 8 - e;
            k = k < -60 ? -60 : (k > 100 ? 100 : k);
            S = __uint_as_float((uint32_t)(127 + k) << 23);
            inv_S = __uint_as_float((uint32_t)(127 - k) << 23);
        }
    }

    if (tid < NW) {
        // =========================== workers: the four epilogues of every tile ===========================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REG_W));
        const int q = tid >> 7;
        const int bar_id = 1 + (row >> 5);
        const bool norm1 = a.s1 != nullptr, norm2 = a.s2 != nullptr;
        const float s1v = norm1 ? __ldg(a.s1) : 1.f, m1v = norm1 ? __ldg(a.m1) : 0.f;
        const float s2v = norm2 ? __ldg(a.s2) : 1.f, m2v = norm2 ? __ldg(a.m2) : 0.f;
        const float inv_s1 = s1v != 0.f ? 1.f / s1v : 0.f;
        const bool act1 = a.act1 != 0, act2 = a.act2 != 0;
        const float eps_s = NORM_EPS * f16::D_SCALE;
        double acc_s1 = 0., acc_m1 = 0., acc_s2 = 0., acc_m2 = 0.;
        float acc_b2 = 0.f;
        long long pt[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tl = 0;
        auto tick = [&](int i) {
            if (PROFILE && tid == 0) { const long long n = clock64(); pt[i] += n - tl; tl = n; }
        };
        if (PROFILE) tl = clock64();

        for (int j = 0; j < my_tiles; ++j) {
            const uint32_t ph = (uint32_t)j & 1u;
            const int tile = (int)blockIdx.x + j * G;
            const int e_my = tile * TM + row;
            const bool valid = e_my < a.n_edges;
            const uint32_t xcol = t_row + ((j & 1) ? COL_X1 : COL_X0);
            const uint32_t r1 = t_row + COL_R1, r3 = t_row + COL_R3;
            const int t_my = valid ? __ldg(a.tgt + e_my) : 0;
            float2 g2[16];          // own 32 columns of d(message) = dagg[target]: requested now, used in E2
            {
                const float* dg = a.dagg + (size_t)t_my * CN + 32 * q;
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    g2[4 * c8] = g2[4 * c8 + 1] = g2[4 * c8 + 2] = g2[4 * c8 + 3] = make_float2(0.f, 0.f);
                    if (valid) ldg256(dg + 8 * c8, g2[4 * c8], g2[4 * c8 + 1], g2[4 * c8 + 2], g2[4 * c8 + 3]);
                }
            }

            // ---------------- E1: z1 (x 4096) -> statistics -> y1 = act(norm(z1)) -> fp16 hi | lo: tensor memory + image ----------------
            tc::mbar_wait(&bars[B_D1_FULL], ph);
            tc::tc_fence_after();
            tick(0);
            float k1, sh1, mean1, sd1 = 0.f;
            // the statistics sweep visits the PARTNER's two 32-column chunks first and this thread's own two last, so that the own
            // chunks are still in registers (va, vb) when the row's mean and sigma are known: no second read of D1
            float2 va[16], vb[16];
            {
                const int co = 64 * (q ^ 1), cm = 64 * q;
                RowStats st;
                st.init();
                tc::tmem_ld16(r1 + co, va);
                tc::tmem_ld16(r1 + co + 16, va + 8);
                tc::tmem_ld16(r1 + co + 32, vb);
                tc::tmem_ld16(r1 + co + 48, vb + 8);
                tc::tmem_wait_ld();
                if (norm1) { st.add_chunk(va); st.add_chunk(vb); }
                tc::tmem_ld16(r1 + cm, va);
                tc::tmem_ld16(r1 + cm + 16, va + 8);
                tc::tmem_ld16(r1 + cm + 32, vb);
                tc::tmem_ld16(r1 + cm + 48, vb + 8);
                tc::tmem_wait_ld();
                if (norm1) {
                    st.add_chunk(va);
                    st.add_chunk(vb);
                    const float sds = st.sigma(H);
                    k1 = f16::A_SCALE * s1v * __frcp_rn(sds + eps_s);
                    sh1 = f16::A_SCALE * m1v;
                    mean1 = st.mean;
                    sd1 = sds * f16::D_UNSCALE;
                } else {
                    k1 = f16::A_SCALE * f16::D_UNSCALE;
                    sh1 = 0.f;
                    mean1 = 0.f;
                }
            }
            tc::tc_fence_before();
            group_sync(bar_id, 64);            // the row partner has read every column: in-place stores may begin
            tc::tc_fence_after();
            if (j > 0) {
                tc::mbar_wait(&bars[B_WE_DONE], ph ^ 1u);      // the previous tile's dWe MMAs have read the y1/dz1 and emb images
                tc::mbar_wait(&bars[B_YIMG_READ], ph ^ 1u);    // and the copy warps have written its dz1 rows out
            }
            {
                const float2 k2 = make_float2(k1, k1), sh2 = make_float2(sh1, sh1), sl = make_float2(LEAKY, LEAKY), nm = make_float2(-mean1, -mean1);
                auto emit = [&](float2 (&v)[16], int c) {
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float2 y = __ffma2_rn(__fadd2_rn(v[i], nm), k2, sh2);
                        if (act1) {
                            const float2 t = __fmul2_rn(y, sl);
                            y.x = fmaxf(y.x, t.x);
                            y.y = fmaxf(y.y, t.y);
                        }
                        f16::split(y, hi[i], lo[i]);
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        yimg[(c / 8 + k) * TM + row] = make_uint4(hi[4 * k], hi[4 * k + 1], hi[4 * k + 2], hi[4 * k + 3]);
                        yimg[YI + (c / 8 + k) * TM + row] = make_uint4(lo[4 * k], lo[4 * k + 1], lo[4 * k + 2], lo[4 * k + 3]);
                    }
                };
                emit(va, 64 * q);
                emit(vb, 64 * q + 32);
            }
            tc::tc_fence_before();
            tc::fence_async_smem();
            warp_arrive(&bars[B_Y1_FULL], lane);
            tick(1);
            // while G2 runs: the emb operand (still in X) -> emb image (q = 0: hi, q = 1: lo); d(message) = dagg[target]
            {
                uint32_t ev[32];
                tc::tc_fence_after();
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                    : "=r"(ev[0]), "=r"(ev[1]), "=r"(ev[2]), "=r"(ev[3]), "=r"(ev[4]), "=r"(ev[5]), "=r"(ev[6]), "=r"(ev[7]), "=r"(ev[8]),
                      "=r"(ev[9]), "=r"(ev[10]), "=r"(ev[11]), "=r"(ev[12]), "=r"(ev[13]), "=r"(ev[14]), "=r"(ev[15]), "=r"(ev[16]),
                      "=r"(ev[17]), "=r"(ev[18]), "=r"(ev[19]), "=r"(ev[20]), "=r"(ev[21]), "=r"(ev[22]), "=r"(ev[23]), "=r"(ev[24]),
                      "=r"(ev[25]), "=r"(ev[26]), "=r"(ev[27]), "=r"(ev[28]), "=r"(ev[29]), "=r"(ev[30]), "=r"(ev[31])
                    : "r"(xcol + 32 * q)
                    : "memory");
                tc::tmem_wait_ld();
#pragma unroll
                for (int k = 0; k < 8; ++k) eimg[q * EI + k * TM + row] = make_uint4(ev[4 * k], ev[4 * k + 1], ev[4 * k + 2], ev[4 * k + 3]);
            }

            // ---------------- E2: statistics and mask of z2; dz2 = norm'(act'(d message)), x S -> hi | lo: tensor memory + image ----------------
            tc::mbar_wait(&bars[B_D2_FULL], ph);
            tc::tc_fence_after();
            tick(2);
            {
                float mean2 = 0.f, sd2 = 0.f, inv_den = 1.f;
                const float2 us = make_float2(f16::D_UNSCALE, f16::D_UNSCALE);
                float2 c2[16];          // own 32 columns of z2 (centred)
                {
                    float2 vo[16];      // the partner's 32 columns (statistics only)
                    const int co = 32 * (q ^ 1), cm = 32 * q;
                    tc::tmem_ld16(r3 + co, vo);
                    tc::tmem_ld16(r3 + co + 16, vo + 8);
                    tc::tmem_ld16(r3 + cm, c2);
                    tc::tmem_ld16(r3 + cm + 16, c2 + 8);
                    tc::tmem_wait_ld();
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        vo[c] = __ffma2_rn(vo[c], us, *reinterpret_cast<const float2*>(bias_s + co + 2 * c));
                        c2[c] = __ffma2_rn(c2[c], us, *reinterpret_cast<const float2*>(bias_s + cm + 2 * c));
                    }
                    if (norm2) {
                        RowStats st;
                        st.init();
                        st.add_chunk(vo);
                        st.add_chunk(c2);
                        mean2 = st.mean;
                        sd2 = st.sigma(CN);
                        inv_den = 1.f / (sd2 + NORM_EPS);
                    }
                }
                const float2 nm = make_float2(-mean2, -mean2);
#pragma unroll
                for (int c = 0; c < 16; ++c) c2[c] = __fadd2_rn(c2[c], nm);
                const float k = s2v * inv_den;
                const float2 k2 = make_float2(k, k), m22 = make_float2(m2v, m2v), id2 = make_float2(inv_den, inv_den);
                const float2 s22 = make_float2(s2v, s2v);
                float2 ps2 = make_float2(0.f, 0.f), pm2 = ps2, sum2 = ps2, dot2 = ps2;
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const float2 y = norm2 ? __ffma2_rn(c2[c], k2, m22) : c2[c];
                    g2[c] = __fmul2_rn(g2[c], make_float2((!act2 || y.x > 0.f) ? 1.f : LEAKY, (!act2 || y.y > 0.f) ? 1.f : LEAKY));
                    if (norm2) {
                        c2[c] = __fmul2_rn(c2[c], id2);              // nv
                        ps2 = __ffma2_rn(g2[c], c2[c], ps2);
                        pm2 = __fadd2_rn(pm2, g2[c]);
                        g2[c] = __fmul2_rn(g2[c], s22);              // dn
                        sum2 = __fadd2_rn(sum2, g2[c]);
                        dot2 = __ffma2_rn(g2[c], c2[c], dot2);
                    }
                }
                if (norm2) {
                    if (valid) { acc_s2 += (double)(ps2.x + ps2.y); acc_m2 += (double)(pm2.x + pm2.y); }
                    xch[row * 2 + q] = make_float2(sum2.x + sum2.y, dot2.x + dot2.y);
                    group_sync(bar_id, 64);
                    const float2 o = xch[row * 2 + (q ^ 1)];
                    const float sum_dn = (sum2.x + sum2.y) + o.x, dot = (dot2.x + dot2.y) + o.y;
                    const float mean_dn = sum_dn / (float)CN;
                    const float coef = sd2 > 0.f ? dot / ((float)(CN - 1) * sd2) : 0.f;
                    const float2 nm2 = make_float2(-mean_dn, -mean_dn), nc2 = make_float2(-coef, -coef);
#pragma unroll
                    for (int c = 0; c < 16; ++c) g2[c] = __ffma2_rn(c2[c], nc2, __fmul2_rn(__fadd2_rn(g2[c], nm2), id2));
                } else {
                    group_sync(bar_id, 64);     // (the partner has finished copying the emb operand out of X)
                }
                // x S, split, store: X hi columns [16 q, 16 q + 16), lo columns [32 + 16 q, ...); image chunks 4 q .. 4 q + 3
                uint32_t hi[16], lo[16];
                float cs[32];
                const float2 S2 = make_float2(S, S);
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const float2 v = __fmul2_rn(g2[c], S2);
                    cs[2 * c] = v.x;
                    cs[2 * c + 1] = v.y;
                    f16::split(v, hi[c], lo[c]);
                }
                f16::tmem_st16u(xcol + 16 * q, hi);
                f16::tmem_st16u(xcol + 32 + 16 * q, lo);
                if (j > 0) tc::mbar_wait(&bars[B_STAGE_FREE], ph ^ 1u);       // E4 of the previous tile has drained its staging tile (same region)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    zimg[(4 * q + k) * TM + row] = make_uint4(hi[4 * k], hi[4 * k + 1], hi[4 * k + 2], hi[4 * k + 3]);
                    zimg[ZI + (4 * q + k) * TM + row] = make_uint4(lo[4 * k], lo[4 * k + 1], lo[4 * k + 2], lo[4 * k + 3]);
                }
                tc::tmem_wait_st();
                tc::tc_fence_before();
                tc::fence_async_smem();
                warp_arrive(&bars[B_Z2_FULL], lane);
                acc_b2 += warp_colsum32(cs, lane);      // bias gradient of msg.1: column sums of dz2 (this warp's rows, column 32 q + lane)
            }
            tick(3);

            // ---------------- E3: d(y1) = D3 / 256 (x S) -> act' -> norm' -> dz1 -> hi | lo in place (tensor memory + image), fp32 scratch ----------------
            tc::mbar_wait(&bars[B_D3_FULL], ph);
            tc::tc_fence_after();
            tick(4);
            {
                float2 dn[32];          // own 64 columns
                const float2 is2 = make_float2(inv_s1, inv_s1), nsh2 = make_float2(-m1v * inv_s1, -m1v * inv_s1);
                const float2 s12 = make_float2(s1v, s1v);
                const float2 ua = make_float2(1.f / f16::A_SCALE, 1.f / f16::A_SCALE), uw = make_float2(1.f / f16::W_SCALE, 1.f / f16::W_SCALE);
                float2 ps2 = make_float2(0.f, 0.f), pm2 = ps2, sum2 = ps2, dot2 = ps2;
                tc::tmem_ld16(r1 + 64 * q, dn);
                tc::tmem_ld16(r1 + 64 * q + 16, dn + 8);
                tc::tmem_ld16(r1 + 64 * q + 32, dn + 16);
                tc::tmem_ld16(r1 + 64 * q + 48, dn + 24);
                tc::tmem_wait_ld();
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int c = 64 * q + 32 * cc;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint4 yh = yimg[(c / 8 + k) * TM + row], yl = yimg[YI + (c / 8 + k) * TM + row];
                        const uint32_t h4[4] = {yh.x, yh.y, yh.z, yh.w}, l4[4] = {yl.x, yl.y, yl.z, yl.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float2& g = dn[16 * cc + 4 * k + i];
                            const float2 y = __fmul2_rn(__fadd2_rn(f16::unpack(h4[i]), f16::unpack(l4[i])), ua);
                            float2 nv;
                            g = __fmul2_rn(g, uw);
                            act_bwd_pair(g, nv, y, act1, is2, nsh2);
                            if (norm1) {
                                ps2 = __ffma2_rn(g, nv, ps2);
                                pm2 = __fadd2_rn(pm2, g);
                                g = __fmul2_rn(g, s12);
                                sum2 = __fadd2_rn(sum2, g);
                                dot2 = __ffma2_rn(g, nv, dot2);
                            }
                        }
                    }
                }
                float mean_dn = 0.f, coef = 0.f, inv_den = 1.f;
                if (norm1) {
                    if (valid) { acc_s1 += (double)(ps2.x + ps2.y); acc_m1 += (double)(pm2.x + pm2.y); }
                    xch[2 * TM + row * 2 + q] = make_float2(sum2.x + sum2.y, dot2.x + dot2.y);
                    group_sync(bar_id, 64);
                    const float2 o = xch[2 * TM + row * 2 + (q ^ 1)];
                    const float sum_dn = (sum2.x + sum2.y) + o.x, dot = (dot2.x + dot2.y) + o.y;
                    inv_den = 1.f / (sd1 + NORM_EPS);
                    mean_dn = sum_dn / (float)H;
                    coef = sd1 > 0.f ? dot / ((float)(H - 1) * sd1) : 0.f;
                }
                tick(5);
                tc::mbar_wait(&bars[B_W2_DONE], ph);        // the dW2 MMAs have read the y1 image: dz1 may replace it
                tick(6);
                const float2 nm2 = make_float2(-mean_dn, -mean_dn), id2 = make_float2(inv_den, inv_den), nc2 = make_float2(-coef, -coef);
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int c = 64 * q + 32 * cc;
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint4 yh = yimg[(c / 8 + k) * TM + row], yl = yimg[YI + (c / 8 + k) * TM + row];
                        const uint32_t h4[4] = {yh.x, yh.y, yh.z, yh.w}, l4[4] = {yl.x, yl.y, yl.z, yl.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float2& g = dn[16 * cc + 4 * k + i];
                            if (norm1) {
                                const float2 y = __fmul2_rn(__fadd2_rn(f16::unpack(h4[i]), f16::unpack(l4[i])), ua);
                                const bool px = !act1 || y.x > 0.f, py = !act1 || y.y > 0.f;
                                const float2 nv = __ffma2_rn(__fmul2_rn(y, make_float2(px ? 1.f : 1.f / LEAKY, py ? 1.f : 1.f / LEAKY)), is2, nsh2);
                                g = __ffma2_rn(nv, nc2, __fmul2_rn(__fadd2_rn(g, nm2), id2));
                            }
                            f16::split(g, hi[4 * k + i], lo[4 * k + i]);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        yimg[(c / 8 + k) * TM + row] = make_uint4(hi[4 * k], hi[4 * k + 1], hi[4 * k + 2], hi[4 * k + 3]);
                        yimg[YI + (c / 8 + k) * TM + row] = make_uint4(lo[4 * k], lo[4 * k + 1], lo[4 * k + 2], lo[4 * k + 3]);
                    }
                }
                tc::tc_fence_before();
                tc::fence_async_smem();
                warp_arrive(&bars[B_Z1_FULL], lane);
                tick(8);
                // (the fp32 copy of dz1 for dproj_gather_kernel is written by the copy warps from the dz1 image)
            }
            tick(9);

            // (E4, d(emb) (+)= D4, runs in the F role: global read-modify-write off the workers' chain)
            tc::tc_fence_before();
        }

        // ---------------- flush: weight-gradient accumulators, bias gradients, norm-scalar gradients ----------------
        if (my_tiles > 0) {
            tc::mbar_wait(&bars[B_WE_DONE], (uint32_t)(my_tiles - 1) & 1u);
            tc::tc_fence_after();
            const float u = inv_S * (1.f / f16::A_SCALE);
            // dW2: lane = y1 channel m, column = dz2 channel n  ->  gW2[n * H + m]
            if (a.gW2 != nullptr) {
#pragma unroll 1
                for (int cc = 0; cc < 2; ++cc) {
                    float v[16];
                    tc::tmem_ld16(t_row + COL_ACC2 + 32 * q + 16 * cc, v);
                    tc::tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; ++i) atomicAdd(a.gW2 + (size_t)(32 * q + 16 * cc + i) * H + row, v[i] * u);
                }
            }
            // dWe: lane = dz1 channel m, column = emb channel n (column CE: the ones column)  ->  gWe[m * ldWe + n], gb1[m]
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                float v[16];
                tc::tmem_ld16(t_row + COL_ACCE + 32 * q + 16 * cc, v);
                tc::tmem_wait_ld();
                if (a.gWe != nullptr) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) atomicAdd(a.gWe + (size_t)row * a.ldWe + 32 * q + 16 * cc + i, v[i] * u);
                }
            }
            if (q == 1 && a.gb1 != nullptr) {
                float v[16];
                tc::tmem_ld16(t_row + COL_ACCE + CE, v);
                tc::tmem_wait_ld();
                atomicAdd(a.gb1 + row, v[0] * u);
            }
            if (a.gb2 != nullptr) atomicAdd(a.gb2 + 32 * q + lane, acc_b2 * inv_S);
        }
        if (PROFILE && tid == 0 && a.prof != nullptr)
            for (int i = 0; i < 12; ++i) a.prof[blockIdx.x * 24 + i] = pt[i];
        {
            double v[4] = {acc_s1 * (double)inv_S, acc_m1 * (double)inv_S, acc_s2, acc_m2};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
                if (lane == 0) red[i * 8 + warp] = v[i];
            }
            group_sync(BAR_WORKERS, NW);
            if (tid < 4) {
                double s = 0.;
                for (int w = 0; w < NW / 32; ++w) s += red[tid * 8 + w];
                float* dst = tid == 0 ? a.g_s1 : tid == 1 ? a.g_m1 : tid == 2 ? a.g_s2 : a.g_m2;
                if (dst != nullptr && my_tiles > 0) atomicAdd(dst, (float)s);
            }
        }
    } else if (tid < NW + 128) {
        // =========================== F: emb operand and accumulator pre-load, one tile ahead ===========================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REG_F));
        const float2 sc = make_float2(f16::D_SCALE, f16::D_SCALE);
        long long ft[6] = {0, 0, 0, 0, 0, 0}, fl = 0;
        auto ftick = [&](int i) {
            if (PROFILE && tid == NW) { const long long n = clock64(); ft[i] += n - fl; fl = n; }
        };
        if (PROFILE) fl = clock64();
        // E4 of tile jj: d(emb) (+)= D4 / (256 S), this thread's whole row (D4_FULL of tile jj has been observed by the caller);
        // the arrive tells the MMA lane that R3 may be overwritten by G2 of the next tile
        auto e4 = [&](int jj) {
            const int e = ((int)blockIdx.x + jj * G) * TM + row;
            const bool valid = e < a.n_edges;
            float2 d[32];
            tc::tmem_ld16(t_row + COL_R3, d);
            tc::tmem_ld16(t_row + COL_R3 + 16, d + 8);
            tc::tmem_ld16(t_row + COL_R3 + 32, d + 16);
            tc::tmem_ld16(t_row + COL_R3 + 48, d + 24);
            tc::tmem_wait_ld();
            tc::tc_fence_before();
            warp_arrive(&bars[B_D4_READ], lane);
            // the row goes through a staging tile in shared memory (the dz2 image's region: its last reader, the dW2 MMAs of this
            // tile, completed before D4_FULL) and from there to global memory with ONE bulk reduce-add per row: no global load,
            // no per-thread sector stores on the load/store pipe
            const float u = inv_S * (1.f / f16::W_SCALE);
            const float2 u2 = make_float2(u, u);
            float* st = stage + row * STAGE_PITCH;
#pragma unroll
            for (int c4 = 0; c4 < 16; ++c4) {
                const float2 v0 = __fmul2_rn(d[2 * c4], u2), v1 = __fmul2_rn(d[2 * c4 + 1], u2);
                *reinterpret_cast<float4*>(st + 4 * c4) = make_float4(v0.x, v0.y, v1.x, v1.y);
            }
            tc::fence_async_smem();
            if (valid) {
                float* o = a.demb + (size_t)e * CE;
                if (a.demb_accumulate)
                    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                                 ::"l"(o), "r"(tc::smem_u32(st)), "n"(CE * 4) : "memory");
                else
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                 ::"l"(o), "r"(tc::smem_u32(st)), "n"(CE * 4) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            warp_arrive(&bars[B_STAGE_FREE], lane);
        };
        for (int j = 0; j < my_tiles; ++j) {
            const int tile = (int)blockIdx.x + j * G;
            const int e = tile * TM + row;
            const bool valid = e < a.n_edges;
            const uint32_t xcol = t_row + ((j & 1) ? COL_X1 : COL_X0);
            int t = 0, s = 0;
            if (valid) { t = __ldg(a.tgt + e); s = __ldg(a.src + e); }
            {   // emb hi | lo -> X (free: the dz2 operand of tile j - 2 was consumed by its G3, which this role has seen complete)
                uint32_t ev[64];
                const long long el = valid ? e : 0;
#pragma unroll
                for (int i = 0; i < 16; ++i) {      // tiled rows (rgnn_f16.cuh): 16 bytes per chunk, a warp reads 512 contiguous bytes
                    const uint4 v = ldg128u(a.emb + emb_tile_word(el, i >> 3, i & 7));
                    ev[4 * i] = v.x; ev[4 * i + 1] = v.y; ev[4 * i + 2] = v.z; ev[4 * i + 3] = v.w;
                }
                if (!valid) {
#pragma unroll
                    for (int i = 0; i < 64; ++i) ev[i] = 0u;
                }
                f16::tmem_st16u(xcol, ev);
                f16::tmem_st16u(xcol + 16, ev + 16);
                f16::tmem_st16u(xcol + 32, ev + 32);
                f16::tmem_st16u(xcol + 48, ev + 48);
            }
            float2 pre[64];          // 4096 (P_t[target] + P_s[source])
            {
                const float* Pt = a.P + (size_t)t * (2 * H);
                const float* Ps = a.P + (size_t)s * (2 * H) + H;
#pragma unroll
                for (int i = 0; i < 16; ++i) ldg256(Pt + 8 * i, pre[4 * i], pre[4 * i + 1], pre[4 * i + 2], pre[4 * i + 3]);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float2 p0, p1, p2, p3;
                    ldg256(Ps + 8 * i, p0, p1, p2, p3);
                    pre[4 * i] = valid ? __fmul2_rn(__fadd2_rn(pre[4 * i], p0), sc) : make_float2(0.f, 0.f);
                    pre[4 * i + 1] = valid ? __fmul2_rn(__fadd2_rn(pre[4 * i + 1], p1), sc) : make_float2(0.f, 0.f);
                    pre[4 * i + 2] = valid ? __fmul2_rn(__fadd2_rn(pre[4 * i + 2], p2), sc) : make_float2(0.f, 0.f);
                    pre[4 * i + 3] = valid ? __fmul2_rn(__fadd2_rn(pre[4 * i + 3], p3), sc) : make_float2(0.f, 0.f);
                }
            }
            ftick(0);           // loads issued and consumed
            if (j > 0) {        // R1 is free once the workers' E3 of the previous tile has read D3 out of it (y1 / dz1 reach the MMAs as images)
                tc::mbar_wait(&bars[B_Z1_FULL], (uint32_t)(j - 1) & 1u);
                tc::tc_fence_after();
            }
            ftick(1);           // idle: waiting for the workers
#pragma unroll
            for (int c = 0; c < 8; ++c) tc::tmem_st16(t_row + COL_R1 + 16 * c, pre + 8 * c);
            tc::tmem_wait_st();
            tc::tc_fence_before();
            warp_arrive(&bars[B_A_FULL], lane);
            ftick(2);           // pre-load stored
            if (j > 0) {
                tc::mbar_wait(&bars[B_D4_FULL], (uint32_t)(j - 1) & 1u);
                tc::tc_fence_after();
                ftick(3);       // rest of G4
                e4(j - 1);
                ftick(4);       // E4
            }
        }
        if (PROFILE && tid == NW && a.prof != nullptr)
            for (int i = 0; i < 6; ++i) a.prof[blockIdx.x * 24 + 12 + i] = ft[i];
        if (my_tiles > 0) {
            tc::mbar_wait(&bars[B_D4_FULL], (uint32_t)(my_tiles - 1) & 1u);
            tc::tc_fence_after();
            e4(my_tiles - 1);
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REG_AUX));
        if (warp > (NW + 128) / 32) {
            // =========================== copy warps: dz1 image -> fp32 rows of the scratch buffer (input of dproj_gather_kernel) ===========================
            // lane = edge row of a 32-row group: 16-byte reads of consecutive rows are conflict free; every lane writes the eight
            // channels of a chunk as one 32-byte sector.  Off the workers' chain: the stores only have to leave before E1 of the
            // next tile overwrites the image.
            const int cw = warp - (NW + 128) / 32 - 1;      // 0..2; row groups: 0 -> {0, 3}, 1 -> {1}, 2 -> {2}
            const float2 iS2 = make_float2(inv_S, inv_S);
            for (int j = 0; j < my_tiles; ++j) {
                const int tile = (int)blockIdx.x + j * G;
                tc::mbar_wait(&bars[B_Z1_FULL], (uint32_t)j & 1u);
                for (int grp = cw; grp < 4; grp += 3) {
                    const int r = 32 * grp + lane;
                    const long long e = (long long)tile * TM + r;
                    if (e < a.n_edges) {
                        float* o = a.dz1_out + (size_t)e * H;
#pragma unroll 4
                        for (int c8 = 0; c8 < H / 8; ++c8) {
                            const uint4 yh = yimg[c8 * TM + r], yl = yimg[YI + c8 * TM + r];
                            const float2 v0 = __fmul2_rn(__fadd2_rn(f16::unpack(yh.x), f16::unpack(yl.x)), iS2);
                            const float2 v1 = __fmul2_rn(__fadd2_rn(f16::unpack(yh.y), f16::unpack(yl.y)), iS2);
                            const float2 v2 = __fmul2_rn(__fadd2_rn(f16::unpack(yh.z), f16::unpack(yl.z)), iS2);
                            const float2 v3 = __fmul2_rn(__fadd2_rn(f16::unpack(yh.w), f16::unpack(yl.w)), iS2);
                            stg256(o + 8 * c8, v0, v1, v2, v3);
                        }
                    }
                }
                warp_arrive(&bars[B_YIMG_READ], lane);
            }
        } else if (warp == (NW + 128) / 32 && lane == 0) {
            // =========================== MMA issue ===========================
            constexpr uint32_t ID_G1 = f16::idesc(TM, H), ID_G2 = f16::idesc(TM, CN);
            constexpr uint32_t ID_G3 = idesc_mn(TM, H, 0, 1), ID_G4 = idesc_mn(TM, CE, 0, 1);
            constexpr uint32_t ID_W2 = idesc_mn(H, CN, 1, 1), ID_WE = idesc_mn(H, NE, 1, 1);
            const uint32_t sWe = tc::smem_u32(wsm), sW2 = tc::smem_u32(wsm + 2 * W1_WORDS);
            const uint32_t sY = tc::smem_u32(yimg), sZ = tc::smem_u32(zimg), sE = tc::smem_u32(eimg);
            bool wacc = false;
            long long mt[6] = {0, 0, 0, 0, 0, 0};
            auto mwait = [&](int slot_i, uint64_t* bar, uint32_t par) {
                if (PROFILE) {
                    const long long t0 = clock64();
                    tc::mbar_wait(bar, par);
                    mt[slot_i] += clock64() - t0;
                } else {
                    tc::mbar_wait(bar, par);
                }
            };
            // ---- G1 of tile jj: R1 (pre-loaded by the F role) += emb W_e^T ----
            auto g1 = [&](int jj) {
                const uint32_t xc = tmem + ((jj & 1) ? COL_X1 : COL_X0);
                mwait(0, &bars[B_A_FULL], (uint32_t)jj & 1u);
                tc::tc_fence_after();
                for (int p = 0; p < np; ++p) {      // small terms first: lo*hi, hi*lo, then hi*hi
                    const int pa = (np == 1) ? 0 : (p == 0 ? 1 : 0), pb = (np == 1) ? 0 : (p == 1 ? 1 : 0);
                    const uint64_t bd = tc::smem_desc(sWe + pb * (W1_WORDS * 4), H * 16, 128);
#pragma unroll
                    for (int ks = 0; ks < CE / 16; ++ks)
                        f16::mma_ts(tmem + COL_R1, xc + (pa ? 32u : 0u) + ks * 8, bd + (uint64_t)((ks * 2 * H * 16) >> 4), ID_G1, true);
                }
                tc::mma_commit(&bars[B_D1_FULL]);
            };
            if (my_tiles > 0) g1(0);
            for (int j = 0; j < my_tiles; ++j) {
                const uint32_t ph = (uint32_t)j & 1u;
                const uint32_t xcol = tmem + ((j & 1) ? COL_X1 : COL_X0);
                // ---- G2: R3 = y1 W_2^T ----
                mwait(1, &bars[B_Y1_FULL], ph);
                if (j > 0) mwait(2, &bars[B_D4_READ], ph ^ 1u);      // the F role has read D4 of the previous tile out of R3
                tc::tc_fence_after();
                {
                    bool acc = false;
                    for (int p = 0; p < np; ++p) {
                        const int pa = (np == 1) ? 0 : (p == 0 ? 1 : 0), pb = (np == 1) ? 0 : (p == 1 ? 1 : 0);
                        const uint64_t bd = tc::smem_desc(sW2 + pb * (W2_WORDS * 4), CN * 16, 128);
                        const uint64_t ad = tc::smem_desc(sY + pa * (Y_WORDS * 4), TM * 16, 128);      // the y1 image read K-major (rows = edges)
#pragma unroll
                        for (int ks = 0; ks < H / 16; ++ks) {
                            f16::mma_ss(tmem + COL_R3, ad + (uint64_t)((ks * 2 * TM * 16) >> 4), bd + (uint64_t)((ks * 2 * CN * 16) >> 4), ID_G2, acc);
                            acc = true;
                        }
                    }
                }
                tc::mma_commit(&bars[B_D2_FULL]);
                // ---- G3: R1 = dz2 W_2 (W_2's forward image read MN-major: N = H, K = CN) ----
                mwait(3, &bars[B_Z2_FULL], ph);
                tc::tc_fence_after();
                {
                    bool acc = false;
                    for (int p = 0; p < np; ++p) {
                        const int pa = (np == 1) ? 0 : (p == 0 ? 1 : 0), pb = (np == 1) ? 0 : (p == 1 ? 1 : 0);
                        const uint64_t bd = tc::smem_desc(sW2 + pb * (W2_WORDS * 4), 128, CN * 16);
#pragma unroll
                        for (int ks = 0; ks < CN / 16; ++ks) {
                            f16::mma_ts(tmem + COL_R1, xcol + (pa ? 32u : 0u) + ks * 8, bd + (uint64_t)((ks * 256) >> 4), ID_G3, acc);
                            acc = true;
                        }
                    }
                }
                tc::mma_commit(&bars[B_D3_FULL]);
                // ---- dW2 += y1^T dz2: A = y1 image (M = H channels), B = dz2 image (N = CN), K = the tile's 128 edges ----
                for (int p = 0; p < np; ++p) {
                    const int pa = (np == 1) ? 0 : (p == 0 ? 1 : 0), pb = (np == 1) ? 0 : (p == 1 ? 1 : 0);
                    const uint64_t ad = tc::smem_desc(sY + pa * (Y_WORDS * 4), 128, TM * 16);
                    const uint64_t bd = tc::smem_desc(sZ + pb * (Z_WORDS * 4), 128, TM * 16);
#pragma unroll
                    for (int ks = 0; ks < TM / 16; ++ks) {
                        f16::mma_ss(tmem + COL_ACC2, ad + (uint64_t)((ks * 256) >> 4), bd + (uint64_t)((ks * 256) >> 4), ID_W2, wacc || p > 0 || ks > 0);
                    }
                }
                tc::mma_commit(&bars[B_W2_DONE]);
                // ---- G4: R3 = dz1 W_e (W_e's forward image read MN-major: N = CE, K = H) ----
                mwait(4, &bars[B_Z1_FULL], ph);
                tc::tc_fence_after();
                {
                    bool acc = false;
                    for (int p = 0; p < np; ++p) {
                        const int pa = (np == 1) ? 0 : (p == 0 ? 1 : 0), pb = (np == 1) ? 0 : (p == 1 ? 1 : 0);
                        const uint64_t bd = tc::smem_desc(sWe + pb * (W1_WORDS * 4), 128, H * 16);
                        const uint64_t ad = tc::smem_desc(sY + pa * (Y_WORDS * 4), TM * 16, 128);      // the dz1 image read K-major
#pragma unroll
                        for (int ks = 0; ks < H / 16; ++ks) {
                            f16::mma_ss(tmem + COL_R3, ad + (uint64_t)((ks * 2 * TM * 16) >> 4), bd + (uint64_t)((ks * 256) >> 4), ID_G4, acc);
                            acc = true;
                        }
                    }
                }
                tc::mma_commit(&bars[B_D4_FULL]);
                // G1 of the NEXT tile goes in front of this tile's dWe: the workers' chain continues with its result, dWe only
                // has to be complete before E1 of the next tile writes the y1 image
                if (j + 1 < my_tiles) g1(j + 1);
                // ---- dWe += dz1^T [emb | 1] ----
                for (int p = 0; p < np; ++p) {
                    const int pa = (np == 1) ? 0 : (p == 0 ? 1 : 0), pb = (np == 1) ? 0 : (p == 1 ? 1 : 0);
                    const uint64_t ad = tc::smem_desc(sY + pa * (Y_WORDS * 4), 128, TM * 16);
                    const uint64_t bd = tc::smem_desc(sE + pb * (E_WORDS * 4), 128, TM * 16);
#pragma unroll
                    for (int ks = 0; ks < TM / 16; ++ks) {
                        f16::mma_ss(tmem + COL_ACCE, ad + (uint64_t)((ks * 256) >> 4), bd + (uint64_t)((ks * 256) >> 4), ID_WE, wacc || p > 0 || ks > 0);
                    }
                }
                tc::mma_commit(&bars[B_WE_DONE]);
                wacc = true;
            }
            if (PROFILE && a.prof != nullptr)
                for (int i = 0; i < 6; ++i) a.prof[blockIdx.x * 24 + 18 + i] = mt[i];
        }
        __syncwarp();
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// max |x| over n floats -> *out (float bits, non-negative: ordered like unsigned integers); *out must be zeroed first
__global__ void absmax_kernel(const float* __restrict__ x, size_t n, unsigned* __restrict__ out) {
    float m = 0.f;
    const size_t n4 = n / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(__ldg(x + i)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f && m < 3.0e38f) atomicMax(out, __float_as_uint(m));
}

__global__ void dproj_gather_kernel(const float* __restrict__ dz1, const int* __restrict__ row_ptr, const int* __restrict__ sptr,
                                    const int* __restrict__ slist, int n_nodes, int H, float* __restrict__ dP);

static int g_f16_bwd = 1;
static int g_f16_bwd_profile = 0;

bool mp_bwd_f16_supported(const ConvDims& d) { return g_f16_bwd && mp_f16_supported(d) && mp_bwd_tc_supported(d); }

int run_conv_edges_bwd_f16(const rgnn_conv& c, const ConvDims& d, const rgnn_graph& g, const uint32_t* emb_hl, const float* P,
                           const float* dagg, float* dP, float* demb, bool first_demb, float* scratch, const int* sptr,
                           const int* slist, cudaStream_t stream) {
    const rgnn_linear& m0 = c.msg.layer[0];
    const rgnn_linear& m1 = c.msg.layer[1];
    const size_t E = (size_t)g.n_edges;
    float* dz1 = scratch;
    unsigned* gmax = reinterpret_cast<unsigned*>(scratch + E * d.h);
    static PerDeviceOnce once;
    if (once.needed()) {
        RGNN_CHECK_CUDA(cudaFuncSetAttribute(mp_edge_bwd_f16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mbf::SMEM));
        RGNN_CHECK_CUDA(cudaFuncSetAttribute(mp_edge_bwd_f16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mbf::SMEM));
        once.mark();
    }
    RGNN_CHECK_CUDA(cudaMemsetAsync(gmax, 0, sizeof(unsigned), stream));
    {
        const size_t n = (size_t)g.n_nodes * d.cn;
        size_t blocks = (n / 4 + 255) / 256;
        if (blocks > (size_t)4 * sm_count()) blocks = (size_t)4 * sm_count();
        if (blocks < 1) blocks = 1;
        absmax_kernel<<<(unsigned)blocks, 256, 0, stream>>>(dagg, n, gmax);
    }
    MpBwdF16Args a;
    a.emb = emb_hl; a.P = P; a.dagg = dagg; a.tgt = g.tgt; a.src = g.src;
    a.wpack = reinterpret_cast<const uint32_t*>(m0.weight_t + conv_msg0_f16_offset(d));
    a.s1 = m0.norm_scale; a.m1 = m0.norm_shift;
    a.b2 = m1.bias; a.s2 = m1.norm_scale; a.m2 = m1.norm_shift;
    a.gmax = reinterpret_cast<const float*>(gmax);
    a.dz1_out = dz1; a.demb = demb;
    a.gW2 = m1.grad_weight; a.gb2 = m1.grad_bias;
    a.gWe = m0.grad_weight ? m0.grad_weight + 2 * d.cn : nullptr;
    a.ldWe = m0.in_features;
    a.gb1 = m0.grad_bias;
    a.g_s1 = m0.grad_norm_scale; a.g_m1 = m0.grad_norm_shift; a.g_s2 = m1.grad_norm_scale; a.g_m2 = m1.grad_norm_shift;
    a.n_edges = g.n_edges; a.act1 = m0.activation; a.act2 = m1.activation;
    a.demb_accumulate = first_demb ? 0 : 1;
    a.passes = mp_f16_passes();
    a.prof = nullptr;
    const int n_tiles = (g.n_edges + mbf::TM - 1) / mbf::TM;
    const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
    if (g_f16_bwd_profile) {     // developer aid (rgnn_set_option("debug", 8)): per-phase cycles of worker thread 0; synchronises
        long long* prof = nullptr;
        RGNN_CHECK_CUDA(cudaMalloc(&prof, sizeof(long long) * 24 * grid));
        RGNN_CHECK_CUDA(cudaMemsetAsync(prof, 0, sizeof(long long) * 24 * grid, stream));
        a.prof = prof;
        mp_edge_bwd_f16_kernel<true><<<grid, mbf::NTHREADS, mbf::SMEM, stream>>>(a);
        RGNN_CHECK_CUDA(cudaStreamSynchronize(stream));
        std::vector<long long> h(24 * grid);
        RGNN_CHECK_CUDA(cudaMemcpy(h.data(), prof, sizeof(long long) * 24 * grid, cudaMemcpyDeviceToHost));
        double tot[24] = {0};
        for (int b = 0; b < grid; ++b) for (int i = 0; i < 24; ++i) tot[i] += (double)h[b * 24 + i];
        static const char* nm[24] = {"waitG1", "E1", "waitG2", "E2", "waitG3", "E3a", "wait_dW2", "-", "E3b", "dz1 stores", "-", "-",
                                     "F:loads", "F:idle", "F:preload", "F:waitG4", "F:E4", "-",
                                     "MMA:wait A", "MMA:wait y1", "MMA:wait D4 read", "MMA:wait dz2", "MMA:wait dz1", "-"};
        fprintf(stderr, "[mp_edge_bwd_f16 profile] cycles per tile:");
        for (int i = 0; i < 24; ++i) if (nm[i][0] != '-') fprintf(stderr, " %s=%.0f", nm[i], tot[i] / (double)n_tiles);
        fprintf(stderr, "\n");
        cudaFree(prof);
    } else {
        mp_edge_bwd_f16_kernel<false><<<grid, mbf::NTHREADS, mbf::SMEM, stream>>>(a);
    }
    RGNN_CHECK_CUDA(cudaGetLastError());
    const int wpb = 8;
    const int blocks = (g.n_nodes + wpb - 1) / wpb;
    dproj_gather_kernel<<<blocks > 8 * sm_count() ? 8 * sm_count() : blocks, 32 * wpb, 0, stream>>>(dz1, g.row_ptr, sptr, slist, g.n_nodes, d.h, dP);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

int mp_bwd_f16_set_option(const char* name, int value) {
    if (strcmp(name, "f16_bwd") == 0 && (value == 0 || value == 1)) { g_f16_bwd = value; return 1; }
    if (strcmp(name, "debug") == 0) { g_f16_bwd_profile = (value & 8) != 0; return 0; }
    return 0;
}
int mp_bwd_f16_get_option(const char* name) {
    if (strcmp(name, "f16_bwd") == 0) return g_f16_bwd;
    return -2;
}

}  // namespace rgnn
