// fp16-split operands for tcgen05.mma.kind::f16 (sm_100a inline PTX), shared by the f16 kernels (rgnn_mp_f16.cu, ...).
//
// fp32 parity on the tensor cores with half the MMA time and half the TMEM columns of 3xTF32:
//   x * S = hi + lo,   hi = fp16(x S) (11 significant bits),  lo = fp16(x S - hi) (the next 11)
//   D = A_lo B_hi + A_hi B_lo + A_hi B_hi    (every product of two fp16 values is exact in the fp32 accumulator)
// S is a power of two that keeps `lo` out of the fp16 subnormal range (activations x 16, weights x 256): the
// representation error is 2^-22 |x| like 3xTF32, the accumulator holds 4096 x the true dot product, and every consumer
// either normalises (scale invariant: sigma and eps scale with it) or multiplies by 2^-12 inside an FFMA it executes anyway.
// One kind::f16 instruction covers K = 16 (32 bytes per row, like K = 8 of tf32), so the hi*hi chain that sets the
// round-toward-zero bias of the accumulator (tools/mma_noise.py) has half as many steps as with tf32.
// `passes = 1` keeps only A_hi B_hi: plain fp16 operands, fp32 accumulate (the reduced-precision mode, ~2.5e-4 per GEMM).
//
// Operand layouts
//   shared memory (B operand, K-major, SWIZZLE_NONE): uint4 op[K/8][rows], i.e. byte offset ((k / 8) * rows + r) * 16 + (k % 8) * 2
//       = the canonical layout ((8,m),(8,2)) with SBO = 128 B (next 8 rows) and LBO = rows * 16 B (next 8 K elements);
//       byte for byte the same descriptor arithmetic as the tf32 images of rgnn_tc.cuh.
//   tensor memory (A operand): lane = row, 32-bit column c holds (A[r][2c], A[r][2c+1]) in its (low, high) half;
//       a K = 16 step reads 8 consecutive columns.
#pragma once
#include <cuda_fp16.h>

#include "rgnn_common.cuh"
#include "rgnn_tc.cuh"

namespace rgnn {
namespace f16 {

constexpr float A_SCALE = 16.f;        // activations
constexpr float W_SCALE = 256.f;       // weights
constexpr float D_SCALE = A_SCALE * W_SCALE;       // what the accumulator holds relative to the true dot product
constexpr float D_UNSCALE = 1.f / D_SCALE;

// instruction descriptor: kind::f16 with fp16 A and B (format 0), fp32 accumulate, both operands K-major, dense
__host__ __device__ constexpr uint32_t idesc(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc_, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc_), "r"((uint32_t)accumulate)
        : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc_, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc_), "r"((uint32_t)accumulate)
        : "memory");
}

// two fp32 values -> packed fp16 pair (x0 in the low half), saturating at +-65504 instead of overflowing to infinity
__device__ __forceinline__ uint32_t pack_sat(float x0, float x1) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x1), "f"(x0));
    return r;
}
__device__ __forceinline__ float2 unpack(uint32_t h) {
    return __half22float2(*reinterpret_cast<const __half2*>(&h));
}
// (x0, x1), already multiplied by the operand scale -> hi and lo pairs
__device__ __forceinline__ void split(float2 x, uint32_t& hi, uint32_t& lo) {
    hi = pack_sat(x.x, x.y);
    const float2 h = unpack(hi);
    const float2 d = __fadd2_rn(x, make_float2(-h.x, -h.y));       // one packed subtraction (same fp32 results as two FADDs)
    lo = pack_sat(d.x, d.y);
}

// 16 packed columns (= 32 fp16 values of this lane's row) <-> tensor memory
__device__ __forceinline__ void tmem_st16u(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8u(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

// 32 bytes per lane and request (a full L2 sector), as eight 32-bit words
__device__ __forceinline__ void ldg256u(const void* p, uint32_t* r) {
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
}
__device__ __forceinline__ void stg256u(void* p, const uint32_t* r) {
    asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

// non-blocking mbarrier probe (the MMA warp polls several barriers and issues whatever is ready)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(tc::smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}

// TS GEMM of one layer, issued by ONE lane: D[dcol] (+)= A-format operand at acol (K columns) * image at sB (N rows).  Shapes are template
// parameters and the loops are unrolled: with runtime K / N the lane spent ~85 cycles of address arithmetic per MMA (64-bit descriptor adds,
// loop control in one dependent instruction stream) where an MMA takes 34 - 68, and the tensor pipe waited for it (measured with cycle
// counters in the lane: 18.9 k -> 16.3 k cycles of issue per tile pair of 30 k).
template <int K, int N>
__device__ __forceinline__ void gemm_ts(uint32_t dcol, uint32_t acol, uint32_t sB, uint32_t idesc_, bool acc0, int np) {
    constexpr uint32_t lbo = (uint32_t)N * 16, img = (uint32_t)K * N * 2;
    bool acc = acc0;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        if (p < np) {
            const int pa = (np == 1) ? 0 : (p == 0 ? 1 : 0);
            const int pb = (np == 1) ? 0 : (p == 1 ? 1 : 0);
            const uint64_t bd0 = tc::smem_desc(sB + pb * img, lbo, 128);
            const uint32_t a0 = acol + (pa ? 16u : 0u);
#pragma unroll
            for (int ks = 0; ks < K / 16; ++ks) {
                mma_ts(dcol, a0 + (ks >> 1) * 32 + (ks & 1) * 8, bd0 + (uint64_t)((ks * 2 * lbo) >> 4), idesc_, acc);
                acc = true;
            }
        }
    }
}

}  // namespace f16

// TILED layout of the pre-split edge rows (what the edge encoder writes and the message kernels read): per tile of 128 edges one
// 32 KB block  [hi | lo][chunk of 8 channels: 8][edge: 128][8 fp16]  -- byte for byte the chunk-major operand image of a tile.
// A thread that owns edge row r moves 16 bytes per chunk and the 32 lanes of a warp touch 512 CONTIGUOUS bytes (4 L1 wavefronts
// per instruction instead of the 32 of a row-per-thread 32-byte access); a whole image is one bulk copy.
constexpr int EMB_TILE_WORDS = 128 * 64;
__host__ __device__ __forceinline__ size_t emb_tile_word(long long e, int img, int c8) {
    return (size_t)(e >> 7) * EMB_TILE_WORDS + (size_t)img * (EMB_TILE_WORDS / 2) + ((size_t)c8 * 128 + (size_t)(e & 127)) * 4;
}
__device__ __forceinline__ uint4 ldg128u(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// streaming variant: the data is used once and should not push re-used rows out of the (small) L1
__device__ __forceinline__ uint4 ldg128u_na(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// one elected lane per warp arrives for its 32 rows (the warp-collective tcgen05.wait / fences come first)
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(bar);
}

// Row statistics of channel_normalization for a row that is visited in chunks of 32 values (16 register pairs): every
// chunk is centred on its own mean (two register-only passes, like the reference's two-pass formula on that chunk) and
// the chunks are merged with the exact pairwise update  mean = mean_a + d n_b / n,  M2 = M2_a + M2_b + d^2 n_a n_b / n
// (d = mean_b - mean_a): no sum-of-squares cancellation, and the loop body stays small enough for the instruction cache.
struct RowStats {
    float mean, m2, n;
    __device__ __forceinline__ void init() { mean = 0.f; m2 = 0.f; n = 0.f; }
    __device__ __forceinline__ void add_chunk(const float2 (&v)[16]) {
        float2 s0 = v[0], s1 = v[1], s2 = v[2], s3 = v[3];
#pragma unroll
        for (int c = 4; c < 16; c += 4) {
            s0 = __fadd2_rn(s0, v[c]); s1 = __fadd2_rn(s1, v[c + 1]); s2 = __fadd2_rn(s2, v[c + 2]); s3 = __fadd2_rn(s3, v[c + 3]);
        }
        const float2 st = __fadd2_rn(__fadd2_rn(s0, s1), __fadd2_rn(s2, s3));
        const float mc = (st.x + st.y) * (1.f / 32.f);
        const float2 nm = make_float2(-mc, -mc);
        float2 q0 = make_float2(0.f, 0.f), q1 = q0, q2 = q0, q3 = q0;
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
            const float2 d0 = __fadd2_rn(v[c], nm), d1 = __fadd2_rn(v[c + 1], nm), d2 = __fadd2_rn(v[c + 2], nm), d3 = __fadd2_rn(v[c + 3], nm);
            q0 = __ffma2_rn(d0, d0, q0); q1 = __ffma2_rn(d1, d1, q1); q2 = __ffma2_rn(d2, d2, q2); q3 = __ffma2_rn(d3, d3, q3);
        }
        const float2 qt = __fadd2_rn(__fadd2_rn(q0, q1), __fadd2_rn(q2, q3));
        const float m2c = qt.x + qt.y;
        const float nn = n + 32.f;
        const float d = mc - mean;
        const float w = 32.f / nn;                   // exact: n is a multiple of 32 up to 128
        m2 = m2 + m2c + d * d * (n * w);
        mean = fmaf(d, w, mean);
        n = nn;
    }
    __device__ __forceinline__ float sigma(int count) const { return __fsqrt_rn(m2 * (1.f / (float)(count - 1))); }
};

}  // namespace rgnn
