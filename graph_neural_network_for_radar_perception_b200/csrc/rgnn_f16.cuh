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
    lo = pack_sat(x.x - h.x, x.y - h.y);
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

}  // namespace f16
}  // namespace rgnn
