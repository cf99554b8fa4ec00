// Device helpers shared by the tensor-core kernels whose epilogues run thread-per-row out of TMEM
// (rgnn_mp_tc.cu, rgnn_rowmlp_tc.cu): named barriers, the cross-thread row reduction and channel_normalization.
#pragma once
#include "rgnn_common.cuh"
#include "rgnn_tc.cuh"

namespace rgnn {

__device__ __forceinline__ void group_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// sum of one value per thread over the NQ threads that share an edge row (same lane, different warps): the partials
// meet in NQ spare TMEM columns of the row's lane.
template <int NQ>
__device__ __forceinline__ float row_allreduce(float v, uint32_t t_cols, int q, int bar_id) {
    tc::tmem_st1(t_cols + q, v);
    tc::tmem_wait_st();
    tc::tc_fence_before();
    group_sync(bar_id, 32 * NQ);
    tc::tc_fence_after();
    float p0, p1, p2 = 0.f, p3 = 0.f;
    if (NQ == 4) tc::tmem_ld4(t_cols, p0, p1, p2, p3); else tc::tmem_ld2(t_cols, p0, p1);
    tc::tmem_wait_ld();
    return (p0 + p1) + (p2 + p3);
}

// channel_normalization + LeakyReLU on a row whose C columns are spread over NQ threads (2*CP each, held as
// register pairs for the packed f32x2 pipe).  Mean / unbiased std as the reference's two-pass formula (common.py:215-220).
// NQ == 2: each thread centres its own half on its own mean (two register-only passes) and the halves are merged by
// the exact pairwise update  mean = (m0 + m1) / 2,  M2 = M2_0 + M2_1 + n (m0 - m1)^2 / 2  -- ONE exchange through TMEM
// instead of two (every exchange is a tcgen05.st / wait / named barrier / tcgen05.ld / wait round trip on the critical
// path of the epilogue), with the accuracy of the two-pass formula (no sum-of-squares cancellation).  The shift from
// the own mean to the row mean is folded into the final affine.
__device__ __forceinline__ void row_allreduce2(float& u, float& v, uint32_t t_cols, int q, int bar_id);
template <int CP, int NQ>
__device__ __forceinline__ void row_norm_act(float2 (&z)[CP], int C, bool has_norm, float scale, float shift,
                                             bool act, uint32_t t_cols /* 2*NQ spare TMEM columns of this row */, int q, int bar_id,
                                             float* sd_out = nullptr /* receives the row's sigma */) {
    // scale / shift are passed BY VALUE: with ~226 KB of shared memory per CTA the L1 is a few KB, and a global load
    // inside this dependent chain costs an L2 round trip per row tile
    if (has_norm) {
        float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < CP; ++c) s2 = __fadd2_rn(s2, z[c]);
        if (NQ == 2 && C == 4 * CP) {      // both threads own exactly C / 2 columns
            const float m_own = (s2.x + s2.y) * (1.f / (float)(2 * CP));
            const float2 nm = make_float2(-m_own, -m_own);
            float2 ss2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                z[c] = __fadd2_rn(z[c], nm);
                ss2 = __ffma2_rn(z[c], z[c], ss2);
            }
            float msum = m_own, m2 = ss2.x + ss2.y;
            row_allreduce2(msum, m2, t_cols, q, bar_id);        // msum = m0 + m1, m2 = M2_0 + M2_1
            const float delta = m_own - 0.5f * msum;            // own mean - row mean = (m_own - m_other) / 2
            const float ss = fmaf((float)(4 * CP) * delta, delta, m2);   // + n (m0 - m1)^2 / 2 with n = 2 CP, (m0 - m1) = 2 delta
            const float sd = sqrtf(ss / (float)(C - 1));
            if (sd_out != nullptr) *sd_out = sd;
            const float k = scale / (sd + NORM_EPS);
            const float2 k2 = make_float2(k, k), sh2 = make_float2(fmaf(delta, k, shift), fmaf(delta, k, shift));
#pragma unroll
            for (int c = 0; c < CP; ++c) z[c] = __ffma2_rn(z[c], k2, sh2);
        } else if (NQ == 4 && C == 8 * CP) {   // four threads own C / 4 columns each: the same merge over four partitions
            const float m_own = (s2.x + s2.y) * (1.f / (float)(2 * CP));
            const float2 nm = make_float2(-m_own, -m_own);
            float2 ss2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                z[c] = __fadd2_rn(z[c], nm);
                ss2 = __ffma2_rn(z[c], z[c], ss2);
            }
            asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(t_cols + 2 * q), "f"(m_own), "f"(ss2.x + ss2.y) : "memory");
            tc::tmem_wait_st();
            tc::tc_fence_before();
            group_sync(bar_id, 128);
            tc::tc_fence_after();
            float m0, v0, m1, v1, m2, v2, m3, v3;
            tc::tmem_ld4(t_cols, m0, v0, m1, v1);
            tc::tmem_ld4(t_cols + 4, m2, v2, m3, v3);
            tc::tmem_wait_ld();
            const float mean = 0.25f * ((m0 + m1) + (m2 + m3));
            const float d0 = m0 - mean, d1 = m1 - mean, d2 = m2 - mean, d3 = m3 - mean;
            const float ss = fmaf((float)(2 * CP), (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3), (v0 + v1) + (v2 + v3));
            const float delta = m_own - mean;
            const float sd = sqrtf(ss / (float)(C - 1));
            if (sd_out != nullptr) *sd_out = sd;
            const float k = scale / (sd + NORM_EPS);
            const float2 k2 = make_float2(k, k), sh2 = make_float2(fmaf(delta, k, shift), fmaf(delta, k, shift));
#pragma unroll
            for (int c = 0; c < CP; ++c) z[c] = __ffma2_rn(z[c], k2, sh2);
        } else {
            const float mean = row_allreduce<NQ>(s2.x + s2.y, t_cols, q, bar_id) / (float)C;
            const float2 nm = make_float2(-mean, -mean);
            float2 ss2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                z[c] = __fadd2_rn(z[c], nm);
                ss2 = __ffma2_rn(z[c], z[c], ss2);
            }
            const float ss = row_allreduce<NQ>(ss2.x + ss2.y, t_cols + NQ, q, bar_id);
            const float sd = sqrtf(ss / (float)(C - 1));
            if (sd_out != nullptr) *sd_out = sd;
            const float k = scale / (sd + NORM_EPS);
            const float2 k2 = make_float2(k, k), sh2 = make_float2(shift, shift);
#pragma unroll
            for (int c = 0; c < CP; ++c) z[c] = __ffma2_rn(z[c], k2, sh2);
        }
    }
    if (act) {   // LeakyReLU(0.01): max(v, 0.01 v)
        const float2 sl = make_float2(LEAKY, LEAKY);
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            const float2 t = __fmul2_rn(z[c], sl);
            z[c].x = fmaxf(z[c].x, t.x);
            z[c].y = fmaxf(z[c].y, t.y);
        }
    }
}

// two sums per row over the 2 threads that share it, exchanged through 4 spare TMEM columns
__device__ __forceinline__ void row_allreduce2(float& u, float& v, uint32_t t_cols, int q, int bar_id) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(t_cols + 2 * q), "f"(u), "f"(v) : "memory");
    tc::tmem_wait_st();
    tc::tc_fence_before();
    group_sync(bar_id, 64);
    tc::tc_fence_after();
    float p0, p1, p2, p3;
    tc::tmem_ld4(t_cols, p0, p1, p2, p3);
    tc::tmem_wait_ld();
    u = p0 + p2;
    v = p1 + p3;
}

// Backward of LeakyReLU + the inverse of the scalar affine for one register pair (packed f32x2):
//   g <- g * act'(y),   nv <- (act^-1(y) - shift) / scale     with y the layer output, neg_shift_is = -shift / scale
__device__ __forceinline__ void act_bwd_pair(float2& g, float2& nv, float2 y, bool act, float2 inv_scale2, float2 neg_shift_is) {
    const bool px = !act || y.x > 0.f, py = !act || y.y > 0.f;
    g = __fmul2_rn(g, make_float2(px ? 1.f : LEAKY, py ? 1.f : LEAKY));
    nv = __ffma2_rn(__fmul2_rn(y, make_float2(px ? 1.f : 1.f / LEAKY, py ? 1.f : 1.f / LEAKY)), inv_scale2, neg_shift_is);
}

__device__ __forceinline__ void stg256(float* p, float2 a, float2 b, float2 c, float2 d) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y), "f"(d.x), "f"(d.y)
                 : "memory");
}

// named barriers (0 = __syncthreads, 1..4 = the NQ warps sharing 32 rows)
constexpr int BAR_WORKERS = 5;   // all worker warps
constexpr int BAR_A_READY = 6;   // workers arrive, MMA warp waits: A operand (smem) written
constexpr int BAR_Y_READY = 7;   // workers arrive, MMA warp waits: y1 (TMEM) written
__device__ __forceinline__ void bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// 32 bytes per lane and request: a full L2 sector (the 16-byte form fetches the sector twice when L1 is tiny)
__device__ __forceinline__ void ldg256(const float* p, float2& a, float2& b, float2& c, float2& d) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(b.x), "=f"(b.y), "=f"(c.x), "=f"(c.y), "=f"(d.x), "=f"(d.y)
                 : "l"(p));
}

}  // namespace rgnn
