// Tile-resident building blocks shared by the forward and backward tile programs.
//
// A CTA (256 threads) owns a tile of TR rows (nodes, edges, undirected links or clusters).  The tile's
// activations live in shared memory from the first load to the last store; a layer is
//     tile_gemm      Y = X * Wt (+bias)            X,Y in smem, Wt streamed from L2 through a double-buffered stage
//     tile_norm_act  per-row channel norm + LeakyReLU in place
// so intermediate activations (in particular the per-edge ones) never reach HBM.
#pragma once
#include "rgnn_common.cuh"

namespace rgnn {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ float leaky(float v) { return v > 0.f ? v : LEAKY * v; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Copy rows [k0, k0+kn) x cols [cb, cb+bw) of the k-major operand Wt (row stride ldw) into a stage buffer laid
// out [KC][CBMAX].  Rows at or beyond k_valid and columns at or beyond c_exist (a multiple of 4) do not exist in memory and
// are zero-filled.
__device__ __forceinline__ void stage_weights(float* ws, const float* __restrict__ Wt, int ldw, int k0, int kn,
                                              int k_valid, int cb, int bw, int c_exist) {
    const int per_row = bw >> 2;
    const int tot = kn * per_row;
    for (int i = threadIdx.x; i < tot; i += NT) {
        const int r = i / per_row, c4 = i - r * per_row;
        float* d = ws + r * CBMAX + 4 * c4;
        if (k0 + r < k_valid && cb + 4 * c4 < c_exist)
            cp_async16(d, Wt + (size_t)(k0 + r) * ldw + cb + 4 * c4);
        else
            *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// One output-column block of Y[TR][ldy] = X[TR][ldx] * Wt[K][ldw] (+ bias[c], c < C).
// K multiple of 8 (X columns [K_true, K) are zero-filled by the producer), bw = 64 or 128.
// Thread (ty,tx) = (tid/16, tid%16) owns rows ty*RPT.. and columns cb + 4*tx.. (+64 when TWO).
template <int TR, bool TWO>
__device__ __forceinline__ void tile_gemm_block(const float* __restrict__ Xs, int ldx, int K, int k_valid,
                                                const float* __restrict__ Wt, int ldw, int cb, int bw,
                                                const float* __restrict__ bias, int C, float* __restrict__ Ys, int ldy,
                                                float* __restrict__ wstage, int c_exist) {
    constexpr int RPT = TR / 16;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[RPT][TWO ? 8 : 4];
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int j = 0; j < (TWO ? 8 : 4); ++j) acc[i][j] = 0.f;

    const int nch = (K + KC - 1) / KC;
    stage_weights(wstage, Wt, ldw, 0, min(KC, K), k_valid, cb, bw, c_exist);
    cp_async_commit();
    for (int ch = 0; ch < nch; ++ch) {
        if (ch + 1 < nch) {
            stage_weights(wstage + ((ch + 1) & 1) * KC * CBMAX, Wt, ldw, (ch + 1) * KC, min(KC, K - (ch + 1) * KC),
                          k_valid, cb, bw, c_exist);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* ws = wstage + (ch & 1) * KC * CBMAX;
        const int kn = min(KC, K - ch * KC);
        const float* xrow = Xs + (ty * RPT) * ldx + ch * KC;
#pragma unroll 2
        for (int kk = 0; kk < kn; kk += 4) {
            float4 xv[RPT];
#pragma unroll
            for (int i = 0; i < RPT; ++i) xv[i] = *reinterpret_cast<const float4*>(xrow + i * ldx + kk);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 w0 = *reinterpret_cast<const float4*>(ws + (kk + q) * CBMAX + 4 * tx);
                float4 w1 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (TWO) w1 = *reinterpret_cast<const float4*>(ws + (kk + q) * CBMAX + 64 + 4 * tx);
#pragma unroll
                for (int i = 0; i < RPT; ++i) {
                    const float x = q == 0 ? xv[i].x : (q == 1 ? xv[i].y : (q == 2 ? xv[i].z : xv[i].w));
                    acc[i][0] = fmaf(x, w0.x, acc[i][0]);
                    acc[i][1] = fmaf(x, w0.y, acc[i][1]);
                    acc[i][2] = fmaf(x, w0.z, acc[i][2]);
                    acc[i][3] = fmaf(x, w0.w, acc[i][3]);
                    if (TWO) {
                        acc[i][4] = fmaf(x, w1.x, acc[i][4]);
                        acc[i][5] = fmaf(x, w1.y, acc[i][5]);
                        acc[i][6] = fmaf(x, w1.z, acc[i][6]);
                        acc[i][7] = fmaf(x, w1.w, acc[i][7]);
                    }
                }
            }
        }
        __syncthreads();
    }
    const int c0 = cb + 4 * tx;
    float b[TWO ? 8 : 4];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = (bias != nullptr && c0 + j < C) ? __ldg(bias + c0 + j) : 0.f;
    if (TWO) {
#pragma unroll
        for (int j = 0; j < 4; ++j) b[4 + j] = (bias != nullptr && c0 + 64 + j < C) ? __ldg(bias + c0 + 64 + j) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        float* y = Ys + (ty * RPT + i) * ldy + c0;
        *reinterpret_cast<float4*>(y) = make_float4(acc[i][0] + b[0], acc[i][1] + b[1], acc[i][2] + b[2], acc[i][3] + b[3]);
        if (TWO)
            *reinterpret_cast<float4*>(y + 64) =
                make_float4(acc[i][4] + b[4], acc[i][5] + b[5], acc[i][6] + b[6], acc[i][7] + b[7]);
    }
}

// Y = X * Wt (+bias) over all Cpad output columns (Cpad multiple of 64).  Ends with __syncthreads().
template <int TR>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ Xs, int ldx, int K, int k_valid,
                                          const float* __restrict__ Wt, int ldw, int Cpad,
                                          const float* __restrict__ bias, int C, float* __restrict__ Ys, int ldy,
                                          float* __restrict__ wstage, int c_exist) {
    for (int cb = 0; cb < Cpad; cb += CBMAX) {
        const int bw = min(CBMAX, Cpad - cb);
        if (bw > 64)
            tile_gemm_block<TR, true>(Xs, ldx, K, k_valid, Wt, ldw, cb, bw, bias, C, Ys, ldy, wstage, c_exist);
        else
            tile_gemm_block<TR, false>(Xs, ldx, K, k_valid, Wt, ldw, cb, bw, bias, C, Ys, ldy, wstage, c_exist);
    }
    __syncthreads();
}

// In-place per-row channel_normalization (reference common.py:215-220: mean, UNBIASED std, eps added to the
// std, scalar affine) and LeakyReLU on a [TR][ld] tile.  C multiple of 32 (<= 256) when normalising.
// One warp per row; lane owns columns lane + 32 q.  sigma_out (nullable): the row's std, kept for the backward.
template <int TR>
__device__ __forceinline__ void tile_norm_act(float* __restrict__ Ys, int ld, int C, const float* __restrict__ scale_p,
                                              const float* __restrict__ shift_p, bool act,
                                              float* __restrict__ sigma_out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (scale_p != nullptr) {
        const float scale = __ldg(scale_p), shift = __ldg(shift_p);
        const int nq = C >> 5;
        for (int r = warp; r < TR; r += NT / 32) {
            float* y = Ys + r * ld;
            float v[16];       // rows of up to 512 channels (hidden width 256: msg.0 is 512 wide)
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                v[q] = q < nq ? y[lane + 32 * q] : 0.f;
                s += v[q];
            }
            const float mean = warp_sum(s) / (float)C;
            float ss = 0.f;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float d = q < nq ? v[q] - mean : 0.f;
                v[q] = d;
                ss = fmaf(d, d, ss);
            }
            const float sd = sqrtf(warp_sum(ss) / (float)(C - 1));
            const float den = sd + NORM_EPS;
            if (sigma_out != nullptr && lane == 0) sigma_out[r] = sd;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                if (q < nq) {
                    float o = fmaf(scale, __fdiv_rn(v[q], den), shift);
                    if (act) o = leaky(o);
                    y[lane + 32 * q] = o;
                }
            }
        }
    } else if (act) {
        for (int i = threadIdx.x; i < TR * C; i += NT) {
            const int r = i / C, j = i - r * C;
            Ys[r * ld + j] = leaky(Ys[r * ld + j]);
        }
    }
    __syncthreads();
}

}  // namespace rgnn
