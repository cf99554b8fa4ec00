// Weight-image packing, batched: after every SGD step the detector re-packs the operand images of its 49 Linears (tf32 hi | lo chunks
// for the row-MLP interpreter, fp16 hi | lo images for the fixed-shape kernels, the round-1 message-kernel images) -- ~214 launches of
// a few microseconds each, which is what a small-batch training step was waiting for.  Inside rgnn_pack_detector the launch sites push
// their arguments into a per-thread queue instead; the queue is flushed as ONE table-driven launch per kind and per 48 entries
// (blockIdx.y = entry).  Entries of one flush run concurrently: a site that rewrites an image an earlier entry wrote flushes first
// (edge_enc_f16_pack).  Element bodies are shared with the single-image kernels, so both paths write identical bytes
// (tests/test_cabi_cpu.py cannot run kernels; tests/test_model_gpu.py::test_batched_packing_is_bit_identical does).
#pragma once
#include <cuda_fp16.h>

#include "rgnn_f16.cuh"
#include "rgnn_tc.cuh"

namespace rgnn {

struct PackTcArgs { const float* W; float* dst; int ldW, n0, Nt, nd0, Np, k0, Kt, Kp, kc, n_loop, transpose; };
struct PackF16Args { const float* W; __half* hi; __half* lo; int off, sn, sk, K, N, n_valid, k_valid; };
struct PackSplitArgs { const float* W; float* hi; float* lo; int off, sn, sk, K, N; };

__device__ __forceinline__ void pack_tc_body(const PackTcArgs& a, int first, int stride) {
    const int tot = a.Kp * a.n_loop;
    for (int i = first; i < tot; i += stride) {
        const int k = i / a.n_loop, n = i - k * a.n_loop;
        float w = 0.f;
        if (k < a.Kt && n < a.Nt) w = a.transpose ? a.W[(size_t)(a.k0 + k) * a.ldW + a.n0 + n] : a.W[(size_t)(a.n0 + n) * a.ldW + a.k0 + k];
        float h, l;
        tc::split_tf32(w, h, l);
        const int chunk = k / a.kc, kk = k - chunk * a.kc;
        float* base = a.dst + (size_t)chunk * (2 * a.kc * a.Np);
        const int off = ((kk >> 2) * a.Np + a.nd0 + n) * 4 + (kk & 3);
        base[off] = h;
        base[a.kc * a.Np + off] = l;
    }
}

// weights: element (n, k) = W[off + n * sn + k * sk], x 256 -> fp16 hi / lo chunk-major images [K/8][N][8]
__device__ __forceinline__ void pack_f16_body(const PackF16Args& a, int first, int stride) {
    const int tot = a.K * a.N;
    for (int i = first; i < tot; i += stride) {
        const int q = i & 7, n = (i >> 3) % a.N, kc = (i >> 3) / a.N;
        const float w = (n < a.n_valid && 8 * kc + q < a.k_valid) ? a.W[(size_t)a.off + (size_t)n * a.sn + (size_t)(8 * kc + q) * a.sk] * f16::W_SCALE : 0.f;
        const uint32_t h2 = f16::pack_sat(w, 0.f);
        const float hf = f16::unpack(h2).x;
        const uint32_t l2 = f16::pack_sat(w - hf, 0.f);
        a.hi[i] = __ushort_as_half((unsigned short)(h2 & 0xFFFFu));
        a.lo[i] = __ushort_as_half((unsigned short)(l2 & 0xFFFFu));
    }
}

__device__ __forceinline__ void pack_split_body(const PackSplitArgs& a, int first, int stride) {
    const int tot = a.K * a.N;
    for (int i = first; i < tot; i += stride) {
        const int q = i & 3, n = (i >> 2) % a.N, kc = (i >> 2) / a.N;
        const float w = a.W[(size_t)a.off + (size_t)n * a.sn + (size_t)(4 * kc + q) * a.sk];
        float h, l;
        tc::split_tf32(w, h, l);
        a.hi[i] = h;
        a.lo[i] = l;
    }
}

// queue (rgnn_packq.cu): push returns false when no queue is open on this thread -- the caller then launches its own kernel
void packq_begin(cudaStream_t stream);
int packq_flush();
int packq_end();
bool packq_push(const PackTcArgs& a);
bool packq_push(const PackF16Args& a);
bool packq_push(const PackSplitArgs& a);
extern int g_pack_batch;        // rgnn_set_option("pack_batch", 0 / 1): 0 = one launch per image (the A / B partner of the test)

}  // namespace rgnn
