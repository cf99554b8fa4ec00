// Message function + sum aggregation of one residual_graph_conv_block on the 5th-generation tensor cores.
//
//   agg[t] = sum_{e: s->t}  ffn2( ffn1( cat(x_t, x_s, emb_e) ) )          (reference gnn_blocks.py:106-113)
//
// with ffn1 = Linear(2Cn+Ce -> H) + channel_normalization + LeakyReLU and ffn2 = Linear(H -> Cn) + norm + act.
// The node part of ffn1's Linear is hoisted out of the edge loop (P = [x W_t^T | x W_s^T], one row per node,
// produced by the node kernel), so per edge
//   z1 = emb_e W_e^T + b1 + P_t[tgt] + P_s[src]
// A CTA owns tiles of 128 target-major edges (UMMA M = 128, thread = TMEM lane = edge row):
//   GEMM1  D1[128 x H]  = A(emb tile, smem) * W_e^T (smem)          tcgen05.mma kind::tf32, SS
//   epi 1  z1 -> mean / unbiased std / scalar affine / LeakyReLU -> y1, written back to TMEM (hi and lo parts)
//   GEMM2  D2[128 x Cn] = y1 (TMEM) * W_2^T (smem)                  tcgen05.mma kind::tf32, TS
//   epi 2  + b2 -> norm -> act -> staged in smem -> segmented sum over equal consecutive targets -> agg
// Per-edge activations never leave the SM.  fp32 parity (rtol 1e-4 through seven residual layers) is kept by
// 3xTF32: every operand is split x = hi + lo (both exact in tf32) and D = A_lo B_hi + A_hi B_lo + A_hi B_hi
// is accumulated in fp32 by the tensor core (relative error ~2^-21 per product); `passes = 1` keeps only the
// hi*hi term (plain TF32, documented tolerance).
#include "rgnn_model.h"
#include "rgnn_tc.cuh"

namespace rgnn {

struct MpTcArgs {
    const float* emb;       // (E, CE) target-major
    const float* P;         // (N, 2H)
    const int* tgt;
    const int* src;
    const int* row_ptr;
    const float* wpack;     // [W1e_hi | W1e_lo | W2_hi | W2_lo], chunk-major (rgnn_tc.cuh)
    const float* b1;
    const float* s1;        // channel_normalization.std of msg.0 (scalar) or nullptr
    const float* m1;
    const float* b2;
    const float* s2;
    const float* m2;
    float* agg;             // (N, CN), zero-initialised by the caller
    int n_edges;
    int act1, act2;
    int passes;             // 3 = 3xTF32 (fp32 parity), 1 = plain TF32
};

static int g_tf32_passes = 3;
static int g_use_tensor_cores = 1;

template <int CE, int H, int CN>
struct MpTcLayout {
    static constexpr int TM = 128;
    static constexpr int W1 = CE * H;          // floats per (hi or lo) copy
    static constexpr int W2 = H * CN;
    static constexpr int A = TM * CE;
    static constexpr int LDS = CN + 4;         // staging row stride (floats): 16 B skew per row, conflict-free float4 rows
    static constexpr int OFF_W1 = 0;
    static constexpr int OFF_W2 = OFF_W1 + 2 * W1;
    static constexpr int OFF_A = OFF_W2 + 2 * W2;
    static constexpr int A_REGION = (2 * A > TM * LDS) ? 2 * A : TM * LDS;
    static constexpr int OFF_B1 = OFF_A + A_REGION;
    static constexpr int OFF_B2 = OFF_B1 + H;
    static constexpr int OFF_TGT = OFF_B2 + CN;
    static constexpr int OFF_BAR = OFF_TGT + TM;           // 2 x uint64
    static constexpr int OFF_SLOT = OFF_BAR + 4;
    static constexpr int FLOATS = OFF_SLOT + 4;
    static constexpr size_t BYTES = (size_t)FLOATS * 4;
    static constexpr int TMEM_COLS = 512;                   // D1/y1_hi [0,H) | y1_lo [H,2H) | D2 [2H, 2H+CN)
    static_assert(CE % 8 == 0 && H % 16 == 0 && CN % 16 == 0, "UMMA shape constraints");
    static_assert(2 * H + CN <= 512, "TMEM budget");
    static_assert(128 % CN == 0, "segmented sum thread mapping");
    static_assert((OFF_BAR % 2) == 0, "mbarrier alignment");
};

// channel_normalization + LeakyReLU on a register-resident row (reference common.py:215-220)
template <int C>
__device__ __forceinline__ void row_norm_act(float (&z)[C], const float* __restrict__ sp, const float* __restrict__ mp, bool act) {
    if (sp != nullptr) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) s += z[c];
        const float mean = s * (1.f / (float)C);
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            z[c] -= mean;
            ss = fmaf(z[c], z[c], ss);
        }
        const float sd = sqrtf(ss * (1.f / (float)(C - 1)));
        const float k = __ldg(sp) / (sd + NORM_EPS);
        const float sh = __ldg(mp);
#pragma unroll
        for (int c = 0; c < C; ++c) z[c] = fmaf(z[c], k, sh);
    }
    if (act) {
#pragma unroll
        for (int c = 0; c < C; ++c) z[c] = z[c] > 0.f ? z[c] : LEAKY * z[c];
    }
}

template <int CE, int H, int CN>
__global__ void __launch_bounds__(128, 1) mp_edge_tc_kernel(const __grid_constant__ MpTcArgs a) {
    using L = MpTcLayout<CE, H, CN>;
    constexpr int TM = L::TM;
    extern __shared__ __align__(1024) float smem[];
    float* w1s = smem + L::OFF_W1;
    float* w2s = smem + L::OFF_W2;
    float* As = smem + L::OFF_A;
    float* stage = As;                       // aliases the A operand (free once GEMM1 has completed)
    float* b1s = smem + L::OFF_B1;
    float* b2s = smem + L::OFF_B2;
    int* tgt_s = reinterpret_cast<int*>(smem + L::OFF_TGT);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem + L::OFF_SLOT);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- one-time setup: weights -> smem (already hi/lo split and chunk-major), barriers, TMEM ----
    {
        const float4* g = reinterpret_cast<const float4*>(a.wpack);
        float4* s = reinterpret_cast<float4*>(w1s);
        constexpr int N4 = (2 * L::W1 + 2 * L::W2) / 4;
        for (int i = tid; i < N4; i += 128) s[i] = __ldg(g + i);
        for (int i = tid; i < H; i += 128) b1s[i] = a.b1 ? __ldg(a.b1 + i) : 0.f;
        for (int i = tid; i < CN; i += 128) b2s[i] = a.b2 ? __ldg(a.b2 + i) : 0.f;
    }
    if (tid == 0) {
        tc::mbar_init(&bars[0], 1);
        tc::mbar_init(&bars[1], 1);
        tc::mbar_init_fence();
    }
    if (warp == 0) tc::tmem_alloc(slot, L::TMEM_COLS);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *slot;
    const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);   // this warp's 32 lanes
    constexpr uint32_t COL_Y1LO = H, COL_D2 = 2 * H;

    constexpr uint32_t IDESC1 = tc::idesc_tf32(TM, H);
    constexpr uint32_t IDESC2 = tc::idesc_tf32(TM, CN);
    constexpr uint32_t LBO_A = TM * 16, LBO_W1 = H * 16, LBO_W2 = CN * 16, SBO = 128;
    const uint32_t sA = tc::smem_u32(As), sW1 = tc::smem_u32(w1s), sW2 = tc::smem_u32(w2s);

    const int n_tiles = (a.n_edges + TM - 1) / TM;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, phase ^= 1) {
        const int row0 = tile * TM;
        const int nvalid = min(TM, a.n_edges - row0);

        // ---- (a) emb tile -> A operand, hi/lo split.  A quarter-warp reads 8 rows x 16 B and writes one
        //      128-byte core matrix; a warp request covers 8 rows x 64 B (fully used sectors).
        {
            const int r8 = lane & 7, kq = lane >> 3;
            float4* Ahi = reinterpret_cast<float4*>(As);
            float4* Alo = reinterpret_cast<float4*>(As + L::A);
#pragma unroll
            for (int rg = 0; rg < 4; ++rg) {
                const int r = warp * 32 + rg * 8 + r8;
                const bool ok = r < nvalid;
                const float4* grow = reinterpret_cast<const float4*>(a.emb + (size_t)(row0 + r) * CE);
#pragma unroll
                for (int kb = 0; kb < CE / 16; ++kb) {
                    const int kc = kb * 4 + kq;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ok) v = __ldg(grow + kc);
                    float4 hi, lo;
                    tc::split_tf32(v.x, hi.x, lo.x);
                    tc::split_tf32(v.y, hi.y, lo.y);
                    tc::split_tf32(v.z, hi.z, lo.z);
                    tc::split_tf32(v.w, hi.w, lo.w);
                    Ahi[kc * TM + r] = hi;
                    Alo[kc * TM + r] = lo;
                }
            }
            tgt_s[tid] = tid < nvalid ? __ldg(a.tgt + row0 + tid) : -1;
        }
        tc::fence_async_smem();
        tc::tc_fence_before();
        __syncthreads();

        // ---- (b) GEMM1: D1 = A * W1e^T ----
        if (tid == 0) {
            tc::tc_fence_after();
            bool acc = false;
            const int np = a.passes == 1 ? 1 : 3;
            for (int p = 0; p < np; ++p) {
                // small terms first: lo*hi, hi*lo, then hi*hi
                const int pa = (np == 1) ? 0 : (p == 0 ? 1 : 0);
                const int pb = (np == 1) ? 0 : (p == 1 ? 1 : 0);
                const uint32_t abase = sA + pa * (L::A * 4), bbase = sW1 + pb * (L::W1 * 4);
#pragma unroll
                for (int ks = 0; ks < CE / 8; ++ks) {
                    const uint64_t ad = tc::smem_desc(abase + ks * 2 * LBO_A, LBO_A, SBO);
                    const uint64_t bd = tc::smem_desc(bbase + ks * 2 * LBO_W1, LBO_W1, SBO);
                    tc::mma_tf32_ss(tmem, ad, bd, IDESC1, acc);
                    acc = true;
                }
            }
            tc::mma_commit(&bars[0]);
        }
        __syncwarp();

        // ---- (c) epilogue 1 ----
        const bool valid = tid < nvalid;
        const int e = row0 + tid;
        const int tg = valid ? tgt_s[tid] : 0;
        const int sr = valid ? __ldg(a.src + e) : 0;
        const float4* Pt = reinterpret_cast<const float4*>(a.P + (size_t)tg * (2 * H));
        const float4* Ps = reinterpret_cast<const float4*>(a.P + (size_t)sr * (2 * H) + H);
        tc::mbar_wait(&bars[0], phase);
        tc::tc_fence_after();
        {
            float z[H];
#pragma unroll
            for (int c = 0; c < H; c += 16) tc::tmem_ld16(t_row + c, z + c);
            tc::tmem_wait_ld();
            if (valid) {
#pragma unroll
                for (int c4 = 0; c4 < H / 4; ++c4) {
                    const float4 pt = __ldg(Pt + c4), ps = __ldg(Ps + c4);
                    const float4 b = *reinterpret_cast<const float4*>(b1s + 4 * c4);
                    z[4 * c4 + 0] += b.x + (pt.x + ps.x);
                    z[4 * c4 + 1] += b.y + (pt.y + ps.y);
                    z[4 * c4 + 2] += b.z + (pt.z + ps.z);
                    z[4 * c4 + 3] += b.w + (pt.w + ps.w);
                }
            }
            row_norm_act<H>(z, a.s1, a.m1, a.act1 != 0);
            // y1 -> TMEM as the A operand of GEMM2: hi part in place over D1, lo part next to it
            if (a.passes == 1) {
#pragma unroll
                for (int c = 0; c < H; ++c) z[c] = tc::tf32_rna(z[c]);
#pragma unroll
                for (int c = 0; c < H; c += 16) tc::tmem_st16(t_row + c, z + c);
            } else {
#pragma unroll
                for (int c = 0; c < H; c += 16) {
                    float hi[16], lo[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) tc::split_tf32(z[c + j], hi[j], lo[j]);
                    tc::tmem_st16(t_row + c, hi);
                    tc::tmem_st16(t_row + COL_Y1LO + c, lo);
                }
            }
            tc::tmem_wait_st();
        }
        tc::tc_fence_before();
        __syncthreads();

        // ---- (d) GEMM2: D2 = y1 * W2^T, A from TMEM ----
        if (tid == 0) {
            tc::tc_fence_after();
            bool acc = false;
            const int np = a.passes == 1 ? 1 : 3;
            for (int p = 0; p < np; ++p) {
                const int pa = (np == 1) ? 0 : (p == 0 ? 1 : 0);
                const int pb = (np == 1) ? 0 : (p == 1 ? 1 : 0);
                const uint32_t acol = tmem + (pa ? COL_Y1LO : 0);
                const uint32_t bbase = sW2 + pb * (L::W2 * 4);
#pragma unroll
                for (int ks = 0; ks < H / 8; ++ks) {
                    const uint64_t bd = tc::smem_desc(bbase + ks * 2 * LBO_W2, LBO_W2, SBO);
                    tc::mma_tf32_ts(tmem + COL_D2, acol + ks * 8, bd, IDESC2, acc);
                    acc = true;
                }
            }
            tc::mma_commit(&bars[1]);
        }
        __syncwarp();

        // ---- (e) epilogue 2 -> staging ----
        tc::mbar_wait(&bars[1], phase);
        tc::tc_fence_after();
        {
            float m[CN];
#pragma unroll
            for (int c = 0; c < CN; c += 16) tc::tmem_ld16(t_row + COL_D2 + c, m + c);
            tc::tmem_wait_ld();
#pragma unroll
            for (int c = 0; c < CN; ++c) m[c] += b2s[c];
            row_norm_act<CN>(m, a.s2, a.m2, a.act2 != 0);
            float4* srow = reinterpret_cast<float4*>(stage + tid * L::LDS);
#pragma unroll
            for (int c4 = 0; c4 < CN / 4; ++c4) srow[c4] = make_float4(m[4 * c4], m[4 * c4 + 1], m[4 * c4 + 2], m[4 * c4 + 3]);
        }
        tc::tc_fence_before();
        __syncthreads();

        // ---- (f) segmented sum over equal consecutive targets.  A target whose whole CSR row lies inside this
        //      thread's row range is stored (source-ascending order, like the reference's index_add_); a row cut
        //      by a range boundary is completed with atomicAdd onto the zero-initialised output.
        {
            constexpr int PARTS = 128 / CN, RP = TM / PARTS;
            const int j = tid % CN, part = tid / CN;
            int r = part * RP;
            const int rend = min(r + RP, nvalid);
            while (r < rend) {
                const int tn = tgt_s[r];
                float s = 0.f;
                int r1 = r;
                while (r1 < rend && tgt_s[r1] == tn) {
                    s += stage[r1 * L::LDS + j];
                    ++r1;
                }
                const bool whole = (__ldg(a.row_ptr + tn) == row0 + r) && (__ldg(a.row_ptr + tn + 1) == row0 + r1);
                float* o = a.agg + (size_t)tn * CN + j;
                if (whole) *o = s; else atomicAdd(o, s);
                r = r1;
            }
        }
        __syncthreads();
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, L::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// weight packing: W[n][koff + k] (row stride ldW) -> hi/lo chunk-major operands [K/4][N][4]
// ---------------------------------------------------------------------------------------------
__global__ void pack_split_kernel(const float* __restrict__ W, int ldW, int koff, int K, int N, float* __restrict__ hi,
                                  float* __restrict__ lo) {
    const int tot = K * N;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += gridDim.x * blockDim.x) {
        const int q = i & 3, n = (i >> 2) % N, kc = (i >> 2) / N;
        const float w = W[(size_t)n * ldW + koff + 4 * kc + q];
        float h, l;
        tc::split_tf32(w, h, l);
        hi[i] = h;
        lo[i] = l;
    }
}

bool mp_tc_supported(const ConvDims& d) { return g_use_tensor_cores && d.cn == 64 && d.ce == 64 && d.h == 128; }

size_t mp_tc_pack_floats(const ConvDims& d) {
    // the buffer exists whenever the shape is one the tensor-core kernel is instantiated for
    return (d.cn == 64 && d.ce == 64 && d.h == 128) ? (size_t)2 * d.ce * d.h + (size_t)2 * d.h * d.cn : 0;
}

int mp_tc_pack(const rgnn_conv& c, const ConvDims& d, float* dst, cudaStream_t stream) {
    if (mp_tc_pack_floats(d) == 0) return RGNN_OK;
    const rgnn_linear& m0 = c.msg.layer[0];
    const rgnn_linear& m1 = c.msg.layer[1];
    const int W1 = d.ce * d.h, W2 = d.h * d.cn;
    pack_split_kernel<<<16, 256, 0, stream>>>(m0.weight, m0.in_features, 2 * d.cn, d.ce, d.h, dst, dst + W1);
    pack_split_kernel<<<16, 256, 0, stream>>>(m1.weight, m1.in_features, 0, d.h, d.cn, dst + 2 * W1, dst + 2 * W1 + W2);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

int run_conv_edges_tc(const rgnn_conv& c, const ConvDims& d, const rgnn_graph& g, const float* emb, const float* P,
                      const float* wpack, float* agg, cudaStream_t stream) {
    using L = MpTcLayout<64, 128, 64>;
    const rgnn_linear& m0 = c.msg.layer[0];
    const rgnn_linear& m1 = c.msg.layer[1];
    MpTcArgs a;
    a.emb = emb; a.P = P; a.tgt = g.tgt; a.src = g.src; a.row_ptr = g.row_ptr;
    a.wpack = wpack;
    a.b1 = m0.bias; a.s1 = m0.norm_scale; a.m1 = m0.norm_shift;
    a.b2 = m1.bias; a.s2 = m1.norm_scale; a.m2 = m1.norm_shift;
    a.agg = agg; a.n_edges = g.n_edges;
    a.act1 = m0.activation; a.act2 = m1.activation;
    a.passes = g_tf32_passes;
    static bool configured = false;
    if (!configured) {
        RGNN_CHECK_CUDA(cudaFuncSetAttribute(mp_edge_tc_kernel<64, 128, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::BYTES));
        configured = true;
    }
    const int n_tiles = (g.n_edges + L::TM - 1) / L::TM;
    const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
    (void)d;
    mp_edge_tc_kernel<64, 128, 64><<<grid, 128, L::BYTES, stream>>>(a);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

}  // namespace rgnn

extern "C" int rgnn_set_option(const char* name, int value) {
    using namespace rgnn;
    if (name != nullptr && strcmp(name, "tf32_passes") == 0 && (value == 1 || value == 3)) { g_tf32_passes = value; return RGNN_OK; }
    if (name != nullptr && strcmp(name, "tensor_cores") == 0 && (value == 0 || value == 1)) { g_use_tensor_cores = value; return RGNN_OK; }
    set_error("rgnn_set_option: unknown option or value (%s = %d)", name ? name : "(null)", value);
    return RGNN_ERR_INVALID;
}

extern "C" int rgnn_get_option(const char* name) {
    using namespace rgnn;
    if (name != nullptr && strcmp(name, "tf32_passes") == 0) return g_tf32_passes;
    if (name != nullptr && strcmp(name, "tensor_cores") == 0) return g_use_tensor_cores;
    return -1;
}
