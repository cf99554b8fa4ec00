// Node-sized halves of the conv-block backward as two fixed-shape kernels on the tensor cores with fp16-split operands
// (rgnn_f16.cuh), each with its weight gradients fused in (the scheme of rgnn_mp_bwd_f16.cu: per-tile chunk-major activation
// images that are read K-major by the data-gradient GEMM and MN-major -- reduction over the tile's 128 NODES -- by the weight-
// gradient GEMMs, whose accumulators live in tensor memory for the whole kernel).  Replace, per conv block, three launches of the
// row-MLP interpreter and four weight-gradient GEMM launches (autograd of gnn_blocks.py:106-113 under reference training.py:81):
//
//   proj_bwd_f16_kernel    dx += dP_t W_t + dP_s W_s ;  dW_msg0[:, 0:cn] += dP_t^T x ;  dW_msg0[:, cn:2cn] += dP_s^T x
//                          (dP = gradient of the hoisted node projection P = [x W_t^T + b | x W_s^T] of msg.0, x = the block's input)
//   upd_bwd_f16_kernel     dz = norm'(act'(d out)) with the saved update output u and its sigma ;  d cat = dz W_u ;
//                          dx = d out + d cat[:, :cn] (identity residual), dagg = d cat[:, cn:] ;  dW_u += dz^T cat(x, agg), db_u += sum dz
//
// The FORWARD weight images serve the transposed products (MN-major reading).  One tile of 128 nodes in flight per CTA, 256 worker
// threads (row, half of the columns) + one MMA-issue lane; gradient operands carry one power-of-two scale per launch from max |.|
// of the incoming gradient (see rgnn_mp_bwd_f16.cu).
#include "rgnn_f16.cuh"
#include "rgnn_model.h"
#include "rgnn_tc_rows.cuh"
#include "rgnn_tile.cuh"

namespace rgnn {

__global__ void absmax_kernel(const float* __restrict__ x, size_t n, unsigned* __restrict__ out);      // rgnn_mp_bwd_f16.cu
const float* f16_weights(const rgnn_linear& L);                                                         // rgnn_model_tc.cu
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane);                              // defined below

namespace nbf {
constexpr int TM = 128, CN = 64, H = 128;
constexpr int NTHREADS = 384, NW = 256;
__host__ __device__ constexpr uint32_t idesc_mn(int M, int N, int a_mn, int b_mn) {
    return f16::idesc(M, N) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}
__device__ __forceinline__ void scale_from(const float* gmax, float& S, float& inv_S) {
    S = 1.f; inv_S = 1.f;
    const float gm = __ldg(gmax);
    if (gm > 0.f && gm < 3.0e38f) {
        int e = (int)((__float_as_uint(gm) >> 23) & 0xFFu) - 126;
        int k = 8 - e;
        k = k < -60 ? -60 : (k > 100 ? 100 : k);
        S = __uint_as_float((uint32_t)(127 + k) << 23);
        inv_S = __uint_as_float((uint32_t)(127 - k) << 23);
    }
}
}  // namespace nbf

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

// =============================================================================================
// projection backward
// =============================================================================================
struct ProjBwdArgs {
    const float* dP;        // (N, 2H)
    const float* x;         // (N, CN)  input of the block whose msg.0 this is
    float* dx;              // (N, CN)  += dP_t W_t + dP_s W_s
    const uint32_t* w_t;    // forward images of the two projection halves: [hi | lo] of (N = H, K = CN), x 256
    const uint32_t* w_s;
    float* gW;              // msg.0 grad weight (H, ldW): columns [0, CN) += dP_t^T x, [CN, 2 CN) += dP_s^T x   (nullable)
    int ldW;
    const float* gmax;      // max |dP|
    int n_rows;
    int passes;
};

namespace pbf {
using namespace nbf;
constexpr int W_WORDS = CN * H;                 // hi + lo of one half (32 KB)
constexpr int P_WORDS = TM * H / 2;             // one (hi or lo) image of a dP half: 32 KB
constexpr int X_WORDS = TM * CN / 2;            // one image of x: 16 KB
constexpr int OFF_WT = 0, OFF_WS = OFF_WT + W_WORDS;
constexpr int OFF_P = OFF_WS + W_WORDS;
constexpr int OFF_X = OFF_P + 2 * P_WORDS;
constexpr int OFF_BAR = OFF_X + 2 * X_WORDS;    // 5 mbarriers
constexpr int OFF_SLOT = OFF_BAR + 2 * 6;
constexpr int WORDS = OFF_SLOT + 2;
constexpr size_t SMEM = (size_t)WORDS * 4;
static_assert(SMEM <= 227 * 1024 && (OFF_BAR % 2) == 0 && (OFF_P % 4) == 0 && (OFF_X % 4) == 0, "shared memory");
constexpr uint32_t COL_D = 0, COL_ACC_T = 64, COL_ACC_S = 128;
enum { B_A0 = 0, B_A1, B_M0, B_M1, B_D };
}  // namespace pbf

__global__ void __launch_bounds__(nbf::NTHREADS, 1) proj_bwd_f16_kernel(const __grid_constant__ ProjBwdArgs a) {
    using namespace pbf;
    extern __shared__ __align__(1024) uint32_t smem_u[];
    uint4* pimg = reinterpret_cast<uint4*>(smem_u + OFF_P);         // [hi | lo][H/8][TM]
    uint4* ximg = reinterpret_cast<uint4*>(smem_u + OFF_X);         // [hi | lo][CN/8][TM]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_u + OFF_BAR);
    uint32_t* slot = smem_u + OFF_SLOT;
    constexpr int PI = P_WORDS / 4, XI = X_WORDS / 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, row = tid & 127;
    const int G = (int)gridDim.x;
    const int n_tiles = (a.n_rows + TM - 1) / TM;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / G + 1 : 0;
    const int np = a.passes == 1 ? 1 : 3;
    {
        uint4* s = reinterpret_cast<uint4*>(smem_u + OFF_WT);
        const uint4* g0 = reinterpret_cast<const uint4*>(a.w_t);
        const uint4* g1 = reinterpret_cast<const uint4*>(a.w_s);
        for (int i = tid; i < W_WORDS / 4; i += NTHREADS) { s[i] = __ldg(g0 + i); s[W_WORDS / 4 + i] = __ldg(g1 + i); }
    }
    if (tid == 0) {
        tc::mbar_init(&bars[B_A0], 8);
        tc::mbar_init(&bars[B_A1], 8);
        tc::mbar_init(&bars[B_M0], 1);
        tc::mbar_init(&bars[B_M1], 1);
        tc::mbar_init(&bars[B_D], 1);
        tc::mbar_init_fence();
    }
    if (warp == 0) tc::tmem_alloc(slot, 256);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *slot;
    const uint32_t t_row = tmem + ((uint32_t)(warp & 3) << 21);
    float S, inv_S;
    scale_from(a.gmax, S, inv_S);

    if (tid < NW) {
        const int q = tid >> 7;
        const float2 S2 = make_float2(S, S), A2 = make_float2(f16::A_SCALE, f16::A_SCALE);
        for (int j = 0; j < my_tiles; ++j) {
            const uint32_t ph = (uint32_t)j & 1u;
            const long long e = ((long long)blockIdx.x + (long long)j * G) * TM + row;
            const bool valid = e < a.n_rows;
            // ---- loads: the own 64 columns of the target half of dP and 32 columns of x now; the source half while the first MMAs
            // run; the own 32 columns of dx just before the accumulator is read (register pressure) ----
            const float* pp = a.dP + (size_t)(valid ? e : 0) * (2 * H) + 64 * q;
            float2 p0[32], xv[16];
            {
                const float* px = a.x + (size_t)(valid ? e : 0) * CN + 32 * q;
#pragma unroll
                for (int i = 0; i < 8; ++i) ldg256(pp + 8 * i, p0[4 * i], p0[4 * i + 1], p0[4 * i + 2], p0[4 * i + 3]);
#pragma unroll
                for (int i = 0; i < 4; ++i) ldg256(px + 8 * i, xv[4 * i], xv[4 * i + 1], xv[4 * i + 2], xv[4 * i + 3]);
            }
            if (j > 0) tc::mbar_wait(&bars[B_M1], ph ^ 1u);      // the previous tile's MMAs have read both images
            auto put_p = [&](float2 (&p)[32]) {      // own 64 columns -> chunks 8 q .. 8 q + 7 of the dP image (x S, hi | lo)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) f16::split(valid ? __fmul2_rn(p[4 * k + i], S2) : make_float2(0.f, 0.f), hi[i], lo[i]);
                    pimg[(8 * q + k) * TM + row] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    pimg[PI + (8 * q + k) * TM + row] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
            };
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) f16::split(valid ? __fmul2_rn(xv[4 * k + i], A2) : make_float2(0.f, 0.f), hi[i], lo[i]);
                ximg[(4 * q + k) * TM + row] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                ximg[XI + (4 * q + k) * TM + row] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            put_p(p0);
            tc::fence_async_smem();
            warp_arrive(&bars[B_A0], lane);
#pragma unroll
            for (int i = 0; i < 8; ++i) ldg256(pp + H + 8 * i, p0[4 * i], p0[4 * i + 1], p0[4 * i + 2], p0[4 * i + 3]);
            tc::mbar_wait(&bars[B_M0], ph);          // the target half's MMAs have read the dP image
            put_p(p0);
            tc::fence_async_smem();
            warp_arrive(&bars[B_A1], lane);
            float2 dxv[16];
            {
                const float* pd = a.dx + (size_t)(valid ? e : 0) * CN + 32 * q;
#pragma unroll
                for (int i = 0; i < 4; ++i) ldg256(pd + 8 * i, dxv[4 * i], dxv[4 * i + 1], dxv[4 * i + 2], dxv[4 * i + 3]);
            }
            // ---- dx += D / (256 S) ----
            tc::mbar_wait(&bars[B_D], ph);
            tc::tc_fence_after();
            {
                float2 d[16];
                tc::tmem_ld16(t_row + COL_D + 32 * q, d);
                tc::tmem_ld16(t_row + COL_D + 32 * q + 16, d + 8);
                tc::tmem_wait_ld();
                const float u = inv_S * (1.f / f16::W_SCALE);
                const float2 u2 = make_float2(u, u);
                if (valid) {
                    float* o = a.dx + (size_t)e * CN + 32 * q;
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8)
                        stg256(o + 8 * c8, __ffma2_rn(d[4 * c8], u2, dxv[4 * c8]), __ffma2_rn(d[4 * c8 + 1], u2, dxv[4 * c8 + 1]),
                               __ffma2_rn(d[4 * c8 + 2], u2, dxv[4 * c8 + 2]), __ffma2_rn(d[4 * c8 + 3], u2, dxv[4 * c8 + 3]));
                }
            }
            tc::tc_fence_before();
        }
        // ---- flush the two weight-gradient accumulators: lane = projection column h', column = node channel c ----
        if (my_tiles > 0 && a.gW != nullptr) {
            tc::mbar_wait(&bars[B_M1], (uint32_t)(my_tiles - 1) & 1u);
            tc::tc_fence_after();
            const float u = inv_S * (1.f / f16::A_SCALE);
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
#pragma unroll 1
                for (int cc = 0; cc < 2; ++cc) {
                    float v[16];
                    tc::tmem_ld16(t_row + (half ? COL_ACC_S : COL_ACC_T) + 32 * q + 16 * cc, v);
                    tc::tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; ++i) atomicAdd(a.gW + (size_t)row * a.ldW + half * CN + 32 * q + 16 * cc + i, v[i] * u);
                }
            }
        }
    } else if (warp == NW / 32 && lane == 0) {
        constexpr uint32_t ID_G = idesc_mn(TM, CN, 0, 1), ID_W = idesc_mn(H, CN, 1, 1);
        const uint32_t sP = tc::smem_u32(pimg), sX = tc::smem_u32(ximg);
        const uint32_t sW[2] = {tc::smem_u32(smem_u + OFF_WT), tc::smem_u32(smem_u + OFF_WS)};
        bool wacc = false;
        for (int j = 0; j < my_tiles; ++j) {
            const uint32_t ph = (uint32_t)j & 1u;
            for (int half = 0; half < 2; ++half) {
                tc::mbar_wait(&bars[half ? B_A1 : B_A0], ph);
                tc::tc_fence_after();
                // D (+)= dP_half W_half: A = the dP image K-major (rows = nodes), B = the forward image (N = H rows, K = CN) read MN-major
                for (int p = 0; p < np; ++p) {
                    const int pa = (np == 1) ? 0 : (p == 0 ? 1 : 0), pb = (np == 1) ? 0 : (p == 1 ? 1 : 0);
                    const uint64_t ad = tc::smem_desc(sP + pa * (P_WORDS * 4), TM * 16, 128);
                    const uint64_t bd = tc::smem_desc(sW[half] + pb * (W_WORDS * 2), 128, H * 16);
#pragma unroll
                    for (int ks = 0; ks < H / 16; ++ks)
                        f16::mma_ss(tmem + COL_D, ad + (uint64_t)((ks * 2 * TM * 16) >> 4), bd + (uint64_t)((ks * 256) >> 4), ID_G, half > 0 || p > 0 || ks > 0);
                }
                if (half == 1) tc::mma_commit(&bars[B_D]);
                // dW_half += dP_half^T x: both images MN-major, K = the tile's 128 nodes
                for (int p = 0; p < np; ++p) {
                    const int pa = (np == 1) ? 0 : (p == 0 ? 1 : 0), pb = (np == 1) ? 0 : (p == 1 ? 1 : 0);
                    const uint64_t ad = tc::smem_desc(sP + pa * (P_WORDS * 4), 128, TM * 16);
                    const uint64_t bd = tc::smem_desc(sX + pb * (X_WORDS * 4), 128, TM * 16);
#pragma unroll
                    for (int ks = 0; ks < TM / 16; ++ks)
                        f16::mma_ss(tmem + (half ? COL_ACC_S : COL_ACC_T), ad + (uint64_t)((ks * 256) >> 4), bd + (uint64_t)((ks * 256) >> 4), ID_W,
                                    wacc || p > 0 || ks > 0);
                }
                tc::mma_commit(&bars[half ? B_M1 : B_M0]);
            }
            wacc = true;
        }
    }
    __syncwarp();
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// =============================================================================================
// node-update backward
// =============================================================================================
struct UpdBwdArgs {
    const float* x;         // (N, CN)  block input
    const float* agg;       // (N, CN)  aggregated messages
    const float* u;         // (N, CN)  update output before the residual (post activation), saved by the forward
    const float* sd;        // (N)      its sigma
    float* dx;              // (N, CN)  d out on entry, d x on exit
    float* dagg;            // (N, CN)
    const uint32_t* w_u;    // forward image of upd.0: [hi | lo] of (N = CN rows, K = 2 CN), x 256
    const float* scale; const float* shift;     // channel_normalization scalars (nullptr = no norm)
    float* gW;              // (CN, 2 CN) += dz^T cat(x, agg)   (nullable)
    float* gb;              // (CN)       += sum dz
    float* g_scale; float* g_shift;
    const float* gmax;      // max |d out|
    int n_rows, act, passes;
};

namespace ubf {
using namespace nbf;
constexpr int K2 = 2 * CN;                      // 128: width of cat(x, agg)
constexpr int WU_WORDS = K2 * CN;               // hi + lo (32 KB)
constexpr int C_WORDS = TM * K2 / 2;            // one image of cat: 32 KB
constexpr int Z_WORDS = TM * CN / 2;            // one image of dz: 16 KB
constexpr int OFF_WU = 0;
constexpr int OFF_C = OFF_WU + WU_WORDS;
constexpr int OFF_Z = OFF_C + 2 * C_WORDS;
constexpr int OFF_XCH = OFF_Z + 2 * Z_WORDS;    // [TM][2] float2
constexpr int OFF_RED = OFF_XCH + TM * 2 * 2;   // 2 x 8 doubles
constexpr int OFF_BAR = OFF_RED + 32;
constexpr int OFF_SLOT = OFF_BAR + 2 * 4;
constexpr int WORDS = OFF_SLOT + 2;
constexpr size_t SMEM = (size_t)WORDS * 4;
static_assert(SMEM <= 227 * 1024 && (OFF_BAR % 2) == 0 && (OFF_RED % 2) == 0 && (OFF_XCH % 2) == 0, "shared memory");
constexpr uint32_t COL_D = 0, COL_ACC = 128;
enum { B_A = 0, B_D, B_W };
}  // namespace ubf

__global__ void __launch_bounds__(nbf::NTHREADS, 1) upd_bwd_f16_kernel(const __grid_constant__ UpdBwdArgs a) {
    using namespace ubf;
    extern __shared__ __align__(1024) uint32_t smem_u[];
    uint4* cimg = reinterpret_cast<uint4*>(smem_u + OFF_C);         // [hi | lo][K2/8][TM]
    uint4* zimg = reinterpret_cast<uint4*>(smem_u + OFF_Z);         // [hi | lo][CN/8][TM]
    float2* xch = reinterpret_cast<float2*>(smem_u + OFF_XCH);
    double* red = reinterpret_cast<double*>(smem_u + OFF_RED);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_u + OFF_BAR);
    uint32_t* slot = smem_u + OFF_SLOT;
    constexpr int CI = C_WORDS / 4, ZI = Z_WORDS / 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, row = tid & 127;
    const int G = (int)gridDim.x;
    const int n_tiles = (a.n_rows + TM - 1) / TM;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / G + 1 : 0;
    const int np = a.passes == 1 ? 1 : 3;
    {
        uint4* s = reinterpret_cast<uint4*>(smem_u + OFF_WU);
        const uint4* g0 = reinterpret_cast<const uint4*>(a.w_u);
        for (int i = tid; i < WU_WORDS / 4; i += NTHREADS) s[i] = __ldg(g0 + i);
    }
    if (tid == 0) {
        tc::mbar_init(&bars[B_A], 8);
        tc::mbar_init(&bars[B_D], 1);
        tc::mbar_init(&bars[B_W], 1);
        tc::mbar_init_fence();
    }
    if (warp == 0) tc::tmem_alloc(slot, 256);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *slot;
    const uint32_t t_row = tmem + ((uint32_t)(warp & 3) << 21);
    float S, inv_S;
    scale_from(a.gmax, S, inv_S);

    if (tid < NW) {
        const int q = tid >> 7;
        const int bar_id = 1 + (row >> 5);
        const bool norm = a.scale != nullptr, act = a.act != 0;
        const float sv = norm ? __ldg(a.scale) : 1.f, mv = norm ? __ldg(a.shift) : 0.f;
        const float inv_s = sv != 0.f ? 1.f / sv : 0.f;
        const float2 is2 = make_float2(inv_s, inv_s), nsh2 = make_float2(-mv * inv_s, -mv * inv_s), s22 = make_float2(sv, sv);
        const float2 S2 = make_float2(S, S), A2 = make_float2(f16::A_SCALE, f16::A_SCALE);
        double acc_s = 0., acc_m = 0.;
        float acc_b = 0.f;
        for (int j = 0; j < my_tiles; ++j) {
            const uint32_t ph = (uint32_t)j & 1u;
            const long long e = ((long long)blockIdx.x + (long long)j * G) * TM + row;
            const bool valid = e < a.n_rows;
            const size_t er = (size_t)(valid ? e : 0);
            // ---- loads: own 32 columns of d out and u; 64 columns of x (q = 0) or agg (q = 1) ----
            float2 g[16], uu[16];
            {
                float2 cv[32];
                const float* pg = a.dx + er * CN + 32 * q;
                const float* pu = a.u + er * CN + 32 * q;
                const float* pc = (q == 0 ? a.x : a.agg) + er * CN;
#pragma unroll
                for (int i = 0; i < 8; ++i) ldg256(pc + 8 * i, cv[4 * i], cv[4 * i + 1], cv[4 * i + 2], cv[4 * i + 3]);
#pragma unroll
                for (int i = 0; i < 4; ++i) ldg256(pg + 8 * i, g[4 * i], g[4 * i + 1], g[4 * i + 2], g[4 * i + 3]);
#pragma unroll
                for (int i = 0; i < 4; ++i) ldg256(pu + 8 * i, uu[4 * i], uu[4 * i + 1], uu[4 * i + 2], uu[4 * i + 3]);
                if (j > 0) tc::mbar_wait(&bars[B_W], ph ^ 1u);       // the previous tile's MMAs have read both images
#pragma unroll
                for (int k = 0; k < 8; ++k) {      // cat(x, agg) x 16 -> chunks 8 q .. 8 q + 7 of the cat image
                    uint32_t h4[4], l4[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) f16::split(valid ? __fmul2_rn(cv[4 * k + i], A2) : make_float2(0.f, 0.f), h4[i], l4[i]);
                    cimg[(8 * q + k) * TM + row] = make_uint4(h4[0], h4[1], h4[2], h4[3]);
                    cimg[CI + (8 * q + k) * TM + row] = make_uint4(l4[0], l4[1], l4[2], l4[3]);
                }
            }
            const float sdv = (norm && valid) ? __ldg(a.sd + er) : 0.f;
#pragma unroll
            for (int c = 0; c < 16; ++c)
                if (!valid) g[c] = make_float2(0.f, 0.f);
            // ---- dz = norm'(act'(d out)) ----
            float2 nv[16];
            float2 ps2 = make_float2(0.f, 0.f), pm2 = ps2, sum2 = ps2, dot2 = ps2;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                act_bwd_pair(g[c], nv[c], uu[c], act, is2, nsh2);
                if (norm) {
                    ps2 = __ffma2_rn(g[c], nv[c], ps2);
                    pm2 = __fadd2_rn(pm2, g[c]);
                    g[c] = __fmul2_rn(g[c], s22);
                    sum2 = __fadd2_rn(sum2, g[c]);
                    dot2 = __ffma2_rn(g[c], nv[c], dot2);
                }
            }
            if (norm) {
                if (valid) { acc_s += (double)(ps2.x + ps2.y); acc_m += (double)(pm2.x + pm2.y); }
                xch[row * 2 + q] = make_float2(sum2.x + sum2.y, dot2.x + dot2.y);
                group_sync(bar_id, 64);
                const float2 o = xch[row * 2 + (q ^ 1)];
                const float sum_dn = (sum2.x + sum2.y) + o.x, dot = (dot2.x + dot2.y) + o.y;
                const float inv_den = 1.f / (sdv + NORM_EPS);
                const float mean_dn = sum_dn / (float)CN;
                const float coef = sdv > 0.f ? dot / ((float)(CN - 1) * sdv) : 0.f;
                const float2 nm2 = make_float2(-mean_dn, -mean_dn), id2 = make_float2(inv_den, inv_den), nc2 = make_float2(-coef, -coef);
#pragma unroll
                for (int c = 0; c < 16; ++c) g[c] = valid ? __ffma2_rn(nv[c], nc2, __fmul2_rn(__fadd2_rn(g[c], nm2), id2)) : make_float2(0.f, 0.f);
            }
            {
                uint32_t hi[16], lo[16];
                float cs[32];
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const float2 v = __fmul2_rn(g[c], S2);
                    cs[2 * c] = v.x; cs[2 * c + 1] = v.y;
                    f16::split(v, hi[c], lo[c]);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    zimg[(4 * q + k) * TM + row] = make_uint4(hi[4 * k], hi[4 * k + 1], hi[4 * k + 2], hi[4 * k + 3]);
                    zimg[ZI + (4 * q + k) * TM + row] = make_uint4(lo[4 * k], lo[4 * k + 1], lo[4 * k + 2], lo[4 * k + 3]);
                }
                tc::fence_async_smem();
                warp_arrive(&bars[B_A], lane);
                acc_b += warp_colsum32(cs, lane);           // bias gradient: column 32 q + lane over this warp's rows
            }
            // the residual d out (whole row, q = 0 only): requested now, while the MMAs run (the row is rewritten by this thread alone)
            float2 gres[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                gres[4 * i] = gres[4 * i + 1] = gres[4 * i + 2] = gres[4 * i + 3] = make_float2(0.f, 0.f);
                if (q == 0 && valid) ldg256(a.dx + er * CN + 8 * i, gres[4 * i], gres[4 * i + 1], gres[4 * i + 2], gres[4 * i + 3]);
            }
            // ---- d cat = D / (256 S): q = 0 -> dx = d out + d cat[:, :CN];  q = 1 -> dagg = d cat[:, CN:] ----
            tc::mbar_wait(&bars[B_D], ph);
            tc::tc_fence_after();
            {
                const float uq = inv_S * (1.f / f16::W_SCALE);
                const float2 u2 = make_float2(uq, uq);
                float* o = (q == 0 ? a.dx : a.dagg) + er * CN;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    float2 d[16];
                    tc::tmem_ld16(t_row + COL_D + 64 * q + 32 * cc, d);
                    tc::tmem_ld16(t_row + COL_D + 64 * q + 32 * cc + 16, d + 8);
                    tc::tmem_wait_ld();
                    if (valid) {
#pragma unroll
                        for (int c8 = 0; c8 < 4; ++c8) {
                            const float2 r0 = gres[16 * cc + 4 * c8], r1 = gres[16 * cc + 4 * c8 + 1], r2 = gres[16 * cc + 4 * c8 + 2],
                                         r3 = gres[16 * cc + 4 * c8 + 3];
                            stg256(o + 32 * cc + 8 * c8, __ffma2_rn(d[4 * c8], u2, r0), __ffma2_rn(d[4 * c8 + 1], u2, r1),
                                   __ffma2_rn(d[4 * c8 + 2], u2, r2), __ffma2_rn(d[4 * c8 + 3], u2, r3));
                        }
                    }
                }
            }
            tc::tc_fence_before();
        }
        // ---- flush: dW_u (lane = cat channel m, column = output channel n -> gW[n * K2 + m]), db_u, the norm scalars ----
        if (my_tiles > 0) {
            tc::mbar_wait(&bars[B_W], (uint32_t)(my_tiles - 1) & 1u);
            tc::tc_fence_after();
            const float u = inv_S * (1.f / f16::A_SCALE);
            if (a.gW != nullptr) {
#pragma unroll 1
                for (int cc = 0; cc < 2; ++cc) {
                    float v[16];
                    tc::tmem_ld16(t_row + COL_ACC + 32 * q + 16 * cc, v);
                    tc::tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; ++i) atomicAdd(a.gW + (size_t)(32 * q + 16 * cc + i) * K2 + row, v[i] * u);
                }
            }
            if (a.gb != nullptr) atomicAdd(a.gb + 32 * q + lane, acc_b * inv_S);
        }
        {
            double v[2] = {acc_s, acc_m};
#pragma unroll
            for (int i = 0; i < 2; ++i) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
                if (lane == 0) red[i * 8 + warp] = v[i];
            }
            group_sync(BAR_WORKERS, NW);
            if (tid < 2) {
                double s = 0.;
                for (int w = 0; w < NW / 32; ++w) s += red[tid * 8 + w];
                float* dst = tid == 0 ? a.g_scale : a.g_shift;
                if (dst != nullptr && my_tiles > 0) atomicAdd(dst, (float)s);
            }
        }
    } else if (warp == NW / 32 && lane == 0) {
        constexpr uint32_t ID_G = idesc_mn(TM, K2, 0, 1), ID_W = idesc_mn(K2, CN, 1, 1);
        const uint32_t sC = tc::smem_u32(cimg), sZ = tc::smem_u32(zimg), sW = tc::smem_u32(smem_u + OFF_WU);
        bool wacc = false;
        for (int j = 0; j < my_tiles; ++j) {
            const uint32_t ph = (uint32_t)j & 1u;
            tc::mbar_wait(&bars[B_A], ph);
            tc::tc_fence_after();
            // D = dz W_u: A = the dz image K-major (K = CN), B = the forward image (N = CN rows, K = K2) read MN-major (N = K2, K = CN)
            {
                bool acc = false;
                for (int p = 0; p < np; ++p) {
                    const int pa = (np == 1) ? 0 : (p == 0 ? 1 : 0), pb = (np == 1) ? 0 : (p == 1 ? 1 : 0);
                    const uint64_t ad = tc::smem_desc(sZ + pa * (Z_WORDS * 4), TM * 16, 128);
                    const uint64_t bd = tc::smem_desc(sW + pb * (WU_WORDS * 2), 128, CN * 16);
#pragma unroll
                    for (int ks = 0; ks < CN / 16; ++ks) {
                        f16::mma_ss(tmem + COL_D, ad + (uint64_t)((ks * 2 * TM * 16) >> 4), bd + (uint64_t)((ks * 256) >> 4), ID_G, acc);
                        acc = true;
                    }
                }
            }
            tc::mma_commit(&bars[B_D]);
            // dW_u += cat^T dz: A = cat image MN-major (M = K2 channels), B = dz image MN-major (N = CN), K = the tile's 128 nodes
            for (int p = 0; p < np; ++p) {
                const int pa = (np == 1) ? 0 : (p == 0 ? 1 : 0), pb = (np == 1) ? 0 : (p == 1 ? 1 : 0);
                const uint64_t ad = tc::smem_desc(sC + pa * (C_WORDS * 4), 128, TM * 16);
                const uint64_t bd = tc::smem_desc(sZ + pb * (Z_WORDS * 4), 128, TM * 16);
#pragma unroll
                for (int ks = 0; ks < TM / 16; ++ks)
                    f16::mma_ss(tmem + COL_ACC, ad + (uint64_t)((ks * 256) >> 4), bd + (uint64_t)((ks * 256) >> 4), ID_W, wacc || p > 0 || ks > 0);
            }
            tc::mma_commit(&bars[B_W]);
            wacc = true;
        }
    }
    __syncwarp();
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// =============================================================================================
// backward of a 64-wide ffn chain (stems, FFN_TaskSpecificHeads, the link head): up to 4 x (Linear 64 -> 64, norm, act) + an
// optional bare Linear 64 -> n_out <= 16, with the layer outputs and sigmas the forward of a training step saved.
// Per tile of 128 rows the layers are walked top down; per layer: dz = norm'(act'(g)) -> dz image, the layer's input rows -> input
// image; D = dz W (forward image read MN-major) -> g of the layer below; dW += input^T dz (both images MN-major, the input image
// zero padded to the 128 rows of the UMMA M), accumulators in tensor memory for the whole kernel, one per layer.
// =============================================================================================
constexpr int C64B_MAX = 4;

struct Chain64BwdArgs {
    int n_rows, n_hidden, tail, n_out;
    const float* x; int ldx;                // input rows of layer 0
    const float* y[C64B_MAX];               // outputs of the hidden layers (post activation), (n, 64)
    const float* sd[C64B_MAX];              // their sigmas
    const float* g_top; int ld_g;           // gradient w.r.t. the stack's output: (n, n_out) with a tail, else (n, 64)
    const uint32_t* w[C64B_MAX + 1];        // forward images [hi | lo]: hidden (K = 64, N = 64), tail (K = 64, N = 16)
    const float* scale[C64B_MAX]; const float* shift[C64B_MAX];
    int act[C64B_MAX];
    float* gW[C64B_MAX + 1]; float* gb[C64B_MAX + 1];
    float* g_scale[C64B_MAX]; float* g_shift[C64B_MAX];
    float* dx; int dx_mode;                 // 0 overwrite, 1 accumulate, 2 atomic scatter onto rows ia[], ib[] (pair sums); nullptr: not wanted
    const int* ia; const int* ib;
    const float* gmax;                      // max |g_top|
    int passes;
};

namespace cbf {
using namespace nbf;
constexpr int W = 64, NTAIL = 16;
constexpr int IMG_WORDS = W * W;            // hi + lo of a 64 x 64 layer (16 KB)
constexpr int TAIL_WORDS = W * NTAIL;          // hi + lo of the tail (4 KB)
constexpr int Y_WORDS = TM * 128 / 2;       // one image of the layer input, padded to 128 channels: 32 KB
constexpr int Z_WORDS = TM * W / 2;         // one image of dz: 16 KB
constexpr int OFF_W = 0;
constexpr int OFF_Y = OFF_W + C64B_MAX * IMG_WORDS + TAIL_WORDS;
constexpr int OFF_Z = OFF_Y + 2 * Y_WORDS;
constexpr int OFF_XCH = OFF_Z + 2 * Z_WORDS;            // [2 phases][TM][2] float2
constexpr int OFF_RED = OFF_XCH + 2 * TM * 2 * 2;    // 8 x 8 doubles
constexpr int OFF_BAR = OFF_RED + 128;
constexpr int OFF_SLOT = OFF_BAR + 2 * 4;
constexpr int WORDS = OFF_SLOT + 2;
constexpr size_t SMEM = (size_t)WORDS * 4;
static_assert(SMEM <= 227 * 1024 && (OFF_BAR % 2) == 0 && (OFF_RED % 2) == 0 && (OFF_XCH % 2) == 0 && (OFF_Y % 4) == 0 && (OFF_Z % 4) == 0,
              "shared memory");
constexpr uint32_t COL_D = 0, COL_ACC = 64, COL_ACC_TAIL = COL_ACC + 64 * C64B_MAX;
enum { B_A = 0, B_D, B_W };
}  // namespace cbf

__global__ void __launch_bounds__(nbf::NTHREADS, 1) chain64_bwd_f16_kernel(const __grid_constant__ Chain64BwdArgs a) {
    using namespace cbf;
    extern __shared__ __align__(1024) uint32_t smem_u[];
    uint4* yimg = reinterpret_cast<uint4*>(smem_u + OFF_Y);         // [hi | lo][16 chunks][TM], chunks 8..15 stay zero
    uint4* zimg = reinterpret_cast<uint4*>(smem_u + OFF_Z);         // [hi | lo][8 chunks][TM]
    float2* xch = reinterpret_cast<float2*>(smem_u + OFF_XCH);
    double* red = reinterpret_cast<double*>(smem_u + OFF_RED);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_u + OFF_BAR);
    uint32_t* slot = smem_u + OFF_SLOT;
    constexpr int YI = Y_WORDS / 4, ZI = Z_WORDS / 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, row = tid & 127;
    const int G = (int)gridDim.x;
    const int n_tiles = (a.n_rows + TM - 1) / TM;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / G + 1 : 0;
    const int np = a.passes == 1 ? 1 : 3;
    const int n_phase = a.n_hidden + (a.tail ? 1 : 0);
    for (int l = 0; l < a.n_hidden; ++l) {
        uint4* s = reinterpret_cast<uint4*>(smem_u + OFF_W + l * IMG_WORDS);
        const uint4* g0 = reinterpret_cast<const uint4*>(a.w[l]);
        for (int i = tid; i < IMG_WORDS / 4; i += NTHREADS) s[i] = __ldg(g0 + i);
    }
    if (a.tail) {
        uint4* s = reinterpret_cast<uint4*>(smem_u + OFF_W + C64B_MAX * IMG_WORDS);
        const uint4* g0 = reinterpret_cast<const uint4*>(a.w[a.n_hidden]);
        for (int i = tid; i < TAIL_WORDS / 4; i += NTHREADS) s[i] = __ldg(g0 + i);
    }
    for (int i = tid; i < 2 * YI; i += NTHREADS) yimg[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < 2 * ZI; i += NTHREADS) zimg[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) {
        tc::mbar_init(&bars[B_A], 8);
        tc::mbar_init(&bars[B_D], 1);
        tc::mbar_init(&bars[B_W], 1);
        tc::mbar_init_fence();
    }
    if (warp == 0) tc::tmem_alloc(slot, 512);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *slot;
    const uint32_t t_row = tmem + ((uint32_t)(warp & 3) << 21);
    float S, inv_S;
    scale_from(a.gmax, S, inv_S);

    if (tid < NW) {
        const int q = tid >> 7;
        const int bar_id = 1 + (row >> 5);
        const float2 S2 = make_float2(S, S), A2 = make_float2(f16::A_SCALE, f16::A_SCALE);
        const float uq = inv_S * (1.f / f16::W_SCALE);
        const float2 u2 = make_float2(uq, uq);
        double acc_s[C64B_MAX], acc_m[C64B_MAX];
        float acc_b[C64B_MAX + 1];
#pragma unroll
        for (int l = 0; l < C64B_MAX; ++l) { acc_s[l] = 0.; acc_m[l] = 0.; acc_b[l] = 0.f; }
        acc_b[C64B_MAX] = 0.f;
        uint32_t use = 0;           // MMA phases so far: the parity of the three barriers
        for (int j = 0; j < my_tiles; ++j) {
            const long long e = ((long long)blockIdx.x + (long long)j * G) * TM + row;
            const bool valid = e < a.n_rows;
            const size_t er = (size_t)(valid ? e : 0);
            float2 g[16];           // gradient w.r.t. the current layer's output, own 32 columns, true scale
            auto put_y = [&](const float* src, int ld) {        // rows of the layer input -> chunks 4 q .. 4 q + 3 of the input image (x 16)
                float2 v[16];
                const float* p = src + er * ld + 32 * q;
#pragma unroll
                for (int i = 0; i < 4; ++i) ldg256(p + 8 * i, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) f16::split(valid ? __fmul2_rn(v[4 * k + i], A2) : make_float2(0.f, 0.f), hi[i], lo[i]);
                    yimg[(4 * q + k) * TM + row] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    yimg[YI + (4 * q + k) * TM + row] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
            };
            auto read_d = [&]() {       // g = D / (256 S), own 32 columns
                tc::mbar_wait(&bars[B_D], (use - 1) & 1u);
                tc::tc_fence_after();
                tc::tmem_ld16(t_row + COL_D + 32 * q, g);
                tc::tmem_ld16(t_row + COL_D + 32 * q + 16, g + 8);
                tc::tmem_wait_ld();
                tc::tc_fence_before();
#pragma unroll
                for (int c = 0; c < 16; ++c) g[c] = __fmul2_rn(g[c], u2);
            };
            if (a.tail) {
                // ---- bare Linear 64 -> n_out: dz = g_top ----
                float gt[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) gt[c] = (valid && c < a.n_out) ? __ldg(a.g_top + er * a.ld_g + c) : 0.f;
                if (use > 0) tc::mbar_wait(&bars[B_W], (use - 1) & 1u);
                {
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        f16::split(make_float2((q ? gt[8 + 2 * i] : gt[2 * i]) * S, (q ? gt[9 + 2 * i] : gt[2 * i + 1]) * S), hi[i], lo[i]);
                    zimg[q * TM + row] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    zimg[ZI + q * TM + row] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
                put_y(a.y[a.n_hidden - 1], W);
                tc::fence_async_smem();
                warp_arrive(&bars[B_A], lane);
                ++use;
                if (q == 0) {
                    float cs[32];
#pragma unroll
                    for (int c = 0; c < 32; ++c) cs[c] = c < 16 ? gt[c] : 0.f;
                    acc_b[C64B_MAX] += warp_colsum32(cs, lane);
                }
                read_d();
            } else {
                const float* p = a.g_top + er * a.ld_g + 32 * q;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    g[4 * i] = g[4 * i + 1] = g[4 * i + 2] = g[4 * i + 3] = make_float2(0.f, 0.f);
                    if (valid) ldg256(p + 8 * i, g[4 * i], g[4 * i + 1], g[4 * i + 2], g[4 * i + 3]);
                }
            }
#pragma unroll 1
            for (int l = a.n_hidden - 1; l >= 0; --l) {
                const bool norm = a.scale[l] != nullptr, act = a.act[l] != 0;
                const float sv = norm ? __ldg(a.scale[l]) : 1.f, mv = norm ? __ldg(a.shift[l]) : 0.f;
                const float inv_s = sv != 0.f ? 1.f / sv : 0.f;
                const float2 is2 = make_float2(inv_s, inv_s), nsh2 = make_float2(-mv * inv_s, -mv * inv_s), s22 = make_float2(sv, sv);
                float2 yl[16], nv[16];
                {
                    const float* p = a.y[l] + er * W + 32 * q;
#pragma unroll
                    for (int i = 0; i < 4; ++i) ldg256(p + 8 * i, yl[4 * i], yl[4 * i + 1], yl[4 * i + 2], yl[4 * i + 3]);
                }
                const float sdv = (norm && valid) ? __ldg(a.sd[l] + er) : 0.f;
                float2 ps2 = make_float2(0.f, 0.f), pm2 = ps2, sum2 = ps2, dot2 = ps2;
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    if (!valid) g[c] = make_float2(0.f, 0.f);
                    act_bwd_pair(g[c], nv[c], yl[c], act, is2, nsh2);
                    if (norm) {
                        ps2 = __ffma2_rn(g[c], nv[c], ps2);
                        pm2 = __fadd2_rn(pm2, g[c]);
                        g[c] = __fmul2_rn(g[c], s22);
                        sum2 = __fadd2_rn(sum2, g[c]);
                        dot2 = __ffma2_rn(g[c], nv[c], dot2);
                    }
                }
                if (norm) {
                    if (valid) {
#pragma unroll
                        for (int k = 0; k < C64B_MAX; ++k)
                            if (k == l) { acc_s[k] += (double)(ps2.x + ps2.y); acc_m[k] += (double)(pm2.x + pm2.y); }
                    }
                    xch[((use & 1u) * TM + row) * 2 + q] = make_float2(sum2.x + sum2.y, dot2.x + dot2.y);
                    group_sync(bar_id, 64);
                    const float2 o = xch[((use & 1u) * TM + row) * 2 + (q ^ 1)];
                    const float sum_dn = (sum2.x + sum2.y) + o.x, dot = (dot2.x + dot2.y) + o.y;
                    const float inv_den = 1.f / (sdv + NORM_EPS);
                    const float mean_dn = sum_dn / (float)W;
                    const float coef = sdv > 0.f ? dot / ((float)(W - 1) * sdv) : 0.f;
                    const float2 nm2 = make_float2(-mean_dn, -mean_dn), id2 = make_float2(inv_den, inv_den), nc2 = make_float2(-coef, -coef);
#pragma unroll
                    for (int c = 0; c < 16; ++c) g[c] = valid ? __ffma2_rn(nv[c], nc2, __fmul2_rn(__fadd2_rn(g[c], nm2), id2)) : make_float2(0.f, 0.f);
                }
                if (use > 0) tc::mbar_wait(&bars[B_W], (use - 1) & 1u);       // the previous phase's MMAs have read both images
                {
                    uint32_t hi[16], lo[16];
                    float cs[32];
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const float2 v = __fmul2_rn(g[c], S2);
                        cs[2 * c] = v.x; cs[2 * c + 1] = v.y;
                        f16::split(v, hi[c], lo[c]);
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        zimg[(4 * q + k) * TM + row] = make_uint4(hi[4 * k], hi[4 * k + 1], hi[4 * k + 2], hi[4 * k + 3]);
                        zimg[ZI + (4 * q + k) * TM + row] = make_uint4(lo[4 * k], lo[4 * k + 1], lo[4 * k + 2], lo[4 * k + 3]);
                    }
                    if (l > 0) put_y(a.y[l - 1], W); else put_y(a.x, a.ldx);
                    tc::fence_async_smem();
                    warp_arrive(&bars[B_A], lane);
                    ++use;
                    const float bsum = warp_colsum32(cs, lane);
#pragma unroll
                    for (int k = 0; k < C64B_MAX; ++k)
                        if (k == l) acc_b[k] += bsum;
                }
                if (l > 0 || a.dx != nullptr) read_d();
            }
            // ---- gradient w.r.t. the stack's input ----
            if (a.dx != nullptr && valid) {
                if (a.dx_mode == 2) {
                    float* oa = a.dx + (size_t)__ldg(a.ia + er) * W + 32 * q;
                    float* ob = a.dx + (size_t)__ldg(a.ib + er) * W + 32 * q;
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        atomicAdd(oa + 2 * c, g[c].x); atomicAdd(oa + 2 * c + 1, g[c].y);
                        atomicAdd(ob + 2 * c, g[c].x); atomicAdd(ob + 2 * c + 1, g[c].y);
                    }
                } else {
                    float* o = a.dx + er * W + 32 * q;
                    if (a.dx_mode == 1) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float2 o0, o1, o2, o3;
                            ldg256(o + 8 * i, o0, o1, o2, o3);
                            g[4 * i] = __fadd2_rn(g[4 * i], o0); g[4 * i + 1] = __fadd2_rn(g[4 * i + 1], o1);
                            g[4 * i + 2] = __fadd2_rn(g[4 * i + 2], o2); g[4 * i + 3] = __fadd2_rn(g[4 * i + 3], o3);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) stg256(o + 8 * i, g[4 * i], g[4 * i + 1], g[4 * i + 2], g[4 * i + 3]);
                }
            }
        }
        // ---- flush: per layer dW (lane = input channel m < 64, column = output channel n -> gW[n * 64 + m]), db, the norm scalars ----
        if (my_tiles > 0) {
            tc::mbar_wait(&bars[B_W], (use - 1) & 1u);
            tc::tc_fence_after();
            const float u = inv_S * (1.f / f16::A_SCALE);
#pragma unroll 1
            for (int l = 0; l < a.n_hidden; ++l) {
                if (a.gW[l] != nullptr) {
#pragma unroll 1
                    for (int cc = 0; cc < 2; ++cc) {
                        float v[16];
                        tc::tmem_ld16(t_row + COL_ACC + 64 * l + 32 * q + 16 * cc, v);
                        tc::tmem_wait_ld();
                        if (row < W) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) atomicAdd(a.gW[l] + (size_t)(32 * q + 16 * cc + i) * W + row, v[i] * u);
                        }
                    }
                }
                float bl = 0.f;
#pragma unroll
                for (int k = 0; k < C64B_MAX; ++k)
                    if (k == l) bl = acc_b[k];
                if (a.gb[l] != nullptr) atomicAdd(a.gb[l] + 32 * q + lane, bl * inv_S);
            }
            if (a.tail) {
                if (a.gW[a.n_hidden] != nullptr && q == 0) {
                    float v[16];
                    tc::tmem_ld16(t_row + COL_ACC_TAIL, v);
                    tc::tmem_wait_ld();
                    if (row < W) {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (i < a.n_out) atomicAdd(a.gW[a.n_hidden] + (size_t)i * W + row, v[i] * u);
                    }
                }
                if (a.gb[a.n_hidden] != nullptr && q == 0 && lane < a.n_out) atomicAdd(a.gb[a.n_hidden] + lane, acc_b[C64B_MAX]);
            }
        }
        {
#pragma unroll 1
            for (int l = 0; l < a.n_hidden; ++l) {
                double v0 = 0., v1 = 0.;
#pragma unroll
                for (int k = 0; k < C64B_MAX; ++k)
                    if (k == l) { v0 = acc_s[k]; v1 = acc_m[k]; }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { v0 += __shfl_xor_sync(0xffffffffu, v0, o); v1 += __shfl_xor_sync(0xffffffffu, v1, o); }
                if (lane == 0) { red[(2 * l) * 8 + warp] = v0; red[(2 * l + 1) * 8 + warp] = v1; }
            }
            group_sync(BAR_WORKERS, NW);
            if (tid < 2 * a.n_hidden && my_tiles > 0) {
                double s = 0.;
                for (int w = 0; w < NW / 32; ++w) s += red[tid * 8 + w];
                float* dst = (tid & 1) ? a.g_shift[tid >> 1] : a.g_scale[tid >> 1];
                if (dst != nullptr) atomicAdd(dst, (float)s);
            }
        }
    } else if (warp == NW / 32 && lane == 0) {
        constexpr uint32_t ID_G = idesc_mn(TM, W, 0, 1), ID_W = idesc_mn(128, W, 1, 1), ID_WT = idesc_mn(128, NTAIL, 1, 1);
        const uint32_t sY = tc::smem_u32(yimg), sZ = tc::smem_u32(zimg), sW0 = tc::smem_u32(smem_u + OFF_W);
        uint32_t use = 0;
        for (int j = 0; j < my_tiles; ++j) {
            for (int ph = 0; ph < n_phase; ++ph) {
                const bool is_tail = a.tail && ph == 0;
                const int l = a.tail ? a.n_hidden - ph : a.n_hidden - 1 - ph;       // hidden layer of this phase (unused for the tail)
                tc::mbar_wait(&bars[B_A], use & 1u);
                tc::tc_fence_after();
                {
                    // D = dz W: A = the dz image K-major, B = the forward image (rows = output channels) read MN-major
                    const uint32_t sW = is_tail ? sW0 + C64B_MAX * IMG_WORDS * 4 : sW0 + (uint32_t)l * IMG_WORDS * 4;
                    const uint32_t lo_off = is_tail ? TAIL_WORDS * 2 : IMG_WORDS * 2;
                    const uint32_t sbo = is_tail ? NTAIL * 16 : W * 16;
                    const int nks = is_tail ? 1 : W / 16;
                    bool acc = false;
                    for (int p = 0; p < np; ++p) {
                        const int pa = (np == 1) ? 0 : (p == 0 ? 1 : 0), pb = (np == 1) ? 0 : (p == 1 ? 1 : 0);
                        const uint64_t ad = tc::smem_desc(sZ + pa * (Z_WORDS * 4), TM * 16, 128);
                        const uint64_t bd = tc::smem_desc(sW + pb * lo_off, 128, sbo);
                        for (int ks = 0; ks < nks; ++ks) {
                            f16::mma_ss(tmem + COL_D, ad + (uint64_t)((ks * 2 * TM * 16) >> 4), bd + (uint64_t)((ks * 256) >> 4), ID_G, acc);
                            acc = true;
                        }
                    }
                }
                tc::mma_commit(&bars[B_D]);
                {
                    // dW += input^T dz: both images MN-major, K = the tile's 128 rows
                    const uint32_t dcol = tmem + (is_tail ? COL_ACC_TAIL : COL_ACC + 64 * (uint32_t)l);
                    for (int p = 0; p < np; ++p) {
                        const int pa = (np == 1) ? 0 : (p == 0 ? 1 : 0), pb = (np == 1) ? 0 : (p == 1 ? 1 : 0);
                        const uint64_t ad = tc::smem_desc(sY + pa * (Y_WORDS * 4), 128, TM * 16);
                        const uint64_t bd = tc::smem_desc(sZ + pb * (Z_WORDS * 4), 128, TM * 16);
#pragma unroll
                        for (int ks = 0; ks < TM / 16; ++ks)
                            f16::mma_ss(dcol, ad + (uint64_t)((ks * 256) >> 4), bd + (uint64_t)((ks * 256) >> 4), is_tail ? ID_WT : ID_W,
                                        j > 0 || p > 0 || ks > 0);
                    }
                }
                tc::mma_commit(&bars[B_W]);
                ++use;
            }
        }
    }
    __syncwarp();
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// =============================================================================================
// host side
// =============================================================================================
static int g_node_bwd_f16 = 1;

bool node_bwd_f16_supported(const rgnn_conv& c, const ConvDims& d) {
    const rgnn_linear& L = c.upd.layer[0];
    return g_node_bwd_f16 && mp_f16_supported(d) && d.cn == 64 && d.h == 128 && c.upd.n == 1 && L.in_features == 2 * d.cn &&
           L.out_features == d.cn && conv_proj_f16_floats(d) > 0 && f16_image_floats(L.in_features, L.out_features) > 0;
}

static int launch_absmax(const float* x, size_t n, unsigned* out, cudaStream_t stream) {
    RGNN_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(unsigned), stream));
    size_t blocks = (n / 4 + 255) / 256;
    if (blocks > (size_t)4 * sm_count()) blocks = (size_t)4 * sm_count();
    if (blocks < 1) blocks = 1;
    absmax_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, n, out);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

// dx += dP [W_t; W_s] and the weight gradient of msg.0's node columns; `scalar` = 4 bytes of scratch for max |dP|
int run_proj_bwd_f16(const rgnn_conv& c, const ConvDims& d, const float* dP, const float* x, int n_nodes, float* dx, float* scalar,
                     cudaStream_t stream) {
    if (n_nodes <= 0) return RGNN_OK;
    const rgnn_linear& m0 = c.msg.layer[0];
    int rc = launch_absmax(dP, (size_t)n_nodes * 2 * d.h, reinterpret_cast<unsigned*>(scalar), stream);
    if (rc) return rc;
    static PerDeviceOnce once;
    if (once.needed()) {
        RGNN_CHECK_CUDA(cudaFuncSetAttribute(proj_bwd_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pbf::SMEM));
        RGNN_CHECK_CUDA(cudaFuncSetAttribute(upd_bwd_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ubf::SMEM));
        once.mark();
    }
    ProjBwdArgs a;
    a.dP = dP; a.x = x; a.dx = dx;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(m0.weight_t + conv_msg0_proj16_offset(d));
    a.w_t = w; a.w_s = w + pbf::W_WORDS;
    a.gW = m0.grad_weight; a.ldW = m0.in_features;
    a.gmax = scalar; a.n_rows = n_nodes; a.passes = mp_f16_passes();
    const int n_tiles = (n_nodes + nbf::TM - 1) / nbf::TM;
    proj_bwd_f16_kernel<<<n_tiles < sm_count() ? n_tiles : sm_count(), nbf::NTHREADS, pbf::SMEM, stream>>>(a);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

int run_upd_bwd_f16(const rgnn_conv& c, const ConvDims& d, int n_nodes, const float* x, const float* agg, const float* u, const float* sd,
                    float* dx, float* dagg, float* scalar, cudaStream_t stream) {
    if (n_nodes <= 0) return RGNN_OK;
    const rgnn_linear& L = c.upd.layer[0];
    int rc = launch_absmax(dx, (size_t)n_nodes * d.cn, reinterpret_cast<unsigned*>(scalar), stream);
    if (rc) return rc;
    static PerDeviceOnce once;
    if (once.needed()) {
        RGNN_CHECK_CUDA(cudaFuncSetAttribute(proj_bwd_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pbf::SMEM));
        RGNN_CHECK_CUDA(cudaFuncSetAttribute(upd_bwd_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ubf::SMEM));
        once.mark();
    }
    UpdBwdArgs a;
    a.x = x; a.agg = agg; a.u = u; a.sd = sd; a.dx = dx; a.dagg = dagg;
    a.w_u = reinterpret_cast<const uint32_t*>(f16_weights(L));
    a.scale = L.norm_scale; a.shift = L.norm_shift;
    a.gW = L.grad_weight; a.gb = L.grad_bias; a.g_scale = L.grad_norm_scale; a.g_shift = L.grad_norm_shift;
    a.gmax = scalar; a.n_rows = n_nodes; a.act = L.activation; a.passes = mp_f16_passes();
    const int n_tiles = (n_nodes + nbf::TM - 1) / nbf::TM;
    upd_bwd_f16_kernel<<<n_tiles < sm_count() ? n_tiles : sm_count(), nbf::NTHREADS, ubf::SMEM, stream>>>(a);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

bool chain64_bwd_f16_supported(const rgnn_stack& s) {
    if (!g_node_bwd_f16 || !chain64_supported(s)) return false;
    const rgnn_linear& last = s.layer[s.n - 1];
    const bool tail = last.norm_scale == nullptr && !last.activation;
    const int n_hidden = tail ? s.n - 1 : s.n;
    if (n_hidden < 1 || n_hidden > C64B_MAX || (tail && last.out_features > cbf::NTAIL)) return false;
    for (int i = 0; i < n_hidden; ++i) {
        const rgnn_linear& L = s.layer[i];
        if (L.in_features != 64 || L.out_features != 64 || L.norm_scale == nullptr) return false;
    }
    return true;
}

// drop-in for tc_stack_bwd on the 64-wide stacks (same saved activations): x = the stack's input rows (or save.x_in), y_out = its
// output when the last layer is a hidden one; `scalar` = 4 bytes of scratch for max |g_top|
int run_chain64_bwd_f16(const rgnn_stack& s, const TcSave& save, const float* x_rows, const float* y_out, const float* g_top, int n_rows,
                        float* dx, int dx_mode, const int* ia, const int* ib, float* scalar, cudaStream_t stream) {
    if (n_rows <= 0) return RGNN_OK;
    const rgnn_linear& last = s.layer[s.n - 1];
    const bool tail = last.norm_scale == nullptr && !last.activation;
    const int n_hidden = tail ? s.n - 1 : s.n;
    Chain64BwdArgs a;
    memset(&a, 0, sizeof(a));
    a.n_rows = n_rows; a.n_hidden = n_hidden; a.tail = tail ? 1 : 0; a.n_out = tail ? last.out_features : 64;
    a.x = x_rows != nullptr ? x_rows : save.x_in; a.ldx = 64;
    RGNN_REQUIRE(a.x != nullptr, "chain64 backward: no input rows");
    for (int i = 0; i < n_hidden; ++i) {
        const rgnn_linear& L = s.layer[i];
        a.y[i] = (i == s.n - 1) ? y_out : save.y[i];
        a.sd[i] = save.sd[i];
        RGNN_REQUIRE(a.y[i] != nullptr && a.sd[i] != nullptr, "chain64 backward: layer %d was not saved by the forward", i);
        a.w[i] = reinterpret_cast<const uint32_t*>(f16_weights(L));
        a.scale[i] = L.norm_scale; a.shift[i] = L.norm_shift; a.act[i] = L.activation;
        a.gW[i] = L.grad_weight; a.gb[i] = L.grad_bias; a.g_scale[i] = L.grad_norm_scale; a.g_shift[i] = L.grad_norm_shift;
    }
    if (tail) {
        a.w[n_hidden] = reinterpret_cast<const uint32_t*>(f16_weights(last));
        a.gW[n_hidden] = last.grad_weight; a.gb[n_hidden] = last.grad_bias;
    }
    a.g_top = g_top; a.ld_g = tail ? last.out_features : 64;
    a.dx = dx; a.dx_mode = dx_mode; a.ia = ia; a.ib = ib;
    a.gmax = scalar; a.passes = mp_f16_passes();
    int rc = launch_absmax(g_top, (size_t)n_rows * a.ld_g, reinterpret_cast<unsigned*>(scalar), stream);
    if (rc) return rc;
    static PerDeviceOnce once;
    if (once.needed()) {
        RGNN_CHECK_CUDA(cudaFuncSetAttribute(chain64_bwd_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cbf::SMEM));
        once.mark();
    }
    const int n_tiles = (n_rows + nbf::TM - 1) / nbf::TM;
    chain64_bwd_f16_kernel<<<n_tiles < sm_count() ? n_tiles : sm_count(), nbf::NTHREADS, cbf::SMEM, stream>>>(a);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

int node_bwd_f16_set_option(const char* name, int value) {
    if (strcmp(name, "f16_node_bwd") == 0 && (value == 0 || value == 1)) { g_node_bwd_f16 = value; return 1; }
    return 0;
}
int node_bwd_f16_get_option(const char* name) { return strcmp(name, "f16_node_bwd") == 0 ? g_node_bwd_f16 : -2; }

}  // namespace rgnn
