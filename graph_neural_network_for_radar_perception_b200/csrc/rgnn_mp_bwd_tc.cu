// Backward of the message function of one residual_graph_conv_block on the tensor cores
// (replaces torch autograd over gnn_blocks.py:106-113, reference gnn/training.py:81).
//
// Per tile of 128 target-major edges the forward is RECOMPUTED on chip (nothing per-edge was kept by the forward pass)
// and the gradient walks back through it; every A operand lives in TMEM, thread = TMEM lane = edge row:
//   G1  D1 = emb W_e^T                       (TS, W_e K-major)        epi1: + P_t[tgt] + P_s[src] -> norm -> act -> y1
//   G2  D2 = y1 W_2^T                        (TS, W_2 K-major)        epi2: + b2 -> norm -> act (mask only);
//                                                                           d(msg) = dagg[tgt] -> act' -> norm' -> dz2
//   G3  D3 = dz2 W_2      = d(y1)            (TS, W_2^T K-major)      epi3: act' -> norm' -> dz1
//   G4  D4 = dz1 W_e      = d(emb)           (TS, W_e^T K-major)      epi4: demb (+)= D4
// Four hi/lo-split weight images (W_e, W_2 and their transposes, 64 KB each) do not fit next to the staging tiles, and
// 32-bit operands have no MN-major form that could read the forward images transposed, so the images are STREAMED:
// a load warp copies each one L2 -> shared memory with one cp.async.bulk into a two-slot ring (mbarrier full / empty),
// always one GEMM ahead of the MMA warp.
// The weight gradients need the edge index as the REDUCTION dimension of both operands, i.e. y1, dz1, dz2 as
// shared-memory operands with hi/lo copies (256 KB per tile): they cannot live on the SM next to all this, so this
// kernel writes y1, dz1 and dz2 to a scratch buffer and rgnn_wgrad_tc.cu contracts them over all edges; the per-node
// gradient of the hoisted projection, dP = [sum_{e: tgt=n} dz1_e | sum_{e: src=n} dz1_e], is a gather over the same scratch
// (deterministic: no floating-point atomics anywhere in the message backward).
#include <vector>

#include "rgnn_model.h"
#include "rgnn_tc.cuh"
#include "rgnn_tile.cuh"
#include "rgnn_tc_rows.cuh"

namespace rgnn {

extern int g_rowmlp_profile;

struct MpBwdArgs {
    const float* emb;       // (E, CE) target-major
    const float* P;         // (N, 2H)
    const float* dagg;      // (N, CN) gradient w.r.t. the aggregated messages
    const int* tgt;
    const int* src;
    const float* wpack;     // [W1e_hi | W1e_lo | W2_hi | W2_lo | W2T_hi | W2T_lo | W1eT_hi | W1eT_lo], chunk-major (rgnn_tc.cuh)
    const float* s1; const float* m1;     // channel_normalization scalars of msg.0 / msg.1 (nullptr = no norm)
    const float* b2; const float* s2; const float* m2;
    float* y1_out;          // (E, H)  scratch
    float* dz1_out;         // (E, H)
    float* dz2_out;         // (E, CN)
    float* demb;            // (E, CE) accumulated over the layers
    float* g_s1; float* g_m1; float* g_s2; float* g_m2;     // gradients of the norm scalars (+=), nullable
    int n_edges;
    int act1, act2;
    int demb_accumulate;
    int passes;
    long long* prof;        // developer aid (rgnn_set_option("debug", 8)): [grid][10] cycle counters of worker thread 0
};

namespace tc {
__device__ __forceinline__ void mbar_expect_tx_b(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_b(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
}  // namespace tc

constexpr int MB_TM = 128, MB_CE = 64, MB_H = 128, MB_CN = 64;
constexpr int MB_NW = 256;                       // worker threads: (row, half of the columns)
constexpr int MB_W1 = MB_CE * MB_H, MB_W2 = MB_H * MB_CN;
static_assert(MB_W1 == MB_W2, "all four weight images have the same size");
constexpr int MB_IMG = 2 * MB_W1;                         // floats of one image (hi | lo): 64 KB
constexpr int MB_OFF_RING = 0;                            // two ring slots
constexpr int MB_OFF_PS = MB_OFF_RING + 2 * MB_IMG;       // P_s[src] rows of a tile (XOR-swizzled 16-byte chunks)
constexpr int MB_OFF_EB = MB_OFF_PS + MB_TM * MB_H;       // raw emb rows of a tile (XOR-swizzled)
constexpr int MB_OFF_RED = MB_OFF_EB + MB_TM * MB_CE;     // 4 x 8 doubles: per-warp partials of the norm-scalar gradients
constexpr int MB_OFF_BAR = MB_OFF_RED + 64;
constexpr int MB_OFF_SLOT = MB_OFF_BAR + 16;            // bars: done[4] | full[2] | empty[2]
constexpr size_t MB_SMEM = (size_t)(MB_OFF_SLOT + 2) * 4;
static_assert(MB_SMEM <= 227 * 1024 && (MB_OFF_BAR % 2) == 0 && (MB_OFF_RED % 2) == 0, "shared memory budget / alignment");
// TMEM columns
constexpr uint32_t MB_COL_Y = 0;         // D1, then y1_hi, then dz1_hi                 [0,128)
constexpr uint32_t MB_COL_YLO = 128;     // y1_lo, then dz1_lo                          [128,256)
constexpr uint32_t MB_COL_D2 = 256;      // D2, then dz2_hi, then D4                    [256,320)
constexpr uint32_t MB_COL_Z2LO = 320;    // dz2_lo; epi3 statistics exchange            [320,384)
constexpr uint32_t MB_COL_D3 = 384;      // emb_hi [384,448) emb_lo [448,512); epi1/epi2 statistics exchange; then D3
constexpr uint32_t MB_COL_ELO = 448;

__global__ void __launch_bounds__(MB_NW + 128, 1) mp_edge_bwd_tc_kernel(const __grid_constant__ MpBwdArgs a) {
    constexpr int TM = MB_TM, CE = MB_CE, H = MB_H, CN = MB_CN, NW = MB_NW;
    constexpr int NMMA = NW + 32;
    extern __shared__ __align__(1024) float smem[];
    float* ring = smem + MB_OFF_RING;
    float* Ps = smem + MB_OFF_PS;
    float* Eb = smem + MB_OFF_EB;
    double* red = reinterpret_cast<double*>(smem + MB_OFF_RED);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MB_OFF_BAR);
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem + MB_OFF_SLOT);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = tid & 127, q = (tid >> 7) & 1;
    const int bar_id = 1 + (row >> 5);

    uint64_t* full = bars + 4;
    uint64_t* empty = bars + 6;
    if (tid == 0) {
        for (int i = 0; i < 8; ++i) tc::mbar_init(&bars[i], 1);
        tc::mbar_init_fence();
    }
    if (warp == 0) tc::tmem_alloc(slot, 512);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *slot;
    const int n_tiles = (a.n_edges + TM - 1) / TM;
    const int np = a.passes == 1 ? 1 : 3;

    if (tid >= NW) {
        // =========================== MMA issue warpgroup (only its first warp works) ===========================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (tid < NW + 32) {
            constexpr uint32_t ID_N128 = tc::idesc_tf32(TM, 128);      // G1 (N = H) and G3 (N = H): K = 64
            constexpr uint32_t ID_N64 = tc::idesc_tf32(TM, 64);        // G2 (N = CN) and G4 (N = CE): K = 128
            static_assert(H == 128 && CN == 64 && CE == 64, "image shapes");
            uint32_t cnt = 0;
            // one GEMM: A (K columns of TMEM, hi at a_hi / lo at a_lo) x the image in the next ring slot ([K/4][N][4], hi | lo)
            auto gemm = [&](uint32_t d_col, uint32_t a_hi, uint32_t a_lo, int N, int K, uint32_t idesc, uint64_t* bar) {
                const uint32_t sl = cnt & 1u, par = (cnt >> 1) & 1u;
                ++cnt;
                tc::mbar_wait(&full[sl], par);
                tc::tc_fence_after();
                const uint32_t sB = tc::smem_u32(ring + sl * MB_IMG);
                const uint32_t lbo = (uint32_t)N * 16u;
                bool acc = false;
                for (int p = 0; p < np; ++p) {     // 3xTF32, small terms first: lo*hi, hi*lo, hi*hi
                    const int pa = (np == 1) ? 0 : (p == 0 ? 1 : 0), pb = (np == 1) ? 0 : (p == 1 ? 1 : 0);
                    const uint32_t acol = tmem + (pa ? a_lo : a_hi);
                    const uint64_t bd0 = tc::smem_desc(sB + pb * (MB_IMG * 2), lbo, 128);
#pragma unroll 8
                    for (int ks = 0; ks < K / 8; ++ks) {      // (one lane issues: keep the k-steps independent, see rgnn_rowmlp_tc.cu)
                        tc::mma_tf32_ts(tmem + d_col, acol + ks * 8, bd0 + (uint64_t)((ks * 2 * lbo) >> 4), idesc, acc);
                        acc = true;
                    }
                }
                tc::mma_commit(&empty[sl]);       // the slot may be refilled once these MMAs have read it
                tc::mma_commit(bar);
            };
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                group_sync(BAR_Y_READY, NMMA);
                tc::tc_fence_after();
                if (lane == 0) gemm(MB_COL_Y, MB_COL_D3, MB_COL_ELO, H, CE, ID_N128, &bars[0]);        // G1: emb x W_e
                __syncwarp();
                group_sync(BAR_Y_READY, NMMA);
                tc::tc_fence_after();
                if (lane == 0) gemm(MB_COL_D2, MB_COL_Y, MB_COL_YLO, CN, H, ID_N64, &bars[1]);        // G2: y1 x W_2
                __syncwarp();
                group_sync(BAR_Y_READY, NMMA);
                tc::tc_fence_after();
                if (lane == 0) gemm(MB_COL_D3, MB_COL_D2, MB_COL_Z2LO, H, CN, ID_N128, &bars[2]);      // G3: dz2 x W_2^T
                __syncwarp();
                group_sync(BAR_Y_READY, NMMA);
                tc::tc_fence_after();
                if (lane == 0) gemm(MB_COL_D2, MB_COL_Y, MB_COL_YLO, CE, H, ID_N64, &bars[3]);        // G4: dz1 x W_e^T
                __syncwarp();
            }
        } else if (tid < NW + 64) {
            // =========================== weight load warp ===========================
            if (lane == 0) {
                uint32_t cnt = 0;
                for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                    for (int j = 0; j < 4; ++j, ++cnt) {      // image order in wpack = order of use
                        const uint32_t sl = cnt & 1u, par = (cnt >> 1) & 1u;
                        tc::mbar_wait(&empty[sl], par ^ 1u);
                        tc::mbar_expect_tx_b(&full[sl], MB_IMG * 4);
                        tc::bulk_g2s_b(ring + sl * MB_IMG, a.wpack + (size_t)j * MB_IMG, MB_IMG * 4, &full[sl]);
                    }
                }
            }
        }
    } else {
        // =========================== worker warps ===========================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        const uint32_t t_row = tmem + ((uint32_t)(row & ~31) << 16);
        const bool norm1 = a.s1 != nullptr, norm2 = a.s2 != nullptr;
        const float s1v = norm1 ? __ldg(a.s1) : 1.f, m1v = norm1 ? __ldg(a.m1) : 0.f;
        const float s2v = norm2 ? __ldg(a.s2) : 1.f, m2v = norm2 ? __ldg(a.m2) : 0.f;
        const float inv_s1 = s1v != 0.f ? 1.f / s1v : 0.f;
        const bool act1 = a.act1 != 0, act2 = a.act2 != 0;
        double acc_s1 = 0., acc_m1 = 0., acc_s2 = 0., acc_m2 = 0.;

        constexpr int RPW = TM / (NW / 32);     // P_s rows staged per warp
        auto load_idx = [&](int tile, int& t_my, int& s_my) {
            const int e = tile * TM + row;
            t_my = e < a.n_edges ? __ldg(a.tgt + e) : -1;
            const int es = tile * TM + warp * RPW + (lane % RPW);
            s_my = es < a.n_edges ? __ldg(a.src + es) : -1;
        };
        auto stage_ps = [&](int s_my) {     // whole 512-byte P_s rows per warp instruction (cp.async, 16 B per lane)
#pragma unroll
            for (int i = 0; i < RPW; ++i) {
                const int r = warp * RPW + i;
                const int sn = __shfl_sync(0xffffffffu, s_my, i);
                if (sn >= 0) cp_async16(Ps + r * H + ((lane ^ (r & 7)) << 2), a.P + (size_t)sn * (2 * H) + H + 4 * lane);
            }
            cp_async_commit();
        };
        auto stage_emb = [&](int tile) {    // 128 rows x 256 B, two rows per warp instruction
            const int row0 = tile * TM;
#pragma unroll
            for (int k = 0; k < (TM * CE / 4) / NW; ++k) {
                const int i = tid + k * NW, r = i >> 4, c4 = i & 15;
                if (row0 + r < a.n_edges) cp_async16(Eb + r * CE + ((c4 ^ (r & 7)) << 2), a.emb + (size_t)(row0 + r) * CE + 4 * c4);
            }
            cp_async_commit();
        };

        int t_my = -1, s_my = -1;       // target of this thread's edge row / source of the row it stages, for the tile in `prepare`
        int n_s = -1;                   // source ids of the tile after that (P_s rows are staged after epilogue 1)
        float2 z[H / 4];                // this thread's half of the hoisted target projection P_t[tgt] (H/2 floats)
        // Everything a tile needs before G1 can be issued.  It runs while G4 of the PREVIOUS tile executes: the emb operand
        // goes to TMEM columns [384, 512), which no MMA touches between G3 and the next G1.
        auto prepare = [&](int tile) {
            const int e = tile * TM + row;
            const bool v_ok = e < a.n_edges;
            cp_async_wait<0>();
            group_sync(BAR_WORKERS, NW);            // this tile's emb and P_s rows have landed (every thread's copies)
#pragma unroll
            for (int c8 = 0; c8 < CE / 2 / 8; ++c8) {
                const int c4 = q * (CE / 8) + 2 * c8;
                const float4 v0 = *reinterpret_cast<const float4*>(Eb + row * CE + ((c4 ^ (row & 7)) << 2));
                const float4 v1 = *reinterpret_cast<const float4*>(Eb + row * CE + (((c4 + 1) ^ (row & 7)) << 2));
                float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w}, hi[8], lo[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (!v_ok) v[j] = 0.f;
                    tc::split_tf32(v[j], hi[j], lo[j]);
                }
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                             ::"r"(t_row + MB_COL_D3 + q * (CE / 2) + 8 * c8), "f"(hi[0]), "f"(hi[1]), "f"(hi[2]), "f"(hi[3]), "f"(hi[4]),
                               "f"(hi[5]), "f"(hi[6]), "f"(hi[7]) : "memory");
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                             ::"r"(t_row + MB_COL_ELO + q * (CE / 2) + 8 * c8), "f"(lo[0]), "f"(lo[1]), "f"(lo[2]), "f"(lo[3]), "f"(lo[4]),
                               "f"(lo[5]), "f"(lo[6]), "f"(lo[7]) : "memory");
            }
            {
                const float* Pt = a.P + (size_t)(v_ok ? t_my : 0) * (2 * H) + q * (H / 2);
#pragma unroll
                for (int c8 = 0; c8 < H / 2 / 8; ++c8) {
                    z[4 * c8] = z[4 * c8 + 1] = z[4 * c8 + 2] = z[4 * c8 + 3] = make_float2(0.f, 0.f);
                    if (v_ok) ldg256(Pt + 8 * c8, z[4 * c8], z[4 * c8 + 1], z[4 * c8 + 2], z[4 * c8 + 3]);
                }
            }
            const int nx = tile + (int)gridDim.x;
            int n_t = -1;
            n_s = -1;
            if (nx < n_tiles) load_idx(nx, n_t, n_s);
            group_sync(BAR_WORKERS, NW);            // everyone has read its emb row: the staging tile may be refilled
            if (nx < n_tiles) stage_emb(nx);
            tc::tmem_wait_st();
            tc::tc_fence_before();
            return n_t;
        };
        uint32_t phase = 0;
        long long pt[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tlast = 0;
        const bool profiling = a.prof != nullptr && tid == 0;
        auto tick = [&](int i) {
            if (profiling) { const long long now = clock64(); pt[i] += now - tlast; tlast = now; }
        };
        if (profiling) tlast = clock64();
        int t_next = -1;
        if ((int)blockIdx.x < n_tiles) {
            load_idx(blockIdx.x, t_my, s_my);
            stage_emb(blockIdx.x);
            stage_ps(s_my);
            t_next = prepare(blockIdx.x);
        }
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, phase ^= 1) {
            const int e_my = tile * TM + row;
            const bool valid = e_my < a.n_edges;
            const int next = tile + (int)gridDim.x;
            const bool has_next = next < n_tiles;
            const int t_cur = t_my;                 // dagg[target] is gathered in epilogue 2
            bar_arrive(BAR_Y_READY, NMMA);          // -> G1 (operand written by prepare())
            tick(0);
            tick(1);

            // ---- epilogue 1: z1 = D1 + P_t + P_s -> norm -> act -> y1 (TMEM hi | lo, and the scratch buffer) ----
            tc::mbar_wait(&bars[0], phase);
            tc::tc_fence_after();
            tick(2);
#pragma unroll
            for (int c = 0; c < H / 2; c += 16) {
                float2 d[8];
                tc::tmem_ld16(t_row + MB_COL_Y + q * (H / 2) + c, d);
                tc::tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c4 = q * (H / 8) + c / 4 + j;
                    const float4 s0 = *reinterpret_cast<const float4*>(Ps + row * H + ((c4 ^ (row & 7)) << 2));
                    z[c / 2 + 2 * j] = __fadd2_rn(__fadd2_rn(z[c / 2 + 2 * j], d[2 * j]), make_float2(s0.x, s0.y));
                    z[c / 2 + 2 * j + 1] = __fadd2_rn(__fadd2_rn(z[c / 2 + 2 * j + 1], d[2 * j + 1]), make_float2(s0.z, s0.w));
                }
            }
            float sd1 = 0.f;
            if (norm1) {
                float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int c = 0; c < H / 4; ++c) s2 = __fadd2_rn(s2, z[c]);
                const float mean = row_allreduce<2>(s2.x + s2.y, t_row + MB_COL_D3, q, bar_id) / (float)H;
                const float2 nm = make_float2(-mean, -mean);
                float2 ss2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int c = 0; c < H / 4; ++c) {
                    z[c] = __fadd2_rn(z[c], nm);
                    ss2 = __ffma2_rn(z[c], z[c], ss2);
                }
                const float ss = row_allreduce<2>(ss2.x + ss2.y, t_row + MB_COL_D3 + 2, q, bar_id);
                sd1 = sqrtf(ss / (float)(H - 1));
                const float k = s1v / (sd1 + NORM_EPS);
                const float2 k2 = make_float2(k, k), sh2 = make_float2(m1v, m1v);
#pragma unroll
                for (int c = 0; c < H / 4; ++c) z[c] = __ffma2_rn(z[c], k2, sh2);
            }
            if (act1) {
                const float2 sl = make_float2(LEAKY, LEAKY);
#pragma unroll
                for (int c = 0; c < H / 4; ++c) {
                    const float2 t = __fmul2_rn(z[c], sl);
                    z[c].x = fmaxf(z[c].x, t.x);
                    z[c].y = fmaxf(z[c].y, t.y);
                }
            }
#pragma unroll
            for (int c = 0; c < H / 2; c += 16) {
                float2 hi[8], lo[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) tc::split_tf32(z[c / 2 + j], hi[j], lo[j]);
                tc::tmem_st16(t_row + MB_COL_Y + q * (H / 2) + c, hi);
                tc::tmem_st16(t_row + MB_COL_YLO + q * (H / 2) + c, lo);
            }
            tc::tmem_wait_st();
            tc::tc_fence_before();
            bar_arrive(BAR_Y_READY, NMMA);          // -> G2
            if (valid) {
                float* o = a.y1_out + (size_t)e_my * H + q * (H / 2);
#pragma unroll
                for (int c8 = 0; c8 < H / 2 / 8; ++c8) stg256(o + 8 * c8, z[4 * c8], z[4 * c8 + 1], z[4 * c8 + 2], z[4 * c8 + 3]);
            }

            // while G2 runs: d(message) = dagg[target] for this thread's half row; next tile's P_s rows
            float2 g2[CN / 4];
            {
                const float* dg = a.dagg + (size_t)(valid ? t_cur : 0) * CN + q * (CN / 2);
#pragma unroll
                for (int c8 = 0; c8 < CN / 2 / 8; ++c8) {
                    g2[4 * c8] = g2[4 * c8 + 1] = g2[4 * c8 + 2] = g2[4 * c8 + 3] = make_float2(0.f, 0.f);
                    if (valid) ldg256(dg + 8 * c8, g2[4 * c8], g2[4 * c8 + 1], g2[4 * c8 + 2], g2[4 * c8 + 3]);
                }
            }
            group_sync(BAR_WORKERS, NW);            // everyone has consumed its P_s row
            if (has_next) stage_ps(n_s);

            // ---- epilogue 2: recompute z2's statistics and activation mask; dz2 = norm'(act'(d message)) ----
            tick(3);
            tc::mbar_wait(&bars[1], phase);
            tc::tc_fence_after();
            tick(4);
            {
                float2 c2[CN / 4];
#pragma unroll
                for (int c = 0; c < CN / 2; c += 16) tc::tmem_ld16(t_row + MB_COL_D2 + q * (CN / 2) + c, c2 + c / 2);
                tc::tmem_wait_ld();
#pragma unroll
                for (int c = 0; c < CN / 4; ++c) {
                    const float2 b = a.b2 ? __ldg(reinterpret_cast<const float2*>(a.b2 + q * (CN / 2)) + c) : make_float2(0.f, 0.f);
                    c2[c] = __fadd2_rn(c2[c], b);
                }
                float sd2 = 0.f, k = 1.f, inv_den = 1.f;
                if (norm2) {
                    float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int c = 0; c < CN / 4; ++c) s2 = __fadd2_rn(s2, c2[c]);
                    const float mean = row_allreduce<2>(s2.x + s2.y, t_row + MB_COL_D3 + 4, q, bar_id) / (float)CN;
                    const float2 nm = make_float2(-mean, -mean);
                    float2 ss2 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int c = 0; c < CN / 4; ++c) {
                        c2[c] = __fadd2_rn(c2[c], nm);
                        ss2 = __ffma2_rn(c2[c], c2[c], ss2);
                    }
                    const float ss = row_allreduce<2>(ss2.x + ss2.y, t_row + MB_COL_D3 + 6, q, bar_id);
                    sd2 = sqrtf(ss / (float)(CN - 1));
                    inv_den = 1.f / (sd2 + NORM_EPS);
                    k = s2v / (sd2 + NORM_EPS);
                }
                // g2 := d(pre-activation); with a norm: dn = s * g, nv = c * inv_den   (packed f32x2: instruction-bound)
                const float2 k2 = make_float2(k, k), m22 = make_float2(m2v, m2v), id2 = make_float2(inv_den, inv_den);
                const float2 s22 = make_float2(s2v, s2v);
                float2 ps2 = make_float2(0.f, 0.f), pm2 = ps2, sum2 = ps2, dot2 = ps2;
#pragma unroll
                for (int c = 0; c < CN / 4; ++c) {
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
                    float sum_dn = sum2.x + sum2.y, dot = dot2.x + dot2.y;
                    row_allreduce2(sum_dn, dot, t_row + MB_COL_D3 + 8, q, bar_id);
                    const float mean_dn = sum_dn / (float)CN;
                    const float coef = sd2 > 0.f ? dot / ((float)(CN - 1) * sd2) : 0.f;
                    const float2 nm2 = make_float2(-mean_dn, -mean_dn), nc2 = make_float2(-coef, -coef);
#pragma unroll
                    for (int c = 0; c < CN / 4; ++c) g2[c] = __ffma2_rn(c2[c], nc2, __fmul2_rn(__fadd2_rn(g2[c], nm2), id2));
                }
            }
#pragma unroll
            for (int c = 0; c < CN / 2; c += 16) {
                float2 hi[8], lo[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) tc::split_tf32(g2[c / 2 + j], hi[j], lo[j]);
                tc::tmem_st16(t_row + MB_COL_D2 + q * (CN / 2) + c, hi);
                tc::tmem_st16(t_row + MB_COL_Z2LO + q * (CN / 2) + c, lo);
            }
            tc::tmem_wait_st();
            tc::tc_fence_before();
            bar_arrive(BAR_Y_READY, NMMA);          // -> G3
            if (valid) {
                float* o = a.dz2_out + (size_t)e_my * CN + q * (CN / 2);
#pragma unroll
                for (int c8 = 0; c8 < CN / 2 / 8; ++c8) stg256(o + 8 * c8, g2[4 * c8], g2[4 * c8 + 1], g2[4 * c8 + 2], g2[4 * c8 + 3]);
            }
            // demb rows to accumulate into (consumed by epilogue 4)
            float2 de[CE / 4];
            if (a.demb_accumulate && valid) {
                const float* o = a.demb + (size_t)e_my * CE + q * (CE / 2);
#pragma unroll
                for (int c8 = 0; c8 < CE / 2 / 8; ++c8) ldg256(o + 8 * c8, de[4 * c8], de[4 * c8 + 1], de[4 * c8 + 2], de[4 * c8 + 3]);
            } else {
#pragma unroll
                for (int c = 0; c < CE / 4; ++c) de[c] = make_float2(0.f, 0.f);
            }

            // ---- epilogue 3: d(y1) = D3 -> act' -> norm' -> dz1 (TMEM hi | lo over y1, and the scratch buffer) ----
            tick(5);
            tc::mbar_wait(&bars[2], phase);
            tc::tc_fence_after();
            tick(6);
            {
                // single pass: dn (= d pre-activation, times the norm scale) and nv (= normalised value recovered from y1) of this
                // thread's half row stay in registers between the row reduction and the final formula (packed f32x2 math)
                float2 dn[H / 4], nv[H / 4];
                const float2 is2 = make_float2(inv_s1, inv_s1), nsh2 = make_float2(-m1v * inv_s1, -m1v * inv_s1);
                const float2 s12 = make_float2(s1v, s1v);
                float2 ps2 = make_float2(0.f, 0.f), pm2 = ps2, sum2 = ps2, dot2 = ps2;
#pragma unroll
                for (int c = 0; c < H / 2; c += 16) {
                    float2 yh[8], yl[8];
                    tc::tmem_ld16(t_row + MB_COL_D3 + q * (H / 2) + c, dn + c / 2);
                    tc::tmem_ld16(t_row + MB_COL_Y + q * (H / 2) + c, yh);
                    tc::tmem_ld16(t_row + MB_COL_YLO + q * (H / 2) + c, yl);
                    tc::tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float2& g = dn[c / 2 + j];
                        act_bwd_pair(g, nv[c / 2 + j], __fadd2_rn(yh[j], yl[j]), act1, is2, nsh2);
                        if (norm1) {
                            ps2 = __ffma2_rn(g, nv[c / 2 + j], ps2);
                            pm2 = __fadd2_rn(pm2, g);
                            g = __fmul2_rn(g, s12);
                            sum2 = __fadd2_rn(sum2, g);
                            dot2 = __ffma2_rn(g, nv[c / 2 + j], dot2);
                        }
                    }
                }
                if (norm1) {
                    if (valid) { acc_s1 += (double)(ps2.x + ps2.y); acc_m1 += (double)(pm2.x + pm2.y); }
                    float sum_dn = sum2.x + sum2.y, dot = dot2.x + dot2.y;
                    row_allreduce2(sum_dn, dot, t_row + MB_COL_Z2LO, q, bar_id);
                    const float inv_den = 1.f / (sd1 + NORM_EPS);
                    const float mean_dn = sum_dn / (float)H;
                    const float coef = sd1 > 0.f ? dot / ((float)(H - 1) * sd1) : 0.f;
                    const float2 nm2 = make_float2(-mean_dn, -mean_dn), id2 = make_float2(inv_den, inv_den), nc2 = make_float2(-coef, -coef);
#pragma unroll
                    for (int c = 0; c < H / 4; ++c) dn[c] = __ffma2_rn(nv[c], nc2, __fmul2_rn(__fadd2_rn(dn[c], nm2), id2));
                }
#pragma unroll
                for (int c = 0; c < H / 2; c += 16) {
                    float2 hi[8], lo[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) tc::split_tf32(dn[c / 2 + j], hi[j], lo[j]);
                    tc::tmem_st16(t_row + MB_COL_Y + q * (H / 2) + c, hi);
                    tc::tmem_st16(t_row + MB_COL_YLO + q * (H / 2) + c, lo);
                }
                tc::tmem_wait_st();
                tc::tc_fence_before();
                bar_arrive(BAR_Y_READY, NMMA);          // -> G4
                if (valid) {
                    float* o = a.dz1_out + (size_t)e_my * H + q * (H / 2);
#pragma unroll
                    for (int c8 = 0; c8 < H / 2 / 8; ++c8) stg256(o + 8 * c8, dn[4 * c8], dn[4 * c8 + 1], dn[4 * c8 + 2], dn[4 * c8 + 3]);
                }
            }


            // ---- while G4 runs: the next tile's operand, projection rows, indices (see prepare) ----
            if (has_next) {
                if (!norm1) group_sync(bar_id, 64);     // the row partner has finished reading D3 (with a norm: row_allreduce2 did that)
                t_my = t_next;
                t_next = prepare(next);
            }
            // ---- epilogue 4: d(emb) (+)= D4 ----
            tick(7);
            tc::mbar_wait(&bars[3], phase);
            tc::tc_fence_after();
            tick(8);
            {
                float2 d[CE / 4];
#pragma unroll
                for (int c = 0; c < CE / 2; c += 16) tc::tmem_ld16(t_row + MB_COL_D2 + q * (CE / 2) + c, d + c / 2);
                tc::tmem_wait_ld();
                if (valid) {
                    float* o = a.demb + (size_t)e_my * CE + q * (CE / 2);
#pragma unroll
                    for (int c8 = 0; c8 < CE / 2 / 8; ++c8)
                        stg256(o + 8 * c8, __fadd2_rn(de[4 * c8], d[4 * c8]), __fadd2_rn(de[4 * c8 + 1], d[4 * c8 + 1]),
                               __fadd2_rn(de[4 * c8 + 2], d[4 * c8 + 2]), __fadd2_rn(de[4 * c8 + 3], d[4 * c8 + 3]));
                }
            }
            tc::tc_fence_before();
            tick(9);
        }
        if (profiling)
            for (int i = 0; i < 10; ++i) a.prof[blockIdx.x * 10 + i] = pt[i];
        // ---- gradients of the channel_normalization scalars: one double per CTA and scalar, published once ----
        {
            double v[4] = {acc_s1, acc_m1, acc_s2, acc_m2};
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
                if (dst != nullptr && n_tiles > (int)blockIdx.x) atomicAdd(dst, (float)s);
            }
        }
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// dP[n] = [ sum_{e: tgt(e) = n} dz1_e | sum_{e: src(e) = n} dz1_e ]: one warp per node, rows of 512 bytes
// ---------------------------------------------------------------------------------------------
__global__ void dproj_gather_kernel(const float* __restrict__ dz1, const int* __restrict__ row_ptr, const int* __restrict__ sptr,
                                    const int* __restrict__ slist, int n_nodes, int H, float* __restrict__ dP) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int n = blockIdx.x * wpb + (threadIdx.x >> 5); n < n_nodes; n += gridDim.x * wpb) {
        for (int c = 4 * lane; c < H; c += 128) {
            float4 at = make_float4(0.f, 0.f, 0.f, 0.f), as = at;
            const int k0 = __ldg(row_ptr + n), k1 = __ldg(row_ptr + n + 1);
            for (int k = k0; k < k1; ++k) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(dz1 + (size_t)k * H + c));
                at.x += v.x; at.y += v.y; at.z += v.z; at.w += v.w;
            }
            const int j0 = __ldg(sptr + n), j1 = __ldg(sptr + n + 1);
            for (int j = j0; j < j1; ++j) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(dz1 + (size_t)__ldg(slist + j) * H + c));
                as.x += v.x; as.y += v.y; as.z += v.z; as.w += v.w;
            }
            *reinterpret_cast<float4*>(dP + (size_t)n * 2 * H + c) = at;
            *reinterpret_cast<float4*>(dP + (size_t)n * 2 * H + H + c) = as;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// source-major index of the target-major edge list (built once per backward call): sptr (N+1), slist (E) = edge positions
// k grouped by src[k], ascending inside a group (deterministic summation order)
// ---------------------------------------------------------------------------------------------
__global__ void src_count_kernel(const int* __restrict__ src, int n_edges, int* __restrict__ cnt) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_edges; k += gridDim.x * blockDim.x) atomicAdd(cnt + __ldg(src + k), 1);
}
__global__ void src_fill_kernel(const int* __restrict__ src, int n_edges, const int* __restrict__ sptr, int* __restrict__ cursor,
                                int* __restrict__ slist) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_edges; k += gridDim.x * blockDim.x) {
        const int s = __ldg(src + k);
        slist[__ldg(sptr + s) + atomicAdd(cursor + s, 1)] = k;
    }
}
__global__ void src_sort_kernel(const int* __restrict__ sptr, int n_nodes, int* __restrict__ slist) {
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < n_nodes; n += gridDim.x * blockDim.x) {
        const int j0 = sptr[n], j1 = sptr[n + 1];
        for (int i = j0 + 1; i < j1; ++i) {       // insertion sort: the lists hold ~k+1 entries
            const int v = slist[i];
            int j = i - 1;
            while (j >= j0 && slist[j] > v) { slist[j + 1] = slist[j]; --j; }
            slist[j + 1] = v;
        }
    }
}

size_t src_index_ints(int n_nodes, int n_edges) {
    return 3 * (align256((size_t)(n_nodes + 2) * 4) / 4) + align256((size_t)(n_edges + 1) * 4) / 4 + align256(scan_ws_ints(n_nodes + 1) * 4) / 4;
}

int build_src_index(const rgnn_graph& g, int* ws, const int** sptr_out, const int** slist_out, cudaStream_t stream) {
    const size_t nn = align256((size_t)(g.n_nodes + 2) * 4) / 4;
    int* cnt = ws;
    int* cursor = cnt + nn;
    int* sptr = cursor + nn;
    int* slist = sptr + nn;
    int* scan_ws = slist + align256((size_t)(g.n_edges + 1) * 4) / 4;
    RGNN_CHECK_CUDA(cudaMemsetAsync(cnt, 0, 2 * nn * sizeof(int), stream));     // cnt and cursor
    if (g.n_edges > 0) {
        const int blocks = (g.n_edges + 255) / 256;
        src_count_kernel<<<blocks > 1184 ? 1184 : blocks, 256, 0, stream>>>(g.src, g.n_edges, cnt);
    }
    int rc = exclusive_scan(cnt, g.n_nodes + 1, sptr, scan_ws, stream);
    if (rc) return rc;
    if (g.n_edges > 0) {
        const int blocks = (g.n_edges + 255) / 256;
        src_fill_kernel<<<blocks > 1184 ? 1184 : blocks, 256, 0, stream>>>(g.src, g.n_edges, sptr, cursor, slist);
        const int nb = (g.n_nodes + 255) / 256;
        src_sort_kernel<<<nb > 1184 ? 1184 : nb, 256, 0, stream>>>(sptr, g.n_nodes, slist);
    }
    RGNN_CHECK_CUDA(cudaGetLastError());
    *sptr_out = sptr;
    *slist_out = slist;
    return RGNN_OK;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool mp_bwd_tc_supported(const ConvDims& d) { return mp_tc_supported(d) && rgnn_get_option("tensor_cores_bwd") != 0; }

size_t mp_bwd_tc_scratch_floats(const ConvDims& d, int n_edges) {
    const size_t E = n_edges > 0 ? (size_t)n_edges : 1;
    return E * (2 * (size_t)d.h + d.cn);
}

int run_conv_edges_bwd_tc(const rgnn_conv& c, const ConvDims& d, const rgnn_graph& g, const float* emb, const float* P,
                          const float* dagg, float* dP, float* demb, bool first_demb, float* scratch, const int* sptr,
                          const int* slist, cudaStream_t stream) {
    const rgnn_linear& m0 = c.msg.layer[0];
    const rgnn_linear& m1 = c.msg.layer[1];
    const size_t E = (size_t)g.n_edges;
    float* y1 = scratch;
    float* dz1 = y1 + E * d.h;
    float* dz2 = dz1 + E * d.h;
    static PerDeviceOnce once;
    if (once.needed()) {
        RGNN_CHECK_CUDA(cudaFuncSetAttribute(mp_edge_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MB_SMEM));
        once.mark();
    }
    MpBwdArgs a;
    a.emb = emb; a.P = P; a.dagg = dagg; a.tgt = g.tgt; a.src = g.src;
    a.wpack = m0.weight_t + conv_msg0_tc_offset(d);
    a.s1 = m0.norm_scale; a.m1 = m0.norm_shift;
    a.b2 = m1.bias; a.s2 = m1.norm_scale; a.m2 = m1.norm_shift;
    a.y1_out = y1; a.dz1_out = dz1; a.dz2_out = dz2; a.demb = demb;
    a.g_s1 = m0.grad_norm_scale; a.g_m1 = m0.grad_norm_shift; a.g_s2 = m1.grad_norm_scale; a.g_m2 = m1.grad_norm_shift;
    a.n_edges = g.n_edges; a.act1 = m0.activation; a.act2 = m1.activation;
    a.demb_accumulate = first_demb ? 0 : 1;
    a.passes = rgnn_get_option("tf32_passes");
    const int n_tiles = (g.n_edges + MB_TM - 1) / MB_TM;
    const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
    a.prof = nullptr;
    if (g_rowmlp_profile) {     // developer aid: per-phase cycles of worker thread 0, printed to stderr (synchronises!)
        long long* prof = nullptr;
        RGNN_CHECK_CUDA(cudaMalloc(&prof, sizeof(long long) * 10 * grid));
        a.prof = prof;
        mp_edge_bwd_tc_kernel<<<grid, MB_NW + 128, MB_SMEM, stream>>>(a);
        RGNN_CHECK_CUDA(cudaStreamSynchronize(stream));
        std::vector<long long> h(10 * grid);
        RGNN_CHECK_CUDA(cudaMemcpy(h.data(), prof, sizeof(long long) * 10 * grid, cudaMemcpyDeviceToHost));
        double tot[10] = {0};
        for (int b = 0; b < grid; ++b) for (int i = 0; i < 10; ++i) tot[i] += (double)h[b * 10 + i];
        static const char* nm[10] = {"in->A", "preG1", "waitG1", "epi1", "waitG2", "epi2", "waitG3", "epi3", "waitG4", "epi4"};
        fprintf(stderr, "[mp_edge_bwd_tc profile] cycles per tile (thread 0):");
        for (int i = 0; i < 10; ++i) fprintf(stderr, " %s=%.0f", nm[i], tot[i] / (double)n_tiles);
        fprintf(stderr, "\n");
        cudaFree(prof);
    } else {
        mp_edge_bwd_tc_kernel<<<grid, MB_NW + 128, MB_SMEM, stream>>>(a);
    }
    RGNN_CHECK_CUDA(cudaGetLastError());
    int rc;
    // dW_2 (cn x h) = dz2^T y1:  D[m = input channel of msg.1][n = output channel] -> dst[n * h + m];  db_2 = column sums of dz2
    if ((rc = launch_wgrad_tc(y1, d.h, d.h, dz2, d.cn, d.cn, g.n_edges, m1.grad_weight, 1, m1.in_features, nullptr, m1.grad_bias, stream)))
        return rc;
    // dW_1[:, 2cn:] (h x ce) = dz1^T emb;  db_1 = column sums of dz1 (the bias rides on the target projection)
    if ((rc = launch_wgrad_tc(dz1, d.h, d.h, emb, d.ce, d.ce, g.n_edges, m0.grad_weight ? m0.grad_weight + 2 * d.cn : nullptr,
                              m0.in_features, 1, m0.grad_bias, nullptr, stream)))
        return rc;
    const int wpb = 8;
    const int blocks = (g.n_nodes + wpb - 1) / wpb;
    dproj_gather_kernel<<<blocks > 8 * sm_count() ? 8 * sm_count() : blocks, 32 * wpb, 0, stream>>>(dz1, g.row_ptr, sptr, slist, g.n_nodes, d.h, dP);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

}  // namespace rgnn
