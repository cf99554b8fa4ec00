// Message function + sum aggregation of one residual_graph_conv_block on the 5th-generation tensor cores.
//
//   agg[t] = sum_{e: s->t}  ffn2( ffn1( cat(x_t, x_s, emb_e) ) )          (reference gnn_blocks.py:106-113)
//
// with ffn1 = Linear(2Cn+Ce -> H) + channel_normalization + LeakyReLU and ffn2 = Linear(H -> Cn) + norm + act.
// The node part of ffn1's Linear is hoisted out of the edge loop (P = [x W_t^T | x W_s^T], one row per node,
// produced by the node kernel), so per edge
//   z1 = emb_e W_e^T + P_t[tgt] + P_s[src]                  (msg.0's bias rides on P_t)
// A CTA owns tiles of 128 target-major edges (UMMA M = 128, thread = TMEM lane = edge row):
//   GEMM1  D1[128 x H]  = A(emb tile, TMEM) * W_e^T (smem)          tcgen05.mma kind::tf32, TS
//   epi 1  z1 -> mean / unbiased std / scalar affine / LeakyReLU -> y1, written back to TMEM (hi and lo parts)
//   GEMM2  D2[128 x Cn] = y1 (TMEM) * W_2^T (smem)                  tcgen05.mma kind::tf32, TS
//   epi 2  + b2 -> norm -> act -> staged in smem -> segmented sum over equal consecutive targets -> agg
// Per-edge activations never leave the SM.  fp32 parity (rtol 1e-4 through seven residual layers) is kept by
// 3xTF32: every operand is split x = hi + lo (both exact in tf32) and D = A_lo B_hi + A_hi B_lo + A_hi B_hi
// is accumulated in fp32 by the tensor core (relative error ~2^-21 per product); `passes = 1` keeps only the
// hi*hi term (plain TF32, documented tolerance).
#include "rgnn_model.h"
#include "rgnn_pack.cuh"
#include "rgnn_tc.cuh"
#include "rgnn_tile.cuh"
#include "rgnn_tc_rows.cuh"

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
    int debug;              // experiment switches (0 in production)
    long long* prof;        // PROFILE builds: [grid][worker warps][12] cycle counters of each worker warp's lane 0
};

extern int g_rowmlp_profile;
static int g_tf32_passes = 3;
static int g_use_tensor_cores = 1;
static int g_use_tensor_cores_bwd = 1;
static int g_wgrad_tma = 1;
static int g_debug = 0;

template <int CE, int H, int CN, int NQ>
struct MpTcLayout {
    static constexpr int TM = 128;             // edges per tile = UMMA M
    static constexpr int NT = 128 * NQ;        // NQ threads share a row, each owns 1/NQ of its columns
    static constexpr int W1 = CE * H;          // floats per (hi or lo) copy
    static constexpr int W2 = H * CN;
    static constexpr int A = TM * CE;
    static constexpr int SEG = TM + 4;
    static constexpr int OFF_W1 = 0;
    static constexpr int OFF_W2 = OFF_W1 + 2 * W1;
    static constexpr int OFF_A = OFF_W2 + 2 * W2;           // A operand (hi | lo); between GEMM1 and the next fill: P_s row staging
    static constexpr int OFF_STAGE = OFF_A + 2 * A;         // messages of one tile, XOR-swizzled 16-byte chunks
    static constexpr int OFF_SEG = OFF_STAGE + TM * CN;     // [2][SEG] int2 {first row, target node} of every target segment (double buffered: deferred segsum)
    static constexpr int OFF_MASK = OFF_SEG + 4 * SEG;      // [2][4] segment-head ballots, [2] segment counts, 3 x [2] boundary flags
    static constexpr int OFF_BAR = OFF_MASK + 16;           // 2 x uint64
    static constexpr int OFF_SLOT = OFF_BAR + 4;
    static constexpr int FLOATS = OFF_SLOT + 2;
    static constexpr size_t BYTES = (size_t)FLOATS * 4;
    static constexpr int TMEM_COLS = 512;                   // D1/y1_hi [0,H) | y1_lo [H,2H) | D2 [2H, 2H+CN) | row statistics (16) | emb hi, lo [512-2CE, 512)
    static constexpr int COL_EHI = 512 - 2 * CE, COL_ELO = 512 - CE;
    static_assert(CE % 16 == 0 && H % 16 == 0 && CN % 16 == 0, "UMMA shape constraints");
    static_assert(2 * H + CN + 16 + 2 * CE <= 512, "TMEM budget");
    static_assert((CE / NQ) % 16 == 0, "emb column split must keep 16-column TMEM accesses");
    static_assert((H / NQ) % 16 == 0 && (CN / NQ) % 16 == 0, "column split must keep 16-column TMEM accesses");
    static_assert(NT % CN == 0 && CN % 4 == 0 && CN / 4 >= 8, "segmented sum thread mapping / stage swizzle");
    static_assert((OFF_BAR % 2) == 0, "mbarrier alignment");
    static_assert(TM * H * 4 <= 2 * A * 4, "P_s staging must fit in the A region");
    static_assert(NQ == 2 || NQ == 4, "row statistics exchange through TMEM");
    static_assert(BYTES <= 227 * 1024, "shared memory budget");
};

// Thread roles (128*NQ + 128 threads, one CTA per SM):
//   workers   128*NQ threads, thread = edge row x column part: emb fill, both epilogues
//   MMA warp  lane 0 issues every tcgen05.mma (the issue loop blocks while the tensor pipe is busy, so it must not sit
//             on a worker's critical path)
//   helpers   the three warps next to the MMA warp: P_s row staging of the next tile, segmented sum of the finished one
// Both A operands live in TMEM (TS MMAs).  An SS tf32 MMA with N = 128 reads 8 KB of shared memory per 64 cycles = the
// whole 128 B/clk of the SM, which starves every other shared-memory user for the duration of GEMM1; with the emb tile
// in TMEM the tensor core reads only W_e, the fill is a tcgen05.st (no proxy fence), and the staging region is no
// operand, so P_s rows can be copied at any time.
// Order of work (default, `debug & 256` = 0): GEMM1 of tile i+1 is issued right behind GEMM2 of tile i and runs under
// the workers' epilogue 2 of tile i:
//     workers   epi1(i)  fill(i+1)  gathers(i+1)  | wait G2(i) |  epi2(i)      epi1(i+1) ...
//     tensor             G2(i) ................................  G1(i+1) ....
// TMEM: D1/y1 of tile i [0, 2H) is dead once G2(i) has completed, which the in-order tensor pipe guarantees before
// G1(i+1) overwrites it; emb(i+1) [512-2CE, 512) is filled while G2(i) runs; D2(i) [2H, 2H+CN) is read while G1(i+1) runs.
// `debug & 256` keeps the serial order fill -> G1 -> epi1 -> G2 -> epi2 per tile (A/B measurements).
template <int CE, int H, int CN, int NQ, bool PROFILE>
__global__ void __launch_bounds__(128 * NQ + 128, 1) mp_edge_tc_kernel(const __grid_constant__ MpTcArgs a) {
    using L = MpTcLayout<CE, H, CN, NQ>;
    constexpr int TM = L::TM, NW = L::NT, NALL = L::NT + 128, HQ = H / NQ, CQ = CN / NQ, EQ = CE / NQ;
    constexpr int NMMA = NW + 32;            // threads that take part in the worker <-> MMA-warp barriers
    // register budget (setmaxnreg): the auxiliary warpgroup (MMA issue warp + three helper warps) keeps AUX_REGS,
    // the workers take the rest
    constexpr int WORKER_REGS = NQ == 4 ? 104 : 216, AUX_REGS = NQ == 4 ? 64 : 72;
    constexpr int LAUNCH_REGS = (65536 / NALL) / 8 * 8;     // what __launch_bounds__(NALL, 1) compiles to (168 / 96)
    static_assert(128 * AUX_REGS + NW * WORKER_REGS <= 65536, "register file");
    // setmaxnreg.inc only draws from what setmaxnreg.dec released (measured: asking for the never-allocated remainder
    // of the register file as well blocks forever)
    static_assert(128 * (LAUNCH_REGS - AUX_REGS) >= NW * (WORKER_REGS - LAUNCH_REGS), "setmaxnreg pool");
    // Hand-over between workers and helpers by four named barriers (workers + helpers):
    //   PS_READY   helpers -> workers : staged P_s rows of the next tile are in shared memory
    //   PS_FREE    workers -> helpers : they have been consumed (added into z), the region may be refilled
    //   STAGE_FULL workers -> helpers : messages + segment table of this tile are in shared memory
    //   STAGE_FREE helpers -> workers : the segmented sum has read them (stage and segment table may be rewritten)
    // (one helper warp stages, the other two sum: the two jobs must not wait for each other)
    constexpr int NSTAGER = 32, NSUMMER = 64, NWS = NW + NSTAGER, NWG = NW + NSUMMER;
    constexpr int BAR_PS_READY = 8, BAR_PS_FREE = 9, BAR_STAGE_FULL = 10, BAR_STAGE_FREE = 11;
    const bool pipelined = !(a.debug & 256);
    extern __shared__ __align__(1024) float smem[];
    float* w1s = smem + L::OFF_W1;
    float* w2s = smem + L::OFF_W2;
    float* Gs = smem + L::OFF_A;             // staged P_s rows of the next tile
    float* stage = smem + L::OFF_STAGE;
    int2* seg_s = reinterpret_cast<int2*>(smem + L::OFF_SEG);
    unsigned* mask_s = reinterpret_cast<unsigned*>(smem + L::OFF_MASK);
    int* nseg_s = reinterpret_cast<int*>(smem + L::OFF_MASK + 8);
    int* cut_s = nseg_s + 2;
    int* cut_first_s = nseg_s + 4;
    int* cut_last_s = nseg_s + 6;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem + L::OFF_SLOT);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = tid & 127, q = tid >> 7;
    const int bar_id = 1 + (row >> 5);       // the NQ warps that share rows [32k, 32k+32)
    const int G = (int)gridDim.x;

    // ---- one-time setup: weights -> smem (already hi/lo split and chunk-major), barriers, TMEM ----
    {
        const float4* g = reinterpret_cast<const float4*>(a.wpack);
        float4* s = reinterpret_cast<float4*>(w1s);
        constexpr int N4 = (2 * L::W1 + 2 * L::W2) / 4;
        for (int i = tid; i < N4; i += NALL) s[i] = __ldg(g + i);
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
    constexpr uint32_t COL_Y1LO = H, COL_D2 = 2 * H, COL_XS = 2 * H + CN;
    const int n_tiles = (a.n_edges + TM - 1) / TM;
    const int np = a.passes == 1 ? 1 : 3;
    if ((int)blockIdx.x >= n_tiles) goto teardown;

    if (tid >= NW + 32) {
        // =========================== helper warps ===========================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(AUX_REGS));
        static_assert(H == 128, "one warp instruction copies one H-float row");
        static_assert(CN == 64, "segsum thread mapping assumes 16 float4 per message row");
        const int ht = tid - (NW + 32);
        if (ht < NSTAGER) {
            // ---- staging warp: P_s[source] rows of the tile after next -> shared memory, whole 512-byte rows per warp
            // instruction (cp.async, 16 B per lane; sector-granular per-thread gathers run ~2.5x slower,
            // tools/micro/bench_gather.cu).  Row r is stored with its 16-byte chunks XOR-swizzled by (r & 7) so that the
            // row-owning threads read it conflict-free.  The copies come from DRAM (P does not fit the L2): issue + landing
            // take a few thousand cycles per tile, which is why this warp does nothing else.
            int id[TM / 32];                                     // source ids of rows lane, lane + 32, ...
            auto load_ids = [&](int tile) {
#pragma unroll
                for (int k = 0; k < TM / 32; ++k) {
                    const int e = tile * TM + k * 32 + lane;
                    id[k] = (tile < n_tiles && e < a.n_edges) ? __ldg(a.src + e) : -1;
                }
            };
            auto hstage = [&]() {
#pragma unroll
                for (int k = 0; k < TM / 32; ++k) {
#pragma unroll 8
                    for (int j = 0; j < 32; ++j) {
                        const int r = k * 32 + j;
                        const int sn = __shfl_sync(0xffffffffu, id[k], j);
                        if (sn >= 0) cp_async16(Gs + r * H + ((lane ^ (r & 7)) << 2), a.P + (size_t)sn * (2 * H) + H + 4 * lane);
                    }
                }
                cp_async_commit();
            };
            load_ids(blockIdx.x + G);
            group_sync(BAR_PS_FREE, NWS);                        // the workers' prologue is done with the staging region
            hstage();
            load_ids(blockIdx.x + 2 * G);
            cp_async_wait<0>();
            bar_arrive(BAR_PS_READY, NWS);
            for (int tile = blockIdx.x; tile < n_tiles; tile += G) {
                group_sync(BAR_PS_FREE, NWS);                    // rows of tile + G consumed -> stage those of tile + 2G
                hstage();
                load_ids(tile + 3 * G);
                cp_async_wait<0>();
                if (tile + G < n_tiles) bar_arrive(BAR_PS_READY, NWS);
            }
        } else {
            // ---- segmented-sum warps: the tile whose messages sit in `stage`.  16 threads (one float4 of columns each)
            // per target segment; interior segments are whole CSR rows by construction (edges are target-major), only the
            // first / last segment of a tile can be cut by its boundary.  Every shared-memory round trip costs several
            // hundred cycles while the tensor core streams its operands, so the dependent chain is kept to: {count,
            // bounds + target of my first three segments} -> {up to RB rows at once} -> store.
            constexpr int RB = 8, NGRP = NSUMMER / 16, SPEC = 3;
            const int hs = ht - NSTAGER, s0 = hs >> 4, c4 = hs & 15;
            bar_arrive(BAR_STAGE_FREE, NWG);
            int hbuf = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += G, hbuf ^= 1) {
                group_sync(BAR_STAGE_FULL, NWG);                 // messages of `tile` are staged
                const int2* sg = seg_s + hbuf * L::SEG;
                int2 h[SPEC + 1];                                // speculative: entries beyond the count are never used
#pragma unroll
                for (int k = 0; k < SPEC; ++k) h[k] = sg[min(s0 + k * NGRP, TM)];
                int2 hn[SPEC];
#pragma unroll
                for (int k = 0; k < SPEC; ++k) hn[k] = sg[min(s0 + k * NGRP + 1, TM)];
                const int nseg = nseg_s[hbuf];
                const int cut = cut_s[hbuf];       // bit 0: first segment continues from the previous tile; bit 1: last one continues
                int k = 0;
                for (int s = s0; s < nseg; s += NGRP, ++k) {
                    int2 a0, a1;
                    if (k == 0) { a0 = h[0]; a1 = hn[0]; }
                    else if (k == 1) { a0 = h[1]; a1 = hn[1]; }
                    else if (k == 2) { a0 = h[2]; a1 = hn[2]; }
                    else { a0 = sg[s]; a1 = sg[s + 1]; }
                    const int rs = a0.x, re = a1.x;
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int r0 = rs; r0 < re; r0 += RB) {
                        float4 v[RB];
#pragma unroll
                        for (int j = 0; j < RB; ++j) {
                            const int r = r0 + j;
                            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (r < re) v[j] = *reinterpret_cast<const float4*>(stage + r * CN + ((c4 ^ (r & 7)) << 2));
                        }
#pragma unroll
                        for (int j = 0; j < RB; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
                    }
                    // whole rows are stored (source-ascending order, as the reference's index_add_); a row cut by a tile
                    // boundary is completed with atomicAdd onto the zero-initialised output
                    const bool whole = !((s == 0 && (cut & 1)) || (s == nseg - 1 && (cut & 2)));
                    float* o = a.agg + (size_t)a0.y * CN + 4 * c4;
                    if (whole) {
                        *reinterpret_cast<float4*>(o) = acc;
                    } else {
                        atomicAdd(o, acc.x); atomicAdd(o + 1, acc.y); atomicAdd(o + 2, acc.z); atomicAdd(o + 3, acc.w);
                    }
                }
                __syncwarp();
                if (tile + G < n_tiles) bar_arrive(BAR_STAGE_FREE, NWG);
            }
        }
    } else if (tid >= NW) {
        // =========================== MMA issue warp ===========================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(AUX_REGS));
        constexpr uint32_t IDESC1 = tc::idesc_tf32(TM, H);
        constexpr uint32_t IDESC2 = tc::idesc_tf32(TM, CN);
        constexpr uint32_t LBO_W1 = H * 16, LBO_W2 = CN * 16, SBO = 128;
        const uint32_t sW1 = tc::smem_u32(w1s), sW2 = tc::smem_u32(w2s);
        // 3xTF32: small terms first (lo*hi, hi*lo), then hi*hi
        auto gemm1 = [&]() {   // D1 = emb (TMEM) * W1e^T
            group_sync(BAR_A_READY, NMMA);
            tc::tc_fence_after();
            if (lane == 0) {
                bool acc = false;
                for (int p = 0; p < np; ++p) {
                    const int pa = (np == 1) ? 0 : (p == 0 ? 1 : 0);
                    const int pb = (np == 1) ? 0 : (p == 1 ? 1 : 0);
                    const uint64_t bd0 = tc::smem_desc(sW1 + pb * (L::W1 * 4), LBO_W1, SBO);
                    const uint32_t ecol = tmem + (pa ? L::COL_ELO : L::COL_EHI);
#pragma unroll
                    for (int ks = 0; ks < CE / 8; ++ks) {
                        tc::mma_tf32_ts(tmem, ecol + ks * 8, bd0 + (uint64_t)((ks * 2 * LBO_W1) >> 4), IDESC1, acc);
                        acc = true;
                    }
                }
                tc::mma_commit(&bars[0]);
            }
            __syncwarp();
        };
        auto gemm2 = [&]() {   // D2 = y1 (TMEM) * W2^T
            group_sync(BAR_Y_READY, NMMA);
            tc::tc_fence_after();
            if (lane == 0) {
                bool acc = false;
                for (int p = 0; p < np; ++p) {
                    const int pa = (np == 1) ? 0 : (p == 0 ? 1 : 0);
                    const int pb = (np == 1) ? 0 : (p == 1 ? 1 : 0);
                    const uint32_t acol = tmem + (pa ? COL_Y1LO : 0);
                    const uint64_t bd0 = tc::smem_desc(sW2 + pb * (L::W2 * 4), LBO_W2, SBO);
#pragma unroll
                    for (int ks = 0; ks < H / 8; ++ks) {
                        tc::mma_tf32_ts(tmem + COL_D2, acol + ks * 8, bd0 + (uint64_t)((ks * 2 * LBO_W2) >> 4), IDESC2, acc);
                        acc = true;
                    }
                }
                tc::mma_commit(&bars[1]);
            }
            __syncwarp();
        };
        if (pipelined) {
            // PROFILE: the tensor pipe's view (lane 0): m0 = wait for y1 (Y_READY), m1 = G2 issue -> G2 complete,
            // m2 = G2 complete -> G1(next) complete (0 when it finished earlier), m3 = G2 issue -> G1(next) issued
            long long m[4] = {0, 0, 0, 0};
            uint32_t mph = 0;
            gemm1();
            for (int tile = blockIdx.x; tile < n_tiles; tile += G, mph ^= 1) {
                const bool nx = tile + G < n_tiles;
                long long t0 = 0, t1 = 0, t2 = 0;
                if (PROFILE) t0 = clock64();
                gemm2();
                if (PROFILE) t1 = clock64();
                if (nx) gemm1();
                if (PROFILE && lane == 0) {
                    t2 = clock64();
                    tc::mbar_wait(&bars[1], mph);
                    const long long t3 = clock64();
                    if (nx) tc::mbar_wait(&bars[0], mph ^ 1);
                    const long long t4 = clock64();
                    m[0] += t1 - t0; m[1] += t3 - t1; m[2] += t4 - t3; m[3] += t2 - t1;
                }
                __syncwarp();
            }
            if (PROFILE && lane == 0 && a.prof != nullptr)
                for (int i = 0; i < 4; ++i) a.prof[(blockIdx.x * (NW / 32)) * 12 + 8 + i] = m[i];
        } else {
            for (int tile = blockIdx.x; tile < n_tiles; tile += G) { gemm1(); gemm2(); }
        }
    } else {
        // =========================== worker warps ===========================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(WORKER_REGS));
        const uint32_t t_row = tmem + ((uint32_t)(row & ~31) << 16);   // this warp's 32 TMEM lanes

        // emb rows of a tile, prefetched into registers one tile ahead: thread = (edge row, column part), EQ contiguous
        // floats of its own row, 32 bytes per request
        float4 pre[EQ / 4];
        auto prefetch = [&](int tile) {
            const int e = tile * TM + row;
            const bool v = tile < n_tiles && e < a.n_edges;
            const float* p = a.emb + (size_t)(v ? e : 0) * CE + q * EQ;
#pragma unroll
            for (int c8 = 0; c8 < EQ / 8; ++c8) {
                float2 x0 = make_float2(0.f, 0.f), x1 = x0, x2 = x0, x3 = x0;
                if (v) ldg256(p + 8 * c8, x0, x1, x2, x3);
                pre[2 * c8] = make_float4(x0.x, x0.y, x1.x, x1.y);
                pre[2 * c8 + 1] = make_float4(x2.x, x2.y, x3.x, x3.y);
            }
        };
        // per-row indices of one tile, fetched one tile ahead:  t : target node of this thread's edge row;  first (lane 0
        // of the first four warps): target of the edge just before this warp's rows;  last (thread TM-1): target of the
        // edge after the tile
        auto load_idx = [&](int tile, int& t, int& first, int& last) {
            const int e = tile * TM + row;
            const bool v = tile < n_tiles && e < a.n_edges;
            t = v ? __ldg(a.tgt + e) : -1;
            first = (tid < TM && lane == 0 && v && e > 0) ? __ldg(a.tgt + e - 1) : -2;
            last = (tid == TM - 1 && v && e + 1 < a.n_edges) ? __ldg(a.tgt + e + 1) : -3;
        };
        // z = P_t[target] + P_s[source] for this thread's part of the row (the hoisted node part of msg.0; the Linear
        // bias is already inside P_t).  P_t rows repeat over consecutive edges: few distinct sectors per warp.
        auto load_pt = [&](float2 (&z)[HQ / 2], int t) {
            const bool v = t >= 0;
            const float* Pt = a.P + (size_t)(v ? t : 0) * (2 * H) + q * HQ;
#pragma unroll
            for (int c8 = 0; c8 < HQ / 8; ++c8) {
                z[4 * c8] = z[4 * c8 + 1] = z[4 * c8 + 2] = z[4 * c8 + 3] = make_float2(0.f, 0.f);
                if (v) ldg256(Pt + 8 * c8, z[4 * c8], z[4 * c8 + 1], z[4 * c8 + 2], z[4 * c8 + 3]);
            }
        };
        auto add_ps = [&](float2 (&z)[HQ / 2]) {
#pragma unroll
            for (int c4 = 0; c4 < HQ / 4; ++c4) {
                const float4 s0 = *reinterpret_cast<const float4*>(Gs + row * H + (((q * (HQ / 4) + c4) ^ (row & 7)) << 2));
                z[2 * c4] = __fadd2_rn(z[2 * c4], make_float2(s0.x, s0.y));
                z[2 * c4 + 1] = __fadd2_rn(z[2 * c4 + 1], make_float2(s0.z, s0.w));
            }
        };

        long long pt[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tlast = 0;
        auto tick = [&](int i) {
            if (PROFILE && lane == 0) { const long long now = clock64(); pt[i] += now - tlast; tlast = now; }
        };
        // per-layer constants once, in registers (no global load may sit in the epilogues' dependent chains)
        const bool norm1 = a.s1 != nullptr, norm2 = a.s2 != nullptr;
        const float s1v = norm1 ? __ldg(a.s1) : 1.f, m1v = norm1 ? __ldg(a.m1) : 0.f;
        const float s2v = norm2 ? __ldg(a.s2) : 1.f, m2v = norm2 ? __ldg(a.m2) : 0.f;
        float2 b2v[CQ / 2];
#pragma unroll
        for (int c = 0; c < CQ / 2; ++c)
            b2v[c] = a.b2 ? __ldg(reinterpret_cast<const float2*>(a.b2 + q * CQ) + c) : make_float2(0.f, 0.f);
        float2 z[HQ / 2];

        // ---- fill: emb tile (hi | lo split) from the prefetched registers -> TMEM; segment-head ballots of the tile ----
        auto fill = [&](int fbuf, int nvalid, int t, int first, int last) {
#pragma unroll
            for (int c = 0; c < EQ; c += 16) {
                float2 hi[8], lo[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 v = pre[c / 4 + j];
                    tc::split_tf32(make_float2(v.x, v.y), hi[2 * j], lo[2 * j]);
                    tc::split_tf32(make_float2(v.z, v.w), hi[2 * j + 1], lo[2 * j + 1]);
                }
                tc::tmem_st16(t_row + L::COL_EHI + q * EQ + c, hi);
                if (np != 1) tc::tmem_st16(t_row + L::COL_ELO + q * EQ + c, lo);
            }
            if (tid < TM) {
                const bool v = tid < nvalid;
                int prev = __shfl_up_sync(0xffffffffu, t, 1);
                if (lane == 0) prev = first;
                const unsigned m = __ballot_sync(0xffffffffu, v && (tid == 0 || t != prev));
                if (lane == 0) mask_s[fbuf * 4 + warp] = m;
                if (tid == 0) cut_first_s[fbuf] = (first == t) ? 1 : 0;
                if (tid == nvalid - 1) cut_last_s[fbuf] = (tid == TM - 1 && last == t) ? 2 : 0;
            }
            tc::tmem_wait_st();
            tc::tc_fence_before();
            bar_arrive(BAR_A_READY, NMMA);          // -> MMA warp issues GEMM1
        };
        // ---- segment table {first row, target node} of a tile from its ballots (after a worker-wide barrier) ----
        auto build_seg = [&](int sbuf, int nvalid, int t) {
            if (tid < TM) {
                const unsigned* mk = mask_s + sbuf * 4;
                const unsigned m = mk[warp];
                int base = 0;
                for (int w = 0; w < warp; ++w) base += __popc(mk[w]);
                if ((m >> lane) & 1u) seg_s[sbuf * L::SEG + base + __popc(m & ((1u << lane) - 1u))] = make_int2(tid, t);
                if (tid == TM - 1) {
                    const int n = base + __popc(m);
                    seg_s[sbuf * L::SEG + n] = make_int2(nvalid, -1);
                    nseg_s[sbuf] = n;
                    cut_s[sbuf] = cut_first_s[sbuf] | cut_last_s[sbuf];
                }
            }
        };
        // ---- epilogue 1: z1 = D1 + (P_t + P_s) -> norm -> act -> y1 (hi | lo) back into TMEM ----
        auto epi1 = [&](uint32_t phase, int tk) {
            tc::mbar_wait(&bars[0], phase);
            tc::tc_fence_after();
            tick(tk);
            {
                float2 d[HQ / 2];
#pragma unroll
                for (int c = 0; c < HQ; c += 16) tc::tmem_ld16(t_row + q * HQ + c, d + c / 2);
                tc::tmem_wait_ld();
#pragma unroll
                for (int c = 0; c < HQ / 2; ++c) z[c] = __fadd2_rn(z[c], d[c]);
            }
            row_norm_act<HQ / 2, NQ>(z, H, norm1, s1v, m1v, a.act1 != 0, t_row + COL_XS, q, bar_id);
            if (np == 1) {
#pragma unroll
                for (int c = 0; c < HQ; c += 16) tc::tmem_st16(t_row + q * HQ + c, z + c / 2);
            } else {
#pragma unroll
                for (int c = 0; c < HQ; c += 16) {
                    float2 hi[8], lo[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) tc::split_tf32(z[c / 2 + j], hi[j], lo[j]);
                    tc::tmem_st16(t_row + q * HQ + c, hi);
                    tc::tmem_st16(t_row + COL_Y1LO + q * HQ + c, lo);
                }
            }
            tc::tmem_wait_st();
            tc::tc_fence_before();
            bar_arrive(BAR_Y_READY, NMMA);          // -> MMA warp issues GEMM2
            tick(tk + 1);
        };
        // ---- epilogue 2: message = act(norm(D2 + b2)) -> stage -> helpers' segmented sum ----
        auto epi2 = [&](uint32_t phase, int tk, bool seg_next, int sbuf, int nv_next, int t_next) {
            float2 m[CQ / 2];
            tc::mbar_wait(&bars[1], phase);
            tc::tc_fence_after();
            tick(tk);
            {
                float2 d[CQ / 2];
#pragma unroll
                for (int c = 0; c < CQ; c += 16) tc::tmem_ld16(t_row + COL_D2 + q * CQ + c, d + c / 2);
                tc::tmem_wait_ld();
#pragma unroll
                for (int c = 0; c < CQ / 2; ++c) m[c] = __fadd2_rn(b2v[c], d[c]);
            }
            row_norm_act<CQ / 2, NQ>(m, CN, norm2, s2v, m2v, a.act2 != 0, t_row + COL_XS + 2 * NQ, q, bar_id);
            tick(tk + 1);                           // PROFILE: p6 = epilogue 2 arithmetic, p7 = wait for STAGE_FREE
            group_sync(BAR_STAGE_FREE, NWG);        // the previous tile's segmented sum has read the stage and its segment table
            tick(7);
            if (seg_next) build_seg(sbuf, nv_next, t_next);
#pragma unroll
            for (int c4 = 0; c4 < CQ / 4; ++c4) {
                const int chunk = (q * (CQ / 4) + c4) ^ (row & 7);
                *reinterpret_cast<float4*>(stage + row * CN + chunk * 4) = make_float4(m[2 * c4].x, m[2 * c4].y, m[2 * c4 + 1].x, m[2 * c4 + 1].y);
            }
            tc::tc_fence_before();                  // D2 has been read: the next GEMM2 may overwrite it
            bar_arrive(BAR_STAGE_FULL, NWG);
            tick(tk + 1);
        };

        // ---- prologue: the first tile's inputs (the workers stage its P_s rows themselves) ----
        int t_my, e_first, e_last;
        {
            constexpr int RPW = TM / (NW / 32);     // P_s rows staged per warp
            const int es = blockIdx.x * TM + warp * RPW + (lane % RPW);
            const int s_my = es < a.n_edges ? __ldg(a.src + es) : -1;
            prefetch(blockIdx.x);
            load_idx(blockIdx.x, t_my, e_first, e_last);
#pragma unroll
            for (int i = 0; i < RPW; ++i) {
                const int r = warp * RPW + i;
                const int sn = __shfl_sync(0xffffffffu, s_my, i);
                if (sn >= 0) cp_async16(Gs + r * H + ((lane ^ (r & 7)) << 2), a.P + (size_t)sn * (2 * H) + H + 4 * lane);
            }
            cp_async_commit();
            load_pt(z, t_my);
            cp_async_wait<0>();
            group_sync(BAR_WORKERS, NW);
            add_ps(z);
        }
        uint32_t phase = 0;
        int buf = 0;
        if (PROFILE) tlast = clock64();
        if (pipelined) {
            const int nv0 = min(TM, a.n_edges - (int)blockIdx.x * TM);
            fill(0, nv0, t_my, e_first, e_last);                 // -> G1 of the first tile
            group_sync(BAR_WORKERS, NW);                         // its ballots are visible
            build_seg(0, nv0, t_my);
            prefetch(blockIdx.x + G);
            load_idx(blockIdx.x + G, t_my, e_first, e_last);     // from here on the index registers describe the NEXT tile
            bar_arrive(BAR_PS_FREE, NWS);
            for (int tile = blockIdx.x; tile < n_tiles; tile += G, phase ^= 1, buf ^= 1) {
                const int next = tile + G;
                const bool has_next = next < n_tiles;
                const int nv_next = min(TM, a.n_edges - next * TM);
                epi1(phase, 0);                                  // ticks 0 (wait G1), 1 (epilogue 1)
                if (has_next) {
                    load_pt(z, t_my);                                // gathers first: their round trip runs under the fill
                    fill(buf ^ 1, nv_next, t_my, e_first, e_last);   // -> G1(next), queued behind G2(tile)
                    prefetch(next + G);
                }
                tick(2);
                group_sync(BAR_PS_READY, NWS);                   // staged P_s rows of `next` (and its ballots) are visible
                tick(3);
                if (has_next) add_ps(z);
                bar_arrive(BAR_PS_FREE, NWS);
                tick(4);
                epi2(phase, 5, has_next, buf ^ 1, nv_next, t_my);    // ticks 5 (wait G2), 6 (epilogue 2 + next segment table)
                load_idx(next + G, t_my, e_first, e_last);
            }
        } else {
            bar_arrive(BAR_PS_FREE, NWS);
            for (int tile = blockIdx.x; tile < n_tiles; tile += G, phase ^= 1, buf ^= 1) {
                const int next = tile + G;
                const bool has_next = next < n_tiles;
                const int nvalid = min(TM, a.n_edges - tile * TM);
                fill(buf, nvalid, t_my, e_first, e_last);
                int n_t, n_first, n_last;
                prefetch(next);
                load_idx(next, n_t, n_first, n_last);
                tick(2);
                epi1(phase, 0);
                if (has_next) load_pt(z, n_t);                   // issue first: their latency overlaps the barrier below
                group_sync(BAR_PS_READY, NWS);                   // the helpers' staged rows (and this tile's ballots) are visible
                tick(3);
                build_seg(buf, nvalid, t_my);
                if (has_next) add_ps(z);
                bar_arrive(BAR_PS_FREE, NWS);
                t_my = n_t; e_first = n_first; e_last = n_last;
                tick(4);
                epi2(phase, 5, false, 0, 0, 0);
            }
        }
        if (PROFILE && lane == 0 && a.prof != nullptr)
            for (int i = 0; i < 8; ++i) a.prof[(blockIdx.x * (NW / 32) + warp) * 12 + i] = pt[i];   // 8..11 of warp 0: MMA warp
    }

teardown:
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, L::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// weight packing: element (n, k) = W[off + n * sn + k * sk] -> hi/lo chunk-major operands [K/4][N][4]
// ---------------------------------------------------------------------------------------------
__global__ void pack_split_kernel(const PackSplitArgs a) { pack_split_body(a, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x); }

static int launch_pack_split(const PackSplitArgs& a, cudaStream_t stream) {
    if (packq_push(a)) return RGNN_OK;          // inside rgnn_pack_detector: batched (rgnn_pack.cuh)
    pack_split_kernel<<<16, 256, 0, stream>>>(a);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

int g_use_tensor_cores_flag() { return g_use_tensor_cores; }

bool mp_tc_supported(const ConvDims& d) { return g_use_tensor_cores && d.cn == 64 && d.ce == 64 && d.h == 128; }

size_t mp_tc_pack_floats(const ConvDims& d) {
    // the buffer exists whenever the shape is one the tensor-core kernels are instantiated for: the two forward images
    // (W_e, W_2) followed by their transposes for the backward (W_2^T, W_e^T), each hi | lo
    return (d.cn == 64 && d.ce == 64 && d.h == 128) ? 2 * ((size_t)2 * d.ce * d.h + (size_t)2 * d.h * d.cn) : 0;
}

int mp_tc_pack(const rgnn_conv& c, const ConvDims& d, float* dst, cudaStream_t stream) {
    if (mp_tc_pack_floats(d) == 0) return RGNN_OK;
    const rgnn_linear& m0 = c.msg.layer[0];
    const rgnn_linear& m1 = c.msg.layer[1];
    const int W1 = d.ce * d.h, W2 = d.h * d.cn;
    float* w1 = dst;                 // (n = h, k = c)  = msg.0.weight[h][2cn + c]
    float* w2 = w1 + 2 * W1;         // (n = i, k = h)  = msg.1.weight[i][h]
    float* w2t = w2 + 2 * W2;        // (n = h, k = i)  = msg.1.weight[i][h]         (d y1 = dz2 W_2)
    float* w1t = w2t + 2 * W2;       // (n = c, k = h)  = msg.0.weight[h][2cn + c]   (d emb = dz1 W_e)
    int rcp = launch_pack_split({m0.weight, w1, w1 + W1, 2 * d.cn, m0.in_features, 1, d.ce, d.h}, stream);
    if (rcp == RGNN_OK) rcp = launch_pack_split({m1.weight, w2, w2 + W2, 0, m1.in_features, 1, d.h, d.cn}, stream);
    if (rcp == RGNN_OK) rcp = launch_pack_split({m1.weight, w2t, w2t + W2, 0, 1, m1.in_features, d.cn, d.h}, stream);
    if (rcp == RGNN_OK) rcp = launch_pack_split({m0.weight, w1t, w1t + W1, 2 * d.cn, 1, m0.in_features, d.h, d.ce}, stream);
    if (rcp) return rcp;
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

template <int NQ>
static int launch_mp_tc(MpTcArgs& a, int n_edges, cudaStream_t stream) {
    using L = MpTcLayout<64, 128, 64, NQ>;
    static PerDeviceOnce once;
    if (once.needed()) {
        RGNN_CHECK_CUDA(cudaFuncSetAttribute(mp_edge_tc_kernel<64, 128, 64, NQ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::BYTES));
        RGNN_CHECK_CUDA(cudaFuncSetAttribute(mp_edge_tc_kernel<64, 128, 64, NQ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::BYTES));
        once.mark();
    }
    const int n_tiles = (n_edges + L::TM - 1) / L::TM;
    const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
    if (g_debug & 8) {   // developer aid: per-phase cycle counters of thread 0, printed to stderr (synchronises!)
        long long* prof = nullptr;
        constexpr int WW = L::NT / 32;
        RGNN_CHECK_CUDA(cudaMalloc(&prof, sizeof(long long) * 12 * grid * WW));
        a.prof = prof;
        mp_edge_tc_kernel<64, 128, 64, NQ, true><<<grid, L::NT + 128, L::BYTES, stream>>>(a);
        RGNN_CHECK_CUDA(cudaStreamSynchronize(stream));
        long long* h = new long long[12 * grid * WW];
        RGNN_CHECK_CUDA(cudaMemcpy(h, prof, sizeof(long long) * 12 * grid * WW, cudaMemcpyDeviceToHost));
        for (int w = 0; w < WW; ++w) {
            double tot[12] = {0};
            for (int b = 0; b < grid; ++b) for (int i = 0; i < 12; ++i) tot[i] += (double)h[(b * WW + w) * 12 + i];
            fprintf(stderr, "[mp_edge_tc profile NQ=%d] cycles per tile, worker warp %d:", NQ, w);
            for (int i = 0; i < (w == 0 ? 12 : 7); ++i) fprintf(stderr, " p%d=%.0f", i, tot[i] / (double)n_tiles);
            fprintf(stderr, "\n");
        }
        delete[] h;
        cudaFree(prof);
        return RGNN_OK;
    }
    mp_edge_tc_kernel<64, 128, 64, NQ, false><<<grid, L::NT + 128, L::BYTES, stream>>>(a);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

int run_conv_edges_tc(const rgnn_conv& c, const ConvDims& d, const rgnn_graph& g, const float* emb, const float* P,
                      const float* wpack, float* agg, cudaStream_t stream) {
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
    a.debug = g_debug;
    a.prof = nullptr;
    (void)d;
    return (g_debug & 16) ? launch_mp_tc<4>(a, g.n_edges, stream) : launch_mp_tc<2>(a, g.n_edges, stream);
}

}  // namespace rgnn

extern "C" int rgnn_set_option(const char* name, int value) {
    using namespace rgnn;
    if (name != nullptr && strcmp(name, "tf32_passes") == 0 && (value == 1 || value == 3)) { g_tf32_passes = value; return RGNN_OK; }
    if (name != nullptr && strcmp(name, "tensor_cores") == 0 && (value == 0 || value == 1)) { g_use_tensor_cores = value; return RGNN_OK; }
    if (name != nullptr && strcmp(name, "tensor_cores_bwd") == 0 && (value == 0 || value == 1)) { g_use_tensor_cores_bwd = value; return RGNN_OK; }
    if (name != nullptr && strcmp(name, "wgrad_tma") == 0 && (value == 0 || value == 1)) { g_wgrad_tma = value; return RGNN_OK; }
    if (name != nullptr && strcmp(name, "pack_batch") == 0 && (value == 0 || value == 1)) { g_pack_batch = value; return RGNN_OK; }
    if (name != nullptr && mp_f16_set_option(name, value)) return RGNN_OK;
    if (name != nullptr && mp_bwd_f16_set_option(name, value)) return RGNN_OK;
    if (name != nullptr && graph_set_option(name, value)) return RGNN_OK;
    if (name != nullptr && node_bwd_f16_set_option(name, value)) return RGNN_OK;
    if (name != nullptr && chain_f16_set_option(name, value)) return RGNN_OK;
    if (name != nullptr && edge_enc_f16_set_option(name, value)) return RGNN_OK;
    if (name != nullptr && strcmp(name, "debug") == 0) { g_debug = value; g_rowmlp_profile = (value & 8) != 0; return RGNN_OK; }
    set_error("rgnn_set_option: unknown option or value (%s = %d)", name ? name : "(null)", value);
    return RGNN_ERR_INVALID;
}

extern "C" int rgnn_get_option(const char* name) {
    using namespace rgnn;
    if (name != nullptr && strcmp(name, "tf32_passes") == 0) return g_tf32_passes;
    if (name != nullptr && strcmp(name, "tensor_cores") == 0) return g_use_tensor_cores;
    if (name != nullptr && strcmp(name, "tensor_cores_bwd") == 0) return g_use_tensor_cores_bwd;
    if (name != nullptr && strcmp(name, "wgrad_tma") == 0) return g_wgrad_tma;
    if (name != nullptr && strcmp(name, "pack_batch") == 0) return g_pack_batch;
    if (name != nullptr && mp_f16_get_option(name) != -2) return mp_f16_get_option(name);
    if (name != nullptr && mp_bwd_f16_get_option(name) != -2) return mp_bwd_f16_get_option(name);
    if (name != nullptr && graph_get_option(name) != -2) return graph_get_option(name);
    if (name != nullptr && node_bwd_f16_get_option(name) != -2) return node_bwd_f16_get_option(name);
    if (name != nullptr && chain_f16_get_option(name) != -2) return chain_f16_get_option(name);
    if (name != nullptr && edge_enc_f16_get_option(name) != -2) return edge_enc_f16_get_option(name);
    return -1;
}
