// Weight gradients on the tensor cores:  D[m][n] = sum_r A[r][m] * B[r][n]   (dW = dZ^T X of an nn.Linear; replaces
// the torch autograd of common.py:185-205 inside gnn/training.py:81).
//
// A (rows x wa, wa <= 128) and B (rows x wb, wb <= 256) are row-major fp32 matrices in HBM with the reduction index
// (graph edges / nodes) as their row.  A CTA streams chunks of 32 rows: 256 loader threads fetch them one chunk ahead
// into registers, split every value into its tf32 hi / lo parts and store the chunk as blocks of 32 columns in which
// every row keeps its 128 contiguous bytes (32-byte granules XORed with row mod 4); with the reduction index along the
// rows that image is the MN-major SWIZZLE_128B_BASE32B operand of tcgen05.mma (the one MN-major layout 32-bit operands
// have; SBO = next group of 4 rows, LBO = next block of 32 columns), so neither operand is ever transposed -- the
// HBM rows go to shared memory as they are.  One elected thread issues
// tcgen05.mma.kind::tf32 (M = 128, N = wb, 3xTF32) into a TMEM accumulator that lives for the whole kernel (split-K
// over the CTAs); at the end the 128 x wb block is added to the gradient buffer with red.global.add.
// Column sums of either operand (bias gradients) fall out of the loader's registers.
#include "rgnn_model.h"
#include "rgnn_tc.cuh"
#include "rgnn_tc_rows.cuh"

namespace rgnn {

constexpr int WG_R = 32;                      // rows (reduction steps) per chunk
constexpr int WG_STAGES = 2;
constexpr int WG_MA = 128;                    // UMMA M (A is zero padded to it)
constexpr int WG_NB_MAX = 256;
constexpr int WG_LOADERS = 256;
constexpr int WG_NT = WG_LOADERS + 32;        // + the MMA warp
constexpr int WG_BLK = WG_R * 32;             // floats of one (32 rows x 32 columns) block image
static_assert(WG_R / 4 == WG_LOADERS / 32, "one warp per group of 4 rows");
// The kernel is instantiated for B operands of up to NBB blocks of 32 columns: NBB = 8 (wb <= 256, 192 KB of shared
// memory, one CTA per SM) and NBB = 2 (wb <= 64, 96 KB, 128 TMEM columns: two CTAs per SM, whose load latencies overlap).
template <int NBB>
struct WgLayout {
    static constexpr int NBLK = WG_MA / 32 + NBB;                 // 32-column blocks per stage: 4 of A, NBB of B
    static constexpr int STAGE_FLOATS = 2 * NBLK * WG_BLK;        // hi + lo images
    static constexpr int OFF_BAR = WG_STAGES * STAGE_FLOATS;
    static constexpr size_t SMEM = (size_t)(OFF_BAR + 16) * 4;
    static constexpr int TMEM_COLS = NBB <= 2 ? 64 : 256;
    static constexpr int CTAS = NBB <= 2 ? 2 : 1;
    static_assert(SMEM * CTAS <= 227 * 1024, "shared memory budget");
};

struct WgradArgs {
    const float* A; const float* B;
    int lda, ldb;
    int wa, wb;             // true widths
    int wa_pad, wb_pad;     // wa_pad = 128 (UMMA M); wb_pad multiple of 16 (UMMA N)
    long long rows;
    float* dst;             // dst[m * sm + n * sn] += D[m][n]
    long long sm, sn;
    float* colsum_a;        // optional: += sum_r A[r][m]
    float* colsum_b;        // optional: += sum_r B[r][n]
    const int* b_ridx;      // optional: row r of B is B[b_ridx[r]]
    int passes;
};

// shared-memory matrix descriptor, MN-major, SWIZZLE_128B_BASE32B (layout type 1): the only MN-major layout of 32-bit
// operands.  Canonical form ((4,8,m),(4,k)) : ((1,4,LBO),(32,SBO)) in elements: a reduction step (K row) holds 32
// consecutive MN elements (128 bytes) whose 32-byte granules are XORed with (K row mod 4); LBO = next block of 32 MN
// elements, SBO = next group of 4 K rows.
__device__ __forceinline__ uint64_t smem_desc_mn32(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return tc::smem_desc(saddr, lbo_bytes, sbo_bytes) | ((uint64_t)1 << 61);
}
// instruction descriptor: kind::tf32, fp32 accumulate, A and B MN-major (bits 15 / 16)
__host__ __device__ constexpr uint32_t idesc_tf32_mn(int M, int N) { return tc::idesc_tf32(M, N) | (1u << 15) | (1u << 16); }

template <int NBB>
__global__ void __launch_bounds__(WG_NT, WgLayout<NBB>::CTAS) wgrad_tc_kernel(const __grid_constant__ WgradArgs a) {
    using L = WgLayout<NBB>;
    constexpr int WG_NBLK = L::NBLK, WG_STAGE_FLOATS = L::STAGE_FLOATS;
    extern __shared__ __align__(1024) float smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* full = bars;                   // [WG_STAGES] loaders -> MMA
    uint64_t* empty = bars + WG_STAGES;      // [WG_STAGES] MMA (tcgen05.commit) -> loaders
    uint64_t* done = bars + 2 * WG_STAGES;
    uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2 * WG_STAGES + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < WG_STAGES; ++s) { tc::mbar_init(&full[s], WG_LOADERS); tc::mbar_init(&empty[s], 1); }
        tc::mbar_init(done, 1);
        tc::mbar_init_fence();
    }
    if (warp == 0) tc::tmem_alloc(slot, L::TMEM_COLS);
    // zero the operand images once: padded columns are never written again
    for (int i = tid; i < WG_STAGES * WG_STAGE_FLOATS / 4; i += WG_NT) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *slot;

    const long long n_chunks = (a.rows + WG_R - 1) / WG_R;
    const long long c_begin = n_chunks * blockIdx.x / gridDim.x, c_end = n_chunks * (blockIdx.x + 1) / gridDim.x;
    const int n_local = (int)(c_end - c_begin);
    const int np = a.passes == 1 ? 1 : 3;
    // stage layout (floats): hi image = [12 blocks][32 rows][32 floats] (blocks 0..3: A, 4..11: B) | lo image
    constexpr int OFF_LO = WG_NBLK * WG_BLK, OFF_B = (WG_MA / 32) * WG_BLK;

    if (warp == WG_LOADERS / 32) {
        // =========================== MMA issue warp ===========================
        if (lane == 0 && n_local > 0) {
            const uint32_t idesc = idesc_tf32_mn(WG_MA, a.wb_pad);
            constexpr uint32_t LBO = WG_BLK * 4, SBO = 4 * 128, KSTEP = 8 * 128;
            bool acc = false;
            for (int c = 0; c < n_local; ++c) {
                const int s = c % WG_STAGES;
                tc::mbar_wait(&full[s], (uint32_t)(c / WG_STAGES) & 1u);
                tc::tc_fence_after();
                const uint32_t base = tc::smem_u32(smem + s * WG_STAGE_FLOATS);
                for (int p = 0; p < np; ++p) {     // 3xTF32, small terms first: lo*hi, hi*lo, hi*hi
                    const int pa = (np == 1) ? 0 : (p == 0 ? 1 : 0), pb = (np == 1) ? 0 : (p == 1 ? 1 : 0);
                    const uint64_t ad0 = smem_desc_mn32(base + (pa ? OFF_LO : 0) * 4, LBO, SBO);
                    const uint64_t bd0 = smem_desc_mn32(base + ((pb ? OFF_LO : 0) + OFF_B) * 4, LBO, SBO);
#pragma unroll
                    for (int ks = 0; ks < WG_R / 8; ++ks) {
                        tc::mma_tf32_ss(tmem, ad0 + (uint64_t)((ks * KSTEP) >> 4), bd0 + (uint64_t)((ks * KSTEP) >> 4), idesc, acc);
                        acc = true;
                    }
                }
                tc::mma_commit(&empty[s]);
            }
            tc::mma_commit(done);
        }
        __syncwarp();
    } else {
        // =========================== loader / splitter warps ===========================
        // warp w owns rows 4w..4w+3 of every chunk; a warp instruction covers those 4 rows x one block of 32 columns
        // (128 contiguous bytes per row in HBM and in the image)
        const int rr = warp * 4 + (lane >> 3), c4 = lane & 7;       // row inside the chunk, float4 inside the 128-byte row
        const int nba = WG_MA / 32, nbb = (a.wb + 31) >> 5;
        const bool va = ((a.lda | a.wa) & 3) == 0 && ((uintptr_t)a.A & 15) == 0;
        const bool vb = ((a.ldb | a.wb) & 3) == 0 && ((uintptr_t)a.B & 15) == 0;
        // position of this thread's float4 inside the swizzled row: 32-byte granule (c4 >> 1) ^ (row & 3)
        const int sw = (rr * 32) + ((((c4 >> 1) ^ (rr & 3)) << 1 | (c4 & 1)) << 2);
        float4 pre[WG_NBLK], csum[WG_NBLK];
#pragma unroll
        for (int i = 0; i < WG_NBLK; ++i) csum[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        auto fetch = [&](long long chunk) {
            const long long r = chunk * WG_R + rr;
#pragma unroll
            for (int b = 0; b < WG_NBLK; ++b) {
                pre[b] = make_float4(0.f, 0.f, 0.f, 0.f);
                const bool isa = b < nba;
                if ((!isa && b - nba >= nbb) || r >= a.rows) continue;
                const int col = (isa ? b : b - nba) * 32 + c4 * 4;
                const int w = isa ? a.wa : a.wb;
                if (col >= w) continue;
                const long long rb = (!isa && a.b_ridx != nullptr) ? (long long)__ldg(a.b_ridx + r) : r;
                const float* p = (isa ? a.A + r * a.lda : a.B + rb * a.ldb) + col;
                if (isa ? va : vb) {
                    pre[b] = __ldg(reinterpret_cast<const float4*>(p));
                } else {
                    pre[b].x = __ldg(p);
                    if (col + 1 < w) pre[b].y = __ldg(p + 1);
                    if (col + 2 < w) pre[b].z = __ldg(p + 2);
                    if (col + 3 < w) pre[b].w = __ldg(p + 3);
                }
            }
        };
        if (n_local > 0) fetch(c_begin);
        for (int c = 0; c < n_local; ++c) {
            const int s = c % WG_STAGES;
            if (c >= WG_STAGES) tc::mbar_wait(&empty[s], (uint32_t)(c / WG_STAGES - 1) & 1u);
            float* st = smem + s * WG_STAGE_FLOATS;
#pragma unroll
            for (int b = 0; b < WG_NBLK; ++b) {
                if (b >= nba && b - nba >= nbb) continue;
                float4 hi, lo;
                tc::split_tf32(pre[b].x, hi.x, lo.x);
                tc::split_tf32(pre[b].y, hi.y, lo.y);
                tc::split_tf32(pre[b].z, hi.z, lo.z);
                tc::split_tf32(pre[b].w, hi.w, lo.w);
                *reinterpret_cast<float4*>(st + b * WG_BLK + sw) = hi;
                *reinterpret_cast<float4*>(st + OFF_LO + b * WG_BLK + sw) = lo;
                csum[b].x += pre[b].x; csum[b].y += pre[b].y; csum[b].z += pre[b].z; csum[b].w += pre[b].w;
            }
            tc::fence_async_smem();
            tc::mbar_arrive(&full[s]);
            if (c + 1 < n_local) fetch(c_begin + c + 1);
        }
        // column sums (bias gradients): reduce over the 4 rows of the warp, then one atomic per warp and column
        if (n_local > 0 && (a.colsum_a != nullptr || a.colsum_b != nullptr)) {
#pragma unroll
            for (int b = 0; b < WG_NBLK; ++b) {
                const bool isa = b < nba;
                float* o = isa ? a.colsum_a : a.colsum_b;
                if (o == nullptr || (!isa && b - nba >= nbb)) continue;      // warp-uniform
                float4 v = csum[b];
#pragma unroll
                for (int off = 8; off < 32; off <<= 1) {
                    v.x += __shfl_xor_sync(0xffffffffu, v.x, off);
                    v.y += __shfl_xor_sync(0xffffffffu, v.y, off);
                    v.z += __shfl_xor_sync(0xffffffffu, v.z, off);
                    v.w += __shfl_xor_sync(0xffffffffu, v.w, off);
                }
                const int col = (isa ? b : b - nba) * 32 + c4 * 4, w = isa ? a.wa : a.wb;
                if (lane < 8) {
                    if (col < w) atomicAdd(o + col, v.x);
                    if (col + 1 < w) atomicAdd(o + col + 1, v.y);
                    if (col + 2 < w) atomicAdd(o + col + 2, v.z);
                    if (col + 3 < w) atomicAdd(o + col + 3, v.w);
                }
            }
        }
        // epilogue: warps 0..3 own the 128 accumulator lanes
        if (warp < 4 && n_local > 0) {
            tc::mbar_wait(done, 0);
            tc::tc_fence_after();
            const int m = tid;      // lane of the accumulator = column of A
            const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);
            for (int n0 = 0; n0 < a.wb_pad; n0 += 16) {
                float v[16];
                tc::tmem_ld16(t_row + n0, v);
                tc::tmem_wait_ld();
                if (m < a.wa && a.dst != nullptr) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (n0 + j < a.wb) atomicAdd(a.dst + m * a.sm + (n0 + j) * a.sn, v[j]);
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, L::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// Same contraction with the operand rows brought in by the bulk-copy engine: a producer warp issues one
// cp.async.bulk per row and operand (32 lanes = the 32 rows of a chunk; row gathers and ragged tails come for free) into a
// ring of RAW chunks several stages deep, completion tracked by mbarrier transaction counts; the 256 splitter threads
// read the raw rows from shared memory, so no thread ever waits on HBM between a proxy fence and its next load (the
// register-prefetch kernel above is latency-bound: ~20 GB/s per SM).  Needs 16-byte aligned rows (lda, ldb, wa, wb
// multiples of 4); everything else takes the kernel above.
// ---------------------------------------------------------------------------------------------
template <int NBB>
struct WgTmaLayout {
    static constexpr int NBLK = WG_MA / 32 + NBB;
    static constexpr int IMG_STAGES = NBB <= 2 ? 2 : 1;              // hi/lo operand images (48 KB or 96 KB each)
    static constexpr int STAGE_FLOATS = 2 * NBLK * WG_BLK;
    static constexpr int OFF_RAW = IMG_STAGES * STAGE_FLOATS;        // raw ring: n_raw stages of 32 x (wa + wb) floats
    static constexpr int RAW_BUDGET_FLOATS = (227 * 1024 - 256) / 4 - OFF_RAW;
    static constexpr int MAX_RAW = 8;
    static constexpr int TMEM_COLS = NBB <= 2 ? 64 : 256;
};

template <int NBB>
__global__ void __launch_bounds__(WG_NT + 32, 1) wgrad_tma_kernel(const __grid_constant__ WgradArgs a, int n_raw) {
    using L = WgTmaLayout<NBB>;
    constexpr int IMG = L::IMG_STAGES, NBLK = L::NBLK;
    extern __shared__ __align__(1024) float smem[];
    const int raw_stage_floats = WG_R * (a.wa + a.wb);
    float* raw = smem + L::OFF_RAW;
    uint64_t* bars = reinterpret_cast<uint64_t*>(raw + n_raw * raw_stage_floats);
    uint64_t* raw_full = bars;                         // [MAX_RAW] bulk copies landed (transaction count)
    uint64_t* raw_empty = bars + L::MAX_RAW;           // [MAX_RAW] 256 splitters are done with the raw chunk
    uint64_t* img_full = bars + 2 * L::MAX_RAW;        // [2] splitters -> MMA
    uint64_t* img_empty = img_full + 2;                // [2] MMA (tcgen05.commit) -> splitters
    uint64_t* done = img_empty + 2;
    uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < L::MAX_RAW; ++s) { tc::mbar_init(&raw_full[s], 1); tc::mbar_init(&raw_empty[s], WG_LOADERS); }
        for (int s = 0; s < 2; ++s) { tc::mbar_init(&img_full[s], WG_LOADERS); tc::mbar_init(&img_empty[s], 1); }
        tc::mbar_init(done, 1);
        tc::mbar_init_fence();
    }
    if (warp == 0) tc::tmem_alloc(slot, L::TMEM_COLS);
    for (int i = tid; i < IMG * L::STAGE_FLOATS / 4; i += WG_NT + 32) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *slot;

    const long long n_chunks = (a.rows + WG_R - 1) / WG_R;
    const long long c_begin = n_chunks * blockIdx.x / gridDim.x, c_end = n_chunks * (blockIdx.x + 1) / gridDim.x;
    const int n_local = (int)(c_end - c_begin);
    const int np = a.passes == 1 ? 1 : 3;
    constexpr int OFF_LO = NBLK * WG_BLK, OFF_B = (WG_MA / 32) * WG_BLK;

    if (warp == WG_LOADERS / 32 + 1) {
        // =========================== producer warp: lane = row of the chunk ===========================
        for (int c = 0; c < n_local; ++c) {
            const int s = c % n_raw;
            if (c >= n_raw) tc::mbar_wait(&raw_empty[s], (uint32_t)(c / n_raw - 1) & 1u);
            const long long r0 = (c_begin + c) * WG_R;
            const int nrow = (int)(a.rows - r0 < WG_R ? a.rows - r0 : WG_R);
            float* dst = raw + s * raw_stage_floats;
            if (lane == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(&raw_full[s])),
                             "r"((uint32_t)(nrow * (a.wa + a.wb) * 4)) : "memory");
            }
            __syncwarp();
            // contiguous operands (leading dimension == width, no gather): ONE bulk copy per chunk; else one per row
            const bool ca = a.lda == a.wa, cb = a.ldb == a.wb && a.b_ridx == nullptr;
            auto bulk = [&](float* d, const float* g, uint32_t bytes) {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(tc::smem_u32(d)), "l"(g), "r"(bytes), "r"(tc::smem_u32(&raw_full[s])) : "memory");
            };
            if (ca) {
                if (lane == 0) bulk(dst, a.A + r0 * a.lda, (uint32_t)(nrow * a.wa * 4));
            } else if (lane < nrow) {
                bulk(dst + lane * a.wa, a.A + (r0 + lane) * a.lda, (uint32_t)(a.wa * 4));
            }
            if (cb) {
                if (lane == 1 % 32) bulk(dst + WG_R * a.wa, a.B + r0 * a.ldb, (uint32_t)(nrow * a.wb * 4));
            } else if (lane < nrow) {
                const long long r = r0 + lane;
                const long long rb = a.b_ridx != nullptr ? (long long)__ldg(a.b_ridx + r) : r;
                bulk(dst + WG_R * a.wa + lane * a.wb, a.B + rb * a.ldb, (uint32_t)(a.wb * 4));
            }
        }
    } else if (warp == WG_LOADERS / 32) {
        // =========================== MMA issue warp ===========================
        if (lane == 0 && n_local > 0) {
            const uint32_t idesc = idesc_tf32_mn(WG_MA, a.wb_pad);
            constexpr uint32_t LBO = WG_BLK * 4, SBO = 4 * 128, KSTEP = 8 * 128;
            bool acc = false;
            for (int c = 0; c < n_local; ++c) {
                const int t = c % IMG;
                tc::mbar_wait(&img_full[t], (uint32_t)(c / IMG) & 1u);
                tc::tc_fence_after();
                const uint32_t base = tc::smem_u32(smem + t * L::STAGE_FLOATS);
                for (int p = 0; p < np; ++p) {     // 3xTF32, small terms first: lo*hi, hi*lo, hi*hi
                    const int pa = (np == 1) ? 0 : (p == 0 ? 1 : 0), pb = (np == 1) ? 0 : (p == 1 ? 1 : 0);
                    const uint64_t ad0 = smem_desc_mn32(base + (pa ? OFF_LO : 0) * 4, LBO, SBO);
                    const uint64_t bd0 = smem_desc_mn32(base + ((pb ? OFF_LO : 0) + OFF_B) * 4, LBO, SBO);
#pragma unroll
                    for (int ks = 0; ks < WG_R / 8; ++ks) {
                        tc::mma_tf32_ss(tmem, ad0 + (uint64_t)((ks * KSTEP) >> 4), bd0 + (uint64_t)((ks * KSTEP) >> 4), idesc, acc);
                        acc = true;
                    }
                }
                tc::mma_commit(&img_empty[t]);
            }
            tc::mma_commit(done);
        }
        __syncwarp();
    } else {
        // =========================== splitter warps ===========================
        const int rr = warp * 4 + (lane >> 3), c4 = lane & 7;
        const int nba = WG_MA / 32, naa = (a.wa + 31) >> 5, nbb = (a.wb + 31) >> 5;
        const int sw = (rr * 32) + ((((c4 >> 1) ^ (rr & 3)) << 1 | (c4 & 1)) << 2);
        float4 csum[NBLK];
#pragma unroll
        for (int i = 0; i < NBLK; ++i) csum[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = 0; c < n_local; ++c) {
            const int s = c % n_raw, t = c % IMG;
            tc::mbar_wait(&raw_full[s], (uint32_t)(c / n_raw) & 1u);
            if (c >= IMG) tc::mbar_wait(&img_empty[t], (uint32_t)(c / IMG - 1) & 1u);
            const float* rs = raw + s * raw_stage_floats;
            float* st = smem + t * L::STAGE_FLOATS;
            const bool row_ok = (c_begin + c) * WG_R + rr < a.rows;
#pragma unroll
            for (int b = 0; b < NBLK; ++b) {
                const bool isa = b < nba;
                const int bl = isa ? b : b - nba;
                if (bl >= (isa ? naa : nbb)) continue;
                const int col = bl * 32 + c4 * 4, w = isa ? a.wa : a.wb;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row_ok && col < w) v = *reinterpret_cast<const float4*>(rs + (isa ? rr * a.wa : WG_R * a.wa + rr * a.wb) + col);
                float4 hi, lo;
                tc::split_tf32(v.x, hi.x, lo.x);
                tc::split_tf32(v.y, hi.y, lo.y);
                tc::split_tf32(v.z, hi.z, lo.z);
                tc::split_tf32(v.w, hi.w, lo.w);
                *reinterpret_cast<float4*>(st + b * WG_BLK + sw) = hi;
                *reinterpret_cast<float4*>(st + OFF_LO + b * WG_BLK + sw) = lo;
                csum[b].x += v.x; csum[b].y += v.y; csum[b].z += v.z; csum[b].w += v.w;
            }
            tc::fence_async_smem();
            tc::mbar_arrive(&img_full[t]);
            tc::mbar_arrive(&raw_empty[s]);
        }
        if (n_local > 0 && (a.colsum_a != nullptr || a.colsum_b != nullptr)) {
#pragma unroll
            for (int b = 0; b < NBLK; ++b) {
                const bool isa = b < nba;
                const int bl = isa ? b : b - nba;
                float* o = isa ? a.colsum_a : a.colsum_b;
                if (o == nullptr || bl >= (isa ? naa : nbb)) continue;      // warp-uniform
                float4 v = csum[b];
#pragma unroll
                for (int off = 8; off < 32; off <<= 1) {
                    v.x += __shfl_xor_sync(0xffffffffu, v.x, off);
                    v.y += __shfl_xor_sync(0xffffffffu, v.y, off);
                    v.z += __shfl_xor_sync(0xffffffffu, v.z, off);
                    v.w += __shfl_xor_sync(0xffffffffu, v.w, off);
                }
                const int col = bl * 32 + c4 * 4, w = isa ? a.wa : a.wb;
                if (lane < 8) {
                    if (col < w) atomicAdd(o + col, v.x);
                    if (col + 1 < w) atomicAdd(o + col + 1, v.y);
                    if (col + 2 < w) atomicAdd(o + col + 2, v.z);
                    if (col + 3 < w) atomicAdd(o + col + 3, v.w);
                }
            }
        }
        if (warp < 4 && n_local > 0) {
            tc::mbar_wait(done, 0);
            tc::tc_fence_after();
            const int m = tid;
            const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);
            for (int n0 = 0; n0 < a.wb_pad; n0 += 16) {
                float v[16];
                tc::tmem_ld16(t_row + n0, v);
                tc::tmem_wait_ld();
                if (m < a.wa && a.dst != nullptr) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (n0 + j < a.wb) atomicAdd(a.dst + m * a.sm + (n0 + j) * a.sn, v[j]);
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, L::TMEM_COLS);
}

template <int NBB>
static int launch_wgrad_tma(const WgradArgs& a, long long n_chunks, cudaStream_t stream) {
    using L = WgTmaLayout<NBB>;
    const int raw_stage_floats = WG_R * (a.wa + a.wb);
    int n_raw = L::RAW_BUDGET_FLOATS / raw_stage_floats;
    if (n_raw > L::MAX_RAW) n_raw = L::MAX_RAW;
    if (n_raw < 2) return -1;                 // not enough room for a ring: caller falls back
    const size_t smem = ((size_t)L::OFF_RAW + (size_t)n_raw * raw_stage_floats) * 4 + (2 * L::MAX_RAW + 6) * 8 + 16;
    static size_t configured = 0;
    if (smem > configured) {
        RGNN_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tma_kernel<NBB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = 227 * 1024;
    }
    long long grid = n_chunks / 8;
    if (grid < 1) grid = 1;
    if (grid > sm_count()) grid = sm_count();
    wgrad_tma_kernel<NBB><<<(int)grid, WG_NT + 32, smem, stream>>>(a, n_raw);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

// D = A^T B added into dst (see WgradArgs); rows may be 0
int launch_wgrad_tc(const float* A, int lda, int wa, const float* B, int ldb, int wb, long long rows, float* dst, long long sm,
                    long long sn, float* colsum_a, float* colsum_b, cudaStream_t stream, const int* b_ridx) {
    if (rows <= 0 || (dst == nullptr && colsum_a == nullptr && colsum_b == nullptr)) return RGNN_OK;
    RGNN_REQUIRE(wa >= 1 && wa <= WG_MA && wb >= 1 && wb <= WG_NB_MAX, "wgrad: operand widths %d x %d outside 128 x 256", wa, wb);
    static PerDeviceOnce once;
    if (once.needed()) {
        RGNN_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WgLayout<2>::SMEM));
        RGNN_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WgLayout<8>::SMEM));
        once.mark();
    }
    WgradArgs a;
    a.A = A; a.B = B; a.lda = lda; a.ldb = ldb; a.wa = wa; a.wb = wb;
    a.wa_pad = WG_MA; a.wb_pad = round_up(wb, 16);
    a.rows = rows; a.dst = dst; a.sm = sm; a.sn = sn; a.colsum_a = colsum_a; a.colsum_b = colsum_b; a.b_ridx = b_ridx;
    a.passes = rgnn_get_option("tf32_passes");
    const long long n_chunks = (rows + WG_R - 1) / WG_R;
    // split-K: enough chunks per CTA to amortise the final 128 x wb reduction into the gradient buffer
    const bool narrow = wb <= 64;
    // bulk-copy fed variant for contiguous, 16-byte aligned operands (one copy per operand and chunk; per-row copies of
    // strided / gathered operands were measured slower than the register path)
    if (lda == wa && ldb == wb && b_ridx == nullptr && ((lda | ldb | wa | wb) & 3) == 0 && ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0 &&
        rgnn_get_option("wgrad_tma") != 0) {
        const int rc = narrow ? launch_wgrad_tma<2>(a, n_chunks, stream) : launch_wgrad_tma<8>(a, n_chunks, stream);
        if (rc >= 0) return rc;
    }
    const long long max_grid = (long long)sm_count() * (narrow ? 2 : 1);
    long long grid = n_chunks / 8;
    if (grid < 1) grid = 1;
    if (grid > max_grid) grid = max_grid;
    if (narrow) wgrad_tc_kernel<2><<<(int)grid, WG_NT, WgLayout<2>::SMEM, stream>>>(a);
    else wgrad_tc_kernel<8><<<(int)grid, WG_NT, WgLayout<8>::SMEM, stream>>>(a);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

}  // namespace rgnn

// Test / integration hook: dst (wa x wb, row-major, leading dimension wb) += A^T B; colsum_* may be NULL.
extern "C" int rgnn_wgrad(const float* A, int lda, int wa, const float* B, int ldb, int wb, long long rows, float* dst,
                          float* colsum_a, float* colsum_b, void* stream) {
    return rgnn::launch_wgrad_tc(A, lda, wa, B, ldb, wb, rows, dst, wb, 1, colsum_a, colsum_b, static_cast<cudaStream_t>(stream), nullptr);
}
