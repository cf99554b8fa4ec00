// Fixed-shape 64-wide ffn_block chains on the tensor cores with fp16-split operands (rgnn_f16.cuh): the stems and
// FFN_TaskSpecificHeads behind the message-passing layers (reference gnn_blocks.py:167-344: node_segmentation,
// node_offset_predictions, edge_formation + link_predictions, the stem of object_classification):
//
//   input rows (n, 64) fp32  [or pair sums h[a] + h[b] for the link head]
//     -> N_HIDDEN x ( Linear 64 -> 64, channel_normalization, LeakyReLU )
//     -> either the activation itself (n, 64) fp32   or   a bare Linear 64 -> n_out (n_out <= 16)
//
// One CTA per SM, persistent.  Thread = TMEM lane = row.  FOUR tiles of 128 rows are in flight per CTA, each owned by one group
// of 4 epilogue warps that carries its tile through all layers (no row statistic crosses a thread, no hand-over between
// groups); the single MMA-issue lane serves whichever group has an operand ready, so one group's epilogue runs under the
// other groups' MMAs.  Every weight image (<= 68 KB as fp16 hi | lo) stays in shared memory for the whole kernel; MMA shapes
// and descriptors are compile-time constants.  Tensor memory: 128 columns per group = two 64-column regions that alternate
// as A operand / accumulator; an epilogue rewrites its accumulator in place into the next A operand (per 32 fp32 columns:
// 16 packed hi columns | 16 packed lo columns).
#include "rgnn_f16.cuh"
#include "rgnn_model.h"
#include "rgnn_tc_rows.cuh"

namespace rgnn {

constexpr int C64_MAX_LAYERS = 5;        // up to 4 hidden + 1 tail Linear

struct Chain64Args {
    int n_rows;
    int in_mode;                // 0: rows of x; 1: pair sums x[ia[r]] + x[ib[r]]
    const float* x;             // (n, 64) fp32, row stride ldx
    int ldx;
    const int* ia;
    const int* ib;
    int n_hidden;               // 1 .. 4
    int tail;                   // 0: store the last activation (n, 64); 1: bare Linear to n_out columns
    int n_out;
    const uint32_t* w[C64_MAX_LAYERS];      // per layer [hi | lo] images (K = 64; N = 64, tail: 16 rows, zero padded)
    const float* bias[C64_MAX_LAYERS];
    const float* scale[C64_MAX_LAYERS];     // channel_normalization gain (device scalar) or nullptr
    const float* shift[C64_MAX_LAYERS];
    int act[C64_MAX_LAYERS];
    float* y;
    int ldy;
    // a training step: what the backward (chain64_bwd_f16_kernel) reads back -- the assembled input rows (pair sums), the outputs of
    // the hidden layers except the stack's own output, the sigmas (each nullable)
    float* save_x;
    float* save_y[C64_MAX_LAYERS];
    float* save_sd[C64_MAX_LAYERS];
    int passes;
    int prefetch;               // input rows of the next tile are requested into L2 one tile ahead (pair mode: indices two tiles ahead)
};

namespace c64 {
constexpr int W = 64, TM = 128, NGROUPS = 4, NT_OUT = 16;
constexpr int NTHREADS = 128 * NGROUPS + 128;          // + the warp group holding the MMA-issue warp
constexpr int IMG_WORDS = W * W;                       // hi + lo of a 64 x 64 layer (2 x 64 x 64 x 2 B)
constexpr int TAIL_WORDS = W * NT_OUT;
constexpr int OFF_W = 0;
constexpr int OFF_CST = OFF_W + 4 * IMG_WORDS + TAIL_WORDS;         // per layer: bias[64], gain, shift (stride 72)
constexpr int CST_LD = 72;
constexpr int OFF_BAR = OFF_CST + C64_MAX_LAYERS * CST_LD;          // a_full[4], d_full[4]
constexpr int OFF_SLOT = OFF_BAR + 2 * 2 * NGROUPS;
constexpr int WORDS = OFF_SLOT + 2;
constexpr size_t SMEM = (size_t)WORDS * 4;
static_assert((OFF_BAR % 2) == 0, "mbarrier alignment");
}  // namespace c64

__global__ void __launch_bounds__(c64::NTHREADS, 1) chain64_f16_kernel(const __grid_constant__ Chain64Args a) {
    using namespace c64;
    extern __shared__ __align__(1024) uint32_t smem_u[];
    float* smem_f = reinterpret_cast<float*>(smem_u);
    uint32_t* wsm = smem_u + OFF_W;
    float* cst = smem_f + OFF_CST;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_u + OFF_BAR);
    uint64_t* a_full = bars;
    uint64_t* d_full = bars + NGROUPS;
    uint32_t* slot = smem_u + OFF_SLOT;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = warp >> 2, w4 = warp & 3, row = tid & 127;
    const int G = (int)gridDim.x;
    const int n_tiles = (a.n_rows + TM - 1) / TM;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / G + 1 : 0;
    const int np = a.passes == 1 ? 1 : 3;
    const int n_mma = a.n_hidden + (a.tail ? 1 : 0);

    // ---- one-time setup ----
    for (int l = 0; l < n_mma; ++l) {
        const int words = l < a.n_hidden ? IMG_WORDS : TAIL_WORDS;
        const uint4* src = reinterpret_cast<const uint4*>(a.w[l]);
        uint4* dst = reinterpret_cast<uint4*>(wsm + l * IMG_WORDS);
        for (int i = tid; i < words / 4; i += NTHREADS) dst[i] = __ldg(src + i);
    }
    for (int i = tid; i < n_mma * CST_LD; i += NTHREADS) {
        const int l = i / CST_LD, c = i - l * CST_LD;
        const int nb = l < a.n_hidden ? W : a.n_out;
        float v = 0.f;
        if (c < W) v = (a.bias[l] != nullptr && c < nb) ? __ldg(a.bias[l] + c) : 0.f;
        else if (c == W) v = a.scale[l] != nullptr ? __ldg(a.scale[l]) : 1.f;
        else if (c == W + 1) v = a.shift[l] != nullptr ? __ldg(a.shift[l]) : 0.f;
        cst[i] = v;
    }
    if (tid == 0) {
        for (int i = 0; i < NGROUPS; ++i) { tc::mbar_init(&a_full[i], 4); tc::mbar_init(&d_full[i], 1); }
        tc::mbar_init_fence();
    }
    if (warp == 0) tc::tmem_alloc(slot, 512);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *slot;

    if (g < NGROUPS) {
        // =========================== epilogue groups ===========================
        const uint32_t t_base = tmem + ((uint32_t)w4 << 21) + (uint32_t)g * 128;    // this group's 128 columns, this warp's lanes
        uint32_t uses = 0;          // completed phases of this group's barriers (one per MMA group)
        const bool pf = a.prefetch != 0;
        auto idx_of = [&](const int* ix, int jj) -> int {
            const long long rq = ((long long)blockIdx.x + (long long)jj * G) * TM + row;
            return (jj < my_tiles && rq < a.n_rows) ? __ldg(ix + rq) : -1;
        };
        auto prefetch_row = [&](const float* q) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(q + 32));
        };
        int a1 = -1, b1 = -1, a2 = -1, b2 = -1;     // pair mode: indices of this group's next tile / the one after
        if (pf && a.in_mode == 1) { a1 = idx_of(a.ia, g); b1 = idx_of(a.ib, g); a2 = idx_of(a.ia, g + NGROUPS); b2 = idx_of(a.ib, g + NGROUPS); }
        for (int j = g; j < my_tiles; j += NGROUPS) {
            const int tile = (int)blockIdx.x + j * G;
            const int r = tile * TM + row;
            const bool valid = r < a.n_rows;
            // ---- input rows -> A operand in region 0 (x 16, hi | lo per 32 columns) ----
            {
                const float* pa = a.x;
                const float* pb = nullptr;
                if (a.in_mode == 1) {
                    int na, nb;
                    if (pf) {
                        na = valid ? a1 : 0; nb = valid ? b1 : 0;
                        a1 = a2; b1 = b2;
                        if (a1 >= 0) { prefetch_row(a.x + (size_t)a1 * a.ldx); prefetch_row(a.x + (size_t)b1 * a.ldx); }
                        a2 = idx_of(a.ia, j + 2 * NGROUPS); b2 = idx_of(a.ib, j + 2 * NGROUPS);
                    } else {
                        na = valid ? __ldg(a.ia + r) : 0; nb = valid ? __ldg(a.ib + r) : 0;
                    }
                    pa = a.x + (size_t)na * a.ldx;
                    pb = a.x + (size_t)nb * a.ldx;
                } else {
                    pa = a.x + (size_t)(valid ? r : 0) * a.ldx;
                    if (pf) {
                        const long long rn = (long long)r + (long long)NGROUPS * G * TM;
                        if (rn < a.n_rows) prefetch_row(a.x + (size_t)rn * a.ldx);
                    }
                }
                const float2 s16 = make_float2(f16::A_SCALE, f16::A_SCALE);
#pragma unroll 1
                for (int c = 0; c < W; c += 32) {
                    float2 v[16];
#pragma unroll
                    for (int i = 0; i < 4; ++i) ldg256(pa + c + 8 * i, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    if (pb != nullptr) {
                        float2 u[16];
#pragma unroll
                        for (int i = 0; i < 4; ++i) ldg256(pb + c + 8 * i, u[4 * i], u[4 * i + 1], u[4 * i + 2], u[4 * i + 3]);
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = __fadd2_rn(v[i], u[i]);
                    }
                    if (a.save_x != nullptr && valid) {
                        float* o = a.save_x + (size_t)r * W + c;
#pragma unroll
                        for (int i = 0; i < 4; ++i) stg256(o + 8 * i, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    }
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) f16::split(valid ? __fmul2_rn(v[i], s16) : make_float2(0.f, 0.f), hi[i], lo[i]);
                    f16::tmem_st16u(t_base + c, hi);
                    if (np != 1) f16::tmem_st16u(t_base + c + 16, lo);
                }
                tc::tmem_wait_st();
                tc::tc_fence_before();
                warp_arrive(&a_full[g], lane);
            }
            // ---- hidden layers ----
            for (int l = 0; l < a.n_hidden; ++l) {
                const float* cb = cst + l * CST_LD;
                const uint32_t dreg = t_base + (uint32_t)(((l + 1) & 1) * W);
                tc::mbar_wait(&d_full[g], uses & 1u);
                ++uses;
                tc::tc_fence_after();
                float2 va[16], vb[16];
                tc::tmem_ld16(dreg, va);
                tc::tmem_ld16(dreg + 16, va + 8);
                tc::tmem_ld16(dreg + 32, vb);
                tc::tmem_ld16(dreg + 48, vb + 8);
                tc::tmem_wait_ld();
                const float2 us = make_float2(f16::D_UNSCALE, f16::D_UNSCALE);
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    va[c] = __ffma2_rn(va[c], us, *reinterpret_cast<const float2*>(cb + 2 * c));
                    vb[c] = __ffma2_rn(vb[c], us, *reinterpret_cast<const float2*>(cb + 32 + 2 * c));
                }
                const bool last = (l + 1 == a.n_hidden) && !a.tail;
                const float osc = last ? 1.f : f16::A_SCALE;          // the next A operand carries x 16
                float k = osc, sh = 0.f, mean = 0.f;
                if (a.scale[l] != nullptr) {
                    RowStats st;
                    st.init();
                    st.add_chunk(va);
                    st.add_chunk(vb);
                    const float sd = st.sigma(W);
                    if (a.save_sd[l] != nullptr && valid) a.save_sd[l][r] = sd;
                    k = osc * cb[W] * __frcp_rn(sd + NORM_EPS);
                    sh = osc * cb[W + 1];
                    mean = st.mean;
                }
                const float2 k2 = make_float2(k, k), sh2 = make_float2(sh, sh), sl = make_float2(LEAKY, LEAKY), nm = make_float2(-mean, -mean);
                const bool act = a.act[l] != 0;
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    va[c] = __ffma2_rn(__fadd2_rn(va[c], nm), k2, sh2);
                    vb[c] = __ffma2_rn(__fadd2_rn(vb[c], nm), k2, sh2);
                    if (act) {
                        const float2 ta = __fmul2_rn(va[c], sl), tb = __fmul2_rn(vb[c], sl);
                        va[c].x = fmaxf(va[c].x, ta.x); va[c].y = fmaxf(va[c].y, ta.y);
                        vb[c].x = fmaxf(vb[c].x, tb.x); vb[c].y = fmaxf(vb[c].y, tb.y);
                    }
                }
                if (last) {
                    if (valid) {
                        float* o = a.y + (size_t)r * a.ldy;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            stg256(o + 8 * i, va[4 * i], va[4 * i + 1], va[4 * i + 2], va[4 * i + 3]);
                            stg256(o + 32 + 8 * i, vb[4 * i], vb[4 * i + 1], vb[4 * i + 2], vb[4 * i + 3]);
                        }
                    }
                } else {
                    if (a.save_y[l] != nullptr && valid) {      // the layer output at its true scale (the operand carries x 16)
                        const float2 un = make_float2(1.f / f16::A_SCALE, 1.f / f16::A_SCALE);
                        float* o = a.save_y[l] + (size_t)r * W;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            stg256(o + 8 * i, __fmul2_rn(va[4 * i], un), __fmul2_rn(va[4 * i + 1], un), __fmul2_rn(va[4 * i + 2], un), __fmul2_rn(va[4 * i + 3], un));
                            stg256(o + 32 + 8 * i, __fmul2_rn(vb[4 * i], un), __fmul2_rn(vb[4 * i + 1], un), __fmul2_rn(vb[4 * i + 2], un), __fmul2_rn(vb[4 * i + 3], un));
                        }
                    }
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) f16::split(va[i], hi[i], lo[i]);
                    f16::tmem_st16u(dreg, hi);
                    if (np != 1) f16::tmem_st16u(dreg + 16, lo);
#pragma unroll
                    for (int i = 0; i < 16; ++i) f16::split(vb[i], hi[i], lo[i]);
                    f16::tmem_st16u(dreg + 32, hi);
                    if (np != 1) f16::tmem_st16u(dreg + 48, lo);
                    tc::tmem_wait_st();
                    tc::tc_fence_before();
                    warp_arrive(&a_full[g], lane);
                }
            }
            // ---- tail Linear (<= 16 outputs, no norm / activation) ----
            if (a.tail) {
                const float* cb = cst + a.n_hidden * CST_LD;
                const uint32_t dreg = t_base + (uint32_t)(((a.n_hidden + 1) & 1) * W);
                tc::mbar_wait(&d_full[g], uses & 1u);
                ++uses;
                tc::tc_fence_after();
                float2 v[8];
                tc::tmem_ld16(dreg, v);
                tc::tmem_wait_ld();
                if (valid) {
                    float* o = a.y + (size_t)r * a.ldy;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        if (2 * c < a.n_out) o[2 * c] = fmaf(v[c].x, f16::D_UNSCALE, cb[2 * c]);
                        if (2 * c + 1 < a.n_out) o[2 * c + 1] = fmaf(v[c].y, f16::D_UNSCALE, cb[2 * c + 1]);
                    }
                }
            }
            tc::tc_fence_before();      // this tile's last TMEM reads precede the next tile's operand stores
        }
    } else if (warp == 4 * NGROUPS) {
        // =========================== MMA issue warp ===========================
        if (lane == 0) {
            constexpr uint32_t IDESC64 = f16::idesc(TM, W), IDESC16 = f16::idesc(TM, NT_OUT);
            const uint32_t sW = tc::smem_u32(wsm);
            int tile_j[NGROUPS], layer[NGROUPS];
            uint32_t uses[NGROUPS];
            int remaining = 0;
            for (int i = 0; i < NGROUPS; ++i) {
                tile_j[i] = i; layer[i] = 0; uses[i] = 0;
                if (i < my_tiles) ++remaining;
            }
            while (remaining > 0) {
                bool did = false;
#pragma unroll
                for (int i = 0; i < NGROUPS; ++i) {
                    if (tile_j[i] >= my_tiles) continue;
                    if (!f16::mbar_test(&a_full[i], uses[i] & 1u)) continue;
                    tc::tc_fence_after();
                    const int l = layer[i];
                    const bool is_tail = l == a.n_hidden;
                    const uint32_t areg = tmem + (uint32_t)i * 128 + (uint32_t)((l & 1) * W);
                    const uint32_t dreg = tmem + (uint32_t)i * 128 + (uint32_t)(((l + 1) & 1) * W);
                    const uint32_t nrows = is_tail ? NT_OUT : W;
                    const uint32_t lbo = nrows * 16, img = W * nrows * 2;       // bytes: next 8 K elements; hi image size
                    const uint32_t wl = sW + (uint32_t)l * (IMG_WORDS * 4);
                    bool acc = false;
                    for (int p = 0; p < np; ++p) {      // small terms first: lo*hi, hi*lo, then hi*hi
                        const int pa = (np == 1) ? 0 : (p == 0 ? 1 : 0);
                        const int pb = (np == 1) ? 0 : (p == 1 ? 1 : 0);
                        const uint64_t bd0 = tc::smem_desc(wl + pb * img, lbo, 128);
#pragma unroll
                        for (int ks = 0; ks < W / 16; ++ks) {
                            const uint32_t acol = areg + (ks >> 1) * 32 + (ks & 1) * 8 + (pa ? 16u : 0u);
                            f16::mma_ts(dreg, acol, bd0 + (uint64_t)((ks * 2 * lbo) >> 4), is_tail ? IDESC16 : IDESC64, acc);
                            acc = true;
                        }
                    }
                    tc::mma_commit(&d_full[i]);
                    ++uses[i];
                    if (++layer[i] == n_mma) {
                        layer[i] = 0;
                        tile_j[i] += NGROUPS;
                        if (tile_j[i] >= my_tiles) --remaining;
                    }
                    did = true;
                }
                if (!did) __nanosleep(20);
            }
        }
        __syncwarp();
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int g_chain_f16 = 1;
static int g_rows_prefetch = 1;     // row-owning kernels request their next tile's rows into L2 one tile ahead (prefetch.global.L2)
int chain_f16_set_option(const char* name, int value) {
    if (strcmp(name, "f16_chain") == 0 && (value == 0 || value == 1)) { g_chain_f16 = value; return 1; }
    if (strcmp(name, "rows_prefetch") == 0 && (value == 0 || value == 1)) { g_rows_prefetch = value; return 1; }
    return 0;
}
int chain_f16_get_option(const char* name) {
    if (strcmp(name, "f16_chain") == 0) return g_chain_f16;
    if (strcmp(name, "rows_prefetch") == 0) return g_rows_prefetch;
    return -2;
}

// floats appended to a Linear's packed buffer for its fp16 image pair (0: the layer has none)
size_t f16_image_floats(int in_features, int out_features) {
    if (in_features > 256 || out_features > 256) return 0;
    if (in_features < 16) return (size_t)16 * round_up(out_features, 16);       // raw feature inputs: K zero padded to 16
    if (in_features % 16 != 0) return 0;
    return (size_t)in_features * round_up(out_features, 16);
}

int pack_f16_image(const float* W, int ldw, int K, int N, int n_valid, int k_valid, uint32_t* dst, cudaStream_t stream);

int f16_pack_linear(const rgnn_linear& L, float* dst, cudaStream_t stream) {
    if (f16_image_floats(L.in_features, L.out_features) == 0) return RGNN_OK;
    const int K = L.in_features < 16 ? 16 : L.in_features;
    return pack_f16_image(L.weight, L.in_features, K, round_up(L.out_features, 16), L.out_features, L.in_features,
                          reinterpret_cast<uint32_t*>(dst), stream);
}

// stack = N_HIDDEN x (64 -> 64, norm, act) [+ bare Linear 64 -> n_out <= 16]
bool chain64_supported(const rgnn_stack& s) {
    if (!g_chain_f16 || s.n < 1 || s.n > C64_MAX_LAYERS) return false;
    int n_hidden = s.n;
    const rgnn_linear& last = s.layer[s.n - 1];
    const bool tail = last.norm_scale == nullptr && !last.activation;
    if (tail) --n_hidden;
    if (n_hidden < 1 || n_hidden > 4) return false;
    for (int i = 0; i < n_hidden; ++i) {
        const rgnn_linear& L = s.layer[i];
        if (L.in_features != 64 || L.out_features != 64 || L.norm_scale == nullptr || L.weight_t == nullptr) return false;
    }
    if (tail && (last.in_features != 64 || last.out_features > 16 || last.weight_t == nullptr)) return false;
    return true;
}

const float* f16_weights(const rgnn_linear& L);      // rgnn_model_tc.cu

int run_chain64(const rgnn_stack& s, const float* x, int ldx, const int* ia, const int* ib, int n_rows, float* y, cudaStream_t stream,
                const TcSave* save) {
    if (n_rows <= 0) return RGNN_OK;
    Chain64Args a;
    memset(&a, 0, sizeof(a));
    const rgnn_linear& last = s.layer[s.n - 1];
    a.tail = (last.norm_scale == nullptr && !last.activation) ? 1 : 0;
    a.n_hidden = s.n - a.tail;
    a.n_out = a.tail ? last.out_features : 64;
    a.n_rows = n_rows;
    a.in_mode = ia != nullptr ? 1 : 0;
    a.x = x; a.ldx = ldx; a.ia = ia; a.ib = ib;
    for (int l = 0; l < s.n; ++l) {
        const rgnn_linear& L = s.layer[l];
        a.w[l] = reinterpret_cast<const uint32_t*>(f16_weights(L));
        a.bias[l] = L.bias; a.scale[l] = L.norm_scale; a.shift[l] = L.norm_shift; a.act[l] = L.activation;
    }
    a.y = y; a.ldy = a.n_out;
    if (save != nullptr) {
        a.save_x = save->x_in;
        for (int l = 0; l < a.n_hidden; ++l) { a.save_y[l] = (l + 1 < s.n) ? save->y[l] : nullptr; a.save_sd[l] = save->sd[l]; }
    }
    a.passes = mp_f16_passes();
    a.prefetch = g_rows_prefetch;
    static PerDeviceOnce once;
    if (once.needed()) {
        RGNN_CHECK_CUDA(cudaFuncSetAttribute(chain64_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c64::SMEM));
        once.mark();
    }
    const int n_tiles = (n_rows + c64::TM - 1) / c64::TM;
    const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
    chain64_f16_kernel<<<grid, c64::NTHREADS, c64::SMEM, stream>>>(a);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

}  // namespace rgnn

// =====================================================================================================================
// Node update of residual_graph_conv_block (reference gnn_blocks.py:96-111) + the hoisted projection of the NEXT block:
//   out = x + ffn_block(128 -> 64)(cat(x, agg))                          (identity residual, reference plan)
//   P_next = [out W_t^T + b | out W_s^T]   (N, 256), W_t / W_s = node columns of the next block's msg.0
// Same structure as chain64_f16_kernel with TWO tiles in flight (256 TMEM columns per group: X = 128 columns for the
// cat(x, agg) operand and later the two projection halves, Y = 64 columns for the update accumulator / `out` operand).
// HBM-bound: 512 B read + 256 B + 1 KB written per node.
// =====================================================================================================================
namespace rgnn {

struct ConvNodesArgs {
    int n_rows;
    const float* x;             // (N, 64)
    const float* agg;           // (N, 64)
    const uint32_t* w_upd;      // [hi | lo] image K = 128, N = 64
    const float* b_upd; const float* s_upd; const float* m_upd;
    int act;
    float* out;                 // (N, 64)
    int has_next;
    const uint32_t* w_pt;       // next block: target half, [hi | lo] image K = 64, N = 128
    const uint32_t* w_ps;       // source half
    const float* b_p;           // msg.0 bias of the next block (rides on the target half) or nullptr
    float* P;                   // (N, 256)
    float* u_save;              // training: the update's output before the residual (N, 64) and its sigma (N), or nullptr
    float* sd_save;
    int passes;
    int prefetch;
};

namespace cnn {
constexpr int TM = 128, NGROUPS = 2;
constexpr int NTHREADS = 128 * NGROUPS + 128;
constexpr int WU_WORDS = 128 * 64;          // hi + lo of the update layer (32 KB)
constexpr int WP_WORDS = 64 * 128;          // hi + lo of one projection half (32 KB)
constexpr int OFF_WU = 0, OFF_WPT = OFF_WU + WU_WORDS, OFF_WPS = OFF_WPT + WP_WORDS;
constexpr int OFF_CST = OFF_WPS + WP_WORDS;             // b_upd[64] | gain | shift | pad | b_p[128]
constexpr int OFF_BAR = OFF_CST + 64 + 8 + 128;
constexpr int OFF_SLOT = OFF_BAR + 2 * 2 * NGROUPS;
constexpr int WORDS = OFF_SLOT + 2;
constexpr size_t SMEM = (size_t)WORDS * 4;
static_assert((OFF_BAR % 2) == 0, "mbarrier alignment");
}  // namespace cnn

__global__ void __launch_bounds__(cnn::NTHREADS, 1) conv_nodes_f16_kernel(const __grid_constant__ ConvNodesArgs a) {
    using namespace cnn;
    extern __shared__ __align__(1024) uint32_t smem_u[];
    float* cst = reinterpret_cast<float*>(smem_u) + OFF_CST;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_u + OFF_BAR);
    uint64_t* a_full = bars;
    uint64_t* d_full = bars + NGROUPS;
    uint32_t* slot = smem_u + OFF_SLOT;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = warp >> 2, w4 = warp & 3, row = tid & 127;
    const int G = (int)gridDim.x;
    const int n_tiles = (a.n_rows + TM - 1) / TM;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / G + 1 : 0;
    const int np = a.passes == 1 ? 1 : 3;
    {
        auto copy = [&](uint32_t* dst, const uint32_t* src, int words) {
            const uint4* s4 = reinterpret_cast<const uint4*>(src);
            uint4* d4 = reinterpret_cast<uint4*>(dst);
            for (int i = tid; i < words / 4; i += NTHREADS) d4[i] = __ldg(s4 + i);
        };
        copy(smem_u + OFF_WU, a.w_upd, WU_WORDS);
        if (a.has_next) {
            copy(smem_u + OFF_WPT, a.w_pt, WP_WORDS);
            copy(smem_u + OFF_WPS, a.w_ps, WP_WORDS);
        }
        for (int i = tid; i < 64 + 8 + 128; i += NTHREADS) {
            float v = 0.f;
            if (i < 64) v = a.b_upd ? __ldg(a.b_upd + i) : 0.f;
            else if (i == 64) v = a.s_upd ? __ldg(a.s_upd) : 1.f;
            else if (i == 65) v = a.m_upd ? __ldg(a.m_upd) : 0.f;
            else if (i >= 72) v = (a.has_next && a.b_p) ? __ldg(a.b_p + (i - 72)) : 0.f;
            cst[i] = v;
        }
    }
    if (tid == 0) {
        for (int i = 0; i < NGROUPS; ++i) { tc::mbar_init(&a_full[i], 4); tc::mbar_init(&d_full[i], 1); }
        tc::mbar_init_fence();
    }
    if (warp == 0) tc::tmem_alloc(slot, 512);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *slot;
    const int n_mma = a.has_next ? 3 : 1;

    if (g < NGROUPS) {
        const uint32_t xr = tmem + ((uint32_t)w4 << 21) + (uint32_t)g * 256, yr = xr + 128;
        uint32_t uses = 0;
        const float2 s16 = make_float2(f16::A_SCALE, f16::A_SCALE), us = make_float2(f16::D_UNSCALE, f16::D_UNSCALE);
        for (int j = g; j < my_tiles; j += NGROUPS) {
            const int tile = (int)blockIdx.x + j * G;
            const int r = tile * TM + row;
            const bool valid = r < a.n_rows;
            const size_t rr = valid ? (size_t)r : 0;
            if (a.prefetch) {       // this group's next tile: its x / agg rows start their way from HBM to L2 one tile ahead
                const long long rn = (long long)r + (long long)NGROUPS * G * TM;
                if (rn < a.n_rows) {
                    const float* px = a.x + (size_t)rn * 64;
                    const float* pa = a.agg + (size_t)rn * 64;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(px));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(px + 32));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(pa));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(pa + 32));
                }
            }
            // ---- cat(x, agg) -> A operand (x 16, hi | lo per 32 columns) ----
#pragma unroll 1
            for (int c = 0; c < 128; c += 32) {
                const float* p = (c < 64 ? a.x + rr * 64 + c : a.agg + rr * 64 + (c - 64));
                float2 v[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) ldg256(p + 8 * i, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) f16::split(valid ? __fmul2_rn(v[i], s16) : make_float2(0.f, 0.f), hi[i], lo[i]);
                f16::tmem_st16u(xr + c, hi);
                if (np != 1) f16::tmem_st16u(xr + c + 16, lo);
            }
            tc::tmem_wait_st();
            tc::tc_fence_before();
            warp_arrive(&a_full[g], lane);
            // ---- update layer epilogue: norm, act, + x, store, next operand ----
            float2 xa[16], xb[16];          // the residual row, requested before the wait
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ldg256(a.x + rr * 64 + 8 * i, xa[4 * i], xa[4 * i + 1], xa[4 * i + 2], xa[4 * i + 3]);
                ldg256(a.x + rr * 64 + 32 + 8 * i, xb[4 * i], xb[4 * i + 1], xb[4 * i + 2], xb[4 * i + 3]);
            }
            tc::mbar_wait(&d_full[g], uses & 1u);
            ++uses;
            tc::tc_fence_after();
            float2 va[16], vb[16];
            tc::tmem_ld16(yr, va);
            tc::tmem_ld16(yr + 16, va + 8);
            tc::tmem_ld16(yr + 32, vb);
            tc::tmem_ld16(yr + 48, vb + 8);
            tc::tmem_wait_ld();
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                va[c] = __ffma2_rn(va[c], us, *reinterpret_cast<const float2*>(cst + 2 * c));
                vb[c] = __ffma2_rn(vb[c], us, *reinterpret_cast<const float2*>(cst + 32 + 2 * c));
            }
            float k = 1.f, sh = 0.f, mean = 0.f;
            if (a.s_upd != nullptr) {
                RowStats st;
                st.init();
                st.add_chunk(va);
                st.add_chunk(vb);
                const float sd = st.sigma(64);
                if (a.sd_save != nullptr && valid) a.sd_save[r] = sd;
                k = cst[64] * __frcp_rn(sd + NORM_EPS);
                sh = cst[65];
                mean = st.mean;
            }
            const float2 k2 = make_float2(k, k), sh2 = make_float2(sh, sh), sl = make_float2(LEAKY, LEAKY), nm = make_float2(-mean, -mean);
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                va[c] = __ffma2_rn(__fadd2_rn(va[c], nm), k2, sh2);
                vb[c] = __ffma2_rn(__fadd2_rn(vb[c], nm), k2, sh2);
                if (a.act) {
                    const float2 ta = __fmul2_rn(va[c], sl), tb = __fmul2_rn(vb[c], sl);
                    va[c].x = fmaxf(va[c].x, ta.x); va[c].y = fmaxf(va[c].y, ta.y);
                    vb[c].x = fmaxf(vb[c].x, tb.x); vb[c].y = fmaxf(vb[c].y, tb.y);
                }
            }
            if (a.u_save != nullptr && valid) {        // a training step keeps the update's output for the backward (upd_bwd_f16_kernel)
                float* o = a.u_save + (size_t)r * 64;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    stg256(o + 8 * i, va[4 * i], va[4 * i + 1], va[4 * i + 2], va[4 * i + 3]);
                    stg256(o + 32 + 8 * i, vb[4 * i], vb[4 * i + 1], vb[4 * i + 2], vb[4 * i + 3]);
                }
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                va[c] = __fadd2_rn(va[c], xa[c]);       // identity residual
                vb[c] = __fadd2_rn(vb[c], xb[c]);
            }
            if (valid) {
                float* o = a.out + (size_t)r * 64;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    stg256(o + 8 * i, va[4 * i], va[4 * i + 1], va[4 * i + 2], va[4 * i + 3]);
                    stg256(o + 32 + 8 * i, vb[4 * i], vb[4 * i + 1], vb[4 * i + 2], vb[4 * i + 3]);
                }
            }
            if (a.has_next) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) f16::split(__fmul2_rn(va[i], s16), hi[i], lo[i]);
                f16::tmem_st16u(yr, hi);
                if (np != 1) f16::tmem_st16u(yr + 16, lo);
#pragma unroll
                for (int i = 0; i < 16; ++i) f16::split(__fmul2_rn(vb[i], s16), hi[i], lo[i]);
                f16::tmem_st16u(yr + 32, hi);
                if (np != 1) f16::tmem_st16u(yr + 48, lo);
                tc::tmem_wait_st();
                tc::tc_fence_before();
                warp_arrive(&a_full[g], lane);
                // ---- the two projection halves, 128 columns each, through the X region ----
                for (int half = 0; half < 2; ++half) {
                    tc::mbar_wait(&d_full[g], uses & 1u);
                    ++uses;
                    tc::tc_fence_after();
                    float* o = a.P + (size_t)rr * 256 + half * 128;
#pragma unroll 1
                    for (int c = 0; c < 128; c += 32) {
                        float2 v[16];
                        tc::tmem_ld16(xr + c, v);
                        tc::tmem_ld16(xr + c + 16, v + 8);
                        tc::tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            v[i] = half == 0 ? __ffma2_rn(v[i], us, *reinterpret_cast<const float2*>(cst + 72 + c + 2 * i)) : __fmul2_rn(v[i], us);
                        if (valid) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) stg256(o + c + 8 * i, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                        }
                    }
                    if (half == 0) {        // the X region has been read: the source half may overwrite it
                        tc::tc_fence_before();
                        warp_arrive(&a_full[g], lane);
                    }
                }
            }
            tc::tc_fence_before();
        }
    } else if (warp == 4 * NGROUPS) {
        if (lane == 0) {
            constexpr uint32_t ID64 = f16::idesc(TM, 64), ID128 = f16::idesc(TM, 128);
            const uint32_t sWU = tc::smem_u32(smem_u + OFF_WU), sWPT = tc::smem_u32(smem_u + OFF_WPT), sWPS = tc::smem_u32(smem_u + OFF_WPS);
            int tile_j[NGROUPS], step[NGROUPS];
            uint32_t uses[NGROUPS];
            int remaining = 0;
            for (int i = 0; i < NGROUPS; ++i) {
                tile_j[i] = i; step[i] = 0; uses[i] = 0;
                if (i < my_tiles) ++remaining;
            }
            while (remaining > 0) {
                bool did = false;
#pragma unroll
                for (int i = 0; i < NGROUPS; ++i) {
                    if (tile_j[i] >= my_tiles) continue;
                    if (!f16::mbar_test(&a_full[i], uses[i] & 1u)) continue;
                    tc::tc_fence_after();
                    const uint32_t xr = tmem + (uint32_t)i * 256, yr = xr + 128;
                    if (step[i] == 0) f16::gemm_ts<128, 64>(yr, xr, sWU, ID64, false, np);            // update layer: Y = cat(x, agg) W_upd^T
                    else if (step[i] == 1) f16::gemm_ts<64, 128>(xr, yr, sWPT, ID128, false, np);      // target half of the projection
                    else f16::gemm_ts<64, 128>(xr, yr, sWPS, ID128, false, np);                        // source half
                    tc::mma_commit(&d_full[i]);
                    ++uses[i];
                    if (++step[i] == n_mma) {
                        step[i] = 0;
                        tile_j[i] += NGROUPS;
                        if (tile_j[i] >= my_tiles) --remaining;
                    }
                    did = true;
                }
                if (!did) __nanosleep(20);
            }
        }
        __syncwarp();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// images of the hoisted projection of a block (node columns of msg.0): target half | source half, each K = 64 x N = 128 [hi | lo]
size_t conv_proj_f16_floats(const ConvDims& d) { return (d.cn == 64 && d.h == 128) ? (size_t)2 * cnn::WP_WORDS : 0; }

int conv_proj_f16_pack(const rgnn_conv& c, const ConvDims& d, float* dst, cudaStream_t stream) {
    if (conv_proj_f16_floats(d) == 0) return RGNN_OK;
    const rgnn_linear& m0 = c.msg.layer[0];
    uint32_t* w = reinterpret_cast<uint32_t*>(dst);
    int rc = pack_f16_image(m0.weight, m0.in_features, d.cn, d.h, d.h, d.cn, w, stream);
    if (rc) return rc;
    return pack_f16_image(m0.weight + d.cn, m0.in_features, d.cn, d.h, d.h, d.cn, w + cnn::WP_WORDS, stream);
}

bool conv_nodes_f16_supported(const rgnn_conv& c, const ConvDims& d) {
    const rgnn_linear& L = c.upd.layer[0];
    return g_chain_f16 && c.upd.n == 1 && d.cn == 64 && d.h == 128 && L.in_features == 128 && L.out_features == 64 && L.weight_t != nullptr;
}

int run_conv_nodes_f16(const rgnn_conv& c, const ConvDims& d, int n_nodes, const float* x, const float* agg, float* out,
                       const rgnn_conv* next, const float* next_proj_images, float* P_next, cudaStream_t stream, float* u_save, float* sd_save) {
    if (n_nodes <= 0) return RGNN_OK;
    (void)d;
    const rgnn_linear& L = c.upd.layer[0];
    ConvNodesArgs a;
    memset(&a, 0, sizeof(a));
    a.n_rows = n_nodes; a.x = x; a.agg = agg;
    a.w_upd = reinterpret_cast<const uint32_t*>(f16_weights(L));
    a.b_upd = L.bias; a.s_upd = L.norm_scale; a.m_upd = L.norm_shift; a.act = L.activation;
    a.out = out;
    a.has_next = next != nullptr;
    if (next != nullptr) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(next_proj_images);
        a.w_pt = w; a.w_ps = w + cnn::WP_WORDS;
        a.b_p = next->msg.layer[0].bias;
        a.P = P_next;
    }
    a.u_save = u_save; a.sd_save = sd_save;
    a.passes = mp_f16_passes();
    a.prefetch = g_rows_prefetch;
    static PerDeviceOnce once;
    if (once.needed()) {
        RGNN_CHECK_CUDA(cudaFuncSetAttribute(conv_nodes_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cnn::SMEM));
        once.mark();
    }
    const int n_tiles = (n_nodes + cnn::TM - 1) / cnn::TM;
    const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
    conv_nodes_f16_kernel<<<grid, cnn::NTHREADS, cnn::SMEM, stream>>>(a);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

}  // namespace rgnn
