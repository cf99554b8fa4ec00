// Forward tile programs: one persistent kernel interprets a short list of steps per tile of TR rows.
// Every reference forward block (ffn stacks, the message-passing layer, the four heads) is a Program.
#include "rgnn_tile.cuh"
#include "rgnn_programs.h"

namespace rgnn {

constexpr int TRF = 64;   // rows per tile in the forward kernels

// ---------------------------------------------------------------------------------------------
// step implementations (each ends with __syncthreads)
// ---------------------------------------------------------------------------------------------
template <int TR>
__device__ __forceinline__ void op_load_rows(float* cur, const Step& st, int row0, int nvalid) {
    const float* __restrict__ src = static_cast<const float*>(st.p0);
    const int* __restrict__ ridx = static_cast<const int*>(st.p1);
    const int ld = st.i0, w = st.i1, dcol = st.i2, padto = st.i3;
    if (((w | ld | dcol | padto) & 3) == 0) {
        const int p4 = padto >> 2, w4 = w >> 2;
        for (int i = threadIdx.x; i < TR * p4; i += NT) {
            const int r = i / p4, j4 = i - r * p4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nvalid && j4 < w4) {
                const size_t row = ridx ? (size_t)__ldg(ridx + row0 + r) : (size_t)(row0 + r);
                v = __ldg(reinterpret_cast<const float4*>(src + row * ld) + j4);
            }
            *reinterpret_cast<float4*>(cur + r * LD + dcol + 4 * j4) = v;
        }
    } else {
        for (int i = threadIdx.x; i < TR * padto; i += NT) {
            const int r = i / padto, j = i - r * padto;
            float v = 0.f;
            if (r < nvalid && j < w) {
                const size_t row = ridx ? (size_t)__ldg(ridx + row0 + r) : (size_t)(row0 + r);
                v = __ldg(src + row * ld + j);
            }
            cur[r * LD + dcol + j] = v;
        }
    }
    __syncthreads();
}

template <int TR>
__device__ __forceinline__ void op_load_pairsum(float* cur, const Step& st, int row0, int nvalid) {
    const float* __restrict__ h = static_cast<const float*>(st.p0);
    const int* __restrict__ ia = static_cast<const int*>(st.p1);
    const int* __restrict__ ib = static_cast<const int*>(st.p2);
    const int ld = st.i0, w4 = st.i1 >> 2;
    for (int i = threadIdx.x; i < TR * w4; i += NT) {
        const int r = i / w4, j4 = i - r * w4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nvalid) {
            const size_t a = (size_t)__ldg(ia + row0 + r), b = (size_t)__ldg(ib + row0 + r);
            const float4 va = __ldg(reinterpret_cast<const float4*>(h + a * ld) + j4);
            const float4 vb = __ldg(reinterpret_cast<const float4*>(h + b * ld) + j4);
            v = make_float4(va.x + vb.x, va.y + vb.y, va.z + vb.z, va.w + vb.w);
        }
        *reinterpret_cast<float4*>(cur + r * LD + 4 * j4) = v;
    }
    __syncthreads();
}

template <int TR>
__device__ __forceinline__ void op_load_segmax(float* cur, const Step& st, int row0, int nvalid) {
    const float* __restrict__ g = static_cast<const float*>(st.p0);
    const int* __restrict__ ptr = static_cast<const int*>(st.p1);
    const int* __restrict__ mem = static_cast<const int*>(st.p2);
    const int ld = st.i0, w = st.i1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < TR; r += NT / 32) {
        int m0 = 0, m1 = 0;
        if (r < nvalid) {
            m0 = __ldg(ptr + row0 + r);
            m1 = __ldg(ptr + row0 + r + 1);
        }
        for (int j = lane; j < w; j += 32) {
            float v = (m1 > m0) ? -INFINITY : 0.f;
            for (int m = m0; m < m1; ++m) v = fmaxf(v, __ldg(g + (size_t)__ldg(mem + m) * ld + j));
            cur[r * LD + j] = v;
        }
    }
    __syncthreads();
}

template <int TR>
__device__ __forceinline__ void op_add_gather2(float* cur, const Step& st, int row0, int nvalid) {
    const float* __restrict__ P = static_cast<const float*>(st.p0);
    const int* __restrict__ it = static_cast<const int*>(st.p1);
    const int* __restrict__ is = static_cast<const int*>(st.p2);
    const int ld = st.i0, w4 = st.i1 >> 2, off = st.i2;
    for (int i = threadIdx.x; i < TR * w4; i += NT) {
        const int r = i / w4, j4 = i - r * w4;
        if (r < nvalid) {
            const size_t t = (size_t)__ldg(it + row0 + r), s = (size_t)__ldg(is + row0 + r);
            const float4 a = __ldg(reinterpret_cast<const float4*>(P + t * ld) + j4);
            const float4 b = __ldg(reinterpret_cast<const float4*>(P + s * ld + off) + j4);
            float4* c = reinterpret_cast<float4*>(cur + r * LD + 4 * j4);
            float4 v = *c;
            // association order of the reference's single dot product is not reproducible anyway;
            // keep (W1c e + b) + (Pt + Ps)
            v.x += a.x + b.x; v.y += a.y + b.y; v.z += a.z + b.z; v.w += a.w + b.w;
            *c = v;
        }
    }
    __syncthreads();
}

template <int TR>
__device__ __forceinline__ void op_add_rows(float* cur, const Step& st, int row0, int nvalid) {
    const float* __restrict__ src = static_cast<const float*>(st.p0);
    const int ld = st.i0, w4 = st.i1 >> 2;
    for (int i = threadIdx.x; i < TR * w4; i += NT) {
        const int r = i / w4, j4 = i - r * w4;
        if (r < nvalid) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src + (size_t)(row0 + r) * ld) + j4);
            float4* c = reinterpret_cast<float4*>(cur + r * LD + 4 * j4);
            float4 v = *c;
            v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
            *c = v;
        }
    }
    __syncthreads();
}

template <int TR>
__device__ __forceinline__ void op_store_rows(const float* cur, const Step& st, int row0, int nvalid) {
    float* __restrict__ dst = static_cast<float*>(const_cast<void*>(st.p0));
    const int ld = st.i0, w = st.i1, dcol = st.i2;
    if (((w | ld | dcol) & 3) == 0) {
        const int w4 = w >> 2;
        for (int i = threadIdx.x; i < TR * w4; i += NT) {
            const int r = i / w4, j4 = i - r * w4;
            if (r < nvalid)
                *reinterpret_cast<float4*>(dst + (size_t)(row0 + r) * ld + dcol + 4 * j4) =
                    *reinterpret_cast<const float4*>(cur + r * LD + 4 * j4);
        }
    } else {
        for (int i = threadIdx.x; i < TR * w; i += NT) {
            const int r = i / w, j = i - r * w;
            if (r < nvalid) dst[(size_t)(row0 + r) * ld + dcol + j] = cur[r * LD + j];
        }
    }
    __syncthreads();
}

// Sum consecutive rows with equal target id.  A target whose whole CSR row lies inside this tile gets a
// plain store (deterministic, and in the same source-ascending order as the reference's index_add_);
// a row cut by a tile boundary is completed with atomicAdd onto the zero-initialised output.
template <int TR>
__device__ __forceinline__ void op_segsum(const float* cur, const Step& st, int row0, int nvalid, int* ibuf) {
    float* __restrict__ agg = static_cast<float*>(const_cast<void*>(st.p0));
    const int* __restrict__ tgt = static_cast<const int*>(st.p1);
    const int* __restrict__ row_ptr = static_cast<const int*>(st.p2);
    const int ld = st.i0, w = st.i1;
    for (int r = threadIdx.x; r < TR; r += NT) ibuf[r] = r < nvalid ? __ldg(tgt + row0 + r) : -1;
    __syncthreads();
    for (int j = threadIdx.x; j < w; j += NT) {
        int r = 0;
        while (r < nvalid) {
            const int t = ibuf[r];
            float s = 0.f;
            int r1 = r;
            while (r1 < nvalid && ibuf[r1] == t) {
                s += cur[r1 * LD + j];
                ++r1;
            }
            const bool whole = (__ldg(row_ptr + t) == row0 + r) && (__ldg(row_ptr + t + 1) == row0 + r1);
            float* o = agg + (size_t)t * ld + j;
            if (whole) *o = s; else atomicAdd(o, s);
            r = r1;
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// interpreter
// ---------------------------------------------------------------------------------------------
template <int TR>
__global__ void __launch_bounds__(NT, 1) chain_fwd_kernel(const __grid_constant__ Program prog) {
    extern __shared__ __align__(16) float smem[];
    float* buf0 = smem;
    float* buf1 = buf0 + TR * LD;
    float* wst = buf1 + TR * LD;
    int* ibuf = reinterpret_cast<int*>(wst + 2 * KC * CBMAX);

    const int n_tiles = (prog.n_rows + TR - 1) / TR;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = tile * TR;
        const int nvalid = min(TR, prog.n_rows - row0);
        float* cur = buf0;
        float* nxt = buf1;
        for (int s = 0; s < prog.n_steps; ++s) {
            const Step& st = prog.steps[s];
            switch (st.op) {
                case OP_LOAD_ROWS: op_load_rows<TR>(cur, st, row0, nvalid); break;
                case OP_LOAD_PAIRSUM: op_load_pairsum<TR>(cur, st, row0, nvalid); break;
                case OP_LOAD_SEGMAX: op_load_segmax<TR>(cur, st, row0, nvalid); break;
                case OP_LINEAR: {
                    tile_gemm<TR>(cur, st.i0, static_cast<const float*>(st.p0), st.i2, st.i2,
                                  static_cast<const float*>(st.p1), st.i1, nxt, wst);
                    float* t = cur; cur = nxt; nxt = t;
                } break;
                case OP_ADD_GATHER2: op_add_gather2<TR>(cur, st, row0, nvalid); break;
                case OP_NORM_ACT:
                    tile_norm_act<TR>(cur, st.i0, static_cast<const float*>(st.p0), static_cast<const float*>(st.p1),
                                      st.i1 != 0, nullptr);
                    break;
                case OP_ADD_ROWS: op_add_rows<TR>(cur, st, row0, nvalid); break;
                case OP_STORE_ROWS: op_store_rows<TR>(cur, st, row0, nvalid); break;
                case OP_SEGSUM: op_segsum<TR>(cur, st, row0, nvalid, ibuf); break;
                default: break;
            }
        }
    }
}

constexpr size_t fwd_smem_bytes() { return (size_t)(2 * TRF * LD + 2 * KC * CBMAX) * sizeof(float) + TRF * sizeof(int); }

int launch_fwd(const Program& p, cudaStream_t stream) {
    if (p.n_rows <= 0) return RGNN_OK;
    static bool configured = false;
    if (!configured) {
        RGNN_CHECK_CUDA(cudaFuncSetAttribute(chain_fwd_kernel<TRF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)fwd_smem_bytes()));
        configured = true;
    }
    const int n_tiles = (p.n_rows + TRF - 1) / TRF;
    const int grid = min(n_tiles, sm_count());
    chain_fwd_kernel<TRF><<<grid, NT, fwd_smem_bytes(), stream>>>(p);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

// ---------------------------------------------------------------------------------------------
// program builder
// ---------------------------------------------------------------------------------------------
bool ProgBuilder::add(int op, int i0, int i1, int i2, int i3, const void* p0, const void* p1, const void* p2,
                      const void* p3) {
    if (p.n_steps >= MAX_STEPS) {
        set_error("tile program too long (%d steps)", p.n_steps);
        ok = false;
        return false;
    }
    Step& s = p.steps[p.n_steps++];
    s.op = op; s.i0 = i0; s.i1 = i1; s.i2 = i2; s.i3 = i3;
    s.p0 = p0; s.p1 = p1; s.p2 = p2; s.p3 = p3;
    return true;
}

bool check_linear(const rgnn_linear& L) {
    if (L.in_features <= 0 || L.in_features > 256 || L.out_features <= 0 || L.out_features > 256) {
        set_error("linear %dx%d outside the supported widths (<=256)", L.out_features, L.in_features);
        return false;
    }
    if (L.norm_scale != nullptr && (L.out_features % 32 != 0 || L.out_features < 2)) {
        set_error("channel_normalization needs out_features %% 32 == 0 (got %d)", L.out_features);
        return false;
    }
    if (L.weight_t == nullptr) {
        set_error("linear without packed weight (call rgnn_pack_*)");
        return false;
    }
    return true;
}

void ProgBuilder::linear(const rgnn_linear& L) {
    if (!check_linear(L)) { ok = false; return; }
    add(OP_LINEAR, round_up(L.in_features, 8), L.out_features, round_up(L.out_features, 64), 0, L.weight_t, L.bias);
    if (L.norm_scale != nullptr || L.activation)
        add(OP_NORM_ACT, L.out_features, L.activation, 0, 0, L.norm_scale, L.norm_shift);
}

void ProgBuilder::stack(const rgnn_stack& s, int first, int last) {
    if (last < 0) last = s.n;
    for (int i = first; i < last; ++i) linear(s.layer[i]);
}

void ProgBuilder::load_rows(const float* src, int ld, int w, int dcol, int padto, const int* ridx) {
    add(OP_LOAD_ROWS, ld, w, dcol, padto < w ? w : padto, src, ridx);
}
void ProgBuilder::store_rows(float* dst, int ld, int w, int dcol) { add(OP_STORE_ROWS, ld, w, dcol, 0, dst); }

}  // namespace rgnn
