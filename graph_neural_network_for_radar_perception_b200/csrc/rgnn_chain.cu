// Tile programs: one persistent kernel interprets a short list of steps per tile of TR rows.
// Every forward block of the reference (ffn stacks, the message-passing layer, the four heads) and its
// backward (which first RECOMPUTES the tile's forward activations into shared memory, so per-edge
// activations are never stored in HBM) is a Program built on the host (rgnn_model*.cu).
#include "rgnn_tile.cuh"
#include "rgnn_programs.h"

namespace rgnn {

struct Tile {
    int row0, nvalid;
    float* smem;
    const Program* pg;          // lives in the kernel parameter space (__grid_constant__)
    __device__ __forceinline__ float* reg(int i) const { return smem + pg->reg_off[i]; }
    __device__ __forceinline__ int ld(int i) const { return pg->reg_ld[i]; }
    float* sigma;   // [MAX_SIGMA][TR]
    float* wst;     // [2][KC][CBMAX]
    int* ibuf;      // [2*TR]
    float* red;     // [32]
    double* dacc;   // [2*MAX_STEPS] per-CTA running sums of the scalar norm-parameter gradients
    float* wacc;    // per-CTA weight / bias gradient accumulators of the OP_WGRAD steps that fit (Step::i5 = offset)
};

// ---------------------------------------------------------------------------------------------
// loads
// ---------------------------------------------------------------------------------------------
template <int TR>
__device__ __forceinline__ void op_load_rows(const Tile& t, const Step& st) {
    const float* __restrict__ src = static_cast<const float*>(st.p0);
    const int* __restrict__ ridx = static_cast<const int*>(st.p1);
    float* buf = t.reg(st.ra);
    const int ldb = t.ld(st.ra);
    const int ld = st.i0, w = st.i1, dcol = st.i2, padto = st.i3, scol = st.i4;
    if (((w | ld | dcol | padto | scol) & 3) == 0) {
        const int p4 = padto >> 2, w4 = w >> 2;
        for (int i = threadIdx.x; i < TR * p4; i += NT) {
            const int r = i / p4, j4 = i - r * p4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < t.nvalid && j4 < w4) {
                const size_t row = ridx ? (size_t)__ldg(ridx + t.row0 + r) : (size_t)(t.row0 + r);
                v = __ldg(reinterpret_cast<const float4*>(src + row * ld + scol) + j4);
            }
            *reinterpret_cast<float4*>(buf + r * ldb + dcol + 4 * j4) = v;
        }
    } else {
        for (int i = threadIdx.x; i < TR * padto; i += NT) {
            const int r = i / padto, j = i - r * padto;
            float v = 0.f;
            if (r < t.nvalid && j < w) {
                const size_t row = ridx ? (size_t)__ldg(ridx + t.row0 + r) : (size_t)(t.row0 + r);
                v = __ldg(src + row * ld + scol + j);
            }
            buf[r * ldb + dcol + j] = v;
        }
    }
    __syncthreads();
}

template <int TR>
__device__ __forceinline__ void op_load_pairsum(const Tile& t, const Step& st) {
    const float* __restrict__ h = static_cast<const float*>(st.p0);
    const int* __restrict__ ia = static_cast<const int*>(st.p1);
    const int* __restrict__ ib = static_cast<const int*>(st.p2);
    float* buf = t.reg(st.ra);
    const int ldb = t.ld(st.ra);
    const int ld = st.i0, w4 = st.i1 >> 2;
    for (int i = threadIdx.x; i < TR * w4; i += NT) {
        const int r = i / w4, j4 = i - r * w4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < t.nvalid) {
            const size_t a = (size_t)__ldg(ia + t.row0 + r), b = (size_t)__ldg(ib + t.row0 + r);
            const float4 va = __ldg(reinterpret_cast<const float4*>(h + a * ld) + j4);
            const float4 vb = __ldg(reinterpret_cast<const float4*>(h + b * ld) + j4);
            v = make_float4(va.x + vb.x, va.y + vb.y, va.z + vb.z, va.w + vb.w);
        }
        *reinterpret_cast<float4*>(buf + r * ldb + 4 * j4) = v;
    }
    __syncthreads();
}

template <int TR>
__device__ __forceinline__ void op_load_segmax(const Tile& t, const Step& st) {
    const float* __restrict__ g = static_cast<const float*>(st.p0);
    const int* __restrict__ ptr = static_cast<const int*>(st.p1);
    const int* __restrict__ mem = static_cast<const int*>(st.p2);
    float* buf = t.reg(st.ra);
    const int ldb = t.ld(st.ra);
    const int ld = st.i0, w = st.i1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < TR; r += NT / 32) {
        int m0 = 0, m1 = 0;
        if (r < t.nvalid) {
            m0 = __ldg(ptr + t.row0 + r);
            m1 = __ldg(ptr + t.row0 + r + 1);
        }
        for (int j = lane; j < w; j += 32) {
            float v = (m1 > m0) ? -INFINITY : 0.f;
            for (int m = m0; m < m1; ++m) v = fmaxf(v, __ldg(g + (size_t)__ldg(mem + m) * ld + j));
            buf[r * ldb + j] = v;
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// element-wise tile ops
// ---------------------------------------------------------------------------------------------
template <int TR>
__device__ __forceinline__ void op_add_gather2(const Tile& t, const Step& st) {
    const float* __restrict__ P = static_cast<const float*>(st.p0);
    const int* __restrict__ it = static_cast<const int*>(st.p1);
    const int* __restrict__ is = static_cast<const int*>(st.p2);
    float* buf = t.reg(st.ra);
    const int ldb = t.ld(st.ra);
    const int ld = st.i0, w4 = st.i1 >> 2, off = st.i2;
    for (int i = threadIdx.x; i < TR * w4; i += NT) {
        const int r = i / w4, j4 = i - r * w4;
        if (r < t.nvalid) {
            const size_t tn = (size_t)__ldg(it + t.row0 + r), sn = (size_t)__ldg(is + t.row0 + r);
            const float4 a = __ldg(reinterpret_cast<const float4*>(P + tn * ld) + j4);
            const float4 b = __ldg(reinterpret_cast<const float4*>(P + sn * ld + off) + j4);
            float4* c = reinterpret_cast<float4*>(buf + r * ldb + 4 * j4);
            float4 v = *c;
            v.x += a.x + b.x; v.y += a.y + b.y; v.z += a.z + b.z; v.w += a.w + b.w;
            *c = v;
        }
    }
    __syncthreads();
}

template <int TR>
__device__ __forceinline__ void op_add_rows(const Tile& t, const Step& st) {
    const float* __restrict__ src = static_cast<const float*>(st.p0);
    float* buf = t.reg(st.ra);
    const int ldb = t.ld(st.ra);
    const int ld = st.i0, w4 = st.i1 >> 2;
    for (int i = threadIdx.x; i < TR * w4; i += NT) {
        const int r = i / w4, j4 = i - r * w4;
        if (r < t.nvalid) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src + (size_t)(t.row0 + r) * ld) + j4);
            float4* c = reinterpret_cast<float4*>(buf + r * ldb + 4 * j4);
            float4 v = *c;
            v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
            *c = v;
        }
    }
    __syncthreads();
}

template <int TR>
__device__ __forceinline__ void op_add_region(const Tile& t, const Step& st) {
    float* a = t.reg(st.ra);
    const float* b = t.reg(st.rb);
    const int lda = t.ld(st.ra), ldb = t.ld(st.rb);
    const int w4 = st.i1 >> 2, ca = st.i2, cb = st.i3;
    for (int i = threadIdx.x; i < TR * w4; i += NT) {
        const int r = i / w4, j4 = i - r * w4;
        float4* pa = reinterpret_cast<float4*>(a + r * lda + ca + 4 * j4);
        const float4 vb = *reinterpret_cast<const float4*>(b + r * ldb + cb + 4 * j4);
        float4 v = *pa;
        v.x += vb.x; v.y += vb.y; v.z += vb.z; v.w += vb.w;
        *pa = v;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// stores / reductions to global memory
// ---------------------------------------------------------------------------------------------
template <int TR>
__device__ __forceinline__ void op_store_rows(const Tile& t, const Step& st) {
    float* __restrict__ dst = static_cast<float*>(const_cast<void*>(st.p0));
    const float* buf = t.reg(st.ra);
    const int ldb = t.ld(st.ra);
    const int ld = st.i0, w = st.i1, dcol = st.i2, scol = st.i4;
    const bool accum = st.i3 != 0;
    if (((w | ld | dcol | scol) & 3) == 0) {
        const int w4 = w >> 2;
        for (int i = threadIdx.x; i < TR * w4; i += NT) {
            const int r = i / w4, j4 = i - r * w4;
            if (r < t.nvalid) {
                float4* o = reinterpret_cast<float4*>(dst + (size_t)(t.row0 + r) * ld + dcol + 4 * j4);
                float4 v = *reinterpret_cast<const float4*>(buf + r * ldb + scol + 4 * j4);
                if (accum) {
                    const float4 u = *o;
                    v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
                }
                *o = v;
            }
        }
    } else {
        for (int i = threadIdx.x; i < TR * w; i += NT) {
            const int r = i / w, j = i - r * w;
            if (r < t.nvalid) {
                float* o = dst + (size_t)(t.row0 + r) * ld + dcol + j;
                const float v = buf[r * ldb + scol + j];
                *o = accum ? *o + v : v;
            }
        }
    }
    __syncthreads();
}

// Sum consecutive rows with equal target id.  A target whose whole CSR row lies inside this tile gets a
// plain store (deterministic, and in the same source-ascending order as the reference's index_add_);
// a row cut by a tile boundary is completed with atomicAdd onto the zero-initialised output.
template <int TR>
__device__ __forceinline__ void op_segsum(const Tile& t, const Step& st) {
    float* __restrict__ agg = static_cast<float*>(const_cast<void*>(st.p0));
    const int* __restrict__ tgt = static_cast<const int*>(st.p1);
    const int* __restrict__ row_ptr = static_cast<const int*>(st.p2);
    const float* buf = t.reg(st.ra);
    const int ldb = t.ld(st.ra);
    const int ld = st.i0, w = st.i1, dcol = st.i2, scol = st.i4;
    for (int r = threadIdx.x; r < TR; r += NT) t.ibuf[r] = r < t.nvalid ? __ldg(tgt + t.row0 + r) : -1;
    __syncthreads();
    for (int j = threadIdx.x; j < w; j += NT) {
        int r = 0;
        while (r < t.nvalid) {
            const int tn = t.ibuf[r];
            float s = 0.f;
            int r1 = r;
            while (r1 < t.nvalid && t.ibuf[r1] == tn) {
                s += buf[r1 * ldb + scol + j];
                ++r1;
            }
            const bool whole = (__ldg(row_ptr + tn) == t.row0 + r) && (__ldg(row_ptr + tn + 1) == t.row0 + r1);
            float* o = agg + (size_t)tn * ld + dcol + j;
            if (whole) *o = s; else atomicAdd(o, s);
            r = r1;
        }
    }
    __syncthreads();
}

template <int TR>
__device__ __forceinline__ void op_scatter_add(const Tile& t, const Step& st) {
    float* __restrict__ dst = static_cast<float*>(const_cast<void*>(st.p0));
    const int* __restrict__ idx = static_cast<const int*>(st.p1);
    const float* buf = t.reg(st.ra);
    const int ldb = t.ld(st.ra);
    const int ld = st.i0, w4 = st.i1 >> 2, dcol = st.i2, scol = st.i4;
    for (int i = threadIdx.x; i < TR * w4; i += NT) {
        const int r = i / w4, j4 = i - r * w4;
        if (r < t.nvalid) {
            const size_t n = (size_t)__ldg(idx + t.row0 + r);
            atomicAdd(reinterpret_cast<float4*>(dst + n * ld + dcol) + j4,
                      *reinterpret_cast<const float4*>(buf + r * ldb + scol + 4 * j4));
        }
    }
    __syncthreads();
}

template <int TR>
__device__ __forceinline__ void op_pair_scatter(const Tile& t, const Step& st) {
    float* __restrict__ dst = static_cast<float*>(const_cast<void*>(st.p0));
    const int* __restrict__ ia = static_cast<const int*>(st.p1);
    const int* __restrict__ ib = static_cast<const int*>(st.p2);
    const float* buf = t.reg(st.ra);
    const int ldb = t.ld(st.ra);
    const int ld = st.i0, w4 = st.i1 >> 2;
    for (int i = threadIdx.x; i < TR * w4; i += NT) {
        const int r = i / w4, j4 = i - r * w4;
        if (r < t.nvalid) {
            const float4 v = *reinterpret_cast<const float4*>(buf + r * ldb + 4 * j4);
            atomicAdd(reinterpret_cast<float4*>(dst + (size_t)__ldg(ia + t.row0 + r) * ld) + j4, v);
            atomicAdd(reinterpret_cast<float4*>(dst + (size_t)__ldg(ib + t.row0 + r) * ld) + j4, v);
        }
    }
    __syncthreads();
}

// torch.max(x[idx], dim=0) backward: the gradient of a pooled column goes to the (first) arg-max member.
template <int TR>
__device__ __forceinline__ void op_segmax_bwd(const Tile& t, const Step& st) {
    float* __restrict__ dst = static_cast<float*>(const_cast<void*>(st.p0));
    const int* __restrict__ ptr = static_cast<const int*>(st.p1);
    const int* __restrict__ mem = static_cast<const int*>(st.p2);
    const float* __restrict__ g = static_cast<const float*>(st.p3);
    const float* buf = t.reg(st.ra);
    const int ldb = t.ld(st.ra);
    const int ld = st.i0, w = st.i1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < t.nvalid; r += NT / 32) {
        const int m0 = __ldg(ptr + t.row0 + r), m1 = __ldg(ptr + t.row0 + r + 1);
        for (int j = lane; j < w; j += 32) {
            float best = -INFINITY;
            int arg = -1;
            for (int m = m0; m < m1; ++m) {
                const int node = __ldg(mem + m);
                const float v = __ldg(g + (size_t)node * ld + j);
                if (v > best) { best = v; arg = node; }
            }
            if (arg >= 0) atomicAdd(dst + (size_t)arg * ld + j, buf[r * ldb + j]);
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// backward of LeakyReLU + channel_normalization, in place on the gradient tile
//   y = leaky(s * n + m), n = (z - mean) / (sigma + eps)
//   dz_j = (dn_j - mean(dn)) / (sigma + eps) - n_j * sum_i(dn_i n_i) / ((C-1) sigma),  dn = s * g
// n is recovered from the saved activation y (n = (leaky^-1(y) - m) / s).  A constant row (sigma = 0) gives
// a zero second term here, where torch autograd produces NaN (documented divergence).
// ---------------------------------------------------------------------------------------------
template <int TR>
__device__ __forceinline__ void op_actnorm_bwd(const Tile& t, const Step& st, int step_idx) {
    float* gbuf = t.reg(st.ra);
    const float* ybuf = t.reg(st.rb);
    const int ldg_ = t.ld(st.ra), ldy = t.ld(st.rb);
    const int C = st.i0, slot = st.i2;
    const bool act = st.i1 != 0;
    const float* scale_p = static_cast<const float*>(st.p0);
    const float* shift_p = static_cast<const float*>(st.p1);
    const bool want_scalar_grads = st.p2 != nullptr;
    const bool has_norm = scale_p != nullptr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float scale = 1.f, shift = 0.f, inv_scale = 0.f;
    if (has_norm) {
        scale = __ldg(scale_p);
        shift = __ldg(shift_p);
        inv_scale = scale != 0.f ? 1.f / scale : 0.f;
    }
    const float* sig = t.sigma + slot * TR;
    float acc_s = 0.f, acc_m = 0.f;
    const int nq = (C + 31) >> 5;
    for (int r = warp; r < TR; r += NT / 32) {
        float g[8], n[8];
        float sum_dn = 0.f, dot = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int c = lane + 32 * q;
            g[q] = 0.f; n[q] = 0.f;
            if (q < nq && c < C) {
                float gv = gbuf[r * ldg_ + c];
                const float yv = ybuf[r * ldy + c];
                if (act && !(yv > 0.f)) gv *= LEAKY;
                if (has_norm) {
                    const float ypre = (act && !(yv > 0.f)) ? yv / LEAKY : yv;
                    const float nv = (ypre - shift) * inv_scale;
                    acc_s = fmaf(gv, nv, acc_s);
                    acc_m += gv;
                    gv *= scale;              // dn
                    n[q] = nv;
                    sum_dn += gv;
                    dot = fmaf(gv, nv, dot);
                }
                g[q] = gv;
            }
        }
        if (has_norm) {
            sum_dn = warp_sum(sum_dn);
            dot = warp_sum(dot);
            const float sd = sig[r];
            const float inv_den = 1.f / (sd + NORM_EPS);
            const float mean_dn = sum_dn / (float)C;
            const float coef = sd > 0.f ? dot / ((float)(C - 1) * sd) : 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) g[q] = (g[q] - mean_dn) * inv_den - n[q] * coef;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int c = lane + 32 * q;
            if (q < nq && c < C) gbuf[r * ldg_ + c] = g[q];
        }
    }
    if (has_norm && want_scalar_grads) {
        acc_s = warp_sum(acc_s);
        acc_m = warp_sum(acc_m);
        if (lane == 0) { t.red[warp] = acc_s; t.red[8 + warp] = acc_m; }
        __syncthreads();
        if (threadIdx.x == 0) {
            // these two scalars are sums over every element of the layer output (heavy cancellation): keep the
            // CTA's running total in double and publish it once, at the end of the kernel
            double a = 0., b = 0.;
            for (int w = 0; w < NT / 32; ++w) { a += (double)t.red[w]; b += (double)t.red[8 + w]; }
            t.dacc[2 * step_idx] += a;
            t.dacc[2 * step_idx + 1] += b;
        }
    }
    __syncthreads();
}

// dW[c][i3 + k] += sum_r dZ[r][i4 + c] X[r][k]   (c < i0, k < i1), db[c] += sum_r dZ[r][i4 + c].
// Work item = 4 output rows (c) x 8 columns (k).  When the step owns a shared-memory accumulator (i5 >= 0; assigned
// by launch_program while they fit) the tile sums are added there -- every element is always updated by the same
// thread, so no atomics are needed -- and the CTA publishes its totals once, at the end of the kernel; one RED per
// element and TILE measured ~4x the cost of the wgrad arithmetic itself.  Otherwise the tile sum goes to global
// memory with one RED per element.
template <int TR>
__device__ __forceinline__ void op_wgrad(const Tile& t, const Step& st) {
    const float* dz = t.reg(st.ra);
    const float* xs = t.reg(st.rb);
    const int ldz = t.ld(st.ra), ldx = t.ld(st.rb);
    const int C = st.i0, K = st.i1, ldW = st.i2, wcol = st.i3, zcol = st.i4;
    float* __restrict__ dW = static_cast<float*>(const_cast<void*>(st.p0));
    float* __restrict__ db = static_cast<float*>(const_cast<void*>(st.p1));
    float* __restrict__ wacc = st.i5 >= 0 ? t.wacc + st.i5 : nullptr;       // [C][K] then [C]
    if (dW != nullptr) {
        const int kg = (K + 7) >> 3, cg = (C + 3) >> 2;
        for (int item = threadIdx.x; item < kg * cg; item += NT) {
            const int kq = item % kg, cq = item / kg;
            float acc[4][8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
            const float* zp = dz + zcol + 4 * cq;
            const float* xp = xs + 8 * kq;
#pragma unroll 4
            for (int r = 0; r < TR; ++r) {
                const float4 z = *reinterpret_cast<const float4*>(zp + r * ldz);
                const float4 x0 = *reinterpret_cast<const float4*>(xp + r * ldx);
                const float4 x1 = *reinterpret_cast<const float4*>(xp + r * ldx + 4);
                const float zz[4] = {z.x, z.y, z.z, z.w};
                const float xx[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(zz[i], xx[j], acc[i][j]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = 4 * cq + i;
                if (c < C) {
                    if (wacc != nullptr) {
                        float* o = wacc + c * K + 8 * kq;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (8 * kq + j < K) o[j] += acc[i][j];
                    } else {
                        float* o = dW + (size_t)c * ldW + wcol + 8 * kq;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (8 * kq + j < K) atomicAdd(o + j, acc[i][j]);
                    }
                }
            }
        }
    }
    if (db != nullptr) {
        for (int c = threadIdx.x; c < C; c += NT) {
            float s = 0.f;
            for (int r = 0; r < TR; ++r) s += dz[r * ldz + zcol + c];
            if (wacc != nullptr) wacc[(dW != nullptr ? C * K : 0) + c] += s; else atomicAdd(db + c, s);
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// interpreter
// ---------------------------------------------------------------------------------------------
template <int TR>
__global__ void __launch_bounds__(NT, 2) tile_program_kernel(const __grid_constant__ Program prog) {
    extern __shared__ __align__(16) float smem[];
    Tile t;
    t.smem = smem;
    t.pg = &prog;
    t.sigma = smem + prog.region_floats;
    t.wst = t.sigma + MAX_SIGMA * TR;
    t.ibuf = reinterpret_cast<int*>(t.wst + 2 * KC * CBMAX);
    t.red = reinterpret_cast<float*>(t.ibuf + 2 * TR);
    t.dacc = reinterpret_cast<double*>(t.red + 32);
    t.wacc = reinterpret_cast<float*>(t.dacc + 2 * MAX_STEPS);
    for (int i = threadIdx.x; i < 2 * MAX_STEPS; i += NT) t.dacc[i] = 0.;
    for (int i = threadIdx.x; i < prog.wacc_floats; i += NT) t.wacc[i] = 0.f;
    __syncthreads();

    const int n_tiles = (prog.n_rows + TR - 1) / TR;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        t.row0 = tile * TR;
        t.nvalid = min(TR, prog.n_rows - t.row0);
        for (int s = 0; s < prog.n_steps; ++s) {
            const Step& st = prog.steps[s];
            switch (st.op) {
                case OP_LOAD_ROWS: op_load_rows<TR>(t, st); break;
                case OP_LOAD_PAIRSUM: op_load_pairsum<TR>(t, st); break;
                case OP_LOAD_SEGMAX: op_load_segmax<TR>(t, st); break;
                case OP_LINEAR:
                    tile_gemm<TR>(t.reg(st.ra), t.ld(st.ra), st.i0, st.i3, static_cast<const float*>(st.p0), st.i4, st.i2,
                                  static_cast<const float*>(st.p1), st.i1, t.reg(st.rb), t.ld(st.rb), t.wst, st.i5 > 0 ? st.i5 : st.i2);
                    break;
                case OP_ADD_GATHER2: op_add_gather2<TR>(t, st); break;
                case OP_NORM_ACT:
                    tile_norm_act<TR>(t.reg(st.ra), t.ld(st.ra), st.i0, static_cast<const float*>(st.p0),
                                      static_cast<const float*>(st.p1), st.i1 != 0,
                                      st.i2 >= 0 ? t.sigma + st.i2 * TR : nullptr);
                    break;
                case OP_ADD_ROWS: op_add_rows<TR>(t, st); break;
                case OP_STORE_ROWS: op_store_rows<TR>(t, st); break;
                case OP_SEGSUM: op_segsum<TR>(t, st); break;
                case OP_ADD_REGION: op_add_region<TR>(t, st); break;
                case OP_ACTNORM_BWD: op_actnorm_bwd<TR>(t, st, s); break;
                case OP_WGRAD: op_wgrad<TR>(t, st); break;
                case OP_SCATTER_ADD: op_scatter_add<TR>(t, st); break;
                case OP_PAIR_SCATTER: op_pair_scatter<TR>(t, st); break;
                case OP_SEGMAX_BWD: op_segmax_bwd<TR>(t, st); break;
                default: break;
            }
        }
    }
    __syncthreads();
    for (int s = 0; s < prog.n_steps; ++s) {      // publish the shared-memory weight-gradient accumulators
        const Step& st = prog.steps[s];
        if (st.op != OP_WGRAD || st.i5 < 0) continue;
        const int C = st.i0, K = st.i1, ldW = st.i2, wcol = st.i3;
        float* dW = static_cast<float*>(const_cast<void*>(st.p0));
        float* db = static_cast<float*>(const_cast<void*>(st.p1));
        const float* wa = t.wacc + st.i5;
        if (dW != nullptr)
            for (int i = threadIdx.x; i < C * K; i += NT) atomicAdd(dW + (size_t)(i / K) * ldW + wcol + (i % K), wa[i]);
        if (db != nullptr)
            for (int c = threadIdx.x; c < C; c += NT) atomicAdd(db + c, wa[(dW != nullptr ? C * K : 0) + c]);
    }
    for (int s = threadIdx.x; s < prog.n_steps; s += NT) {
        const Step& st = prog.steps[s];
        if (st.op == OP_ACTNORM_BWD && st.p0 != nullptr && st.p2 != nullptr) {
            atomicAdd(static_cast<float*>(const_cast<void*>(st.p2)), (float)t.dacc[2 * s]);
            atomicAdd(static_cast<float*>(const_cast<void*>(st.p3)), (float)t.dacc[2 * s + 1]);
        }
    }
}

static size_t program_smem_bytes(const Program& p) {
    return ((size_t)p.region_floats + MAX_SIGMA * p.tr + 2 * KC * CBMAX + 32 + p.wacc_floats) * sizeof(float) + 2 * p.tr * sizeof(int) +
           2 * MAX_STEPS * sizeof(double);
}

constexpr size_t SMEM_LIMIT = 227 * 1024;

int launch_program(const Program& p_in, cudaStream_t stream) {
    if (p_in.n_rows <= 0) return RGNN_OK;
    Program p = p_in;
    // Two CTAs share an SM when their shared memory allows it (one CTA's global-memory phases overlap the other's
    // math); that budget takes precedence over the weight-gradient accumulators, which get what is left.
    p.wacc_floats = 0;
    const size_t half = (SMEM_LIMIT - 2048) / 2;
    const size_t budget = program_smem_bytes(p) <= half ? half : SMEM_LIMIT;
    for (int s = 0; s < p.n_steps; ++s) {
        Step& st = p.steps[s];
        if (st.op != OP_WGRAD) continue;
        st.i5 = -1;
        const int need = (st.p0 != nullptr ? st.i0 * st.i1 : 0) + (st.p1 != nullptr ? st.i0 : 0);
        Program q = p;
        q.wacc_floats = p.wacc_floats + need;
        if (need > 0 && program_smem_bytes(q) <= budget) {
            st.i5 = p.wacc_floats;
            p.wacc_floats += need;
        }
    }
    const size_t smem = program_smem_bytes(p);
    RGNN_REQUIRE(smem <= SMEM_LIMIT, "tile program needs %zu bytes of shared memory (> %zu)", smem, SMEM_LIMIT);
    RGNN_REQUIRE(p.tr == 64 || p.tr == 32, "tile program with tr=%d", p.tr);
    static PerDeviceOnce once;
    if (once.needed()) {
        RGNN_CHECK_CUDA(cudaFuncSetAttribute(tile_program_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
        RGNN_CHECK_CUDA(cudaFuncSetAttribute(tile_program_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
        once.mark();
    }
    const int n_tiles = (p.n_rows + p.tr - 1) / p.tr;
    // two CTAs share an SM when their shared memory allows it: one CTA's global-memory phases overlap the other's math
    const int per_sm = smem <= half ? 2 : 1;
    const int grid = min(n_tiles, per_sm * sm_count());
    if (p.tr == 64)
        tile_program_kernel<64><<<grid, NT, smem, stream>>>(p);
    else
        tile_program_kernel<32><<<grid, NT, smem, stream>>>(p);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

// ---------------------------------------------------------------------------------------------
// program builder
// ---------------------------------------------------------------------------------------------
ProgBuilder::ProgBuilder(int n_rows, int tr) {
    memset(&p, 0, sizeof(p));
    p.n_rows = n_rows;
    p.tr = tr;
}

int ProgBuilder::region(int width) {
    if (n_regions >= MAX_REGIONS) {
        set_error("tile program uses too many shared-memory regions");
        ok = false;
        return 0;
    }
    const int ld = round_up(width, 4) + 4;      // +4 floats: rows start 16 B apart modulo the 128 B bank window
    p.reg_off[n_regions] = p.region_floats;
    p.reg_ld[n_regions] = ld;
    p.region_floats += p.tr * ld;
    return n_regions++;
}

int ProgBuilder::sigma_slot() {
    if (n_sigma >= MAX_SIGMA) {
        set_error("tile program uses too many normalisation layers");
        ok = false;
        return 0;
    }
    return n_sigma++;
}

Step* ProgBuilder::add(int op, int ra, int rb) {
    static Step dummy;
    if (p.n_steps >= MAX_STEPS) {
        set_error("tile program too long (%d steps)", p.n_steps);
        ok = false;
        return &dummy;
    }
    Step& s = p.steps[p.n_steps++];
    memset(&s, 0, sizeof(s));
    s.op = op;
    s.ra = (short)ra;
    s.rb = (short)rb;
    return &s;
}

bool check_linear(const rgnn_linear& L) {
    if (L.in_features <= 0 || L.in_features > 512 || L.out_features <= 0 || L.out_features > 512) {
        set_error("linear %dx%d outside the supported widths (<=512)", L.out_features, L.in_features);
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

void ProgBuilder::load_rows(int ra, const float* src, int ld, int w, int dcol, int padto, const int* ridx, int scol) {
    Step* s = add(OP_LOAD_ROWS, ra);
    s->p0 = src; s->p1 = ridx;
    s->i0 = ld; s->i1 = w; s->i2 = dcol; s->i3 = padto < w ? w : padto; s->i4 = scol;
}

void ProgBuilder::gemm(int ra, int rb, const float* Wt, int ldw, int K, int k_valid, int C, int Cpad, const float* bias, int c_exist) {
    Step* s = add(OP_LINEAR, ra, rb);
    s->p0 = Wt; s->p1 = bias;
    s->i0 = K; s->i1 = C; s->i2 = Cpad; s->i3 = k_valid; s->i4 = ldw; s->i5 = c_exist;
}

void ProgBuilder::norm_act(int ra, const rgnn_linear& L, int slot) {
    if (L.norm_scale == nullptr && !L.activation) return;
    Step* s = add(OP_NORM_ACT, ra);
    s->p0 = L.norm_scale; s->p1 = L.norm_shift;
    s->i0 = L.out_features; s->i1 = L.activation; s->i2 = slot;
}

void ProgBuilder::linear(int ra, int rb, const rgnn_linear& L, int slot) {
    if (!check_linear(L)) { ok = false; return; }
    const int Kp = round_up(L.in_features, 8), Cp = round_up(L.out_features, 64);
    gemm(ra, rb, L.weight_t, Cp, Kp, Kp, L.out_features, Cp, L.bias);
    norm_act(rb, L, slot);
}

void ProgBuilder::store_rows(int ra, float* dst, int ld, int w, int dcol, bool accumulate, int scol) {
    Step* s = add(OP_STORE_ROWS, ra);
    s->p0 = dst;
    s->i0 = ld; s->i1 = w; s->i2 = dcol; s->i3 = accumulate ? 1 : 0; s->i4 = scol;
}

}  // namespace rgnn
