// Microbenchmark (B200): issue cost of the instructions of the fp16-split epilogues, one warp per SM sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t pack_sat(float x0, float x1) {
    uint32_t r;
    asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x1), "f"(x0));
    return r;
}
__device__ __forceinline__ uint32_t pack_rn(float x0, float x1) {
    uint32_t r;
    asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x1), "f"(x0));
    return r;
}
template <int MODE>
__global__ void ops_kernel(int reps, long long* out, float* sink, float seed) {
    float2 v[8];
    uint32_t u[8];
    for (int i = 0; i < 8; ++i) { v[i] = make_float2(seed + i + threadIdx.x, seed * 0.5f + i); u[i] = i * 977u + threadIdx.x; }
    const float2 k = make_float2(1.0001f, 0.9999f), c = make_float2(0.01f, 0.01f);
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) v[i] = __ffma2_rn(v[i], k, c);                                   // FFMA2
            else if (MODE == 1) { v[i].x = fmaf(v[i].x, k.x, c.x); v[i].y = fmaf(v[i].y, k.y, c.y); }   // 2 x FFMA
            else if (MODE == 2) { v[i].x = fmaxf(v[i].x, c.x); v[i].y = fmaxf(v[i].y, c.y); }         // 2 x FMNMX
            else if (MODE == 3) { u[i] = pack_sat(v[i].x, v[i].y); v[i].x = __uint_as_float(u[i]); }   // F2FP satfinite (+ dependency)
            else if (MODE == 4) { u[i] = pack_rn(v[i].x, v[i].y); v[i].x = __uint_as_float(u[i]); }    // F2FP
            else if (MODE == 5) { const float2 h = __half22float2(*reinterpret_cast<__half2*>(&u[i])); v[i] = __fadd2_rn(v[i], h); }  // unpack x2 + FADD2
            else if (MODE == 6) {       // the whole split
                const uint32_t hi = pack_sat(v[i].x, v[i].y);
                const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi));
                const uint32_t lo = pack_sat(v[i].x - h.x, v[i].y - h.y);
                u[i] ^= hi + lo;
                v[i].x += 1.f;
            } else if (MODE == 7) {     // LeakyReLU on a pair: FMUL2 + 2 FMNMX
                const float2 t = __fmul2_rn(v[i], c);
                v[i].x = fmaxf(v[i].x, t.x); v[i].y = fmaxf(v[i].y, t.y);
            } else if (MODE == 8) {     // split with the integer rounding trick for hi (no F2FP for the fp32 image of hi)
                const float hx = __uint_as_float((__float_as_uint(v[i].x) + 0x1000u) & 0xFFFFE000u);
                const float hy = __uint_as_float((__float_as_uint(v[i].y) + 0x1000u) & 0xFFFFE000u);
                const uint32_t hi = pack_rn(hx, hy);
                const uint32_t lo = pack_rn(v[i].x - hx, v[i].y - hy);
                u[i] ^= hi + lo;
                v[i].x += 1.f;
            }
        }
    }
    const long long t1 = clock64();
    float acc = 0.f;
    for (int i = 0; i < 8; ++i) acc += v[i].x + v[i].y + (float)u[i];
    if (acc == 12345.678f) sink[0] = acc;
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * 32 + (threadIdx.x >> 5)] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps, long long* out, float* sink) {
    const int reps = 4000;
    ops_kernel<MODE><<<148, 32 * warps>>>(reps, out, sink, 1.5f);
    cudaDeviceSynchronize();
    long long h[148 * 32];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int b = 0; b < 148; ++b) for (int w = 0; w < warps; ++w) mx = h[b * 32 + w] > mx ? h[b * 32 + w] : mx;
    printf("%-44s warps/SM %2d: %.2f cycles per loop body element (8 independent per rep) [%s]\n", name, warps, mx / reps / 8, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    long long* out; float* sink;
    cudaMalloc(&out, 148 * 32 * 8); cudaMalloc(&sink, 4);
    for (int warps : {4, 8, 16}) {
        run<0>("FFMA2", warps, out, sink);
        run<1>("2 x FFMA", warps, out, sink);
        run<2>("2 x FMNMX", warps, out, sink);
        run<3>("F2FP satfinite", warps, out, sink);
        run<4>("F2FP rn", warps, out, sink);
        run<5>("unpack f16x2 -> 2 f32 + FADD2", warps, out, sink);
        run<6>("whole split (pack, unpack, 2 sub, pack)", warps, out, sink);
        run<7>("LeakyReLU pair (FMUL2 + 2 FMNMX)", warps, out, sink);
        run<8>("split with integer hi rounding", warps, out, sink);
    }
    return 0;
}
