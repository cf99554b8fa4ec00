// Microbenchmark: how fast can one SM gather 512-byte rows (P_s[src]) from L2 with different access shapes?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench_gather bench_gather.cu && ./bench_gather
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int ROWF = 256;   // floats per P row (2H)
constexpr int H = 128;

__device__ __forceinline__ void ldg256(const float* p, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}

// A: thread = (row, quarter): 4 x 32-byte loads of its 128-byte piece
__global__ void __launch_bounds__(512) gatherA(const float* __restrict__ P, const int* __restrict__ src, int n_edges, float* out) {
    const int row = threadIdx.x & 127, q = threadIdx.x >> 7;
    float acc = 0.f;
    const int n_tiles = n_edges / 128;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int s = src[t * 128 + row];
        const float* p = P + (size_t)s * ROWF + H + q * 32;
        float4 a[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) ldg256(p + 8 * i, a[2 * i], a[2 * i + 1]);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += a[i].x + a[i].y + a[i].z + a[i].w;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
// C: same mapping, 8 x 16-byte loads
__global__ void __launch_bounds__(512) gatherC(const float* __restrict__ P, const int* __restrict__ src, int n_edges, float* out) {
    const int row = threadIdx.x & 127, q = threadIdx.x >> 7;
    float acc = 0.f;
    const int n_tiles = n_edges / 128;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int s = src[t * 128 + row];
        const float4* p = reinterpret_cast<const float4*>(P + (size_t)s * ROWF + H + q * 32);
        float4 a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = __ldg(p + i);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += a[i].x + a[i].y + a[i].z + a[i].w;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
// B: a warp reads whole rows: lane = 16-byte chunk of the 512-byte half row; 16 warps x 8 rows per tile
__global__ void __launch_bounds__(512) gatherB(const float* __restrict__ P, const int* __restrict__ src, int n_edges, float* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0.f;
    const int n_tiles = n_edges / 128;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int s_l = src[t * 128 + warp * 8 + (lane & 7)];
        float4 a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int s = __shfl_sync(0xffffffffu, s_l, i);
            a[i] = __ldg(reinterpret_cast<const float4*>(P + (size_t)s * ROWF + H) + lane);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += a[i].x + a[i].y + a[i].z + a[i].w;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
// S: streaming read of the same number of bytes (coalesced), for reference
__global__ void __launch_bounds__(512) streamS(const float* __restrict__ P, int n_edges, float* out) {
    float acc = 0.f;
    const int n_tiles = n_edges / 128;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const float4* p = reinterpret_cast<const float4*>(P + ((size_t)t * 128 * H) % ((size_t)90000 * ROWF));
        float4 a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = __ldg(p + i * 512 + threadIdx.x);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += a[i].x + a[i].y + a[i].z + a[i].w;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
    const int n_nodes = 96000, n_edges = 1174528 / 128 * 128, per_frame = 3000;
    std::vector<int> src(n_edges);
    srand(1);
    for (int e = 0; e < n_edges; ++e) {
        const int frame = (int)((long long)e * 32 / n_edges);
        src[e] = frame * per_frame + rand() % per_frame;
    }
    float *P, *out; int* dsrc;
    cudaMalloc(&P, (size_t)n_nodes * ROWF * 4); cudaMemset(P, 0, (size_t)n_nodes * ROWF * 4);
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&dsrc, n_edges * 4);
    cudaMemcpy(dsrc, src.data(), n_edges * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto launch) {
        for (int i = 0; i < 3; ++i) launch();
        cudaEventRecord(e0);
        for (int i = 0; i < 10; ++i) launch();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
        const double tiles_per_sm = n_edges / 128.0 / 148.0;
        printf("%-8s %.3f ms  -> %.0f cycles/tile/SM @1.9GHz, %.1f GB/s useful\n", name, ms, ms * 1e-3 * 1.9e9 / tiles_per_sm,
               (double)n_edges * 512 / (ms * 1e-3) / 1e9);
    };
    run("A 32B", [&] { gatherA<<<148, 512>>>(P, dsrc, n_edges, out); });
    run("C 16B", [&] { gatherC<<<148, 512>>>(P, dsrc, n_edges, out); });
    run("B row", [&] { gatherB<<<148, 512>>>(P, dsrc, n_edges, out); });
    run("stream", [&] { streamS<<<148, 512>>>(P, n_edges, out); });
    run("A x2occ", [&] { gatherA<<<296, 512>>>(P, dsrc, n_edges, out); });
    run("B x2occ", [&] { gatherB<<<296, 512>>>(P, dsrc, n_edges, out); });
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
