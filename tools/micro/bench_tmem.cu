// Microbenchmark (B200): throughput of tcgen05.ld / tcgen05.st (32x32b.x16) as a function of the number of warps, and of
// LDGSTS (cp.async 16 B) row gathers issued by 1 / 2 / 4 warps.   nvcc -arch=sm_100a -O3 -o bench_tmem bench_tmem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../graph_neural_network_for_radar_perception_b200/csrc/rgnn_tc.cuh"
using namespace rgnn;

__global__ void __launch_bounds__(512, 1) tmem_kernel(int nwarps, int reps, int mode, long long* out, float* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tc::tmem_alloc(&slot, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = slot;
    const uint32_t t_row = tmem + ((uint32_t)(warp & 3) << 21);
    float v[16];
    for (int i = 0; i < 16; ++i) v[i] = (float)(threadIdx.x + i);
    // initialise all columns
    if (warp < 4) for (int c = 0; c < 512; c += 16) tc::tmem_st16(t_row + c, v);
    tc::tmem_wait_st();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    long long t0 = clock64();
    float acc = 0.f;
    if (warp < nwarps) {
        for (int r = 0; r < reps; ++r) {
            if (mode == 0) {            // 128 columns read, then wait (like an epilogue)
#pragma unroll
                for (int c = 0; c < 128; c += 16) { tc::tmem_ld16(t_row + ((warp >> 2) * 128) % 512 + c, v); }
                tc::tmem_wait_ld();
                acc += v[0] + v[15];
            } else if (mode == 1) {     // 128 columns written, then wait
#pragma unroll
                for (int c = 0; c < 128; c += 16) tc::tmem_st16(t_row + ((warp >> 2) * 128) % 512 + c, v);
                tc::tmem_wait_st();
            } else {                    // one 16-column read + wait (latency)
                tc::tmem_ld16(t_row, v);
                tc::tmem_wait_ld();
                acc += v[3];
            }
        }
    }
    long long t1 = clock64();
    if (lane == 0 && warp < nwarps) out[blockIdx.x * 16 + warp] = t1 - t0;
    if (acc == 12345.f) sink[0] = acc;
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

__device__ __forceinline__ void cp16(void* s, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(s)), "l"(g));
}
// gather of 128 rows x 512 B per "tile" from a big table with random row ids, by nw warps (whole row per warp instruction)
__global__ void __launch_bounds__(512, 1) gather_kernel(const float* P, const int* ids, int n_ids, int nw, int tiles, int mode, long long* out) {
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long t0 = clock64();
    if (warp < nw) {
        for (int t = 0; t < tiles; ++t) {
            float* dst = sm + (t & 1) * 128 * 128;
            const int base = ((blockIdx.x * tiles + t) * 128) % (n_ids - 128);
            const int rows = 128 / nw;
            if (mode == 0) {
                for (int k = 0; k < rows / 32 + (rows < 32); ++k) {
                    const int myid = ids[base + warp * rows + k * 32 + (lane % (rows < 32 ? rows : 32))];
                    const int n = rows < 32 ? rows : 32;
#pragma unroll 8
                    for (int i = 0; i < n; ++i) {
                        const int r = warp * rows + k * 32 + i;
                        const int sn = __shfl_sync(0xffffffffu, myid, i);
                        cp16(dst + r * 128 + ((lane ^ (r & 7)) << 2), P + (size_t)sn * 256 + 128 + 4 * lane);
                    }
                }
            } else {                    // one lane = one row: 32 x 16-byte copies of its own row (no shuffles)
                for (int k = 0; k < rows / 32 + (rows < 32); ++k) {
                    if (rows >= 32 || lane < rows) {
                        const int r = warp * rows + k * 32 + lane;
                        const int sn = ids[base + r];
#pragma unroll 8
                        for (int c = 0; c < 32; ++c) cp16(dst + r * 128 + ((c ^ (r & 7)) << 2), P + (size_t)sn * 256 + 128 + 4 * c);
                    }
                }
            }
            asm volatile("cp.async.commit_group;\n" ::);
            asm volatile("cp.async.wait_group 1;\n" ::);
        }
        asm volatile("cp.async.wait_group 0;\n" ::);
    }
    long long t1 = clock64();
    if (lane == 0 && warp < nw) out[blockIdx.x * 16 + warp] = t1 - t0;
}

int main() {
    long long* out; float* sink;
    cudaMalloc(&out, 148 * 16 * 8); cudaMalloc(&sink, 4);
    long long h[148 * 16];
    const int reps = 2000;
    for (int mode = 0; mode < 3; ++mode)
        for (int nw : {4, 8, 16}) {
            tmem_kernel<<<148, 512>>>(nw, reps, mode, out, sink);
            cudaDeviceSynchronize();
            cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            double mx = 0; for (int b = 0; b < 148; ++b) for (int w = 0; w < nw; ++w) mx = h[b * 16 + w] > mx ? h[b * 16 + w] : mx;
            const double bytes = mode == 2 ? 32.0 * 16 * 4 * nw : 32.0 * 128 * 4 * nw;
            printf("tmem mode %d (%s) warps %2d: %.1f cycles per rep per warp, %.1f B/clk/SM  [%s]\n", mode, mode == 0 ? "ld 128 cols + wait" : mode == 1 ? "st 128 cols + wait" : "ld 16 cols + wait",
                   nw, mx / reps, bytes * reps / mx, cudaGetErrorString(cudaGetLastError()));
        }
    // gather
    const int N = 768000;
    float* P; int* ids;
    cudaMalloc(&P, (size_t)N * 256 * 4); cudaMemset(P, 0, (size_t)N * 256 * 4);
    const int n_ids = 1 << 22;
    int* hid = new int[n_ids];
    unsigned s = 12345;
    for (int i = 0; i < n_ids; ++i) { s = s * 1664525u + 1013904223u; hid[i] = (int)((s >> 8) % 6000) + (i / 35000) * 3000 % (N - 6000); }   // frame-local neighbourhoods
    cudaMalloc(&ids, n_ids * 4); cudaMemcpy(ids, hid, n_ids * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 128 * 128 * 4);
    const int tiles = 200;
    for (int mode = 0; mode < 2; ++mode)
        for (int nw : {1, 2, 4, 8}) {
            gather_kernel<<<148, 512, 2 * 128 * 128 * 4>>>(P, ids, n_ids, nw, tiles, mode, out);
            cudaDeviceSynchronize();
            cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            double mx = 0; for (int b = 0; b < 148; ++b) for (int w = 0; w < nw; ++w) mx = h[b * 16 + w] > mx ? h[b * 16 + w] : mx;
            printf("gather mode %d (%s) warps %d: %.0f cycles per tile (64 KB), %.1f B/clk/SM [%s]\n", mode, mode == 0 ? "row per warp instr" : "row per lane", nw, mx / tiles, 65536.0 * tiles / mx,
                   cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
