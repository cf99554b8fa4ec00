// Micro test (B200): the "chunk-major" fp16 shared-memory images of rgnn_f16.cuh read as MN-major operands of
// tcgen05.mma.kind::f16 (instruction descriptor bits 15 / 16), which is what lets the fused message backward use ONE weight
// image for W and W^T and take the weight gradients (reduction over EDGES) straight from the per-tile activation images.
//   test 1 (dgrad, TS):  D[128 x 128] = A[128 x 64] (TMEM)  x  W2 (64 x 128), W2 stored as the K-major image of the FORWARD
//                        GEMM (N = 64 rows, K = 128) and read MN-major here (N = 128, K = 64)
//   test 2 (wgrad, SS):  D[128 x 80]  = Y^T Z,  Y image [16 chunks][128 edges][8], Z image [10 chunks][128 edges][8], both MN-major
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o test_mnmajor test_mnmajor.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../../graph_neural_network_for_radar_perception_b200/csrc/rgnn_f16.cuh"
using namespace rgnn;

__host__ __device__ constexpr uint32_t idesc_mn(int M, int N, int a_mn, int b_mn) {
    return f16::idesc(M, N) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}

__global__ void __launch_bounds__(128, 1) k_test(const __half* w2img, const __half* a1, const __half* yimg, const __half* zimg,
                                                 float* d1, float* d2) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __half* sW = reinterpret_cast<__half*>(sm);                  // 16 x 64 x 8 halves = 16 KB
    __half* sY = sW + 16 * 64 * 8;                               // 16 x 128 x 8     = 32 KB
    __half* sZ = sY + 16 * 128 * 8;                              // 10 x 128 x 8     = 20 KB
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 16 * 64 * 8; i += 128) sW[i] = w2img[i];
    for (int i = tid; i < 16 * 128 * 8; i += 128) sY[i] = yimg[i];
    for (int i = tid; i < 10 * 128 * 8; i += 128) sZ[i] = zimg[i];
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_init_fence(); }
    if (warp == 0) tc::tmem_alloc(&slot, 512);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = slot;
    const uint32_t t_row = tmem + ((uint32_t)warp << 21);
    // A operand of test 1: row = tid, K = 64 halves = 32 packed columns at [256, 288)
    {
        uint32_t v[32];
        const uint32_t* src = reinterpret_cast<const uint32_t*>(a1 + (size_t)tid * 64);
        for (int i = 0; i < 32; ++i) v[i] = src[i];
        f16::tmem_st16u(t_row + 256, v);
        f16::tmem_st16u(t_row + 272, v + 16);
        tc::tmem_wait_st();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    if (tid == 0) {
        // test 1: N = 128 (h), K = 64 (cn); image element (h, cn) at ((h/8)*64 + cn)*8 + h%8 halves -> LBO (next 8 k) = 128 B, SBO (next 8 n) = 1024 B
        const uint32_t id1 = idesc_mn(128, 128, 0, 1);
        const uint64_t b0 = tc::smem_desc(tc::smem_u32(sW), 128, 1024);
        for (int ks = 0; ks < 4; ++ks) f16::mma_ts(tmem + 0, tmem + 256 + ks * 8, b0 + (uint64_t)((ks * 256) >> 4), id1, ks > 0);
        // test 2: A = Y image MN-major (M = 128 channels, K = 128 edges): LBO (next 8 edges) = 128 B, SBO (next 8 channels) = 2048 B
        const uint32_t id2 = idesc_mn(128, 80, 1, 1);
        const uint64_t ya = tc::smem_desc(tc::smem_u32(sY), 128, 2048);
        const uint64_t zb = tc::smem_desc(tc::smem_u32(sZ), 128, 2048);
        for (int ks = 0; ks < 8; ++ks)
            f16::mma_ss(tmem + 128, ya + (uint64_t)((ks * 256) >> 4), zb + (uint64_t)((ks * 256) >> 4), id2, ks > 0);
        tc::mma_commit(&bar);
    }
    __syncwarp();
    tc::mbar_wait(&bar, 0);
    tc::tc_fence_after();
    for (int c = 0; c < 128; c += 16) {
        float v[16];
        tc::tmem_ld16(t_row + c, v);
        tc::tmem_wait_ld();
        for (int i = 0; i < 16; ++i) d1[(size_t)tid * 128 + c + i] = v[i];
    }
    for (int c = 0; c < 80; c += 16) {
        float v[16];
        tc::tmem_ld16(t_row + 128 + c, v);
        tc::tmem_wait_ld();
        for (int i = 0; i < 16; ++i) d2[(size_t)tid * 80 + c + i] = v[i];
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

int main() {
    srand(1);
    auto rnd = []() { return (float)((rand() % 17) - 8) * 0.25f; };
    // W2 (cn = 64 x h = 128), forward image [h/8][cn][8]
    std::vector<float> W2(64 * 128), A1(128 * 64), Y(128 * 128), Z(128 * 80);
    for (auto& v : W2) v = rnd();
    for (auto& v : A1) v = rnd();
    for (auto& v : Y) v = rnd();
    for (auto& v : Z) v = rnd();
    std::vector<__half> w2img(16 * 64 * 8), a1h(128 * 64), yimg(16 * 128 * 8), zimg(10 * 128 * 8);
    for (int cn = 0; cn < 64; ++cn) for (int h = 0; h < 128; ++h) w2img[((h / 8) * 64 + cn) * 8 + h % 8] = __float2half(W2[cn * 128 + h]);
    for (int i = 0; i < 128 * 64; ++i) a1h[i] = __float2half(A1[i]);
    for (int e = 0; e < 128; ++e) for (int c = 0; c < 128; ++c) yimg[((c / 8) * 128 + e) * 8 + c % 8] = __float2half(Y[e * 128 + c]);
    for (int e = 0; e < 128; ++e) for (int c = 0; c < 80; ++c) zimg[((c / 8) * 128 + e) * 8 + c % 8] = __float2half(Z[e * 80 + c]);
    __half *dw, *da, *dy, *dz;
    float *d1, *d2;
    cudaMalloc(&dw, w2img.size() * 2); cudaMalloc(&da, a1h.size() * 2); cudaMalloc(&dy, yimg.size() * 2); cudaMalloc(&dz, zimg.size() * 2);
    cudaMalloc(&d1, 128 * 128 * 4); cudaMalloc(&d2, 128 * 80 * 4);
    cudaMemcpy(dw, w2img.data(), w2img.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(da, a1h.data(), a1h.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dy, yimg.data(), yimg.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dz, zimg.data(), zimg.size() * 2, cudaMemcpyHostToDevice);
    const int smem = (16 * 64 * 8 + 16 * 128 * 8 + 10 * 128 * 8) * 2;
    cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_test<<<1, 128, smem>>>(dw, da, dy, dz, d1, d2);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<float> h1(128 * 128), h2(128 * 80);
    cudaMemcpy(h1.data(), d1, h1.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(h2.data(), d2, h2.size() * 4, cudaMemcpyDeviceToHost);
    double e1 = 0, e2 = 0;
    for (int r = 0; r < 128; ++r) for (int h = 0; h < 128; ++h) {
        double s = 0;
        for (int cn = 0; cn < 64; ++cn) s += (double)A1[r * 64 + cn] * W2[cn * 128 + h];
        e1 = fmax(e1, fabs(s - h1[r * 128 + h]));
    }
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 80; ++n) {
        double s = 0;
        for (int ed = 0; ed < 128; ++ed) s += (double)Y[ed * 128 + m] * Z[ed * 80 + n];
        e2 = fmax(e2, fabs(s - h2[m * 80 + n]));
    }
    printf("test1 (TS, B MN-major) max abs err %.3e   sample got %.3f\n", e1, h1[5 * 128 + 7]);
    printf("test2 (SS, A and B MN-major, N = 80) max abs err %.3e   sample got %.3f\n", e2, h2[5 * 80 + 7]);
    return 0;
}
