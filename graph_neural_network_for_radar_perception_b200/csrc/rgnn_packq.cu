// Batched weight-image packing: see rgnn_pack.cuh.
#include <vector>

#include "rgnn_common.cuh"
#include "rgnn_pack.cuh"

namespace rgnn {

int g_pack_batch = 1;

constexpr int PACKQ_BATCH = 48;       // entries per launch: 48 x 64 B of kernel parameters
template <typename A>
struct PackBatch { A e[PACKQ_BATCH]; };

__global__ void pack_tc_batch_kernel(const __grid_constant__ PackBatch<PackTcArgs> t) {
    pack_tc_body(t.e[blockIdx.y], blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}
__global__ void pack_f16_batch_kernel(const __grid_constant__ PackBatch<PackF16Args> t) {
    pack_f16_body(t.e[blockIdx.y], blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}
__global__ void pack_split_batch_kernel(const __grid_constant__ PackBatch<PackSplitArgs> t) {
    pack_split_body(t.e[blockIdx.y], blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

namespace {
struct PackQueue {
    bool open = false;
    cudaStream_t stream = nullptr;
    std::vector<PackTcArgs> tc;
    std::vector<PackF16Args> f16;
    std::vector<PackSplitArgs> split;
};
thread_local PackQueue g_q;

template <typename A, typename K>
int flush_kind(std::vector<A>& v, K kernel, cudaStream_t stream) {
    for (size_t i = 0; i < v.size(); i += PACKQ_BATCH) {
        PackBatch<A> b;
        const int n = (int)(v.size() - i < (size_t)PACKQ_BATCH ? v.size() - i : (size_t)PACKQ_BATCH);
        for (int j = 0; j < n; ++j) b.e[j] = v[i + j];
        kernel<<<dim3(16, n), 256, 0, stream>>>(b);
        RGNN_CHECK_CUDA(cudaGetLastError());
    }
    v.clear();
    return RGNN_OK;
}
}  // namespace

void packq_begin(cudaStream_t stream) {
    g_q.open = true;
    g_q.stream = stream;
    g_q.tc.clear(); g_q.f16.clear(); g_q.split.clear();
}

int packq_flush() {
    if (!g_q.open) return RGNN_OK;
    int rc = flush_kind(g_q.tc, pack_tc_batch_kernel, g_q.stream);
    if (rc == RGNN_OK) rc = flush_kind(g_q.f16, pack_f16_batch_kernel, g_q.stream);
    if (rc == RGNN_OK) rc = flush_kind(g_q.split, pack_split_batch_kernel, g_q.stream);
    return rc;
}

int packq_end() {
    const int rc = packq_flush();
    g_q.open = false;
    return rc;
}

bool packq_push(const PackTcArgs& a) { if (!g_q.open) return false; g_q.tc.push_back(a); return true; }
bool packq_push(const PackF16Args& a) { if (!g_q.open) return false; g_q.f16.push_back(a); return true; }
bool packq_push(const PackSplitArgs& a) { if (!g_q.open) return false; g_q.split.push_back(a); return true; }

}  // namespace rgnn
