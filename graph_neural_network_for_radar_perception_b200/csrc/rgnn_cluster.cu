// Proposal clustering on the device: the step between the link / offset heads and the object-class head of
// Model_Inference.forward when no cluster list is given (reference gnn_detector.py:164-187, Simple_DBSCAN in
// modules/inference/clustering.py:8-92).
//
// The reference builds a dense N x N adjacency (predicted links that survive a distance gate, or all pairs of predicted
// centres within eps) and grows clusters by a Python breadth-first search.  A BFS over every unvisited node in index order
// labels the CONNECTED COMPONENTS of that graph, numbered by their smallest member; members are listed in ascending order
// (np.nonzero).  Here: lock-free union-find with "larger root hooks under smaller root" (so every root is its
// component's smallest node), rank of the roots = cluster id, stable radix sort of the nodes by cluster id = member lists.
#include <cub/device/device_radix_sort.cuh>

#include "rgnn_model.h"

namespace rgnn {

__device__ __forceinline__ int uf_find(int* parent, int i) {
    int p = __ldcg(parent + i);  // L2 reads: other CTAs hook roots concurrently, an L1 line may be stale
    while (p != i) {            // path halving; concurrent hooks only ever lower parent[] towards smaller ids
        const int gp = __ldcg(parent + p);
        if (gp != p) parent[i] = gp;
        i = p;
        p = __ldcg(parent + i);
    }
    return i;
}

__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    while (a != b) {
        if (a < b) { const int t = a; a = b; b = t; }          // a > b: hook a under b
        const int seen = atomicCAS(parent + a, a, b);
        if (seen == a) return;
        // a is no longer a root: `seen` is the parent another thread gave it.  Continue from the value the atomic itself
        // returned (an L2-coherent read) instead of re-reading parent[] through L1, where a stale line could keep
        // returning the old root and the loop would only end when that line happens to be evicted.
        a = uf_find(parent, seen);
        b = uf_find(parent, b);
    }
}

__global__ void uf_init_kernel(int* parent, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) parent[i] = i;
}

// links mode (clustering.py:9-24): the undirected pair (r, c) is an edge iff its predicted class is 1 and NOT
// (sqrt(dx^2 + dy^2) >= eps); float32 like the reference's arrays, products and sum rounded separately
__global__ void uf_links_kernel(const float* __restrict__ xy, const int* __restrict__ und_a, const int* __restrict__ und_b,
                                const float* __restrict__ logits, int n_und, float eps, int* parent) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_und; k += gridDim.x * blockDim.x) {
        const float2 l = __ldg(reinterpret_cast<const float2*>(logits) + k);
        if (!(l.y > l.x)) continue;                           // torch.max(softmax, -1): class 1 only if strictly larger
        const int a = __ldg(und_a + k), b = __ldg(und_b + k);
        const float dx = xy[2 * a] - xy[2 * b], dy = xy[2 * a + 1] - xy[2 * b + 1];
        const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
        if (d >= eps) continue;
        uf_union(parent, a, b);
    }
}

// offsets mode (clustering.py:27-41): all pairs of one frame with (a - b)^T (a - b) <= eps
__global__ void uf_radius_kernel(const float* __restrict__ xy, const int* __restrict__ frame_ptr, int n_frames, float eps,
                                 int* parent) {
    __shared__ float2 tile[256];
    const int f = blockIdx.y;
    const int p0 = frame_ptr[f], p1 = frame_ptr[f + 1];
    for (int i0 = p0 + blockIdx.x * blockDim.x; i0 < p1; i0 += gridDim.x * blockDim.x) {
        const int i = i0 + threadIdx.x;
        const float2 me = i < p1 ? reinterpret_cast<const float2*>(xy)[i] : make_float2(0.f, 0.f);
        for (int j0 = i0; j0 < p1; j0 += 256) {               // only j > i: the relation is symmetric
            __syncthreads();
            if (j0 + (int)threadIdx.x < p1) tile[threadIdx.x] = reinterpret_cast<const float2*>(xy)[j0 + threadIdx.x];
            __syncthreads();
            const int nj = min(256, p1 - j0);
            if (i < p1) {
                for (int jj = 0; jj < nj; ++jj) {
                    const int j = j0 + jj;
                    if (j <= i) continue;
                    const float dx = me.x - tile[jj].x, dy = me.y - tile[jj].y;
                    if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) <= eps) uf_union(parent, i, j);
                }
            }
        }
    }
}

__global__ void uf_flatten_kernel(int* parent, int n, int* is_root) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) is_root[i] = parent[i] == i ? 1 : 0;
}

__global__ void uf_label_kernel(const int* __restrict__ parent, const int* __restrict__ root_rank, int n, int* cluster_id,
                                int* node_idx, int* count) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int r = i;
        while (parent[r] != r) r = parent[r];
        const int c = root_rank[r];
        cluster_id[i] = c;
        node_idx[i] = i;
        atomicAdd(count + c, 1);
    }
}

// workspace (ints): parent (n) | is_root (n+1) | root_rank (n+1) | count (n+1) | node_idx (n) | keys_out (n) | scan ws | cub temp
static size_t cluster_ints(int n) { return 6 * (align256((size_t)(n + 2) * 4) / 4) + align256(scan_ws_ints(n + 1) * 4) / 4; }

static size_t cub_sort_bytes(int n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int*)nullptr, (int*)nullptr, (const int*)nullptr, (int*)nullptr, n);
    return align256(bytes);
}

static int finish_clusters(int* ws_i, int n, size_t ws_bytes, int* cluster_id, int* n_clusters_out, int* cl_ptr, int* cl_members,
                           cudaStream_t stream) {
    const size_t nn = align256((size_t)(n + 2) * 4) / 4;
    int* parent = ws_i;
    int* is_root = parent + nn;
    int* root_rank = is_root + nn;
    int* count = root_rank + nn;
    int* node_idx = count + nn;
    int* keys_out = node_idx + nn;
    int* scan_ws = keys_out + nn;
    void* cub_tmp = scan_ws + align256(scan_ws_ints(n + 1) * 4) / 4;
    const int blocks = (n + 255) / 256 > 1184 ? 1184 : (n + 255) / 256;
    uf_flatten_kernel<<<blocks, 256, 0, stream>>>(parent, n, is_root);
    int rc = exclusive_scan(is_root, n, root_rank, scan_ws, stream);      // root_rank[n] = number of clusters
    if (rc) return rc;
    RGNN_CHECK_CUDA(cudaMemcpyAsync(n_clusters_out, root_rank + n, sizeof(int), cudaMemcpyDeviceToDevice, stream));
    RGNN_CHECK_CUDA(cudaMemsetAsync(count, 0, (size_t)(n + 1) * sizeof(int), stream));
    uf_label_kernel<<<blocks, 256, 0, stream>>>(parent, root_rank, n, cluster_id, node_idx, count);
    rc = exclusive_scan(count, n, cl_ptr, scan_ws, stream);               // entries past the last cluster repeat n
    if (rc) return rc;
    size_t cub_bytes = cub_sort_bytes(n);
    RGNN_REQUIRE((char*)cub_tmp + cub_bytes <= (char*)ws_i + ws_bytes, "cluster workspace too small");
    // stable LSD radix sort by cluster id keeps the nodes of a cluster in ascending order (np.nonzero order)
    int bits = 1;
    while ((1 << bits) < n + 1 && bits < 31) ++bits;
    RGNN_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, (const int*)cluster_id, keys_out, (const int*)node_idx, cl_members, n,
                                                    0, bits, stream));
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

}  // namespace rgnn

using namespace rgnn;

extern "C" size_t rgnn_cluster_workspace_bytes(int n_nodes) {
    return cluster_ints(n_nodes) * sizeof(int) + cub_sort_bytes(n_nodes > 0 ? n_nodes : 1) + 256;
}

extern "C" int rgnn_cluster_links(const float* xy, const int32_t* und_a, const int32_t* und_b, const float* link_logits, int n_nodes,
                                  int n_und, float eps, int32_t* cluster_id, int32_t* n_clusters_out, int32_t* cl_ptr,
                                  int32_t* cl_members, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    RGNN_REQUIRE(n_nodes > 0, "cluster: empty graph");
    RGNN_REQUIRE(workspace_bytes >= rgnn_cluster_workspace_bytes(n_nodes), "cluster workspace too small");
    int* parent = static_cast<int*>(workspace);
    const int nb = (n_nodes + 255) / 256 > 1184 ? 1184 : (n_nodes + 255) / 256;
    uf_init_kernel<<<nb, 256, 0, stream>>>(parent, n_nodes);
    if (n_und > 0) {
        const int eb = (n_und + 255) / 256 > 1184 ? 1184 : (n_und + 255) / 256;
        uf_links_kernel<<<eb, 256, 0, stream>>>(xy, und_a, und_b, link_logits, n_und, eps, parent);
    }
    return finish_clusters(parent, n_nodes, workspace_bytes, cluster_id, n_clusters_out, cl_ptr, cl_members, stream);
}

extern "C" int rgnn_cluster_radius(const float* xy, const int32_t* frame_ptr_dev, int n_frames, int n_nodes, int max_frame_nodes,
                                   float eps, int32_t* cluster_id, int32_t* n_clusters_out, int32_t* cl_ptr, int32_t* cl_members,
                                   void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    RGNN_REQUIRE(n_nodes > 0 && n_frames > 0, "cluster: empty graph");
    RGNN_REQUIRE(workspace_bytes >= rgnn_cluster_workspace_bytes(n_nodes), "cluster workspace too small");
    int* parent = static_cast<int*>(workspace);
    const int nb = (n_nodes + 255) / 256 > 1184 ? 1184 : (n_nodes + 255) / 256;
    uf_init_kernel<<<nb, 256, 0, stream>>>(parent, n_nodes);
    int gx = (max_frame_nodes + 255) / 256;
    if (gx < 1) gx = 1;
    if (gx > 64) gx = 64;
    uf_radius_kernel<<<dim3(gx, n_frames), 256, 0, stream>>>(xy, frame_ptr_dev, n_frames, eps, parent);
    return finish_clusters(parent, n_nodes, workspace_bytes, cluster_id, n_clusters_out, cl_ptr, cl_members, stream);
}

// ---------------------------------------------------------------------------------------------
// Proposals from clusters (reference modules/inference/inference.py:23-47 and output.py:111-118): per cluster the sample
// mean of the member positions, the sample covariance of (mean - x) plus the measurement-noise covariance (noise only for
// a single member), the size, and the majority vote over the members' arg-max segmentation class (torch.bincount +
// argmax: ties go to the lowest class).  One thread per cluster, members visited in list order with separately rounded
// float32 operations: NumPy's axis-0 reductions over (n,2) / (n,2,2) arrays add row by row in that order.
// ---------------------------------------------------------------------------------------------
namespace rgnn {
__global__ void proposals_kernel(const float* __restrict__ px, const float* __restrict__ py, const float* __restrict__ node_cls,
                                 int n_classes, const int* __restrict__ cl_ptr, const int* __restrict__ cl_members, int n_clusters,
                                 float noise_xx, float noise_xy, float noise_yy, float* __restrict__ mean, float* __restrict__ cov,
                                 int* __restrict__ size, int* __restrict__ vote) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_clusters; c += gridDim.x * blockDim.x) {
        const int m0 = cl_ptr[c], m1 = cl_ptr[c + 1], n = m1 - m0;
        float sx = 0.f, sy = 0.f;
        int cnt[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) cnt[k] = 0;
        for (int m = m0; m < m1; ++m) {
            const int i = cl_members[m];
            sx = __fadd_rn(sx, px[i]);
            sy = __fadd_rn(sy, py[i]);
            if (node_cls != nullptr) {
                const float* l = node_cls + (size_t)i * n_classes;
                int best = 0;
                for (int k = 1; k < n_classes; ++k) if (l[k] > l[best]) best = k;
#pragma unroll
                for (int k = 0; k < 16; ++k) cnt[k] += (k == best);
            }
        }
        const float mx = __fdiv_rn(sx, (float)n), my = __fdiv_rn(sy, (float)n);
        float cxx = 0.f, cxy = 0.f, cyy = 0.f;
        if (n > 1) {
            for (int m = m0; m < m1; ++m) {
                const int i = cl_members[m];
                const float ex = __fsub_rn(mx, px[i]), ey = __fsub_rn(my, py[i]);
                cxx = __fadd_rn(cxx, __fmul_rn(ex, ex));
                cxy = __fadd_rn(cxy, __fmul_rn(ex, ey));
                cyy = __fadd_rn(cyy, __fmul_rn(ey, ey));
            }
            const float d = (float)(n - 1);
            cxx = __fadd_rn(__fdiv_rn(cxx, d), noise_xx);
            cxy = __fadd_rn(__fdiv_rn(cxy, d), noise_xy);
            cyy = __fadd_rn(__fdiv_rn(cyy, d), noise_yy);
        } else {
            cxx = noise_xx; cxy = noise_xy; cyy = noise_yy;
        }
        mean[2 * c] = mx; mean[2 * c + 1] = my;
        cov[4 * c] = cxx; cov[4 * c + 1] = cxy; cov[4 * c + 2] = cxy; cov[4 * c + 3] = cyy;
        size[c] = n;
        if (vote != nullptr) {
            int best = 0;
#pragma unroll
            for (int k = 1; k < 16; ++k) if (k < n_classes && cnt[k] > cnt[best]) best = k;
            vote[c] = best;
        }
    }
}
}  // namespace rgnn

extern "C" int rgnn_cluster_proposals(const float* px, const float* py, const float* node_cls, int n_classes, const int32_t* cl_ptr,
                                      const int32_t* cl_members, int n_clusters, const float* noise_cov_host, float* mean, float* cov,
                                      int32_t* size, int32_t* vote, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n_clusters <= 0) return RGNN_OK;
    RGNN_REQUIRE(node_cls == nullptr || (n_classes >= 1 && n_classes <= 16), "proposals: %d classes", n_classes);
    const int blocks = (n_clusters + 255) / 256 > 1184 ? 1184 : (n_clusters + 255) / 256;
    proposals_kernel<<<blocks, 256, 0, stream>>>(px, py, node_cls, n_classes, cl_ptr, cl_members, n_clusters, noise_cov_host[0],
                                                 noise_cov_host[1], noise_cov_host[3], mean, cov, size, vote);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

// ---------------------------------------------------------------------------------------------
// Per-cluster max over member rows (object_classification used on its own, gnn_blocks.py:384-386: torch.max(x[idx], dim=0) per cluster):
// one warp per cluster, lane = channel (mod 32); the row that supplied each maximum (first in member order on ties) is kept for the
// backward, which routes d pooled to exactly that row like autograd of torch.max(dim).  The detector-level forward fuses this
// reduction into the class-head kernel; this is the block-level API path.
// ---------------------------------------------------------------------------------------------
namespace rgnn {
__global__ void segment_max_kernel(const float* __restrict__ x, int W, const int* __restrict__ cl_ptr, const int* __restrict__ members,
                                   int n_clusters, float* __restrict__ pooled, int* __restrict__ argrow) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int c = blockIdx.x * wpb + (threadIdx.x >> 5); c < n_clusters; c += gridDim.x * wpb) {
        const int m0 = __ldg(cl_ptr + c), m1 = __ldg(cl_ptr + c + 1);
        for (int col = lane; col < W; col += 32) {
            float best = -INFINITY;
            int arg = -1;
            for (int m = m0; m < m1; ++m) {
                const int r = __ldg(members + m);
                const float v = __ldg(x + (size_t)r * W + col);
                if (v > best || arg < 0) { best = v; arg = r; }       // first maximum in member order; NaN-free inputs
            }
            pooled[(size_t)c * W + col] = best;
            if (argrow != nullptr) argrow[(size_t)c * W + col] = arg;
        }
    }
}
__global__ void segment_max_bwd_kernel(const float* __restrict__ d_pooled, const int* __restrict__ argrow, long long total, int W,
                                       float* __restrict__ dx) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = __ldg(argrow + i);
        if (r >= 0) atomicAdd(dx + (size_t)r * W + (int)(i % W), __ldg(d_pooled + i));      // clusters may share nodes: accumulate
    }
}
}  // namespace rgnn

extern "C" int rgnn_segment_max_fwd(const float* x, int width, const int32_t* cl_ptr, const int32_t* cl_members, int n_clusters,
                                    float* pooled, int32_t* argrow, void* stream_) {
    using namespace rgnn;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n_clusters <= 0) return RGNN_OK;
    RGNN_REQUIRE(width >= 1, "segment_max: width %d", width);
    const int wpb = 8;
    const int blocks = (n_clusters + wpb - 1) / wpb;
    segment_max_kernel<<<blocks > 8 * 148 ? 8 * 148 : blocks, 32 * wpb, 0, stream>>>(x, width, cl_ptr, cl_members, n_clusters, pooled, argrow);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

extern "C" int rgnn_segment_max_bwd(const float* d_pooled, const int32_t* argrow, int n_clusters, int width, int n_rows, float* dx,
                                    void* stream_) {
    using namespace rgnn;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    RGNN_CHECK_CUDA(cudaMemsetAsync(dx, 0, (size_t)n_rows * width * sizeof(float), stream));
    const long long total = (long long)n_clusters * width;
    if (total <= 0) return RGNN_OK;
    const long long blocks = (total + 255) / 256;
    segment_max_bwd_kernel<<<(unsigned)(blocks > 8 * 148 ? 8 * 148 : blocks), 256, 0, stream>>>(d_pooled, argrow, total, width, dx);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}
