// Multi-task loss forward + gradient w.r.t. the logits, accuracies, and the SGD update.
// Replaces reference modules/neural_net/gnn/loss.py:37-76 (Loss_Graph.forward), lossfunc.py:19-55,
// gnn_detector.py:23-28 (compute_accuracy) and torch.optim.SGD (set_param_for_training_gnn.py:46).
#include "rgnn_common.cuh"

namespace rgnn {

struct LossArgs {
    rgnn_loss_cfg cfg;
    const float *node_cls, *node_off, *link_cls, *obj_cls;
    const int64_t *node_gt, *link_gt, *obj_gt;
    const float* off_gt;
    int n_nodes, n_und, n_clusters;
    float inv_nodes, inv_und, inv_clusters;
    float *g_node_cls, *g_node_off, *g_link, *g_obj;
    double* losses;
    int* correct;
    const double* counts_dev;   // optional (3): device-resident global counts (all-reduced on the stream); overrides inv_*
    float* dp_tail;             // optional (5): [0] = 1 when a loss term is NaN, [1..4] += this rank's weighted loss shares
};

__device__ __forceinline__ float softplus(float z) { return fmaxf(z, 0.f) + log1pf(expf(-fabsf(z))); }

// soft-target cross entropy with one-hot target (lossfunc.py:24-26): -w_y log_softmax(z)_y
__device__ __forceinline__ float ce_row(const float* __restrict__ z, int C, int y, float w, float gscale,
                                        float* __restrict__ g, bool* hit) {
    float m = z[0];
    int am = 0;
    for (int c = 1; c < C; ++c)
        if (z[c] > m) { m = z[c]; am = c; }
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(z[c] - m);
    const float lse = m + logf(s);
    if (g != nullptr)
        for (int c = 0; c < C; ++c) g[c] = gscale * w * (expf(z[c] - lse) - (c == y ? 1.f : 0.f));
    *hit = am == y;
    return w * (lse - z[y]);
}

// torchvision.ops.sigmoid_focal_loss on a one-hot target, summed over the classes (loss.py:56-58)
__device__ __forceinline__ float focal_row(const float* __restrict__ z, int C, int y, float alpha, float gamma,
                                           float gscale, float* __restrict__ g, bool* hit) {
    float tot = 0.f, m = z[0];
    int am = 0;
    for (int c = 0; c < C; ++c) {
        if (z[c] > m) { m = z[c]; am = c; }
        const float p = 1.f / (1.f + expf(-z[c]));
        float l, d;
        if (c == y) {
            const float nlogp = softplus(-z[c]);                 // -log p
            const float q = 1.f - p;
            const float qg = gamma == 2.f ? q * q : powf(q, gamma);
            l = alpha * qg * nlogp;
            d = alpha * qg * (-gamma * p * nlogp - q);
        } else {
            const float nlogq = softplus(z[c]);                  // -log(1-p)
            const float pg = gamma == 2.f ? p * p : powf(p, gamma);
            l = (1.f - alpha) * pg * nlogq;
            d = (1.f - alpha) * pg * (gamma * (1.f - p) * nlogq + p);
        }
        tot += l;
        if (g != nullptr) g[c] = gscale * d;
    }
    *hit = am == y;
    return tot;
}

__global__ void loss_kernel(const __grid_constant__ LossArgs a) {
    __shared__ double sh[4][8];
    __shared__ int shc[3][8];
    double acc[4] = {0., 0., 0., 0.};
    int hits[3] = {0, 0, 0};
    const int stride = gridDim.x * blockDim.x;
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int C = a.cfg.n_classes, CE = a.cfg.n_edge_classes;
    float inv_nodes = a.inv_nodes, inv_und = a.inv_und, inv_clusters = a.inv_clusters;
    if (a.counts_dev != nullptr) {      // data parallel: the counts were summed over the ranks on this stream, never seen by the host
        const double cn = a.counts_dev[0], cu = a.counts_dev[1], cc = a.counts_dev[2];
        inv_nodes = cn > 0. ? (float)(1.0 / cn) : 0.f;
        inv_und = cu > 0. ? (float)(1.0 / cu) : 0.f;
        inv_clusters = cc > 0. ? (float)(1.0 / cc) : 0.f;
    }
    for (int i = t0; i < a.n_nodes; i += stride) {
        bool hit;
        const int y = (int)a.node_gt[i];
        acc[0] += ce_row(a.node_cls + (size_t)i * C, C, y, a.cfg.class_weights[y], a.cfg.w_node_cls * inv_nodes,
                         a.g_node_cls ? a.g_node_cls + (size_t)i * C : nullptr, &hit);
        hits[0] += hit;
        const float dx = a.node_off[2 * i] - a.off_gt[2 * i], dy = a.node_off[2 * i + 1] - a.off_gt[2 * i + 1];
        acc[1] += 0.5f * (dx * dx + dy * dy);
        if (a.g_node_off) {
            a.g_node_off[2 * i] = a.cfg.w_node_reg * inv_nodes * dx;
            a.g_node_off[2 * i + 1] = a.cfg.w_node_reg * inv_nodes * dy;
        }
    }
    for (int i = t0; i < a.n_und; i += stride) {
        bool hit;
        acc[2] += focal_row(a.link_cls + (size_t)i * CE, CE, (int)a.link_gt[i], a.cfg.focal_alpha, a.cfg.focal_gamma,
                            a.cfg.w_edge_cls * inv_und, a.g_link ? a.g_link + (size_t)i * CE : nullptr, &hit);
        hits[1] += hit;
    }
    for (int i = t0; i < a.n_clusters; i += stride) {
        bool hit;
        acc[3] += ce_row(a.obj_cls + (size_t)i * C, C, (int)a.obj_gt[i], 1.f, a.cfg.w_obj_cls * inv_clusters,
                         a.g_obj ? a.g_obj + (size_t)i * C : nullptr, &hit);
        hits[2] += hit;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = 0; k < 4; ++k) {
        double v = acc[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sh[k][warp] = v;
    }
    for (int k = 0; k < 3; ++k) {
        int v = hits[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) shc[k][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double v = 0.;
        for (int w = 0; w < 8; ++w) v += sh[threadIdx.x][w];
        const double scale = threadIdx.x == 0 ? (double)a.cfg.w_node_cls * inv_nodes
                           : threadIdx.x == 1 ? (double)a.cfg.w_node_reg * inv_nodes
                           : threadIdx.x == 2 ? (double)a.cfg.w_edge_cls * inv_und
                                              : (double)a.cfg.w_obj_cls * inv_clusters;
        atomicAdd(a.losses + threadIdx.x, v * scale);
        if (a.dp_tail != nullptr) {
            // a NaN total has a NaN partial: the reference's skip_batch test (gnn/training.py:40-45) without a host read
            if (v != v) a.dp_tail[0] = 1.f;
            atomicAdd(a.dp_tail + 1 + threadIdx.x, (float)(v * scale));
        }
    } else if (threadIdx.x >= 32 && threadIdx.x < 35) {
        int v = 0;
        for (int w = 0; w < 8; ++w) v += shc[threadIdx.x - 32][w];
        atomicAdd(a.correct + (threadIdx.x - 32), v);
    }
}

__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, size_t n, float lr,
                           float mu, float wd, float gscale, int first, const float* __restrict__ skip) {
    // device-side version of the reference's skip_batch (gnn/training.py:40-45): a non-zero or NaN flag leaves parameters
    // and momentum untouched, without the host ever reading the loss
    if (skip != nullptr) {
        const float f = __ldg(skip);
        if (f != 0.f || f != f) return;
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float pi = p[i];
        const float gi = fmaf(wd, pi, g[i] * gscale);
        const float b = first ? gi : fmaf(mu, buf[i], gi);
        buf[i] = b;
        p[i] = pi - lr * b;
    }
}

}  // namespace rgnn

using namespace rgnn;

extern "C" int rgnn_losses_fwdbwd(const rgnn_loss_cfg* cfg, const float* node_cls, const float* node_off,
                                  const float* link_cls, const float* obj_cls, const int64_t* node_cls_gt,
                                  const float* node_off_gt, const int64_t* link_gt, const int64_t* obj_gt, int n_nodes,
                                  int n_und, int n_clusters, double count_nodes, double count_und, double count_clusters,
                                  float* grad_node_cls, float* grad_node_off, float* grad_link_cls, float* grad_obj_cls,
                                  double* losses_out, int32_t* correct_out, const double* counts_dev, float* dp_tail,
                                  void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    RGNN_REQUIRE(cfg->n_classes >= 1 && cfg->n_classes <= 16 && cfg->n_edge_classes >= 1 && cfg->n_edge_classes <= 16,
                 "losses: class counts out of range");
    LossArgs a;
    a.cfg = *cfg;
    a.node_cls = node_cls; a.node_off = node_off; a.link_cls = link_cls; a.obj_cls = obj_cls;
    a.node_gt = node_cls_gt; a.link_gt = link_gt; a.obj_gt = obj_gt; a.off_gt = node_off_gt;
    a.n_nodes = n_nodes; a.n_und = n_und; a.n_clusters = n_clusters;
    a.inv_nodes = count_nodes > 0 ? (float)(1.0 / count_nodes) : 0.f;
    a.inv_und = count_und > 0 ? (float)(1.0 / count_und) : 0.f;
    a.inv_clusters = count_clusters > 0 ? (float)(1.0 / count_clusters) : 0.f;
    a.g_node_cls = grad_node_cls; a.g_node_off = grad_node_off; a.g_link = grad_link_cls; a.g_obj = grad_obj_cls;
    a.losses = losses_out; a.correct = correct_out;
    a.counts_dev = counts_dev; a.dp_tail = dp_tail;
    RGNN_CHECK_CUDA(cudaMemsetAsync(losses_out, 0, 4 * sizeof(double), stream));
    RGNN_CHECK_CUDA(cudaMemsetAsync(correct_out, 0, 3 * sizeof(int32_t), stream));
    int m = n_nodes > n_und ? n_nodes : n_und;
    m = m > n_clusters ? m : n_clusters;
    int grid = (m + 255) / 256;
    grid = grid < 1 ? 1 : (grid > 4 * sm_count() ? 4 * sm_count() : grid);
    loss_kernel<<<grid, 256, 0, stream>>>(a);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

extern "C" int rgnn_sgd_step(float* params, const float* grads, float* momentum_buf, size_t n, float lr, float momentum,
                             float weight_decay, float grad_scale, int first_step, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n == 0) return RGNN_OK;
    size_t grid = (n + 255) / 256;
    if (grid > (size_t)4 * sm_count()) grid = (size_t)4 * sm_count();
    sgd_kernel<<<(unsigned)grid, 256, 0, stream>>>(params, grads, momentum_buf, n, lr, momentum, weight_decay, grad_scale, first_step, nullptr);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

extern "C" int rgnn_sgd_step_guarded(float* params, const float* grads, float* momentum_buf, size_t n, float lr, float momentum,
                                     float weight_decay, float grad_scale, int first_step, const float* skip_flag, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n == 0) return RGNN_OK;
    size_t grid = (n + 255) / 256;
    if (grid > (size_t)4 * sm_count()) grid = (size_t)4 * sm_count();
    sgd_kernel<<<(unsigned)grid, 256, 0, stream>>>(params, grads, momentum_buf, n, lr, momentum, weight_decay, grad_scale, first_step, skip_flag);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}
