// Graph construction on the GPU: batched exact kNN + radius degree, symmetrised row-major adjacency,
// CSR utilities and the raw node / edge features.
// Replaces reference modules/compute_features/graph_features.py (compute_adjacency_information :58-84,
// _v2 :87-114, compute_knn :25-44, compute_ball_query :11-22, compute_node_features :117-144,
// compute_edge_features :147-164) without ever forming the N x N matrices.
#include <float.h>
#include <limits.h>

#include "rgnn_common.cuh"

namespace rgnn {

// ---------------------------------------------------------------------------------------------
// device-wide exclusive scan of int32 (three small kernels; n up to 2^31)
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_BLOCK = 1024;   // elements per block (256 threads x 4)

__global__ void scan_reduce_kernel(const int* __restrict__ in, int n, int* __restrict__ block_sums) {
    __shared__ int sh[8];
    const int base = blockIdx.x * SCAN_BLOCK;
    int s = 0;
    for (int i = threadIdx.x; i < SCAN_BLOCK; i += 256) {
        const int g = base + i;
        if (g < n) s += in[g];
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += sh[w];
        block_sums[blockIdx.x] = t;
    }
}

__global__ void scan_sums_kernel(int* __restrict__ block_sums, int nb, int* __restrict__ total_out) {
    // single block: sequential over chunks of 1024 with an in-block scan
    __shared__ int sh[1024];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nb ? block_sums[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const int t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        const int incl = sh[threadIdx.x];
        if (i < nb) block_sums[i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += incl;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out != nullptr) *total_out = carry;
}

__global__ void scan_apply_kernel(const int* __restrict__ in, int n, const int* __restrict__ block_sums,
                                  int* __restrict__ out /* n+1 entries; out[n] = total */) {
    __shared__ int sh[256];
    const int base = blockIdx.x * SCAN_BLOCK + threadIdx.x * 4;
    int v[4];
    int s = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        v[q] = (base + q < n) ? in[base + q] : 0;
        s += v[q];
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        const int t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    int run = block_sums[blockIdx.x] + sh[threadIdx.x] - s;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (base + q < n) out[base + q] = run;
        run += v[q];
        if (base + q == n - 1) out[n] = run;
    }
}

size_t scan_ws_ints(int n) { return (size_t)(n + SCAN_BLOCK - 1) / SCAN_BLOCK + 1; }

// out has n+1 entries.  in and out may alias only if identical pointers are NOT used (they must differ).
int exclusive_scan(const int* in, int n, int* out, int* ws, cudaStream_t stream) {
    if (n <= 0) {
        RGNN_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(int), stream));
        return RGNN_OK;
    }
    const int nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    scan_reduce_kernel<<<nb, 256, 0, stream>>>(in, n, ws);
    scan_sums_kernel<<<1, 1024, 0, stream>>>(ws, nb, nullptr);
    scan_apply_kernel<<<nb, 256, 0, stream>>>(in, n, ws, out);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

// ---------------------------------------------------------------------------------------------
// phase A: exact k nearest (by (d2, index)) + radius count, brute force per frame with smem tiles
// ---------------------------------------------------------------------------------------------
constexpr int KNN_THREADS = 128;
constexpr int KNN_TILE = 512;

__device__ __forceinline__ float dist2(float xi, float yi, float xj, float yj) {
    // fl(fl(dx*dx) + fl(dy*dy)): what NumPy's batched (1x2)@(2x1) float32 matmul yields (graph_features.py:70-75)
    const float dx = __fsub_rn(xi, xj), dy = __fsub_rn(yi, yj);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

template <int KCAP>
__global__ void __launch_bounds__(KNN_THREADS) knn_kernel(const float* __restrict__ px, const float* __restrict__ py,
                                                          const int* __restrict__ frame_ptr, float eps2, int knn,
                                                          int ks /* row stride of knn_idx */,
                                                          int* __restrict__ knn_idx, int* __restrict__ degree) {
    // candidates of a tile as {x, y} pairs: one 16-byte shared-memory broadcast serves two candidates (the loop is bound by
    // instruction issue: 2 LDS + 9 ALU per candidate before, 0.5 LDS + 8 ALU now); candidates are still visited in index order
    __shared__ float4 sxy[KNN_TILE / 2];
    float2* sxy2 = reinterpret_cast<float2*>(sxy);
    const int f = blockIdx.y;
    const int f0 = frame_ptr[f], f1 = frame_ptr[f + 1];
    const int nf = f1 - f0;
    if ((int)blockIdx.x * KNN_THREADS >= nf) return;
    const int li = blockIdx.x * KNN_THREADS + threadIdx.x;
    const bool active = li < nf;
    const int gi = f0 + li;
    const float xi = active ? px[gi] : 0.f, yi = active ? py[gi] : 0.f;
    float bd[KCAP];
    int bj[KCAP];
#pragma unroll
    for (int q = 0; q < KCAP; ++q) { bd[q] = FLT_MAX; bj[q] = -1; }
    int deg = -1;                      // the point itself (d = 0 <= eps2) is counted by the loop and taken out here
    auto visit = [&](float xj, float yj, int j) {
        const float d = dist2(xi, yi, xj, yj);
        deg += d <= eps2 ? 1 : 0;
        if (d < bd[KCAP - 1]) {
#pragma unroll
            for (int q = KCAP - 1; q > 0; --q) {
                if (d < bd[q - 1]) { bd[q] = bd[q - 1]; bj[q] = bj[q - 1]; }
                else if (d < bd[q]) { bd[q] = d; bj[q] = j; }
            }
            if (d < bd[0]) { bd[0] = d; bj[0] = j; }
        }
    };
    for (int t0 = 0; t0 < nf; t0 += KNN_TILE) {
        const int tn = min(KNN_TILE, nf - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < tn; i += KNN_THREADS) sxy2[i] = make_float2(px[f0 + t0 + i], py[f0 + t0 + i]);
        __syncthreads();
        if (active) {
            int jj = 0;
#pragma unroll 2
            for (; jj + 1 < tn; jj += 2) {
                const float4 c = sxy[jj >> 1];
                visit(c.x, c.y, t0 + jj);
                visit(c.z, c.w, t0 + jj + 1);
            }
            if (jj < tn) { const float2 c = sxy2[jj]; visit(c.x, c.y, t0 + jj); }
        }
    }
    if (active) {
        degree[gi] = max(deg, 0);      // (eps2 < 0 or a NaN position: the point did not count itself either)
        const int kp1 = knn >= nf ? nf : knn + 1;   // graph_features.py:35
#pragma unroll
        for (int q = 0; q < KCAP; ++q)
            if (q < ks) knn_idx[(size_t)gi * ks + q] = (q < kp1) ? f0 + bj[q] : -1;
    }
}

// ---------------------------------------------------------------------------------------------
// phase A with a uniform grid: the same exact result (k + 1 nearest by (d2, index), radius count), candidates from the cells
// around the query instead of the whole frame.
//   cell_sort_kernel (one CTA per frame): bounding box -> square cells (about 8 points per cell, at most 32 x 32) -> counting
//   sort of the frame's points by cell; a sorted entry is {x, y, local index} (16 bytes: one load per candidate).
//   knn_grid_kernel: a thread owns a query (taken in SORTED order, so the lanes of a warp sit in the same or adjacent cells and
//   read the same candidates: broadcast loads) and visits the cells ring by ring.  After ring r every point outside the
//   (2r+1) x (2r+1) block is at least `bound` away (distance to the nearest side of the block that is not the grid's border); the
//   search stops when the worst kept candidate and the radius both lie strictly inside bound (less 1e-4 of a cell, which covers the
//   float rounding of the cell assignment and of d2).  The distance is the reference's fl(fl(dx^2) + fl(dy^2)) and the kept list is
//   ordered by (d2, index) with an explicit index tie-break, so the result is bit-identical to the brute-force kernel whatever the
//   order in which candidates are visited.
// ---------------------------------------------------------------------------------------------
constexpr int GRID_MAXN = 32;
constexpr int GRID_CELLS = GRID_MAXN * GRID_MAXN;
struct FrameGrid { float x0, y0, cell, inv_cell; int nx, ny, pad0, pad1; };

__device__ __forceinline__ int grid_coord(float v, float v0, float inv_cell, int n) {
    const float t = (v - v0) * inv_cell;
    int c = t > 0.f ? (t < (float)n ? (int)t : n - 1) : 0;      // (a NaN position lands in cell 0)
    return c;
}

__global__ void __launch_bounds__(256) cell_sort_kernel(const float* __restrict__ px, const float* __restrict__ py,
                                                        const int* __restrict__ frame_ptr, FrameGrid* __restrict__ grids,
                                                        int* __restrict__ cell_start, float4* __restrict__ sorted) {
    __shared__ int cnt[GRID_CELLS + 1];
    __shared__ int cur[GRID_CELLS];
    __shared__ float red[4][8];
    __shared__ FrameGrid gsh;
    const int f = blockIdx.x;
    const int f0 = frame_ptr[f], nf = frame_ptr[f + 1] - f0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float xmin = FLT_MAX, xmax = -FLT_MAX, ymin = FLT_MAX, ymax = -FLT_MAX;
    for (int i = tid; i < nf; i += 256) {
        const float x = px[f0 + i], y = py[f0 + i];
        xmin = fminf(xmin, x); xmax = fmaxf(xmax, x); ymin = fminf(ymin, y); ymax = fmaxf(ymax, y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        xmin = fminf(xmin, __shfl_xor_sync(0xffffffffu, xmin, o)); xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymin = fminf(ymin, __shfl_xor_sync(0xffffffffu, ymin, o)); ymax = fmaxf(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    if (lane == 0) { red[0][warp] = xmin; red[1][warp] = xmax; red[2][warp] = ymin; red[3][warp] = ymax; }
    for (int i = tid; i <= GRID_CELLS; i += 256) cnt[i] = 0;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 8; ++w) {
            xmin = fminf(xmin, red[0][w]); xmax = fmaxf(xmax, red[1][w]); ymin = fminf(ymin, red[2][w]); ymax = fmaxf(ymax, red[3][w]);
        }
        int n = (int)ceilf(sqrtf((float)nf * 0.125f));
        n = n < 1 ? 1 : (n > GRID_MAXN ? GRID_MAXN : n);
        const float w = xmax - xmin, h = ymax - ymin;
        float cell = fmaxf(w, h) / (float)n;
        if (!(cell > 1e-12f) || !(cell < 1e30f)) cell = 1.f;      // all points coincide (or non-finite input): one column of cells
        cell *= 1.0001f;
        FrameGrid g;
        g.x0 = xmin; g.y0 = ymin; g.cell = cell; g.inv_cell = 1.f / cell;
        g.nx = min(n, (int)(w * g.inv_cell) + 1);
        g.ny = min(n, (int)(h * g.inv_cell) + 1);
        if (g.nx < 1) g.nx = 1;
        if (g.ny < 1) g.ny = 1;
        g.pad0 = g.pad1 = 0;
        gsh = g;
        grids[f] = g;
    }
    __syncthreads();
    const FrameGrid g = gsh;
    const int ncell = g.nx * g.ny;
    for (int i = tid; i < nf; i += 256) {
        const int c = grid_coord(py[f0 + i], g.y0, g.inv_cell, g.ny) * g.nx + grid_coord(px[f0 + i], g.x0, g.inv_cell, g.nx);
        atomicAdd(&cnt[c], 1);
    }
    __syncthreads();
    if (warp == 0) {        // exclusive scan of <= 1024 counters by one warp: 32 per lane, then a warp scan of the lane totals
        int base = lane * 32, tot = 0;
        for (int k = 0; k < 32; ++k) tot += (base + k < ncell) ? cnt[base + k] : 0;
        int incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        int run = incl - tot;
        for (int k = 0; k < 32; ++k) {
            if (base + k < ncell) {
                const int c = cnt[base + k];
                cnt[base + k] = run;
                cur[base + k] = run;
                run += c;
            }
        }
        if (lane == 31) cnt[ncell] = nf;
    }
    __syncthreads();
    for (int i = tid; i <= ncell; i += 256) cell_start[(size_t)f * (GRID_CELLS + 1) + i] = cnt[i];
    for (int i = tid; i < nf; i += 256) {
        const float x = px[f0 + i], y = py[f0 + i];
        const int c = grid_coord(y, g.y0, g.inv_cell, g.ny) * g.nx + grid_coord(x, g.x0, g.inv_cell, g.nx);
        const int pos = atomicAdd(&cur[c], 1);
        sorted[f0 + pos] = make_float4(x, y, __int_as_float(i), 0.f);
    }
}

template <int KCAP>
__global__ void __launch_bounds__(KNN_THREADS) knn_grid_kernel(const float4* __restrict__ sorted, const int* __restrict__ frame_ptr,
                                                               const FrameGrid* __restrict__ grids, const int* __restrict__ cell_start,
                                                               float eps2, int knn, int ks, int* __restrict__ knn_idx,
                                                               int* __restrict__ degree) {
    const int f = blockIdx.y;
    const int f0 = frame_ptr[f], nf = frame_ptr[f + 1] - f0;
    const int sp = blockIdx.x * KNN_THREADS + threadIdx.x;
    if (sp >= nf) return;
    const FrameGrid g = grids[f];
    const int* __restrict__ cs = cell_start + (size_t)f * (GRID_CELLS + 1);
    const float4* __restrict__ pts = sorted + f0;
    const float4 me = pts[sp];
    const float xi = me.x, yi = me.y;
    const int li = __float_as_int(me.z);
    const int cx = grid_coord(xi, g.x0, g.inv_cell, g.nx), cy = grid_coord(yi, g.y0, g.inv_cell, g.ny);
    float bd[KCAP];
    int bj[KCAP];
#pragma unroll
    for (int q = 0; q < KCAP; ++q) { bd[q] = FLT_MAX; bj[q] = INT_MAX; }
    int deg = -1;
    auto before = [](float d, int j, float d2, int j2) { return d < d2 || (d == d2 && j < j2); };
    auto visit_span = [&](int p0, int p1) {
        for (int p = p0; p < p1; ++p) {
            const float4 c = __ldg(pts + p);
            const int j = __float_as_int(c.z);
            const float d = dist2(xi, yi, c.x, c.y);
            deg += d <= eps2 ? 1 : 0;
            if (before(d, j, bd[KCAP - 1], bj[KCAP - 1])) {
#pragma unroll
                for (int q = KCAP - 1; q > 0; --q) {
                    if (before(d, j, bd[q - 1], bj[q - 1])) { bd[q] = bd[q - 1]; bj[q] = bj[q - 1]; }
                    else if (before(d, j, bd[q], bj[q])) { bd[q] = d; bj[q] = j; }
                }
                if (before(d, j, bd[0], bj[0])) { bd[0] = d; bj[0] = j; }
            }
        }
    };
    const float slack = 1e-4f * g.cell;
    for (int r = 0;; ++r) {
        const int xlo = max(cx - r, 0), xhi = min(cx + r, g.nx - 1);
        for (int dy = -r; dy <= r; ++dy) {
            const int yy = cy + dy;
            if (yy < 0 || yy >= g.ny) continue;
            if (dy == -r || dy == r) {
                visit_span(cs[yy * g.nx + xlo], cs[yy * g.nx + xhi + 1]);       // a whole row of the block: contiguous in the sorted order
            } else {
                if (cx - r >= 0) visit_span(cs[yy * g.nx + cx - r], cs[yy * g.nx + cx - r + 1]);
                if (cx + r < g.nx) visit_span(cs[yy * g.nx + cx + r], cs[yy * g.nx + cx + r + 1]);
            }
        }
        const bool open_l = cx - r > 0, open_r = cx + r < g.nx - 1, open_d = cy - r > 0, open_u = cy + r < g.ny - 1;
        if (!(open_l || open_r || open_d || open_u)) break;         // the block covers the whole grid
        float bound = FLT_MAX;
        if (open_l) bound = fminf(bound, xi - (g.x0 + (float)(cx - r) * g.cell));
        if (open_r) bound = fminf(bound, (g.x0 + (float)(cx + r + 1) * g.cell) - xi);
        if (open_d) bound = fminf(bound, yi - (g.y0 + (float)(cy - r) * g.cell));
        if (open_u) bound = fminf(bound, (g.y0 + (float)(cy + r + 1) * g.cell) - yi);
        bound -= slack;
        if (bound > 0.f) {
            const float b2 = bound * bound * 0.9999f;
            if (bd[KCAP - 1] < b2 && eps2 < b2) break;
        }
    }
    const int gi = f0 + li;
    degree[gi] = max(deg, 0);
    const int kp1 = knn >= nf ? nf : knn + 1;
#pragma unroll
    for (int q = 0; q < KCAP; ++q)
        if (q < ks) knn_idx[(size_t)gi * ks + q] = (q < kp1) ? f0 + bj[q] : -1;
}

static int g_knn_grid = 1;
int graph_set_option(const char* name, int value) {
    if (strcmp(name, "knn_grid") == 0 && (value == 0 || value == 1)) { g_knn_grid = value; return 1; }
    return 0;
}
int graph_get_option(const char* name) { return strcmp(name, "knn_grid") == 0 ? g_knn_grid : -2; }

// ---------------------------------------------------------------------------------------------
// phase B: symmetrise.  row(i) = (kNN(i) \ {i})  U  {j : i in kNN(j)}   [U radius(i) for _v2]
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool in_list(const int* __restrict__ lst, int n, int v) {
    for (int q = 0; q < n; ++q)
        if (lst[q] == v) return true;
    return false;
}

// own[i] = entries row i writes itself; extra[j] += entries pushed into row j by non-mutual neighbours
__global__ void sym_count_kernel(const float* __restrict__ px, const float* __restrict__ py,
                                 const int* __restrict__ knn_idx, int ks, const int* __restrict__ degree, int n,
                                 float eps2, int union_radius, int* __restrict__ own, int* __restrict__ extra) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int* mine = knn_idx + (size_t)i * ks;
    int cnt = union_radius ? degree[i] : 0;
    const float xi = px[i], yi = py[i];
    for (int q = 0; q < ks; ++q) {
        const int j = mine[q];
        if (j < 0) break;
        if (j == i) continue;
        const bool in_ball = union_radius && dist2(xi, yi, px[j], py[j]) <= eps2;
        if (!in_ball) {
            ++cnt;
            if (!in_list(knn_idx + (size_t)j * ks, ks, i)) atomicAdd(extra + j, 1);
        }
    }
    own[i] = cnt;
}

__global__ void add_kernel(const int* __restrict__ a, const int* __restrict__ b, int n, int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + b[i];
}

// _v2 only: list the radius neighbours of i (ascending j) at the head of its row
__global__ void __launch_bounds__(KNN_THREADS) radius_fill_kernel(const float* __restrict__ px, const float* __restrict__ py,
                                                                  const int* __restrict__ frame_ptr, float eps2,
                                                                  const int* __restrict__ row_ptr, int n_total, int capacity,
                                                                  int* __restrict__ col) {
    __shared__ float sx[KNN_TILE], sy[KNN_TILE];
    if (row_ptr[n_total] > capacity) return;
    const int f = blockIdx.y;
    const int f0 = frame_ptr[f], f1 = frame_ptr[f + 1];
    const int nf = f1 - f0;
    if ((int)blockIdx.x * KNN_THREADS >= nf) return;
    const int li = blockIdx.x * KNN_THREADS + threadIdx.x;
    const bool active = li < nf;
    const int gi = f0 + li;
    const float xi = active ? px[gi] : 0.f, yi = active ? py[gi] : 0.f;
    int w = active ? row_ptr[gi] : 0;
    for (int t0 = 0; t0 < nf; t0 += KNN_TILE) {
        const int tn = min(KNN_TILE, nf - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < tn; i += KNN_THREADS) { sx[i] = px[f0 + t0 + i]; sy[i] = py[f0 + t0 + i]; }
        __syncthreads();
        if (active)
            for (int jj = 0; jj < tn; ++jj)
                if (t0 + jj != li && dist2(xi, yi, sx[jj], sy[jj]) <= eps2) col[w++] = f0 + t0 + jj;
    }
}

__global__ void sym_fill_kernel(const float* __restrict__ px, const float* __restrict__ py,
                                const int* __restrict__ knn_idx, int ks, const int* __restrict__ degree,
                                const int* __restrict__ own, const int* __restrict__ row_ptr, int n, float eps2,
                                int union_radius, int capacity, int* __restrict__ cursor, int* __restrict__ col) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || row_ptr[n] > capacity) return;
    const int* mine = knn_idx + (size_t)i * ks;
    int w = row_ptr[i] + (union_radius ? degree[i] : 0);
    const float xi = px[i], yi = py[i];
    for (int q = 0; q < ks; ++q) {
        const int j = mine[q];
        if (j < 0) break;
        if (j == i) continue;
        const bool in_ball = union_radius && dist2(xi, yi, px[j], py[j]) <= eps2;
        if (!in_ball) {
            col[w++] = j;
            if (!in_list(knn_idx + (size_t)j * ks, ks, i)) {
                const int slot = atomicAdd(cursor + j, 1);
                col[row_ptr[j] + own[j] + slot] = i;
            }
        }
    }
}

// ascending insertion sort of every CSR row (rows are short: ~k..2k entries, radius rows a few dozen)
__global__ void sort_rows_kernel(const int* __restrict__ row_ptr, int n, int capacity, int* __restrict__ col,
                                 int* __restrict__ aux /* optional second array permuted alongside, or nullptr */) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || row_ptr[n] > capacity) return;
    const int a = row_ptr[i], b = row_ptr[i + 1];
    for (int p = a + 1; p < b; ++p) {
        const int v = col[p];
        const int x = aux ? aux[p] : 0;
        int q = p - 1;
        while (q >= a && col[q] > v) {
            col[q + 1] = col[q];
            if (aux) aux[q + 1] = aux[q];
            --q;
        }
        col[q + 1] = v;
        if (aux) aux[q + 1] = x;
    }
}

__global__ void copy_last_kernel(const int* __restrict__ row_ptr, int n, int* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = row_ptr[n];
}

// ---------------------------------------------------------------------------------------------
// finalize: row ids, reverse-edge permutation, undirected list
// ---------------------------------------------------------------------------------------------
__global__ void finalize_rows_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, int n,
                                     int* __restrict__ row_of_edge, int* __restrict__ perm, int* __restrict__ und_cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int a = row_ptr[i], b = row_ptr[i + 1];
    int cnt = 0;
    for (int k = a; k < b; ++k) {
        const int c = col[k];
        row_of_edge[k] = i;
        cnt += (c > i) ? 1 : 0;
        // position of the reverse edge (c -> i) inside row c (binary search, row sorted ascending)
        int lo = row_ptr[c], hi = row_ptr[c + 1] - 1, pos = -1;
        while (lo <= hi) {
            const int mid = (lo + hi) >> 1;
            const int v = col[mid];
            if (v == i) { pos = mid; break; }
            if (v < i) lo = mid + 1; else hi = mid - 1;
        }
        perm[k] = pos;   // -1 would mean the adjacency is not symmetric
    }
    und_cnt[i] = cnt;
}

__global__ void finalize_und_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, int n,
                                    const int* __restrict__ und_ptr, int* __restrict__ und_a, int* __restrict__ und_b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int w = und_ptr[i];
    for (int k = row_ptr[i]; k < row_ptr[i + 1]; ++k) {
        const int c = col[k];
        if (c > i) { und_a[w] = i; und_b[w] = c; ++w; }
    }
}

// ---------------------------------------------------------------------------------------------
// general path: target-major CSR from an arbitrary int64 edge_index
// ---------------------------------------------------------------------------------------------
__global__ void ei_count_kernel(const int64_t* __restrict__ es, const int64_t* __restrict__ ed, int n_edges,
                                int* __restrict__ cnt, int* __restrict__ und_flag) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    atomicAdd(cnt + (int)ed[e], 1);
    und_flag[e] = es[e] < ed[e] ? 1 : 0;
}

__global__ void ei_fill_kernel(const int64_t* __restrict__ ed, int n_edges, const int* __restrict__ row_ptr,
                               int* __restrict__ cursor, int* __restrict__ perm) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    const int t = (int)ed[e];
    perm[row_ptr[t] + atomicAdd(cursor + t, 1)] = e;
}

__global__ void ei_expand_kernel(const int64_t* __restrict__ es, const int64_t* __restrict__ ed, int n_edges,
                                 const int* __restrict__ perm, const int* __restrict__ und_flag,
                                 const int* __restrict__ und_pos, int* __restrict__ src, int* __restrict__ tgt,
                                 int* __restrict__ und_a, int* __restrict__ und_b) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_edges) return;
    const int e = perm[k];
    src[k] = (int)es[e];
    tgt[k] = (int)ed[e];
    if (und_flag[k]) {   // k used as an edge id here (same index range)
        und_a[und_pos[k]] = (int)es[k];
        und_b[und_pos[k]] = (int)ed[k];
    }
}

// ---------------------------------------------------------------------------------------------
// raw node / edge features
// ---------------------------------------------------------------------------------------------
__global__ void node_features_kernel(const float* __restrict__ px, const float* __restrict__ py,
                                     const float* __restrict__ vr, const float* __restrict__ rcs,
                                     const int64_t* __restrict__ ts, const int* __restrict__ degree,
                                     const int* __restrict__ frame_ptr, double min_range, double max_range,
                                     double min_az, double max_az, int range_f64, int az_f64,
                                     float* __restrict__ out /* (n,6) */) {
    __shared__ long long s_min[8], s_max[8];
    const int f = blockIdx.x;
    const int f0 = frame_ptr[f], f1 = frame_ptr[f + 1];
    long long lo = LLONG_MAX, hi = LLONG_MIN;
    for (int i = f0 + threadIdx.x; i < f1; i += blockDim.x) {
        const long long t = ts[i];
        lo = t < lo ? t : lo;
        hi = t > hi ? t : hi;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const long long l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = l2 < lo ? l2 : lo;
        hi = h2 > hi ? h2 : hi;
    }
    if ((threadIdx.x & 31) == 0) { s_min[threadIdx.x >> 5] = lo; s_max[threadIdx.x >> 5] = hi; }
    __syncthreads();
    lo = s_min[0]; hi = s_max[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
        lo = s_min[w] < lo ? s_min[w] : lo;
        hi = s_max[w] > hi ? s_max[w] : hi;
    }
    for (int i = f0 + threadIdx.x; i < f1; i += blockDim.x) {
        float* o = out + (size_t)i * 6;
        o[0] = vr[i];
        o[1] = rcs[i];
        // normalize_time (graph_features.py:47-55): int64 / int64 true division in float64
        o[2] = hi == lo ? 0.f : (float)((double)(ts[i] - lo) / (double)(hi - lo));
        o[3] = (float)((double)degree[i] / 10.0);                       // :130
        const float x = px[i], y = py[i];
        const float r = __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));   // :133
        const float th = fabsf((float)atan2((double)y, (double)x));     // :134 (float32 arctan2, correctly rounded here)
        o[4] = range_f64 ? (float)(((double)r - max_range) / (min_range - max_range))
                         : __fdiv_rn(__fsub_rn(r, (float)max_range), (float)(min_range - max_range));
        o[5] = az_f64 ? (float)(((double)th - max_az) / (min_az - max_az))
                      : __fdiv_rn(__fsub_rn(th, (float)max_az), (float)(min_az - max_az));
    }
}

__global__ void edge_features_kernel(const float* __restrict__ px, const float* __restrict__ py,
                                     const float* __restrict__ vx, const float* __restrict__ vy,
                                     const int64_t* __restrict__ ts, const int* __restrict__ er,
                                     const int* __restrict__ ec, int n_edges, float* __restrict__ out /* (E,7) */) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    const int s = er[e], t = ec[e];   // feature(source) - feature(target), source = edge_index[0] (graph_features.py:153)
    const float ex = __fdiv_rn(__fsub_rn(px[s], px[t]), 10.f);
    const float ey = __fdiv_rn(__fsub_rn(py[s], py[t]), 10.f);
    const float el = __fdiv_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey))), 10.f);
    const float evx = __fsub_rn(vx[s], vx[t]), evy = __fsub_rn(vy[s], vy[t]);
    const float ev = __fsqrt_rn(__fadd_rn(__fmul_rn(evx, evx), __fmul_rn(evy, evy)));
    const float et = (float)((double)(ts[s] - ts[t]) * 1e-6);
    float* o = out + (size_t)e * 7;
    o[0] = ex; o[1] = ey; o[2] = el; o[3] = evx; o[4] = evy; o[5] = ev; o[6] = et;
}

static int knn_stride(int knn) { return knn + 1; }

}  // namespace rgnn

using namespace rgnn;

// workspace: knn_idx (n*ks) | own (n) | extra (n) | total (n) | cursor (n) | scan ws | cell-sorted points (n) | frame grids | cell starts
extern "C" size_t rgnn_graph_build_workspace_bytes(int n_points, int n_frames, int knn) {
    const size_t n = (size_t)n_points;
    return align256(n * knn_stride(knn) * 4) + 4 * align256(n * 4) + align256(scan_ws_ints(n_points) * 4) + 256 +
           align256(n * sizeof(float4)) + align256((size_t)n_frames * sizeof(FrameGrid)) + align256((size_t)n_frames * (GRID_CELLS + 1) * 4);
}

extern "C" int rgnn_graph_build(const float* px, const float* py, const int32_t* frame_ptr_dev,
                                const int32_t* frame_ptr_host, int n_frames, int n_points, float eps2, int knn,
                                int union_radius, int32_t* degree, int32_t* row_ptr, int32_t* col,
                                int32_t edge_capacity, int32_t* n_edges_out, void* workspace, size_t workspace_bytes,
                                void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    RGNN_REQUIRE(n_frames >= 1 && n_points >= 1 && knn >= 1, "graph_build: empty input");
    RGNN_REQUIRE(knn <= 67, "graph_build: knn=%d above the supported maximum of 67", knn);
    RGNN_REQUIRE(workspace_bytes >= rgnn_graph_build_workspace_bytes(n_points, n_frames, knn), "graph_build: workspace too small");
    int max_nf = 0;
    for (int f = 0; f < n_frames; ++f) {
        const int nf = frame_ptr_host[f + 1] - frame_ptr_host[f];
        RGNN_REQUIRE(nf >= 1, "graph_build: frame %d is empty", f);
        max_nf = nf > max_nf ? nf : max_nf;
    }
    RGNN_REQUIRE(frame_ptr_host[n_frames] - frame_ptr_host[0] == n_points && n_frames <= 65535, "graph_build: bad frame_ptr");
    const int ks = knn_stride(knn);
    char* w = static_cast<char*>(workspace);
    int* knn_idx = reinterpret_cast<int*>(w); w += align256((size_t)n_points * ks * 4);
    int* own = reinterpret_cast<int*>(w); w += align256((size_t)n_points * 4);
    int* extra = reinterpret_cast<int*>(w); w += align256((size_t)n_points * 4);
    int* total = reinterpret_cast<int*>(w); w += align256((size_t)n_points * 4);
    int* cursor = reinterpret_cast<int*>(w); w += align256((size_t)n_points * 4);
    int* scan_ws = reinterpret_cast<int*>(w); w += align256(scan_ws_ints(n_points) * 4) + 256;
    float4* sorted = reinterpret_cast<float4*>(w); w += align256((size_t)n_points * sizeof(float4));
    FrameGrid* grids = reinterpret_cast<FrameGrid*>(w); w += align256((size_t)n_frames * sizeof(FrameGrid));
    int* cell_start = reinterpret_cast<int*>(w);

    const dim3 grid_a((max_nf + KNN_THREADS - 1) / KNN_THREADS, n_frames);
    if (g_knn_grid && max_nf > 4 * (knn + 1)) {
        // uniform grid per frame: the same result from ~10 x (k + 1) candidates per query instead of the whole frame
        cell_sort_kernel<<<n_frames, 256, 0, stream>>>(px, py, frame_ptr_dev, grids, cell_start, sorted);
        if (ks <= 12) knn_grid_kernel<12><<<grid_a, KNN_THREADS, 0, stream>>>(sorted, frame_ptr_dev, grids, cell_start, eps2, knn, ks, knn_idx, degree);
        else if (ks <= 20) knn_grid_kernel<20><<<grid_a, KNN_THREADS, 0, stream>>>(sorted, frame_ptr_dev, grids, cell_start, eps2, knn, ks, knn_idx, degree);
        else if (ks <= 36) knn_grid_kernel<36><<<grid_a, KNN_THREADS, 0, stream>>>(sorted, frame_ptr_dev, grids, cell_start, eps2, knn, ks, knn_idx, degree);
        else knn_grid_kernel<68><<<grid_a, KNN_THREADS, 0, stream>>>(sorted, frame_ptr_dev, grids, cell_start, eps2, knn, ks, knn_idx, degree);
    } else if (ks <= 12) knn_kernel<12><<<grid_a, KNN_THREADS, 0, stream>>>(px, py, frame_ptr_dev, eps2, knn, ks, knn_idx, degree);
    else if (ks <= 20) knn_kernel<20><<<grid_a, KNN_THREADS, 0, stream>>>(px, py, frame_ptr_dev, eps2, knn, ks, knn_idx, degree);
    else if (ks <= 36) knn_kernel<36><<<grid_a, KNN_THREADS, 0, stream>>>(px, py, frame_ptr_dev, eps2, knn, ks, knn_idx, degree);
    else knn_kernel<68><<<grid_a, KNN_THREADS, 0, stream>>>(px, py, frame_ptr_dev, eps2, knn, ks, knn_idx, degree);
    RGNN_CHECK_CUDA(cudaGetLastError());

    RGNN_CHECK_CUDA(cudaMemsetAsync(extra, 0, (size_t)n_points * 4, stream));
    RGNN_CHECK_CUDA(cudaMemsetAsync(cursor, 0, (size_t)n_points * 4, stream));
    const int nb = (n_points + 255) / 256;
    sym_count_kernel<<<nb, 256, 0, stream>>>(px, py, knn_idx, ks, degree, n_points, eps2, union_radius, own, extra);
    add_kernel<<<nb, 256, 0, stream>>>(own, extra, n_points, total);
    RGNN_CHECK_CUDA(cudaGetLastError());
    int rc = exclusive_scan(total, n_points, row_ptr, scan_ws, stream);
    if (rc) return rc;
    copy_last_kernel<<<1, 32, 0, stream>>>(row_ptr, n_points, n_edges_out);
    if (union_radius)
        radius_fill_kernel<<<grid_a, KNN_THREADS, 0, stream>>>(px, py, frame_ptr_dev, eps2, row_ptr, n_points, edge_capacity, col);
    sym_fill_kernel<<<nb, 256, 0, stream>>>(px, py, knn_idx, ks, degree, own, row_ptr, n_points, eps2, union_radius,
                                            edge_capacity, cursor, col);
    sort_rows_kernel<<<nb, 256, 0, stream>>>(row_ptr, n_points, edge_capacity, col, nullptr);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

extern "C" size_t rgnn_graph_finalize_workspace_bytes(int n_points, int n_edges) {
    (void)n_edges;
    return 2 * align256((size_t)(n_points + 1) * 4) + align256(scan_ws_ints(n_points) * 4) + 256;
}

extern "C" int rgnn_graph_finalize(const int32_t* row_ptr, const int32_t* col, int n_points, int n_edges,
                                   int32_t* row_of_edge, int32_t* perm, int32_t* und_a, int32_t* und_b,
                                   int32_t* n_und_out, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    (void)n_edges;
    RGNN_REQUIRE(workspace_bytes >= rgnn_graph_finalize_workspace_bytes(n_points, n_edges), "graph_finalize: workspace too small");
    char* w = static_cast<char*>(workspace);
    int* und_cnt = reinterpret_cast<int*>(w); w += align256((size_t)(n_points + 1) * 4);
    int* und_ptr = reinterpret_cast<int*>(w); w += align256((size_t)(n_points + 1) * 4);
    int* scan_ws = reinterpret_cast<int*>(w);
    const int nb = (n_points + 255) / 256;
    finalize_rows_kernel<<<nb, 256, 0, stream>>>(row_ptr, col, n_points, row_of_edge, perm, und_cnt);
    RGNN_CHECK_CUDA(cudaGetLastError());
    int rc = exclusive_scan(und_cnt, n_points, und_ptr, scan_ws, stream);
    if (rc) return rc;
    finalize_und_kernel<<<nb, 256, 0, stream>>>(row_ptr, col, n_points, und_ptr, und_a, und_b);
    copy_last_kernel<<<1, 32, 0, stream>>>(und_ptr, n_points, n_und_out);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

// workspace: cnt (n+1) | cursor (n) | und_flag (E) | und_pos (E+1) | scan ws
extern "C" size_t rgnn_csr_from_edge_index_workspace_bytes(int n_nodes, int n_edges) {
    const int m = n_nodes > n_edges ? n_nodes : n_edges;
    return 2 * align256((size_t)(n_nodes + 1) * 4) + 2 * align256((size_t)(n_edges + 1) * 4) + align256(scan_ws_ints(m) * 4) + 256;
}

extern "C" int rgnn_csr_from_edge_index(const int64_t* edge_src, const int64_t* edge_dst, int n_nodes, int n_edges,
                                        int32_t* row_ptr, int32_t* src, int32_t* tgt, int32_t* perm, int32_t* und_a,
                                        int32_t* und_b, int32_t* n_und_out, void* workspace, size_t workspace_bytes,
                                        void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    RGNN_REQUIRE(n_nodes >= 1 && n_edges >= 0, "csr_from_edge_index: empty graph");
    RGNN_REQUIRE(workspace_bytes >= rgnn_csr_from_edge_index_workspace_bytes(n_nodes, n_edges), "csr_from_edge_index: workspace too small");
    char* w = static_cast<char*>(workspace);
    int* cnt = reinterpret_cast<int*>(w); w += align256((size_t)(n_nodes + 1) * 4);
    int* cursor = reinterpret_cast<int*>(w); w += align256((size_t)(n_nodes + 1) * 4);
    int* und_flag = reinterpret_cast<int*>(w); w += align256((size_t)(n_edges + 1) * 4);
    int* und_pos = reinterpret_cast<int*>(w); w += align256((size_t)(n_edges + 1) * 4);
    int* scan_ws = reinterpret_cast<int*>(w);
    RGNN_CHECK_CUDA(cudaMemsetAsync(cnt, 0, (size_t)(n_nodes + 1) * 4, stream));
    RGNN_CHECK_CUDA(cudaMemsetAsync(cursor, 0, (size_t)(n_nodes + 1) * 4, stream));
    if (n_edges == 0) {
        RGNN_CHECK_CUDA(cudaMemsetAsync(row_ptr, 0, (size_t)(n_nodes + 1) * 4, stream));
        RGNN_CHECK_CUDA(cudaMemsetAsync(n_und_out, 0, 4, stream));
        return RGNN_OK;
    }
    const int nbe = (n_edges + 255) / 256, nbn = (n_nodes + 255) / 256;
    ei_count_kernel<<<nbe, 256, 0, stream>>>(edge_src, edge_dst, n_edges, cnt, und_flag);
    RGNN_CHECK_CUDA(cudaGetLastError());
    int rc = exclusive_scan(cnt, n_nodes, row_ptr, scan_ws, stream);
    if (rc) return rc;
    rc = exclusive_scan(und_flag, n_edges, und_pos, scan_ws, stream);
    if (rc) return rc;
    ei_fill_kernel<<<nbe, 256, 0, stream>>>(edge_dst, n_edges, row_ptr, cursor, perm);
    // order each row by caller edge id: deterministic, and source-ascending for reference-ordered input
    sort_rows_kernel<<<nbn, 256, 0, stream>>>(row_ptr, n_nodes, n_edges, perm, nullptr);
    ei_expand_kernel<<<nbe, 256, 0, stream>>>(edge_src, edge_dst, n_edges, perm, und_flag, und_pos, src, tgt, und_a, und_b);
    copy_last_kernel<<<1, 32, 0, stream>>>(und_pos, n_edges, n_und_out);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}

extern "C" int rgnn_graph_features(const float* px, const float* py, const float* vx, const float* vy, const float* vr,
                                   const float* rcs, const int64_t* timestamp_us, const int32_t* degree,
                                   const int32_t* frame_ptr_dev, int n_frames, int n_points, const int32_t* edge_row,
                                   const int32_t* edge_col, int n_edges, double min_range, double max_range,
                                   double min_azimuth, double max_azimuth, int range_in_f64, int azimuth_in_f64,
                                   float* node_features, float* edge_features, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    (void)n_points;
    if (node_features != nullptr) {
        node_features_kernel<<<n_frames, 256, 0, stream>>>(px, py, vr, rcs, timestamp_us, degree, frame_ptr_dev, min_range,
                                                           max_range, min_azimuth, max_azimuth, range_in_f64,
                                                           azimuth_in_f64, node_features);
    }
    if (edge_features != nullptr && n_edges > 0) {
        edge_features_kernel<<<(n_edges + 255) / 256, 256, 0, stream>>>(px, py, vx, vy, timestamp_us, edge_row, edge_col,
                                                                        n_edges, edge_features);
    }
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}
