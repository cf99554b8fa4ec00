// Shared host/device declarations for the radar-GNN kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/rgnn.h"

namespace rgnn {

constexpr int NT = 256;          // threads per CTA for the tile kernels
constexpr int KC = 32;           // K-chunk of the staged weight operand
constexpr int CBMAX = 128;       // output-column block of the tile GEMM
constexpr float NORM_EPS = 1e-5f;     // reference modules/neural_net/constants.py:9
constexpr float LEAKY = 0.01f;        // reference modules/neural_net/constants.py:10

void set_error(const char* fmt, ...);
int sm_count();          // of the CURRENT device
int graph_set_option(const char* name, int value);     // rgnn_graph.cu ("knn_grid")
int graph_get_option(const char* name);

// Per-device one-time setup at a launch site (cudaFuncSetAttribute is a per-device setting): `needed()` is true until
// `mark()` has run on the current device.  The attribute calls are idempotent, so two threads racing through the same
// site is harmless; the bit mask is atomic.
struct PerDeviceOnce {
    unsigned long long done = 0;
    static int device() { int d = 0; cudaGetDevice(&d); return d & 63; }
    bool needed() const { return ((__atomic_load_n(&done, __ATOMIC_ACQUIRE) >> device()) & 1ull) == 0; }
    void mark() { __atomic_fetch_or(&done, 1ull << device(), __ATOMIC_RELEASE); }
};

#define RGNN_CHECK_CUDA(expr)                                                          \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess) {                                                       \
            rgnn::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return RGNN_ERR_CUDA;                                                      \
        }                                                                              \
    } while (0)

#define RGNN_REQUIRE(cond, ...)                                                        \
    do {                                                                               \
        if (!(cond)) {                                                                 \
            rgnn::set_error(__VA_ARGS__);                                              \
            return RGNN_ERR_INVALID;                                                   \
        }                                                                              \
    } while (0)

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

// ---------------------------------------------------------------------------------------------
// tile programs (interpreted by rgnn_chain.cu): a CTA owns a tile of TR rows; its activations live in
// numbered shared-memory regions; a step reads/writes explicit regions (ra, rb).
// ---------------------------------------------------------------------------------------------
enum Op : int {
    OP_END = 0,
    OP_LOAD_ROWS,      // ra[r][i2 + j] = src[row(r)*i0 + i4 + j], j < i1; zero-fill up to i3; p1 = optional int32 row index
    OP_LOAD_PAIRSUM,   // ra[r][j] = h[a[row]][j] + h[b[row]][j]
    OP_LOAD_SEGMAX,    // ra[r][j] = max_{m in members[ptr[row]..ptr[row+1])} g[m][j]
    OP_LINEAR,         // rb = ra * Wt (+ bias): i0 = K (mult of 8), i1 = C, i2 = Cpad (mult of 64), i3 = valid rows of Wt, i4 = ldw
    OP_ADD_GATHER2,    // ra[r][j] += P[t[row]][j] + P[s[row]][i2 + j], j < i1
    OP_NORM_ACT,       // ra: per-row channel norm (p0/p1 scale/shift, nullable) + LeakyReLU (i1); i2 = sigma slot or -1
    OP_ADD_ROWS,       // ra[r][j] += src[row*i0 + j], j < i1
    OP_STORE_ROWS,     // dst[row*i0 + i2 + j] (= or +=, i3) ra[r][i4 + j], j < i1
    OP_SEGSUM,         // dst[t[row]*i0 + i2 + j] += ra[r][i4 + j], j < i1, segmented over equal consecutive t
    OP_ADD_REGION,     // ra[r][i2 + j] += rb[r][i3 + j], j < i1
    OP_ACTNORM_BWD,    // ra (dY -> dZ in place) through LeakyReLU (i1) + channel norm of the layer whose output is rb
    OP_WGRAD,          // dW[c*i2 + i3 + k] += sum_r ra[r][i4 + c] * rb[r][k], c < i0, k < i1; db[c] += sum_r ra[r][i4 + c]
    OP_SCATTER_ADD,    // atomicAdd(dst[idx[row]*i0 + i2 + j], ra[r][i4 + j]), j < i1
    OP_PAIR_SCATTER,   // atomicAdd into dst[a[row]] and dst[b[row]]
    OP_SEGMAX_BWD,     // route d(pooled) rows in ra to the arg-max member rows of dst
};

struct Step {
    int op;
    short ra, rb;
    int i0, i1, i2, i3, i4, i5;
    const void* p0;
    const void* p1;
    const void* p2;
    const void* p3;
};

constexpr int MAX_STEPS = 44;
constexpr int MAX_REGIONS = 16;
constexpr int MAX_SIGMA = 8;

struct Program {
    int n_steps;
    int n_rows;
    int tr;                 // rows per tile (64 forward, 32 backward)
    int region_floats;      // total shared-memory floats of all regions
    int wacc_floats;        // shared-memory weight-gradient accumulators (assigned by launch_program)
    int reg_off[MAX_REGIONS];
    int reg_ld[MAX_REGIONS];
    Step steps[MAX_STEPS];
};

}  // namespace rgnn
