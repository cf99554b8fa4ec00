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
int sm_count();

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
// tile-program steps (see rgnn_chain.cu)
// ---------------------------------------------------------------------------------------------
enum Op : int {
    OP_END = 0,
    OP_LOAD_ROWS,      // cur[r][i2 + j] = src[row(r)*i0 + j], j < i1; zero up to i3; p1 = optional int32 row index
    OP_LOAD_PAIRSUM,   // cur[r][j] = h[a[row]][j] + h[b[row]][j]
    OP_LOAD_SEGMAX,    // cur[r][j] = max_{m in members[ptr[row]..ptr[row+1])} g[m][j]
    OP_LINEAR,         // nxt = cur * Wt (+ bias); swap
    OP_ADD_GATHER2,    // cur[r][j] += P[t[row]][j] + P[s[row]][i2 + j]
    OP_NORM_ACT,       // per-row channel norm (optional) + LeakyReLU (optional), in place
    OP_ADD_ROWS,       // cur[r][j] += src[row*i0 + j]
    OP_STORE_ROWS,     // dst[row*i0 + i2 + j] = cur[r][j], j < i1
    OP_SEGSUM,         // agg[t[row]][j] += cur[r][j]  (segmented by equal consecutive t)
    // ---- backward-only steps (rgnn_chain_bwd.cu) ----
    OPB_SAVE_INPUT,    // remember the current buffer as the input of linear #i0
    OPB_LOAD_GRAD,     // work[r][j] = g[row(r)*i0 + j] (p1 optional row index), j < i1
    OPB_ACTNORM_BWD,   // work (dY) -> dZ in place through LeakyReLU + channel norm of layer i0
    OPB_WGRAD,         // dW[c][k] += sum_r dZ[r][c] X[r][k];  db[c] += sum_r dZ[r][c]
    OPB_DGRAD,         // nxt_work = dZ * W ; swap
    OPB_STORE_GRAD,    // dst[row*i0 + i2 + j] (=|+=) work[r][j]
    OPB_SCATTER_GRAD,  // atomicAdd(dst[idx[row]*i0 + i2 + j], work[r][j])
    OPB_SEGSUM_GRAD,   // like OP_SEGSUM but from the work buffer
    OPB_SEGMAX_BWD,    // route d(pooled) to the arg-max member rows
    OPB_PAIR_SCATTER,  // atomicAdd dst[a[row]] and dst[b[row]]
};

struct Step {
    int op;
    int i0, i1, i2, i3;
    const void* p0;
    const void* p1;
    const void* p2;
    const void* p3;
};

constexpr int MAX_STEPS = 40;

struct Program {
    int n_steps;
    int n_rows;
    Step steps[MAX_STEPS];
};

}  // namespace rgnn
