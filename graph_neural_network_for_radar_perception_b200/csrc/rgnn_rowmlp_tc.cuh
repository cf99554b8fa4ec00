// Row-MLP chains on the tensor cores: the program format shared by the host-side builders (rgnn_model.cu) and the
// interpreter kernel (rgnn_rowmlp_tc.cu).
//
// A chain (graph_feature_encoding, the stems, FFN_TaskSpecificHead, the update MLP of residual_graph_conv_block, ...;
// reference gnn_blocks.py) is evaluated for tiles of 128 rows.  A tile's activations never leave the SM: the input rows
// are written to TMEM as the A operand (hi | lo tf32 split), every Linear is a group of tcgen05.mma.kind::tf32 with the
// A operand in TMEM and the weights streamed L2 -> shared memory in 64 KB chunks (cp.async.bulk + mbarrier ring), the
// epilogue (bias, channel_normalization, LeakyReLU, residual, optional store) runs thread-per-row out of TMEM and writes
// the next layer's A operand back into TMEM.  A program is a list of STAGES = a few MMA chunk operations + one epilogue;
// the host assigns TMEM columns.  Wide no-norm layers (the 256-wide first encoder layer) are split into column blocks that
// the next layer consumes as K chunks, so no activation wider than 128 ever has to be resident.
#pragma once
#include "rgnn_common.cuh"

namespace rgnn {

constexpr int TC_MAX_STAGES = 12;
constexpr int TC_MAX_MMA = 3;
constexpr int TC_SLOT_FLOATS = 16384;     // 64 KB weight ring slot: hi + lo copies of a (K chunk x N) block, K*N <= 8192
constexpr int TC_SLOTS = 3;               // ring slots; the third one doubles as the input staging tile of programs that stage rows
constexpr int TC_XS_COL = 496;            // 16 spare TMEM columns for the per-row statistics exchange

enum TcInMode : int { TC_IN_ROWS = 0, TC_IN_PAIRSUM = 1, TC_IN_SEGMAX = 2, TC_IN_LIN0 = 3, TC_IN_BWD = 4 };

// Backward of (channel_normalization + LeakyReLU) of one layer applied to a gradient row: dz = norm'(act'(g)).  The
// forward pass of a training step saved the layer's output rows y (post-activation) and the per-row sigma, so nothing is
// recomputed; the normalised value is recovered from y like the FFMA backward does (rgnn_chain.cu, OP_ACTNORM_BWD).
struct TcBwd {
    const float* y;         // forward output rows of the layer (row stride y_ld); nullptr with lin0 = 0: plain Linear, dz = g
    const float* sd;        // per-row sigma of the norm (nullptr: no norm)
    const float* scale;     // channel_normalization scalars (nullptr: no norm)
    const float* shift;
    float* g_scale;         // gradients of the scalars (+=), nullable
    float* g_shift;
    float* y_store;         // lin0 only: the recomputed first-layer activation is written here (operand of a weight gradient)
    int y_ld, act;
    int lin0;               // 1: y = columns [lin0_off, lin0_off + n) of act(W0 f + b0), recomputed from the raw feature row
    int lin0_off;
};

struct TcMma {
    int a_hi, a_lo;     // TMEM columns of the A operand (K columns each)
    int d;              // TMEM column of the accumulator
    int N, K;           // MMA shape: N multiple of 16 (<= 128 here), K multiple of 8 (this chunk)
    int ldn, n_off;     // the chunk holds ldn output rows (K * ldn <= 8192); this MMA uses rows [n_off, n_off + N)
    int acc;            // 0 = overwrite D, 1 = accumulate
    const float* w;     // chunk in the packed weight stream: [hi: (K/4, ldn, 4)] [lo: same]
};

struct TcEpi {
    int d;              // TMEM column of the accumulator to read
    int n_cols;         // padded width handled (multiple of 16 * threads-per-row)
    int n_true;         // real width (bias / norm / store)
    int act;
    const float* bias;  // n_true floats or nullptr
    const float* scale; // channel_normalization scalars or nullptr
    const float* shift;
    int y_hi, y_lo;     // TMEM columns receiving the result as the next A operand (-1: nothing follows)
    int store_ld, store_w;
    float* store;       // optional: rows written to global memory (first store_w columns)
    const float* resid; // optional residual row added to the result (identity residual of the conv block)
    int resid_ld;
    int refill;         // 1: no epilogue; once the stage's MMAs are done the workers write the second LIN0 half into the A operand
    int is_bwd;         // 1: backward epilogue: z = norm'(act'(D)) described by `bwd` (no bias); store / y_hi / y_lo as usual
    float* sd_store;    // forward of a training step: per-row sigma of this layer's norm (nullable)
    float* pre_store;   // forward of a training step: the result BEFORE the residual is added (row stride n_true), nullable
    int store_mode;     // 0 = overwrite, 1 = accumulate (+=), 2 = atomic pair scatter onto rows ia[row], ib[row] of `store`
    int pad2;
    const int* ia;      // store_mode 2
    const int* ib;
    TcBwd bwd;
};

struct TcStage {
    int n_mma, pad;
    TcMma mma[TC_MAX_MMA];
    TcEpi epi;
};

struct TcInput {
    int mode;
    int k_pad;          // padded input width (multiple of 8)
    int a_hi, a_lo;     // TMEM columns of the input operand
    const float* p0;    // ROWS: first source, row stride ld0, w0 columns; PAIRSUM / SEGMAX: the gathered matrix
    const float* p1;    // ROWS: optional second source (concatenated), row stride ld1, w1 columns
    int ld0, w0, ld1, w1;
    const int* i0;      // ROWS: optional row index; PAIRSUM: first node; SEGMAX: segment pointer
    const int* i1;      // PAIRSUM: second node; SEGMAX: members
    // LIN0: the A operand is act(W0 f + b0), 128 columns at a time, computed by the workers on the CUDA cores from the raw
    // feature row f (p0, <= 7 columns, optional row index i0); W0 = p1 (256 x w0, row-major), b0 = lin_b
    const float* lin_b;
    int lin_act, pad;
    // BWD: p0 = gradient rows (ld0, w0 columns, optional row index i0); the A operand is norm'(act'(g)) described by `bwd`,
    // also written to bwd_store (row stride k_pad) when that is set
    TcBwd bwd;
    float* bwd_store;
    float* in_store;        // ROWS / PAIRSUM in a training step: the assembled input rows are written here (row stride k_pad)
};

// raw feature rows + first encoder Linear for epilogues with bwd.lin0 (the LIN0 input mode keeps its own copy in TcInput)
struct TcLin0 {
    const float* f;         // (rows, w) features, optional row index
    const int* ridx;
    const float* W;         // (256, w) row-major
    const float* b;
    int ld, w, act, pad;
};

struct TcProgram {
    int n_rows, n_stages;
    int n_slots, pad;       // 2 when the input rows are staged through shared memory (the staging tile aliases slot 2), else 3
    TcInput in;
    TcLin0 lin0;
    TcStage st[TC_MAX_STAGES];
};

// K-chunk size of a packed layer: the largest multiple of 8 dividing Kp with kc * Np <= 8192 floats per copy
inline int tc_chunk_k(int Kp, int Np) {
    int best = 8;
    for (int kc = 8; kc <= Kp; kc += 8)
        if (Kp % kc == 0 && kc * Np <= TC_SLOT_FLOATS / 2) best = kc;
    return best;
}
inline int tc_np(int out_features) { return round_up(out_features, 32); }
inline int tc_kp(int in_features) { return round_up(in_features, 8); }
inline size_t tc_pack_floats(int in_features, int out_features) { return (size_t)2 * tc_kp(in_features) * tc_np(out_features); }

int launch_rowmlp_tc(const TcProgram& pg, cudaStream_t stream);
// rows [n0, n0+Nt) x columns [k0, k0+Kt) of W -> chunked hi/lo pack at output rows [nd0, nd0+Nt) of a (Kp x Np) layer
int pack_tc(const float* W, int ldW, int n0, int Nt, int nd0, int Np, int k0, int Kt, int Kp, int kc, bool pad_rows, float* dst,
            cudaStream_t stream, bool transpose = false);

}  // namespace rgnn
