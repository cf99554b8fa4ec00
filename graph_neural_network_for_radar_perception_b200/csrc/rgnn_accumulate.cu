// Sliding-window accumulation pre-pass: the step that turns the raw detections of the
// scans of a window into the `meas_*` arrays the graph construction reads -- on the device, for many windows per launch.
//
//   reference modules/data_utils/read_data.py:227-303     extract_and_sync_radar_data (per-scan loop)
//             modules/data_utils/meas_selection.py:39-70  stationary gate |vr_pred - vr| <= 1.5
//             modules/data_utils/meas_sync.py:15-21,44-70 vr_cartesian_vf, ego compensation R p + t
//             modules/compute_groundtruth/compute_node_labels.py:71-86   class label per detection
//             modules/compute_features/grid_features.py:162-173          region-of-interest filter
//             modules/compute_features/graph_features.py:167-182         dynamic-detection filter
//
// The per-scan scalars (relative pose T = inv(T_curr) T_prev, sensor velocity, mount yaw) are computed by the host
// mirror in float64 exactly as the reference computes them (a 3x3 inverse per window is not device work) and arrive
// as a table of 9 doubles per scan; everything per detection happens here.  HBM-bound streaming work: 39 B in and
// <= 45 B out per detection, two passes (count, write) over the inputs of a window; one CTA per window keeps the
// compaction stable (the reference's boolean-mask order) without atomics.
//
// Arithmetic contract (tests/test_accumulate_gpu.py): positions bit-exact (float64 fma(R01, y, R00 * x) + t, the order
// of the OpenBLAS dgemm micro-kernel behind the reference's 2x2 @ 2xN matmul, then one rounding to float32); stationary
// flags, class labels, selection and order exact; velocities within 2 ulp (float32 cos / sin: the reference uses NumPy's
// SIMD float32 routines, this kernel the correctly rounded value of the float64 function).
#include "rgnn_common.cuh"

namespace rgnn {

constexpr int ACC_NT = 256;
constexpr int SCAN_PARAMS = 9;      // R00 R01 R10 R11 tx ty mount_yaw vxs vys
__constant__ int c_old_to_new[12] = {0, 4, 4, 4, 4, 3, 3, 1, 2, 5, 5, 7};   // reference modules/data_utils/labels.py:18-31,90-100

struct AccArgs {
    const float* x_cc; const float* y_cc; const float* vr; const float* vr_comp; const float* azimuth; const float* rcs;
    const long long* timestamp; const unsigned char* label_id; const unsigned char* has_track; const int* point_scan;
    const double* scan_params; const int* window_ptr; const unsigned char* flip;
    int n_windows; int select;
    float min_x, max_x, min_y, max_y;
    float* px; float* py; float* vx; float* vy; float* vr_out; float* rcs_out; long long* ts_out; float* label_out;
    int* src_index; unsigned char* stationary_out; int* out_ptr; int* counts;
};

struct Det {
    float px, py, vx, vy, label;
    bool stationary, keep;
};

__device__ __forceinline__ Det transform(const AccArgs& a, int i, bool flip) {
    Det d;
    const double* sp = a.scan_params + (size_t)a.point_scan[i] * SCAN_PARAMS;
    const float az = a.azimuth[i];
    // stationary gate: float32 cos / sin of the azimuth, float64 products (the sensor velocity is a float64 scalar)
    const float ca = (float)cos((double)az), sa = (float)sin((double)az);
    const double vr_pred = -__dadd_rn(__dmul_rn(sp[7], (double)ca), __dmul_rn(sp[8], (double)sa));
    d.stationary = fabs(vr_pred - (double)a.vr[i]) <= 1.5;
    // compensated range rate as a vehicle-frame vector: float32 throughout (the mount yaw is a weak Python scalar)
    const float angle = __fadd_rn(az, (float)sp[6]);
    const float vrc = a.vr_comp[i];
    d.vx = __fmul_rn(vrc, (float)cos((double)angle));
    d.vy = __fmul_rn(vrc, (float)sin((double)angle));
    // ego compensation of the position into the frame of the window's last scan
    const double x = (double)a.x_cc[i], y = (double)a.y_cc[i];
    d.px = (float)__dadd_rn(__fma_rn(sp[1], y, __dmul_rn(sp[0], x)), sp[4]);
    d.py = (float)__dadd_rn(__fma_rn(sp[3], y, __dmul_rn(sp[2], x)), sp[5]);
    if (flip) { d.py = -d.py; d.vy = -d.vy; }
    const int lab = a.has_track[i] ? c_old_to_new[min((int)a.label_id[i], 11)] : (d.stationary ? 7 : 6);
    d.label = (float)lab;
    d.keep = !a.select || (d.px >= a.min_x && d.px < a.max_x && d.py >= a.min_y && d.py < a.max_y && lab != 7);
    return d;
}

// pass 1: kept detections per window
__global__ void __launch_bounds__(ACC_NT) acc_count_kernel(const AccArgs a) {
    const int w = blockIdx.x;
    const int p0 = a.window_ptr[w], p1 = a.window_ptr[w + 1];
    const bool flip = a.flip != nullptr && a.flip[w] != 0;
    int n = 0;
    for (int i = p0 + threadIdx.x; i < p1; i += ACC_NT) n += transform(a, i, flip).keep ? 1 : 0;
    __shared__ int part[ACC_NT / 32];
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < ACC_NT / 32; ++k) t += part[k];
        a.counts[w] = t;
    }
}

// exclusive scan of the window counts (a few hundred windows: one CTA)
__global__ void __launch_bounds__(1024) acc_scan_kernel(const int* __restrict__ counts, int n, int* __restrict__ out_ptr) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n ? counts[i] : 0;
        int x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int t = warp_tot[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, t, o);
                if (threadIdx.x >= o) t += y;
            }
            warp_tot[threadIdx.x] = t;
        }
        __syncthreads();
        const int before = carry + (threadIdx.x >= 32 ? warp_tot[(threadIdx.x >> 5) - 1] : 0) + x - v;
        if (i < n) out_ptr[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) out_ptr[n] = carry;
}

// pass 2: recompute, compact in order, write
__global__ void __launch_bounds__(ACC_NT) acc_write_kernel(const AccArgs a) {
    const int w = blockIdx.x;
    const int p0 = a.window_ptr[w], p1 = a.window_ptr[w + 1];
    const bool flip = a.flip != nullptr && a.flip[w] != 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ int part[ACC_NT / 32];
    int base = a.out_ptr[w];
    for (int c = p0; c < p1; c += ACC_NT) {
        const int i = c + threadIdx.x;
        Det d;
        d.keep = false;
        if (i < p1) {
            d = transform(a, i, flip);
            if (a.stationary_out != nullptr) a.stationary_out[i] = d.stationary ? 1 : 0;
        }
        const unsigned m = __ballot_sync(0xffffffffu, d.keep);
        if (lane == 0) part[warp] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int k = 0; k < ACC_NT / 32; ++k) {
            const int v = part[k];
            before += k < warp ? v : 0;
            total += v;
        }
        if (d.keep) {
            const int o = base + before + __popc(m & ((1u << lane) - 1u));
            a.px[o] = d.px; a.py[o] = d.py; a.vx[o] = d.vx; a.vy[o] = d.vy;
            a.vr_out[o] = a.vr_comp[i]; a.rcs_out[o] = a.rcs[i]; a.ts_out[o] = a.timestamp[i];
            a.label_out[o] = d.label; a.src_index[o] = i - p0;
        }
        base += total;
        __syncthreads();
    }
}

}  // namespace rgnn

extern "C" size_t rgnn_accumulate_workspace_bytes(int n_windows) {
    return rgnn::align256(sizeof(int) * (size_t)(n_windows > 0 ? n_windows : 1));
}

extern "C" int rgnn_accumulate_windows(const float* x_cc, const float* y_cc, const float* vr, const float* vr_compensated,
                                       const float* azimuth_sc, const float* rcs, const int64_t* timestamp,
                                       const uint8_t* label_id, const uint8_t* has_track, const int32_t* point_scan,
                                       const double* scan_params, const int32_t* window_ptr_dev, const uint8_t* flip_dev,
                                       int n_windows, int n_points, int select, float min_x, float max_x, float min_y,
                                       float max_y, float* meas_px, float* meas_py, float* meas_vx, float* meas_vy,
                                       float* meas_vr, float* meas_rcs, int64_t* meas_timestamp, float* class_labels,
                                       int32_t* src_index, uint8_t* stationary_flag, int32_t* out_ptr_dev, void* workspace,
                                       size_t workspace_bytes, void* stream) {
    using namespace rgnn;
    RGNN_REQUIRE(n_windows >= 0 && n_points >= 0, "rgnn_accumulate_windows: negative sizes");
    RGNN_REQUIRE(out_ptr_dev != nullptr, "rgnn_accumulate_windows: out_ptr_dev is required");
    cudaStream_t s = (cudaStream_t)stream;
    if (n_windows == 0) {
        RGNN_CHECK_CUDA(cudaMemsetAsync(out_ptr_dev, 0, sizeof(int), s));
        return RGNN_OK;
    }
    RGNN_REQUIRE(workspace != nullptr && workspace_bytes >= rgnn_accumulate_workspace_bytes(n_windows),
                 "rgnn_accumulate_windows: workspace too small");
    RGNN_REQUIRE(n_points == 0 || (x_cc && y_cc && vr && vr_compensated && azimuth_sc && rcs && timestamp && label_id && has_track &&
                                   point_scan && scan_params && meas_px && meas_py && meas_vx && meas_vy && meas_vr && meas_rcs &&
                                   meas_timestamp && class_labels && src_index),
                 "rgnn_accumulate_windows: null array");
    RGNN_REQUIRE(window_ptr_dev != nullptr, "rgnn_accumulate_windows: window_ptr_dev is required");
    AccArgs a;
    a.x_cc = x_cc; a.y_cc = y_cc; a.vr = vr; a.vr_comp = vr_compensated; a.azimuth = azimuth_sc; a.rcs = rcs;
    a.timestamp = (const long long*)timestamp; a.label_id = label_id; a.has_track = has_track; a.point_scan = point_scan;
    a.scan_params = scan_params; a.window_ptr = window_ptr_dev; a.flip = flip_dev;
    a.n_windows = n_windows; a.select = select;
    a.min_x = min_x; a.max_x = max_x; a.min_y = min_y; a.max_y = max_y;
    a.px = meas_px; a.py = meas_py; a.vx = meas_vx; a.vy = meas_vy; a.vr_out = meas_vr; a.rcs_out = meas_rcs;
    a.ts_out = (long long*)meas_timestamp; a.label_out = class_labels; a.src_index = src_index;
    a.stationary_out = stationary_flag; a.out_ptr = out_ptr_dev; a.counts = (int*)workspace;
    acc_count_kernel<<<n_windows, ACC_NT, 0, s>>>(a);
    RGNN_CHECK_CUDA(cudaGetLastError());
    acc_scan_kernel<<<1, 1024, 0, s>>>(a.counts, n_windows, out_ptr_dev);
    RGNN_CHECK_CUDA(cudaGetLastError());
    acc_write_kernel<<<n_windows, ACC_NT, 0, s>>>(a);
    RGNN_CHECK_CUDA(cudaGetLastError());
    return RGNN_OK;
}
