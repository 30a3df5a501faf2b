// LearningModule.estimateDisturbance / learn preprocessing on the device (Learning_module.py:46-59, :63-120;
// SURVEY §8f rank 2): box-filtered positions -> np.gradient on the recorded time stamps -> box-filtered velocities
// -> drift means, a0 = median(speed / freq), GP training inputs / targets.  A few thousand samples: one CTA walks
// the phases with barriers, intermediate arrays live in a caller-provided workspace; what it buys is that the
// trajectory recorded by the fused rollout never leaves HBM on its way into mr_gp_fit.
#include <cuda_runtime.h>

#include "mr_common.cuh"

namespace mr {

constexpr int kLearnThreads = 1024;

// scipy.ndimage.uniform_filter1d(x, size, mode="nearest"), origin 0: mean of x[i - size/2 .. i - size/2 + size - 1]
// with indices clamped to the array (scipy keeps a running sum; summing the window directly differs by rounding only)
__device__ void box_filter_nearest(const double* x, double* y, int n, int size) {
    const int left = size / 2;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < size; ++k) {
            int j = i - left + k;
            j = j < 0 ? 0 : (j >= n ? n - 1 : j);
            s += x[j];
        }
        y[i] = s / (double)size;
    }
    __syncthreads();
}

// np.gradient(f, t) (edge_order 1): second-order interior differences on a non-uniform grid, one-sided at the ends
__device__ void gradient_nonuniform(const double* f, const double* t, double t0, double* g, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double v;
        if (n < 2) v = 0.0;
        else if (i == 0) v = (f[1] - f[0]) / ((t[1] - t0) - (t[0] - t0));
        else if (i == n - 1) v = (f[n - 1] - f[n - 2]) / ((t[n - 1] - t0) - (t[n - 2] - t0));
        else {
            const double hs = (t[i] - t0) - (t[i - 1] - t0), hd = (t[i + 1] - t0) - (t[i] - t0);
            const double a = -hd / (hs * (hd + hs)), b = (hd - hs) / (hd * hs), c = hs / (hd * (hd + hs));
            v = a * f[i - 1] + b * f[i] + c * f[i + 1];
        }
        g[i] = v;
    }
    __syncthreads();
}

__device__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0) { for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w]; red[32] = s; }
    __syncthreads();
    s = red[32];
    __syncthreads();
    return s;
}

__global__ void __launch_bounds__(kLearnThreads, 1)
learn_preprocess_kernel(const double* __restrict__ px, const double* __restrict__ py, const double* __restrict__ time, int n,
                        int filter_n, int subtract_t0, double drift_x, double drift_y, const double* __restrict__ alpha_sim,
                        double freq, int n_valid, double* vx, double* vy, double* __restrict__ x_out,
                        double* __restrict__ yx_out, double* __restrict__ yy_out, double* __restrict__ scalars, double* ws) {
    __shared__ double red[33];
    __shared__ double s_med[2];
    double* fx = ws;             // filtered positions, later speed / freq
    double* fy = ws + n;
    double* gx = ws + 2 * (int64_t)n;
    double* gy = ws + 3 * (int64_t)n;
    const double t0 = subtract_t0 ? time[0] : 0.0;          // learn() shifts the clock first (:70); only differences matter
    const int N = filter_n;

    box_filter_nearest(px, fx, n, N);
    box_filter_nearest(py, fy, n, N);
    gradient_nonuniform(fx, time, t0, gx, n);
    gradient_nonuniform(fy, time, t0, gy, n);
    box_filter_nearest(gx, vx, n, N / 2);
    box_filter_nearest(gy, vy, n, N / 2);

    // drift estimate: mean filtered velocity away from the filter's boundary effect (:58-59)
    const int lo = N, hi = n - N;
    double sx = 0.0, sy = 0.0;
    for (int i = lo + (int)threadIdx.x; i < hi; i += blockDim.x) { sx += vx[i]; sy += vy[i]; }
    sx = block_sum(sx, red);
    sy = block_sum(sy, red);
    const int cnt = hi > lo ? hi - lo : 0;
    if (threadIdx.x == 0) { scalars[0] = cnt ? sx / cnt : 0.0; scalars[1] = cnt ? sy / cnt : 0.0; scalars[2] = 0.0; scalars[3] = 0.0; }
    if (!alpha_sim) return;

    // learn(): frames [N, n_valid - N) of the controller-on part; a0 = median(speed / freq) (:86-112)
    const int m = n_valid - 2 * N;
    if (m <= 0) return;
    double* sp = fx;                                           // reuse: speed / freq
    for (int k = threadIdx.x; k < m; k += blockDim.x) {
        const double dx = vx[N + k] - drift_x, dy = vy[N + k] - drift_y;
        sp[k] = sqrt(dx * dx + dy * dy) / freq;
    }
    __syncthreads();
    // median by rank counting (ties broken by index): ranks (m-1)/2 and m/2, averaged like np.median
    const int r0 = (m - 1) / 2, r1 = m / 2;
    for (int k = threadIdx.x; k < m; k += blockDim.x) {
        const double v = sp[k];
        int rank = 0;
        for (int j = 0; j < m; ++j) { const double w = sp[j]; rank += (w < v) || (w == v && j < k); }
        if (rank == r0) s_med[0] = v;
        if (rank == r1) s_med[1] = v;
    }
    __syncthreads();
    const double a0 = r0 == r1 ? s_med[0] : 0.5 * (s_med[0] + s_med[1]);
    for (int k = threadIdx.x; k < m; k += blockDim.x) {
        const double al = alpha_sim[N + k];
        x_out[k] = al;
        yx_out[k] = vx[N + k] - a0 * freq * cos(al);           // residual targets of the two GPs (:117-118)
        yy_out[k] = vy[N + k] - a0 * freq * sin(al);
    }
    if (threadIdx.x == 0) { scalars[2] = a0; scalars[3] = (double)m; }
}

}  // namespace mr

extern "C" {

int64_t mr_learn_workspace_bytes(int32_t n) { return n > 0 ? (int64_t)n * 4 * 8 : 0; }

int mr_learn_preprocess(const double* px, const double* py, const double* time, int32_t n, int32_t filter_n, int32_t subtract_t0,
                        double drift_x, double drift_y, const double* alpha_sim, double freq, int32_t n_valid, double* vx_out,
                        double* vy_out, double* x_out, double* yx_out, double* yy_out, double* scalars_out, void* workspace,
                        int64_t workspace_bytes, void* stream) {
    mr::NvtxRange nvtx_range("mr_learn_preprocess");
    using namespace mr;
    if (!px || !py || !time || !vx_out || !vy_out || !scalars_out) return fail(MR_ERR_ARG, "mr_learn_preprocess: null argument");
    if (n < 2 || n > (1 << 20)) return fail(MR_ERR_ARG, "mr_learn_preprocess: need 2 <= n <= 2^20 samples");
    if (filter_n < 2 || 2 * filter_n >= n) return fail(MR_ERR_ARG, "mr_learn_preprocess: filter length must satisfy 2 <= N < n/2");
    if (!workspace || workspace_bytes < mr_learn_workspace_bytes(n)) return fail(MR_ERR_ARG, "mr_learn_preprocess: workspace too small");
    if (alpha_sim) {
        if (!x_out || !yx_out || !yy_out) return fail(MR_ERR_ARG, "mr_learn_preprocess: null output for learn()");
        if (!(freq != 0.0)) return fail(MR_ERR_ARG, "mr_learn_preprocess: freq must be non-zero");
        if (n_valid > n || n_valid <= 2 * filter_n) return fail(MR_ERR_ARG, "mr_learn_preprocess: need 2 N < n_valid <= n");
        if (n_valid - 2 * filter_n > 65536) return fail(MR_ERR_UNSUPPORTED, "mr_learn_preprocess: the rank-counting median handles up to 65536 frames");
    }
    learn_preprocess_kernel<<<1, kLearnThreads, 0, (cudaStream_t)stream>>>(px, py, time, n, filter_n, subtract_t0, drift_x, drift_y,
                                                                           alpha_sim, freq, n_valid, vx_out, vy_out, x_out, yx_out,
                                                                           yy_out, scalars_out, (double*)workspace);
    return check_launch("mr_learn_preprocess");
}

}  // extern "C"
