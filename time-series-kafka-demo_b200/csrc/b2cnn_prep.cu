// b2cnn_prep.cu -- the two steps in FRONT of the model call, on the device (SURVEY.md section 8, f2 + f1):
//
//   f2  bin/processStream.py:196-208  per (patient, signal): mean of the samples of a 180 s window that
//                                     slides by 5 s (Spark `window(...)` + avg: nulls are skipped)
//       bin/processStream.py:62-123   forward-fill, back-fill, then 0-fill of that 5-second grid
//   f1  bin/predictStream.py:245-259  600 s windows sliding by 60 s -> 120 grid points per signal
//       bin/predictStream.py:105-139  x_arr[0, signal_index, :] = the 120 points, absent signals = zeros
//
// Input is one WFDB format-16 numerics record as it lies on disk (interleaved little-endian int16,
// -32768 == missing; physical = (adc - baseline) / gain, bin/sendStream.py:46 via wfdb.rdrecord), output
// is the [n_windows, n_channels, 120] batch the model consumes, written in f32 or bf16 straight into the
// tensor that b2cnn_forward reads -- the Kafka/Spark hop and the per-row numpy assembly disappear for replay.
// All sums are fp64 like Spark's avg; numpy's restatement (time-series-kafka-demo_b200/stream.py, the oracle of
// these kernels) takes prefix-sum differences instead of direct window sums, so the two agree to ~1e-12.
// This is byte shuffling around a few thousand samples per signal: launch-latency work, not a roofline kernel.
#include <cuda_bf16.h>

#include <cmath>

#include "b2cnn_internal.cuh"

namespace b2cnn {

constexpr int kPrepThreads = 1024;

struct PrepDims {
    int64_t n_samples, n_grid, n_windows;
    int n_sig, n_sel, n_channels, window_points, step;
    double period, grid_s, smooth_s;
};

struct PrepSignals {            // per selected signal: column in the record, gain, baseline
    int col[16];
    double gain[16], baseline[16];
};

__device__ __forceinline__ double phys_value(const int16_t *raw, int64_t i, int n_sig, int col, double gain, double base,
                                             bool *ok) {
    const int16_t a = raw[i * n_sig + col];
    *ok = a != (int16_t)-32768;
    return ((double)a - base) / gain;
}

// number of sample times t_i = i * period (i < n) that are <= v  (numpy.searchsorted(t, v, side="right"))
__device__ __forceinline__ int64_t count_le(int64_t n, double period, double v) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((double)mid * period <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// grid value k of signal ch: mean of the valid samples with time in (tau - smooth, tau], tau = k * grid_s.
// The window's samples are summed directly, in time order (what Spark's avg over the window rows does); a
// prefix-sum difference would leave cancellation residue (1e-12 instead of an exact 0 for an all-zero window).
__global__ void prep_grid_kernel(const int16_t *__restrict__ raw, PrepDims d, PrepSignals sg, double *__restrict__ grid) {
    const int ch = blockIdx.y;
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= d.n_grid) return;
    const double tau = (double)k * d.grid_s;
    const int64_t lo = count_le(d.n_samples, d.period, tau - d.smooth_s);
    const int64_t hi = count_le(d.n_samples, d.period, tau);
    const int col = sg.col[ch];
    const double gain = sg.gain[ch], base = sg.baseline[ch];
    double s = 0.0; int cnt = 0;
    for (int64_t i = lo; i < hi; ++i) {
        bool ok;
        const double v = phys_value(raw, i, d.n_sig, col, gain, base, &ok);
        if (ok) { s += v; ++cnt; }
    }
    grid[(int64_t)ch * d.n_grid + k] = cnt > 0 ? s / (double)cnt : nan("");
}

// one CTA per signal: forward fill (last valid value), back fill of the leading gap, zeros if nothing is valid
__global__ void __launch_bounds__(kPrepThreads)
prep_fill_kernel(PrepDims d, double *__restrict__ grid) {
    __shared__ long long sh_last[kPrepThreads];
    __shared__ long long sh_first;                                // first valid index of the signal (n_grid: none)
    const int ch = blockIdx.x, tid = threadIdx.x;
    double *g = grid + (int64_t)ch * d.n_grid;
    const int64_t per = (d.n_grid + kPrepThreads - 1) / kPrepThreads;
    const int64_t k0 = (int64_t)tid * per, k1 = min(k0 + per, d.n_grid);
    if (tid == 0) sh_first = (long long)d.n_grid;
    __syncthreads();
    long long last = -1, first_here = (long long)d.n_grid;
    for (int64_t k = k0; k < k1; ++k)
        if (g[k] == g[k]) { last = k; if (first_here > k) first_here = k; }
    sh_last[tid] = last;
    if (first_here < (long long)d.n_grid) atomicMin(&sh_first, first_here);
    __syncthreads();
    for (int off = 1; off < kPrepThreads; off <<= 1) {           // inclusive max-scan of the chunk results
        long long v = -1;
        if (tid >= off) v = sh_last[tid - off];
        __syncthreads();
        if (tid >= off && v > sh_last[tid]) sh_last[tid] = v;
        __syncthreads();
    }
    long long carry = tid > 0 ? sh_last[tid - 1] : -1;            // last valid index before this chunk
    const long long first = sh_first;
    const double first_val = first < (long long)d.n_grid ? g[first] : 0.0;
    // Values are read before any thread of this chunk overwrites them: a chunk only writes its own range and
    // reads g[carry] from an EARLIER chunk's last valid entry, which that chunk leaves unchanged (valid stays).
    const double carry_val = carry >= 0 ? g[carry] : 0.0;
    __syncthreads();
    double cur = carry_val; bool have = carry >= 0;
    for (int64_t k = k0; k < k1; ++k) {
        const double v = g[k];
        if (v == v) { cur = v; have = true; }
        else g[k] = have ? cur : first_val;                      // leading gap: back fill (or 0 when all missing)
    }
}

template <typename Tout>
__device__ __forceinline__ Tout cast_out(double v);
template <>
__device__ __forceinline__ float cast_out<float>(double v) { return (float)v; }                   // predictStream.py:155 .float()
template <>
__device__ __forceinline__ __nv_bfloat16 cast_out<__nv_bfloat16>(double v) { return __float2bfloat16_rn((float)v); }

// x[w][c][p] = grid[c][w * step + p] for the signals the record has, zeros for the rest (predictStream.py:131)
template <typename Tout>
__global__ void prep_assemble_kernel(PrepDims d, const double *__restrict__ grid, Tout *__restrict__ x, double *__restrict__ t0) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per_w = (int64_t)d.n_channels * d.window_points;
    if (e >= d.n_windows * per_w) return;
    const int64_t w = e / per_w;
    const int r = (int)(e - w * per_w);
    const int c = r / d.window_points, pnt = r - c * d.window_points;
    double v = 0.0;
    if (c < d.n_sel) v = grid[(int64_t)c * d.n_grid + w * d.step + pnt];
    x[e] = cast_out<Tout>(v);
    if (t0 && r == 0) t0[w] = (double)(w * d.step) * d.grid_s;
}

static bool prep_dims(int64_t n_samples, int n_sig, int n_sel, double fs, const b2cnn_prep_config *cfg, PrepDims *d, const char **err) {
    if (!cfg || n_samples < 1 || n_sig < 1 || n_sig > 64 || n_sel < 0 || n_sel > 16 || !(fs > 0.0)) { *err = "bad record shape"; return false; }
    if (cfg->n_channels < 1 || cfg->n_channels > 16 || n_sel > cfg->n_channels || cfg->window_points < 1 || cfg->grid_s < 1 ||
        cfg->smooth_s < 1 || cfg->stride_s < cfg->grid_s || cfg->stride_s % cfg->grid_s) { *err = "bad preprocessing configuration"; return false; }
    d->n_samples = n_samples; d->n_sig = n_sig; d->n_sel = n_sel;
    d->n_channels = cfg->n_channels; d->window_points = cfg->window_points; d->step = cfg->stride_s / cfg->grid_s;
    d->period = 1.0 / fs; d->grid_s = (double)cfg->grid_s; d->smooth_s = (double)cfg->smooth_s;
    const double t_last = (double)(n_samples - 1) * d->period;
    d->n_grid = (int64_t)floor(t_last / d->grid_s) + 1;
    const int64_t span = d->n_grid - d->window_points + 1;
    d->n_windows = span > 0 ? (span + d->step - 1) / d->step : 0;
    return true;
}

int64_t prep_window_count(int64_t n_samples, double fs, const b2cnn_prep_config *cfg) {
    PrepDims d; const char *e = "";
    if (!prep_dims(n_samples, 1, 0, fs, cfg, &d, &e)) return -1;
    return d.n_windows;
}

static int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }

int64_t prep_workspace_bytes(int64_t n_samples, double fs, int n_sel, const b2cnn_prep_config *cfg) {
    PrepDims d; const char *e = "";
    if (!prep_dims(n_samples, 1, n_sel, fs, cfg, &d, &e)) return -1;
    const int64_t ns = n_sel > 0 ? n_sel : 1;
    return align256(ns * d.n_grid * 8);                        // the 5-second grid of every selected signal, fp64
}

int prep_windows(const int16_t *raw, int64_t n_samples, int n_sig, const int *sel, int n_sel, const double *gains,
                 const double *baselines, double fs, const b2cnn_prep_config *cfg, void *x_out, int dtype, double *t0_out,
                 void *workspace, int64_t ws_bytes, cudaStream_t st, const char **err) {
    PrepDims d;
    if (!prep_dims(n_samples, n_sig, n_sel, fs, cfg, &d, err)) return B2CNN_EINVAL;
    if (!raw || !x_out || (n_sel > 0 && (!sel || !gains || !baselines))) { *err = "null pointer"; return B2CNN_EINVAL; }
    if (dtype != B2CNN_DTYPE_F32 && dtype != B2CNN_DTYPE_BF16) { *err = "dtype must be f32 or bf16"; return B2CNN_EINVAL; }
    if (ws_bytes < prep_workspace_bytes(n_samples, fs, n_sel, cfg) || !workspace) { *err = "workspace too small (b2cnn_prep_workspace_bytes)"; return B2CNN_ESTATE; }
    if (d.n_windows == 0) return B2CNN_OK;
    PrepSignals sg;
    for (int i = 0; i < n_sel; ++i) {
        if (sel[i] < 0 || sel[i] >= n_sig || !(gains[sel[i]] != 0.0)) { *err = "bad signal selection / zero gain"; return B2CNN_EINVAL; }
        sg.col[i] = sel[i]; sg.gain[i] = gains[sel[i]]; sg.baseline[i] = baselines[sel[i]];
    }
    double *grid = reinterpret_cast<double *>(workspace);
    if (n_sel > 0) {
        dim3 gg((unsigned)((d.n_grid + 255) / 256), n_sel);
        prep_grid_kernel<<<gg, 256, 0, st>>>(raw, d, sg, grid);
        prep_fill_kernel<<<n_sel, kPrepThreads, 0, st>>>(d, grid);
    }
    const int64_t total = d.n_windows * d.n_channels * d.window_points;
    const unsigned nb = (unsigned)((total + 255) / 256);
    if (dtype == B2CNN_DTYPE_F32)
        prep_assemble_kernel<float><<<nb, 256, 0, st>>>(d, grid, reinterpret_cast<float *>(x_out), t0_out);
    else
        prep_assemble_kernel<__nv_bfloat16><<<nb, 256, 0, st>>>(d, grid, reinterpret_cast<__nv_bfloat16 *>(x_out), t0_out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return B2CNN_ECUDA; }
    return B2CNN_OK;
}

}  // namespace b2cnn
