// b2cnn_api.cu -- the extern "C" boundary declared in include/b2cnn.h.
// Replaces the reference's `model = torch.load(...); model.eval()` (bin/predictStream.py:36-37)
// and `output = model(x, age)` (bin/predictStream.py:157) with plain-pointer entry points.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "b2cnn_internal.cuh"
#include "b2cnn_tc.cuh"

using namespace b2cnn;

static thread_local std::string g_err;

static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
static int cuda_fail(cudaError_t e, const char *what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return B2CNN_ECUDA;
}
#define CU_TRY(expr)                                         \
    do {                                                     \
        cudaError_t e__ = (expr);                            \
        if (e__ != cudaSuccess) return cuda_fail(e__, #expr); \
    } while (0)

// Every entry point that touches the handle's device switches to it for the duration of the call only and
// restores the caller's current device on every exit path (a forward on cuda:1 must not leave the calling
// thread on cuda:1 -- later `device="cuda"` allocations of the host framework would land on the wrong GPU).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        err = cudaGetDevice(&prev);
        if (err != cudaSuccess) { prev = -1; return; }
        if (prev != dev) err = cudaSetDevice(dev); else prev = -1;     // nothing to restore
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};
#define DEVICE_GUARD(dev)                   \
    DeviceGuard guard__(dev);               \
    if (guard__.err != cudaSuccess) return cuda_fail(guard__.err, "cudaSetDevice")

struct b2cnn_handle {
    b2cnn_config cfg;
    Dims d;
    int device = 0;
    int num_sms = 148;
    bool weights_set = false;
    ConvWeights cw;
    HeadWeights hw;
    float *d_blob = nullptr;    // packed blob as given
    float *d_wih0T = nullptr;   // [L][64]
    int64_t n_weights = 0;
    int64_t opt_path = B2CNN_PATH_AUTO;
    int64_t opt_stream = 1;      // fp32 windows: streaming kernel instead of the generic one
    int64_t opt_tc_splits = 3;   // bf16 pieces per conv1 weight in the fused kernels
    int64_t last_launches = 0;
    int last_path = 0;
    int64_t opt_profile = 0;
    int64_t opt_small = 1;      // single-launch kernel for short windows / small batches
    cudaEvent_t ev_stage[3] = {nullptr, nullptr, nullptr};   // start, after front end, after head
    bool ev_valid = false;
    TcState tc;                 // tensor-core path state (b2cnn_tc.cu)
    // ---- host-path staging (b2cnn_forward_host only)
    void *st_x[2] = {nullptr, nullptr};
    size_t st_x_bytes = 0;
    float *st_age = nullptr, *st_out = nullptr;
    void *st_ws = nullptr;
    size_t st_ws_bytes = 0, st_vec_elems = 0;
    cudaStream_t s_copy = nullptr, s_comp = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
};

static int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

static bool derive_dims(const b2cnn_config &c, Dims &d) {
    d.C = c.in_channels; d.K1 = c.k1; d.K2 = c.k2; d.PK = c.pool_k; d.PS = c.pool_s; d.W = c.window;
    d.act = c.act; d.has_affine = (c.flags & B2CNN_FLAG_AFFINE) ? 1 : 0; d.age_coef = c.age_coef; d.XP = c.window;
    if (d.C < 1 || d.K1 < 1 || d.K2 < 1 || d.PK < 1 || d.PS < 1 || d.W < 1) return false;
    d.L1 = d.W - d.K1 + 1;
    if (d.L1 < d.PK) return false;
    d.P1 = (d.L1 - d.PK) / d.PS + 1;
    d.L2 = d.P1 - d.K2 + 1;
    if (d.L2 < d.PK) return false;
    d.L = (d.L2 - d.PK) / d.PS + 1;
    return d.L >= 1;
}

extern "C" int64_t b2cnn_l_out(const b2cnn_config *cfg) {
    Dims d;
    if (!cfg || !derive_dims(*cfg, d)) return -1;
    return d.L;
}

extern "C" int64_t b2cnn_weight_count(const b2cnn_config *cfg) {
    Dims d;
    if (!cfg || !derive_dims(*cfg, d)) return -1;
    int64_t n = (int64_t)kCMid * d.C * d.K1 + kCMid + kCMid * d.K2 + 1;
    n += (int64_t)kGates * d.L + kGates * kHidden + 2 * kGates;   // layer 0
    n += 2 * kGates * kHidden + 2 * kGates;                       // layer 1
    n += kHidden + 1;                                             // out
    if (d.has_affine) n += 2 * kCMid + 2;
    return n;
}

extern "C" const char *b2cnn_last_error(void) { return g_err.c_str(); }

// ---- one training step (b2cnn_train.cu) ----
extern "C" int64_t b2cnn_train_workspace_bytes(const b2cnn_config *cfg, int64_t B) {
    const int64_t n = train_workspace_bytes(cfg, B);
    if (n < 0) fail(B2CNN_EINVAL, "b2cnn_train_workspace_bytes: bad configuration / batch");
    return n;
}
extern "C" int b2cnn_train_step(const b2cnn_config *cfg, float *params, float *adam_m, float *adam_v, float *grads, int64_t step,
                                const b2cnn_adam *opt, int apply_update, const float *x, int64_t B, const float *age,
                                const float *target, int mode, const float *mask1, const float *mask2, float *loss_out,
                                void *workspace, int64_t workspace_bytes, void *stream) {
    if (!cfg || !opt) return fail(B2CNN_EINVAL, "b2cnn_train_step: null configuration");
    if (mode != B2CNN_MODE_INDEPENDENT && mode != B2CNN_MODE_SEQUENCE) return fail(B2CNN_EINVAL, "b2cnn_train_step: bad mode");
    int prev = -1;
    if (cfg->device >= 0) {
        if (cudaGetDevice(&prev) != cudaSuccess || cudaSetDevice(cfg->device) != cudaSuccess) return fail(B2CNN_ECUDA, "b2cnn_train_step: cudaSetDevice");
    }
    const char *err = "";
    const int rc = train_step(cfg, params, adam_m, adam_v, grads, step, opt->lr, opt->beta1, opt->beta2, opt->eps, apply_update, x, B, age,
                              target, mode == B2CNN_MODE_SEQUENCE ? 1 : 0, mask1, mask2, loss_out, workspace, workspace_bytes,
                              reinterpret_cast<cudaStream_t>(stream), &err);
    if (prev >= 0) cudaSetDevice(prev);
    return rc == B2CNN_OK ? rc : fail(rc, std::string("b2cnn_train_step: ") + err);
}

// ---- preprocessing + window assembly (b2cnn_prep.cu) ----
extern "C" int64_t b2cnn_prep_window_count(int64_t n_samples, double fs, const b2cnn_prep_config *cfg) {
    const int64_t n = prep_window_count(n_samples, fs, cfg);
    if (n < 0) fail(B2CNN_EINVAL, "b2cnn_prep_window_count: bad record shape / configuration");
    return n;
}
extern "C" int64_t b2cnn_prep_workspace_bytes(int64_t n_samples, double fs, int32_t n_sel, const b2cnn_prep_config *cfg) {
    const int64_t n = prep_workspace_bytes(n_samples, fs, n_sel, cfg);
    if (n < 0) fail(B2CNN_EINVAL, "b2cnn_prep_workspace_bytes: bad record shape / configuration");
    return n;
}
extern "C" int b2cnn_prep_windows(const int16_t *raw, int64_t n_samples, int32_t n_sig, const int32_t *sel, int32_t n_sel,
                                  const double *gains, const double *baselines, double fs, const b2cnn_prep_config *cfg,
                                  void *x_out, int dtype, double *t0_out, void *workspace, int64_t workspace_bytes, void *stream) {
    const char *err = "";
    const int rc = prep_windows(raw, n_samples, n_sig, sel, n_sel, gains, baselines, fs, cfg, x_out, dtype, t0_out, workspace,
                                workspace_bytes, reinterpret_cast<cudaStream_t>(stream), &err);
    return rc == B2CNN_OK ? rc : fail(rc, std::string("b2cnn_prep_windows: ") + err);
}
// ---- streaming form: per-patient device ring buffers (b2cnn_prep.cu) ----
struct b2cnn_ring { Ring *r; int device; };

extern "C" int b2cnn_ring_create(const b2cnn_prep_config *cfg, int32_t n_patients, int32_t n_sig, double fs, int32_t device,
                                 b2cnn_ring **out) {
    if (!cfg || !out) return fail(B2CNN_EINVAL, "b2cnn_ring_create: null argument");
    *out = nullptr;
    int dev = device;
    if (dev < 0) CU_TRY(cudaGetDevice(&dev));
    DEVICE_GUARD(dev);
    const char *err = "";
    Ring *r = nullptr;
    const int rc = ring_create(cfg, n_patients, n_sig, fs, dev, &r, &err);
    if (rc != B2CNN_OK) return fail(rc, std::string("b2cnn_ring_create: ") + err);
    b2cnn_ring *h = new (std::nothrow) b2cnn_ring{r, dev};
    if (!h) { ring_destroy(r); return fail(B2CNN_ESTATE, "out of host memory"); }
    *out = h;
    return B2CNN_OK;
}
extern "C" void b2cnn_ring_destroy(b2cnn_ring *h) {
    if (!h) return;
    DeviceGuard guard(h->device);
    ring_destroy(h->r);
    delete h;
}
extern "C" int b2cnn_ring_reset(b2cnn_ring *h, void *stream) {
    if (!h) return fail(B2CNN_EINVAL, "b2cnn_ring_reset: null argument");
    DEVICE_GUARD(h->device);
    const char *err = "";
    const int rc = ring_reset(h->r, reinterpret_cast<cudaStream_t>(stream), &err);
    return rc == B2CNN_OK ? rc : fail(rc, std::string("b2cnn_ring_reset: ") + err);
}
extern "C" int b2cnn_ring_set_signals(b2cnn_ring *h, int32_t patient, const int32_t *sel, int32_t n_sel, const double *gains,
                                      const double *baselines, void *stream) {
    if (!h) return fail(B2CNN_EINVAL, "b2cnn_ring_set_signals: null argument");
    DEVICE_GUARD(h->device);
    const char *err = "";
    const int rc = ring_set_signals(h->r, patient, sel, n_sel, gains, baselines, reinterpret_cast<cudaStream_t>(stream), &err);
    return rc == B2CNN_OK ? rc : fail(rc, std::string("b2cnn_ring_set_signals: ") + err);
}
extern "C" int b2cnn_ring_push(b2cnn_ring *h, const void *new_samples, int sample_kind, int64_t n_new, void *x_out, int dtype,
                               int32_t *emitted, int64_t *window_index, double *t0_seconds, void *stream) {
    if (!h) return fail(B2CNN_EINVAL, "b2cnn_ring_push: null argument");
    if (sample_kind != B2CNN_SAMPLES_ADC16 && sample_kind != B2CNN_SAMPLES_F64 && sample_kind != B2CNN_SAMPLES_GRID)
        return fail(B2CNN_EINVAL, "b2cnn_ring_push: sample_kind must be B2CNN_SAMPLES_ADC16, _F64 or _GRID");
    DEVICE_GUARD(h->device);
    const char *err = "";
    int em = 0;
    const int rc = ring_push(h->r, new_samples, sample_kind, n_new, x_out, dtype, &em, window_index,
                             t0_seconds, reinterpret_cast<cudaStream_t>(stream), &err);
    if (rc != B2CNN_OK) return fail(rc, std::string("b2cnn_ring_push: ") + err);
    if (emitted) *emitted = em;
    return B2CNN_OK;
}

// ---- wire formats (b2cnn_wire.cu) ----
extern "C" int b2cnn_decode_sample_messages(const void *bytes, const int64_t *offsets, int64_t n_msgs, int32_t *idx_out, double *val_out,
                                            const int64_t *row_of_msg, double *frame, int64_t frame_rows, int32_t n_sig, int32_t *n_bad,
                                            void *stream) {
    const char *err = "";
    const int rc = wire_decode_pairs(reinterpret_cast<const uint8_t *>(bytes), offsets, n_msgs, idx_out, val_out, row_of_msg, frame,
                                     frame_rows, n_sig, n_bad, reinterpret_cast<cudaStream_t>(stream), &err);
    return rc == B2CNN_OK ? rc : fail(rc, std::string("b2cnn_decode_sample_messages: ") + err);
}
extern "C" int b2cnn_decode_array_messages(const void *bytes, const int64_t *offsets, int64_t n_msgs, int32_t max_vals, double *vals_out,
                                           int32_t *counts_out, int32_t *n_bad, void *stream) {
    const char *err = "";
    const int rc = wire_decode_arrays(reinterpret_cast<const uint8_t *>(bytes), offsets, n_msgs, max_vals, vals_out, counts_out, n_bad,
                                      reinterpret_cast<cudaStream_t>(stream), &err);
    return rc == B2CNN_OK ? rc : fail(rc, std::string("b2cnn_decode_array_messages: ") + err);
}
extern "C" int b2cnn_frame_check(const void *frame, int64_t bytes, b2cnn_frame_header *header_out, int64_t *ids_offset,
                                 int64_t *samples_offset) {
    if (!frame || bytes < (int64_t)sizeof(b2cnn_frame_header)) return fail(B2CNN_EINVAL, "b2cnn_frame_check: buffer shorter than a header");
    b2cnn_frame_header hd;
    memcpy(&hd, frame, sizeof hd);
    if (hd.magic != B2CNN_FRAME_MAGIC || hd.version != 1) return fail(B2CNN_EINVAL, "b2cnn_frame_check: bad magic / version");
    if (hd.kind > B2CNN_SAMPLES_GRID || hd.n_patients < 1 || hd.n_new < 1 || hd.n_sig < 1 || hd.n_sig > 64)
        return fail(B2CNN_EINVAL, "b2cnn_frame_check: bad kind / shape");
    const int64_t esz = hd.kind == B2CNN_SAMPLES_ADC16 ? 2 : 8;
    const int64_t ids = sizeof hd, smp = (ids + 4ll * hd.n_patients + 7) / 8 * 8;
    // the header is untrusted input: three 32-bit counts can wrap a 64-bit product, so the size is formed in 128 bits
    const unsigned __int128 need128 = (unsigned __int128)smp + (unsigned __int128)esz * hd.n_patients * hd.n_new * hd.n_sig;
    const int64_t need = need128 > (unsigned __int128)INT64_MAX ? INT64_MAX : (int64_t)need128;
    if (bytes != need) {
        char buf[128];
        snprintf(buf, sizeof buf, "b2cnn_frame_check: frame is %lld bytes, its header describes %lld", (long long)bytes, (long long)need);
        return fail(B2CNN_EINVAL, buf);
    }
    if (header_out) *header_out = hd;
    if (ids_offset) *ids_offset = ids;
    if (samples_offset) *samples_offset = smp;
    return B2CNN_OK;
}
extern "C" double b2cnn_parse_decimal(const char *s, int64_t len, int32_t *status) {
    int st = 0;
    const double v = wire_parse_decimal_host(s, len, &st);
    if (status) *status = st;
    return v;
}

extern "C" const char *b2cnn_version(void) { return "b2cnn 0.3 (sm_100a; tcgen05 fused bf16 path, fp32 streaming path, generic path, device preprocessing + patient ring buffers)"; }

extern "C" int b2cnn_create(const b2cnn_config *cfg, b2cnn_handle **out) {
    if (!cfg || !out) return fail(B2CNN_EINVAL, "b2cnn_create: null argument");
    *out = nullptr;
    Dims d;
    if (!derive_dims(*cfg, d)) return fail(B2CNN_EINVAL, "b2cnn_create: window too short for this conv/pool stack");
    if (cfg->c_mid != kCMid || cfg->hidden != kHidden || cfg->layers != 2)
        return fail(B2CNN_EARCH, "b2cnn_create: only conv1 out_channels=4, LSTM hidden=16, layers=2 (bin/models.py:10,16) are supported");
    if (kCMid * d.C * d.K1 > kMaxW1) return fail(B2CNN_EARCH, "b2cnn_create: in_channels*k1 > 200 not supported");
    if (d.K2 > kMaxK2) return fail(B2CNN_EARCH, "b2cnn_create: k2 > 8 not supported");
    if (d.C > 16) return fail(B2CNN_EARCH, "b2cnn_create: in_channels > 16 not supported");
    if (cfg->act < 0 || cfg->act > 2) return fail(B2CNN_EINVAL, "b2cnn_create: bad activation");
    if (cfg->lstm_input != d.L) {
        char buf[256];
        snprintf(buf, sizeof buf,
                 "b2cnn_create: L_out(window=%d)=%d != lstm_input=%d: x.view(-1, MAGICNUM) would straddle windows (bin/models.py:29)",
                 d.W, d.L, cfg->lstm_input);
        return fail(B2CNN_EVIEW, buf);
    }
    int dev = cfg->device;
    if (dev < 0) CU_TRY(cudaGetDevice(&dev));
    DEVICE_GUARD(dev);
    b2cnn_handle *h = new (std::nothrow) b2cnn_handle();
    if (!h) return fail(B2CNN_ESTATE, "out of host memory");
    h->cfg = *cfg; h->d = d; h->device = dev;
    cudaError_t e = cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) { delete h; return cuda_fail(e, "cudaDeviceGetAttribute"); }
    h->n_weights = b2cnn_weight_count(cfg);
    e = cudaMalloc(&h->d_blob, sizeof(float) * h->n_weights);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_wih0T, sizeof(float) * (size_t)d.L * kGates);
    if (e != cudaSuccess) { b2cnn_destroy(h); return cuda_fail(e, "cudaMalloc(weights)"); }
    *out = h;
    return B2CNN_OK;
}

extern "C" void b2cnn_destroy(b2cnn_handle *h) {
    if (!h) return;
    DeviceGuard guard(h->device);
    tc_release(h->tc);
    cudaFree(h->d_blob); cudaFree(h->d_wih0T);
    for (int i = 0; i < 3; ++i)
        if (h->ev_stage[i]) cudaEventDestroy(h->ev_stage[i]);
    for (int i = 0; i < 2; ++i) {
        cudaFree(h->st_x[i]);
        if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]);
        if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
    }
    cudaFree(h->st_age); cudaFree(h->st_out); cudaFree(h->st_ws);
    if (h->s_copy) cudaStreamDestroy(h->s_copy);
    if (h->s_comp) cudaStreamDestroy(h->s_comp);
    delete h;
}

extern "C" int b2cnn_set_weights(b2cnn_handle *h, const float *blob, int64_t n, int on_device, void *stream) {
    if (!h || !blob) return fail(B2CNN_EINVAL, "b2cnn_set_weights: null argument");
    if (n != h->n_weights) {
        char buf[160];
        snprintf(buf, sizeof buf, "b2cnn_set_weights: expected %lld floats, got %lld", (long long)h->n_weights, (long long)n);
        return fail(B2CNN_EINVAL, buf);
    }
    cudaStream_t st = (cudaStream_t)stream;
    DEVICE_GUARD(h->device);
    const Dims &d = h->d;
    CU_TRY(cudaMemcpyAsync(h->d_blob, blob, sizeof(float) * n, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    // conv weights + affine -> host copy for the kernel-parameter constant bank
    const int64_t n_conv = (int64_t)kCMid * d.C * d.K1 + kCMid + kCMid * d.K2 + 1;
    std::vector<float> conv(n_conv), aff(2 * kCMid + 2, 0.f);
    const int64_t aff_off = n - (2 * kCMid + 2);
    if (on_device) {
        CU_TRY(cudaMemcpyAsync(conv.data(), blob, sizeof(float) * n_conv, cudaMemcpyDeviceToHost, st));
        if (d.has_affine) CU_TRY(cudaMemcpyAsync(aff.data(), blob + aff_off, sizeof(float) * aff.size(), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
    } else {
        memcpy(conv.data(), blob, sizeof(float) * n_conv);
        if (d.has_affine) memcpy(aff.data(), blob + aff_off, sizeof(float) * aff.size());
    }
    ConvWeights &cw = h->cw;
    memset(&cw, 0, sizeof cw);
    const float *w1 = conv.data(), *b1 = w1 + (int64_t)kCMid * d.C * d.K1, *w2 = b1 + kCMid, *b2 = w2 + kCMid * d.K2;
    for (int o = 0; o < kCMid; ++o)
        for (int c = 0; c < d.C; ++c)
            for (int k = 0; k < d.K1; ++k) cw.w1[(c * d.K1 + k) * kCMid + o] = w1[((int64_t)o * d.C + c) * d.K1 + k];
    for (int o = 0; o < kCMid; ++o) cw.b1[o] = b1[o];
    for (int i = 0; i < kCMid * d.K2; ++i) cw.w2[i] = w2[i];
    cw.b2 = b2[0];
    for (int o = 0; o < kCMid; ++o) { cw.s1[o] = d.has_affine ? aff[o] : 1.f; cw.t1[o] = d.has_affine ? aff[kCMid + o] : 0.f; }
    cw.s2 = d.has_affine ? aff[2 * kCMid] : 1.f;
    cw.t2 = d.has_affine ? aff[2 * kCMid + 1] : 0.f;
    // head pointers into the device blob
    const float *p = h->d_blob + n_conv;
    const float *wih0 = p; p += (int64_t)kGates * d.L;
    h->hw.whh0 = p; p += kGates * kHidden;
    h->hw.bih0 = p; p += kGates;
    h->hw.bhh0 = p; p += kGates;
    h->hw.wih1 = p; p += kGates * kHidden;
    h->hw.whh1 = p; p += kGates * kHidden;
    h->hw.bih1 = p; p += kGates;
    h->hw.bhh1 = p; p += kGates;
    h->hw.wo = p; p += kHidden;
    h->hw.bo = p;
    h->hw.wih0T = h->d_wih0T;
    launch_transpose_wih(wih0, h->d_wih0T, d.L, st);
    CU_TRY(cudaGetLastError());
    int rc = tc_prepare(h->tc, d, cw, wih0, h->hw, (int)h->opt_tc_splits, h->num_sms, st);
    if (rc != 0) return fail(B2CNN_ECUDA, std::string("tc_prepare: ") + tc_error());
    h->weights_set = true;
    return B2CNN_OK;
}

// Which kernels a (dtype, B, mode) call takes decides its scratch: the streaming tensor-core kernels need the range
// partials [slices][B][64] (+ [B][64] gates for a sequence scan) and a few ints per window; only the generic and the
// unfused tensor-core paths round-trip feature rows [B][L] through HBM.
struct WsLayout { int64_t feats, partial, gates, tc, total; int ks_ws; };

static bool use_tc(b2cnn_handle *h, int dtype, int64_t B, int mode);

static WsLayout ws_layout(b2cnn_handle *h, int64_t B, int mode, int dtype) {
    const Dims &d = h->d;
    WsLayout w;
    const int ks = choose_ksplit(B, d.L, h->num_sms);
    w.ks_ws = tc_partial_slices(h->tc) > ks ? tc_partial_slices(h->tc) : ks;
    bool streaming = false;
    if (dtype >= 0 && h->opt_path != B2CNN_PATH_GENERIC) {
        if (dtype == B2CNN_DTYPE_F32) streaming = h->opt_stream && tc_stream_supported(h->tc, d, dtype);
        else streaming = use_tc(h, dtype, B, mode) && tc_fused_supported(h->tc, d, dtype);
    }
    w.feats = streaming ? 0 : align_up(B * d.L * 4, 256);
    w.partial = align_up((int64_t)w.ks_ws * B * kGates * 4, 256);
    w.gates = align_up(B * kGates * 4, 256);
    w.tc = tc_workspace_bytes(h->tc, d, B);
    w.total = w.feats + w.partial + w.gates + w.tc;
    return w;
}

// dtype-blind size: enough for whatever path a call may take (the generic path's feature rows included)
extern "C" int64_t b2cnn_workspace_bytes(b2cnn_handle *h, int64_t B, int mode) {
    if (!h || B < 1) return -1;
    return ws_layout(h, B, mode, -1).total;
}

// exact size for windows of `dtype`: 307 MB smaller at [4096,3,75000] bf16, where the features never leave the SM
extern "C" int64_t b2cnn_workspace_bytes_for(b2cnn_handle *h, int64_t B, int mode, int dtype) {
    if (!h || B < 1 || (dtype != B2CNN_DTYPE_F32 && dtype != B2CNN_DTYPE_BF16)) return -1;
    return ws_layout(h, B, mode, dtype).total;
}

static bool use_tc(b2cnn_handle *h, int dtype, int64_t B, int mode) {
    if (h->opt_path == B2CNN_PATH_GENERIC) return false;
    return tc_supported(h->tc, h->d, dtype, B, mode);
}

static int forward_device(b2cnn_handle *h, const void *x, int dtype, int64_t B, int64_t xpitch, const float *age, int64_t n_age,
                          int mode, int apply_sigmoid, float *out, void *ws, int64_t ws_bytes, cudaStream_t st) {
    Dims d = h->d;
    if (xpitch < d.W || xpitch > 0x7fffffff) return fail(B2CNN_EINVAL, "b2cnn_forward: x_pitch must be >= window");
    d.XP = (int)xpitch;
    const int ks = choose_ksplit(B, d.L, h->num_sms);
    const WsLayout wl = ws_layout(h, B, mode, dtype);
    if (ws_bytes < wl.total) return fail(B2CNN_ESTATE, "b2cnn_forward: workspace smaller than b2cnn_workspace_bytes_for()");
    char *base = (char *)ws;
    float *feats = wl.feats ? (float *)base : nullptr; base += wl.feats;
    float *partial = (float *)base; base += wl.partial;
    float *gates = (float *)base; base += wl.gates;
    void *tc_ws = base;
    const char *err = "";
    int launches = 0;
    const bool tc = use_tc(h, dtype, B, mode);
    const bool stream = h->opt_path != B2CNN_PATH_GENERIC && h->opt_stream && tc_stream_supported(h->tc, d, dtype);
    if (!wl.feats && !stream && !(tc && tc_fused_supported(h->tc, d, dtype)) &&
        !(h->opt_path != B2CNN_PATH_TENSORCORE && h->opt_small && (mode == B2CNN_MODE_INDEPENDENT || B == 1) && B <= 256 &&
          (int64_t)d.C * d.W <= 8192 && small_supported(d)) &&
        !(h->opt_path != B2CNN_PATH_TENSORCORE && h->opt_small && mode == B2CNN_MODE_INDEPENDENT && B >= 8 && batch_supported(d)))
        return fail(B2CNN_ESTATE, "b2cnn_forward: this x_pitch takes a path that needs feature rows; size the workspace with b2cnn_workspace_bytes()");
    if (!tc && !stream && h->opt_path == B2CNN_PATH_TENSORCORE)
        return fail(B2CNN_EARCH, "path=tensorcore requested but this shape/dtype/mode is not supported by the tcgen05 kernel");
    const bool prof = h->opt_profile != 0;
    if (prof) {
        for (int i = 0; i < 3; ++i)
            if (!h->ev_stage[i]) CU_TRY(cudaEventCreate(&h->ev_stage[i]));
        CU_TRY(cudaEventRecord(h->ev_stage[0], st));
    }
    // short windows, many of them (all patients of a trigger, [P,10,120]): one launch, one warp per window
    if (h->opt_path != B2CNN_PATH_TENSORCORE && h->opt_small && mode == B2CNN_MODE_INDEPENDENT && B >= 8 && batch_supported(d)) {
        int n1 = launch_short_batch(d, h->cw, h->hw, x, dtype, B, age, n_age, apply_sigmoid, out, h->num_sms, st, &err);
        if (n1 < 0) return fail(B2CNN_ECUDA, std::string("short-window batch kernel: ") + err);
        if (prof) { CU_TRY(cudaEventRecord(h->ev_stage[1], st)); CU_TRY(cudaEventRecord(h->ev_stage[2], st)); h->ev_valid = true; }
        h->last_launches = n1; h->last_path = B2CNN_PATH_GENERIC;
        return B2CNN_OK;
    }
    // short windows, few of them (the production call is [1,10,120]): one launch does everything
    if (h->opt_path != B2CNN_PATH_TENSORCORE && h->opt_small && (mode == B2CNN_MODE_INDEPENDENT || B == 1) && B <= 256 &&
        (int64_t)d.C * d.W <= 8192 && small_supported(d)) {
        int n1 = launch_small_forward(d, h->cw, h->hw, x, dtype, B, age, n_age, apply_sigmoid, out, st, &err);
        if (n1 < 0) return fail(B2CNN_ECUDA, std::string("small-window kernel: ") + err);
        if (prof) { CU_TRY(cudaEventRecord(h->ev_stage[1], st)); CU_TRY(cudaEventRecord(h->ev_stage[2], st)); h->ev_valid = true; }
        h->last_launches = n1; h->last_path = B2CNN_PATH_GENERIC;
        return B2CNN_OK;
    }
    // feature layout: the generic kernel writes rows [B][L]; the tensor-core kernel's threads
    // are windows, so it writes the transpose [L][B] (coalesced across lanes).
    int64_t sB = d.L, sP = 1;
    int n;
    if (h->opt_path != B2CNN_PATH_GENERIC && h->opt_stream && tc_stream_supported(h->tc, d, dtype)) {
        // fp32 windows: TMA-streamed CUDA-core conv1 + tcgen05 projection, features never leave the SM
        const bool indep = mode == B2CNN_MODE_INDEPENDENT;
        int slices = 0;
        n = tc_stream_gates(h->tc, d, h->cw, h->hw, x, B, feats, partial, gates, tc_ws, h->num_sms, st, &err, !indep, &slices);
        if (n < 0) return fail(B2CNN_ECUDA, std::string("fp32 stream kernel: ") + err);
        launches += n;
        if (prof) CU_TRY(cudaEventRecord(h->ev_stage[1], st));
        // the head kernel is the last reader of the call's exception list: it also puts the handle's flag state back to zero
        const bool cleans = indep && h->tc.cur_own;
        n = indep ? launch_reduce_lstm_head(d, h->hw, partial, slices, B, age, n_age, apply_sigmoid, out, st, &err,
                                            cleans ? h->tc.cur_count : nullptr, h->tc.cur_flags, h->tc.cur_list)
                  : launch_lstm_head(d, h->hw, gates, B, age, n_age, mode, apply_sigmoid, out, st, &err);
        if (n < 0) return fail(B2CNN_ECUDA, std::string("head: ") + err);
        if (cleans) h->tc.flags_clean = true;
        launches += n;
        if (prof) { CU_TRY(cudaEventRecord(h->ev_stage[2], st)); h->ev_valid = true; }
        h->last_launches = launches; h->last_path = B2CNN_PATH_STREAM;
        return B2CNN_OK;
    }
    if (tc && tc_fused_supported(h->tc, d, dtype)) {
        // conv + pool + projection fused on the tensor cores: the features never leave the SM
        const bool indep = mode == B2CNN_MODE_INDEPENDENT;
        int slices = 0;
        n = tc_fused_gates(h->tc, d, h->cw, h->hw, x, B, feats, partial, gates, tc_ws, h->num_sms, st, &err, !indep, &slices);
        if (n < 0) return fail(B2CNN_ECUDA, std::string("tensor-core fused kernel: ") + err);
        launches += n;
        if (prof) CU_TRY(cudaEventRecord(h->ev_stage[1], st));
        // the head kernel is the last reader of the call's exception list: it also puts the handle's flag state back to zero
        const bool cleans = indep && h->tc.cur_own;
        n = indep ? launch_reduce_lstm_head(d, h->hw, partial, slices, B, age, n_age, apply_sigmoid, out, st, &err,
                                            cleans ? h->tc.cur_count : nullptr, h->tc.cur_flags, h->tc.cur_list)
                  : launch_lstm_head(d, h->hw, gates, B, age, n_age, mode, apply_sigmoid, out, st, &err);
        if (n < 0) return fail(B2CNN_ECUDA, std::string("head: ") + err);
        if (cleans) h->tc.flags_clean = true;
        launches += n;
        if (prof) { CU_TRY(cudaEventRecord(h->ev_stage[2], st)); h->ev_valid = true; }
        h->last_launches = launches; h->last_path = B2CNN_PATH_TENSORCORE;
        return B2CNN_OK;
    }
    if (tc) {
        sB = 1; sP = B;
        n = tc_frontend(h->tc, d, h->cw, x, B, feats, sB, sP, tc_ws, h->num_sms, st, &err);
        if (n < 0) return fail(B2CNN_ECUDA, std::string("tensor-core front end: ") + err);
    } else {
        n = launch_frontend_generic(d, h->cw, x, dtype, B, feats, sB, sP, st, h->num_sms, &err);
        if (n < 0) return fail(B2CNN_ECUDA, std::string("front end: ") + err);
    }
    launches += n;
    if (prof) CU_TRY(cudaEventRecord(h->ev_stage[1], st));
    n = launch_head(d, h->hw, feats, sB, sP, B, age, n_age, mode, apply_sigmoid, out, gates, partial, ks, st, &err);
    if (n < 0) return fail(B2CNN_ECUDA, std::string("head: ") + err);
    launches += n;
    if (prof) { CU_TRY(cudaEventRecord(h->ev_stage[2], st)); h->ev_valid = true; }
    h->last_launches = launches; h->last_path = tc ? B2CNN_PATH_TENSORCORE : B2CNN_PATH_GENERIC;
    return B2CNN_OK;
}

static int check_call(b2cnn_handle *h, const void *x, int dtype, int64_t B, const float *age, int64_t n_age, int mode, float *out) {
    if (!h || !x || !age || !out) return fail(B2CNN_EINVAL, "b2cnn_forward: null argument");
    if (!h->weights_set) return fail(B2CNN_ESTATE, "b2cnn_forward: weights not set (call b2cnn_set_weights)");
    if (B < 1 || B > (int64_t)0x7fffffff / 64) return fail(B2CNN_EINVAL, "b2cnn_forward: batch size out of range");
    if (dtype != B2CNN_DTYPE_F32 && dtype != B2CNN_DTYPE_BF16) return fail(B2CNN_EINVAL, "b2cnn_forward: dtype must be f32 (0) or bf16 (1)");
    if (mode != B2CNN_MODE_INDEPENDENT && mode != B2CNN_MODE_SEQUENCE) return fail(B2CNN_EINVAL, "b2cnn_forward: bad mode");
    if (n_age != 1 && n_age != B) return fail(B2CNN_EINVAL, "b2cnn_forward: age must have 1 or B elements");
    return B2CNN_OK;
}

extern "C" int b2cnn_forward(b2cnn_handle *h, const void *x, int dtype, int64_t B, const float *age, int64_t n_age,
                             int mode, int apply_sigmoid, float *out, void *workspace, int64_t workspace_bytes, void *stream) {
    int rc = check_call(h, x, dtype, B, age, n_age, mode, out);
    if (rc) return rc;
    if (!workspace || workspace_bytes < ws_layout(h, B, mode, dtype).total)
        return fail(B2CNN_ESTATE, "b2cnn_forward: workspace missing or smaller than b2cnn_workspace_bytes_for()");
    DEVICE_GUARD(h->device);
    return forward_device(h, x, dtype, B, h->d.W, age, n_age, mode, apply_sigmoid, out, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int b2cnn_forward_pitched(b2cnn_handle *h, const void *x, int dtype, int64_t B, int64_t x_pitch, const float *age,
                                     int64_t n_age, int mode, int apply_sigmoid, float *out, void *workspace,
                                     int64_t workspace_bytes, void *stream) {
    int rc = check_call(h, x, dtype, B, age, n_age, mode, out);
    if (rc) return rc;
    if (!workspace || workspace_bytes < ws_layout(h, B, mode, dtype).total)
        return fail(B2CNN_ESTATE, "b2cnn_forward_pitched: workspace missing or smaller than b2cnn_workspace_bytes_for()");
    DEVICE_GUARD(h->device);
    return forward_device(h, x, dtype, B, x_pitch, age, n_age, mode, apply_sigmoid, out, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int b2cnn_features(b2cnn_handle *h, const void *x, int dtype, int64_t B, float *feats, void *stream) {
    if (!h || !x || !feats) return fail(B2CNN_EINVAL, "b2cnn_features: null argument");
    if (!h->weights_set) return fail(B2CNN_ESTATE, "b2cnn_features: weights not set");
    if (B < 1) return fail(B2CNN_EINVAL, "b2cnn_features: B < 1");
    DEVICE_GUARD(h->device);
    const char *err = "";
    int n;
    if (use_tc(h, dtype, B, B2CNN_MODE_INDEPENDENT) && tc_can_emit_features(h->tc)) {
        n = tc_features(h->tc, h->d, h->cw, x, B, feats, h->num_sms, (cudaStream_t)stream, &err);
        h->last_path = B2CNN_PATH_TENSORCORE;
    } else {
        if (h->opt_path == B2CNN_PATH_TENSORCORE) return fail(B2CNN_EARCH, "b2cnn_features: tensor-core path unavailable for this shape");
        n = launch_frontend_generic(h->d, h->cw, x, dtype, B, feats, h->d.L, 1, (cudaStream_t)stream, h->num_sms, &err);
        h->last_path = B2CNN_PATH_GENERIC;
    }
    if (n < 0) return fail(B2CNN_ECUDA, std::string("front end: ") + err);
    h->last_launches = n;
    return B2CNN_OK;
}

// ---- host-pointer entry: chunked H2D overlapped with compute -----------------------------
static int ensure_host_staging(b2cnn_handle *h, size_t x_chunk_bytes, int64_t B, size_t ws_bytes) {
    if (!h->s_copy) {
        CU_TRY(cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking));
        CU_TRY(cudaStreamCreateWithFlags(&h->s_comp, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CU_TRY(cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming));
            CU_TRY(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
        }
    }
    if (x_chunk_bytes > h->st_x_bytes) {
        for (int i = 0; i < 2; ++i) { cudaFree(h->st_x[i]); h->st_x[i] = nullptr; CU_TRY(cudaMalloc(&h->st_x[i], x_chunk_bytes)); }
        h->st_x_bytes = x_chunk_bytes;
    }
    if ((size_t)B > h->st_vec_elems) {
        cudaFree(h->st_age); cudaFree(h->st_out); h->st_age = h->st_out = nullptr;
        CU_TRY(cudaMalloc(&h->st_age, sizeof(float) * B));
        CU_TRY(cudaMalloc(&h->st_out, sizeof(float) * B));
        h->st_vec_elems = (size_t)B;
    }
    if (ws_bytes > h->st_ws_bytes) {
        cudaFree(h->st_ws); h->st_ws = nullptr;
        CU_TRY(cudaMalloc(&h->st_ws, ws_bytes));
        h->st_ws_bytes = ws_bytes;
    }
    return B2CNN_OK;
}

extern "C" int b2cnn_forward_host(b2cnn_handle *h, const void *x_host, int dtype, int64_t B, const float *age_host,
                                  int64_t n_age, int mode, int apply_sigmoid, float *out_host) {
    int rc = check_call(h, x_host, dtype, B, age_host, n_age, mode, out_host);
    if (rc) return rc;
    DEVICE_GUARD(h->device);
    const Dims &d = h->d;
    const size_t esz = dtype == B2CNN_DTYPE_BF16 ? 2 : 4;
    const size_t win_bytes = (size_t)d.C * d.W * esz;
    // independent windows: stream in chunks of ~64 MiB; a sequence scan needs the whole batch.
    int64_t chunk = B;
    if (mode == B2CNN_MODE_INDEPENDENT) {
        chunk = (int64_t)((64u << 20) / win_bytes);
        if (chunk < 1) chunk = 1;
        if (chunk > B) chunk = B;
    }
    const int64_t ws_bytes = ws_layout(h, chunk, mode, dtype).total;
    rc = ensure_host_staging(h, (size_t)chunk * win_bytes, B, (size_t)ws_bytes);
    if (rc) return rc;
    CU_TRY(cudaMemcpyAsync(h->st_age, age_host, sizeof(float) * n_age, cudaMemcpyHostToDevice, h->s_copy));
    int64_t launches = 0;
    int idx = 0;
    for (int64_t b0 = 0; b0 < B; b0 += chunk, ++idx) {
        const int64_t nb = (B - b0 < chunk) ? (B - b0) : chunk;
        const int s = idx & 1;
        if (idx >= 2) CU_TRY(cudaStreamWaitEvent(h->s_copy, h->ev_done[s], 0));
        CU_TRY(cudaMemcpyAsync(h->st_x[s], (const char *)x_host + (size_t)b0 * win_bytes, (size_t)nb * win_bytes,
                               cudaMemcpyHostToDevice, h->s_copy));
        CU_TRY(cudaEventRecord(h->ev_copied[s], h->s_copy));
        CU_TRY(cudaStreamWaitEvent(h->s_comp, h->ev_copied[s], 0));
        rc = forward_device(h, h->st_x[s], dtype, nb, d.W, n_age == 1 ? h->st_age : h->st_age + b0, n_age == 1 ? 1 : nb, mode,
                            apply_sigmoid, h->st_out + b0, h->st_ws, ws_bytes, h->s_comp);
        if (rc) return rc;
        launches += h->last_launches;
        CU_TRY(cudaEventRecord(h->ev_done[s], h->s_comp));
    }
    CU_TRY(cudaMemcpyAsync(out_host, h->st_out, sizeof(float) * B, cudaMemcpyDeviceToHost, h->s_comp));
    CU_TRY(cudaStreamSynchronize(h->s_comp));
    CU_TRY(cudaStreamSynchronize(h->s_copy));
    h->last_launches = launches;
    return B2CNN_OK;
}

extern "C" int b2cnn_set_option(b2cnn_handle *h, const char *key, int64_t value) {
    if (!h || !key) return fail(B2CNN_EINVAL, "b2cnn_set_option: null argument");
    if (!strcmp(key, "path")) {
        if (value < 0 || value > 2) return fail(B2CNN_EINVAL, "path must be 0 (auto), 1 (generic) or 2 (tensorcore)");
        h->opt_path = value;
        return B2CNN_OK;
    }
    if (!strcmp(key, "small_kernel")) { h->opt_small = value ? 1 : 0; return B2CNN_OK; }
    if (!strcmp(key, "tc_fused")) { h->tc.opt_fused = value ? 1 : 0; return B2CNN_OK; }
    if (!strcmp(key, "stream_f32")) { h->opt_stream = value ? 1 : 0; return B2CNN_OK; }
    if (!strcmp(key, "profile")) { h->opt_profile = value ? 1 : 0; h->ev_valid = false; return B2CNN_OK; }
    if (!strcmp(key, "tc_splits")) {
        if (value != 2 && value != 3) return fail(B2CNN_EINVAL, "tc_splits must be 2 or 3");
        if (h->weights_set && value != h->opt_tc_splits) return fail(B2CNN_ESTATE, "set tc_splits before b2cnn_set_weights");
        h->opt_tc_splits = value;
        return B2CNN_OK;
    }
    return fail(B2CNN_EINVAL, std::string("unknown option: ") + key);
}

extern "C" int64_t b2cnn_get_option(b2cnn_handle *h, const char *key) {
    if (!h || !key) return -1;
    if (!strcmp(key, "path")) return h->opt_path;
    if (!strcmp(key, "tc_splits")) return h->opt_tc_splits;
    if (!strcmp(key, "num_sms")) return h->num_sms;
    if (!strcmp(key, "profile")) return h->opt_profile;
    if (!strcmp(key, "tc_available")) return tc_supported(h->tc, h->d, B2CNN_DTYPE_BF16, 128, B2CNN_MODE_INDEPENDENT) ? 1 : 0;
    return -1;
}

extern "C" int64_t b2cnn_last_launch_count(b2cnn_handle *h) { return h ? h->last_launches : -1; }
extern "C" int b2cnn_last_path(b2cnn_handle *h) { return h ? h->last_path : -1; }

extern "C" double b2cnn_last_stage_ms(b2cnn_handle *h, int stage) {
    if (!h || !h->ev_valid || stage < 0 || stage > 1) return -1.0;
    if (cudaEventSynchronize(h->ev_stage[stage + 1]) != cudaSuccess) return -1.0;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev_stage[stage], h->ev_stage[stage + 1]) != cudaSuccess) return -1.0;
    return (double)ms;
}
