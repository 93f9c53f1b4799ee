/*
 * b2cnn.h -- C ABI of the B200-native MyCNN forward pass (libb2cnn.so).
 *
 * The reference (travistangvh/time-series-kafka-demo) has no plugin/FFI seam: the only
 * boundary of its hot path is the Python call `output = model(x_arr, a_arr)`
 * (bin/predictStream.py:157; also bin/utils.py:204,249,682) into `MyCNN.forward`
 * (bin/models.py:22-36).  These entry points are what a ctypes/cffi binding for that call
 * binds; INTEGRATION.md shows the stub.  Plain pointers and sizes only -- no torch types.
 *
 * Ownership: the caller owns x / age / out / workspace; the library owns only its packed
 * weight buffers, 512 KB of NaN-exception state allocated with them (flags of the windows the
 * tensor-core kernels hand to the exact path; zero between calls, used by calls on the first
 * stream a handle sees -- other streams use a copy in the workspace) and, for b2cnn_forward_host,
 * its pinned/device staging buffers.  A handle serves one call at a time.
 * b2cnn_forward makes no allocation and is asynchronous on `stream`.
 * Errors: every call returns 0 on success or a B2CNN_E* code; b2cnn_last_error() returns a
 * thread-local message.  There is no CPU fallback anywhere in this library.
 */
#ifndef B2CNN_H_
#define B2CNN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2CNN_OK 0
#define B2CNN_EINVAL 1      /* bad argument / shape / dtype                                */
#define B2CNN_EARCH 2       /* architecture outside what the kernels support               */
#define B2CNN_EVIEW 3       /* L_out(window) != lstm_input: x.view(-1, MAGICNUM) would      */
                            /* straddle windows (bin/models.py:29) -- rejected, not guessed */
#define B2CNN_ECUDA 4       /* CUDA runtime / driver error                                  */
#define B2CNN_ESTATE 5      /* weights not set, workspace too small, ...                    */

enum { B2CNN_DTYPE_F32 = 0, B2CNN_DTYPE_BF16 = 1 };
/* INDEPENDENT: every window starts from the zero LSTM state == looping the reference one
 *   window at a time (bin/predictStream.py:70,157).
 * SEQUENCE: bit-for-bit the structure of model(x_batch): the LSTM scans the batch axis
 *   (bin/models.py:29-30 with B>1; bin/utils.py:249). */
enum { B2CNN_MODE_INDEPENDENT = 0, B2CNN_MODE_SEQUENCE = 1 };
enum { B2CNN_ACT_TANH = 0, B2CNN_ACT_RELU = 1, B2CNN_ACT_IDENTITY = 2 };
enum { B2CNN_PATH_AUTO = 0, B2CNN_PATH_GENERIC = 1, B2CNN_PATH_TENSORCORE = 2,
       B2CNN_PATH_STREAM = 3 /* reported by b2cnn_last_path only: fp32 windows, streamed CUDA-core conv1 + tcgen05 projection */ };

#define B2CNN_FLAG_AFFINE 1 /* per-channel scale/shift after each conv (folded eval-BatchNorm) */

/* Mirrors the constructor of MyCNN (bin/models.py:6-20). */
typedef struct b2cnn_config {
    int32_t in_channels; /* conv1 in_channels            models.py:10 (10; 7 in MyCNN2/3)    */
    int32_t k1;          /* conv1 kernel_size            models.py:10 (10; 5 in MyCNN2/3/4)  */
    int32_t c_mid;       /* conv1 out_channels           models.py:10 (must be 4)            */
    int32_t k2;          /* conv2 kernel_size            models.py:11 (5)                    */
    int32_t pool_k;      /* MaxPool1d kernel_size        models.py:12 (3; 2 in MyCNN2/3/4)   */
    int32_t pool_s;      /* MaxPool1d stride             models.py:12 (2)                    */
    int32_t hidden;      /* LSTM hidden_size             models.py:16 (must be 16)           */
    int32_t layers;      /* LSTM num_layers              models.py:16 (must be 2)            */
    int32_t window;      /* samples per window W         config.cfg:23 (120)                 */
    int32_t lstm_input;  /* MAGICNUM                     models.py:8  (must equal L_out(W))  */
    int32_t act;         /* B2CNN_ACT_*; the reference uses tanh (models.py:23,26)           */
    int32_t flags;       /* B2CNN_FLAG_*                                                     */
    float age_coef;      /* models.py:32 (1e-8)                                              */
    int32_t device;      /* CUDA device ordinal, or -1 for the current device                */
} b2cnn_config;

typedef struct b2cnn_handle b2cnn_handle;

/* L_out(W): conv1 -> pool -> conv2 -> pool output length (floor pooling); <=0 if invalid. */
int64_t b2cnn_l_out(const b2cnn_config *cfg);

/* Number of floats in the packed weight blob, in this order (== state_dict order of the
 * used tensors, bin/models.py:10-17):
 *   conv1.weight[4][C][K1], conv1.bias[4], conv2.weight[1][4][K2], conv2.bias[1],
 *   lstm.weight_ih_l0[64][L], lstm.weight_hh_l0[64][16], lstm.bias_ih_l0[64], lstm.bias_hh_l0[64],
 *   lstm.weight_ih_l1[64][16], lstm.weight_hh_l1[64][16], lstm.bias_ih_l1[64], lstm.bias_hh_l1[64],
 *   out.weight[16], out.bias[1]
 *   (+ if B2CNN_FLAG_AFFINE: scale1[4], shift1[4], scale2[1], shift2[1]) */
int64_t b2cnn_weight_count(const b2cnn_config *cfg);

/* Replaces `model = torch.load(path); model.eval()` (bin/predictStream.py:36-37). */
int b2cnn_create(const b2cnn_config *cfg, b2cnn_handle **out);
void b2cnn_destroy(b2cnn_handle *h);

/* Replaces load_state_dict: copies the packed blob (host or device memory) into the
 * library's device buffers, enqueued on `stream` (a cudaStream_t, may be NULL). */
int b2cnn_set_weights(b2cnn_handle *h, const float *blob, int64_t n_floats, int blob_on_device,
                      void *stream);

/* Bytes of caller-provided device scratch b2cnn_forward needs for a batch of B windows.
 * b2cnn_workspace_bytes is dtype-blind (enough for any path, the generic path's [B][L_out] feature rows included);
 * b2cnn_workspace_bytes_for is exact for windows of `dtype`: where the streaming tensor-core kernels apply, the
 * features never leave the SM and the scratch is the range partials only (39 MB instead of 346 MB at [4096,3,75000]). */
int64_t b2cnn_workspace_bytes(b2cnn_handle *h, int64_t B, int mode);
int64_t b2cnn_workspace_bytes_for(b2cnn_handle *h, int64_t B, int mode, int dtype);

/* Replaces `output = model(x, age)` (bin/predictStream.py:157).  All pointers are DEVICE
 * pointers.  x: [B][C][W] contiguous, dtype f32 or bf16.  age: n_age == B or 1 (broadcast).
 * out: [B] floats: the logit (bin/models.py:34), or sigmoid(logit) if apply_sigmoid
 * (bin/predictStream.py:160). */
int b2cnn_forward(b2cnn_handle *h, const void *x, int dtype, int64_t B, const float *age,
                  int64_t n_age, int mode, int apply_sigmoid, float *out, void *workspace,
                  int64_t workspace_bytes, void *stream);

/* b2cnn_forward for windows whose channel rows are `x_pitch` ELEMENTS apart (x_pitch >= W; window b starts at
 * x + b * C * x_pitch): a producer that pads its rows to a multiple of 16 bytes (8 bf16 / 4 fp32 samples) lets TMA
 * stream windows of ANY length straight from `x`; a contiguous tensor with W % 8 != 0 (7500, 37500 ...) has to be
 * re-pitched into scratch first (one extra read + write of the input).  The pad is never read. */
int b2cnn_forward_pitched(b2cnn_handle *h, const void *x, int dtype, int64_t B, int64_t x_pitch, const float *age,
                          int64_t n_age, int mode, int apply_sigmoid, float *out, void *workspace,
                          int64_t workspace_bytes, void *stream);

/* Same call with HOST pointers (ideally pinned): chunked H2D copy of x overlapped with
 * compute, D2H of the B results; synchronous on return.  Uses library-owned staging. */
int b2cnn_forward_host(b2cnn_handle *h, const void *x_host, int dtype, int64_t B,
                       const float *age_host, int64_t n_age, int mode, int apply_sigmoid,
                       float *out_host);

/* Intermediate of bin/models.py:29 (after the second pool, before the LSTM):
 * feats[B][L_out] floats on the device.  For parity tests. */
int b2cnn_features(b2cnn_handle *h, const void *x, int dtype, int64_t B, float *feats,
                   void *stream);

/* Options: "path" = B2CNN_PATH_*; "tc_splits" = 2|3: bf16 pieces per fp32 conv1 weight on the
 * tensor cores (3, default: exact fp32 weights; 2: weights rounded to 16 mantissa bits);
 * "stream_f32" = 0|1 (default 1): fp32 windows take the streaming kernel instead of the generic one;
 * "tc_fused" = 0|1 (default 1): bf16 windows take the fused conv+projection kernel; "small_kernel" = 0|1;
 * "profile" = 0|1: record CUDA events around the stages of each b2cnn_forward on its stream. */
int b2cnn_set_option(b2cnn_handle *h, const char *key, int64_t value);
int64_t b2cnn_get_option(b2cnn_handle *h, const char *key);

/* Kernel launches issued by the most recent forward on this handle (bench: gpu_launches),
 * and which path it took (B2CNN_PATH_GENERIC / B2CNN_PATH_TENSORCORE / B2CNN_PATH_STREAM). */
int64_t b2cnn_last_launch_count(b2cnn_handle *h);
int b2cnn_last_path(b2cnn_handle *h);

/* With option "profile"=1: device time in ms of a stage of the most recent b2cnn_forward
 * (0 = front end conv/pool kernel(s), the dominant kernel; 1 = projection + LSTM head).
 * Synchronises on the stage's end event.  <0 if unavailable. */
double b2cnn_last_stage_ms(b2cnn_handle *h, int stage);

/* ---- The two steps in front of the model call, on the device (SURVEY.md section 8, rows f2 + f1) ----
 * b2cnn_prep_windows replaces, for replay, what reaches bin/predictStream.py:105-139 through Kafka and
 * Spark: the 180 s / 5 s sliding mean with nulls skipped (bin/processStream.py:196-208), forward-fill,
 * back-fill and 0-fill of that grid (bin/processStream.py:62-123), and the 600 s / 60 s window assembly
 * x_arr[0, signal_index, :] with zeros for absent signals (bin/predictStream.py:105-139,245-259).
 *   raw        device pointer, int16 [n_samples][n_sig]: a WFDB format-16 numerics record as on disk
 *              (-32768 = missing; physical = (adc - baseline) / gain, what wfdb.rdrecord returns,
 *              bin/sendStream.py:46)
 *   sel        host array [n_sel]: record columns of the model's signals; position i becomes model
 *              channel i (the message index of bin/sendStream.py:59-64)
 *   gains / baselines  host arrays [n_sig]
 *   x_out      device pointer, [n_windows][n_channels][window_points] in `dtype` (f32 or bf16)
 *   t0_out     device pointer or NULL, [n_windows] window start times in seconds
 * The caller owns every buffer; the call allocates nothing and is asynchronous on `stream`. */
typedef struct b2cnn_prep_config {
    int32_t n_channels;    /* model input channels                 config.cfg CHANNEL_NAMES (10)      */
    int32_t window_points; /* points per window                    config.cfg WINDOWSIZE (120)        */
    int32_t grid_s;        /* slide of the smoothing window        processStream.py:199 (5 s)         */
    int32_t smooth_s;      /* length of the smoothing window       processStream.py:199 (180 s)       */
    int32_t stride_s;      /* slide of the model window            predictStream.py:252 (60 s)        */
} b2cnn_prep_config;
int64_t b2cnn_prep_window_count(int64_t n_samples, double fs, const b2cnn_prep_config *cfg);
int64_t b2cnn_prep_workspace_bytes(int64_t n_samples, double fs, int32_t n_sel, const b2cnn_prep_config *cfg);
int b2cnn_prep_windows(const int16_t *raw, int64_t n_samples, int32_t n_sig, const int32_t *sel, int32_t n_sel,
                       const double *gains, const double *baselines, double fs, const b2cnn_prep_config *cfg,
                       void *x_out, int dtype, double *t0_out, void *workspace, int64_t workspace_bytes, void *stream);

/* ---- The same two steps as a STREAM: per-patient device ring buffers ----
 * Replaces the per-trigger Python loop of bin/predictStream.py:70-156 (one row per patient, B = 1 each, numpy
 * assembly on the host) by device-resident state for P patients: every trigger appends the new samples of all
 * patients (b2cnn_ring_push), the grid points whose 180 s window is now complete are finalised and forward-filled
 * (bin/processStream.py:62-123,196-208), and the [P][n_channels][window_points] batch of the 600 s window that just
 * completed is written for ONE b2cnn_forward call.  Pushing a record trigger by trigger reproduces
 * b2cnn_prep_windows on the whole record bit-for-bit (tests/test_stream.py).
 *   n_sig        signals per sample frame (columns of the pushed arrays)
 *   fs           sampling rate shared by the ring's patients (1/60 Hz for MIMIC numerics)
 *   new_samples  DEVICE pointer [n_patients][n_new][n_sig]: B2CNN_SAMPLES_ADC16 = int16 ADC units as in a WFDB
 *                format-16 file (-32768 = missing; gain / baseline from b2cnn_ring_set_signals), or
 *                B2CNN_SAMPLES_F64 = physical values as fp64 (NaN = missing) -- what bin/sendStream.py:59-64 publishes
 *   emitted      host int: 1 when x_out was written (from the 10th trigger on), 0 while the first window fills
 * One push may carry at most stride_s seconds of samples.  The ring owns its device buffers; b2cnn_ring_push allocates
 * nothing and is asynchronous on `stream`. */
enum { B2CNN_SAMPLES_ADC16 = 0, B2CNN_SAMPLES_F64 = 1,
       B2CNN_SAMPLES_GRID = 2 /* fp64 5-second grid points [n_patients][n_new][n_sig] as bin/processStream.py:126-131 publishes
                                 them on `call-stream` (already smoothed and filled, 12 per trigger): appended as they are */ };
typedef struct b2cnn_ring b2cnn_ring;
int b2cnn_ring_create(const b2cnn_prep_config *cfg, int32_t n_patients, int32_t n_sig, double fs, int32_t device,
                      b2cnn_ring **out);
void b2cnn_ring_destroy(b2cnn_ring *ring);
int b2cnn_ring_reset(b2cnn_ring *ring, void *stream);
/* sel[n_sel]: frame columns of the model's signals for this patient (position i -> model channel i);
 * gains / baselines: host arrays [n_sig] or NULL (physical input). */
int b2cnn_ring_set_signals(b2cnn_ring *ring, int32_t patient, const int32_t *sel, int32_t n_sel, const double *gains,
                           const double *baselines, void *stream);
int b2cnn_ring_push(b2cnn_ring *ring, const void *new_samples, int sample_kind, int64_t n_new, void *x_out, int dtype,
                    int32_t *emitted, int64_t *window_index, double *t0_seconds, void *stream);

/* ---- The reference's wire formats, decoded on the device ----
 * A trigger's Kafka messages as one DEVICE byte buffer + offsets [n_msgs + 1] (message t = bytes[offsets[t] .. offsets[t+1])).
 * b2cnn_decode_sample_messages: value = json.dumps([i, val]) (bin/sendStream.py:62): idx_out[t] = i, val_out[t] = val
 *   (either may be NULL); with `frame` [frame_rows][n_sig] fp64 (first filled with NaN = missing) and row_of_msg[t]
 *   (the frame row = patient * n_new + sample the message belongs to, from its key / arrival order; < 0 = skip) the
 *   value is scattered to frame[row][i] -- the array b2cnn_ring_push(B2CNN_SAMPLES_F64) takes.
 * b2cnn_decode_array_messages: value = "[v0,v1,...]" (bin/processStream.py:128, read back at bin/predictStream.py:241):
 *   vals_out[t][0 .. max_vals) (NaN-padded), counts_out[t] = number of values (-1: malformed).
 * Numbers are converted with correct rounding (== json.loads / float() / Double.parseDouble) for up to 19 significant
 * digits and |decimal exponent| <= 27; NaN / Infinity / null are accepted; anything else counts in *n_bad (device int)
 * and yields NaN.  b2cnn_parse_decimal is the same parser compiled for the host (tests; status 0 ok, 1 malformed,
 * 2 out of range). */
int b2cnn_decode_sample_messages(const void *bytes, const int64_t *offsets, int64_t n_msgs, int32_t *idx_out, double *val_out,
                                 const int64_t *row_of_msg, double *frame, int64_t frame_rows, int32_t n_sig, int32_t *n_bad,
                                 void *stream);
int b2cnn_decode_array_messages(const void *bytes, const int64_t *offsets, int64_t n_msgs, int32_t max_vals, double *vals_out,
                                int32_t *counts_out, int32_t *n_bad, void *stream);
double b2cnn_parse_decimal(const char *s, int64_t len, int32_t *status);

/* The B200-native alternative to one JSON message per (sample, signal): ONE binary frame per trigger for all patients.
 *   header (32 bytes, little-endian)  |  int32 subject_id[n_patients]  |  pad to 8 bytes  |  samples[n_patients][n_new][n_sig]
 * `kind` = B2CNN_SAMPLES_ADC16 (int16), _F64 or _GRID (fp64): the payload is exactly the array b2cnn_ring_push takes, so
 * "decoding" is one H2D copy.  b2cnn_frame_check validates a HOST buffer and returns the byte offsets of the two arrays. */
#define B2CNN_FRAME_MAGIC 0x46573242u /* "B2WF" */
typedef struct b2cnn_frame_header {
    uint32_t magic;              /* B2CNN_FRAME_MAGIC */
    uint16_t version;            /* 1 */
    uint16_t kind;               /* B2CNN_SAMPLES_* */
    uint32_t n_patients, n_new, n_sig;
    uint32_t reserved;
    uint64_t first_index;        /* index of the frame's first sample (grid point) in the stream: gaps are detectable */
} b2cnn_frame_header;
int b2cnn_frame_check(const void *frame, int64_t bytes, b2cnn_frame_header *header_out, int64_t *ids_offset, int64_t *samples_offset);

const char *b2cnn_last_error(void);
const char *b2cnn_version(void);

/* ---- One training step on the device ----
 * Replaces the body of the reference's training loop (bin/utils.py:200-208):
 *     optimizer.zero_grad(); output = model(input, age); loss = criterion(output, target); loss.backward(); optimizer.step()
 * with criterion = nn.BCEWithLogitsLoss() (bin/utils.py:663) and torch.optim.Adam (bin/explore_torch.ipynb:3204-3205;
 * no amsgrad, no weight decay), for the model in train() mode: B2CNN_MODE_SEQUENCE is what model(input_batch, age)
 * computes (the LSTM scans the batch axis, bin/models.py:29-30), B2CNN_MODE_INDEPENDENT treats every window as its own
 * sequence.  All pointers are DEVICE pointers, everything is fp32:
 *   params          the packed blob of b2cnn_weight_count() floats (no affine), updated in place when apply_update != 0
 *   adam_m, adam_v  optimizer state, same size (zero before step 1); grads: same size, receives d loss / d params
 *   step            1-based count of optimizer steps (bias correction)
 *   x [B][C][W], age [B], target [B] (0 / 1)
 *   mask1 [B][4][P1], mask2 [B][L_out]: the two nn.Dropout(0.1) masks of bin/models.py:25,28 ALREADY scaled by 1/(1-p)
 *                   (torch's Philox stream cannot be reproduced here, so the caller draws them); NULL = no dropout
 *   loss_out        one float: the mean BCE-with-logits loss of this batch (before the update)
 * The call allocates nothing and is asynchronous on `stream`. */
typedef struct b2cnn_adam {
    float lr, beta1, beta2, eps;   /* torch defaults: 1e-3, 0.9, 0.999, 1e-8 */
} b2cnn_adam;
int64_t b2cnn_train_workspace_bytes(const b2cnn_config *cfg, int64_t B);
int b2cnn_train_step(const b2cnn_config *cfg, float *params, float *adam_m, float *adam_v, float *grads, int64_t step,
                     const b2cnn_adam *opt, int apply_update, const float *x, int64_t B, const float *age, const float *target,
                     int mode, const float *mask1, const float *mask2, float *loss_out, void *workspace, int64_t workspace_bytes,
                     void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B2CNN_H_ */
