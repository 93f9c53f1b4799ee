// b2cnn_internal.cuh -- shared declarations of libb2cnn (sm_100a only).
//
// The hot path being replaced is MyCNN.forward (/root/reference/bin/models.py:22-36):
//   conv1+tanh (:23) -> pool (:24) -> [dropout = identity in eval (:25)] -> conv2+tanh (:26)
//   -> pool (:27) -> view(-1, MAGICNUM) (:29) -> 2-layer LSTM (:30) -> Linear (:31)
//   -> * relu(age*1e-8+1) (:32-33) -> squeeze (:34)
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b2cnn.h"

namespace b2cnn {

constexpr int kCMid = 4;      // conv1 out_channels (models.py:10)
constexpr int kHidden = 16;   // LSTM hidden (models.py:16)
constexpr int kGates = 64;    // 4 * hidden, PyTorch gate order i,f,g,o
constexpr int kMaxW1 = 800;   // 4 * C * K1 floats kept in the kernel-parameter constant bank
constexpr int kMaxK2 = 8;

// Everything derived from b2cnn_config (see RefArch in oracle/mycnn_torch.py for the mirror).
struct Dims {
    int C, K1, K2, PK, PS, W;
    int L1, P1, L2, L;       // conv1 out, pool1 out, conv2 out, pool2 out (= LSTM input size)
    int act, has_affine;
    float age_coef;
    int XP;                  // row pitch of x in elements (>= W): channel rows start XP apart, windows C*XP apart.
                             // == W for a contiguous [B][C][W] tensor; set per call (b2cnn_forward_pitched)
};

// Conv weights travel as a by-value kernel parameter: they land in the constant bank, so the
// fully unrolled FFMAs of the templated kernels read them as c[0x0][imm] operands.
struct ConvWeights {
    float w1[kMaxW1];          // [c][k][o]  -> (c*K1 + k)*4 + o
    float b1[kCMid];
    float w2[kCMid * kMaxK2];  // [c][k]     -> c*K2 + k
    float b2;
    float s1[kCMid], t1[kCMid], s2, t2;   // optional conv-epilogue affine (folded eval-BN)
};

// Small LSTM / head tensors: device pointers into the handle's copy of the packed blob.
struct HeadWeights {
    const float *wih0T;   // [L][64]  transposed copy of lstm.weight_ih_l0
    const float *whh0;    // [64][16]
    const float *bih0, *bhh0;
    const float *wih1, *whh1;   // [64][16]
    const float *bih1, *bhh1;
    const float *wo, *bo;       // out.weight[16], out.bias[1]
};

struct FrontParams {
    const void *x;        // [B][C][W] f32 or bf16
    float *feats;         // features, element (b, p) at feats[b*sB + p*sP]
    int64_t sB, sP;
    int B;
    int tile_p;           // final positions per tile
    int n_tiles;
    Dims d;
    int xs_stride, a1_stride;   // padded smem row strides (floats)
    const int *win_list;        // optional: indices of the windows to process (device memory)
    const int *win_count;       //           and how many (device memory)
    // gate mode (the exact re-computation behind the streaming tensor-core kernels): instead of writing feature
    // rows, every CTA multiplies the features of its tiles by W_ih_l0^T and writes gate_part[slice][b][64];
    // slice = blockIdx.x covers tiles [slice * tiles_per_slice, ...), slices >= gate_slices_used are zero-filled
    float *gate_part;           // [gate_slices][B][64] or nullptr
    const float *wih0T;         // [L][64]
    int gate_slices, tiles_per_slice;
    ConvWeights cw;
};

// NaN-propagating max, as ATen's max_pool1d ((v > m) || isnan(v)); fmaxf would drop NaNs.
__device__ __forceinline__ float max_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == B2CNN_ACT_TANH) return tanhf(v);
    if (act == B2CNN_ACT_RELU) return (v > 0.f || v != v) ? v : 0.f;
    return v;
}

__device__ __forceinline__ float sigmoid_acc(float v) { return 1.0f / (1.0f + expf(-v)); }

// conflict-free smem index for "thread r reads a window starting at stride*r": one pad word
// every 32 keeps lanes of a warp on distinct banks for strides 4 and 8.
__host__ __device__ __forceinline__ int padi(int g) { return g + (g >> 5); }

// ---- launchers (b2cnn_generic.cu / b2cnn_head.cu) ----------------------------------------
// Each returns the number of kernels it launched, or <0 with the message in `err`.
int launch_frontend_generic(const Dims &d, const ConvWeights &cw, const void *x, int dtype,
                            int64_t B, float *feats, int64_t sB, int64_t sP, cudaStream_t st,
                            int num_sms, const char **err);

int launch_frontend_generic_listed(const Dims &d, const ConvWeights &cw, const void *x, int dtype,
                                   int64_t B, float *feats, int64_t sB, int64_t sP, const int *win_list,
                                   const int *win_count, cudaStream_t st, int num_sms, const char **err);

// exact gate partials of the listed windows straight into the range-partial buffer the head sums (no feature rows)
int launch_frontend_generic_gates_listed(const Dims &d, const ConvWeights &cw, const void *x, int dtype, int64_t B,
                                         const float *wih0T, float *gate_part, int gate_slices, const int *win_list,
                                         const int *win_count, cudaStream_t st, int num_sms, const char **err);

int launch_head(const Dims &d, const HeadWeights &hw, const float *feats, int64_t sB, int64_t sP,
                int64_t B, const float *age, int64_t n_age, int mode, int apply_sigmoid,
                float *out, float *gates_ws, float *partial_ws, int ksplit, cudaStream_t st,
                const char **err);

int launch_reduce_gates(const float *partial, int slices, int64_t B, const HeadWeights &hw, float *gates,
                        cudaStream_t st, const char **err);
int launch_lstm_head(const Dims &d, const HeadWeights &hw, const float *gates, int64_t B, const float *age,
                     int64_t n_age, int mode, int apply_sigmoid, float *out, cudaStream_t st, const char **err);

int launch_reduce_lstm_head(const Dims &d, const HeadWeights &hw, const float *partial, int slices, int64_t B,
                            const float *age, int64_t n_age, int apply_sigmoid, float *out, cudaStream_t st, const char **err,
                            int *clean_count = nullptr, int *clean_flags = nullptr, const int *clean_list = nullptr);

int choose_ksplit(int64_t B, int L, int num_sms);

// b2cnn_small.cu: whole forward pass of short windows in one launch (independent windows only)
bool small_supported(const Dims &d);
int launch_small_forward(const Dims &d, const ConvWeights &cw, const HeadWeights &hw, const void *x, int dtype, int64_t B,
                         const float *age, int64_t n_age, int apply_sigmoid, float *out, cudaStream_t st, const char **err);

// b2cnn_batch.cu: many short windows per launch, one warp per window (the production shape [P, 10, 120])
bool batch_supported(const Dims &d);
int launch_short_batch(const Dims &d, const ConvWeights &cw, const HeadWeights &hw, const void *x, int dtype, int64_t B,
                       const float *age, int64_t n_age, int apply_sigmoid, float *out, int num_sms, cudaStream_t st, const char **err);

// b2cnn_prep.cu: preprocessing + window assembly in front of the model call (f2 + f1)
int64_t prep_window_count(int64_t n_samples, double fs, const b2cnn_prep_config *cfg);
int64_t prep_workspace_bytes(int64_t n_samples, double fs, int n_sel, const b2cnn_prep_config *cfg);
int prep_windows(const int16_t *raw, int64_t n_samples, int n_sig, const int *sel, int n_sel, const double *gains,
                 const double *baselines, double fs, const b2cnn_prep_config *cfg, void *x_out, int dtype, double *t0_out,
                 void *workspace, int64_t ws_bytes, cudaStream_t st, const char **err);

// streaming form of the same preprocessing: per-patient device ring buffers (b2cnn_prep.cu)
struct Ring;
int ring_create(const b2cnn_prep_config *cfg, int n_patients, int n_sig, double fs, int device, Ring **out, const char **err);
void ring_destroy(Ring *r);
int ring_device(const Ring *r);
int ring_reset(Ring *r, cudaStream_t st, const char **err);
int ring_set_signals(Ring *r, int patient, const int *sel, int n_sel, const double *gains, const double *baselines,
                     cudaStream_t st, const char **err);
int ring_push(Ring *r, const void *new_samples, int sample_kind, int64_t n_new, void *x_out, int dtype, int *emitted,
              int64_t *window_out, double *t0_out, cudaStream_t st, const char **err);

// b2cnn_wire.cu: the reference's JSON wire formats decoded on the device (row f3)
int wire_decode_pairs(const uint8_t *bytes, const int64_t *offsets, int64_t n_msgs, int *idx_out, double *val_out,
                      const int64_t *row_of_msg, double *frame, int64_t frame_rows, int n_sig, int *n_bad, cudaStream_t st,
                      const char **err);
int wire_decode_arrays(const uint8_t *bytes, const int64_t *offsets, int64_t n_msgs, int max_vals, double *vals_out, int *counts_out,
                       int *n_bad, cudaStream_t st, const char **err);
double wire_parse_decimal_host(const char *s, int64_t len, int *status);

// b2cnn_train.cu: one training step (row f4)
int64_t train_workspace_bytes(const b2cnn_config *cfg, int64_t B);
int train_step(const b2cnn_config *cfg, float *params, float *adam_m, float *adam_v, float *grads, int64_t step, float lr, float beta1,
               float beta2, float eps, int apply_update, const float *x, int64_t B, const float *age, const float *target, int sequence,
               const float *mask1, const float *mask2, float *loss_out, void *workspace, int64_t ws_bytes, cudaStream_t st,
               const char **err);

void launch_transpose_wih(const float *wih0, float *wih0T, int L, cudaStream_t st);

}  // namespace b2cnn
