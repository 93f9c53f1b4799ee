/*
 * TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference forward pass.
 *
 * Follows /root/reference/bin/models.py:22-36 (MyCNN.forward) with the PyTorch operator
 * semantics the reference relies on (torch 2.11 CPU; third-party, not vendored by the
 * reference): valid cross-correlation Conv1d, tanh, MaxPool1d(floor, no padding,
 * NaN-propagating), eval-mode Dropout (identity), nn.LSTM (gate order i,f,g,o, separate
 * b_ih + b_hh, zero initial state), Linear, age scale relu(age*coef + 1).
 *
 * Two instantiations: *_f64 (double accumulation: the "ground truth" used to judge whether
 * the CUDA path or torch-CPU fp32 is closer) and *_f32 (float arithmetic).
 * Pinned against the reference's golden vector in tests/test_oracle.py.
 *
 * Never linked into, imported by, or executed from the product path.
 *
 * Packed weight blob order (floats), identical to the C-ABI's b2cnn_set_weights():
 *   conv1.weight[CM][C][K1], conv1.bias[CM], conv2.weight[1][CM][K2], conv2.bias[1],
 *   lstm.weight_ih_l0[4H][L], lstm.weight_hh_l0[4H][H], lstm.bias_ih_l0[4H], lstm.bias_hh_l0[4H],
 *   lstm.weight_ih_l1[4H][H], lstm.weight_hh_l1[4H][H], lstm.bias_ih_l1[4H], lstm.bias_hh_l1[4H],
 *   out.weight[H], out.bias[1]
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t in_channels, k1, c_mid, k2, pool_k, pool_s, hidden, window;
    int32_t act;        /* 0 tanh (reference), 1 relu, 2 identity */
    int32_t reserved;
    double age_coef;
} mycnn_ref_arch;

static int conv_len(int n, int k) { return n - k + 1; }
static int pool_len(int n, int pk, int ps) { return (n - pk) / ps + 1; }

int mycnn_ref_l_out(const mycnn_ref_arch *a) {
    int l1 = conv_len(a->window, a->k1);
    int p1 = pool_len(l1, a->pool_k, a->pool_s);
    int l2 = conv_len(p1, a->k2);
    return pool_len(l2, a->pool_k, a->pool_s);
}

int64_t mycnn_ref_weight_count(const mycnn_ref_arch *a) {
    int64_t L = mycnn_ref_l_out(a), H = a->hidden, G = 4 * H;
    return (int64_t)a->c_mid * a->in_channels * a->k1 + a->c_mid + (int64_t)a->c_mid * a->k2 + 1 +
           G * L + G * H + G + G + G * H + G * H + G + G + H + 1;
}

#define DEFINE_FORWARD(REAL, SUFFIX, TANH, EXP)                                                   \
    static REAL act_##SUFFIX(REAL v, int act) {                                                   \
        if (act == 0) return TANH(v);                                                             \
        if (act == 1) return (v > 0 || v != v) ? v : (REAL)0;                                     \
        return v;                                                                                 \
    }                                                                                             \
    static REAL sigmoid_##SUFFIX(REAL v) { return (REAL)1 / ((REAL)1 + EXP(-v)); }                \
    /* MaxPool1d: models.py:24,27.  (v > m) || isnan(v) keeps the first NaN like ATen. */         \
    static void pool_##SUFFIX(const REAL *in, int n, int pk, int ps, REAL *out) {                 \
        int m = pool_len(n, pk, ps);                                                              \
        for (int j = 0; j < m; ++j) {                                                             \
            REAL best = in[j * ps];                                                               \
            for (int k = 1; k < pk; ++k) {                                                        \
                REAL v = in[j * ps + k];                                                          \
                if ((v > best) || (v != v)) best = v;                                             \
            }                                                                                     \
            out[j] = best;                                                                        \
        }                                                                                         \
    }                                                                                             \
    /* one LSTM cell step for one layer: models.py:30 */                                          \
    static void lstm_cell_##SUFFIX(const float *w_ih, const float *w_hh, const float *b_ih,       \
                                   const float *b_hh, const REAL *x, int nx, int H, REAL *h,      \
                                   REAL *c) {                                                     \
        REAL g[256];                                                                              \
        for (int r = 0; r < 4 * H; ++r) {                                                         \
            REAL s = 0;                                                                           \
            for (int k = 0; k < nx; ++k) s += (REAL)w_ih[(int64_t)r * nx + k] * x[k];            \
            REAL s2 = 0;                                                                          \
            for (int k = 0; k < H; ++k) s2 += (REAL)w_hh[r * H + k] * h[k];                       \
            g[r] = (s + (REAL)b_ih[r]) + (s2 + (REAL)b_hh[r]);                                    \
        }                                                                                         \
        for (int u = 0; u < H; ++u) {                                                             \
            REAL ig = sigmoid_##SUFFIX(g[u]), fg = sigmoid_##SUFFIX(g[H + u]);                    \
            REAL gg = TANH(g[2 * H + u]), og = sigmoid_##SUFFIX(g[3 * H + u]);                    \
            c[u] = fg * c[u] + ig * gg;                                                           \
        }                                                                                         \
        for (int u = 0; u < H; ++u) h[u] = sigmoid_##SUFFIX(g[3 * H + u]) * TANH(c[u]);           \
    }                                                                                             \
    /* mode 0: every window starts from the zero state (predictStream.py:157, B=1 per call);  */  \
    /* mode 1: the LSTM scans the batch axis (models.py:29-30 with B>1).                      */  \
    int mycnn_ref_forward_##SUFFIX(const mycnn_ref_arch *a, const float *blob, const float *x,    \
                                   const float *age, int64_t B, int mode, REAL *logits,           \
                                   REAL *feats_out) {                                             \
        const int C = a->in_channels, K1 = a->k1, CM = a->c_mid, K2 = a->k2, H = a->hidden;       \
        const int W = a->window, PK = a->pool_k, PS = a->pool_s;                                  \
        const int L1 = conv_len(W, K1), P1 = pool_len(L1, PK, PS), L2 = conv_len(P1, K2);         \
        const int L = pool_len(L2, PK, PS), G = 4 * H;                                            \
        if (L1 < PK || L2 < PK || L < 1 || H > 64) return 1;                                      \
        const float *w1 = blob, *b1 = w1 + (int64_t)CM * C * K1, *w2 = b1 + CM;                   \
        const float *b2 = w2 + CM * K2, *wih0 = b2 + 1, *whh0 = wih0 + (int64_t)G * L;            \
        const float *bih0 = whh0 + G * H, *bhh0 = bih0 + G, *wih1 = bhh0 + G;                     \
        const float *whh1 = wih1 + G * H, *bih1 = whh1 + G * H, *bhh1 = bih1 + G;                 \
        const float *wo = bhh1 + G, *bo = wo + H;                                                 \
        REAL *c1 = (REAL *)malloc(sizeof(REAL) * (size_t)CM * L1);                                \
        REAL *q1 = (REAL *)malloc(sizeof(REAL) * (size_t)CM * P1);                                \
        REAL *c2 = (REAL *)malloc(sizeof(REAL) * (size_t)L2);                                     \
        REAL *f = (REAL *)malloc(sizeof(REAL) * (size_t)L);                                       \
        REAL h0[64], cc0[64], h1[64], cc1[64];                                                    \
        memset(h0, 0, sizeof h0); memset(cc0, 0, sizeof cc0);                                     \
        memset(h1, 0, sizeof h1); memset(cc1, 0, sizeof cc1);                                     \
        for (int64_t b = 0; b < B; ++b) {                                                         \
            const float *xb = x + b * (int64_t)C * W;                                             \
            for (int o = 0; o < CM; ++o)          /* conv1 + act: models.py:23 */                 \
                for (int t = 0; t < L1; ++t) {                                                    \
                    REAL s = 0;                                                                   \
                    for (int c = 0; c < C; ++c)                                                   \
                        for (int k = 0; k < K1; ++k)                                              \
                            s += (REAL)w1[((int64_t)o * C + c) * K1 + k] * (REAL)xb[(int64_t)c * W + t + k]; \
                    c1[(int64_t)o * L1 + t] = act_##SUFFIX(s + (REAL)b1[o], a->act);              \
                }                                                                                 \
            for (int o = 0; o < CM; ++o)          /* pool: models.py:24; dropout :25 = id */      \
                pool_##SUFFIX(c1 + (int64_t)o * L1, L1, PK, PS, q1 + (int64_t)o * P1);            \
            for (int t = 0; t < L2; ++t) {        /* conv2 + act: models.py:26 */                 \
                REAL s = 0;                                                                       \
                for (int c = 0; c < CM; ++c)                                                      \
                    for (int k = 0; k < K2; ++k) s += (REAL)w2[c * K2 + k] * q1[(int64_t)c * P1 + t + k]; \
                c2[t] = act_##SUFFIX(s + (REAL)b2[0], a->act);                                    \
            }                                                                                     \
            pool_##SUFFIX(c2, L2, PK, PS, f);     /* models.py:27; view :29 */                    \
            if (feats_out) memcpy(feats_out + b * (int64_t)L, f, sizeof(REAL) * (size_t)L);       \
            if (mode == 0) {                                                                      \
                memset(h0, 0, sizeof h0); memset(cc0, 0, sizeof cc0);                             \
                memset(h1, 0, sizeof h1); memset(cc1, 0, sizeof cc1);                             \
            }                                                                                     \
            lstm_cell_##SUFFIX(wih0, whh0, bih0, bhh0, f, L, H, h0, cc0);   /* models.py:30 */    \
            lstm_cell_##SUFFIX(wih1, whh1, bih1, bhh1, h0, H, H, h1, cc1);                        \
            REAL y = 0;                                                                           \
            for (int k = 0; k < H; ++k) y += (REAL)wo[k] * h1[k];           /* models.py:31 */    \
            y += (REAL)bo[0];                                                                     \
            REAL s = (REAL)age[b] * (REAL)a->age_coef + (REAL)1;            /* models.py:32 */    \
            if (!(s > 0) && s == s) s = 0;                                                        \
            logits[b] = y * s;                                              /* models.py:33 */    \
        }                                                                                         \
        free(c1); free(q1); free(c2); free(f);                                                    \
        return 0;                                                                                 \
    }

DEFINE_FORWARD(double, f64, tanh, exp)
DEFINE_FORWARD(float, f32, tanhf, expf)
