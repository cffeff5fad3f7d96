/*
 * TEST INFRASTRUCTURE - NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, double accumulation, OpenMP over independent rows) of the
 * VoiceSplit / VoiceFilter mask-estimation forward pass, used only as the checker by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Nothing in
 * voicesplit_b200/ or models/ may call into this file.
 *
 * The arithmetic of the reference lives in PyTorch (pinned torch==1.0.1 in
 * /root/reference/requirements.txt:7; not vendored), so the functions below restate the
 * published semantics of the torch.nn layers at the reference's own call sites:
 *
 *   conv stack   /root/reference/models/voicesplit/model.py:15-54   (ZeroPad2d, Conv2d, BatchNorm2d eval, Mish)
 *                /root/reference/models/voicefilter/model.py:17-56  (same, ReLU)
 *   Mish         /root/reference/utils/generic_utils.py:395-399     (x * tanh(softplus(x)), softplus threshold 20)
 *   reshape/cat  /root/reference/models/voicesplit/model.py:72-81   (index c*F+f, d-vector tiled over T)
 *   BiLSTM       /root/reference/models/voicesplit/model.py:57-61,82 (1 layer, gates i,f,g,o, zero state)
 *   head         /root/reference/models/voicesplit/model.py:83-87   (relu, fc1, relu, fc2, sigmoid)
 *   mask apply   /root/reference/train.py:95
 *
 * Pinning: the reference has no golden vectors of its own, so this file is
 * pinned against outputs of the unmodified reference module imported from /root/reference
 * (tests/golden/make_golden.py writes tests/golden/ npz files; tests/test_oracle.py checks them).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#define VS_ACT_MISH 0
#define VS_ACT_RELU 1

typedef struct {
    int num_freq, emb_dim, lstm_dim, fc1_dim, fc2_dim, activation;
    /* conv layer l = 0..7: weight [Cout][Cin][kh][kw], bias [Cout], bn gamma/beta/mean/var [Cout] */
    const float *conv_w[8], *conv_b[8], *bn_g[8], *bn_b[8], *bn_m[8], *bn_v[8];
    /* lstm direction d = 0 (forward), 1 (reverse) */
    const float *w_ih[2], *w_hh[2], *b_ih[2], *b_hh[2];
    const float *fc1_w, *fc1_b, *fc2_w, *fc2_b;
} vs_oracle_params;

static const int L_CIN[8]  = {1, 64, 64, 64, 64, 64, 64, 64};
static const int L_COUT[8] = {64, 64, 64, 64, 64, 64, 64, 8};
static const int L_KH[8]   = {1, 7, 5, 5, 5, 5, 5, 1};
static const int L_KW[8]   = {7, 1, 5, 5, 5, 5, 5, 1};
static const int L_DIL[8]  = {1, 1, 1, 2, 4, 8, 16, 1};

static double act_fn(double x, int kind) {
    if (kind == VS_ACT_RELU) return x > 0.0 ? x : 0.0;
    /* F.softplus(beta=1, threshold=20): identity above the threshold */
    double sp = x > 20.0 ? x : log1p(exp(x));
    return x * tanh(sp);
}

/* One conv layer + eval BatchNorm (eps 1e-5) + activation on one utterance.
 * in [Cin][T][F] -> out [Cout][T][F]; "same" zero padding ((kh-1)/2*dil rows, (kw-1)/2 cols),
 * dilation on T only (model.py:26-48). */
static void conv_bn_act(const float* in, float* out, int T, int F, int l, const vs_oracle_params* p) {
    const int Cin = L_CIN[l], Cout = L_COUT[l], kh = L_KH[l], kw = L_KW[l], dil = L_DIL[l];
    const int pt = (kh - 1) / 2 * dil, pf = (kw - 1) / 2;
    const float* W = p->conv_w[l];
#pragma omp parallel
    {
        double* acc = (double*)malloc(sizeof(double) * (size_t)F);
#pragma omp for collapse(2) schedule(static)
        for (int co = 0; co < Cout; ++co) {
            for (int t = 0; t < T; ++t) {
                for (int f = 0; f < F; ++f) acc[f] = (double)p->conv_b[l][co];
                for (int ci = 0; ci < Cin; ++ci) {
                    for (int i = 0; i < kh; ++i) {
                        int ti = t + i * dil - pt;
                        if (ti < 0 || ti >= T) continue;
                        const float* row = in + ((size_t)ci * T + ti) * F;
                        for (int j = 0; j < kw; ++j) {
                            double w = (double)W[(((size_t)co * Cin + ci) * kh + i) * kw + j];
                            int sh = j - pf;            /* reads row[f + sh] */
                            int f0 = sh < 0 ? -sh : 0;
                            int f1 = sh > 0 ? F - sh : F;
                            for (int f = f0; f < f1; ++f) acc[f] += w * (double)row[f + sh];
                        }
                    }
                }
                double inv = (double)p->bn_g[l][co] / sqrt((double)p->bn_v[l][co] + 1e-5);
                double mu = (double)p->bn_m[l][co], be = (double)p->bn_b[l][co];
                float* o = out + ((size_t)co * T + t) * F;
                for (int f = 0; f < F; ++f) o[f] = (float)act_fn((acc[f] - mu) * inv + be, p->activation);
            }
        }
        free(acc);
    }
}

static double sigmoid_d(double x) { return 1.0 / (1.0 + exp(-x)); }

/* y[n] = b[n] + sum_k W[n][k] x[k] */
static void gemv(const float* W, const float* b, const float* x, int N, int K, double* y) {
    for (int n = 0; n < N; ++n) {
        double s = b ? (double)b[n] : 0.0;
        const float* w = W + (size_t)n * K;
        for (int k = 0; k < K; ++k) s += (double)w[k] * (double)x[k];
        y[n] = s;
    }
}

/*
 * Full forward for a batch.  x [B][T][F], emb [B][E] -> mask [B][T][F]; optional outputs:
 * masked [B][T][F] (= x * mask), conv_out [B][T][8F] (LSTM input without the d-vector),
 * lstm_out [B][T][2H], act_l [B][64][T][F] after conv layer `dump_layer` (0-based, <7) if dump != NULL.
 */
int vs_oracle_forward(const vs_oracle_params* p, const float* x, const float* emb, int B, int T,
                      float* mask, float* masked, float* conv_out, float* lstm_out,
                      int dump_layer, float* dump) {
    const int F = p->num_freq, E = p->emb_dim, H = p->lstm_dim, N1 = p->fc1_dim, N2 = p->fc2_dim;
    const int I = 8 * F + E;
    if (N2 != F) return -1;
    const size_t plane = (size_t)T * F;
    float* a = (float*)malloc(sizeof(float) * 64 * plane);
    float* b = (float*)malloc(sizeof(float) * 64 * plane);
    float* xin = (float*)malloc(sizeof(float) * (size_t)T * I);
    float* hout = (float*)malloc(sizeof(float) * (size_t)T * 2 * H);
    if (!a || !b || !xin || !hout) return -2;

    for (int u = 0; u < B; ++u) {
        /* x.unsqueeze(1): one input channel (model.py:68) */
        memcpy(a, x + (size_t)u * plane, sizeof(float) * plane);
        float *src = a, *dst = b;
        for (int l = 0; l < 8; ++l) {
            conv_bn_act(src, dst, T, F, l, p);
            if (dump && l == dump_layer)
                memcpy(dump + (size_t)u * L_COUT[l] * plane, dst, sizeof(float) * L_COUT[l] * plane);
            float* tmp = src; src = dst; dst = tmp;
        }
        /* src = [8][T][F]; transpose(1,2).view -> [T][8F] index c*F+f, then cat the d-vector (model.py:72-81) */
        for (int t = 0; t < T; ++t) {
            float* row = xin + (size_t)t * I;
            for (int c = 0; c < 8; ++c)
                memcpy(row + (size_t)c * F, src + ((size_t)c * T + t) * F, sizeof(float) * F);
            memcpy(row + 8 * F, emb + (size_t)u * E, sizeof(float) * E);
            if (conv_out) memcpy(conv_out + ((size_t)u * T + t) * 8 * F, row, sizeof(float) * 8 * F);
        }
        /* BiLSTM, gate order i,f,g,o, h0=c0=0; reverse output aligned to the input time index */
#pragma omp parallel for schedule(static) num_threads(2)
        for (int d = 0; d < 2; ++d) {
            double* gx = (double*)malloc(sizeof(double) * 4 * H);
            double* gh = (double*)malloc(sizeof(double) * 4 * H);
            float* h = (float*)calloc(H, sizeof(float));
            double* c = (double*)calloc(H, sizeof(double));
            for (int s = 0; s < T; ++s) {
                int t = d ? T - 1 - s : s;
                gemv(p->w_ih[d], p->b_ih[d], xin + (size_t)t * I, 4 * H, I, gx);
                gemv(p->w_hh[d], p->b_hh[d], h, 4 * H, H, gh);
                for (int j = 0; j < H; ++j) {
                    double ig = sigmoid_d(gx[j] + gh[j]);
                    double fg = sigmoid_d(gx[H + j] + gh[H + j]);
                    double gg = tanh(gx[2 * H + j] + gh[2 * H + j]);
                    double og = sigmoid_d(gx[3 * H + j] + gh[3 * H + j]);
                    c[j] = fg * c[j] + ig * gg;
                    double hv = og * tanh(c[j]);
                    hout[(size_t)t * 2 * H + (size_t)d * H + j] = (float)hv;
                }
                for (int j = 0; j < H; ++j) h[j] = hout[(size_t)t * 2 * H + (size_t)d * H + j];
            }
            free(gx); free(gh); free(h); free(c);
        }
        if (lstm_out) memcpy(lstm_out + (size_t)u * T * 2 * H, hout, sizeof(float) * (size_t)T * 2 * H);
        /* relu -> fc1 -> relu -> fc2 -> sigmoid (model.py:83-87), mask apply (train.py:95) */
#pragma omp parallel
        {
            float* r = (float*)malloc(sizeof(float) * 2 * H);
            double* y1 = (double*)malloc(sizeof(double) * N1);
            float* r1 = (float*)malloc(sizeof(float) * N1);
            double* y2 = (double*)malloc(sizeof(double) * N2);
#pragma omp for schedule(static)
            for (int t = 0; t < T; ++t) {
                for (int j = 0; j < 2 * H; ++j) { float v = hout[(size_t)t * 2 * H + j]; r[j] = v > 0.f ? v : 0.f; }
                gemv(p->fc1_w, p->fc1_b, r, N1, 2 * H, y1);
                for (int j = 0; j < N1; ++j) r1[j] = y1[j] > 0.0 ? (float)y1[j] : 0.f;
                gemv(p->fc2_w, p->fc2_b, r1, N2, N1, y2);
                for (int f = 0; f < F; ++f) {
                    float m = (float)sigmoid_d(y2[f]);
                    size_t o = ((size_t)u * T + t) * F + f;
                    mask[o] = m;
                    if (masked) masked[o] = x[o] * m;
                }
            }
            free(r); free(y1); free(r1); free(y2);
        }
    }
    free(a); free(b); free(xin); free(hout);
    return 0;
}

/* Stand-alone activation for unit tests of the fused epilogues. */
void vs_oracle_activation(const float* in, float* out, long n, int kind) {
    for (long i = 0; i < n; ++i) out[i] = (float)act_fn((double)in[i], kind);
}
