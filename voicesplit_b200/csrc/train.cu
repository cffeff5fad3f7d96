// Training path orchestration and C ABI: forward with batch-statistics BatchNorm that keeps what the
// backward needs, and the backward producing gradients for every parameter (and the d-vector).
// Arithmetic is fp32 on CUDA cores (the VS_PREC_FP32 kernels + train_kernels.cu); the dgrad of the conv
// layers reuses the forward conv kernel with flipped / transposed weights.
// Reference: autograd through models/voicesplit/model.py:66-89 as driven by train.py:94-111.
#include "train.cuh"
#include "tc.cuh"

namespace vs {

__global__ void k_sum0_to_float(const double* __restrict__ sums, float* out, int C) {
    int c = threadIdx.x;
    if (c < C && out) out[c] = (float)sums[c];
}
// hprev[m][d*H + j] = h of the forward step before (b,t) in direction d (zero at the sequence start)
__global__ void k_shift_h(const float* __restrict__ hout, float* __restrict__ hprev, int T, int H, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;   // n = M * 2H
    int col = (int)(i % (2 * H));
    long long m = i / (2 * H);
    int t = (int)(m % T), d = col / H;
    int tp = d ? t + 1 : t - 1;
    hprev[i] = (tp >= 0 && tp < T) ? hout[(m + (tp - t)) * 2 * H + col] : 0.f;
}
// dsum[b][n] = sum_t da[(b*T + t)][n]
__global__ void k_sum_over_t(const float* __restrict__ da, float* __restrict__ dsum, int T, int N, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;   // n = B * N
    int col = (int)(i % N);
    long long b = i / N;
    double s = 0.0;
    for (int t = 0; t < T; ++t) s += da[((size_t)b * T + t) * N + col];
    dsum[i] = (float)s;
}
__global__ void k_relu_copy(const float* __restrict__ in, float* __restrict__ out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = fmaxf(in[i], 0.f);
}
__global__ void k_fill(float* p, float v, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

struct TrainWs {
    float* z[7];            // pre-BatchNorm conv outputs, planes [B][T][Fp][64]
    float *P, *G1, *G2;     // activation / gradient scratch planes
    elt16 *Ahi, *Alo;       // a_l as fp16 hi/lo planes (operand of the tensor-core forward conv)
    elt16 *Dhi, *Dlo;       // dz_l as bf16 hi/lo planes (operand of the tensor-core data-gradient conv)
    float *z7, *xcat, *dxcat;           // [M][8F]
    float *gates, *bias_u, *hout, *cseq, *hprev, *dh;   // LSTM
    float *y1, *dy1, *dz2;              // head
    float* stat;            // [8][256]: mean, rstd, scale, shift per layer
    double* sums;           // [128]
    float* dwp;             // packed conv weight gradient [25*64*64]
    float* dsum;            // [B][8H]
    float* lstm_scratch;    // forward exchange + backward exchange (max)
    void* gemm_tc;          // 16-bit operand planes of the tensor-core GEMMs (tc_gemm.cu)
    unsigned int* barrier;
    size_t total;
};

static TrainWs train_carve(const vs_engine* e, int B, int T, void* base) {
    const int F = e->d.num_freq, H = e->d.lstm_dim, N1 = e->d.fc1_dim, Fp = padded_freq(F);
    char* p = (char*)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += align_up(bytes, 1024); return r; };
    const size_t M = (size_t)B * T, plane = M * Fp * 64 * sizeof(float);
    TrainWs w{};
    for (int l = 0; l < 7; ++l) w.z[l] = (float*)take(plane);
    w.P = (float*)take(plane); w.G1 = (float*)take(plane); w.G2 = (float*)take(plane);
    w.Ahi = (elt16*)take(plane / 2); w.Alo = (elt16*)take(plane / 2);
    w.Dhi = (elt16*)take(plane / 2); w.Dlo = (elt16*)take(plane / 2);
    w.z7 = (float*)take(M * 8 * F * 4); w.xcat = (float*)take(M * 8 * F * 4); w.dxcat = (float*)take(M * 8 * F * 4);
    w.gates = (float*)take(M * 8 * H * 4); w.bias_u = (float*)take((size_t)B * 8 * H * 4);
    w.hout = (float*)take(M * 2 * H * 4); w.cseq = (float*)take(M * 2 * H * 4); w.hprev = (float*)take(M * 2 * H * 4); w.dh = (float*)take(M * 2 * H * 4);
    w.y1 = (float*)take(M * N1 * 4); w.dy1 = (float*)take(M * N1 * 4); w.dz2 = (float*)take(M * F * 4);
    w.stat = (float*)take(8 * 256 * 4); w.sums = (double*)take(128 * 8);
    w.dwp = (float*)take((size_t)49 * 64 * 64 * 4);
    w.dsum = (float*)take((size_t)B * 8 * H * 4);
    size_t a = lstm_rec_scratch_bytes(e, B), b = tr_lstm_bwd_scratch_bytes(H, B);
    w.lstm_scratch = (float*)take(a > b ? a : b);
    w.barrier = (unsigned int*)take(256);
    w.gemm_tc = take(tc_train_gemm_workspace_bytes(e, B, T));
    w.total = off;
    return w;
}

#define TR(e, st, call) VS_LAUNCH(e, KID_TR_MISC, st, call)
#define TRK(e, id, st, call) VS_LAUNCH(e, id, st, call)

}  // namespace vs

using namespace vs;

extern "C" {

int vs_engine_set_sync_bn(vs_engine* e, vs_stat_allreduce_fn fn, void* user, int32_t world_size) {
    if (!e) { set_error("null engine"); return VS_ERR_INVALID; }
    if (fn && world_size < 1) { set_error("world_size must be >= 1"); return VS_ERR_INVALID; }
    e->sync_fn = fn; e->sync_user = user; e->sync_world = fn ? world_size : 1;
    return VS_OK;
}
int vs_engine_set_backward_hook(vs_engine* e, vs_backward_hook_fn fn, void* user) {
    if (!e) { set_error("null engine"); return VS_ERR_INVALID; }
    e->bwd_hook = fn; e->bwd_hook_user = user;
    return VS_OK;
}

int vs_engine_set_train_tensor_cores(vs_engine* e, int32_t enabled) {
    if (!e) { set_error("null engine"); return VS_ERR_INVALID; }
    e->train_tc = enabled != 0;
    return VS_OK;
}

size_t vs_train_workspace_bytes(const vs_engine* e, int32_t B, int32_t T) {
    if (!e || B < 1 || T < 1) return 0;
    return train_carve(e, B, T, nullptr).total;
}

int vs_train_forward(vs_engine* e, const vs_train_state* bn, const float* x, const float* emb, float* mask, int32_t B, int32_t T,
                     void* workspace, size_t workspace_bytes, void* stream) {
    if (!e || !e->loaded) { set_error("parameters not loaded"); return VS_ERR_STATE; }
    if (!x || !emb || !mask || !workspace || B < 1 || T < 1) { set_error("bad argument"); return VS_ERR_INVALID; }
    TrainWs w = train_carve(e, B, T, workspace);
    if (workspace_bytes < w.total) { set_error("workspace too small"); return VS_ERR_STATE; }
    cudaStream_t st = (cudaStream_t)stream;
    const int F = e->d.num_freq, H = e->d.lstm_dim, E = e->d.emb_dim, N1 = e->d.fc1_dim, Fp = padded_freq(F), act = e->d.activation;
    const long long M = (long long)B * T;
    const float mom = bn ? bn->momentum : 0.1f;
    const BnSync sync = bn_sync_of(e);
    e->launches = 0;
    prof_begin(e, st);
    // conv stack with batch statistics: z_l = conv(a_{l-1}) + bias, a_l = act(BN_batch(z_l))
    for (int l = 0; l < 7; ++l) {
        if (l == 0) {
            TRK(e, KID_TR_CONV_FWD, st, launch_front_fp32_ex(e, x, w.z[0], e->conv_w32[0], e->ones64, e->conv_bias[0], VS_ACT_NONE, B, T, st));
        } else if (e->train_tc) {
            int rc = tc_train_conv(e, l, false, w.Ahi, w.Alo, e->conv_bias[l], w.z[l], B, T, 1, KID_TR_CONV_FWD, st);
            if (rc != VS_OK) return rc;
        } else {
            TRK(e, KID_TR_CONV_FWD, st, launch_conv_fp32_ex(e, l, w.P, w.z[l], e->conv_w32[l], e->ones64, e->conv_bias[l], VS_ACT_NONE, B, T, st));
        }
        TRK(e, KID_TR_BN_STATS, st, tr_bn_stats_plane(w.z[l], w.sums, F, Fp, M, e->num_sms, st));
        if (sync.fn) TR(e, st, tr_bn_sync(sync, w.sums, st));
        TR(e, st, tr_bn_finalize(w.sums, (double)M * F * sync.world, e->bn_gamma[l], e->bn_beta[l], w.stat + l * 256, bn ? bn->running_mean[l] : nullptr,
                                 bn ? bn->running_var[l] : nullptr, bn ? (long long*)bn->num_batches_tracked[l] : nullptr, mom, 64, st));
        const bool tc_next = e->train_tc && l < 6;     // the next layer's conv reads the 16-bit planes
        TRK(e, KID_TR_BN_ACT, st, tr_bn_act_plane(act, w.z[l], tc_next ? nullptr : w.P, w.stat + l * 256, F, Fp, M * Fp, st,
                                                  tc_next ? w.Ahi : nullptr, tc_next ? w.Alo : nullptr, 1));
    }
    TR(e, st, launch_point8_fp32_ex(e, w.P, w.z7, e->conv_w32[7], e->ones64, e->conv_bias[7], VS_ACT_NONE, B, T, st));
    TR(e, st, tr_bn_stats_cols(w.z7, w.sums, 8, F, M, e->num_sms, st));
    if (sync.fn) TR(e, st, tr_bn_sync(sync, w.sums, st));
    TR(e, st, tr_bn_finalize(w.sums, (double)M * F * sync.world, e->bn_gamma[7], e->bn_beta[7], w.stat + 7 * 256, bn ? bn->running_mean[7] : nullptr,
                             bn ? bn->running_var[7] : nullptr, bn ? (long long*)bn->num_batches_tracked[7] : nullptr, mom, 8, st));
    TR(e, st, tr_bn_act_cols(act, w.z7, w.xcat, w.stat + 7 * 256, 8, F, M, st));
    // BiLSTM (gate activations and cell states are kept for the backward) and head
    VS_LAUNCH(e, KID_EMB_BIAS, st, launch_gemm_fp32(emb, E, e->wih_e, E, e->b_lstm, nullptr, 1, w.bias_u, 8 * H, B, 8 * H, E, false, EPI_NONE, nullptr, nullptr, st));
    if (e->train_tc && (H % 4) == 0) {
        int rc = tc_train_inproj(e, w.xcat, w.bias_u, w.gates, w.gemm_tc, B, T, st);
        if (rc != VS_OK) return rc;
    } else {
        VS_LAUNCH(e, KID_INPROJ, st, launch_gemm_fp32(w.xcat, 8 * F, e->wih_x, 8 * F, nullptr, w.bias_u, T, w.gates, 8 * H, (int)M, 8 * H, 8 * F, false, EPI_NONE, nullptr, nullptr, st));
    }
    VS_LAUNCH(e, KID_LSTM_REC, st, launch_lstm_rec_fp32(e, w.gates, w.hout, w.lstm_scratch, w.barrier, B, T, st, nullptr, nullptr, 0, w.gates, w.cseq));
    VS_LAUNCH(e, KID_FC1, st, launch_gemm_fp32(w.hout, 2 * H, e->fc1_w, 2 * H, e->fc1_b, nullptr, 1, w.y1, N1, (int)M, N1, 2 * H, true, EPI_RELU, nullptr, nullptr, st));
    VS_LAUNCH(e, KID_FC2, st, launch_gemm_fp32(w.y1, N1, e->fc2_w, N1, e->fc2_b, nullptr, 1, mask, F, (int)M, F, N1, false, EPI_SIGMOID_MASK, nullptr, nullptr, st));
    return VS_OK;
}

int vs_train_backward(vs_engine* e, const float* x, const float* emb, const float* mask, const float* grad_mask, const vs_grads* g,
                      float* grad_emb, float* grad_x, int32_t B, int32_t T, void* workspace, size_t workspace_bytes, void* stream) {
    if (!e || !e->loaded) { set_error("parameters not loaded"); return VS_ERR_STATE; }
    if (!x || !emb || !mask || !grad_mask || !g || !workspace) { set_error("bad argument"); return VS_ERR_INVALID; }
    TrainWs w = train_carve(e, B, T, workspace);
    if (workspace_bytes < w.total) { set_error("workspace too small"); return VS_ERR_STATE; }
    cudaStream_t st = (cudaStream_t)stream;
    const int F = e->d.num_freq, H = e->d.lstm_dim, E = e->d.emb_dim, N1 = e->d.fc1_dim, Fp = padded_freq(F), act = e->d.activation;
    const long long M = (long long)B * T;
    const int Mi = (int)M, KI = 8 * F + E;
    const BnSync sync = bn_sync_of(e);
    e->launches = 0;
    prof_begin(e, st);
    // ---- head: mask = sigmoid(z2), z2 = y1 W2^T + b2, y1 = relu(rh W1^T + b1), rh = relu(h)
    TR(e, st, tr_sigmoid_bwd(grad_mask, mask, w.dz2, M * F, st));
    TRK(e, KID_TR_GEMM, st, tr_gemm(w.dz2, 1, F, w.y1, N1, 1, g->fc2_w, N1, F, N1, Mi, false, st));               // dW2 [F][N1] = dz2^T y1
    TR(e, st, tr_colsum(w.dz2, F, Mi, F, g->fc2_b, st));
    TRK(e, KID_TR_GEMM, st, tr_gemm(w.dz2, F, 1, e->fc2_w, N1, 1, w.dy1, N1, Mi, N1, F, false, st));              // dy1 = dz2 W2
    TR(e, st, tr_relu_mask(w.dy1, w.y1, M * N1, st));
    {
        long long n = M * 2 * H;
        k_relu_copy<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w.hout, w.hprev, n);                 // rh in hprev (reused below)
        TR(e, st, cudaGetLastError());
    }
    TRK(e, KID_TR_GEMM, st, tr_gemm(w.dy1, 1, N1, w.hprev, 2 * H, 1, g->fc1_w, 2 * H, N1, 2 * H, Mi, false, st)); // dW1 [N1][2H] = dy1^T rh
    TR(e, st, tr_colsum(w.dy1, N1, Mi, N1, g->fc1_b, st));
    TRK(e, KID_TR_GEMM, st, tr_gemm(w.dy1, N1, 1, e->fc1_w, 2 * H, 1, w.dh, 2 * H, Mi, 2 * H, N1, false, st));    // d rh = dy1 W1
    TR(e, st, tr_relu_mask(w.dh, w.hout, M * 2 * H, st));                                            // -> d lstm_out
    // ---- BiLSTM: recurrence backward turns the saved gate activations into pre-activation gradients (in place)
    VS_LAUNCH(e, KID_TR_LSTM_BWD, st, tr_lstm_bwd(e, w.gates, w.cseq, w.dh, w.lstm_scratch, B, T, st));
    {
        long long n = M * 2 * H;
        k_shift_h<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w.hout, w.hprev, T, H, n);
        TR(e, st, cudaGetLastError());
        long long nb = (long long)B * 8 * H;
        k_sum_over_t<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(w.gates, w.dsum, T, 8 * H, nb);
        TR(e, st, cudaGetLastError());
    }
    const bool tc_gemms = e->train_tc && (H % 4) == 0;    // 16-byte operand rows for TMA
    for (int d = 0; d < 2; ++d) {
        const float* da = w.gates + (size_t)d * 4 * H;     // [M][4H] view with row stride 8H
        TRK(e, KID_TR_GEMM, st, tr_gemm(da, 1, 8 * H, w.hprev + (size_t)d * H, 2 * H, 1, g->w_hh[d], H, 4 * H, H, Mi, false, st));     // dW_hh = da^T h_prev
        if (!tc_gemms) TRK(e, KID_TR_GEMM, st, tr_gemm(da, 1, 8 * H, w.xcat, 8 * F, 1, g->w_ih[d], KI, 4 * H, 8 * F, Mi, false, st));  // dW_ih[:, :8F] = da^T X
        TRK(e, KID_TR_GEMM, st, tr_gemm(w.dsum + (size_t)d * 4 * H, 1, 8 * H, emb, E, 1, g->w_ih[d] + 8 * F, KI, 4 * H, E, B, false, st)); // dW_ih[:, 8F:] = (sum_t da)^T emb
        TR(e, st, tr_colsum(da, 8 * H, Mi, 4 * H, g->b_ih[d], st));
        TR(e, st, cudaMemcpyAsync(g->b_hh[d], g->b_ih[d], sizeof(float) * 4 * H, cudaMemcpyDeviceToDevice, st));
    }
    if (grad_emb) TRK(e, KID_TR_GEMM, st, tr_gemm(w.dsum, 8 * H, 1, e->wih_e, E, 1, grad_emb, E, B, E, 8 * H, false, st));          // d emb = (sum_t da) W_ih[:, 8F:]
    if (tc_gemms) {   // d X = da W_ih[:, :8F] and dW_ih[:, :8F] = da^T X on the tensor-core GEMM (bf16x3)
        int rc = tc_train_lstm_input_grads(e, w.gates, w.xcat, w.dxcat, g->w_ih[0], g->w_ih[1], KI, w.gemm_tc, B, T, st);
        if (rc != VS_OK) return rc;
    } else {
        TRK(e, KID_TR_GEMM, st, tr_gemm(w.gates, 8 * H, 1, e->wih_x, 8 * F, 1, w.dxcat, 8 * F, Mi, 8 * F, 8 * H, false, st));       // d X = da W_ih[:, :8F]
    }
    // every LSTM / FC parameter gradient is enqueued: a data-parallel caller can start reducing them now (97 % of the bytes)
    if (e->bwd_hook && e->bwd_hook(e->bwd_hook_user, VS_BWD_STAGE_LSTM_FC_DONE, stream) != 0) {
        set_error("backward hook failed"); return VS_ERR_STATE;
    }
    // ---- cnn8 (64 -> 8, BN, act) backward
    TR(e, st, tr_bn_bwd_cols(act, w.dxcat, w.z7, w.stat + 7 * 256, e->bn_gamma[7], w.sums, w.dxcat, 8, F, M, e->num_sms, st,
                             g->bn_gamma[7], g->bn_beta[7], sync));   // dxcat <- dz7 (in place)
    TR(e, st, tr_bn_stats_cols(w.dxcat, w.sums, 8, F, M, e->num_sms, st));
    k_sum0_to_float<<<1, 64, 0, st>>>(w.sums, g->conv_b[7], 8);
    TR(e, st, cudaGetLastError());
    TR(e, st, tr_bn_act_plane(act, w.z[6], w.P, w.stat + 6 * 256, F, Fp, M * Fp, st));                // a_6
    TR(e, st, tr_point8_bwd(w.P, w.dxcat, e->conv_w32[7], w.G1, w.dwp, F, Fp, M, e->num_sms, st));  // G1 = d a_6, dwp = dW8 packed [64][8]
    TR(e, st, tr_unpack_conv_grad(w.dwp, g->conv_w[7], 8, 64, 1, st));
    // ---- cnn7 .. cnn1
    for (int l = 6; l >= 0; --l) {
        const ConvGeom cg = kConv[l];
        const bool tc_dgrad = e->train_tc && l >= 1;
        // tensor-core path: dz_l only as bf16 hi/lo planes (what wgrad and dgrad read); fp32 path: fp32 plane G2
        TRK(e, KID_TR_BN_BWD, st, tr_bn_bwd_plane(act, w.G1, w.z[l], w.stat + l * 256, e->bn_gamma[l], w.sums, tc_dgrad ? nullptr : w.G2, F, Fp, M,
                                                  e->num_sms, st, tc_dgrad ? w.Dhi : nullptr, tc_dgrad ? w.Dlo : nullptr, g->bn_gamma[l], g->bn_beta[l],
                                                  sync));
        if (tc_dgrad) {
            // a conv bias in front of a BatchNorm has an exactly zero gradient (sum dz = -gamma rstd S2 sum(xhat) / N, sum(xhat) = 0)
            TR(e, st, cudaMemsetAsync(g->conv_b[l], 0, 64 * sizeof(float), st));
        } else {
            TRK(e, KID_TR_BN_STATS, st, tr_bn_stats_plane(w.G2, w.sums, F, Fp, M, e->num_sms, st));
            k_sum0_to_float<<<1, 64, 0, st>>>(w.sums, g->conv_b[l], 64);
            TR(e, st, cudaGetLastError());
        }
        if (l == 0) {
            TR(e, st, tr_front_wgrad(x, w.G2, w.dwp, F, Fp, M, e->num_sms, st));
            TR(e, st, tr_unpack_conv_grad(w.dwp, g->conv_w[0], 64, 1, 7, st));
            if (grad_x) TRK(e, KID_TR_DGRAD, st, tr_front_dgrad(w.G2, e->conv_w32[0], grad_x, F, Fp, M, st));   // d loss / d spectrogram
            break;
        }
        if (e->train_tc) {   // a_{l-1} recomputed as bf16 hi/lo planes; weight gradient on tensor cores
            TRK(e, KID_TR_BN_ACT, st, tr_bn_act_plane(act, w.z[l - 1], nullptr, w.stat + (l - 1) * 256, F, Fp, M * Fp, st, w.Ahi, w.Alo, 0));
            int rc = tc_train_wgrad(e, l, w.Ahi, w.Alo, w.Dhi, w.Dlo, w.dwp, g->conv_w[l], B, T, KID_TR_WGRAD, st);
            if (rc != VS_OK) return rc;
        } else {
            TRK(e, KID_TR_BN_ACT, st, tr_bn_act_plane(act, w.z[l - 1], w.P, w.stat + (l - 1) * 256, F, Fp, M * Fp, st));  // a_{l-1}
            TRK(e, KID_TR_WGRAD, st, tr_conv_wgrad(w.P, w.G2, w.dwp, T, F, Fp, cg.kh, cg.kw, cg.dil, M, st));
            TR(e, st, tr_unpack_conv_grad(w.dwp, g->conv_w[l], 64, 64, cg.kh * cg.kw, st));
        }
        if (e->train_tc) {   // G1 = d a_{l-1}: conv of dz_l with the flipped / transposed weights
            int rc = tc_train_conv(e, l, true, w.Dhi, w.Dlo, e->zeros64, w.G1, B, T, 0, KID_TR_DGRAD, st);
            if (rc != VS_OK) return rc;
        } else {
            TRK(e, KID_TR_DGRAD, st, launch_conv_fp32_ex(e, l, w.G2, w.G1, e->conv_wT32[l], e->ones64, e->zeros64, VS_ACT_NONE, B, T, st));
        }
    }
    return VS_OK;
}

}  // extern "C"

namespace vs {
// called from vs_engine_load_params (engine.cu): raw per-channel vectors and data-gradient weights
int train_pack(vs_engine* e, const vs_params* p, cudaStream_t st) {
    if (!e->ones64) {
        VS_CUDA_TRY(cudaMalloc(&e->ones64, 64 * sizeof(float)));
        VS_CUDA_TRY(cudaMalloc(&e->zeros64, 64 * sizeof(float)));
        for (int l = 0; l < 8; ++l) {
            VS_CUDA_TRY(cudaMalloc(&e->conv_bias[l], 64 * sizeof(float)));
            VS_CUDA_TRY(cudaMalloc(&e->bn_gamma[l], 64 * sizeof(float)));
            VS_CUDA_TRY(cudaMalloc(&e->bn_beta[l], 64 * sizeof(float)));
            if (l >= 1 && l <= 6) VS_CUDA_TRY(cudaMalloc(&e->conv_wT32[l], sizeof(float) * kConv[l].kh * kConv[l].kw * 64 * 64));
        }
        k_fill<<<1, 64, 0, st>>>(e->ones64, 1.f, 64);
        k_fill<<<1, 64, 0, st>>>(e->zeros64, 0.f, 64);
    }
    for (int l = 0; l < 8; ++l) {
        const int c = kConv[l].cout;
        VS_CUDA_TRY(cudaMemcpyAsync(e->conv_bias[l], p->conv_b[l], c * sizeof(float), cudaMemcpyDeviceToDevice, st));
        VS_CUDA_TRY(cudaMemcpyAsync(e->bn_gamma[l], p->bn_gamma[l], c * sizeof(float), cudaMemcpyDeviceToDevice, st));
        VS_CUDA_TRY(cudaMemcpyAsync(e->bn_beta[l], p->bn_beta[l], c * sizeof(float), cudaMemcpyDeviceToDevice, st));
        if (l >= 1 && l <= 6) {
            cudaError_t ce = tr_pack_conv_dgrad(e->conv_w32[l], e->conv_wT32[l], kConv[l].kh, kConv[l].kw, st);
            if (ce != cudaSuccess) { set_error(cudaGetErrorString(ce)); return VS_ERR_CUDA; }
        }
    }
    VS_CUDA_TRY(cudaGetLastError());
    return VS_OK;
}
void train_free(vs_engine* e) {
    cudaFree(e->ones64); cudaFree(e->zeros64);
    for (int l = 0; l < 8; ++l) { cudaFree(e->conv_bias[l]); cudaFree(e->bn_gamma[l]); cudaFree(e->bn_beta[l]); cudaFree(e->conv_wT32[l]); }
}
}  // namespace vs
