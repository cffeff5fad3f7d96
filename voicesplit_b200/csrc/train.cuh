// Training path (fp32): declarations shared by train_kernels.cu and train.cu.
#pragma once
#include "common.cuh"

#define VS_ACT_NONE 2   // internal: raw conv output (pre-BatchNorm), used by the training forward and dgrad

namespace vs {

cudaError_t tr_bn_stats_plane(const float* z, double* sums, int F, int Fp, long long nrows, int num_sms, cudaStream_t st);
cudaError_t tr_bn_stats_cols(const float* z, double* sums, int C, int F, long long nrows, int num_sms, cudaStream_t st);
cudaError_t tr_bn_finalize(const double* sums, double count, const float* gamma, const float* beta, float* stat, float* rmean, float* rvar,
                           long long* nb, float momentum, int C, cudaStream_t st);
cudaError_t tr_bn_act_plane(int act, const float* z, float* a, const float* stat, int F, int Fp, long long npix, cudaStream_t st,
                            elt16* ahi = nullptr, elt16* alo = nullptr, int elt = 1);
cudaError_t tr_bn_act_cols(int act, const float* z, float* a, const float* stat, int C, int F, long long nrows, cudaStream_t st);
// SyncBN (vs_engine_set_sync_bn): all-reduce of the 128 per-channel double sums across ranks; count_scale = world size
struct BnSync {
    vs_stat_allreduce_fn fn = nullptr;
    void* user = nullptr;
    int world = 1;
};
inline BnSync bn_sync_of(const vs_engine* e) { BnSync s; s.fn = e->sync_fn; s.user = e->sync_user; s.world = e->sync_fn ? e->sync_world : 1; return s; }
// returns cudaSuccess, or cudaErrorUnknown after set_error() if the callback failed
cudaError_t tr_bn_sync(const BnSync& sync, double* sums, cudaStream_t st);
// BN + activation backward.  dgamma / dbeta (optional) receive this rank's sums BEFORE the SyncBN all-reduce: the parameter
// gradient is a per-rank quantity that the gradient all-reduce averages like every other gradient.
cudaError_t tr_bn_bwd_plane(int act, const float* da, const float* z, const float* stat, const float* gamma, double* sums, float* dz,
                            int F, int Fp, long long nrows, int num_sms, cudaStream_t st, elt16* dhi, elt16* dlo, float* dgamma, float* dbeta,
                            const BnSync& sync);
cudaError_t tr_bn_bwd_cols(int act, const float* da, const float* z, const float* stat, const float* gamma, double* sums, float* dz,
                           int C, int F, long long nrows, int num_sms, cudaStream_t st, float* dgamma, float* dbeta, const BnSync& sync);
// d loss / d x through cnn1: dx[row][f] = sum_c sum_j w[j][c] dz0[row][f - j + 3][c]   (dz0: fp32 plane [rows][Fp][64])
cudaError_t tr_front_dgrad(const float* dz0, const float* w /*[7][64]*/, float* dx, int F, int Fp, long long nrows, cudaStream_t st);
cudaError_t tr_conv_wgrad(const float* a, const float* dz, float* dwp, int T, int F, int Fp, int kh, int kw, int dil, long long nrows, cudaStream_t st);
cudaError_t tr_front_wgrad(const float* x, const float* dz, float* dwp, int F, int Fp, long long nrows, int num_sms, cudaStream_t st);
cudaError_t tr_point8_bwd(const float* a, const float* dz7, const float* w8p, float* da, float* dw8p, int F, int Fp, long long nrows, int num_sms, cudaStream_t st);
cudaError_t tr_unpack_conv_grad(const float* dwp, float* dw, int cout, int cin, int taps, cudaStream_t st);
cudaError_t tr_pack_conv_dgrad(const float* wp, float* wt, int kh, int kw, cudaStream_t st);
cudaError_t tr_gemm(const float* A, long long sai, long long sak, const float* B, long long sbk, long long sbj, float* C, long long ldc,
                    int I, int J, int K, bool accumulate, cudaStream_t st);
cudaError_t tr_colsum(const float* A, long long lda, int I, int J, float* out, cudaStream_t st);
cudaError_t tr_sigmoid_bwd(const float* g, const float* m, float* out, long long n, cudaStream_t st);
cudaError_t tr_relu_mask(float* g, const float* y, long long n, cudaStream_t st);
size_t tr_lstm_bwd_scratch_bytes(int H, int B);
cudaError_t tr_lstm_bwd(const vs_engine* e, float* gates, const float* cseq, const float* dhout, void* scratch, int B, int T, cudaStream_t st);

// raw-output variants of the fp32 forward kernels (fp32_kernels.cu)
cudaError_t launch_front_fp32_ex(const vs_engine* e, const float* x, float* plane, const float* w, const float* scale, const float* shift,
                                 int act, int B, int T, cudaStream_t st);
cudaError_t launch_conv_fp32_ex(const vs_engine* e, int layer, const float* in, float* out, const float* w, const float* scale,
                                const float* shift, int act, int B, int T, cudaStream_t st);
cudaError_t launch_point8_fp32_ex(const vs_engine* e, const float* plane, float* xcat, const float* w, const float* scale,
                                  const float* shift, int act, int B, int T, cudaStream_t st);

}  // namespace vs
