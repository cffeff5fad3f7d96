// On-device SDR for the evaluation driver: the number the reference's test.py reports,
// `bss_eval_sources(clean_wav, est_wav, False)[0][0]` (utils/generic_utils.py:511), for one source per utterance.
// mir_eval's algorithm (BSS Eval v3, 512-tap time-invariant filter):
//     s_filt  = projection of the zero-padded estimate onto span{ reference delayed by 0..511 samples }
//               -> normal equations  G c = d,  G = Toeplitz(r),  r[k] = sum ref[n] ref[n+k],  d[k] = sum ref[n] est[n+k]
//     SDR     = 10 log10( |s_filt|^2 / |est - s_filt|^2 ),  s_filt = c * ref  (length L + 511)
// Speech autocorrelation matrices are ill conditioned and mir_eval works in float64, so everything here accumulates in
// double: the correlations (k_sdr_xcorr: one thread per lag, signal chunks staged in shared memory, per-chunk partial sums
// reduced in a fixed order - deterministic), the Toeplitz solve (k_sdr_solve: Levinson recursion, one CTA per utterance)
// and the projection energies (k_sdr_energy).  HBM traffic is the two waveforms; the work is 3 x 512 x L double FMAs.
#include "common.cuh"

namespace vs {

constexpr int kFlen = 512;        // mir_eval.separation.bss_eval_sources default filter length
constexpr int kSdrChunk = 2048;   // samples per CTA

// partial[b][chunk][0][k] = sum_{n in chunk} ref[n] ref[n+k],  [1][k] = sum ref[n] est[n+k]
__global__ void __launch_bounds__(kFlen) k_sdr_xcorr(const float* __restrict__ ref, const float* __restrict__ est, int L, int nchunks,
                                                     double* __restrict__ partial) {
    __shared__ float s_ref[kSdrChunk + kFlen], s_est[kSdrChunk + kFlen];
    const int b = blockIdx.y, ch = blockIdx.x, n0 = ch * kSdrChunk, k = threadIdx.x;
    const float* r = ref + (size_t)b * L;
    const float* e = est + (size_t)b * L;
    for (int i = threadIdx.x; i < kSdrChunk + kFlen; i += blockDim.x) {
        const int n = n0 + i;
        s_ref[i] = n < L ? r[n] : 0.f;
        s_est[i] = n < L ? e[n] : 0.f;
    }
    __syncthreads();
    double ar = 0.0, ad = 0.0;
    const int cnt = min(kSdrChunk, L - n0);
#pragma unroll 4
    for (int i = 0; i < cnt; ++i) {
        const double x = (double)s_ref[i];
        ar = fma(x, (double)s_ref[i + k], ar);
        ad = fma(x, (double)s_est[i + k], ad);
    }
    double* dst = partial + ((size_t)b * nchunks + ch) * 2 * kFlen;
    dst[k] = ar;
    dst[kFlen + k] = ad;
}

__device__ __forceinline__ double block_sum_512(double v, double* sh /*[16]*/) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kFlen / 32; ++w) s += sh[w];
    return s;
}

// Levinson recursion for the symmetric positive definite Toeplitz system G c = d, one CTA (512 threads) per utterance
__global__ void __launch_bounds__(kFlen) k_sdr_solve(const double* __restrict__ partial, int nchunks, double* __restrict__ coef) {
    __shared__ double r[kFlen], d[kFlen], a[kFlen], an[kFlen], x[kFlen], sh[16];
    const int b = blockIdx.x, t = threadIdx.x;
    {
        double sr = 0.0, sd = 0.0;
        for (int ch = 0; ch < nchunks; ++ch) {          // fixed order: deterministic
            const double* src = partial + ((size_t)b * nchunks + ch) * 2 * kFlen;
            sr += src[t];
            sd += src[kFlen + t];
        }
        r[t] = sr; d[t] = sd; a[t] = t == 0 ? 1.0 : 0.0; x[t] = 0.0;
    }
    __syncthreads();
    double E = r[0];
    if (!(E > 0.0)) {                                   // silent reference: no projection
        coef[(size_t)b * kFlen + t] = 0.0;
        return;
    }
    if (t == 0) x[0] = d[0] / E;
    __syncthreads();
    for (int m = 1; m < kFlen; ++m) {
        // reflection coefficient: k = -(sum_{i<m} a[i] r[m-i]) / E      (a[0] = 1)
        const double acc = block_sum_512(t < m ? a[t] * r[m - t] : 0.0, sh);
        const double kf = -acc / E;
        if (t <= m) an[t] = (t < m ? a[t] : 0.0) + kf * (t >= 1 ? a[m - t] : 0.0);
        if (t == 0) an[0] = 1.0;
        E *= (1.0 - kf * kf);
        __syncthreads();
        if (t <= m) a[t] = an[t];
        __syncthreads();
        if (!(E > 0.0)) break;                          // numerically singular: keep the order-(m-1) solution
        // solution update: q = (d[m] - sum_{i<m} x[i] r[m-i]) / E;  x[i] += q a[m-i]
        const double acc2 = block_sum_512(t < m ? x[t] * r[m - t] : 0.0, sh);
        const double q = (d[m] - acc2) / E;
        if (t <= m) x[t] += q * a[m - t];
        __syncthreads();
    }
    coef[(size_t)b * kFlen + t] = x[t];
}

// energies of s_filt = c * ref and of est_pad - s_filt over output samples [n0, n0 + chunk) of L + 511
__global__ void __launch_bounds__(256) k_sdr_energy(const float* __restrict__ ref, const float* __restrict__ est, const double* __restrict__ coef, int L,
                                                    int nchunks, double* __restrict__ epart /*[B][nchunks][2]*/) {
    __shared__ double c[kFlen];
    __shared__ float s_ref[kSdrChunk + kFlen];
    __shared__ double sh[2][8];
    const int b = blockIdx.y, ch = blockIdx.x, n0 = ch * kSdrChunk;
    for (int i = threadIdx.x; i < kFlen; i += blockDim.x) c[i] = coef[(size_t)b * kFlen + i];
    const float* r = ref + (size_t)b * L;
    for (int i = threadIdx.x; i < kSdrChunk + kFlen; i += blockDim.x) {      // s_ref[i] = ref[n0 - (kFlen - 1) + i - ... ]
        const int n = n0 - kFlen + i;
        s_ref[i] = (n >= 0 && n < L) ? r[n] : 0.f;
    }
    __syncthreads();
    const int total = L + kFlen - 1;
    double es = 0.0, ee = 0.0;
    for (int j = threadIdx.x; j < kSdrChunk; j += blockDim.x) {
        const int n = n0 + j;
        if (n >= total) break;
        double s = 0.0;
#pragma unroll 4
        for (int k = 0; k < kFlen; ++k) s = fma(c[k], (double)s_ref[j + kFlen - k], s);   // ref[n - k]
        const double err = (n < L ? (double)est[(size_t)b * L + n] : 0.0) - s;
        es = fma(s, s, es);
        ee = fma(err, err, ee);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { es += __shfl_xor_sync(0xffffffffu, es, o); ee += __shfl_xor_sync(0xffffffffu, ee, o); }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = es; sh[1][threadIdx.x >> 5] = ee; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a0 = 0.0, a1 = 0.0;
        for (int w = 0; w < 8; ++w) { a0 += sh[0][w]; a1 += sh[1][w]; }
        epart[((size_t)b * nchunks + ch) * 2] = a0;
        epart[((size_t)b * nchunks + ch) * 2 + 1] = a1;
    }
}

__global__ void k_sdr_final(const double* __restrict__ epart, int nchunks, int B, float* __restrict__ sdr) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double s = 0.0, e = 0.0;
    for (int ch = 0; ch < nchunks; ++ch) { s += epart[((size_t)b * nchunks + ch) * 2]; e += epart[((size_t)b * nchunks + ch) * 2 + 1]; }
    sdr[b] = (float)(10.0 * log10(s / e));              // _bss_source_crit: inf for a perfect estimate, nan for silence, as mir_eval
}

struct SdrWs {
    double *partial, *coef, *epart;
    size_t total;
    int nch_x, nch_e;
};
static SdrWs sdr_carve(int B, int L, void* base) {
    char* p = (char*)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += align_up(bytes, 1024); return r; };
    SdrWs w{};
    w.nch_x = (L + kSdrChunk - 1) / kSdrChunk;
    w.nch_e = (L + kFlen - 1 + kSdrChunk - 1) / kSdrChunk;
    w.partial = (double*)take((size_t)B * w.nch_x * 2 * kFlen * 8);
    w.coef = (double*)take((size_t)B * kFlen * 8);
    w.epart = (double*)take((size_t)B * w.nch_e * 2 * 8);
    w.total = off;
    return w;
}

}  // namespace vs

using namespace vs;

extern "C" {

size_t vs_sdr_workspace_bytes(int32_t B, int32_t L) {
    if (B < 1 || L < 1) return 0;
    return sdr_carve(B, L, nullptr).total;
}

int vs_sdr(vs_engine* e, const float* ref_wav, const float* est_wav, float* sdr_out, int32_t B, int32_t L, void* workspace, size_t workspace_bytes,
           void* stream) {
    if (!e || !ref_wav || !est_wav || !sdr_out || !workspace || B < 1 || L < 1) { set_error("bad argument"); return VS_ERR_INVALID; }
    if (B > 65535) { set_error("vs_sdr: at most 65535 utterances per call"); return VS_ERR_INVALID; }
    SdrWs w = sdr_carve(B, L, workspace);
    if (workspace_bytes < w.total) { set_error("workspace too small"); return VS_ERR_STATE; }
    cudaStream_t st = (cudaStream_t)stream;
    e->launches = 0;
    prof_begin(e, st);
    k_sdr_xcorr<<<dim3(w.nch_x, B), kFlen, 0, st>>>(ref_wav, est_wav, L, w.nch_x, w.partial);
    VS_LAUNCH(e, KID_TR_MISC, st, cudaGetLastError());
    k_sdr_solve<<<B, kFlen, 0, st>>>(w.partial, w.nch_x, w.coef);
    VS_LAUNCH(e, KID_TR_MISC, st, cudaGetLastError());
    k_sdr_energy<<<dim3(w.nch_e, B), 256, 0, st>>>(ref_wav, est_wav, w.coef, L, w.nch_e, w.epart);
    VS_LAUNCH(e, KID_TR_MISC, st, cudaGetLastError());
    k_sdr_final<<<(B + 127) / 128, 128, 0, st>>>(w.epart, w.nch_e, B, sdr_out);
    VS_LAUNCH(e, KID_TR_MISC, st, cudaGetLastError());
    return VS_OK;
}

}  // extern "C"
