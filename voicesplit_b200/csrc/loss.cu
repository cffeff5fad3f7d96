// Training-loss chain on the device: the reference's differentiable iSTFT
// (openVoiceFilterAudioProcessor.torch_spec2wav, utils/audio_processor.py:498-509) and SiSNR_With_Pit
// (utils/generic_utils.py:403-474) as train.py:95-109 chains them, plus the analytic backward of both.
//
//   forward : spec, phase -> complex spectrum rows (fp16 hi/lo) -> frames = spectrum x synthesis matrix on the tcgen05
//             GEMM of tc_gemm.cu (estimate and target batched as one 2B-utterance GEMM) -> overlap-add / window envelope
//   Si-SNR  : one CTA per utterance; four passes over the two waveforms (L2 resident) with double block reductions:
//             means, <ze, zt> and energies, noise energy, then the closed-form d loss / d wav_est
//                 g_i = alpha zt_i + beta noise_i,   d/d est_i = m_i (g_i - sum_j m_j g_j / n)
//             written straight into the operand layout of the backward GEMM (divided by the window envelope)
//   backward: overlap-add^T is a strided gather, so d frames[t][n] = u[t hop + n] are OVERLAPPING ROWS of one signal u
//             (TMA row stride = hop, never materialised); d spectrum = frames(u) x synthesis^T on the same GEMM (bf16
//             hi/lo: gradients span many decades), whose epilogue chains (dRe, dIm) -> d magnitude -> d dB -> d spec.
#include "tc.cuh"

namespace vs {

struct LossState {
    int n_fft = 0, hop = 0, win = 0, bins = 0, ldk = 0, mode = 0;
    float min_db = -100.f, ref_db = 20.f;
    elt16 *syn_hi = nullptr, *syn_lo = nullptr;     // fp16 [win][ldk]      col 2k = c_k w cos / N, 2k+1 = -c_k w sin / N
    elt16 *synT_hi = nullptr, *synT_lo = nullptr;   // bf16 [2 bins][win]   the same matrix transposed (backward operand)
    float* wsq = nullptr;                           // [win] window squared
};

__device__ __forceinline__ double hann_any(int n, int win, int periodic) { return 0.5 - 0.5 * cos(2.0 * M_PI * n / (periodic ? win : win - 1)); }

__global__ void k_make_synthesis(int n_fft, int win, int bins, int ldk, int periodic, elt16* shi, elt16* slo, elt16* thi, elt16* tlo, float* wsq) {
    const int lp = (n_fft - win) / 2;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < win) wsq[i] = (float)(hann_any((int)i, win, periodic) * hann_any((int)i, win, periodic));
    if (i >= (long long)win * ldk) return;
    const int c = (int)(i % ldk), n = (int)(i / ldk), k = c >> 1;
    double v = 0.0;
    if (c < 2 * bins) {
        const double ang = 2.0 * M_PI * ((long long)k * (lp + n) % n_fft) / n_fft;
        const double ck = (k == 0 || 2 * k == n_fft) ? 1.0 : 2.0;       // one-sided spectrum: interior bins count twice
        v = hann_any(n, win, periodic) * ck / n_fft * ((c & 1) ? -sin(ang) : cos(ang));
    }
    // fp16 hi/lo of the forward operand: stored x n_fft (values up to 2 instead of 1.7e-3), so that the `lo` halves stay in
    // fp16's normal range (unscaled they are ~1e-6, subnormal: the split would carry ~15 bits, not 22); the overlap-add
    // multiplies by 1 / n_fft.  The transposed bf16 copy of the backward keeps the true values (bf16 has the fp32 exponent).
    split16<1>((float)(v * n_fft), shi[i], slo[i]);
    if (c < 2 * bins) split16<0>((float)v, thi[(size_t)c * win + n], tlo[(size_t)c * win + n]);
}

// utils/audio_processor.py:500-509: clamp -> denormalise -> 10^(S/20) -> complex with the given phase angle.
// Rows [0, M1) come from spec0, rows [M1, 2 M1) from spec1 (both use the same phase), written as fp16 hi/lo [rows][ldk].
__global__ void k_loss_complex(const float* __restrict__ spec0, const float* __restrict__ spec1, const float* __restrict__ phase,
                               elt16* __restrict__ hi, elt16* __restrict__ lo, long long M1, int bins, int ldk, float min_db, float ref_db, int q1,
                               long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over [rows][ldk / 2] bin slots
    if (i >= n) return;
    const int k = (int)(i % (ldk / 2));
    const long long m = i / (ldk / 2);
    float re = 0.f, im = 0.f;
    if (k < bins) {
        const long long ms = m < M1 ? m : m - M1;
        const float s = fminf(fmaxf((m < M1 ? spec0 : spec1)[ms * bins + k], 0.f), 1.f);
        const float amp = exp10f(((s - 1.f) * -min_db + ref_db) * 0.05f);
        float sn, cs;
        sincosf(phase[ms * bins + k], &sn, &cs);
        re = amp * (q1 ? expf(cs) : cs);
        im = amp * (q1 ? expf(sn) : sn);
    }
    elt16 h0, l0, h1, l1;
    split16<1>(re, h0, l0); split16<1>(im, h1, l1);
    const size_t o = (size_t)m * ldk + 2 * k;
    hi[o] = h0; hi[o + 1] = h1; lo[o] = l0; lo[o + 1] = l1;
}

// overlap-added squared window at output sample s (padded position s + n_fft / 2); at most ceil(win / hop) frames touch it
__device__ __forceinline__ void covering_frames(int s, int T, int win, int hop, int lp, int half, int& pos, int& t0, int& t1) {
    pos = s + half - lp;                           // relative to the window support of frame 0
    t1 = pos / hop;
    t0 = pos - win + hop < 0 ? 0 : (pos - win + hop) / hop;
    if (t1 > T - 1) t1 = T - 1;
}
__device__ __forceinline__ float envelope_at(int s, int T, int win, int hop, int lp, int half, const float* __restrict__ wsq) {
    int pos, t0, t1;
    covering_frames(s, T, win, hop, lp, half, pos, t0, t1);
    float wss = 0.f;
    for (int t = t0; t <= t1; ++t) {
        const int nidx = pos - t * hop;
        if (nidx >= 0 && nidx < win) wss += wsq[nidx];
    }
    return wss > 1e-11f ? wss : 1.f;               // torch.istft's envelope guard
}

__global__ void k_loss_overlap_add(const float* __restrict__ frames /*[rows][win]*/, const float* __restrict__ wsq, float* __restrict__ out, int T, int win,
                                   int hop, int lp, int half, int Lout, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = (int)(i % Lout);
    const long long b = i / Lout;
    int pos, t0, t1;
    covering_frames(s, T, win, hop, lp, half, pos, t0, t1);
    float acc = 0.f, wss = 0.f;
    for (int t = t0; t <= t1; ++t) {
        const int nidx = pos - t * hop;
        if (nidx >= 0 && nidx < win) { acc += frames[((size_t)b * T + t) * win + nidx]; wss += wsq[nidx]; }
    }
    acc *= 0.5f / (float)half;                     // the synthesis operand is stored x n_fft (k_make_synthesis); half = n_fft / 2
    out[i] = wss > 1e-11f ? acc / wss : acc;
}

// d loss / d wav -> the backward GEMM operand: u[b][p] = grad[b][p - half] / envelope, zero outside, bf16 hi/lo [B][Lp]
__global__ void k_grad_to_operand(const float* __restrict__ grad, const float* __restrict__ wsq, elt16* __restrict__ uhi, elt16* __restrict__ ulo, int T,
                                  int win, int hop, int lp, int half, int Lout, int Lp, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = (int)(i % Lp);
    const long long b = i / Lp;
    const int s = p - half;
    float v = 0.f;
    if (s >= 0 && s < Lout) v = grad[b * Lout + s] / envelope_at(s, T, win, hop, lp, half, wsq);
    split16<0>(v, uhi[i], ulo[i]);
}

template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* sh /*[32 * NV]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int q = 0; q < NV; ++q)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
    __syncthreads();                                // protects sh against the previous call's readers
    if (lane == 0)
#pragma unroll
        for (int q = 0; q < NV; ++q) sh[warp * NV + q] = v[q];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        double s = 0.0;
        for (int w = 0; w < nw; ++w) s += sh[w * NV + q];   // same order in every thread: deterministic, identical result
        v[q] = s;
    }
}

// SiSNR_With_Pit for one source per utterance (utils/generic_utils.py:421-473 with C = 1) and its gradient.
// est_wav / tgt_wav: [B][L].  One CTA per utterance.
__global__ void __launch_bounds__(1024) k_sisnr(const float* __restrict__ est_wav, const float* __restrict__ tgt_wav, const long long* __restrict__ lens, int B,
                                                int L, float* __restrict__ snr_out,
                                                const float* __restrict__ wsq, elt16* __restrict__ uhi, elt16* __restrict__ ulo, int T, int win, int hop,
                                                int lp, int half, int Lp) {
    __shared__ double sh[32 * 4];
    const int b = blockIdx.x;
    const float* est = est_wav + (size_t)b * L;
    const float* tgt = tgt_wav + (size_t)b * L;
    const long long len = lens[b];
    const int valid = len < 0 ? 0 : (len > L ? L : (int)len);       // get_mask: mask[i, :, len:] = 0
    const double n = (double)len;                                   // :431 num_samples is the raw length
    const double eps = 1e-16;
    double a2[2] = {0.0, 0.0};
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        if (i < valid) a2[0] += est[i];
        a2[1] += tgt[i];                                            // :432 the target mean sums the UNMASKED source
    }
    block_sum<2>(a2, sh);
    const double me = a2[0] / n, mt = a2[1] / n;
    double a4[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < valid; i += blockDim.x) {
        const double ze = est[i] - me, zt = tgt[i] - mt;
        a4[0] += ze * zt; a4[1] += zt * zt; a4[2] += ze; a4[3] += zt;
    }
    block_sum<4>(a4, sh);
    const double dot = a4[0], e0 = a4[1], sze = a4[2], szt = a4[3];
    const double en = e0 + eps, a = dot / en;                       // proj = a zt (:451-453)
    double a1[1] = {0.0};
    for (int i = threadIdx.x; i < valid; i += blockDim.x) {
        const double nz = (est[i] - me) - a * (tgt[i] - mt);
        a1[0] += nz * nz;
    }
    block_sum<1>(a1, sh);
    const double nn = a1[0] + eps, P = a * a * e0, r = P / nn;
    const double snr = 10.0 * log10(r + eps);                       // :457-458
    if (threadIdx.x == 0 && snr_out) snr_out[b] = (float)snr;
    if (!uhi) return;
    // loss = 20 - mean_b snr_b  =>  d loss / d ze = -(1/B) d snr / d ze
    const double kk = 10.0 / (2.302585092994046 * (r + eps));
    const double c = (dot - a * e0) / en;
    const double alpha = kk * (2.0 * a * e0 / (en * nn) + 2.0 * P * c / (nn * nn));
    const double beta = -kk * 2.0 * P / (nn * nn);
    const double gsum = alpha * szt + beta * (sze - a * szt);       // sum_j m_j g_j
    const double scale = -1.0 / B;
    for (int p = threadIdx.x; p < Lp; p += blockDim.x) {
        const int s = p - half;
        float v = 0.f;
        if (s >= 0 && s < valid) {
            const double zt = tgt[s] - mt, nz = (est[s] - me) - a * zt;
            v = (float)(scale * (alpha * zt + beta * nz - gsum / n)) / envelope_at(s, T, win, hop, lp, half, wsq);
        }
        split16<0>(v, uhi[(size_t)b * Lp + p], ulo[(size_t)b * Lp + p]);
    }
}

__global__ void k_loss_mean(const float* __restrict__ snr, int B, float* __restrict__ loss) {
    __shared__ double sh[32];
    double v[1] = {0.0};
    for (int i = threadIdx.x; i < B; i += blockDim.x) v[0] += snr[i];
    block_sum<1>(v, sh);
    if (threadIdx.x == 0) *loss = (float)(20.0 - v[0] / B);         // generic_utils.py:470-473, max_snr / C with C = 1
}

struct LossWs {
    elt16 *c_hi, *c_lo, *u_hi, *u_lo;
    float *frames, *wav, *snr;
    size_t total;
    int Tp, Lp, Lout;
};
static LossWs loss_carve(const LossState* s, int B, int T, void* base) {
    char* p = (char*)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += align_up(bytes, 1024); return r; };
    LossWs w{};
    w.Lout = s->hop * (T - 1);
    w.Tp = (w.Lout + s->n_fft + s->hop - 1) / s->hop;        // rows per utterance of the overlapping-row view of u
    w.Lp = w.Tp * s->hop;
    const size_t rows = (size_t)2 * B * T;                   // estimate + target
    w.c_hi = (elt16*)take(rows * s->ldk * 2);
    w.c_lo = (elt16*)take(rows * s->ldk * 2);
    w.frames = (float*)take(rows * s->win * 4);
    w.wav = (float*)take((size_t)2 * B * w.Lout * 4);
    w.u_hi = (elt16*)take((size_t)B * w.Lp * 2 + (size_t)s->n_fft * 4);   // slack: the last (ignored) rows read past the last utterance
    w.u_lo = (elt16*)take((size_t)B * w.Lp * 2 + (size_t)s->n_fft * 4);
    w.snr = (float*)take((size_t)B * 4);
    w.total = off;
    return w;
}

// spectrogram rows -> waveforms; nsrc = 1 (spec0 only) or 2 (spec0 then spec1)
static int loss_forward(vs_engine* e, const LossState* s, const LossWs& w, const float* spec0, const float* spec1, const float* phase, float* wav,
                        int B, int T, cudaStream_t st) {
    const int nsrc = spec1 ? 2 : 1;
    const long long M1 = (long long)B * T, rows = M1 * nsrc;
    {
        const long long n = rows * (s->ldk / 2);
        k_loss_complex<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(spec0, spec1, phase, w.c_hi, w.c_lo, M1, s->bins, s->ldk, s->min_db, s->ref_db,
                                                                   s->mode == VS_ISTFT_Q1, n);
        VS_LAUNCH(e, KID_TR_MISC, st, cudaGetLastError());
    }
    GemmTcArgs a{};
    a.M = (int)rows; a.N = s->win; a.K = 2 * s->bins; a.lda = s->ldk; a.ldw = s->ldk; a.out32 = w.frames; a.ld_out = s->win;
    int rc = launch_gemm_tc(e, GEPI_PLAIN, KID_TR_GEMM, w.c_hi, w.c_lo, s->syn_hi, s->syn_lo, a, VS_PREC_FP16X3, st);
    if (rc != VS_OK) return rc;
    {
        const long long n = (long long)nsrc * B * w.Lout;
        k_loss_overlap_add<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w.frames, s->wsq, wav, T, s->win, s->hop, (s->n_fft - s->win) / 2, s->n_fft / 2,
                                                                       w.Lout, n);
        VS_LAUNCH(e, KID_TR_MISC, st, cudaGetLastError());
    }
    return VS_OK;
}

// u (already in the workspace) -> d loss / d spec
static int loss_backward_gemm(vs_engine* e, const LossState* s, const LossWs& w, const float* spec, const float* phase, float* grad_spec, int B, int T,
                              cudaStream_t st) {
    GemmTcArgs a{};
    a.M = B * w.Tp; a.N = 2 * s->bins; a.K = s->win; a.lda = s->hop; a.ldw = s->win;
    a.out32 = grad_spec; a.rows_per_utt = w.Tp; a.t_valid = T; a.n_bins = s->bins; a.min_db = s->min_db; a.ref_db = s->ref_db;
    a.g_spec = spec; a.g_phase = phase; a.q1 = s->mode == VS_ISTFT_Q1;
    const int lp = (s->n_fft - s->win) / 2;                   // frame t starts at padded position t hop + lp
    return launch_gemm_tc(e, GEPI_ISTFT_BWD, KID_TR_GEMM, w.u_hi + lp, w.u_lo + lp, s->synT_hi, s->synT_lo, a, VS_PREC_BF16X3, st);
}

static const LossState* loss_state(vs_engine* e, int B, int T) {
    if (!e || !e->loss) { set_error("call vs_loss_configure first"); return nullptr; }
    if (!e->tc) { set_error("parameters must be loaded before the loss kernels run (tensor-core state)"); return nullptr; }
    if (B < 1 || T < 2) { set_error("loss: need B >= 1 and T >= 2 frames"); return nullptr; }
    return (const LossState*)e->loss;
}

void loss_free(vs_engine* e) {
    LossState* s = (LossState*)e->loss;
    if (!s) return;
    cudaFree(s->syn_hi); cudaFree(s->syn_lo); cudaFree(s->synT_hi); cudaFree(s->synT_lo); cudaFree(s->wsq);
    delete s;
    e->loss = nullptr;
}

}  // namespace vs

using namespace vs;

extern "C" {

int vs_loss_configure(vs_engine* e, const vs_loss_params* lp, void* stream) {
    if (!e || !lp) { set_error("null argument"); return VS_ERR_INVALID; }
    if (lp->n_fft / 2 + 1 != e->d.num_freq) { set_error("n_fft / 2 + 1 must equal num_freq"); return VS_ERR_INVALID; }
    if (lp->win_length > lp->n_fft || lp->win_length % 8 || lp->hop_length % 8 || lp->hop_length < 8 || (lp->n_fft - lp->win_length) % 2 ||
        lp->win_length < 16) {
        set_error("loss: win_length and hop_length must be multiples of 8 and n_fft - win_length even"); return VS_ERR_INVALID;
    }
    if (lp->phase_mode != VS_ISTFT_Q1 && lp->phase_mode != VS_ISTFT_CORRECTED) { set_error("unknown phase_mode"); return VS_ERR_INVALID; }
    loss_free(e);
    LossState* s = new LossState();
    e->loss = s;
    s->n_fft = lp->n_fft; s->hop = lp->hop_length; s->win = lp->win_length; s->bins = lp->n_fft / 2 + 1; s->mode = lp->phase_mode;
    s->min_db = lp->min_level_db; s->ref_db = lp->ref_level_db;
    s->ldk = (2 * s->bins + 7) / 8 * 8;
    const size_t ns = (size_t)s->win * s->ldk, nt = (size_t)2 * s->bins * s->win;
    VS_CUDA_TRY(cudaMalloc(&s->syn_hi, ns * 2)); VS_CUDA_TRY(cudaMalloc(&s->syn_lo, ns * 2));
    VS_CUDA_TRY(cudaMalloc(&s->synT_hi, nt * 2)); VS_CUDA_TRY(cudaMalloc(&s->synT_lo, nt * 2));
    VS_CUDA_TRY(cudaMalloc(&s->wsq, s->win * 4));
    // Q1-faithful: torch.hamming_window(win, periodic=False, alpha=.5, beta=.5) = symmetric Hann (audio_processor.py:509)
    k_make_synthesis<<<(unsigned)((ns + 255) / 256), 256, 0, (cudaStream_t)stream>>>(s->n_fft, s->win, s->bins, s->ldk, s->mode == VS_ISTFT_CORRECTED,
                                                                                    s->syn_hi, s->syn_lo, s->synT_hi, s->synT_lo, s->wsq);
    VS_CUDA_TRY(cudaGetLastError());
    return VS_OK;
}

size_t vs_loss_workspace_bytes(const vs_engine* e, int32_t B, int32_t T) {
    if (!e || !e->loss || B < 1 || T < 2) return 0;
    return loss_carve((const LossState*)e->loss, B, T, nullptr).total;
}

int vs_loss_spec2wav(vs_engine* e, const float* spec, const float* phase, float* wav_out, int32_t B, int32_t T, void* workspace, size_t workspace_bytes,
                     void* stream) {
    const LossState* s = loss_state(e, B, T);
    if (!s) return VS_ERR_STATE;
    if (!spec || !phase || !wav_out || !workspace) { set_error("bad argument"); return VS_ERR_INVALID; }
    LossWs w = loss_carve(s, B, T, workspace);
    if (workspace_bytes < w.total) { set_error("workspace too small"); return VS_ERR_STATE; }
    e->launches = 0;
    prof_begin(e, (cudaStream_t)stream);
    return loss_forward(e, s, w, spec, nullptr, phase, wav_out, B, T, (cudaStream_t)stream);
}

int vs_loss_spec2wav_backward(vs_engine* e, const float* spec, const float* phase, const float* grad_wav, float* grad_spec, int32_t B, int32_t T,
                              void* workspace, size_t workspace_bytes, void* stream) {
    const LossState* s = loss_state(e, B, T);
    if (!s) return VS_ERR_STATE;
    if (!spec || !phase || !grad_wav || !grad_spec || !workspace) { set_error("bad argument"); return VS_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    LossWs w = loss_carve(s, B, T, workspace);
    if (workspace_bytes < w.total) { set_error("workspace too small"); return VS_ERR_STATE; }
    e->launches = 0;
    prof_begin(e, st);
    {
        const long long n = (long long)B * w.Lp;
        k_grad_to_operand<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(grad_wav, s->wsq, w.u_hi, w.u_lo, T, s->win, s->hop, (s->n_fft - s->win) / 2,
                                                                      s->n_fft / 2, w.Lout, w.Lp, n);
        VS_LAUNCH(e, KID_TR_MISC, st, cudaGetLastError());
    }
    return loss_backward_gemm(e, s, w, spec, phase, grad_spec, B, T, st);
}

int vs_sisnr_loss(vs_engine* e, const float* est_spec, const float* target_spec, const float* phase, const int64_t* seq_len, float* loss_out,
                  float* snr_out, float* grad_est, int32_t B, int32_t T, void* workspace, size_t workspace_bytes, void* stream) {
    const LossState* s = loss_state(e, B, T);
    if (!s) return VS_ERR_STATE;
    if (!est_spec || !target_spec || !phase || !seq_len || !loss_out || !workspace) { set_error("bad argument"); return VS_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    LossWs w = loss_carve(s, B, T, workspace);
    if (workspace_bytes < w.total) { set_error("workspace too small"); return VS_ERR_STATE; }
    e->launches = 0;
    prof_begin(e, st);
    int rc = loss_forward(e, s, w, est_spec, target_spec, phase, w.wav, B, T, st);
    if (rc != VS_OK) return rc;
    float* snr = snr_out ? snr_out : w.snr;
    k_sisnr<<<B, 1024, 0, st>>>(w.wav, w.wav + (size_t)B * w.Lout, (const long long*)seq_len, B, w.Lout, snr, s->wsq, grad_est ? w.u_hi : nullptr, w.u_lo, T, s->win, s->hop,
                                (s->n_fft - s->win) / 2, s->n_fft / 2, w.Lp);
    VS_LAUNCH(e, KID_TR_MISC, st, cudaGetLastError());
    k_loss_mean<<<1, 256, 0, st>>>(snr, B, loss_out);
    VS_LAUNCH(e, KID_TR_MISC, st, cudaGetLastError());
    if (!grad_est) return VS_OK;
    return loss_backward_gemm(e, s, w, est_spec, phase, grad_est, B, T, st);
}


int vs_sisnr_wav(vs_engine* e, const float* est_wav, const float* target_wav, const int64_t* seq_len, float* loss_out, float* snr_out, int32_t B,
                 int32_t L, void* stream) {
    if (!e || !est_wav || !target_wav || !seq_len || !loss_out || !snr_out || B < 1 || L < 1) { set_error("bad argument"); return VS_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    e->launches = 0;
    prof_begin(e, st);
    k_sisnr<<<B, 1024, 0, st>>>(est_wav, target_wav, (const long long*)seq_len, B, L, snr_out, nullptr, nullptr, nullptr, 0, 0, 0, 0, 0, 0);
    VS_LAUNCH(e, KID_TR_MISC, st, cudaGetLastError());
    k_loss_mean<<<1, 256, 0, st>>>(snr_out, B, loss_out);
    VS_LAUNCH(e, KID_TR_MISC, st, cudaGetLastError());
    return VS_OK;
}

}  // extern "C"
