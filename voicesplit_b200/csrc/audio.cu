// STFT front end and phase-preserving iSTFT back end on the device:
// reference utils/audio_processor.py:469-496,537-547 (wav2spec / spec2wav with the mixture phase),
// whose STFT arithmetic is librosa.stft / librosa.istft (n_fft 1200, hop 160, hann(400), center, reflect).
//
// Both transforms are GEMMs on the tcgen05 kernel of tc_gemm.cu (fp16 hi/lo operands, 3 passes):
//   STFT : the window is zero outside its win_length-sample support, so
//          X_t[k] = sum_{n < win} w[n] y_pad[t hop + lp + n] e^{-2 pi i k (lp + n) / n_fft},  lp = (n_fft - win) / 2
//          is  frames[T][win] x DFT[2 bins][win]^T.  The frames are never materialised: the A operand is a TMA
//          tensor map over the padded signal whose ROW STRIDE is the hop (overlapping rows).  The fused epilogue
//          turns (re, im) into the normalised dB magnitude and the unit phasor D / |D|.
//   iSTFT: frames[T][win] = spectrum[T][2 bins] x (w[n] c_k / n_fft) (cos, -sin)^T, then an overlap-add gather
//          divided by the window sum-square (each output sample sees at most ceil(win / hop) frames).
#include "tc.cuh"

namespace vs {

struct AudioState {
    int n_fft = 0, hop = 0, win = 0, bins = 0;
    float min_db = -100.f, ref_db = 20.f;
    elt16 *dft_hi = nullptr, *dft_lo = nullptr;     // [2 bins][win]           row 2k = w cos, 2k+1 = -w sin
    elt16 *idft_hi = nullptr, *idft_lo = nullptr;   // [win][ldk = 2 bins pad] col 2k = c_k w cos, 2k+1 = -c_k w sin (x n_fft: 1 / n_fft in the overlap-add)
    float* wsq = nullptr;                           // [win] window squared
    int ldk = 0;
};

__device__ __forceinline__ double hann_p(int n, int win) { return 0.5 - 0.5 * cos(2.0 * M_PI * n / win); }

__global__ void k_make_dft(int n_fft, int win, int bins, int ldk, elt16* dhi, elt16* dlo, elt16* ihi, elt16* ilo, float* wsq) {
    const int lp = (n_fft - win) / 2;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < win) wsq[i] = (float)(hann_p((int)i, win) * hann_p((int)i, win));
    const long long nd = (long long)2 * bins * win;
    if (i < nd) {   // forward matrix [2 bins][win]
        int n = (int)(i % win), r = (int)(i / win), k = r >> 1;
        double ang = 2.0 * M_PI * ((long long)k * (lp + n) % n_fft) / n_fft;
        double v = hann_p(n, win) * ((r & 1) ? -sin(ang) : cos(ang));
        split16<1>((float)v, dhi[i], dlo[i]);
    }
    const long long ni = (long long)win * ldk;
    if (i < ni) {   // inverse matrix [win][ldk]
        int c = (int)(i % ldk), n = (int)(i / ldk), k = c >> 1;
        double v = 0.0;
        if (c < 2 * bins) {
            double ang = 2.0 * M_PI * ((long long)k * (lp + n) % n_fft) / n_fft;
            double ck = (k == 0 || 2 * k == n_fft) ? 1.0 : 2.0;       // one-sided spectrum: interior bins count twice
            // stored x n_fft (1 / n_fft is applied by the overlap-add): keeps the fp16 `lo` halves out of the subnormal range
            v = hann_p(n, win) * ck * ((c & 1) ? -sin(ang) : cos(ang));
        }
        split16<1>((float)v, ihi[i], ilo[i]);
    }
}

// wav [B][L] fp32 -> reflect-padded (n_fft/2 each side) fp16 hi/lo planes [B][Lp], zero beyond L + n_fft
__global__ void k_wav_prep(const float* __restrict__ wav, elt16* __restrict__ hi, elt16* __restrict__ lo, int L, int Lp, int half, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int p = (int)(i % Lp);
    long long b = i / Lp;
    float v = 0.f;
    if (p < L + 2 * half) {
        int j = p - half;
        if (j < 0) j = -j;
        if (j >= L) j = 2 * (L - 1) - j;
        if (j >= 0 && j < L) v = wav[b * L + j];
    }
    elt16 h, l;
    split16<1>(v, h, l);
    hi[i] = h; lo[i] = l;
}

// normalised (masked) spectrogram + phasor -> complex spectrum rows [M][ldk] as fp16 hi/lo (utils/audio_processor.py:489,545-547)
__global__ void k_spec_to_complex(const float* __restrict__ spec, const float* __restrict__ phasor, elt16* __restrict__ hi, elt16* __restrict__ lo,
                                  int bins, int ldk, float min_db, float ref_db, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over [M][ldk / 2] bin slots
    if (i >= n) return;
    int k = (int)(i % (ldk / 2));
    long long m = i / (ldk / 2);
    float re = 0.f, im = 0.f;
    if (k < bins) {
        float s = fminf(fmaxf(spec[m * bins + k], 0.f), 1.f);
        float amp = exp10f(((s - 1.f) * -min_db + ref_db) * 0.05f);
        float2 ph = reinterpret_cast<const float2*>(phasor)[m * bins + k];
        re = amp * ph.x; im = amp * ph.y;
    }
    elt16 h0, l0, h1, l1;
    split16<1>(re, h0, l0); split16<1>(im, h1, l1);
    const size_t o = (size_t)m * ldk + 2 * k;
    hi[o] = h0; hi[o + 1] = h1; lo[o] = l0; lo[o + 1] = l1;
}

// overlap-add gather + window sum-square normalisation; output sample i of utterance b is padded position i + n_fft/2
__global__ void k_overlap_add(const float* __restrict__ frames /*[B*T][win]*/, const float* __restrict__ wsq, float* __restrict__ out,
                              int T, int win, int hop, int lp, int half, int Lout, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int s = (int)(i % Lout);
    long long b = i / Lout;
    int pos = s + half - lp;                       // position relative to the window support of frame 0
    int t1 = pos / hop, t0 = (pos - win + hop) / hop;
    if (pos - win + hop < 0) t0 = 0;
    if (t1 > T - 1) t1 = T - 1;
    float acc = 0.f, wss = 0.f;
    for (int t = t0; t <= t1; ++t) {
        int nidx = pos - t * hop;
        if (nidx >= 0 && nidx < win) { acc += frames[((size_t)b * T + t) * win + nidx]; wss += wsq[nidx]; }
    }
    acc *= 0.5f / (float)half;                     // the inverse-DFT operand is stored x n_fft (k_make_dft); half = n_fft / 2
    out[i] = wss > 1.17549435e-38f ? acc / wss : acc;
}

struct AudioWs {
    elt16 *y_hi, *y_lo, *c_hi, *c_lo;
    float* frames;
    size_t total;
    int Tp, Lp, T;
};
static AudioWs audio_carve(const AudioState* s, int B, int L, int T, void* base) {
    char* p = (char*)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += align_up(bytes, 1024); return r; };
    AudioWs w{};
    w.T = T;
    w.Tp = (L + s->n_fft + s->hop - 1) / s->hop;           // rows per utterance of the overlapping-row frame view
    w.Lp = w.Tp * s->hop;
    w.y_hi = (elt16*)take((size_t)B * w.Lp * 2 + 4096);    // slack: the last (ignored) rows read past the last utterance
    w.y_lo = (elt16*)take((size_t)B * w.Lp * 2 + 4096);
    w.c_hi = (elt16*)take((size_t)B * T * s->ldk * 2);
    w.c_lo = (elt16*)take((size_t)B * T * s->ldk * 2);
    w.frames = (float*)take((size_t)B * T * s->win * 4);
    w.total = off;
    return w;
}

// |STFT|^2 of wav [B][L] as bf16 hi/lo planes [B * T][ld16] (T = 1 + L / hop), for the mel front end of encoder.cu
size_t audio_stft_scratch_bytes(const vs_engine* e, int B, int L) {
    if (!e->audio) return 0;
    const AudioState* s = (const AudioState*)e->audio;
    const size_t Lp = (size_t)((L + s->n_fft + s->hop - 1) / s->hop) * s->hop;
    return 2 * align_up((size_t)B * Lp * 2 + 4096, 1024);
}
int audio_stft_power(vs_engine* e, const float* wav, elt16* pw_hi, elt16* pw_lo, int ld16, int B, int L, void* scratch, cudaStream_t st) {
    if (!e->audio) { set_error("call vs_audio_configure first"); return VS_ERR_STATE; }
    const AudioState* s = (const AudioState*)e->audio;
    if (L <= s->n_fft / 2) { set_error("signal shorter than n_fft / 2 cannot be reflect-padded"); return VS_ERR_INVALID; }
    const int Tp = (L + s->n_fft + s->hop - 1) / s->hop, Lp = Tp * s->hop, T = 1 + L / s->hop;
    elt16* y_hi = (elt16*)scratch;
    elt16* y_lo = (elt16*)((char*)scratch + align_up((size_t)B * Lp * 2 + 4096, 1024));
    {
        const long long n = (long long)B * Lp;
        k_wav_prep<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(wav, y_hi, y_lo, L, Lp, s->n_fft / 2, n);
        VS_LAUNCH(e, KID_CONVERT, st, cudaGetLastError());
    }
    GemmTcArgs a{};
    a.M = B * Tp; a.N = 2 * s->bins; a.K = s->win; a.lda = s->hop; a.ldw = s->win;
    a.out_hi = pw_hi; a.out_lo = pw_lo; a.ld16 = ld16; a.rows_per_utt = Tp; a.t_valid = T; a.n_bins = s->bins;
    const int lp = (s->n_fft - s->win) / 2;
    return launch_gemm_tc(e, GEPI_STFT_POWER, KID_HEAD, y_hi + lp, y_lo + lp, s->dft_hi, s->dft_lo, a, VS_PREC_FP16X3, st);
}
void audio_geometry(const vs_engine* e, int* n_fft, int* hop, int* win) {
    const AudioState* s = (const AudioState*)e->audio;
    *n_fft = s ? s->n_fft : 0; *hop = s ? s->hop : 0; *win = s ? s->win : 0;
}

}  // namespace vs

using namespace vs;

extern "C" {

int vs_audio_configure(vs_engine* e, const vs_audio_params* ap, void* stream) {
    if (!e || !ap) { set_error("null argument"); return VS_ERR_INVALID; }
    if (ap->n_fft / 2 + 1 != e->d.num_freq) { set_error("n_fft / 2 + 1 must equal num_freq"); return VS_ERR_INVALID; }
    if (ap->win_length > ap->n_fft || ap->win_length % 8 || ap->hop_length % 8 || ap->hop_length < 8 || (ap->n_fft - ap->win_length) % 2) {
        set_error("audio: win_length and hop_length must be multiples of 8 and n_fft - win_length even"); return VS_ERR_INVALID;
    }
    AudioState* s = (AudioState*)e->audio;
    if (!s) { s = new AudioState(); e->audio = s; }
    cudaFree(s->dft_hi); cudaFree(s->dft_lo); cudaFree(s->idft_hi); cudaFree(s->idft_lo); cudaFree(s->wsq);
    s->n_fft = ap->n_fft; s->hop = ap->hop_length; s->win = ap->win_length; s->bins = ap->n_fft / 2 + 1;
    s->min_db = ap->min_level_db; s->ref_db = ap->ref_level_db;
    s->ldk = (2 * s->bins + 7) / 8 * 8;
    const size_t nd = (size_t)2 * s->bins * s->win, ni = (size_t)s->win * s->ldk;
    VS_CUDA_TRY(cudaMalloc(&s->dft_hi, nd * 2)); VS_CUDA_TRY(cudaMalloc(&s->dft_lo, nd * 2));
    VS_CUDA_TRY(cudaMalloc(&s->idft_hi, ni * 2)); VS_CUDA_TRY(cudaMalloc(&s->idft_lo, ni * 2));
    VS_CUDA_TRY(cudaMalloc(&s->wsq, s->win * 4));
    const size_t n = nd > ni ? nd : ni;
    k_make_dft<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(s->n_fft, s->win, s->bins, s->ldk, s->dft_hi, s->dft_lo, s->idft_hi,
                                                                             s->idft_lo, s->wsq);
    VS_CUDA_TRY(cudaGetLastError());
    return VS_OK;
}

size_t vs_audio_workspace_bytes(const vs_engine* e, int32_t B, int32_t L) {
    if (!e || !e->audio || B < 1 || L < 1) return 0;
    const AudioState* s = (const AudioState*)e->audio;
    return audio_carve(s, B, L, 1 + L / s->hop, nullptr).total;
}

int vs_wav2spec(vs_engine* e, const float* wav, float* spec, float* phasor, int32_t B, int32_t L, void* workspace, size_t workspace_bytes,
                void* stream) {
    if (!e || !e->audio) { set_error("call vs_audio_configure first"); return VS_ERR_STATE; }
    if (!wav || !spec || !phasor || !workspace || B < 1) { set_error("bad argument"); return VS_ERR_INVALID; }
    const AudioState* s = (const AudioState*)e->audio;
    if (L <= s->n_fft / 2) { set_error("signal shorter than n_fft / 2 cannot be reflect-padded"); return VS_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const int T = 1 + L / s->hop;
    AudioWs w = audio_carve(s, B, L, T, workspace);
    if (workspace_bytes < w.total) { set_error("workspace too small"); return VS_ERR_STATE; }
    e->launches = 0;
    prof_begin(e, st);
    {
        const long long n = (long long)B * w.Lp;
        k_wav_prep<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(wav, w.y_hi, w.y_lo, L, w.Lp, s->n_fft / 2, n);
        VS_LAUNCH(e, KID_CONVERT, st, cudaGetLastError());
    }
    GemmTcArgs a{};
    a.M = B * w.Tp; a.N = 2 * s->bins; a.K = s->win; a.lda = s->hop; a.ldw = s->win;
    a.out32 = spec; a.phasor = phasor; a.rows_per_utt = w.Tp; a.t_valid = T; a.n_bins = s->bins; a.min_db = s->min_db; a.ref_db = s->ref_db;
    const int lp = (s->n_fft - s->win) / 2;     // the frames start at the window support
    return launch_gemm_tc(e, GEPI_STFT, KID_HEAD, w.y_hi + lp, w.y_lo + lp, s->dft_hi, s->dft_lo, a, VS_PREC_FP16X3, st);
}

int vs_spec2wav(vs_engine* e, const float* spec, const float* phasor, float* wav_out, int32_t B, int32_t T, void* workspace,
                size_t workspace_bytes, void* stream) {
    if (!e || !e->audio) { set_error("call vs_audio_configure first"); return VS_ERR_STATE; }
    if (!spec || !phasor || !wav_out || !workspace || B < 1 || T < 2) { set_error("bad argument"); return VS_ERR_INVALID; }
    const AudioState* s = (const AudioState*)e->audio;
    cudaStream_t st = (cudaStream_t)stream;
    const int Lout = s->hop * (T - 1);
    AudioWs w = audio_carve(s, B, Lout, T, workspace);
    if (workspace_bytes < w.total) { set_error("workspace too small"); return VS_ERR_STATE; }
    e->launches = 0;
    prof_begin(e, st);
    {
        const long long n = (long long)B * T * (s->ldk / 2);
        k_spec_to_complex<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(spec, phasor, w.c_hi, w.c_lo, s->bins, s->ldk, s->min_db, s->ref_db, n);
        VS_LAUNCH(e, KID_CONVERT, st, cudaGetLastError());
    }
    GemmTcArgs a{};
    a.M = B * T; a.N = s->win; a.K = 2 * s->bins; a.lda = s->ldk; a.ldw = s->ldk; a.out32 = w.frames; a.ld_out = s->win;
    int rc = launch_gemm_tc(e, GEPI_PLAIN, KID_HEAD, w.c_hi, w.c_lo, s->idft_hi, s->idft_lo, a, VS_PREC_FP16X3, st);
    if (rc != VS_OK) return rc;
    {
        const long long n = (long long)B * Lout;
        k_overlap_add<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w.frames, s->wsq, wav_out, T, s->win, s->hop, (s->n_fft - s->win) / 2, s->n_fft / 2,
                                                                  Lout, n);
        VS_LAUNCH(e, KID_CONVERT, st, cudaGetLastError());
    }
    return VS_OK;
}

}  // extern "C"

namespace vs {
void audio_free(vs_engine* e) {
    AudioState* s = (AudioState*)e->audio;
    if (!s) return;
    cudaFree(s->dft_hi); cudaFree(s->dft_lo); cudaFree(s->idft_hi); cudaFree(s->idft_lo); cudaFree(s->wsq);
    delete s;
    e->audio = nullptr;
}
}  // namespace vs
