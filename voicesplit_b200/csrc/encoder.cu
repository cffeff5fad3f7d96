// GE2E speaker encoder - the producer of the d-vector input:
//   mel   = openVoiceFilterAudioProcessor.get_mel(wav)          utils/audio_processor.py:456-468
//   dvec  = SpeakerEncoder(mel)                                  notebooks/GE2E-Seungwonpark-...-openvoicefilter.py:63-85
// i.e. |STFT|^2 -> 40-band mel -> log10, 80-frame windows every 40 frames, 3 x LSTM(768) per window, last frame,
// Linear(768 -> 256), L2 normalise, mean over windows.
//
//   mel front end : the STFT GEMM of audio.cu with a power epilogue (bf16 hi/lo planes: the power spans many decades),
//                   then mel = power x basis^T on the same tcgen05 GEMM (bf16x3) with a log10 epilogue that writes the
//                   fp16 hi/lo operand planes of the first LSTM layer.
//   LSTM stack    : per layer  gates_x = X W_ih^T + b  as one GEMM (fp16x3), then the recurrence as a persistent
//                   warp-specialised tcgen05 kernel (k_lstm_uni_tc).  The windows overlap by half, but W_ih x_t does not
//                   depend on the window, so layer 0 projects every mel FRAME once and the recurrence indexes it by
//                   (utterance, window, step).
//   k_lstm_uni_tc : one CTA per (slice of 8 hidden units, group of 128 sequences); its 32 rows of W_hh (4 gates x 8 units,
//                   K = H, fp16 hi/lo) stay in shared memory for all steps; every step it TMA-loads h_{t-1} of its 128
//                   sequences in 64-wide K blocks, issues D[128 seq][32] = h W_slice^T (3 passes, fp32 in TMEM), adds the
//                   input projection, applies the gates in registers (one thread = one sequence x 8 units), publishes h_t
//                   (fp16 hi/lo) and meets the other slices of its group at a global-memory barrier.
#include "tc.cuh"
#include "sm100_ptx.cuh"

namespace vs {
using namespace ptx;

constexpr int kEU = 8;          // hidden units per CTA
constexpr int kEB = 128;        // sequences per CTA (MMA M)
constexpr int kEStages = 3;     // h K-blocks in flight
constexpr int kEncMaxLayers = 4;

struct EncoderState {
    vs_encoder_dims d{};
    int n_fft = 0, hop = 0, win = 0, bins = 0, ldp = 0;     // mel front end geometry (from the audio state)
    elt16 *mel_hi = nullptr, *mel_lo = nullptr;             // bf16 [num_mels][ldp] mel basis
    elt16 *wih_hi[kEncMaxLayers] = {}, *wih_lo[kEncMaxLayers] = {};   // fp16 [4H][Kp]
    elt16 *whh_hi[kEncMaxLayers] = {}, *whh_lo[kEncMaxLayers] = {};   // fp16 [nslices * 32][Hp], row = slice*32 + gate*8 + j
    float* bias[kEncMaxLayers] = {};                        // [4H] b_ih + b_hh
    elt16 *proj_hi = nullptr, *proj_lo = nullptr;           // fp16 [emb][Hp]
    float* proj_b = nullptr;
    bool loaded = false;
    int max_smem = 0;
};

// librosa.filters.mel defaults (Slaney scale, slaney area normalisation); double arithmetic, then bf16 hi/lo
__device__ __forceinline__ double mel_to_hz_d(double m) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}
__device__ __forceinline__ double hz_to_mel_d(double f) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
__global__ void k_make_mel_basis(int sr, int n_fft, int n_mels, int bins, int ldp, elt16* hi, elt16* lo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_mels * ldp) return;
    const int k = i % ldp, m = i / ldp;
    double v = 0.0;
    if (k < bins) {
        const double mmax = hz_to_mel_d(sr / 2.0), step = mmax / (n_mels + 1);
        const double f0 = mel_to_hz_d(step * m), f1 = mel_to_hz_d(step * (m + 1)), f2 = mel_to_hz_d(step * (m + 2));
        const double f = (sr / 2.0) * k / (bins - 1);
        const double lower = (f - f0) / (f1 - f0), upper = (f2 - f) / (f2 - f1);
        v = fmax(0.0, fmin(lower, upper)) * 2.0 / (f2 - f0);
    }
    split16<0>((float)v, hi[i], lo[i]);
}

// fp32 [rows][cols] -> fp16 hi/lo [rows][ldo] (zero padded columns)
__global__ void k_enc_split(const float* __restrict__ src, int rows, int cols, int ldo, elt16* __restrict__ hi, elt16* __restrict__ lo) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)rows * ldo) return;
    const int c = (int)(i % ldo);
    const long long r = i / ldo;
    split16<1>(c < cols ? src[r * cols + c] : 0.f, hi[i], lo[i]);
}
// W_hh [4H][H] -> [nslices * 32][Hp], row = slice*32 + gate*8 + j
__global__ void k_enc_pack_whh(const float* __restrict__ whh, int H, int Hp, int nslices, elt16* __restrict__ hi, elt16* __restrict__ lo) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)nslices * 32 * Hp) return;
    const int k = (int)(i % Hp);
    const long long r = i / Hp;
    const int row = (int)(r % 32), sl = (int)(r / 32), g = row / kEU, j = row % kEU, u = sl * kEU + j;
    split16<1>((u < H && k < H) ? whh[((size_t)g * H + u) * H + k] : 0.f, hi[i], lo[i]);
}
__global__ void k_enc_bias(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + b[i];
}

struct EncLstmArgs {
    int N, S, H, nslices, ngroups, group0, nkb, nk16;
    int Np, Hp;
    const float* gates_x;         // row(n, s) = (n / nwin) * utt_rows + (n % nwin) * win_rows + s, row stride 4H
    int nwin, utt_rows, win_rows;
    elt16 *hx_hi, *hx_lo;         // [2 parity][Np][Hp] exchange buffer
    elt16 *seq_hi, *seq_lo;       // [N * S][H] layer output (operand of the next layer's projection), or null
    elt16 *last_hi, *last_lo;     // [N][H] h of the final step, or null
    unsigned int* barrier;        // [ngroups_total]
};

__global__ void __launch_bounds__(192, 1) k_lstm_uni_tc(const EncLstmArgs a, const __grid_constant__ CUtensorMap tm_w_hi,
                                                        const __grid_constant__ CUtensorMap tm_w_lo, const __grid_constant__ CUtensorMap tm_h_hi,
                                                        const __grid_constant__ CUtensorMap tm_h_lo) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* w_smem = smem;                                            // [plane][kb][32 rows][128 B]
    uint8_t* a_ring = smem + (size_t)2 * a.nkb * 4096;                 // [stage][plane][128 rows][128 B]
    constexpr int stage_bytes = 2 * 16384;
    uint64_t* bars = reinterpret_cast<uint64_t*>(a_ring + (size_t)kEStages * stage_bytes);
    uint64_t* a_full = bars;
    uint64_t* a_empty = bars + kEStages;
    uint64_t* w_full = a_empty + kEStages;
    uint64_t* acc_full = w_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int slice = blockIdx.x % a.nslices;
    const int grp = a.group0 + blockIdx.x / a.nslices;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned int* counter = a.barrier + grp;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kEStages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        mbar_init(w_full, 1);
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 32);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // producer / issuer warps: the whole warp runs the warp-uniform loops, polls and waits; one elected lane issues the TMA /
    // tcgen05 instructions (an `if (lane == 0)` region wraps each of them in an elect-and-loop sequence of ~100 cycles)
    if (warp == 0) {
        // ---------------- producer: W slice once, then h K-blocks every step ----------------
        if (elect_one()) {
            mbar_arrive_expect_tx(w_full, (uint32_t)(2 * a.nkb * 4096));
            for (int p = 0; p < 2; ++p)
                for (int kb = 0; kb < a.nkb; ++kb)
                    tma_load_2d(w_smem + (size_t)(p * a.nkb + kb) * 4096, p == 0 ? &tm_w_hi : &tm_w_lo, w_full, kb * 64, slice * 32);
        }
        __syncwarp();
        int st = 0, ph = 0;
        for (int s = 1; s < a.S; ++s) {
            const unsigned int target = (unsigned int)s * a.nslices;     // every slice of the group has published h_{s-1}
            unsigned int spins = 0;
            for (;;) {      // acquire loads: no separate gpu-scope fence between the flag and the TMA issue (as in k_lstm_tc)
                unsigned int seen;
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
                if (seen >= target) break;
                if (++spins > (1u << 28)) __trap();
            }
            asm volatile("fence.proxy.async;" ::: "memory");             // generic-proxy flag read -> async-proxy (TMA) data reads
            const int row0 = ((s - 1) & 1) * a.Np + grp * kEB;
            for (int kb = 0; kb < a.nkb; ++kb) {
                mbar_wait(&a_empty[st], ph ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&a_full[st], (uint32_t)stage_bytes);
                    uint8_t* dst = a_ring + (size_t)st * stage_bytes;
                    tma_load_2d(dst, &tm_h_hi, &a_full[st], kb * 64, row0);
                    tma_load_2d(dst + 16384, &tm_h_lo, &a_full[st], kb * 64, row0);
                }
                __syncwarp();
                if (++st == kEStages) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        const uint32_t idesc = make_idesc_bf16(128, 32, 1);
        mbar_wait(w_full, 0);
        tc_fence_after();
        const uint32_t w_addr = smem_u32(w_smem);
        int st = 0, ph = 0;
        for (int s = 1; s < a.S; ++s) {
            uint32_t accumulate = 0;
            for (int kb = 0; kb < a.nkb; ++kb) {
                mbar_wait(&a_full[st], ph);
                tc_fence_after();
                const uint32_t h_hi = smem_u32(a_ring + (size_t)st * stage_bytes), h_lo = h_hi + 16384;
                const uint32_t w_hi = w_addr + (uint32_t)kb * 4096, w_lo = w_addr + (uint32_t)(a.nkb + kb) * 4096;
                if (elect_one()) {
                    const uint64_t d_hh = make_smem_desc(h_hi, 16, 1024, 2), d_hl = make_smem_desc(h_lo, 16, 1024, 2);
                    const uint64_t d_wh = make_smem_desc(w_hi, 16, 1024, 2), d_wl = make_smem_desc(w_lo, 16, 1024, 2);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (kb * 4 + k < a.nk16) {
                            umma_bf16(tmem, d_hh + 2 * k, d_wh + 2 * k, idesc, k == 0 ? accumulate : 1u);
                            umma_bf16(tmem, d_hl + 2 * k, d_wh + 2 * k, idesc, 1);
                            umma_bf16(tmem, d_hh + 2 * k, d_wl + 2 * k, idesc, 1);
                        }
                    }
                    umma_commit(&a_empty[st]);
                    if (kb == a.nkb - 1) umma_commit(acc_full);
                }
                __syncwarp();
                accumulate = 1;
                if (++st == kEStages) { st = 0; ph ^= 1; }
            }
        }
    } else {
        // ---------------- cell update: thread = one sequence, 8 units ----------------
        const int quad = warp & 3;
        const int n = grp * kEB + quad * 32 + lane;
        const bool valid = n < a.N;
        const int nv = valid ? n : 0;
        const int u0 = slice * kEU;
        float c[kEU];
#pragma unroll
        for (int j = 0; j < kEU; ++j) c[j] = 0.f;
        const uint32_t t_base = tmem + ((uint32_t)(quad * 32) << 16);
        const size_t gx_row0 = (size_t)(nv / a.nwin) * a.utt_rows + (size_t)(nv % a.nwin) * a.win_rows;
        for (int s = 0; s < a.S; ++s) {
            float gx[4][kEU];
            const float* gsrc = a.gates_x + (gx_row0 + s) * 4 * a.H + u0;
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
                for (int j = 0; j < kEU; j += 4) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(gsrc + (size_t)g * a.H + j));
                    gx[g][j] = v.x; gx[g][j + 1] = v.y; gx[g][j + 2] = v.z; gx[g][j + 3] = v.w;
                }
            if (s > 0) {
                mbar_wait(acc_full, (s - 1) & 1);
                tc_fence_after();
                uint32_t r[32];
                tmem_ld_32x32(t_base, r);
                tmem_ld_wait();
#pragma unroll
                for (int g = 0; g < 4; ++g)
#pragma unroll
                    for (int j = 0; j < kEU; ++j) gx[g][j] += __uint_as_float(r[g * kEU + j]);
                tc_fence_before();
            }
            __align__(16) elt16 vh[kEU], vl[kEU];
#pragma unroll
            for (int j = 0; j < kEU; ++j) {
                const float ig = sigmoid_fast(gx[0][j]), fg = sigmoid_fast(gx[1][j]);
                const float gg = tanh_fast(gx[2][j]), og = sigmoid_fast(gx[3][j]);
                c[j] = fmaf(fg, c[j], ig * gg);
                split16<1>(og * tanh_fast(c[j]), vh[j], vl[j]);
            }
            if (valid) {     // the exchange copy first: only it is on the critical path of the step
                const size_t xo = ((size_t)(s & 1) * a.Np + n) * a.Hp + u0;
                *reinterpret_cast<uint4*>(a.hx_hi + xo) = *reinterpret_cast<const uint4*>(vh);
                *reinterpret_cast<uint4*>(a.hx_lo + xo) = *reinterpret_cast<const uint4*>(vl);
            }
            if (s + 1 < a.S) {
                // publish h_s: CTA barrier of the 128 cell threads, then ONE gpu-scope release + counter bump
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (threadIdx.x == 64) {
                    __threadfence();
                    atomicAdd(counter, 1u);
                }
            }
            if (valid) {     // the next layer's operand / the projection's input drain behind the release
                if (a.seq_hi) {
                    const size_t so = ((size_t)n * a.S + s) * a.H + u0;
                    *reinterpret_cast<uint4*>(a.seq_hi + so) = *reinterpret_cast<const uint4*>(vh);
                    *reinterpret_cast<uint4*>(a.seq_lo + so) = *reinterpret_cast<const uint4*>(vl);
                }
                if (a.last_hi && s == a.S - 1) {
                    *reinterpret_cast<uint4*>(a.last_hi + (size_t)n * a.H + u0) = *reinterpret_cast<const uint4*>(vh);
                    *reinterpret_cast<uint4*>(a.last_lo + (size_t)n * a.H + u0) = *reinterpret_cast<const uint4*>(vl);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 32);
}

// proj output [nseq][emb] (+ bias) -> L2 normalise every window -> mean over the windows of an utterance (notebook :82-84)
__global__ void __launch_bounds__(256) k_dvector_finish(const float* __restrict__ proj, const float* __restrict__ bias, float* __restrict__ dvec, int nwin,
                                                        int emb) {
    __shared__ float sh[8];
    const int b = blockIdx.x;
    for (int j0 = 0; j0 < emb; j0 += blockDim.x) {      // emb <= blockDim.x in practice: one pass
        const int j = j0 + threadIdx.x;
        float acc = 0.f;
        for (int w = 0; w < nwin; ++w) {
            // the norm needs the whole vector: loop over emb in blockDim chunks
            float ss = 0.f;
            for (int q = threadIdx.x; q < emb; q += blockDim.x) {
                const float v = proj[((size_t)b * nwin + w) * emb + q] + bias[q];
                ss = fmaf(v, v, ss);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            __syncthreads();
            if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = ss;
            __syncthreads();
            float tot = 0.f;
            for (int q = 0; q < (int)(blockDim.x >> 5); ++q) tot += sh[q];
            if (j < emb) acc += (proj[((size_t)b * nwin + w) * emb + j] + bias[j]) / sqrtf(tot);
        }
        if (j < emb) dvec[(size_t)b * emb + j] = acc / nwin;
    }
}

struct EncWs {
    elt16 *pw_hi, *pw_lo;           // bf16 [B*T][ldp] power spectrum
    elt16 *x_hi, *x_lo;             // fp16 [B*T][melp] log-mel (layer-0 operand)
    float* gates;                   // [max(B*T, nseq*S)][4H]
    elt16 *seq_hi[2], *seq_lo[2];   // fp16 [nseq*S][H] ping-pong layer outputs
    elt16 *last_hi, *last_lo;       // fp16 [nseq][H]
    elt16 *hx_hi, *hx_lo;           // [2][Np][Hp]
    unsigned int* barrier;
    float* proj;                    // [nseq][emb]
    void* stft_scratch;
    size_t total;
    int nwin, nseq, Np, melp;
};
static EncWs enc_carve(const vs_engine* e, const EncoderState* s, int B, int T, int L, void* base) {
    char* p = (char*)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += align_up(bytes, 1024); return r; };
    EncWs w{};
    const int H = s->d.lstm_hidden, S = s->d.window;
    w.nwin = T >= S ? (T - S) / s->d.stride + 1 : 0;
    w.nseq = B * w.nwin;
    w.Np = (int)align_up((size_t)(w.nseq > 0 ? w.nseq : 1), kEB);
    w.melp = (s->d.num_mels + 7) / 8 * 8;
    const size_t rows0 = (size_t)B * T, rows1 = (size_t)w.nseq * S;
    w.pw_hi = (elt16*)take(rows0 * s->ldp * 2);
    w.pw_lo = (elt16*)take(rows0 * s->ldp * 2);
    w.x_hi = (elt16*)take(rows0 * w.melp * 2);
    w.x_lo = (elt16*)take(rows0 * w.melp * 2);
    w.gates = (float*)take((rows0 > rows1 ? rows0 : rows1) * 4 * H * 4);
    for (int i = 0; i < 2; ++i) { w.seq_hi[i] = (elt16*)take(rows1 * H * 2); w.seq_lo[i] = (elt16*)take(rows1 * H * 2); }
    w.last_hi = (elt16*)take((size_t)w.Np * H * 2);
    w.last_lo = (elt16*)take((size_t)w.Np * H * 2);
    w.hx_hi = (elt16*)take((size_t)2 * w.Np * H * 2);
    w.hx_lo = (elt16*)take((size_t)2 * w.Np * H * 2);
    w.barrier = (unsigned int*)take(4096);
    w.proj = (float*)take((size_t)w.Np * s->d.emb_dim * 4);
    w.stft_scratch = take(L > 0 ? audio_stft_scratch_bytes(e, B, L) : 0);
    w.total = off;
    return w;
}

static int enc_recurrence(vs_engine* e, const EncoderState* s, const EncWs& w, int layer, int utt_rows, int win_rows, elt16* seq_hi, elt16* seq_lo,
                          bool last, cudaStream_t st) {
    const int H = s->d.lstm_hidden, nslices = H / kEU;
    EncLstmArgs a{};
    a.N = w.nseq; a.S = s->d.window; a.H = H; a.nslices = nslices; a.nkb = (H + 63) / 64; a.nk16 = (H + 15) / 16;
    a.Np = w.Np; a.Hp = H; a.gates_x = w.gates; a.nwin = w.nwin; a.utt_rows = utt_rows; a.win_rows = win_rows;
    a.hx_hi = w.hx_hi; a.hx_lo = w.hx_lo; a.seq_hi = seq_hi; a.seq_lo = seq_lo;
    a.last_hi = last ? w.last_hi : nullptr; a.last_lo = last ? w.last_lo : nullptr; a.barrier = w.barrier;
    const int ngroups_total = w.Np / kEB;
    if (ngroups_total * (int)sizeof(unsigned int) > 4096) { set_error("encoder batch too large for the barrier table"); return VS_ERR_INVALID; }
    const int smem = 1024 + 2 * a.nkb * 4096 + kEStages * 2 * 16384 + 256;
    if (smem > s->max_smem) { set_error("lstm_hidden too large for the recurrent kernel"); return VS_ERR_UNSUPPORTED; }
    const int max_groups = e->num_sms / nslices;
    if (max_groups < 1) { set_error("lstm_hidden too large: all slices of a step must be co-resident"); return VS_ERR_UNSUPPORTED; }
    CUtensorMap tm_w_hi, tm_w_lo, tm_h_hi, tm_h_lo;
    {
        uint64_t wd[2] = {(uint64_t)H, (uint64_t)nslices * 32}, ws[1] = {(uint64_t)H * sizeof(elt16)};
        uint32_t wb[2] = {64, 32};
        uint64_t hd[2] = {(uint64_t)H, (uint64_t)2 * w.Np}, hs[1] = {(uint64_t)H * sizeof(elt16)};
        uint32_t hb[2] = {64, 128};
        bool ok = make_tmap_bf16(&tm_w_hi, s->whh_hi[layer], 2, wd, ws, wb, CU_TENSOR_MAP_SWIZZLE_128B);
        ok = ok && make_tmap_bf16(&tm_w_lo, s->whh_lo[layer], 2, wd, ws, wb, CU_TENSOR_MAP_SWIZZLE_128B);
        ok = ok && make_tmap_bf16(&tm_h_hi, w.hx_hi, 2, hd, hs, hb, CU_TENSOR_MAP_SWIZZLE_128B);
        ok = ok && make_tmap_bf16(&tm_h_lo, w.hx_lo, 2, hd, hs, hb, CU_TENSOR_MAP_SWIZZLE_128B);
        if (!ok) { set_error("cuTensorMapEncodeTiled failed (encoder lstm)"); return VS_ERR_CUDA; }
    }
    cudaError_t ce = cudaMemsetAsync(a.barrier, 0, 4096, st);
    if (ce != cudaSuccess) { set_error(cudaGetErrorString(ce)); return VS_ERR_CUDA; }
    ce = cudaFuncSetAttribute(k_lstm_uni_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (ce != cudaSuccess) { set_error(cudaGetErrorString(ce)); return VS_ERR_CUDA; }
    for (int g0 = 0; g0 < ngroups_total; g0 += max_groups) {     // groups are independent: as many as are co-resident per launch
        a.group0 = g0;
        a.ngroups = ngroups_total - g0 < max_groups ? ngroups_total - g0 : max_groups;
        void* args[] = {(void*)&a, (void*)&tm_w_hi, (void*)&tm_w_lo, (void*)&tm_h_hi, (void*)&tm_h_lo};
        ce = cudaLaunchCooperativeKernel((const void*)k_lstm_uni_tc, dim3(a.ngroups * nslices), dim3(192), args, (size_t)smem, st);
        if (ce != cudaSuccess) { set_error(std::string("k_lstm_uni_tc launch: ") + cudaGetErrorString(ce)); return VS_ERR_CUDA; }
        e->launches++;
    }
    if (e->profiling) prof_after(e, KID_LSTM_REC, st);
    return VS_OK;
}

// log-mel operand planes (w.x_hi/lo, [B*T][melp]) -> d-vectors
static int enc_stack(vs_engine* e, const EncoderState* s, const EncWs& w, float* dvec, int B, int T, cudaStream_t st) {
    const int H = s->d.lstm_hidden, S = s->d.window;
    for (int l = 0; l < s->d.lstm_layers; ++l) {
        GemmTcArgs a{};
        a.N = 4 * H; a.out32 = w.gates; a.ld_out = 4 * H; a.bias_group = s->bias[l];
        int rc;
        if (l == 0) {          // every mel frame once; the recurrence picks (utterance, window, step)
            a.M = B * T; a.K = s->d.num_mels; a.lda = w.melp; a.ldw = w.melp; a.group_rows = a.M;
            rc = launch_gemm_tc(e, GEPI_GATES, KID_INPROJ, w.x_hi, w.x_lo, s->wih_hi[0], s->wih_lo[0], a, VS_PREC_FP16X3, st);
        } else {
            a.M = w.nseq * S; a.K = H; a.lda = H; a.ldw = H; a.group_rows = a.M;
            rc = launch_gemm_tc(e, GEPI_GATES, KID_INPROJ, w.seq_hi[(l - 1) & 1], w.seq_lo[(l - 1) & 1], s->wih_hi[l], s->wih_lo[l], a, VS_PREC_FP16X3, st);
        }
        if (rc != VS_OK) return rc;
        const bool last = l == s->d.lstm_layers - 1;
        rc = enc_recurrence(e, s, w, l, l == 0 ? T : w.nwin * S, l == 0 ? s->d.stride : S, last ? nullptr : w.seq_hi[l & 1], last ? nullptr : w.seq_lo[l & 1],
                            last, st);
        if (rc != VS_OK) return rc;
    }
    GemmTcArgs a{};
    a.M = w.nseq; a.N = s->d.emb_dim; a.K = H; a.lda = H; a.ldw = H; a.out32 = w.proj; a.ld_out = s->d.emb_dim;
    int rc = launch_gemm_tc(e, GEPI_PLAIN, KID_HEAD, w.last_hi, w.last_lo, s->proj_hi, s->proj_lo, a, VS_PREC_FP16X3, st);
    if (rc != VS_OK) return rc;
    k_dvector_finish<<<B, 256, 0, st>>>(w.proj, s->proj_b, dvec, w.nwin, s->d.emb_dim);
    VS_LAUNCH(e, KID_HEAD, st, cudaGetLastError());
    return VS_OK;
}

// wav -> power planes -> log-mel planes (and optionally fp32 mel_out [B][T][num_mels])
static int enc_mel(vs_engine* e, const EncoderState* s, const EncWs& w, const float* wav, float* mel_out, int B, int L, int T, cudaStream_t st) {
    int rc = audio_stft_power(e, wav, w.pw_hi, w.pw_lo, s->ldp, B, L, w.stft_scratch, st);
    if (rc != VS_OK) return rc;
    GemmTcArgs a{};
    a.M = B * T; a.N = s->d.num_mels; a.K = s->bins; a.lda = s->ldp; a.ldw = s->ldp;
    a.out32 = mel_out; a.ld_out = s->d.num_mels; a.out_hi = w.x_hi; a.out_lo = w.x_lo; a.ld16 = w.melp;
    return launch_gemm_tc(e, GEPI_LOGMEL, KID_HEAD, w.pw_hi, w.pw_lo, s->mel_hi, s->mel_lo, a, VS_PREC_BF16X3, st);
}

static EncoderState* enc_state(vs_engine* e, bool need_params) {
    if (!e || !e->encoder) { set_error("call vs_encoder_configure first"); return nullptr; }
    EncoderState* s = (EncoderState*)e->encoder;
    if (need_params && !s->loaded) { set_error("call vs_encoder_load_params first"); return nullptr; }
    if (!e->tc) { set_error("the mask engine's parameters must be loaded before the encoder kernels run (tensor-core state)"); return nullptr; }
    return s;
}

void encoder_free(vs_engine* e) {
    EncoderState* s = (EncoderState*)e->encoder;
    if (!s) return;
    cudaFree(s->mel_hi); cudaFree(s->mel_lo); cudaFree(s->proj_hi); cudaFree(s->proj_lo); cudaFree(s->proj_b);
    for (int l = 0; l < kEncMaxLayers; ++l) {
        cudaFree(s->wih_hi[l]); cudaFree(s->wih_lo[l]); cudaFree(s->whh_hi[l]); cudaFree(s->whh_lo[l]); cudaFree(s->bias[l]);
    }
    delete s;
    e->encoder = nullptr;
}

}  // namespace vs

using namespace vs;

extern "C" {

int vs_encoder_configure(vs_engine* e, const vs_encoder_dims* d, void* stream) {
    if (!e || !d) { set_error("null argument"); return VS_ERR_INVALID; }
    int n_fft, hop, win;
    audio_geometry(e, &n_fft, &hop, &win);
    if (!n_fft) { set_error("call vs_audio_configure first (the mel front end shares its STFT)"); return VS_ERR_STATE; }
    if (d->lstm_layers < 1 || d->lstm_layers > kEncMaxLayers || d->lstm_hidden % 8 || d->lstm_hidden < 64 || d->num_mels < 1 || d->emb_dim < 1 ||
        d->window < 2 || d->stride < 1 || d->sample_rate < 1) {
        set_error("encoder: need 1..4 layers, lstm_hidden a multiple of 8 (>= 64), window >= 2, stride >= 1"); return VS_ERR_INVALID;
    }
    encoder_free(e);
    EncoderState* s = new EncoderState();
    e->encoder = s;
    s->d = *d; s->n_fft = n_fft; s->hop = hop; s->win = win; s->bins = n_fft / 2 + 1; s->ldp = (s->bins + 7) / 8 * 8;
    cudaDeviceGetAttribute(&s->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, e->device);
    const int n = d->num_mels * s->ldp;
    VS_CUDA_TRY(cudaMalloc(&s->mel_hi, (size_t)n * 2)); VS_CUDA_TRY(cudaMalloc(&s->mel_lo, (size_t)n * 2));
    k_make_mel_basis<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d->sample_rate, n_fft, d->num_mels, s->bins, s->ldp, s->mel_hi, s->mel_lo);
    VS_CUDA_TRY(cudaGetLastError());
    return VS_OK;
}

int vs_encoder_load_params(vs_engine* e, const vs_encoder_params* p, void* stream) {
    if (!e || !e->encoder || !p) { set_error("call vs_encoder_configure first"); return VS_ERR_STATE; }
    EncoderState* s = (EncoderState*)e->encoder;
    cudaStream_t st = (cudaStream_t)stream;
    const int H = s->d.lstm_hidden, nslices = H / kEU, melp = (s->d.num_mels + 7) / 8 * 8;
    for (int l = 0; l < s->d.lstm_layers; ++l) {
        if (!p->w_ih[l] || !p->w_hh[l] || !p->b_ih[l] || !p->b_hh[l]) { set_error("encoder: missing LSTM parameter"); return VS_ERR_INVALID; }
        const int K = l == 0 ? s->d.num_mels : H, Kp = l == 0 ? melp : H;
        const size_t ni = (size_t)4 * H * Kp, nh = (size_t)nslices * 32 * H;
        if (!s->wih_hi[l]) {
            VS_CUDA_TRY(cudaMalloc(&s->wih_hi[l], ni * 2)); VS_CUDA_TRY(cudaMalloc(&s->wih_lo[l], ni * 2));
            VS_CUDA_TRY(cudaMalloc(&s->whh_hi[l], nh * 2)); VS_CUDA_TRY(cudaMalloc(&s->whh_lo[l], nh * 2));
            VS_CUDA_TRY(cudaMalloc(&s->bias[l], (size_t)4 * H * 4));
        }
        k_enc_split<<<(unsigned)((ni + 255) / 256), 256, 0, st>>>(p->w_ih[l], 4 * H, K, Kp, s->wih_hi[l], s->wih_lo[l]);
        k_enc_pack_whh<<<(unsigned)((nh + 255) / 256), 256, 0, st>>>(p->w_hh[l], H, H, nslices, s->whh_hi[l], s->whh_lo[l]);
        k_enc_bias<<<(4 * H + 255) / 256, 256, 0, st>>>(p->b_ih[l], p->b_hh[l], s->bias[l], 4 * H);
    }
    if (!p->proj_w || !p->proj_b) { set_error("encoder: missing projection parameter"); return VS_ERR_INVALID; }
    const size_t np = (size_t)s->d.emb_dim * H;
    if (!s->proj_hi) {
        VS_CUDA_TRY(cudaMalloc(&s->proj_hi, np * 2)); VS_CUDA_TRY(cudaMalloc(&s->proj_lo, np * 2));
        VS_CUDA_TRY(cudaMalloc(&s->proj_b, (size_t)s->d.emb_dim * 4));
    }
    k_enc_split<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(p->proj_w, s->d.emb_dim, H, H, s->proj_hi, s->proj_lo);
    VS_CUDA_TRY(cudaMemcpyAsync(s->proj_b, p->proj_b, (size_t)s->d.emb_dim * 4, cudaMemcpyDeviceToDevice, st));
    VS_CUDA_TRY(cudaGetLastError());
    s->loaded = true;
    return VS_OK;
}

size_t vs_encoder_workspace_bytes(const vs_engine* e, int32_t B, int32_t L_or_frames, int32_t from_wav) {
    if (!e || !e->encoder || B < 1 || L_or_frames < 1) return 0;
    const EncoderState* s = (const EncoderState*)e->encoder;
    const int L = from_wav ? L_or_frames : 0, T = from_wav ? 1 + L_or_frames / s->hop : L_or_frames;
    return enc_carve(e, s, B, T, L, nullptr).total;
}

int vs_encoder_mel(vs_engine* e, const float* wav, float* mel_out, int32_t B, int32_t L, void* workspace, size_t workspace_bytes, void* stream) {
    EncoderState* s = enc_state(e, false);
    if (!s) return VS_ERR_STATE;
    if (!wav || !mel_out || !workspace || B < 1 || L < 1) { set_error("bad argument"); return VS_ERR_INVALID; }
    const int T = 1 + L / s->hop;
    EncWs w = enc_carve(e, s, B, T, L, workspace);
    if (workspace_bytes < w.total) { set_error("workspace too small"); return VS_ERR_STATE; }
    e->launches = 0;
    prof_begin(e, (cudaStream_t)stream);
    return enc_mel(e, s, w, wav, mel_out, B, L, T, (cudaStream_t)stream);
}

int vs_encoder_forward(vs_engine* e, const float* mel, float* dvec, int32_t B, int32_t T, void* workspace, size_t workspace_bytes, void* stream) {
    EncoderState* s = enc_state(e, true);
    if (!s) return VS_ERR_STATE;
    if (!mel || !dvec || !workspace || B < 1) { set_error("bad argument"); return VS_ERR_INVALID; }
    if (T < s->d.window) { set_error("reference audio shorter than one encoder window"); return VS_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    EncWs w = enc_carve(e, s, B, T, 0, workspace);
    if (workspace_bytes < w.total) { set_error("workspace too small"); return VS_ERR_STATE; }
    e->launches = 0;
    prof_begin(e, st);
    const size_t n = (size_t)B * T * w.melp;
    k_enc_split<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(mel, B * T, s->d.num_mels, w.melp, w.x_hi, w.x_lo);
    VS_LAUNCH(e, KID_CONVERT, st, cudaGetLastError());
    return enc_stack(e, s, w, dvec, B, T, st);
}

int vs_encoder_dvector(vs_engine* e, const float* wav, float* dvec, int32_t B, int32_t L, void* workspace, size_t workspace_bytes, void* stream) {
    EncoderState* s = enc_state(e, true);
    if (!s) return VS_ERR_STATE;
    if (!wav || !dvec || !workspace || B < 1 || L < 1) { set_error("bad argument"); return VS_ERR_INVALID; }
    const int T = 1 + L / s->hop;
    if (T < s->d.window) { set_error("reference audio shorter than one encoder window"); return VS_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    EncWs w = enc_carve(e, s, B, T, L, workspace);
    if (workspace_bytes < w.total) { set_error("workspace too small"); return VS_ERR_STATE; }
    e->launches = 0;
    prof_begin(e, st);
    int rc = enc_mel(e, s, w, wav, nullptr, B, L, T, st);
    if (rc != VS_OK) return rc;
    return enc_stack(e, s, w, dvec, B, T, st);
}

}  // extern "C"
