/*
 * voicesplit_b200 - C ABI of the B200 (sm_100a) mask-estimation engine.
 *
 * The reference (Edresson/VoiceSplit) has no FFI or plugin interface: its boundary for this
 * path is the Python nn.Module contract.  This header is therefore the
 * seam *beneath* that contract: the repo's own models/voicesplit/model.py and
 * models/voicefilter/model.py (same class names, constructor, forward signature and state_dict
 * as /root/reference/models/voicesplit/model.py:9-89 and
 * /root/reference/models/voicefilter/model.py:11-90) bind exactly these entry points through
 * ctypes (voicesplit_b200/_cabi.py).  INTEGRATION.md shows the stub a reference maintainer adds.
 *
 * Conventions: plain pointers and sizes only; every function returns VS_OK (0) or a negative
 * VS_ERR_* code and never throws; vs_last_error() gives the message for the calling thread.
 * Device pointers are CUDA device memory of the current device; `stream` is a cudaStream_t
 * passed as void* (NULL = legacy default stream).  No function allocates device memory on the
 * hot call (vs_forward / vs_conv_stack): parameters are packed once by vs_engine_load_params
 * and scratch space is a caller-provided workspace of vs_workspace_bytes().
 */
#ifndef VOICESPLIT_B200_H
#define VOICESPLIT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VS_ABI_VERSION 2

#define VS_OK 0
#define VS_ERR_INVALID (-1)     /* bad argument / unsupported shape */
#define VS_ERR_CUDA (-2)        /* a CUDA runtime/driver call failed */
#define VS_ERR_STATE (-3)       /* parameters not loaded, workspace too small, ... */
#define VS_ERR_UNSUPPORTED (-4) /* feature not built for this device (needs sm_100a) */

/* activation after each BatchNorm: Mish = VoiceSplit (reference utils/generic_utils.py:395-399),
 * ReLU = VoiceFilter (reference models/voicefilter/model.py:21) */
#define VS_ACT_MISH 0
#define VS_ACT_RELU 1

/* arithmetic of the contractions (conv stack, LSTM input projection, FC head):
 *   FP32    fp32 FFMA on CUDA cores - exact-order-independent fp32, the on-GPU ground truth
 *   BF16X3  tcgen05 tensor cores, operands split a = hi + lo (bf16 each), three MMAs
 *           (hi*hi + hi*lo + lo*hi) into one fp32 TMEM accumulator: ~2^-16 relative operand error
 *   BF16    tcgen05 tensor cores, single bf16 pass (fast mode; error reported, not hidden) */
#define VS_PREC_FP32 0
#define VS_PREC_BF16X3 1
#define VS_PREC_BF16 2
/*   FP16X3 / FP16: the same two schemes with IEEE half operands (11-bit significand instead of 8):
 *           hi + lo carries ~22 bits, i.e. fp32-grade products.  Half has a narrow exponent, so
 *           weights are pre-scaled by a per-layer power of two (undone exactly in the epilogue) and
 *           activations are clamped to +-60000 before conversion. */
#define VS_PREC_FP16X3 3
#define VS_PREC_FP16 4
/*   FP16_F8C: the conv stack (97 % of the tensor-core work) keeps every activation as fp16 `hi` plus a 128-byte e4m3
 *           correction row [2^8 (x - hi) | 2^-2 hi] and every weight as fp16 `hi` plus [2^-8 w_hi | 2^2 w_lo]; per filter-tap
 *           pair it issues the fp16 product (4 kind::f16 MMAs) and ONE fp8 product over that 128-byte K axis (4 kind::f8f6f4
 *           MMAs at twice the rate) = x_lo w_hi + x_hi w_lo, into the same fp32 TMEM accumulator: 8 MMA slots instead of the
 *           12 of FP16X3, correction terms carried with 4 significant bits (relative operand error ~2^-15).  LSTM input
 *           projection, recurrence and FC head run as FP16X3. */
#define VS_PREC_FP16_F8C 5

typedef struct vs_engine vs_engine;

/* mirrors the config.json keys the reference module reads (reference config.json:36-41,86;
 * models/voicesplit/model.py:13,58-64) */
typedef struct vs_dims {
    int32_t num_freq; /* audio[backend].num_freq, F */
    int32_t emb_dim;  /* model.emb_dim, E */
    int32_t lstm_dim; /* model.lstm_dim, H */
    int32_t fc1_dim;  /* model.fc1_dim */
    int32_t fc2_dim;  /* model.fc2_dim, must equal num_freq */
    int32_t activation; /* VS_ACT_* */
} vs_dims;

/* Device pointers to the fp32 parameters in the reference's state_dict layout:
 * conv l=0..7 are Sequential positions 1,5,9,13,17,21,25,28; bn l the BatchNorm after it. */
typedef struct vs_params {
    const float* conv_w[8];  /* [Cout][Cin][kh][kw] */
    const float* conv_b[8];  /* [Cout] */
    const float* bn_gamma[8];
    const float* bn_beta[8];
    const float* bn_mean[8]; /* running_mean */
    const float* bn_var[8];  /* running_var  */
    const float* w_ih[2];    /* [4H][8F+E], direction 0 = forward, 1 = reverse */
    const float* w_hh[2];    /* [4H][H] */
    const float* b_ih[2];    /* [4H] */
    const float* b_hh[2];    /* [4H] */
    const float* fc1_w;      /* [fc1][2H] */
    const float* fc1_b;
    const float* fc2_w;      /* [F][fc1] */
    const float* fc2_b;
} vs_params;

int vs_abi_version(void);
const char* vs_last_error(void);

/* Engine life cycle.  vs_engine_create selects the current CUDA device and fails with
 * VS_ERR_UNSUPPORTED if it is not compute capability 10.x. */
int vs_engine_create(const vs_dims* dims, vs_engine** out);
int vs_engine_destroy(vs_engine* e);

/* Fold eval-mode BatchNorm into per-channel scale/shift, repack the weights into the kernels'
 * layouts (fp32 + bf16 hi/lo planes, tap-pair-major for the tensor-core conv).  Must be called
 * again whenever the parameters change (optimizer step, load_state_dict). */
int vs_engine_load_params(vs_engine* e, const vs_params* device_params, void* stream);

/* Scratch bytes vs_forward / vs_conv_stack need for a batch of B utterances of T frames. */
size_t vs_workspace_bytes(const vs_engine* e, int32_t B, int32_t T, int32_t precision);

/* The hot path: reference VoiceSplit.forward (models/voicesplit/model.py:66-89) plus the
 * caller-side mask apply (train.py:95).
 *   x      [B][T][F] fp32 device, spectrogram magnitudes
 *   emb    [B][E]    fp32 device, d-vectors
 *   mask   [B][T][F] fp32 device, out
 *   masked [B][T][F] fp32 device, out, x*mask; may be NULL */
int vs_forward(vs_engine* e, const float* x, const float* emb, float* mask, float* masked,
               int32_t B, int32_t T, int32_t precision, void* workspace, size_t workspace_bytes,
               void* stream);

/* Same call with HOST buffers (pinned or pageable): copies x/emb to the device, runs vs_forward,
 * copies mask (and masked) back and synchronises the stream.  The engine keeps a grow-only
 * device staging area for this entry point, so repeated calls of the same shape do not allocate. */
int vs_forward_host(vs_engine* e, const float* x_host, const float* emb_host, float* mask_host,
                    float* masked_host, int32_t B, int32_t T, int32_t precision, void* stream);
/* The host entry points are the only forward calls that own device memory: the FIRST call with a larger (B, T, precision)
 * than seen before does one cudaMalloc (after synchronising the stream) - every later call of that or a smaller shape
 * allocates nothing.  vs_forward_host_reserve does that allocation up front (staging of vs_forward_host and both slots +
 * the shared workspace of vs_forward_host_submit), so that a serving loop never allocates on a request. */
int vs_forward_host_reserve(vs_engine* e, int32_t B, int32_t T, int32_t precision);

/* Pipelined form of vs_forward_host for serving loops: two slots (0, 1), each with its own device
 * staging.  submit enqueues  H2D (copy stream) -> vs_forward (compute stream) -> D2H (copy stream)
 * for one batch and returns immediately; wait blocks until that slot's outputs are in host memory.
 * Submitting slot s while slot 1-s is in flight overlaps its copies with the other slot's compute.
 * Host buffers must be pinned (cudaHostAlloc / torch pin_memory) for the copies to be asynchronous,
 * and must not be touched between submit and wait.  A slot must be waited on before it is reused. */
int vs_forward_host_submit(vs_engine* e, int32_t slot, const float* x_host, const float* emb_host,
                           float* mask_host, float* masked_host, int32_t B, int32_t T, int32_t precision);
int vs_forward_host_wait(vs_engine* e, int32_t slot);

/* Conv stack only (reference model.conv + the transpose/view of model.py:70-74):
 * x [B][T][F] -> conv_out [B][T][8F] fp32 device (index c*F+f). */
int vs_conv_stack(vs_engine* e, const float* x, float* conv_out, int32_t B, int32_t T,
                  int32_t precision, void* workspace, size_t workspace_bytes, void* stream);

/* ---- training (BASELINE config 4): fp32, BatchNorm with batch statistics, full backward -------------
 * vs_train_forward is VoiceSplit.forward in .train() mode (models/voicesplit/model.py:66-89 under
 * train.py:84,94): every BatchNorm normalises with the statistics of this batch and, when `bn` is
 * given, updates running_mean / running_var (momentum, unbiased variance) and num_batches_tracked in
 * place - the tensors of the module's own buffers.  It keeps what the backward needs in `workspace`
 * (vs_train_workspace_bytes), which must stay untouched until vs_train_backward has run.
 * vs_train_backward consumes d(loss)/d(mask) and writes the gradient of every parameter in the
 * reference's layouts (vs_grads mirrors vs_params) plus, optionally, d(loss)/d(emb) and d(loss)/d(x).  Parameters must
 * have been loaded with vs_engine_load_params since their last change. */
typedef struct vs_train_state {
    float* running_mean[8];
    float* running_var[8];
    int64_t* num_batches_tracked[8];
    float momentum; /* nn.BatchNorm2d default 0.1 */
} vs_train_state;
typedef struct vs_grads {
    float* conv_w[8];
    float* conv_b[8];
    float* bn_gamma[8];
    float* bn_beta[8];
    float* w_ih[2];
    float* w_hh[2];
    float* b_ih[2];
    float* b_hh[2];
    float* fc1_w;
    float* fc1_b;
    float* fc2_w;
    float* fc2_b;
} vs_grads;
/* The conv forward and data-gradient of the training path run on the tcgen05 conv kernel by default
 * (fp16x3 activations, bf16x3 gradients: fp32-grade, no loss scaling); 0 selects the fp32 CUDA-core convs. */
int vs_engine_set_train_tensor_cores(vs_engine* e, int32_t enabled);

/* Data-parallel training hooks.  Both callbacks run on the calling host thread in the middle of
 * vs_train_forward / vs_train_backward and must only ENQUEUE work that is ordered after everything already enqueued on `stream`
 * (an NCCL collective issued from that stream, or from a side stream that waits on an event recorded there); they return 0 or
 * a non-zero error code, which aborts the call with VS_ERR_STATE.
 *   vs_stat_allreduce_fn  SyncBN: sum `count` doubles on the device, in place, across the `world_size` ranks.  Called once per
 *                         BatchNorm layer in the forward (per-channel sum z, sum z^2) and once in the backward (sum du,
 *                         sum du xhat); the engine then divides by world_size x the local element count, so every rank
 *                         normalises with the statistics of the CONCATENATED batch - the reference's single-process semantics
 *                         (train.py:84-111 with plain nn.BatchNorm2d).  fn = NULL (default): per-rank statistics.
 *   vs_backward_hook_fn   stage VS_BWD_STAGE_LSTM_FC_DONE: the gradients of the LSTM and FC parameters (97 % of the 75.5 MB)
 *                         are enqueued, the conv-stack backward follows - the point where their all-reduce can start. */
typedef int (*vs_stat_allreduce_fn)(void* user, double* device_sums, int32_t count, void* stream);
typedef int (*vs_backward_hook_fn)(void* user, int32_t stage, void* stream);
#define VS_BWD_STAGE_LSTM_FC_DONE 1
int vs_engine_set_sync_bn(vs_engine* e, vs_stat_allreduce_fn fn, void* user, int32_t world_size);
int vs_engine_set_backward_hook(vs_engine* e, vs_backward_hook_fn fn, void* user);
size_t vs_train_workspace_bytes(const vs_engine* e, int32_t B, int32_t T);
int vs_train_forward(vs_engine* e, const vs_train_state* bn, const float* x, const float* emb, float* mask,
                     int32_t B, int32_t T, void* workspace, size_t workspace_bytes, void* stream);
/* grad_emb [B][E] and grad_x [B][T][F] (d loss / d spectrogram, through cnn1) are optional outputs (NULL = not wanted). */
int vs_train_backward(vs_engine* e, const float* x, const float* emb, const float* mask, const float* grad_mask,
                      const vs_grads* grads, float* grad_emb, float* grad_x, int32_t B, int32_t T, void* workspace,
                      size_t workspace_bytes, void* stream);

/* ---- audio front / back end (BASELINE config 5) --------------------------
 * The reference's openVoiceFilterAudioProcessor.wav2spec / spec2wav with the mixture phase
 * (utils/audio_processor.py:469-496): librosa.stft(n_fft, hop, win, hann, center, reflect) -> |D| -> dB ->
 * clip-normalise to [0,1], and back through librosa.istft.  vs_audio_configure builds the (i)DFT operands
 * for the given parameters (n_fft / 2 + 1 must equal num_freq).
 *   vs_wav2spec: wav [B][L] fp32 device -> spec [B][T][F] in [0,1], phasor [B][T][F][2] = D / |D|, T = 1 + L / hop
 *   vs_spec2wav: (masked) spec + phasor -> wav_out [B][hop * (T - 1)]
 * A whole separation is  vs_wav2spec -> vs_forward (masked output) -> vs_spec2wav. */
typedef struct vs_audio_params {
    int32_t n_fft, hop_length, win_length;
    float min_level_db, ref_level_db;   /* reference config.json:90-95: -100, 20 */
} vs_audio_params;
int vs_audio_configure(vs_engine* e, const vs_audio_params* params, void* stream);
size_t vs_audio_workspace_bytes(const vs_engine* e, int32_t B, int32_t L);
int vs_wav2spec(vs_engine* e, const float* wav, float* spec, float* phasor, int32_t B, int32_t L, void* workspace,
                size_t workspace_bytes, void* stream);
int vs_spec2wav(vs_engine* e, const float* spec, const float* phasor, float* wav_out, int32_t B, int32_t T,
                void* workspace, size_t workspace_bytes, void* stream);

/* ---- training-loss chain on the device (BASELINE config 4) -----------------
 * What train.py:95-109 runs after the mask every step:
 *   wav = ap.torch_inv_spectrogram(spec, spec_phase)      utils/audio_processor.py:498-509 (torchaudio istft)
 *   loss = SiSNR_With_Pit()(wav_est, wav_target, seq_len)  utils/generic_utils.py:403-474, C = 1 source
 * phase_mode VS_ISTFT_Q1 reproduces the reference verbatim (real = mag e^{cos phi}, imag =
 * mag e^{sin phi}, symmetric Hann); VS_ISTFT_CORRECTED uses mag (cos phi, sin phi) and the periodic Hann of
 * the analysis side.  All pointers are device pointers; spec / phase are [B][T][F] fp32 (phase = angle in
 * radians), waveforms [B][hop (T - 1)], seq_len [B] int64 (valid samples, the dataset's wav length).
 *   vs_loss_spec2wav           spec, phase -> wav                        (the differentiable iSTFT, forward)
 *   vs_loss_spec2wav_backward  d loss / d wav -> d loss / d spec         (its backward; spec/phase = forward inputs)
 *   vs_sisnr_loss              est_spec, target_spec, phase, seq_len -> loss (device scalar), optional per-utterance
 *                              Si-SNR [B] and optional grad_est = d loss / d est_spec [B][T][F]: both iSTFTs, the
 *                              masked zero-mean Si-SNR, its analytic gradient and the iSTFT backward in one call,
 *                              no host synchronisation, no Python loop over the batch (generic_utils.py:413-414). */
enum { VS_ISTFT_Q1 = 0, VS_ISTFT_CORRECTED = 1 };
typedef struct vs_loss_params {
    int32_t n_fft, hop_length, win_length;
    float min_level_db, ref_level_db;
    int32_t phase_mode;
} vs_loss_params;
int vs_loss_configure(vs_engine* e, const vs_loss_params* params, void* stream);
size_t vs_loss_workspace_bytes(const vs_engine* e, int32_t B, int32_t T);
int vs_loss_spec2wav(vs_engine* e, const float* spec, const float* phase, float* wav_out, int32_t B, int32_t T,
                     void* workspace, size_t workspace_bytes, void* stream);
int vs_loss_spec2wav_backward(vs_engine* e, const float* spec, const float* phase, const float* grad_wav,
                              float* grad_spec, int32_t B, int32_t T, void* workspace, size_t workspace_bytes,
                              void* stream);
int vs_sisnr_loss(vs_engine* e, const float* est_spec, const float* target_spec, const float* phase,
                  const int64_t* seq_len, float* loss_out, float* snr_out, float* grad_est, int32_t B, int32_t T,
                  void* workspace, size_t workspace_bytes, void* stream);

/* ---- evaluation metrics on the device (utils/generic_utils.py:476-533, test.py:71) ----
 *   vs_sisnr_wav  SiSNR_With_Pit on waveforms, one source per utterance (utils/generic_utils.py:417-474): est / target
 *                 [B][L] fp32, seq_len [B] int64 -> loss = 20 - mean(snr) (device scalar) and snr [B] (required).
 *                 NOTE validation() calls criterion(clean, est) - arguments swapped w.r.t. training;
 *                 the caller chooses the order, the kernel treats its first waveform as the estimate.
 *   vs_sdr        mir_eval.separation.bss_eval_sources(ref, est, compute_permutation=False)[0][0] per utterance
 *                 (utils/generic_utils.py:511): 512-tap projection SDR in double precision; ref / est [B][L] fp32 -> [B]. */
int vs_sisnr_wav(vs_engine* e, const float* est_wav, const float* target_wav, const int64_t* seq_len, float* loss_out,
                 float* snr_out, int32_t B, int32_t L, void* stream);
size_t vs_sdr_workspace_bytes(int32_t B, int32_t L);
int vs_sdr(vs_engine* e, const float* ref_wav, const float* est_wav, float* sdr_out, int32_t B, int32_t L,
           void* workspace, size_t workspace_bytes, void* stream);

/* ---- GE2E speaker encoder: the producer of the d-vector ---------------------
 * Replaces, for the extraction loop of notebooks/GE2E-Seungwonpark-ExtractSpeakerEmbedding-adaptado-para-openvoicefilter.py
 * (:141-143):  mel = ap.get_mel(wav)   (utils/audio_processor.py:456-468: |librosa.stft|^2 -> mel basis -> log10(. + 1e-6))
 *              dvec = embedder(mel)    (SpeakerEncoder.forward, notebook :75-85: windows of `window` frames every `stride`,
 *                                       `lstm_layers` x LSTM(lstm_hidden), last frame, Linear -> emb_dim, L2 normalise, mean)
 * The mel front end shares the STFT of vs_audio_configure (call that first; n_fft / hop / win come from it).
 * Device layouts: wav [B][L] fp32; mel [B][T][num_mels] fp32 (frames are rows: the transpose of the reference's
 * [num_mels, T]); dvec [B][emb_dim] fp32.  All utterances of a call have the same length; T = 1 + L / hop >= window.
 * Parameters are PyTorch nn.LSTM / nn.Linear tensors (gate order i, f, g, o), device pointers, per layer. */
typedef struct vs_encoder_dims {
    int32_t num_mels, lstm_layers, lstm_hidden, emb_dim, window, stride; /* notebook :34-41: 40, 3, 768, 256, 80, 40 */
    int32_t sample_rate;                                                  /* 16000 */
} vs_encoder_dims;
typedef struct vs_encoder_params {
    const float* w_ih[4]; /* lstm.weight_ih_l{k} [4H][num_mels | H] */
    const float* w_hh[4]; /* lstm.weight_hh_l{k} [4H][H] */
    const float* b_ih[4]; /* lstm.bias_ih_l{k} [4H] */
    const float* b_hh[4]; /* lstm.bias_hh_l{k} [4H] */
    const float* proj_w;  /* proj.linear_layer.weight [emb_dim][H] */
    const float* proj_b;  /* proj.linear_layer.bias [emb_dim] */
} vs_encoder_params;
int vs_encoder_configure(vs_engine* e, const vs_encoder_dims* dims, void* stream);
int vs_encoder_load_params(vs_engine* e, const vs_encoder_params* params, void* stream);
/* from_wav != 0: `n` is the waveform length L (vs_encoder_mel / vs_encoder_dvector); else the frame count T (vs_encoder_forward) */
size_t vs_encoder_workspace_bytes(const vs_engine* e, int32_t B, int32_t n, int32_t from_wav);
int vs_encoder_mel(vs_engine* e, const float* wav, float* mel_out, int32_t B, int32_t L, void* workspace,
                   size_t workspace_bytes, void* stream);
int vs_encoder_forward(vs_engine* e, const float* mel, float* dvec, int32_t B, int32_t T, void* workspace,
                       size_t workspace_bytes, void* stream);
int vs_encoder_dvector(vs_engine* e, const float* wav, float* dvec, int32_t B, int32_t L, void* workspace,
                       size_t workspace_bytes, void* stream);

/* Test hooks: run a single conv layer l (0..6 -> 64-channel output) on an fp32 NCHW input
 * in [B][Cin][T][F] -> out [B][64][T][F], and the BiLSTM + head on a given conv_out.
 * They allocate internally and synchronise; not for the hot path. */
int vs_debug_conv_layer(vs_engine* e, int32_t layer, const float* in_nchw, float* out_nchw,
                        int32_t B, int32_t T, int32_t precision, void* stream);
int vs_debug_lstm_head(vs_engine* e, const float* conv_out, const float* emb, const float* x,
                       float* lstm_out, float* mask, int32_t B, int32_t T, int32_t precision,
                       void* stream);

/* Phase timers of the last tensor-core LSTM recurrence (SM cycles summed over all steps, CTA 0):
 * [0] producer waits on the group barrier, [1] producer issues TMA, [2] MMA thread waits for h blocks,
 * [3] MMA issue, [4] cell thread waits for the accumulator, [5] TMEM load + gate math, [6] stores,
 * [7] CTA barrier + fence + atomic.  Synchronises the device. */
int vs_debug_lstm_timing(vs_engine* e, int64_t* cycles8);

/* Number of kernels this library launched during the last vs_forward / vs_conv_stack. */
int vs_last_launch_count(const vs_engine* e);

/* Per-kernel device timing of the last vs_forward / vs_conv_stack: when enabled, a CUDA event is
 * recorded on the launch stream after every kernel.  vs_profile_read waits for the last event and
 * returns the number of entries written (kernel id, milliseconds); ids: 0 cnn1, 1..6 cnn2..cnn7,
 * 7 cnn8+reshape, 8 d-vector gate bias, 9 LSTM input projection, 10 LSTM recurrence, 11 fc1,
 * 12 fc2+sigmoid+mask, 13 layout/precision conversion, 14 fused head. */
int vs_engine_set_profiling(vs_engine* e, int32_t enabled);
int vs_profile_read(vs_engine* e, int32_t max_entries, int32_t* kernel_ids, float* milliseconds);

#ifdef __cplusplus
}
#endif
#endif /* VOICESPLIT_B200_H */
