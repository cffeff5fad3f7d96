// Tensor-core (tcgen05 / TMEM / TMA) path of the engine: interface used by engine.cu.
#pragma once
#include "common.cuh"

namespace vs {

// precision -> (number of MMA passes, 16-bit element type: 0 = bf16, 1 = fp16)
// FP16_F8C keeps two operand planes like the x3 modes (fp16 hi + the e4m3 correction plane, common.cuh), so it reports "3":
// every buffer is sized as for FP16X3; the conv kernels then issue 4 f16 + 4 f8f6f4 MMAs per tap pair instead of 12 f16.
inline int tc_passes(int precision) { return (precision == VS_PREC_BF16X3 || precision == VS_PREC_FP16X3 || precision == VS_PREC_FP16_F8C) ? 3 : 1; }
inline int tc_elt(int precision) { return (precision == VS_PREC_FP16X3 || precision == VS_PREC_FP16 || precision == VS_PREC_FP16_F8C) ? 1 : 0; }
inline bool tc_f8c(int precision) { return precision == VS_PREC_FP16_F8C; }
// arithmetic of the LSTM input projection / recurrence / FC head under each conv precision
inline int tc_head_precision(int precision) { return precision == VS_PREC_FP16_F8C ? VS_PREC_FP16X3 : precision; }

struct TcLstmBuffers {  // recurrent-kernel buffers shared with the fp32 path (carved by engine.cu)
    float* gates;       // [B*T][8H]
    float* bias_u;      // [B][8H]
    float* hout;        // [B][T][2H]
    float* hx;          // recurrent exchange + cell state
    unsigned int* barrier;
};

int tc_create(vs_engine* e);
void tc_destroy(vs_engine* e);
int tc_pack(vs_engine* e, cudaStream_t st);  // after the fp32 packing of vs_engine_load_params
size_t tc_workspace_bytes(const vs_engine* e, int B, int T, int precision);
int tc_forward(vs_engine* e, const float* x, const float* emb, float* mask, float* masked, int B, int T,
               int precision, void* ws, const TcLstmBuffers& lb, cudaStream_t st);
int tc_conv_stack(vs_engine* e, const float* x, float* conv_out, int B, int T, int precision, void* ws, cudaStream_t st);
// single layer on fp32 planes (converted to/from the bf16 hi/lo planes internally)
int tc_debug_layer(vs_engine* e, int layer, const float* x, const float* plane_in, float* plane_out, int B, int T,
                   int precision, cudaStream_t st);
int tc_debug_lstm_head(vs_engine* e, const float* conv_out, const float* emb, const float* x, float* mask, int B, int T,
                       int precision, const TcLstmBuffers& lb, cudaStream_t st);

// gate nonlinearities of the recurrent kernels (tc_lstm.cu, encoder.cu)
__device__ __forceinline__ float sigmoid_fast(float x) {
    // 1 / (1 + 2^(-x log2 e)) with the raw MUFU approximations (no range/denormal slow paths):
    // ex2.approx and rcp.approx are each accurate to ~2^-22 relative
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
    return r;
}
__device__ __forceinline__ float tanh_fast(float x) { return fmaf(2.f, sigmoid_fast(2.f * x), -1.f); }


// ---- tc_gemm.cu: the tensor-core GEMM (also used by audio.cu) -----------------------------------
enum { GEPI_GATES = 0, GEPI_FC1 = 1, GEPI_FC2 = 2, GEPI_PLAIN = 3, GEPI_STFT = 4, GEPI_ISTFT_BWD = 5, GEPI_STFT_POWER = 6, GEPI_LOGMEL = 7 };

struct GemmTcArgs {
    int M, N, K;
    int lda, ldw;               // row strides (elements) of the 16-bit A and W planes, multiples of 8
    int tn;                     // 1: C[M][N] = sum_k A[k][m] * W[k][n] (both operands [K rows][cols], MN-major MMAs; N tile % 64 == 0)
    int n_tile, n_tiles_n, n_tiles_m, total_tiles, n_kb;
    int passes, stages;
    int csz;                    // CTAs per cluster (1 or 2): pairs of M blocks share every W tile through TMA multicast (launch_gemm_tc decides)
    int group_rows;             // GATES: rows per utterance (T)
    const float* bias;          // [N] (FC1/FC2) or null
    const float* bias_group;    // [M / group_rows][N] (GATES)
    float* out32;               // GATES: [M][N]; FC2: mask [M][N]
    int ld_out;
    elt16* out_hi;              // FC1: [M][ld16]
    elt16* out_lo;
    int ld16;
    const float* xmul;          // FC2: spectrogram [M][N]
    float* masked;              // FC2: optional
    // STFT epilogue (audio.cu): out32 = normalised dB magnitude [utt][t_valid][n_bins], phasor = D / |D| as float2
    float* phasor;
    int rows_per_utt, t_valid, n_bins;
    float min_db, ref_db;
    // iSTFT-backward epilogue (loss.cu): columns are (dRe, dIm) of a bin; out32 = d loss / d normalised spectrogram
    // [utt][t_valid][n_bins], chained through the complex build of torch_spec2wav (g_spec = its input, g_phase = angles)
    const float* g_spec;
    const float* g_phase;
    int q1;                     // 1: reference-verbatim exp(cos), exp(sin) weights; 0: cos, sin
    // encoder.cu: STFT_POWER writes |D|^2 of bin k as bf16 hi/lo into out_hi/out_lo [utt * t_valid + t][ld16];
    //             LOGMEL writes log10(acc + 1e-6) as fp32 (out32, optional) and fp16 hi/lo (out_hi/out_lo, [M][ld16])
};
// a.M/N/K, lda, ldw and the epilogue fields must be set; tile shape and pipeline depth are derived here
int launch_gemm_tc(vs_engine* e, int epi, int kid, const elt16* a_hi, const elt16* a_lo, const elt16* w_hi, const elt16* w_lo,
                   GemmTcArgs a, int precision, cudaStream_t st);



// audio.cu services used by encoder.cu
size_t audio_stft_scratch_bytes(const vs_engine* e, int B, int L);
int audio_stft_power(vs_engine* e, const float* wav, elt16* pw_hi, elt16* pw_lo, int ld16, int B, int L, void* scratch, cudaStream_t st);
void audio_geometry(const vs_engine* e, int* n_fft, int* hop, int* win);

// training GEMMs on the tensor-core kernel (tc_gemm.cu)
size_t tc_train_gemm_workspace_bytes(const vs_engine* e, int B, int T);
int tc_train_inproj(vs_engine* e, const float* xcat, const float* bias_u, float* gates, void* ws, int B, int T, cudaStream_t st);
int tc_train_lstm_input_grads(vs_engine* e, const float* da /*[M][8H]*/, const float* xcat, float* dxcat, float* dw_ih0, float* dw_ih1, int ld_dw,
                              void* ws, int B, int T, cudaStream_t st);

// ---- tc_gemm.cu: LSTM input projection / recurrence / FC head of the tensor-core path ------------
int tc_gemm_pack(vs_engine* e, cudaStream_t st);
void tc_gemm_destroy(vs_engine* e);
size_t tc_gemm_workspace_bytes(const vs_engine* e, int B, int T, int precision);
// Everything after the 64-channel conv planes: cnn8 (+reshape), LSTM, head.  If conv_out32 is given
// the planes are ignored and the LSTM input is taken from it (debug hook).
int tc_lstm_head(vs_engine* e, const elt16* plane_hi, const elt16* plane_lo, bool plane_f8c, const float* conv_out32,
                 const float* emb, const float* x, float* mask, float* masked, int B, int T, int precision, void* gemm_ws,
                 const TcLstmBuffers& lb, cudaStream_t st);
// training (train.cu): raw forward conv / data-gradient conv of layer 1..6 on k_conv_tc, 3 passes, fp32 output plane
int tc_train_conv(vs_engine* e, int layer, bool dgrad, const elt16* in_hi, const elt16* in_lo, const float* shift, float* out32, int B, int T,
                  int elt, int kid, cudaStream_t st);

// tc_wgrad.cu: weight gradient of layer 1..6 from bf16 hi/lo activation and dz planes; writes the reference layout
int tc_train_wgrad(vs_engine* e, int layer, const elt16* a_hi, const elt16* a_lo, const elt16* d_hi, const elt16* d_lo, float* dwp,
                   float* dw_out, int B, int T, int kid, cudaStream_t st);

// ---- tc_lstm.cu: tensor-core recurrent kernel ---------------------------------------------------
int tc_lstm_pack(vs_engine* e, void** slot, cudaStream_t st);
void tc_lstm_destroy(void* slot);
size_t tc_lstm_scratch_bytes(const vs_engine* e, int B);
int tc_lstm_recurrence(vs_engine* e, void* slot, const float* gates_x, float* hout, void* scratch, elt16* hr_hi, elt16* hr_lo,
                       int B, int T, int precision, cudaStream_t st);
void* tc_lstm_slot(vs_engine* e);  // LstmState pointer kept in TcState (tc_conv.cu)
int tc_lstm_read_timing(void* slot, long long* out8);

cudaError_t tc_launch_point8(const vs_engine* e, const elt16* hi, const elt16* lo, int elt, bool f8c, float* x32, elt16* xhi,
                             elt16* xlo, int ldx, int B, int T, cudaStream_t st);

}  // namespace vs
