// Shared device helpers and the engine's internal declarations.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/voicesplit_b200.h"

namespace vs {

// ---- conv stack geometry: reference models/voicesplit/model.py:15-52 --------------------------
struct ConvGeom { int cin, cout, kh, kw, dil; };
static constexpr ConvGeom kConv[8] = {
    {1, 64, 1, 7, 1},  {64, 64, 7, 1, 1}, {64, 64, 5, 5, 1}, {64, 64, 5, 5, 2},
    {64, 64, 5, 5, 4}, {64, 64, 5, 5, 8}, {64, 64, 5, 5, 16}, {64, 8, 1, 1, 1}};

constexpr int kC = 64;  // channels of the hidden conv planes

// Padded row length of the channels-last activation planes [B][T][Fp][64]: at least two zero
// pixels after every row so that a +-2 shift along F in the flattened pixel index reads zeros
// (the reference's ZeroPad2d), rounded to 8 pixels for alignment.
__host__ __device__ inline int padded_freq(int F) { return ((F + 2 + 7) / 8) * 8; }

// ---- activations -------------------------------------------------------------------------------
// Mish: x * tanh(softplus(x)), softplus threshold 20 (reference utils/generic_utils.py:399).
// tanh(log(1+e)) = ((1+e)^2 - 1) / ((1+e)^2 + 1) = (e^2 + 2e) / (e^2 + 2e + 2); one exp, one divide.
// Branch-free, 9 instructions (FMNMX, FMUL, EX2, FADD, FMUL, FADD, RCP, FMUL, FMUL), raw MUFU approximations (each ~1 ulp):
// e is taken at min(x, 30), where n / (n + 2) is already exactly 1.0f (it is for x > 8.7), so large x return x without the
// e^2 overflow; no range / denormal fix-up code, no divergent branch - the fused epilogues interleave many of these.
__device__ __forceinline__ float mish_f(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(x, 30.f) * 1.4426950408889634f));
    const float n = e * (e + 2.f);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(n + 2.f));
    return x * (n * r);
}
__device__ __forceinline__ float mish_precise(float x) {
    if (x > 20.f) return x;
    float e = expf(x);
    float n = e * (e + 2.f);
    return x * (n / (n + 2.f));
}
template <int ACT>
__device__ __forceinline__ float activate(float x) {
    if (ACT == 2) return x;                       // VS_ACT_NONE (train.cuh): raw pre-BatchNorm output
    if (ACT == VS_ACT_RELU) return fmaxf(x, 0.f);
    return mish_precise(x);
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

typedef unsigned short elt16;  // raw bits of a bf16 or fp16 value
// 16-bit element conversion: ELT 0 = bf16, 1 = fp16 (clamped: half overflows at 65504)
template <int ELT>
__device__ __forceinline__ void split16(float y, elt16& hi, elt16& lo) {
    if (ELT == 0) {
        __nv_bfloat16 h = __float2bfloat16(y);
        hi = __bfloat16_as_ushort(h);
        lo = __bfloat16_as_ushort(__float2bfloat16(y - __bfloat162float(h)));
    } else {
        y = fminf(fmaxf(y, -60000.f), 60000.f);
        __half h = __float2half_rn(y);
        hi = __half_as_ushort(h);
        lo = __half_as_ushort(__float2half_rn(y - __half2float(h)));
    }
}
// two activations (bounded below: Mish / ReLU outputs) -> packed hi pair and packed lo pair, one pack instruction each
template <int ELT>
__device__ __forceinline__ void split16_pair(float v0, float v1, uint32_t& hi2, uint32_t& lo2) {
    if (ELT == 0) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
        hi2 = *reinterpret_cast<const uint32_t*>(&h);
        const float2 hf = __bfloat1622float2(h);
        const __nv_bfloat162 l = __floats2bfloat162_rn(v0 - hf.x, v1 - hf.y);
        lo2 = *reinterpret_cast<const uint32_t*>(&l);
    } else {
        v0 = fminf(v0, 60000.f);
        v1 = fminf(v1, 60000.f);
        const __half2 h = __floats2half2_rn(v0, v1);
        hi2 = *reinterpret_cast<const uint32_t*>(&h);
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
        lo2 = *reinterpret_cast<const uint32_t*>(&l);
    }
}
__device__ __forceinline__ void split16_rt(float y, int elt, elt16& hi, elt16& lo) {
    if (elt == 0) split16<0>(y, hi, lo); else split16<1>(y, hi, lo);
}
// ---- VS_PREC_FP16_F8C: the fp8 correction plane ("c8") --------------------------------------------
// An activation x is kept as hi = fp16(x) (the operand of the kind::f16 main pass) plus 128 bytes per pixel of e4m3
// correction operands, channels innermost:   [ l8 = e4m3(2^8 * (x - hi)) x 64 | x8 = e4m3(2^-2 * hi) x 64 ].
// Weights carry the matching row   [ e4m3(2^-8 * w_hi) x 64 | e4m3(2^2 * w_lo) x 64 ],   so ONE kind::f8f6f4 product over the
// 128-byte K axis is  x_lo * w_hi + x_hi * w_lo  - the two correction terms of the split product - with the power-of-two
// scales cancelling inside each product (the fp32 TMEM accumulator is shared with the unscaled main pass).
constexpr float kF8cLoScale = 256.f, kF8cHiScale = 0.25f;
__device__ __forceinline__ unsigned short e4m3x2(float a, float b) {   // low byte = e4m3(a), high byte = e4m3(b), saturating
    return (unsigned short)__nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
}
__device__ __forceinline__ float e4m3_to_float(unsigned int byte) {
    __half_raw h = __nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)(byte & 0xffu), __NV_E4M3);
    return __half2float(*reinterpret_cast<__half*>(&h));
}
// hi (fp16 bits) and the float residual of y
__device__ __forceinline__ void split_f8c(float y, elt16& hi, float& lo) {
    y = fminf(fmaxf(y, -60000.f), 60000.f);
    const __half h = __float2half_rn(y);
    hi = __half_as_ushort(h);
    lo = y - __half2float(h);
}
// The same for the two channels a lane stores together, straight to the stored words: hi2 = packed fp16 pair (one F2FP), l8 / x8 =
// the two e4m3 pairs of the c8 row.  Only the upper clamp: the values are activations (Mish >= -0.31, ReLU >= 0).
__device__ __forceinline__ void split_f8c_pair(float v0, float v1, uint32_t& hi2, unsigned short& l8, unsigned short& x8) {
    v0 = fminf(v0, 60000.f);
    v1 = fminf(v1, 60000.f);
    const __half2 h = __floats2half2_rn(v0, v1);
    hi2 = *reinterpret_cast<const uint32_t*>(&h);
    const float2 hf = __half22float2(h);
    l8 = e4m3x2(kF8cLoScale * (v0 - hf.x), kF8cLoScale * (v1 - hf.y));
    const __half2 q = __hmul2(h, __float2half2_rn(kF8cHiScale));      // x8 = e4m3(2^-2 * hi): one HMUL2 + one conversion for the pair
    x8 = (unsigned short)__nv_cvt_halfraw2_to_fp8x2(*reinterpret_cast<const __half2_raw*>(&q), __NV_SATFINITE, __NV_E4M3);
}
__device__ __forceinline__ float join16(elt16 hi, elt16 lo, int elt) {
    return elt == 0 ? __bfloat162float(__ushort_as_bfloat16(hi)) + __bfloat162float(__ushort_as_bfloat16(lo))
                    : __half2float(__ushort_as_half(hi)) + __half2float(__ushort_as_half(lo));
}


// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const std::string& msg);
#define VS_CUDA_TRY(expr)                                                                      \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            vs::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                          ":" + std::to_string(__LINE__) + ")");                               \
            return VS_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace vs

// ---- the engine --------------------------------------------------------------------------------
struct vs_engine {
    vs_dims d;
    int device = 0, num_sms = 0;
    bool loaded = false;
    int launches = 0;

    // packed parameters (device, owned)
    float* conv_w32[8] = {};    // fp32 [tap][ci][co]
    float* conv_scale[8] = {};  // gamma / sqrt(var + eps)
    float* conv_shift[8] = {};  // (bias - mean) * scale + beta
    float* wih_x = nullptr;     // fp32 [8H][8F]   both directions stacked, spectrogram columns
    float* wih_e = nullptr;     // fp32 [8H][E]    d-vector columns
    float* b_lstm = nullptr;    // fp32 [8H]       b_ih + b_hh
    float* whh = nullptr;       // fp32 [2][4H][H]
    float* fc1_w = nullptr, *fc1_b = nullptr, *fc2_w = nullptr, *fc2_b = nullptr;
    // training path: data-gradient conv weights [tap'][co][ci] (layers 1..6), raw per-channel vectors
    float* conv_wT32[8] = {};
    float* conv_bias[8] = {};   // conv bias (not folded)
    float* bn_gamma[8] = {};
    float* bn_beta[8] = {};
    float* ones64 = nullptr, *zeros64 = nullptr;
    // data-parallel hooks (train.cu): SyncBN statistics all-reduce and the mid-backward notification
    vs_stat_allreduce_fn sync_fn = nullptr;
    void* sync_user = nullptr;
    int sync_world = 1;
    vs_backward_hook_fn bwd_hook = nullptr;
    void* bwd_hook_user = nullptr;
    bool train_tc = true;       // training: forward and data-gradient convs on the tcgen05 conv kernel (else fp32 CUDA cores)

    // host staging for vs_forward_host (grow-only)
    void* stage = nullptr;
    size_t stage_bytes = 0;
    // pipelined host entry (vs_forward_host_submit / _wait): HostPipe in engine.cu
    void* pipe = nullptr;
    // STFT / iSTFT state (audio.cu)
    void* audio = nullptr;
    // differentiable iSTFT + Si-SNR state (loss.cu)
    void* loss = nullptr;
    // GE2E speaker encoder (encoder.cu)
    void* encoder = nullptr;

    // tensor-core path state (tc_*.cu)
    void* tc = nullptr;

    // optional per-kernel timing (vs_engine_set_profiling): events recorded after every launch
    bool profiling = false;
    void* prof = nullptr;
};

namespace vs {

// kernel ids reported by vs_profile_read
enum KernelId {
    KID_FRONT = 0, KID_CONV1 = 1 /* +layer-1 for layers 1..6 */, KID_POINT8 = 7, KID_EMB_BIAS = 8, KID_INPROJ = 9,
    KID_LSTM_REC = 10, KID_FC1 = 11, KID_FC2 = 12, KID_CONVERT = 13, KID_HEAD = 14,
    // training path
    KID_TR_CONV_FWD = 20, KID_TR_BN_STATS = 21, KID_TR_BN_ACT = 22, KID_TR_BN_BWD = 23, KID_TR_WGRAD = 24, KID_TR_DGRAD = 25,
    KID_TR_GEMM = 26, KID_TR_LSTM_BWD = 27, KID_TR_MISC = 28
};
void prof_begin(vs_engine* e, cudaStream_t st);
void prof_after(vs_engine* e, int id, cudaStream_t st);
#define VS_LAUNCH(e, id, st, call)                                                             \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess) {                                                               \
            vs::set_error(std::string(#call) + ": " + cudaGetErrorString(_e));                 \
            return VS_ERR_CUDA;                                                                \
        }                                                                                      \
        (e)->launches++;                                                                       \
        if ((e)->profiling) vs::prof_after((e), (id), (st));                                   \
    } while (0)

// training path (train.cu): raw per-channel vectors + data-gradient weights, refreshed by vs_engine_load_params
int train_pack(vs_engine* e, const vs_params* p, cudaStream_t st);
void train_free(vs_engine* e);
void audio_free(vs_engine* e);
void loss_free(vs_engine* e);
void encoder_free(vs_engine* e);

// fp32 kernels (fp32_kernels.cu); all launch on `st` and return a cudaError_t
cudaError_t launch_front_fp32(const vs_engine* e, const float* x, float* plane, int B, int T, cudaStream_t st);
cudaError_t launch_conv_fp32(const vs_engine* e, int layer, const float* in, float* out, int B, int T, cudaStream_t st);
cudaError_t launch_point8_fp32(const vs_engine* e, const float* plane, float* xcat, int B, int T, cudaStream_t st);
// C[M][N] = op(A[M][K]) * W[N][K]^T + bias ; see fp32_kernels.cu
enum GemmEpi { EPI_NONE = 0, EPI_RELU = 1, EPI_SIGMOID_MASK = 2 };
cudaError_t launch_gemm_fp32(const float* A, int lda, const float* W, int ldw, const float* bias,
                             const float* bias_group, int group_rows, float* C, int ldc, int M, int N, int K,
                             bool relu_a, GemmEpi epi, const float* xmul, float* masked, cudaStream_t st);
// hr_hi/hr_lo (optional): relu(h) as 16-bit hi/lo planes [B*T][2H] - the fc1 operand of the tensor-core head
// gates_save/cseq (optional, training): gate activations i,f,g,o written over gates_x in place, cell states [B*T][2H]
cudaError_t launch_lstm_rec_fp32(const vs_engine* e, const float* gates_x, float* hout, float* hx,
                                 unsigned int* barrier, int B, int T, cudaStream_t st, elt16* hr_hi = nullptr,
                                 elt16* hr_lo = nullptr, int elt = 0, float* gates_save = nullptr, float* cseq = nullptr);
size_t lstm_rec_scratch_bytes(const vs_engine* e, int B);
// layout converters for the debug hooks
cudaError_t launch_nchw_to_plane(const float* nchw, float* plane, int B, int C, int T, int F, cudaStream_t st);
cudaError_t launch_plane_to_nchw(const float* plane, float* nchw, int B, int C, int T, int F, cudaStream_t st);

}  // namespace vs
