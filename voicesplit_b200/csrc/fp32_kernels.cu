// FP32 (CUDA-core FFMA) kernels of the mask path: the VS_PREC_FP32 arithmetic and the on-GPU
// ground truth the tensor-core kernels are validated against at full size.
//
// Activation planes are channels-last [B][T][Fp][64] fp32 with Fp = padded_freq(F); pixels
// f in [F, Fp) are kept at zero by every producer, which is the reference's ZeroPad2d along F
// (models/voicesplit/model.py:16-47) in the flattened pixel index.
#include "common.cuh"
#include "train.cuh"

namespace vs {

// ---------------------------------------------------------------------------------------------
// cnn1: ZeroPad2d((3,3,0,0)) + Conv2d(1,64,(1,7)) + BN + act   (model.py:17-19)
// ---------------------------------------------------------------------------------------------
template <int ACT>
__global__ void __launch_bounds__(256) k_front_fp32(const float* __restrict__ x, float* __restrict__ plane,
                                                    const float* __restrict__ w /*[7][64]*/,
                                                    const float* __restrict__ scale, const float* __restrict__ shift,
                                                    int T, int F, int Fp) {
    __shared__ float xs[32 + 6];
    __shared__ float ws[7 * 64];
    __shared__ float sc[64], sh[64];
    const int f0 = blockIdx.x * 32, t = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x;
    const float* xrow = x + ((size_t)b * T + t) * F;
    if (tid < 38) {
        int f = f0 + tid - 3;
        xs[tid] = (f >= 0 && f < F) ? xrow[f] : 0.f;
    }
    for (int i = tid; i < 7 * 64; i += 256) ws[i] = w[i];
    if (tid < 64) { sc[tid] = scale[tid]; sh[tid] = shift[tid]; }
    __syncthreads();
    const int px = tid >> 3, cg = tid & 7;
    const int f = f0 + px;
    if (f >= Fp) return;
    float out[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        int co = cg * 8 + c;
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 7; ++j) a = fmaf(ws[j * 64 + co], xs[px + j], a);
        out[c] = (f < F) ? activate<ACT>(fmaf(a, sc[co], sh[co])) : 0.f;
    }
    float4* dst = reinterpret_cast<float4*>(plane + (((size_t)b * T + t) * Fp + f) * 64 + cg * 8);
    dst[0] = make_float4(out[0], out[1], out[2], out[3]);
    dst[1] = make_float4(out[4], out[5], out[6], out[7]);
}

cudaError_t launch_front_fp32_ex(const vs_engine* e, const float* x, float* plane, const float* w, const float* scale, const float* shift,
                                 int act, int B, int T, cudaStream_t st) {
    const int F = e->d.num_freq, Fp = padded_freq(F);
    dim3 grid((Fp + 31) / 32, T, B);
    if (act == VS_ACT_RELU) k_front_fp32<VS_ACT_RELU><<<grid, 256, 0, st>>>(x, plane, w, scale, shift, T, F, Fp);
    else if (act == VS_ACT_NONE) k_front_fp32<VS_ACT_NONE><<<grid, 256, 0, st>>>(x, plane, w, scale, shift, T, F, Fp);
    else k_front_fp32<VS_ACT_MISH><<<grid, 256, 0, st>>>(x, plane, w, scale, shift, T, F, Fp);
    return cudaGetLastError();
}

cudaError_t launch_front_fp32(const vs_engine* e, const float* x, float* plane, int B, int T, cudaStream_t st) {
    const int F = e->d.num_freq, Fp = padded_freq(F);
    dim3 grid((Fp + 31) / 32, T, B);
    if (e->d.activation == VS_ACT_RELU)
        k_front_fp32<VS_ACT_RELU><<<grid, 256, 0, st>>>(x, plane, e->conv_w32[0], e->conv_scale[0], e->conv_shift[0], T, F, Fp);
    else
        k_front_fp32<VS_ACT_MISH><<<grid, 256, 0, st>>>(x, plane, e->conv_w32[0], e->conv_scale[0], e->conv_shift[0], T, F, Fp);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// cnn2..cnn7: "same" zero pad + Conv2d(64,64,(kh,kw),dilation=(dil,1)) + BN + act  (model.py:21-48)
// Tile: 64 pixels along F x 64 output channels per block; thread = 4 pixels x 4 channels.
// ---------------------------------------------------------------------------------------------
template <int ACT>
__global__ void __launch_bounds__(256) k_conv_fp32(const float* __restrict__ in, float* __restrict__ out,
                                                   const float* __restrict__ w /*[kh*kw][64][64]*/,
                                                   const float* __restrict__ scale, const float* __restrict__ shift,
                                                   int T, int F, int Fp, int kh, int kw, int dil) {
    __shared__ __align__(16) float strip[(64 + 6) * 64];
    __shared__ __align__(16) float wt[64 * 64];
    const int f0 = blockIdx.x * 64, t = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int hw = kw / 2, nstrip = 64 + kw - 1;
    float acc[4][4];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[p][c] = 0.f;

    for (int i = 0; i < kh; ++i) {
        const int tr = t + (i - kh / 2) * dil;
        if (tr < 0 || tr >= T) continue;  // uniform per block: ZeroPad2d rows
        const float* row = in + ((size_t)b * T + tr) * Fp * 64;
        __syncthreads();
        for (int idx = tid; idx < nstrip * 16; idx += 256) {
            int px = idx >> 4, q = idx & 15;
            int f = f0 + px - hw;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (f >= 0 && f < Fp) v = *reinterpret_cast<const float4*>(row + (size_t)f * 64 + q * 4);
            *reinterpret_cast<float4*>(strip + px * 64 + q * 4) = v;
        }
        for (int j = 0; j < kw; ++j) {
            __syncthreads();
            const float* wsrc = w + (size_t)(i * kw + j) * 64 * 64;
            for (int idx = tid; idx < 64 * 16; idx += 256)
                *reinterpret_cast<float4*>(wt + idx * 4) = *reinterpret_cast<const float4*>(wsrc + idx * 4);
            __syncthreads();
#pragma unroll 8
            for (int ci = 0; ci < 64; ++ci) {
                float4 wv = *reinterpret_cast<const float4*>(wt + ci * 64 + tx * 4);
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    float a = strip[(ty * 4 + p + j) * 64 + ci];
                    acc[p][0] = fmaf(a, wv.x, acc[p][0]);
                    acc[p][1] = fmaf(a, wv.y, acc[p][1]);
                    acc[p][2] = fmaf(a, wv.z, acc[p][2]);
                    acc[p][3] = fmaf(a, wv.w, acc[p][3]);
                }
            }
        }
    }
    float4 sc = *reinterpret_cast<const float4*>(scale + tx * 4);
    float4 sh = *reinterpret_cast<const float4*>(shift + tx * 4);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        int f = f0 + ty * 4 + p;
        if (f >= Fp) continue;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (f < F) {
            o.x = activate<ACT>(fmaf(acc[p][0], sc.x, sh.x));
            o.y = activate<ACT>(fmaf(acc[p][1], sc.y, sh.y));
            o.z = activate<ACT>(fmaf(acc[p][2], sc.z, sh.z));
            o.w = activate<ACT>(fmaf(acc[p][3], sc.w, sh.w));
        }
        *reinterpret_cast<float4*>(out + (((size_t)b * T + t) * Fp + f) * 64 + tx * 4) = o;
    }
}

cudaError_t launch_conv_fp32_ex(const vs_engine* e, int layer, const float* in, float* out, const float* w, const float* scale,
                                const float* shift, int act, int B, int T, cudaStream_t st) {
    const int F = e->d.num_freq, Fp = padded_freq(F);
    const ConvGeom g = kConv[layer];
    dim3 grid((Fp + 63) / 64, T, B);
    if (act == VS_ACT_RELU) k_conv_fp32<VS_ACT_RELU><<<grid, 256, 0, st>>>(in, out, w, scale, shift, T, F, Fp, g.kh, g.kw, g.dil);
    else if (act == VS_ACT_NONE) k_conv_fp32<VS_ACT_NONE><<<grid, 256, 0, st>>>(in, out, w, scale, shift, T, F, Fp, g.kh, g.kw, g.dil);
    else k_conv_fp32<VS_ACT_MISH><<<grid, 256, 0, st>>>(in, out, w, scale, shift, T, F, Fp, g.kh, g.kw, g.dil);
    return cudaGetLastError();
}

cudaError_t launch_conv_fp32(const vs_engine* e, int layer, const float* in, float* out, int B, int T, cudaStream_t st) {
    const int F = e->d.num_freq, Fp = padded_freq(F);
    const ConvGeom g = kConv[layer];
    dim3 grid((Fp + 63) / 64, T, B);
    if (e->d.activation == VS_ACT_RELU)
        k_conv_fp32<VS_ACT_RELU><<<grid, 256, 0, st>>>(in, out, e->conv_w32[layer], e->conv_scale[layer], e->conv_shift[layer], T, F, Fp, g.kh, g.kw, g.dil);
    else
        k_conv_fp32<VS_ACT_MISH><<<grid, 256, 0, st>>>(in, out, e->conv_w32[layer], e->conv_scale[layer], e->conv_shift[layer], T, F, Fp, g.kh, g.kw, g.dil);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// cnn8 + reshape: Conv2d(64,8,1x1) + BN + act, written as the LSTM input row [B*T][8F] with
// column c*F+f (model.py:51-52,72-74); the d-vector concat is folded into a gate bias instead.
// ---------------------------------------------------------------------------------------------
template <int ACT>
__global__ void __launch_bounds__(256) k_point8_fp32(const float* __restrict__ plane, float* __restrict__ xcat,
                                                     const float* __restrict__ w /*[64][8]*/,
                                                     const float* __restrict__ scale, const float* __restrict__ shift,
                                                     int T, int F, int Fp, long long npix) {
    __shared__ float ws[64 * 8];
    __shared__ float sc[8], sh[8];
    for (int i = threadIdx.x; i < 512; i += 256) ws[i] = w[i];
    if (threadIdx.x < 8) { sc[threadIdx.x] = scale[threadIdx.x]; sh[threadIdx.x] = shift[threadIdx.x]; }
    __syncthreads();
    long long p = (long long)blockIdx.x * 256 + threadIdx.x;
    if (p >= npix) return;
    int f = (int)(p % F);
    long long bt = p / F;
    const float4* src = reinterpret_cast<const float4*>(plane + ((size_t)bt * Fp + f) * 64);
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
#pragma unroll 4
    for (int q = 0; q < 16; ++q) {
        float4 v = src[q];
        float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = fmaf(vv[k], ws[(q * 4 + k) * 8 + c], acc[c]);
    }
    float* dst = xcat + (size_t)bt * 8 * F + f;
#pragma unroll
    for (int c = 0; c < 8; ++c) dst[(size_t)c * F] = activate<ACT>(fmaf(acc[c], sc[c], sh[c]));
}

cudaError_t launch_point8_fp32_ex(const vs_engine* e, const float* plane, float* xcat, const float* w, const float* scale,
                                  const float* shift, int act, int B, int T, cudaStream_t st) {
    const int F = e->d.num_freq, Fp = padded_freq(F);
    long long npix = (long long)B * T * F;
    unsigned grid = (unsigned)((npix + 255) / 256);
    if (act == VS_ACT_RELU) k_point8_fp32<VS_ACT_RELU><<<grid, 256, 0, st>>>(plane, xcat, w, scale, shift, T, F, Fp, npix);
    else if (act == VS_ACT_NONE) k_point8_fp32<VS_ACT_NONE><<<grid, 256, 0, st>>>(plane, xcat, w, scale, shift, T, F, Fp, npix);
    else k_point8_fp32<VS_ACT_MISH><<<grid, 256, 0, st>>>(plane, xcat, w, scale, shift, T, F, Fp, npix);
    return cudaGetLastError();
}

cudaError_t launch_point8_fp32(const vs_engine* e, const float* plane, float* xcat, int B, int T, cudaStream_t st) {
    const int F = e->d.num_freq, Fp = padded_freq(F);
    long long npix = (long long)B * T * F;
    unsigned grid = (unsigned)((npix + 255) / 256);
    if (e->d.activation == VS_ACT_RELU)
        k_point8_fp32<VS_ACT_RELU><<<grid, 256, 0, st>>>(plane, xcat, e->conv_w32[7], e->conv_scale[7], e->conv_shift[7], T, F, Fp, npix);
    else
        k_point8_fp32<VS_ACT_MISH><<<grid, 256, 0, st>>>(plane, xcat, e->conv_w32[7], e->conv_scale[7], e->conv_shift[7], T, F, Fp, npix);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Generic fp32 GEMM  C[M][N] = op(A)[M][K] * W[N][K]^T + bias[n] + bias_group[m / group_rows][n]
// used for the LSTM input projection, the d-vector gate bias, fc1 and fc2 (+sigmoid, mask apply).
// ---------------------------------------------------------------------------------------------
template <int EPI, bool RELU_A>
__global__ void __launch_bounds__(256) k_gemm_fp32(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                                                   const float* __restrict__ bias, const float* __restrict__ bias_group,
                                                   int group_rows, float* __restrict__ C, int ldc, int M, int N, int K,
                                                   const float* __restrict__ xmul, float* __restrict__ masked) {
    __shared__ float As[16][64 + 4];
    __shared__ float Ws[16][64 + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int lr = tid >> 2, lk = (tid & 3) * 4;  // each thread stages 4 consecutive k of one row
    for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int k = k0 + lk + q;
            int m = m0 + lr, n = n0 + lr;
            float a = (m < M && k < K) ? A[(size_t)m * lda + k] : 0.f;
            if (RELU_A) a = fmaxf(a, 0.f);
            float w = (n < N && k < K) ? W[(size_t)n * ldw + k] : 0.f;
            As[lk + q][lr] = a;
            Ws[lk + q][lr] = w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            float4 wv = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
            float a4[4] = {av.x, av.y, av.z, av.w}, w4[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], w4[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (bias) v += bias[n];
            if (bias_group) v += bias_group[(size_t)(m / group_rows) * N + n];
            if (EPI == EPI_RELU) v = fmaxf(v, 0.f);
            if (EPI == EPI_SIGMOID_MASK) {
                v = sigmoid_f(v);
                if (masked) masked[(size_t)m * ldc + n] = xmul[(size_t)m * ldc + n] * v;
            }
            C[(size_t)m * ldc + n] = v;
        }
    }
}

cudaError_t launch_gemm_fp32(const float* A, int lda, const float* W, int ldw, const float* bias,
                             const float* bias_group, int group_rows, float* C, int ldc, int M, int N, int K,
                             bool relu_a, GemmEpi epi, const float* xmul, float* masked, cudaStream_t st) {
    dim3 grid((N + 63) / 64, (M + 63) / 64);
    if (group_rows <= 0) group_rows = 1;
#define VS_GEMM(E, R) k_gemm_fp32<E, R><<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, bias_group, group_rows, C, ldc, M, N, K, xmul, masked)
    if (epi == EPI_NONE && !relu_a) VS_GEMM(EPI_NONE, false);
    else if (epi == EPI_NONE && relu_a) VS_GEMM(EPI_NONE, true);
    else if (epi == EPI_RELU && !relu_a) VS_GEMM(EPI_RELU, false);
    else if (epi == EPI_RELU && relu_a) VS_GEMM(EPI_RELU, true);
    else if (epi == EPI_SIGMOID_MASK && !relu_a) VS_GEMM(EPI_SIGMOID_MASK, false);
    else VS_GEMM(EPI_SIGMOID_MASK, true);
#undef VS_GEMM
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// BiLSTM recurrence (model.py:57-61,82): persistent kernel, one CTA per (direction, slice of 8
// hidden units).  The CTA keeps its 32 rows of W_hh (4 gates x 8 units) in shared memory for the
// whole sequence, each thread owns one hidden unit for two utterances (all four gates, so the
// cell update is thread-local), and the CTAs of one direction exchange h through a transposed
// global buffer hx[dir][parity][H][Bp] with a per-direction arrive/spin barrier every step.
// gates_x [B*T][8H] already holds W_ih x + W_ih_e emb + b_ih + b_hh (direction d at column d*4H).
// ---------------------------------------------------------------------------------------------
constexpr int kHS = 8;    // hidden units per CTA
constexpr int kBT = 64;   // utterances per batch tile

__device__ __forceinline__ void dir_barrier(unsigned int* counter, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned int spins = 0;
        while (*reinterpret_cast<volatile unsigned int*>(counter) < target) {
            if (++spins > (1u << 30)) __trap();  // never hang the GPU: a lost CTA is a bug
        }
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256, 1) k_lstm_rec_fp32(const float* gates_x /* may alias gates_save */, const float* __restrict__ whh,
                                                          float* __restrict__ hout, float* hx, float* cstate,
                                                          unsigned int* barrier, unsigned int barrier_base,
                                                          int B, int Bp, int T, int H, int nslices,
                                                          elt16* __restrict__ hr_hi, elt16* __restrict__ hr_lo, int elt,
                                                          float* gates_save, float* cseq) {
    extern __shared__ __align__(16) float smem[];
    float* wt = smem;                  // [H][kHS][4]   (k, unit, gate)
    float* ht = smem + (size_t)H * 32; // [H][kBT]
    const int d = blockIdx.x / nslices, sl = blockIdx.x % nslices;
    const int tid = threadIdx.x, j = tid & 7, bp = tid >> 3;  // unit in slice, utterance pair in tile
    const int hj = sl * kHS + j;
    const bool unit_ok = hj < H;
    // stage this CTA's W_hh rows: wt[k][j][g] = whh[d][g*H + hj][k]
    for (int idx = tid; idx < H * 32; idx += 256) {
        int k = idx >> 5, r = idx & 31, jj = r >> 2, g = r & 3;
        int row = g * H + sl * kHS + jj;
        wt[idx] = (sl * kHS + jj < H) ? whh[((size_t)d * 4 * H + row) * H + k] : 0.f;
    }
    __syncthreads();
    const int ntile = (B + kBT - 1) / kBT;
    for (int s = 0; s < T; ++s) {
        const int t = d ? T - 1 - s : s;
        const int par = s & 1;
        const float* hprev = hx + ((size_t)(d * 2 + par) * H) * Bp;
        float* hnext = hx + ((size_t)(d * 2 + (par ^ 1)) * H) * Bp;
        for (int bt = 0; bt < ntile; ++bt) {
            const int b0 = bt * kBT;
            if (s > 0) {
                __syncthreads();
                for (int idx = tid; idx < H * (kBT / 4); idx += 256) {
                    int k = idx / (kBT / 4), q = idx % (kBT / 4);
                    float4 v = __ldcg(reinterpret_cast<const float4*>(hprev + (size_t)k * Bp + b0 + q * 4));
                    *reinterpret_cast<float4*>(ht + k * kBT + q * 4) = v;
                }
                __syncthreads();
            }
            const int b_a = b0 + bp * 2, b_b = b_a + 1;
            float acc[2][4];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                int b = b_a + u;
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    acc[u][g] = (unit_ok && b < B) ? gates_x[((size_t)b * T + t) * 8 * H + (size_t)d * 4 * H + g * H + hj] : 0.f;
            }
            if (s > 0) {
#pragma unroll 4
                for (int k = 0; k < H; ++k) {
                    float4 w = *reinterpret_cast<const float4*>(wt + k * 32 + j * 4);
                    float2 h2 = *reinterpret_cast<const float2*>(ht + k * kBT + bp * 2);
                    acc[0][0] = fmaf(w.x, h2.x, acc[0][0]); acc[0][1] = fmaf(w.y, h2.x, acc[0][1]);
                    acc[0][2] = fmaf(w.z, h2.x, acc[0][2]); acc[0][3] = fmaf(w.w, h2.x, acc[0][3]);
                    acc[1][0] = fmaf(w.x, h2.y, acc[1][0]); acc[1][1] = fmaf(w.y, h2.y, acc[1][1]);
                    acc[1][2] = fmaf(w.z, h2.y, acc[1][2]); acc[1][3] = fmaf(w.w, h2.y, acc[1][3]);
                }
            }
            if (unit_ok) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    int b = u ? b_b : b_a;
                    if (b >= B) continue;
                    float* cp = cstate + ((size_t)d * H + hj) * Bp + b;
                    float c_old = s > 0 ? *cp : 0.f;
                    float ig = sigmoid_f(acc[u][0]), fg = sigmoid_f(acc[u][1]);
                    float gg = tanhf(acc[u][2]), og = sigmoid_f(acc[u][3]);
                    float c_new = fmaf(fg, c_old, ig * gg);
                    float h_new = og * tanhf(c_new);
                    *cp = c_new;
                    hnext[(size_t)hj * Bp + b] = h_new;
                    if (gates_save) {   // training: keep what the backward recurrence needs
                        const size_t gi = ((size_t)b * T + t) * 8 * H + (size_t)d * 4 * H + hj;
                        gates_save[gi] = ig; gates_save[gi + H] = fg; gates_save[gi + 2 * H] = gg; gates_save[gi + 3 * H] = og;
                        cseq[((size_t)b * T + t) * 2 * H + (size_t)d * H + hj] = c_new;
                    }
                    const size_t oidx = ((size_t)b * T + t) * 2 * H + (size_t)d * H + hj;
                    hout[oidx] = h_new;
                    if (hr_hi) {
                        elt16 vh, vl;
                        split16_rt(fmaxf(h_new, 0.f), elt, vh, vl);
                        hr_hi[oidx] = vh;
                        if (hr_lo) hr_lo[oidx] = vl;
                    }
                }
            }
        }
        if (s + 1 < T) dir_barrier(barrier + d, barrier_base + (unsigned int)(s + 1) * nslices);
    }
}

size_t lstm_rec_scratch_bytes(const vs_engine* e, int B) {
    const int H = e->d.lstm_dim;
    const size_t Bp = align_up((size_t)B, kBT);
    // hx [2][2][H][Bp] + cstate [2][H][Bp]
    return (size_t)(4 + 2) * H * Bp * sizeof(float);
}


cudaError_t launch_lstm_rec_fp32(const vs_engine* e, const float* gates_x, float* hout, float* hx,
                                 unsigned int* barrier, int B, int T, cudaStream_t st, elt16* hr_hi, elt16* hr_lo, int elt,
                                 float* gates_save, float* cseq) {
    const int H = e->d.lstm_dim;
    const int nslices = (H + kHS - 1) / kHS;
    const int Bp = (int)align_up((size_t)B, kBT);
    if (2 * nslices > e->num_sms) return cudaErrorInvalidConfiguration;
    size_t smem = (size_t)H * (32 + kBT) * sizeof(float);
    cudaError_t err = cudaFuncSetAttribute(k_lstm_rec_fp32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    float* cstate = hx + (size_t)4 * H * Bp;
    // barrier counters are zeroed per launch (stream ordered) so the base is always 0
    err = cudaMemsetAsync(barrier, 0, 2 * sizeof(unsigned int), st);
    if (err != cudaSuccess) return err;
    unsigned int base = 0;
    const float* whh = e->whh;
    int Bv = B, Bpv = Bp, Tv = T, Hv = H, ns = nslices;
    void* args[] = {(void*)&gates_x, (void*)&whh, (void*)&hout, (void*)&hx, (void*)&cstate, (void*)&barrier,
                    (void*)&base, (void*)&Bv, (void*)&Bpv, (void*)&Tv, (void*)&Hv, (void*)&ns,
                    (void*)&hr_hi, (void*)&hr_lo, (void*)&elt, (void*)&gates_save, (void*)&cseq};
    return cudaLaunchCooperativeKernel((const void*)k_lstm_rec_fp32, dim3(2 * nslices), dim3(256), args, smem, st);
}

// ---------------------------------------------------------------------------------------------
// layout converters for the debug hooks
// ---------------------------------------------------------------------------------------------
__global__ void k_nchw_to_plane(const float* __restrict__ nchw, float* __restrict__ plane, int C, int T, int F, int Fp, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;  // n = B*T*Fp*64
    int c = (int)(i & 63);
    long long p = i >> 6;
    int f = (int)(p % Fp);
    long long bt = p / Fp;
    int t = (int)(bt % T);
    long long b = bt / T;
    float v = 0.f;
    if (f < F && c < C) v = nchw[(((size_t)b * C + c) * T + t) * F + f];
    plane[i] = v;
}
__global__ void k_plane_to_nchw(const float* __restrict__ plane, float* __restrict__ nchw, int C, int T, int F, int Fp, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;  // n = B*C*T*F
    int f = (int)(i % F);
    long long r = i / F;
    int t = (int)(r % T);
    r /= T;
    int c = (int)(r % C);
    long long b = r / C;
    nchw[i] = plane[(((size_t)b * T + t) * Fp + f) * 64 + c];
}
cudaError_t launch_nchw_to_plane(const float* nchw, float* plane, int B, int C, int T, int F, cudaStream_t st) {
    int Fp = padded_freq(F);
    long long n = (long long)B * T * Fp * 64;
    k_nchw_to_plane<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(nchw, plane, C, T, F, Fp, n);
    return cudaGetLastError();
}
cudaError_t launch_plane_to_nchw(const float* plane, float* nchw, int B, int C, int T, int F, cudaStream_t st) {
    int Fp = padded_freq(F);
    long long n = (long long)B * C * T * F;
    k_plane_to_nchw<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(plane, nchw, C, T, F, Fp, n);
    return cudaGetLastError();
}

}  // namespace vs
