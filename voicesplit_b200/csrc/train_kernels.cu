// Training-mode kernels (fp32, CUDA cores): BatchNorm with batch statistics, the backward of the
// conv / BN / activation stack, the LSTM backward recurrence and a strided GEMM for the weight and
// input gradients.  Reference semantics: torch.nn layers of models/voicesplit/model.py:15-64 under
// autograd, as driven by train.py:94-111.
#include "train.cuh"

namespace vs {

// ---------------------------------------------------------------------------------------------
// activation derivative w.r.t. the pre-activation u
// ---------------------------------------------------------------------------------------------
template <int ACT>
__device__ __forceinline__ float act_grad(float u) {
    if (ACT == VS_ACT_RELU) return u > 0.f ? 1.f : 0.f;
    // mish(u) = u * tanh(sp), sp = softplus(u) (identity above 20, utils/generic_utils.py:399 / F.softplus):
    // d/du = t + u * (1 - t^2) * sigmoid(u).  With e = exp(u), n = (1+e)^2 - 1 = e(e+2):
    // t = tanh(log(1+e)) = n / (n+2),  1 - t^2 = 4(n+1) / (n+2)^2,  sigmoid(u) = e / (1+e)  -> one exp, two divides
    if (u > 20.f) return 1.f;            // t == 1 and 1 - t^2 == 0 in fp32 (and avoids e^2 overflow)
    const float e = __expf(u);
    const float n = e * (e + 2.f);
    const float inv = __fdividef(1.f, n + 2.f);
    const float t = n * inv;
    const float sech2 = 4.f * (n + 1.f) * inv * inv;
    return fmaf(u * sech2, __fdividef(e, 1.f + e), t);
}

// ---------------------------------------------------------------------------------------------
// BatchNorm statistics.  "plane" layout: [rows = B*T][Fp][C] channels-last, valid pixels f < F.
// "cols" layout (cnn8 output): [rows][C*F], channel = column / F.
// sums[0][c] = sum z, sums[1][c] = sum z^2 (double, atomically accumulated)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bn_stats_plane(const float* __restrict__ z, double* __restrict__ sums, int F, int Fp, long long nrows) {
    const int c = threadIdx.x & 63, lane4 = threadIdx.x >> 6;
    double s = 0.0, q = 0.0;
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
        const float* zr = z + (size_t)row * Fp * 64;
        float fs = 0.f, fq = 0.f;
        for (int f = lane4; f < F; f += 4) {
            float v = zr[(size_t)f * 64 + c];
            fs += v; fq = fmaf(v, v, fq);
        }
        s += fs; q += fq;
    }
    __shared__ double sh[2][256];
    sh[0][threadIdx.x] = s; sh[1][threadIdx.x] = q;
    __syncthreads();
    if (threadIdx.x < 64) {
        double a = sh[0][c] + sh[0][c + 64] + sh[0][c + 128] + sh[0][c + 192];
        double b = sh[1][c] + sh[1][c + 64] + sh[1][c + 128] + sh[1][c + 192];
        atomicAdd(&sums[c], a);
        atomicAdd(&sums[64 + c], b);
    }
}
__global__ void __launch_bounds__(256) k_bn_stats_cols(const float* __restrict__ z, double* __restrict__ sums, int C, int F, long long nrows) {
    // grid.y = channel; threads stride over (row, f)
    const int c = blockIdx.y;
    double s = 0.0, q = 0.0;
    const long long n = nrows * F;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        long long row = i / F; int f = (int)(i - row * F);
        float v = z[(size_t)row * C * F + (size_t)c * F + f];
        s += v; q += (double)v * v;
    }
    __shared__ double sh[2][256];
    sh[0][threadIdx.x] = s; sh[1][threadIdx.x] = q;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { atomicAdd(&sums[c], sh[0][0]); atomicAdd(&sums[64 + c], sh[1][0]); }
}
// mean / rstd / folded scale+shift from the sums; running-stat update (momentum, unbiased variance),
// exactly nn.BatchNorm2d in training mode (eps 1e-5)
__global__ void k_bn_finalize(const double* __restrict__ sums, double count, const float* __restrict__ gamma, const float* __restrict__ beta,
                              float* __restrict__ stat /*[4][64]: mean, rstd, scale, shift*/, float* running_mean, float* running_var,
                              long long* num_batches, float momentum, int C) {
    int c = threadIdx.x;
    if (c >= C) return;
    double mean = sums[c] / count;
    double var = sums[64 + c] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    float rstd = (float)(1.0 / sqrt(var + 1e-5));
    stat[c] = (float)mean; stat[64 + c] = rstd;
    float sc = gamma[c] * rstd;
    stat[128 + c] = sc; stat[192 + c] = beta[c] - (float)mean * sc;
    if (running_mean) {
        double unb = count > 1.0 ? var * count / (count - 1.0) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
        if (c == 0 && num_batches) *num_batches += 1;
    }
}

// a = act(z * scale + shift), pads zeroed (plane) / cols layout
template <int ACT>
__global__ void __launch_bounds__(256) k_bn_act_plane(const float* __restrict__ z, float* __restrict__ a, const float* __restrict__ stat, int F, int Fp, long long n4,
                                                      elt16* __restrict__ ahi, elt16* __restrict__ alo, int elt) {
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;   // float4 index
    if (i >= n4) return;
    int c4 = (int)(i & 15) * 4;
    long long pix = i >> 4;
    int f = (int)(pix % Fp);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (f < F) {
        float4 v = reinterpret_cast<const float4*>(z)[i];
        const float* sc = stat + 128 + c4; const float* sh = stat + 192 + c4;
        o.x = activate<ACT>(fmaf(v.x, sc[0], sh[0])); o.y = activate<ACT>(fmaf(v.y, sc[1], sh[1]));
        o.z = activate<ACT>(fmaf(v.z, sc[2], sh[2])); o.w = activate<ACT>(fmaf(v.w, sc[3], sh[3]));
    }
    if (a) reinterpret_cast<float4*>(a)[i] = o;
    if (ahi) {   // also as 16-bit hi/lo planes: the operand of the tensor-core conv
        __align__(8) elt16 h[4], l[4];
        split16_rt(o.x, elt, h[0], l[0]); split16_rt(o.y, elt, h[1], l[1]);
        split16_rt(o.z, elt, h[2], l[2]); split16_rt(o.w, elt, h[3], l[3]);
        reinterpret_cast<uint2*>(ahi)[i] = *reinterpret_cast<const uint2*>(h);
        reinterpret_cast<uint2*>(alo)[i] = *reinterpret_cast<const uint2*>(l);
    }
}
template <int ACT>
__global__ void __launch_bounds__(256) k_bn_act_cols(const float* __restrict__ z, float* __restrict__ a, const float* __restrict__ stat, int C, int F, long long n) {
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    int c = (int)((i % ((long long)C * F)) / F);
    a[i] = activate<ACT>(fmaf(z[i], stat[128 + c], stat[192 + c]));
}

// ---------------------------------------------------------------------------------------------
// BN + activation backward.  Pass 1: S1[c] = sum du, S2[c] = sum du * xhat with du = da * act'(u).
// Pass 2: dz = gamma * rstd * (du - S1/N - xhat * S2/N); also accumulates sum dz (conv bias grad).
// ---------------------------------------------------------------------------------------------
template <int ACT>
__global__ void __launch_bounds__(256) k_bn_bwd_reduce_plane(const float* __restrict__ da, const float* __restrict__ z, const float* __restrict__ stat,
                                                             double* __restrict__ sums, int F, int Fp, long long nrows) {
    // thread = 4 channels (float4) x every 16th pixel of a row; block-level reduction, then one double atomic per channel
    const int c4 = (threadIdx.x & 15) * 4, plane_lane = threadIdx.x >> 4;
    float mean[4], rstd[4], sc[4], sh[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { mean[k] = stat[c4 + k]; rstd[k] = stat[64 + c4 + k]; sc[k] = stat[128 + c4 + k]; sh[k] = stat[192 + c4 + k]; }
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    double d1[4] = {0, 0, 0, 0}, d2[4] = {0, 0, 0, 0};
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
        const size_t base = (size_t)row * Fp * 64;
#pragma unroll 4
        for (int f = plane_lane; f < F; f += 16) {
            const float4 zv = *reinterpret_cast<const float4*>(z + base + (size_t)f * 64 + c4);
            const float4 dv = *reinterpret_cast<const float4*>(da + base + (size_t)f * 64 + c4);
            const float zz[4] = {zv.x, zv.y, zv.z, zv.w}, dd[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float du = dd[k] * act_grad<ACT>(fmaf(zz[k], sc[k], sh[k]));
                s1[k] += du; s2[k] = fmaf(du, (zz[k] - mean[k]) * rstd[k], s2[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) { d1[k] += s1[k]; d2[k] += s2[k]; s1[k] = 0.f; s2[k] = 0.f; }
    }
    __shared__ double shm[2][16][64];
#pragma unroll
    for (int k = 0; k < 4; ++k) { shm[0][plane_lane][c4 + k] = d1[k]; shm[1][plane_lane][c4 + k] = d2[k]; }
    __syncthreads();
    if (threadIdx.x < 128) {
        const int which = threadIdx.x >> 6, c = threadIdx.x & 63;
        double t = 0.0;
#pragma unroll
        for (int l = 0; l < 16; ++l) t += shm[which][l][c];
        atomicAdd(&sums[which * 64 + c], t);
    }
}
template <int ACT>
__global__ void __launch_bounds__(256) k_bn_bwd_apply_plane(const float* __restrict__ da, const float* __restrict__ z, const float* __restrict__ stat,
                                                            const float* __restrict__ gamma, const double* __restrict__ sums, double count,
                                                            float* __restrict__ dz, int F, int Fp, long long npix,
                                                            elt16* __restrict__ dhi, elt16* __restrict__ dlo) {
    // per-channel constants once per block (the double divisions used to run per element): m1 = S1/N, m2 = S2/N, g = gamma rstd
    __shared__ float cm1[64], cm2[64], cg[64];
    if (threadIdx.x < 64) {
        const int c = threadIdx.x;
        cm1[c] = (float)(sums[c] / count); cm2[c] = (float)(sums[64 + c] / count); cg[c] = gamma[c] * stat[64 + c];
    }
    __syncthreads();
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;   // float4 index over [pixels][16]
    if (i >= npix * 16) return;
    const int c4 = (int)(i & 15) * 4;
    const int f = (int)((i >> 4) % Fp);
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    if (f < F) {
        const float4 zv = reinterpret_cast<const float4*>(z)[i];
        const float4 dv = reinterpret_cast<const float4*>(da)[i];
        const float zz[4] = {zv.x, zv.y, zv.z, zv.w}, dd[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = c4 + k;
            const float mean = stat[c], rstd = stat[64 + c];
            const float du = dd[k] * act_grad<ACT>(fmaf(zz[k], stat[128 + c], stat[192 + c]));
            const float xh = (zz[k] - mean) * rstd;
            o[k] = cg[c] * (du - cm1[c] - xh * cm2[c]);
        }
    }
    if (dz) reinterpret_cast<float4*>(dz)[i] = make_float4(o[0], o[1], o[2], o[3]);
    if (dhi) {   // bf16 hi/lo (fp32 exponent range: gradients need no loss scaling) for the tensor-core data gradient
        __align__(8) elt16 h[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) split16<0>(o[k], h[k], l[k]);
        reinterpret_cast<uint2*>(dhi)[i] = *reinterpret_cast<const uint2*>(h);
        reinterpret_cast<uint2*>(dlo)[i] = *reinterpret_cast<const uint2*>(l);
    }
}
template <int ACT>
__global__ void __launch_bounds__(256) k_bn_bwd_reduce_cols(const float* __restrict__ da, const float* __restrict__ z, const float* __restrict__ stat,
                                                            double* __restrict__ sums, int C, int F, long long nrows) {
    const int c = blockIdx.y;
    const float mean = stat[c], rstd = stat[64 + c], sc = stat[128 + c], sh = stat[192 + c];
    double s1 = 0.0, s2 = 0.0;
    const long long n = nrows * F;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        long long row = i / F; int f = (int)(i - row * F);
        size_t o = (size_t)row * C * F + (size_t)c * F + f;
        float zv = z[o];
        float du = da[o] * act_grad<ACT>(fmaf(zv, sc, sh));
        s1 += du; s2 += (double)du * ((zv - mean) * rstd);
    }
    __shared__ double shm[2][256];
    shm[0][threadIdx.x] = s1; shm[1][threadIdx.x] = s2;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { shm[0][threadIdx.x] += shm[0][threadIdx.x + o]; shm[1][threadIdx.x] += shm[1][threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { atomicAdd(&sums[c], shm[0][0]); atomicAdd(&sums[64 + c], shm[1][0]); }
}
template <int ACT>
__global__ void __launch_bounds__(256) k_bn_bwd_apply_cols(const float* __restrict__ da, const float* __restrict__ z, const float* __restrict__ stat,
                                                           const float* __restrict__ gamma, const double* __restrict__ sums, double count,
                                                           float* __restrict__ dz, int C, int F, long long n) {
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    int c = (int)((i % ((long long)C * F)) / F);
    const float mean = stat[c], rstd = stat[64 + c];
    float zv = z[i];
    float du = da[i] * act_grad<ACT>(fmaf(zv, stat[128 + c], stat[192 + c]));
    float xh = (zv - mean) * rstd;
    dz[i] = gamma[c] * rstd * (du - (float)(sums[c] / count) - xh * (float)(sums[64 + c] / count));
}

// ---------------------------------------------------------------------------------------------
// conv weight gradient: dWp[tap][ci][co] += sum over pixels a[p + off(tap)][ci] * dz[p][co]
// grid = (pixel chunks of `rows_per_block` (b,t) rows, taps); 256 threads, each a 4x4 (ci, co) block.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_conv_wgrad_fp32(const float* __restrict__ a, const float* __restrict__ dz, float* __restrict__ dwp,
                                                         int T, int F, int Fp, int kh, int kw, int dil, int rows_per_block, long long nrows) {
    __shared__ __align__(16) float As[32][64];
    __shared__ __align__(16) float Ds[32][64];
    const int tap = blockIdx.y, i = tap / kw, j = tap % kw;
    const int dt = (i - kh / 2) * dil, df = j - kw / 2;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;   // tx: co group, ty: ci group
    float acc[4][4];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = 0.f;
    const long long row0 = (long long)blockIdx.x * rows_per_block;
    for (long long row = row0; row < row0 + rows_per_block && row < nrows; ++row) {
        const int t = (int)(row % T);
        const int ts = t + dt;
        if (ts < 0 || ts >= T) continue;   // block-uniform: the shifted input row is zero padding
        const float* arow = a + (size_t)(row + dt) * Fp * 64;
        const float* drow = dz + (size_t)row * Fp * 64;
        for (int f0 = 0; f0 < F; f0 += 32) {
            __syncthreads();
            for (int idx = tid; idx < 32 * 16; idx += 256) {
                int px = idx >> 4, q = idx & 15;
                int f = f0 + px, fs = f + df;
                float4 av = make_float4(0.f, 0.f, 0.f, 0.f), dv = av;
                if (f < F) {
                    dv = *reinterpret_cast<const float4*>(drow + (size_t)f * 64 + q * 4);
                    if (fs >= 0 && fs < F) av = *reinterpret_cast<const float4*>(arow + (size_t)fs * 64 + q * 4);
                }
                *reinterpret_cast<float4*>(&As[px][q * 4]) = av;
                *reinterpret_cast<float4*>(&Ds[px][q * 4]) = dv;
            }
            __syncthreads();
#pragma unroll 8
            for (int px = 0; px < 32; ++px) {
                float4 av = *reinterpret_cast<const float4*>(&As[px][ty * 4]);
                float4 dv = *reinterpret_cast<const float4*>(&Ds[px][tx * 4]);
                float a4[4] = {av.x, av.y, av.z, av.w}, d4[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(a4[p], d4[q], acc[p][q]);
            }
        }
    }
    float* dst = dwp + (size_t)tap * 64 * 64;
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) atomicAdd(&dst[(ty * 4 + p) * 64 + tx * 4 + q], acc[p][q]);
}
// cnn1 weight gradient: dWp[j][co] += sum x[row][f + j - 3] * dz[row][f][co]
__global__ void __launch_bounds__(256) k_front_wgrad_fp32(const float* __restrict__ x, const float* __restrict__ dz, float* __restrict__ dwp,
                                                          int F, int Fp, long long nrows) {
    const int co = threadIdx.x & 63, part = threadIdx.x >> 6;
    float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
        const float* xr = x + (size_t)row * F;
        const float* dr = dz + (size_t)row * Fp * 64;
        for (int f = part; f < F; f += 4) {
            float d = dr[(size_t)f * 64 + co];
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                int fs = f + j - 3;
                if (fs >= 0 && fs < F) acc[j] = fmaf(xr[fs], d, acc[j]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 7; ++j) atomicAdd(&dwp[j * 64 + co], acc[j]);
}
// cnn8: dW8p[ci][c] += sum a[pix][ci] * dz7[row][c*F+f];  da[pix][ci] = sum_c W8p[ci][c] * dz7
__global__ void __launch_bounds__(256) k_point8_bwd_fp32(const float* __restrict__ a, const float* __restrict__ dz7, const float* __restrict__ w8p,
                                                         float* __restrict__ da, float* __restrict__ dw8p, int F, int Fp, long long nrows) {
    __shared__ float ws[64 * 8];
    __shared__ float red[64 * 8];
    for (int i = threadIdx.x; i < 512; i += 256) { ws[i] = w8p[i]; red[i] = 0.f; }
    __syncthreads();
    const int ci = threadIdx.x & 63, part = threadIdx.x >> 6;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
        const float* ar = a + (size_t)row * Fp * 64;
        float* dar = da + (size_t)row * Fp * 64;
        const float* dzr = dz7 + (size_t)row * 8 * F;
        for (int f = part; f < Fp; f += 4) {
            float g = 0.f;
            if (f < F) {
                float av = ar[(size_t)f * 64 + ci];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float d = dzr[(size_t)c * F + f];
                    acc[c] = fmaf(av, d, acc[c]);
                    g = fmaf(ws[ci * 8 + c], d, g);
                }
            }
            dar[(size_t)f * 64 + ci] = g;
        }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) atomicAdd(&red[ci * 8 + c], acc[c]);
    __syncthreads();
    for (int i = threadIdx.x; i < 512; i += 256) atomicAdd(&dw8p[i], red[i]);
}

// packed [tap][ci][co] gradient -> reference layout [co][ci][kh][kw]
__global__ void k_unpack_conv_grad(const float* __restrict__ dwp, float* __restrict__ dw, int cout, int cin, int taps) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cout * cin * taps) return;
    int co = i % cout, ci = (i / cout) % cin, tap = i / (cout * cin);
    dw[((size_t)co * cin + ci) * taps + tap] = dwp[i];
}
// forward weights [tap][ci][co] -> data-gradient weights [tap'][co][ci] with the taps flipped
__global__ void k_pack_conv_dgrad(const float* __restrict__ wp, float* __restrict__ wt, int kh, int kw) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int n = kh * kw * 64 * 64;
    if (i >= n) return;
    int ci = i & 63, co = (i >> 6) & 63, tap = i >> 12;          // output index: [tap'][co (in)][ci (out)]
    int it = tap / kw, jt = tap % kw;
    int src_tap = (kh - 1 - it) * kw + (kw - 1 - jt);
    wt[i] = wp[((size_t)src_tap * 64 + ci) * 64 + co];
}

// ---------------------------------------------------------------------------------------------
// strided fp32 GEMM: C[i][j] (+)= sum_k A[i*sai + k*sak] * B[k*sbk + j*sbj]; optional ReLU mask on A
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gemm_strided(const float* __restrict__ A, long long sai, long long sak, const float* __restrict__ Bm, long long sbk,
                                                      long long sbj, float* __restrict__ C, long long ldc, int I, int J, int K, int accumulate) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
    float acc[4][4];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = 0.f;
    // pick the faster-varying index of each operand for the loading threads
    const bool a_k_fast = (sak == 1), b_k_fast = (sbk == 1);
    for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int e = r * 256 + tid;                     // 1024 elements per operand tile
            int ak = a_k_fast ? (e & 15) : (e >> 6), ai = a_k_fast ? (e >> 4) : (e & 63);
            int bk = b_k_fast ? (e & 15) : (e >> 6), bj = b_k_fast ? (e >> 4) : (e & 63);
            int gi = i0 + ai, gk = k0 + ak;
            As[ak][ai] = (gi < I && gk < K) ? A[(size_t)gi * sai + (size_t)gk * sak] : 0.f;
            int gj = j0 + bj, gk2 = k0 + bk;
            Bs[bk][bj] = (gj < J && gk2 < K) ? Bm[(size_t)gk2 * sbk + (size_t)gj * sbj] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(a4[p], b4[q], acc[p][q]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        int i = i0 + ty * 4 + p;
        if (i >= I) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int j = j0 + tx * 4 + q;
            if (j >= J) continue;
            float* c = C + (size_t)i * ldc + j;
            *c = accumulate ? *c + acc[p][q] : acc[p][q];
        }
    }
}
// column sums: out[j] = sum_i A[i][j] (lda)
__global__ void __launch_bounds__(256) k_colsum(const float* __restrict__ A, long long lda, int I, int J, float* __restrict__ out) {
    int j = blockIdx.x * 32 + (threadIdx.x & 31);
    int part = threadIdx.x >> 5;
    double s = 0.0;
    if (j < J)
        for (int i = part; i < I; i += 8) s += A[(size_t)i * lda + j];
    __shared__ double sh[8][32];
    sh[part][threadIdx.x & 31] = s;
    __syncthreads();
    if (part == 0 && j < J) {
        double t = 0.0;
        for (int p = 0; p < 8; ++p) t += sh[p][threadIdx.x & 31];
        out[j] = (float)t;
    }
}
// elementwise helpers of the head backward
__global__ void k_sigmoid_bwd(const float* __restrict__ g, const float* __restrict__ m, float* __restrict__ out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float mv = m[i]; out[i] = g[i] * mv * (1.f - mv); }
}
__global__ void k_relu_mask(float* __restrict__ g, const float* __restrict__ y, long long n) {   // g *= (y > 0)
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !(y[i] > 0.f)) g[i] = 0.f;
}

// ---------------------------------------------------------------------------------------------
// LSTM backward recurrence (BPTT).  Same CTA layout as the forward kernel (direction x slice of 8
// hidden units; thread = one unit x two utterances).  Going backwards in processing order, a CTA
//   1. forms the pre-activation gradients da_{i,f,g,o} of ITS units for step s from dh (the incoming
//      lstm_out gradient plus the recurrent term it computed in the previous iteration) and the
//      saved gate activations / cell states, writes them to `gates` (in place, [B*T][8H]) and to the
//      transposed exchange buffer,
//   2. waits for the other slices, then computes its slice of dh_{s-1} = W_hh^T da_s (K = 4H).
// ---------------------------------------------------------------------------------------------
constexpr int kHS = 8, kBT = 64;
__device__ __forceinline__ void dir_barrier2(unsigned int* counter, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned int spins = 0;
        while (*reinterpret_cast<volatile unsigned int*>(counter) < target) {
            if (++spins > (1u << 30)) __trap();
        }
        __threadfence();
    }
    __syncthreads();
}
__global__ void __launch_bounds__(256, 1) k_lstm_bwd_fp32(float* __restrict__ gates /* in: i,f,g,o activations; out: da */, const float* __restrict__ cseq,
                                                          const float* __restrict__ dhout, const float* __restrict__ whh, float* dgx /*[2][2][4H][Bp]*/,
                                                          float* state /* dc, dh_rec: [2][2][H][Bp] */, unsigned int* barrier,
                                                          int B, int Bp, int T, int H, int nslices) {
    extern __shared__ __align__(16) float smem[];
    float* wt = smem;                         // [4H][kHS]: wt[row][j] = whh[d][row][sl*8 + j]
    float* dt_ = smem + (size_t)4 * H * kHS;  // [KC][kBT] chunk of da
    const int KC = H;                         // rows of da staged per chunk (4 chunks of H rows)
    const int d = blockIdx.x / nslices, sl = blockIdx.x % nslices;
    const int tid = threadIdx.x, j = tid & 7, bp = tid >> 3;
    const int hj = sl * kHS + j;
    const bool unit_ok = hj < H;
    for (int idx = tid; idx < 4 * H * kHS; idx += 256) {
        int row = idx >> 3, jj = idx & 7;
        wt[idx] = (sl * kHS + jj < H) ? whh[((size_t)d * 4 * H + row) * H + sl * kHS + jj] : 0.f;
    }
    __syncthreads();
    float* dc_state = state + ((size_t)(d * 2 + 0) * H) * Bp;
    float* dh_state = state + ((size_t)(d * 2 + 1) * H) * Bp;
    const int ntile = (B + kBT - 1) / kBT;
    for (int s = T - 1; s >= 0; --s) {
        const int t = d ? T - 1 - s : s;
        const int tprev = d ? t + 1 : t - 1;          // time index of forward step s-1
        const int par = s & 1;
        float* dg_out = dgx + ((size_t)(d * 2 + par) * 4 * H) * Bp;
        // ---- 1. gate gradients of this CTA's units
        for (int bt = 0; bt < ntile; ++bt) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int b = bt * kBT + bp * 2 + u;
                if (!unit_ok || b >= B) continue;
                const size_t gi = ((size_t)b * T + t) * 8 * H + (size_t)d * 4 * H + hj;
                const float ig = gates[gi], fg = gates[gi + H], gg = gates[gi + 2 * H], og = gates[gi + 3 * H];
                const size_t ci = ((size_t)b * T + t) * 2 * H + (size_t)d * H + hj;
                const float c = cseq[ci];
                const float cprev = s > 0 ? cseq[((size_t)b * T + tprev) * 2 * H + (size_t)d * H + hj] : 0.f;
                const size_t si = (size_t)hj * Bp + b;
                float dh = dhout[ci] + (s < T - 1 ? dh_state[si] : 0.f);
                float dc_in = s < T - 1 ? dc_state[si] : 0.f;
                const float tc = tanhf(c);
                const float dc = fmaf(dh * og, 1.f - tc * tc, dc_in);
                const float da_i = dc * gg * ig * (1.f - ig);
                const float da_f = dc * cprev * fg * (1.f - fg);
                const float da_g = dc * ig * (1.f - gg * gg);
                const float da_o = dh * tc * og * (1.f - og);
                dc_state[si] = dc * fg;                     // dc flowing to step s-1
                gates[gi] = da_i; gates[gi + H] = da_f; gates[gi + 2 * H] = da_g; gates[gi + 3 * H] = da_o;
                dg_out[(size_t)(hj) * Bp + b] = da_i;
                dg_out[(size_t)(H + hj) * Bp + b] = da_f;
                dg_out[(size_t)(2 * H + hj) * Bp + b] = da_g;
                dg_out[(size_t)(3 * H + hj) * Bp + b] = da_o;
            }
        }
        if (s == 0) break;
        dir_barrier2(barrier + d, (unsigned int)(T - s) * nslices);
        // ---- 2. dh_{s-1}[b][hj] = sum_row da_s[row][b] * whh[row][hj]
        for (int bt = 0; bt < ntile; ++bt) {
            const int b0 = bt * kBT;
            float acc0 = 0.f, acc1 = 0.f;
            for (int kc = 0; kc < 4 * H; kc += KC) {
                __syncthreads();
                for (int idx = tid; idx < KC * (kBT / 4); idx += 256) {
                    int k = idx / (kBT / 4), q = idx % (kBT / 4);
                    float4 v = __ldcg(reinterpret_cast<const float4*>(dg_out + (size_t)(kc + k) * Bp + b0 + q * 4));
                    *reinterpret_cast<float4*>(dt_ + k * kBT + q * 4) = v;
                }
                __syncthreads();
#pragma unroll 4
                for (int k = 0; k < KC; ++k) {
                    float w = wt[(kc + k) * kHS + j];
                    float2 g2 = *reinterpret_cast<const float2*>(dt_ + k * kBT + bp * 2);
                    acc0 = fmaf(w, g2.x, acc0); acc1 = fmaf(w, g2.y, acc1);
                }
            }
            if (unit_ok) {
                int b = b0 + bp * 2;
                if (b < B) dh_state[(size_t)hj * Bp + b] = acc0;
                if (b + 1 < B) dh_state[(size_t)hj * Bp + b + 1] = acc1;
            }
        }
        __syncthreads();
    }
}

__global__ void k_sums_to_grads(const double* __restrict__ sums, float* dgamma, float* dbeta, int C) {
    int c = threadIdx.x;
    if (c < C) { if (dbeta) dbeta[c] = (float)sums[c]; if (dgamma) dgamma[c] = (float)sums[64 + c]; }
}

// ---------------------------------------------------------------------------------------------
// d loss / d spectrogram: the transposed 1x7 filter of cnn1 applied to dz0, summed over the 64 channels.
// One warp per output pixel pair is overkill: a thread owns one pixel and walks 7 taps x 16 float4 of channels;
// neighbouring threads read neighbouring pixels (the 7-pixel windows overlap in L1).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_front_dgrad(const float* __restrict__ dz0, const float* __restrict__ w, float* __restrict__ dx,
                                                     int F, int Fp, long long nrows) {
    __shared__ __align__(16) float ws[7 * 64];
    for (int i = threadIdx.x; i < 7 * 64; i += 256) ws[i] = w[i];
    __syncthreads();
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= nrows * F) return;
    const long long row = i / F;
    const int f = (int)(i - row * F);
    const float* zr = dz0 + (size_t)row * Fp * 64;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        const int fs = f - j + 3;              // z0[fs] used x[fs + j - 3] = x[f] with tap j
        if (fs < 0 || fs >= F) continue;
        const float4* zp = reinterpret_cast<const float4*>(zr + (size_t)fs * 64);
        const float4* wp = reinterpret_cast<const float4*>(ws + j * 64);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float4 a = zp[q], b = wp[q];
            acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
        }
    }
    dx[i] = acc;
}

// ---------------------------------------------------------------------------------------------
// launch wrappers
// ---------------------------------------------------------------------------------------------
#define VS_ACT_DISPATCH(act, CALL_MISH, CALL_RELU) do { if ((act) == VS_ACT_RELU) { CALL_RELU; } else { CALL_MISH; } } while (0)

cudaError_t tr_bn_stats_plane(const float* z, double* sums, int F, int Fp, long long nrows, int num_sms, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(sums, 0, 128 * sizeof(double), st);
    if (e != cudaSuccess) return e;
    int grid = (int)(nrows < (long long)num_sms * 8 ? nrows : (long long)num_sms * 8);
    k_bn_stats_plane<<<grid, 256, 0, st>>>(z, sums, F, Fp, nrows);
    return cudaGetLastError();
}
cudaError_t tr_bn_stats_cols(const float* z, double* sums, int C, int F, long long nrows, int num_sms, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(sums, 0, 128 * sizeof(double), st);
    if (e != cudaSuccess) return e;
    long long n = nrows * F;
    int gx = (int)((n + 255) / 256 < (long long)num_sms * 2 ? (n + 255) / 256 : (long long)num_sms * 2);
    k_bn_stats_cols<<<dim3(gx, C), 256, 0, st>>>(z, sums, C, F, nrows);
    return cudaGetLastError();
}
cudaError_t tr_bn_finalize(const double* sums, double count, const float* gamma, const float* beta, float* stat, float* rmean, float* rvar,
                           long long* nb, float momentum, int C, cudaStream_t st) {
    k_bn_finalize<<<1, 64, 0, st>>>(sums, count, gamma, beta, stat, rmean, rvar, nb, momentum, C);
    return cudaGetLastError();
}
cudaError_t tr_bn_act_plane(int act, const float* z, float* a, const float* stat, int F, int Fp, long long npix, cudaStream_t st,
                            elt16* ahi, elt16* alo, int elt) {
    long long n4 = npix * 16;
    unsigned grid = (unsigned)((n4 + 255) / 256);
    VS_ACT_DISPATCH(act, (k_bn_act_plane<VS_ACT_MISH><<<grid, 256, 0, st>>>(z, a, stat, F, Fp, n4, ahi, alo, elt)),
                    (k_bn_act_plane<VS_ACT_RELU><<<grid, 256, 0, st>>>(z, a, stat, F, Fp, n4, ahi, alo, elt)));
    return cudaGetLastError();
}
cudaError_t tr_bn_act_cols(int act, const float* z, float* a, const float* stat, int C, int F, long long nrows, cudaStream_t st) {
    long long n = nrows * C * F;
    unsigned grid = (unsigned)((n + 255) / 256);
    VS_ACT_DISPATCH(act, (k_bn_act_cols<VS_ACT_MISH><<<grid, 256, 0, st>>>(z, a, stat, C, F, n)),
                    (k_bn_act_cols<VS_ACT_RELU><<<grid, 256, 0, st>>>(z, a, stat, C, F, n)));
    return cudaGetLastError();
}
cudaError_t tr_bn_sync(const BnSync& sync, double* sums, cudaStream_t st) {
    if (!sync.fn) return cudaSuccess;
    if (sync.fn(sync.user, sums, 128, (void*)st) != 0) { set_error("SyncBN statistics all-reduce callback failed"); return cudaErrorUnknown; }
    return cudaSuccess;
}
cudaError_t tr_front_dgrad(const float* dz0, const float* w, float* dx, int F, int Fp, long long nrows, cudaStream_t st) {
    const long long n = nrows * F;
    k_front_dgrad<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dz0, w, dx, F, Fp, nrows);
    return cudaGetLastError();
}
cudaError_t tr_bn_bwd_plane(int act, const float* da, const float* z, const float* stat, const float* gamma, double* sums, float* dz,
                            int F, int Fp, long long nrows, int num_sms, cudaStream_t st, elt16* dhi, elt16* dlo, float* dgamma, float* dbeta,
                            const BnSync& sync) {
    cudaError_t e = cudaMemsetAsync(sums, 0, 128 * sizeof(double), st);
    if (e != cudaSuccess) return e;
    int grid = (int)(nrows < (long long)num_sms * 8 ? nrows : (long long)num_sms * 8);
    VS_ACT_DISPATCH(act, (k_bn_bwd_reduce_plane<VS_ACT_MISH><<<grid, 256, 0, st>>>(da, z, stat, sums, F, Fp, nrows)),
                    (k_bn_bwd_reduce_plane<VS_ACT_RELU><<<grid, 256, 0, st>>>(da, z, stat, sums, F, Fp, nrows)));
    if (dgamma || dbeta) k_sums_to_grads<<<1, 64, 0, st>>>(sums, dgamma, dbeta, 64);
    e = tr_bn_sync(sync, sums, st);
    if (e != cudaSuccess) return e;
    const long long npix = nrows * Fp;
    const double count = (double)nrows * F * sync.world;
    unsigned g2 = (unsigned)((npix * 16 + 255) / 256);
    VS_ACT_DISPATCH(act, (k_bn_bwd_apply_plane<VS_ACT_MISH><<<g2, 256, 0, st>>>(da, z, stat, gamma, sums, count, dz, F, Fp, npix, dhi, dlo)),
                    (k_bn_bwd_apply_plane<VS_ACT_RELU><<<g2, 256, 0, st>>>(da, z, stat, gamma, sums, count, dz, F, Fp, npix, dhi, dlo)));
    return cudaGetLastError();
}
cudaError_t tr_bn_bwd_cols(int act, const float* da, const float* z, const float* stat, const float* gamma, double* sums, float* dz,
                           int C, int F, long long nrows, int num_sms, cudaStream_t st, float* dgamma, float* dbeta, const BnSync& sync) {
    cudaError_t e = cudaMemsetAsync(sums, 0, 128 * sizeof(double), st);
    if (e != cudaSuccess) return e;
    long long nf = nrows * F;
    int gx = (int)((nf + 255) / 256 < (long long)num_sms * 2 ? (nf + 255) / 256 : (long long)num_sms * 2);
    VS_ACT_DISPATCH(act, (k_bn_bwd_reduce_cols<VS_ACT_MISH><<<dim3(gx, C), 256, 0, st>>>(da, z, stat, sums, C, F, nrows)),
                    (k_bn_bwd_reduce_cols<VS_ACT_RELU><<<dim3(gx, C), 256, 0, st>>>(da, z, stat, sums, C, F, nrows)));
    if (dgamma || dbeta) k_sums_to_grads<<<1, 64, 0, st>>>(sums, dgamma, dbeta, C);
    e = tr_bn_sync(sync, sums, st);
    if (e != cudaSuccess) return e;
    long long n = nrows * C * F;
    unsigned g2 = (unsigned)((n + 255) / 256);
    const double count = (double)nf * sync.world;
    VS_ACT_DISPATCH(act, (k_bn_bwd_apply_cols<VS_ACT_MISH><<<g2, 256, 0, st>>>(da, z, stat, gamma, sums, count, dz, C, F, n)),
                    (k_bn_bwd_apply_cols<VS_ACT_RELU><<<g2, 256, 0, st>>>(da, z, stat, gamma, sums, count, dz, C, F, n)));
    return cudaGetLastError();
}
cudaError_t tr_conv_wgrad(const float* a, const float* dz, float* dwp, int T, int F, int Fp, int kh, int kw, int dil, long long nrows, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(dwp, 0, (size_t)kh * kw * 64 * 64 * sizeof(float), st);
    if (e != cudaSuccess) return e;
    const int rpb = 8;
    dim3 grid((unsigned)((nrows + rpb - 1) / rpb), kh * kw);
    k_conv_wgrad_fp32<<<grid, 256, 0, st>>>(a, dz, dwp, T, F, Fp, kh, kw, dil, rpb, nrows);
    return cudaGetLastError();
}
cudaError_t tr_front_wgrad(const float* x, const float* dz, float* dwp, int F, int Fp, long long nrows, int num_sms, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(dwp, 0, 7 * 64 * sizeof(float), st);
    if (e != cudaSuccess) return e;
    int grid = (int)(nrows < (long long)num_sms * 4 ? nrows : (long long)num_sms * 4);
    k_front_wgrad_fp32<<<grid, 256, 0, st>>>(x, dz, dwp, F, Fp, nrows);
    return cudaGetLastError();
}
cudaError_t tr_point8_bwd(const float* a, const float* dz7, const float* w8p, float* da, float* dw8p, int F, int Fp, long long nrows, int num_sms, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(dw8p, 0, 512 * sizeof(float), st);
    if (e != cudaSuccess) return e;
    int grid = (int)(nrows < (long long)num_sms * 4 ? nrows : (long long)num_sms * 4);
    k_point8_bwd_fp32<<<grid, 256, 0, st>>>(a, dz7, w8p, da, dw8p, F, Fp, nrows);
    return cudaGetLastError();
}
cudaError_t tr_unpack_conv_grad(const float* dwp, float* dw, int cout, int cin, int taps, cudaStream_t st) {
    int n = cout * cin * taps;
    k_unpack_conv_grad<<<(n + 255) / 256, 256, 0, st>>>(dwp, dw, cout, cin, taps);
    return cudaGetLastError();
}
cudaError_t tr_pack_conv_dgrad(const float* wp, float* wt, int kh, int kw, cudaStream_t st) {
    int n = kh * kw * 64 * 64;
    k_pack_conv_dgrad<<<(n + 255) / 256, 256, 0, st>>>(wp, wt, kh, kw);
    return cudaGetLastError();
}
cudaError_t tr_gemm(const float* A, long long sai, long long sak, const float* B, long long sbk, long long sbj, float* C, long long ldc,
                    int I, int J, int K, bool accumulate, cudaStream_t st) {
    dim3 grid((J + 63) / 64, (I + 63) / 64);
    k_gemm_strided<<<grid, 256, 0, st>>>(A, sai, sak, B, sbk, sbj, C, ldc, I, J, K, accumulate ? 1 : 0);
    return cudaGetLastError();
}
cudaError_t tr_colsum(const float* A, long long lda, int I, int J, float* out, cudaStream_t st) {
    k_colsum<<<(J + 31) / 32, 256, 0, st>>>(A, lda, I, J, out);
    return cudaGetLastError();
}
cudaError_t tr_sigmoid_bwd(const float* g, const float* m, float* out, long long n, cudaStream_t st) {
    k_sigmoid_bwd<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g, m, out, n);
    return cudaGetLastError();
}
cudaError_t tr_relu_mask(float* g, const float* y, long long n, cudaStream_t st) {
    k_relu_mask<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g, y, n);
    return cudaGetLastError();
}
size_t tr_lstm_bwd_scratch_bytes(int H, int B) {
    const size_t Bp = align_up((size_t)B, kBT);
    return ((size_t)2 * 2 * 4 * H + (size_t)2 * 2 * H) * Bp * sizeof(float) + 256;
}
cudaError_t tr_lstm_bwd(const vs_engine* e, float* gates, const float* cseq, const float* dhout, void* scratch, int B, int T, cudaStream_t st) {
    const int H = e->d.lstm_dim;
    const int nslices = (H + kHS - 1) / kHS;
    const int Bp = (int)align_up((size_t)B, kBT);
    float* dgx = (float*)scratch;
    float* state = dgx + (size_t)2 * 2 * 4 * H * Bp;
    unsigned int* barrier = (unsigned int*)(state + (size_t)2 * 2 * H * Bp);
    size_t smem = ((size_t)4 * H * kHS + (size_t)H * kBT) * sizeof(float);
    cudaError_t err = cudaFuncSetAttribute(k_lstm_bwd_fp32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    err = cudaMemsetAsync(barrier, 0, 2 * sizeof(unsigned int), st);
    if (err != cudaSuccess) return err;
    const float* whh = e->whh;
    int Bv = B, Bpv = Bp, Tv = T, Hv = H, ns = nslices;
    void* args[] = {(void*)&gates, (void*)&cseq, (void*)&dhout, (void*)&whh, (void*)&dgx, (void*)&state, (void*)&barrier,
                    (void*)&Bv, (void*)&Bpv, (void*)&Tv, (void*)&Hv, (void*)&ns};
    return cudaLaunchCooperativeKernel((const void*)k_lstm_bwd_fp32, dim3(2 * nslices), dim3(256), args, smem, st);
}

}  // namespace vs
