// Conv weight gradient on tensor cores (training backward of cnn2..cnn7):
//     dW[tap][ci][co] = sum over pixels p of  a[p + off(tap)][ci] * dz[p][co]
// The contraction runs over PIXELS, which are the rows of the channels-last planes, so both MMA
// operands are MN-major (tools/umma_probe.cu "mn_*" cases):
//     D[128][64] (TMEM) += A[K = 16 pixels][M = 128] * B[K = 16 pixels][N = 64 co]
// A is a window into an activation strip; its two 64-wide M blocks are the same strip one pixel
// apart (descriptor leading byte offset = 128 B), i.e. the taps df and df+1 of one filter row -
// the same tap pairing as the forward kernel.  B is the dz tile of the chunk.
// CTA (kind dt, index i) owns filter row dt for every 128-pixel chunk c = i, i + cpk, ...; the n_j
// tap pairs of that row accumulate in n_j x 64 TMEM columns.  Because tcgen05 accumulation truncates,
// the accumulators are flushed (atomicAdd into the fp32 global gradient) every kFlush chunks,
// alternating two TMEM buffers so the MMAs never wait for the flush.
// Operands are bf16 hi/lo planes (fp32 exponent range: no loss scaling); 3 passes hi*hi + lo*hi + hi*lo.
#include "tc.cuh"
#include "sm100_ptx.cuh"

namespace vs {
using namespace ptx;

constexpr int kWgChunk = 128;      // pixels per chunk (MMA K per chunk)
constexpr int kWgStrip = 136;      // strip rows: chunk + halo, multiple of 8
constexpr int kWgStages = 3;
constexpr int kWgFlush = 16;       // chunks accumulated in TMEM before promotion to global fp32
constexpr int kWgStageBytes = 2 * kWgStrip * 128 + 2 * kWgChunk * 128;

struct WgradTcArgs {
    int Q, B, chunks_per_utt, total_chunks, cpk;   // cpk: CTAs per kind
    int n_dt, n_j, halo, dt_stride;
    float* dwp;                                     // [n_dt][n_j][2][64 ci][64 co]
};

__global__ void __launch_bounds__(192, 1) k_wgrad_tc(const WgradTcArgs a, const __grid_constant__ CUtensorMap tm_a_hi,
                                                     const __grid_constant__ CUtensorMap tm_a_lo,
                                                     const __grid_constant__ CUtensorMap tm_d_hi,
                                                     const __grid_constant__ CUtensorMap tm_d_lo) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kWgStages * kWgStageBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kWgStages;
    uint64_t* acc_full = empty + kWgStages;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dt = blockIdx.x / a.cpk, idx = blockIdx.x % a.cpk;
    const int my_chunks = idx < a.total_chunks ? (a.total_chunks - idx + a.cpk - 1) / a.cpk : 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kWgStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // producer / issuer: the whole warp runs the warp-uniform loops and waits, one elected lane issues the TMA / tcgen05
    // instructions (inside an `if (lane == 0)` region every uniform-datapath instruction gets an elect-and-loop wrapper)
    if (warp == 0) {
        {
            int st = 0, ph = 0;
            for (int it = 0; it < my_chunks; ++it) {
                const int c = idx + it * a.cpk;
                const int b = c / a.chunks_per_utt, q0 = (c - b * a.chunks_per_utt) * kWgChunk;
                const int qs = q0 - a.halo + (dt - a.n_dt / 2) * a.dt_stride;
                mbar_wait(&empty[st], ph ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full[st], (uint32_t)kWgStageBytes);
                    uint8_t* dst = smem + (size_t)st * kWgStageBytes;
                    tma_load_3d(dst, &tm_a_hi, &full[st], 0, qs, b);
                    tma_load_3d(dst + kWgStrip * 128, &tm_a_lo, &full[st], 0, qs, b);
                    tma_load_3d(dst + 2 * kWgStrip * 128, &tm_d_hi, &full[st], 0, q0, b);
                    tma_load_3d(dst + 2 * kWgStrip * 128 + kWgChunk * 128, &tm_d_lo, &full[st], 0, q0, b);
                }
                __syncwarp();
                if (++st == kWgStages) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        {
            // bf16 operands, fp32 accumulate, A and B both MN-major
            const uint32_t idesc = make_idesc_bf16(128, 64) | (1u << 15) | (1u << 16);
            int st = 0, ph = 0;
            for (int it = 0; it < my_chunks; ++it) {
                const int grp = it / kWgFlush, buf = grp & 1;
                const bool first_of_group = (it % kWgFlush) == 0;
                if (first_of_group) {
                    mbar_wait(&acc_empty[buf], ((grp >> 1) & 1) ^ 1);
                    tc_fence_after();
                }
                mbar_wait(&full[st], ph);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(smem + (size_t)st * kWgStageBytes), a_lo = a_hi + kWgStrip * 128;
                const uint32_t d_hi = a_hi + 2 * kWgStrip * 128, d_lo = d_hi + kWgChunk * 128;
                if (elect_one()) {
                    for (int pass = 0; pass < 3; ++pass) {      // hi*hi, lo*hi, hi*lo
                        const uint32_t ap = pass == 1 ? a_lo : a_hi, dp = pass == 2 ? d_lo : d_hi;
                        for (int j = 0; j < a.n_j; ++j) {
                            const uint32_t d_tmem = tmem + (uint32_t)(buf * 256 + j * 64);
                            // K step = 16 pixel rows = 2048 bytes = 128 descriptor address units
                            const uint64_t da0 = make_smem_desc(ap + (uint32_t)(2 * j) * 128, 128, 1024, 2);
                            const uint64_t db0 = make_smem_desc(dp, 128, 1024, 2);
#pragma unroll
                            for (int k = 0; k < kWgChunk / 16; ++k)
                                umma_bf16(d_tmem, da0 + 128 * k, db0 + 128 * k, idesc, (first_of_group && pass == 0 && k == 0) ? 0u : 1u);
                        }
                    }
                    umma_commit(&empty[st]);
                    if ((it % kWgFlush) == kWgFlush - 1 || it == my_chunks - 1) umma_commit(&acc_full[buf]);
                }
                __syncwarp();
                if (++st == kWgStages) { st = 0; ph ^= 1; }
            }
        }
    } else {
        // flush: TMEM lane m = (h = m / 64, ci = m % 64); column j*64 + co
        const int quad = warp & 3;
        const int m = quad * 32 + lane, h = m >> 6, ci = m & 63;
        const int ngroups = (my_chunks + kWgFlush - 1) / kWgFlush;
        for (int grp = 0; grp < ngroups; ++grp) {
            const int buf = grp & 1;
            mbar_wait(&acc_full[buf], (grp >> 1) & 1);
            tc_fence_after();
            for (int j = 0; j < a.n_j; ++j) {
                float* dst = a.dwp + ((((size_t)dt * a.n_j + j) * 2 + h) * 64 + ci) * 64;
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 256 + j * 64 + c0), r);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 32; ++q) atomicAdd(dst + c0 + q, __uint_as_float(r[q]));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

// dwp [n_dt][n_j][2][ci][co] -> reference layout dW[co][ci][kh][kw] (tap slots beyond kw are padding)
__global__ void k_unpack_wgrad_tc(const float* __restrict__ dwp, float* __restrict__ dw, int kh, int kw, int n_j) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int n = 64 * 64 * kh * kw;
    if (i >= n) return;
    int df = i % kw, dt_ = (i / kw) % kh, ci = (i / (kw * kh)) % 64, co = i / (kw * kh * 64);
    int j = df >> 1, h = df & 1;
    dw[i] = dwp[((((size_t)dt_ * n_j + j) * 2 + h) * 64 + ci) * 64 + co];
}

int tc_train_wgrad(vs_engine* e, int layer, const elt16* a_hi, const elt16* a_lo, const elt16* d_hi, const elt16* d_lo, float* dwp,
                   float* dw_out, int B, int T, int kid, cudaStream_t st) {
    const ConvGeom g = kConv[layer];
    const int F = e->d.num_freq, Fp = padded_freq(F);
    WgradTcArgs a{};
    a.Q = T * Fp; a.B = B;
    a.chunks_per_utt = (a.Q + kWgChunk - 1) / kWgChunk;
    a.total_chunks = B * a.chunks_per_utt;
    a.n_dt = g.kh; a.n_j = (g.kw + 1) / 2; a.halo = g.kw / 2; a.dt_stride = g.dil * Fp;
    a.cpk = e->num_sms / a.n_dt;
    if (a.cpk < 1) { set_error("too few SMs for the wgrad kernel"); return VS_ERR_UNSUPPORTED; }
    a.dwp = dwp;
    CUtensorMap tm_a_hi, tm_a_lo, tm_d_hi, tm_d_lo;
    {
        uint64_t dims[3] = {64, (uint64_t)a.Q, (uint64_t)B};
        uint64_t str[2] = {128, (uint64_t)a.Q * 128};
        uint32_t boxa[3] = {64, (uint32_t)kWgStrip, 1}, boxd[3] = {64, (uint32_t)kWgChunk, 1};
        bool ok = make_tmap_bf16(&tm_a_hi, (void*)a_hi, 3, dims, str, boxa, CU_TENSOR_MAP_SWIZZLE_128B);
        ok = ok && make_tmap_bf16(&tm_a_lo, (void*)a_lo, 3, dims, str, boxa, CU_TENSOR_MAP_SWIZZLE_128B);
        ok = ok && make_tmap_bf16(&tm_d_hi, (void*)d_hi, 3, dims, str, boxd, CU_TENSOR_MAP_SWIZZLE_128B);
        ok = ok && make_tmap_bf16(&tm_d_lo, (void*)d_lo, 3, dims, str, boxd, CU_TENSOR_MAP_SWIZZLE_128B);
        if (!ok) { set_error("cuTensorMapEncodeTiled failed (wgrad)"); return VS_ERR_CUDA; }
    }
    const size_t nacc = (size_t)a.n_dt * a.n_j * 2 * 64 * 64;
    cudaError_t ce = cudaMemsetAsync(dwp, 0, nacc * sizeof(float), st);
    const int smem = 1024 + kWgStages * kWgStageBytes + 256;
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (ce == cudaSuccess) {
        k_wgrad_tc<<<a.n_dt * a.cpk, 192, smem, st>>>(a, tm_a_hi, tm_a_lo, tm_d_hi, tm_d_lo);
        ce = cudaGetLastError();
    }
    if (ce == cudaSuccess) {
        const int n = 64 * 64 * g.kh * g.kw;
        k_unpack_wgrad_tc<<<(n + 255) / 256, 256, 0, st>>>(dwp, dw_out, g.kh, g.kw, a.n_j);
        ce = cudaGetLastError();
    }
    if (ce != cudaSuccess) { set_error(std::string("k_wgrad_tc: ") + cudaGetErrorString(ce)); return VS_ERR_CUDA; }
    e->launches += 2;
    if (e->profiling) prof_after(e, kid, st);
    return VS_OK;
}

}  // namespace vs
