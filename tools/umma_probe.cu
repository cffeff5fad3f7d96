// Stand-alone probe of the tcgen05 / TMA building blocks the conv kernel relies on.  Built here
// (nvcc sm_100a), run on the B200 box via gpurun.  Each case computes D[128][N] = A[128][64] * Bwin^T
// where Bwin is an N-row window of a (N + 8)-row strip starting `shift` rows into the strip, and
// compares with a CPU product of the same bf16 values.  Cases vary: swizzle mode (128B / none), how
// shared memory was filled (software layout vs TMA), row shift of the window and the descriptor
// base_offset.  A final timing case measures cycles per MMA for back-to-back issue.
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#include "../voicesplit_b200/csrc/sm100_ptx.cuh"

using namespace vs;
using namespace vs::ptx;

struct ProbeArgs {
    const __nv_bfloat16* A;   // [128][64]
    const __nv_bfloat16* B;   // [strip_rows][64]
    float* D;                 // [128][N]
    int N, strip_rows, shift;
    int layout;               // 2 = SW128, 0 = none
    int fill;                 // 0 = software, 1 = TMA
    int base_offset_mode;     // 0: 0, 1: shift & 7, 2: (start_addr >> 7) & 7
    int iters;                // timing: repeat the 4-MMA group this many times
    long long* cycles;
    int m_dim;                // MMA M (128 default, 64 for the half-height timing case)
};

__global__ void __launch_bounds__(128, 1) probe_kernel(ProbeArgs a, const __grid_constant__ CUtensorMap tmA,
                                                       const __grid_constant__ CUtensorMap tmB) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;               // 128 rows x 128 B = 16 KB
    uint8_t* sB = smem + 16384;       // strip_rows x 128 B
    __shared__ uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;

    if (tid == 0) {
        mbar_init(&bar_load, 1);
        mbar_init(&bar_mma, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(&tmem_base_s, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (a.fill == 0) {
        // software fill: 16-byte chunks
        const uint4* gA = reinterpret_cast<const uint4*>(a.A);
        const uint4* gB = reinterpret_cast<const uint4*>(a.B);
        for (int idx = tid; idx < 128 * 8; idx += 128) {
            int r = idx >> 3, c = idx & 7;
            uint32_t off = a.layout == 2 ? (uint32_t)(r * 128 + ((c ^ (r & 7)) * 16)) : (uint32_t)(c * (128 * 16) + r * 16);
            *reinterpret_cast<uint4*>(sA + off) = gA[idx];
        }
        for (int idx = tid; idx < a.strip_rows * 8; idx += 128) {
            int r = idx >> 3, c = idx & 7;
            uint32_t off = a.layout == 2 ? (uint32_t)(r * 128 + ((c ^ (r & 7)) * 16)) : (uint32_t)(c * (a.strip_rows * 16) + r * 16);
            *reinterpret_cast<uint4*>(sB + off) = gB[idx];
        }
        fence_proxy_async();
        __syncthreads();
    } else {
        if (tid == 0) {
            mbar_arrive_expect_tx(&bar_load, (uint32_t)((128 + a.strip_rows) * 128));
            tma_load_2d(sA, &tmA, &bar_load, 0, 0);
            int done = 0;
            while (done < a.strip_rows) {   // box is at most 256 rows (here 136)
                tma_load_2d(sB + done * 128, &tmB, &bar_load, 0, done);
                done += 136;
            }
        }
        mbar_wait(&bar_load, 0);
    }

    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_bf16(a.m_dim ? a.m_dim : 128, a.N);
        const uint32_t aaddr = smem_u32(sA), baddr = smem_u32(sB);
        tc_fence_after();
        t0 = clock64();
        for (int it = 0; it < a.iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint64_t da, db;
                if (a.layout == 2) {
                    uint32_t bstart = baddr + a.shift * 128 + k * 32;
                    uint32_t bo = a.base_offset_mode == 0 ? 0 : (a.base_offset_mode == 1 ? (a.shift & 7) : ((bstart >> 7) & 7));
                    da = make_smem_desc(aaddr + k * 32, 16, 1024, 2, 0);
                    db = make_smem_desc(bstart, 16, 1024, 2, bo);
                } else {
                    da = make_smem_desc(aaddr + k * 2 * (128 * 16), 128 * 16, 128, 0, 0);
                    db = make_smem_desc(baddr + a.shift * 16 + k * 2 * (a.strip_rows * 16), a.strip_rows * 16, 128, 0, 0);
                }
                umma_bf16(tmem, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
            }
        }
        umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    if (tid == 0) {
        t1 = clock64();
        if (a.cycles) *a.cycles = t1 - t0;
    }
    tc_fence_after();
    // epilogue: warp w reads lanes 32w..32w+31
    for (int c0 = 0; c0 < a.N; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
        int row = warp * 32 + (tid & 31);
        for (int j = 0; j < 32; ++j)
            if (c0 + j < a.N) a.D[(size_t)row * a.N + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

// MN-major operands (the weight-gradient GEMM contracts over PIXELS, which are the smem rows):
//   D[128][64] = sum_k A[k][m] * B[k][n],  A rows m < 64 read strip row (k + shift), m >= 64 read strip row
//   (k + shift + 1): the second 64-wide MN block of A is the same strip one pixel later, addressed through the
//   descriptor's leading byte offset (= 128 B).  Tiles are [rows][64] 16-bit, 128-byte swizzle, filled by TMA.
struct ProbeMnArgs {
    float* D;          // [128][64]
    int shift, K;      // K multiple of 16, <= 64
    int lbo_bytes, sbo_bytes, swap;   // descriptor variants
};
__global__ void __launch_bounds__(128, 1) probe_mn_kernel(ProbeMnArgs a, const __grid_constant__ CUtensorMap tmA,
                                                          const __grid_constant__ CUtensorMap tmB) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;               // 128 rows x 128 B (strip: rows = pixels)
    uint8_t* sB = smem + 16384;       // 136 rows x 128 B (dz tile: rows = pixels)
    __shared__ uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar_load, 1); mbar_init(&bar_mma, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(&tmem_base_s, 64); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        mbar_arrive_expect_tx(&bar_load, (uint32_t)((128 + 136) * 128));
        tma_load_2d(sA, &tmA, &bar_load, 0, 0);
        tma_load_2d(sB, &tmB, &bar_load, 0, 0);
    }
    mbar_wait(&bar_load, 0);
    if (tid == 0) {
        // idesc: bf16 A/B, fp32 D, both MN-major
        const uint32_t idesc = make_idesc_bf16(128, 64) | (1u << 15) | (1u << 16);
        tc_fence_after();
        for (int k = 0; k < a.K / 16; ++k) {
            uint32_t astart = smem_u32(sA) + (uint32_t)(a.shift + 16 * k) * 128;
            uint32_t bstart = smem_u32(sB) + (uint32_t)(16 * k) * 128;
            uint64_t da = a.swap ? make_smem_desc(astart, a.sbo_bytes, a.lbo_bytes, 2) : make_smem_desc(astart, a.lbo_bytes, a.sbo_bytes, 2);
            uint64_t db = a.swap ? make_smem_desc(bstart, a.sbo_bytes, a.lbo_bytes, 2) : make_smem_desc(bstart, a.lbo_bytes, a.sbo_bytes, 2);
            umma_bf16(tmem, da, db, idesc, k > 0 ? 1u : 0u);
        }
        umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
        int row = warp * 32 + (tid & 31);
        for (int j = 0; j < 32; ++j) a.D[(size_t)row * 64 + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

static float bf(float v) { return __bfloat162float(__float2bfloat16(v)); }

int main() {
    const int maxN = 256, strip = maxN + 16;
    std::vector<float> hA(128 * 64), hB(strip * 64);
    srand(1);
    for (auto& v : hA) v = bf((rand() % 2001 - 1000) / 1000.f);
    for (auto& v : hB) v = bf((rand() % 2001 - 1000) / 1000.f);
    std::vector<__nv_bfloat16> bA(hA.size()), bB(hB.size());
    for (size_t i = 0; i < hA.size(); ++i) bA[i] = __float2bfloat16(hA[i]);
    for (size_t i = 0; i < hB.size(); ++i) bB[i] = __float2bfloat16(hB[i]);
    __nv_bfloat16 *dA, *dB;
    float* dD;
    long long* dcyc;
    CK(cudaMalloc(&dA, bA.size() * 2)); CK(cudaMalloc(&dB, bB.size() * 2));
    CK(cudaMalloc(&dD, 128 * maxN * 4)); CK(cudaMalloc(&dcyc, 8));
    CK(cudaMemcpy(dA, bA.data(), bA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, bB.data(), bB.size() * 2, cudaMemcpyHostToDevice));

    CUtensorMap tmA, tmB;
    {
        uint64_t dimsA[2] = {64, 128}, strA[1] = {128};
        uint32_t boxA[2] = {64, 128};
        uint64_t dimsB[2] = {64, (uint64_t)strip}, strB[1] = {128};
        uint32_t boxB[2] = {64, 136};
        if (!make_tmap_bf16(&tmA, dA, 2, dimsA, strA, boxA, CU_TENSOR_MAP_SWIZZLE_128B) ||
            !make_tmap_bf16(&tmB, dB, 2, dimsB, strB, boxB, CU_TENSOR_MAP_SWIZZLE_128B)) {
            printf("PROBE tensor map encode FAILED\n");
            return 3;
        }
    }
    const int smem_bytes = 1024 + 16384 + strip * 128;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));

    struct Case { const char* name; int N, shift, layout, fill, bom; };
    std::vector<Case> cases = {
        {"sw128_soft_N64_shift0", 64, 0, 2, 0, 0},   {"sw128_soft_N256_shift0", 256, 0, 2, 0, 0},
        {"sw128_tma_N256_shift0", 256, 0, 2, 1, 0},  {"sw128_tma_N208_shift0", 208, 0, 2, 1, 0},
        {"sw128_tma_N256_shift8_bo0", 256, 8, 2, 1, 0},
        {"sw128_tma_N256_shift1_bo0", 256, 1, 2, 1, 0}, {"sw128_tma_N256_shift1_boS", 256, 1, 2, 1, 1},
        {"sw128_tma_N256_shift2_bo0", 256, 2, 2, 1, 0}, {"sw128_tma_N256_shift2_boS", 256, 2, 2, 1, 1},
        {"sw128_tma_N256_shift4_bo0", 256, 4, 2, 1, 0}, {"sw128_tma_N256_shift4_boS", 256, 4, 2, 1, 1},
        {"sw128_tma_N256_shift5_bo0", 256, 5, 2, 1, 0}, {"sw128_tma_N256_shift5_boS", 256, 5, 2, 1, 1},
        {"sw128_tma_N256_shift5_boA", 256, 5, 2, 1, 2},
        {"none_soft_N64_shift0", 64, 0, 0, 0, 0},    {"none_soft_N256_shift0", 256, 0, 0, 0, 0},
        {"none_soft_N256_shift1", 256, 1, 0, 0, 0},  {"none_soft_N256_shift2", 256, 2, 0, 0, 0},
        {"none_soft_N256_shift5", 256, 5, 0, 0, 0},  {"none_soft_N208_shift3", 208, 3, 0, 0, 0},
    };
    std::vector<float> hD(128 * maxN);
    for (const Case& c : cases) {
        ProbeArgs a{dA, dB, dD, c.N, strip, c.shift, c.layout, c.fill, c.bom, 1, dcyc};
        CK(cudaMemset(dD, 0xff, 128 * maxN * 4));
        probe_kernel<<<1, 128, smem_bytes>>>(a, tmA, tmB);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("PROBE %-32s CUDA-ERROR %s\n", c.name, cudaGetErrorString(e)); return 4; }
        CK(cudaMemcpy(hD.data(), dD, 128 * c.N * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < c.N; ++n) {
                double s = 0;
                for (int k = 0; k < 64; ++k) s += (double)hA[m * 64 + k] * hB[(n + c.shift) * 64 + k];
                double d = fabs(s - hD[m * c.N + n]);
                if (!(d <= maxerr)) maxerr = d;  // NaN-propagating
            }
        printf("PROBE %-32s max_err=%.3e %s\n", c.name, maxerr, maxerr < 1e-3 ? "PASS" : "FAIL");
    }
    // timing: cycles per MMA (K=16 step) when issued back to back from fixed smem operands
    struct TCase { const char* name; int N, layout; };
    for (const TCase& t : {TCase{"time_sw128_N256", 256, 2}, TCase{"time_sw128_N192", 192, 2}, TCase{"time_sw128_N128", 128, 2},
                           TCase{"time_sw128_N64", 64, 2}, TCase{"time_none_N256", 256, 0}, TCase{"time_none_N192", 192, 0}}) {
        ProbeArgs a{dA, dB, dD, t.N, strip, 0, t.layout, 0, 0, 2000, dcyc};
        probe_kernel<<<1, 128, smem_bytes>>>(a, tmA, tmB);
        CK(cudaDeviceSynchronize());
        long long cyc;
        CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
        printf("PROBE %-32s cycles_per_mma=%.1f (ideal %d)\n", t.name, (double)cyc / (2000.0 * 4), t.N / 2);
    }
    // MN-major operands + two-tap M=128 through the leading byte offset (weight-gradient GEMM)
    {
        CUtensorMap tmAs, tmBs;
        uint64_t dA2[2] = {64, 128}, sA2[1] = {128};
        uint32_t bA2[2] = {64, 128};
        uint64_t dB2[2] = {64, (uint64_t)strip}, sB2[1] = {128};
        uint32_t bB2[2] = {64, 136};
        if (!make_tmap_bf16(&tmAs, dA, 2, dA2, sA2, bA2, CU_TENSOR_MAP_SWIZZLE_128B) || !make_tmap_bf16(&tmBs, dB, 2, dB2, sB2, bB2, CU_TENSOR_MAP_SWIZZLE_128B)) {
            printf("PROBE mn tensor map encode FAILED\n");
            return 3;
        }
        CK(cudaFuncSetAttribute(probe_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        struct MnCase { const char* name; int shift, K, lbo, sbo, swap; };
        for (const MnCase& c : {MnCase{"mn_K16_shift0_lbo128_sbo1024", 0, 16, 128, 1024, 0}, MnCase{"mn_K64_shift0_lbo128_sbo1024", 0, 64, 128, 1024, 0},
                                MnCase{"mn_K64_shift2_lbo128_sbo1024", 2, 64, 128, 1024, 0}, MnCase{"mn_K64_shift5_lbo128_sbo1024", 5, 64, 128, 1024, 0},
                                MnCase{"mn_K64_shift2_swapped", 2, 64, 128, 1024, 1}, MnCase{"mn_K64_shift0_lbo8192(same tap twice?)", 0, 64, 8192, 1024, 0}}) {
            ProbeMnArgs a{dD, c.shift, c.K, c.lbo, c.sbo, c.swap};
            CK(cudaMemset(dD, 0xff, 128 * 64 * 4));
            probe_mn_kernel<<<1, 128, smem_bytes>>>(a, tmAs, tmBs);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("PROBE %-40s CUDA-ERROR %s\n", c.name, cudaGetErrorString(e)); return 4; }
            CK(cudaMemcpy(hD.data(), dD, 128 * 64 * 4, cudaMemcpyDeviceToHost));
            double maxerr = 0, maxerr_lo = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 64; ++n) {
                    double s = 0;
                    for (int k = 0; k < c.K; ++k) s += (double)hA[(k + c.shift + (m >= 64)) * 64 + (m & 63)] * hB[k * 64 + n];
                    double d = fabs(s - hD[m * 64 + n]);
                    if (!(d <= maxerr)) maxerr = d;
                    if (m < 64 && !(d <= maxerr_lo)) maxerr_lo = d;
                }
            printf("PROBE %-40s max_err=%.3e (rows<64: %.3e) %s\n", c.name, maxerr, maxerr_lo, maxerr < 1e-3 ? "PASS" : "FAIL");
        }
    }
    // accumulation behaviour: repeat the same K=64 product `it` times into one TMEM accumulator and compare
    // with it * (exact product): round-to-nearest accumulation errs like sqrt(steps) ulp with random sign,
    // truncation errs like 0.5 * steps ulp, always toward zero
    for (int iters : {1, 16, 64, 256, 1024}) {
        ProbeArgs a{dA, dB, dD, 64, strip, 0, 2, 0, 0, iters, dcyc};
        probe_kernel<<<1, 128, smem_bytes>>>(a, tmA, tmB);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hD.data(), dD, 128 * 64 * 4, cudaMemcpyDeviceToHost));
        double rel_sum = 0, rel_max = 0, toward_zero = 0;
        int cnt = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 64; ++n) {
                double s = 0;
                for (int k = 0; k < 64; ++k) s += (double)hA[m * 64 + k] * hB[n * 64 + k];
                double want = s * iters, got = hD[m * 64 + n];
                if (fabs(want) < 1.0) continue;
                double rel = (got - want) / want;       // negative = magnitude too small = toward zero
                rel_sum += rel; if (fabs(rel) > rel_max) rel_max = fabs(rel);
                toward_zero += rel < 0; ++cnt;
            }
        printf("PROBE accumulate_x%-5d mma_steps=%-5d mean_rel_err=%+.3e max_rel_err=%.3e frac_toward_zero=%.2f\n", iters, iters * 4,
               rel_sum / cnt, rel_max, toward_zero / cnt);
    }
    // does a half-height MMA (M = 64) take half the cycles?  (it would let the unpaired fifth filter tap run at M = 64)
    for (int n : {256, 128}) {
        ProbeArgs a{dA, dB, dD, n, strip, 0, 2, 0, 0, 2000, dcyc, 64};
        probe_kernel<<<1, 128, smem_bytes>>>(a, tmA, tmB);
        CK(cudaDeviceSynchronize());
        long long cyc;
        CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
        printf("PROBE time_sw128_M64_N%-3d              cycles_per_mma=%.1f (M=128 takes %d)\n", n, (double)cyc / (2000.0 * 4), n / 2);
    }
    printf("PROBE done\n");
    return 0;
}
