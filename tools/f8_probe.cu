// Stand-alone probe for the fp8-correction candidate: can the two 2^-11 correction passes of
// the conv run as tcgen05.mma kind::f8f6f4 (e4m3 operands, 2x the rate of kind::f16) INTO THE SAME fp32 TMEM accumulator as
// the fp16 main pass, with the row-shifted window trick of k_conv_tc on 64-byte pixel rows?
// Built here (nvcc sm_100a), run on the B200 box via gpurun; results of the round-1 run are in profiles/r01_f8_probe.txt
// (all checked cases exact; ~129 cycles per e4m3 MMA at N = 256, 171 back to back on 64-byte rows):
//   case 1  e4m3 x e4m3, no-swizzle K-major layout, K = 64 (two K = 32 MMAs), N = 64 and 256            -> exact vs CPU
//   case 2  128-byte rows (128 fp8 per row), 128B swizzle, window start shifted by 0..9 rows, base_offset 0 -> exact vs CPU
//   case 3  64-byte rows (one pixel = 64 channels), 64B swizzle (layout code 4, 8-row atom = 512 B), shifted windows
//   case 4  kind::f16 MMA followed by kind::f8f6f4 MMAs into one accumulator (accumulate = 1)            -> sum of both products
//   case 6  the 64-byte-row strip filled by TMA (UINT16 x 32 per row, SWIZZLE_64B, box of 136 rows, negative start row = zero fill)
//   case 5  cycles per MMA: f8 at N = 256 (K = 32 per instruction, so equal cycles = twice the fp16 rate), and f16 / f8 interleaved
// Operands are small integers / powers of two so that every product and sum is exact in fp32 (no tolerance needed).
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../voicesplit_b200/csrc/sm100_ptx.cuh"

using namespace vs;
using namespace vs::ptx;

#define make_idesc_f8 make_idesc_e4m3   // shared with the product kernels (sm100_ptx.cuh)

struct F8Args {
    const uint8_t* A8;    // [128][K] e4m3
    const uint8_t* B8;    // [strip_rows][K] e4m3
    const __half* A16;    // [128][64] fp16 (case 4)
    const __half* B16;    // [N][64]
    float* D;             // [128][N]
    int N, K;             // K = bytes per row of the fp8 operands (64 or 128)
    int strip_rows, shift;
    int layout;           // 0 none, 2 = 128B swizzle (K = 128), 4 = 64B swizzle (K = 64)
    int with_f16;         // case 4: one fp16 K = 64 product first
    int iters, interleave;
    long long* cycles;
    int tma_fill;         // case 6: the B strip comes from TMA (64B swizzle, 64-byte rows) starting at global row `tma_row0` (may be < 0)
    int tma_row0;
};

// byte offset of 16-byte chunk c of row r in the given layout (rows of `rowbytes` bytes, `rows` rows in the tile)
__device__ __forceinline__ uint32_t chunk_off(int layout, int r, int c, int rowbytes, int rows) {
    if (layout == 2) return (uint32_t)(r * 128 + ((c ^ (r & 7)) * 16));             // Swizzle<3,4,3>
    if (layout == 4) return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) * 16));        // Swizzle<2,4,3>
    return (uint32_t)(c * (rows * 16) + r * 16);                                     // no swizzle: [chunk][row] core-matrix columns
}

__global__ void __launch_bounds__(128, 1) f8_probe_kernel(F8Args a, const __grid_constant__ CUtensorMap tmB) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                 // 128 x K bytes (<= 16 KB)
    uint8_t* sB = smem + 16384;         // strip_rows x K bytes (<= 34 KB)
    uint8_t* sA16 = smem + 16384 + 36864;   // 128 x 128 B fp16, 128B swizzle
    uint8_t* sB16 = sA16 + 16384;           // N x 128 B
    __shared__ uint64_t bar_mma, bar_load;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar_mma, 1); mbar_init(&bar_load, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(&tmem_base_s, 256); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const int cpr = a.K / 16;           // 16-byte chunks per row
    for (int idx = tid; idx < 128 * cpr; idx += 128) {
        const int r = idx / cpr, c = idx % cpr;
        *reinterpret_cast<uint4*>(sA + chunk_off(a.layout, r, c, a.K, 128)) = reinterpret_cast<const uint4*>(a.A8)[idx];
    }
    if (!a.tma_fill) {
        for (int idx = tid; idx < a.strip_rows * cpr; idx += 128) {
            const int r = idx / cpr, c = idx % cpr;
            *reinterpret_cast<uint4*>(sB + chunk_off(a.layout, r, c, a.K, a.strip_rows)) = reinterpret_cast<const uint4*>(a.B8)[idx];
        }
    }
    if (a.with_f16) {
        for (int idx = tid; idx < 128 * 8; idx += 128) {
            const int r = idx >> 3, c = idx & 7;
            *reinterpret_cast<uint4*>(sA16 + chunk_off(2, r, c, 128, 128)) = reinterpret_cast<const uint4*>(a.A16)[idx];
        }
        for (int idx = tid; idx < a.N * 8; idx += 128) {
            const int r = idx >> 3, c = idx & 7;
            *reinterpret_cast<uint4*>(sB16 + chunk_off(2, r, c, 128, a.N)) = reinterpret_cast<const uint4*>(a.B16)[idx];
        }
    }
    fence_proxy_async();
    __syncthreads();
    if (a.tma_fill) {   // rows tma_row0 .. tma_row0 + strip_rows of the global [rows][64 B] plane; rows outside are zero-filled
        if (tid == 0) {
            mbar_arrive_expect_tx(&bar_load, (uint32_t)(a.strip_rows * 64));
            for (int done = 0; done < a.strip_rows; done += 136) tma_load_2d(sB + done * 64, &tmB, &bar_load, 0, a.tma_row0 + done);
        }
        mbar_wait(&bar_load, 0);
    }

    long long t0 = 0;
    if (tid == 0) {
        const uint32_t id8 = make_idesc_f8(128, a.N), id16 = make_idesc_bf16(128, a.N, 1);
        const uint32_t aaddr = smem_u32(sA), baddr = smem_u32(sB);
        tc_fence_after();
        t0 = clock64();
        uint32_t acc = 0;
        for (int it = 0; it < a.iters; ++it) {
            if (a.with_f16 && (it == 0 || a.interleave)) {
                for (int k = 0; k < 4; ++k) {      // fp16 main pass: K = 64 as four K = 16 MMAs on 128B-swizzled tiles
                    umma_bf16(tmem, make_smem_desc(smem_u32(sA16) + k * 32, 16, 1024, 2), make_smem_desc(smem_u32(sB16) + k * 32, 16, 1024, 2), id16, acc);
                    acc = 1;
                }
            }
            for (int k = 0; k < a.K / 32; ++k) {   // fp8: K = 32 bytes per MMA
                uint64_t da, db;
                if (a.layout == 2) {
                    da = make_smem_desc(aaddr + k * 32, 16, 1024, 2, 0);
                    db = make_smem_desc(baddr + a.shift * 128 + k * 32, 16, 1024, 2, 0);
                } else if (a.layout == 4) {
                    da = make_smem_desc(aaddr + k * 32, 16, 512, 4, 0);
                    db = make_smem_desc(baddr + a.shift * 64 + k * 32, 16, 512, 4, 0);
                } else {   // no swizzle: LBO = distance between K-adjacent core matrices, SBO = between 8-row groups
                    da = make_smem_desc(aaddr + k * 2 * (128 * 16), 128 * 16, 128, 0, 0);
                    db = make_smem_desc(baddr + a.shift * 16 + k * 2 * (a.strip_rows * 16), a.strip_rows * 16, 128, 0, 0);
                }
                umma_f8(tmem, da, db, id8, acc);
                acc = 1;
            }
        }
        umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    if (tid == 0 && a.cycles) *a.cycles = clock64() - t0;
    tc_fence_after();
    for (int c0 = 0; c0 < a.N; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
        const int row = warp * 32 + (tid & 31);
        for (int j = 0; j < 32; ++j)
            if (c0 + j < a.N) a.D[(size_t)row * a.N + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

static uint8_t to_e4m3(float v) { return (uint8_t)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3); }
static float from_e4m3(uint8_t b) {
    __half_raw h = __nv_cvt_fp8_to_halfraw(b, __NV_E4M3);
    return __half2float(*reinterpret_cast<__half*>(&h));
}

static int run_case(const char* name, int N, int K, int layout, int shift, int with_f16, int iters, int interleave, bool check, int tma_fill = 0,
                    int tma_row0 = 0) {
    const int strip = tma_fill ? 272 : N + 16;          // TMA case: two boxes of 136 rows
    std::vector<uint8_t> A8((size_t)128 * K), B8((size_t)strip * K);
    std::vector<__half> A16(128 * 64), B16((size_t)N * 64);
    srand(1234 + N + K + layout + shift);
    auto small = [](int lim) { return (float)((rand() % (2 * lim + 1)) - lim) * 0.5f; };   // multiples of 0.5 up to +-lim/2: exact in e4m3
    for (auto& v : A8) v = to_e4m3(small(6));
    for (auto& v : B8) v = to_e4m3(small(6));
    for (auto& v : A16) v = __float2half(small(40));
    for (auto& v : B16) v = __float2half(small(40));
    uint8_t *dA8, *dB8; __half *dA16, *dB16; float* dD; long long* dC;
    cudaMalloc(&dA8, A8.size()); cudaMalloc(&dB8, B8.size()); cudaMalloc(&dA16, A16.size() * 2); cudaMalloc(&dB16, B16.size() * 2);
    cudaMalloc(&dD, (size_t)128 * N * 4); cudaMalloc(&dC, 8);
    cudaMemcpy(dA8, A8.data(), A8.size(), cudaMemcpyHostToDevice); cudaMemcpy(dB8, B8.data(), B8.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dA16, A16.data(), A16.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB16, B16.data(), B16.size() * 2, cudaMemcpyHostToDevice);
    F8Args a{dA8, dB8, dA16, dB16, dD, N, K, strip, shift, layout, with_f16, iters, interleave, dC, tma_fill, tma_row0};
    CUtensorMap tmB;
    memset(&tmB, 0, sizeof(tmB));
    if (tma_fill) {   // TMA only moves bits: 64 bytes per row = 32 uint16, 64B swizzle, box of 136 rows
        uint64_t dims[2] = {32, (uint64_t)strip}, strides[1] = {64};
        uint32_t box[2] = {32, 136};
        if (!make_tmap_bf16(&tmB, dB8, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B)) { printf("tensor map failed\n"); return 1; }
    }
    const int smem = 1024 + 16384 + 36864 + 16384 + 32768;
    cudaFuncSetAttribute(f8_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    f8_probe_kernel<<<1, 128, smem>>>(a, tmB);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-44s CUDA error: %s\n", name, cudaGetErrorString(e)); return 1; }
    std::vector<float> D((size_t)128 * N);
    long long cyc = 0;
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    if (check) {
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
                double s = 0.0;
                const int brow = n + shift + (tma_fill ? tma_row0 : 0);      // global row of the plane this window row maps to
                for (int k = 0; k < K; ++k)
                    s += (double)from_e4m3(A8[(size_t)m * K + k]) * ((brow >= 0 && brow < strip) ? from_e4m3(B8[(size_t)brow * K + k]) : 0.f);
                s *= iters;
                if (with_f16)
                    for (int k = 0; k < 64; ++k) s += (double)__half2float(A16[m * 64 + k]) * __half2float(B16[(size_t)n * 64 + k]) * (interleave ? iters : 1);
                if (fabs(s - D[(size_t)m * N + n]) > 1e-3 * (1.0 + fabs(s))) ++bad;
            }
    }
    const int n_mma = iters * (K / 32) + (with_f16 ? (interleave ? iters : 1) * 4 : 0);
    printf("%-44s N=%3d K=%3d shift=%d  %s  cycles/MMA %.1f (%d MMAs)\n", name, N, K, shift, check ? (bad ? "MISMATCH" : "ok") : "(timing)", (double)cyc / n_mma, n_mma);
    if (bad) printf("    %d of %d elements differ\n", bad, 128 * N);
    cudaFree(dA8); cudaFree(dB8); cudaFree(dA16); cudaFree(dB16); cudaFree(dD); cudaFree(dC);
    return bad ? 1 : 0;
}

int main() {
    int fails = 0;
    fails += run_case("1 e4m3 no-swizzle", 64, 64, 0, 0, 0, 1, 0, true);
    fails += run_case("1 e4m3 no-swizzle", 256, 64, 0, 0, 0, 1, 0, true);
    for (int sh : {0, 1, 2, 5, 9}) fails += run_case("2 e4m3 128B swizzle, shifted window", 256, 128, 2, sh, 0, 1, 0, true);
    for (int sh : {0, 1, 2, 3, 4, 7, 9}) fails += run_case("3 e4m3 64B swizzle (64-byte pixel rows)", 256, 64, 4, sh, 0, 1, 0, true);
    fails += run_case("4 fp16 main + e4m3 correction, one accumulator", 256, 64, 4, 2, 1, 1, 0, true);
    fails += run_case("4 fp16 + e4m3, 128B rows", 256, 128, 2, 2, 1, 1, 0, true);
    for (int sh : {0, 3, 8}) fails += run_case("6 TMA-filled 64B-swizzle strip, rows from -5", 256, 64, 4, sh, 0, 1, 0, true, 1, -5);
    fails += run_case("6 TMA-filled strip + fp16 main pass", 256, 64, 4, 2, 1, 1, 0, true, 1, 3);
    run_case("5 timing e4m3 N=256 (64B rows)", 256, 64, 4, 0, 0, 256, 0, false);
    run_case("5 timing e4m3 N=256 (128B rows)", 256, 128, 2, 0, 0, 128, 0, false);
    run_case("5 timing fp16 + e4m3 interleaved", 256, 64, 4, 0, 1, 128, 1, false);
    printf(fails ? "FAILED cases: %d\n" : "all checked cases ok\n", fails);
    return fails ? 1 : 0;
}
