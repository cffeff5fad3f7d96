// Tensor-core conv stack: cnn2..cnn7 as a tcgen05 implicit GEMM, plus the CUDA-core cnn1 / cnn8
// kernels that feed and drain the bf16 activation planes.
//
// Data layout.  Activation planes are channels-last bf16 [B][Q][64] with Q = T*Fp flattened
// pixels per utterance (Fp = padded_freq(F); pixels f >= F of every row are zero), stored as a
// `hi` plane and - in the fp32-faithful BF16X3 mode - a `lo` plane with x ~= hi + lo.
// A +-2 shift along F is a +-2 shift of the flat pixel index (it lands in the zero pad), a shift of
// dt rows along T is a shift of dt*Fp, and anything outside [0, Q) is zero-filled by TMA, which
// together are the reference's ZeroPad2d (models/voicesplit/model.py:16-47).
//
// GEMM mapping (per CTA tile of N consecutive flat pixels):
//     D[128][N] (TMEM, fp32) += A[128][64] (weights, smem) * B[N][64]^T (pixels, smem)
// A row 2*co+h holds tap (dt, df = 2j+h) of output channel co: one MMA applies TWO filter taps
// to the same N input pixels (M = 128 keeps the tensor pipe at full rate although the layer has
// only 64 output channels).  The two halves belong to output pixels one apart, so the epilogue
// forms out[co][p] = D[2co][p] + D[2co+1][p+1] with one warp shuffle (adjacent lanes).
// B is an N-row window into a strip of N+8 pixel rows that TMA loaded once per (dt, plane): the
// five df taps reuse the strip by moving the window start (row-shifted 128B-swizzle descriptor).
//
// Warp roles (576 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-17 = epilogue
// (TMEM -> registers -> BN-fold + activation -> 16-bit hi/lo -> global; four warps per TMEM lane
// quadrant, each taking every fourth 32-column chunk, so every SM sub-partition has four warps to
// hide the MUFU/shuffle latency of the Mish epilogue).  Two TMEM accumulators of N columns
// double-buffer MMA against the epilogue; the CTA is persistent over tiles.
//
// Weight multicast.  Every CTA walks the same sequence of weight tiles (15 steps x 32 KB per 5x5 tile = 480 KB, more than the
// 338 KB of pixel strips), and once the MMA work per tile drops to 8 slots per step (FP16_F8C) the L2 -> SM fill, not the tensor
// pipe, bounds the kernel (chip-wide L2 throughput ~6.3 KB/clk = 43 B/clk/SM).  CTAs are therefore launched as thread-block
// clusters that share ONE fetch of each weight tile: CTA r of the cluster loads rows [r, r+1) * 128/csz of the tile and TMA
// multicasts them into every member's ring slot; a slot is refilled only after the MMA threads of ALL members have committed
// it (tcgen05.commit multicast onto every member's w_empty barrier, count = cluster size).
#include "tc.cuh"
#include "sm100_ptx.cuh"
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <stdlib.h>
#include <type_traits>

namespace vs {
using namespace ptx;

constexpr int kWStages = 5;             // weight tiles (16 KB each) in flight
// Epilogue warps: 4 per TMEM lane quadrant for the flat tiles (N = 256: 8 chunks of 32 columns, two per warp).  The 2-D tiles of
// cnn2 (N = 224: 7 chunks, 4 paired steps) have a quarter of the MMA work per tile, so the epilogue sets their pace; a same-box
// sweep of 8 / 12 / 16 / 20 / 28 warps gave 10.65 / 10.87 / 10.15 / 10.65 / 10.92 M cycles per launch: 16 (<= 113 registers) it is.
#ifndef VS_EPI_WARPS_2D
#define VS_EPI_WARPS_2D 16
#endif
constexpr int kEpiWarpsFlat = 16, kEpiWarps2D = VS_EPI_WARPS_2D;
constexpr int conv_threads(int ew) { return 64 + 32 * ew; }
constexpr int kWTileBytes = 128 * 128;  // 128 rows x 64 bf16

struct ConvTcArgs {
    int Q, F, Fp, B;
    int N;                // MMA N (pixels per tile incl. the one lost to the pair shift)
    int tiles_per_utt, total_tiles;
    int n_dt, n_j, halo;  // taps along T, tap pairs along F, strip halo (2 for 5x5, 0 for 7x1)
    int dt_stride;        // dilation * Fp: flat-pixel offset of one tap step along T
    int passes;           // 1 (single 16-bit pass) or 3 (two operand planes: hi/lo split, or hi + fp8 correction plane)
    int f8c;              // VS_PREC_FP16_F8C: second plane = e4m3 correction operands, 4 f16 + 4 f8f6f4 MMAs per tap pair
    int strip_rows, box_rows, n_boxes, s_stages;
    int csz, n_iter;      // cluster size sharing the weight fetches; tile iterations every CTA runs (the same for all: lock step)
    // 2-D tiles for the kw = 1 layer (cnn2, 7 taps along T): a tile is `tr` frames x 8 bins (N = 8 tr pixels), its input strip the
    // (tr + kh - 1) x 8 block around it, loaded ONCE per plane with a 4-D TMA box and shared by all kh taps: tap dt is the window
    // starting dt * 8 rows (= dt swizzle atoms) into the strip.  Flat tiles would fetch a fresh strip per tap (7x the bytes).
    int tile2d, tr, n_ft, T;
    unsigned int n_ft_magic;   // 2^32 / n_ft + 1: tin / n_ft as a multiply-high (tin < tiles per utterance)
    int tr_out, t_halo;   // 2-D tiles: output frames per tile (tr - 1: the upper tap of a pair lands one frame up), (kh - 1) / 2
    int l2_prefetch;      // flat tiles: prefetch the next tile's strips into L2 (VOICESPLIT_CONV_L2PREFETCH=1; off by default: measured no gain)
    int act;
    const float* scale;
    const float* shift;
    elt16* out_hi;
    elt16* out_lo;  // may be null (single-pass modes)
    float* out32;   // OUT32 kernels (training): fp32 plane [B][Q][64] instead of the 16-bit planes
};

template <int ACT>
__device__ __forceinline__ float act_fast(float x) {
    if (ACT == 2) return x;   // pass-through: raw conv output (training forward before BatchNorm, data gradients)
    if (ACT == VS_ACT_RELU) return fmaxf(x, 0.f);
    return mish_f(x);
}

// GEO = 1 / 2: the flat 5x5 layers in a two-plane mode (fp16_f8c, fp16x3, bf16x3; cnn3..7 = 89 % of the FLOPs, also the training
// forward and data gradient) get an MMA-issue loop whose schedule is a compile-time constant (see the issuer below), without (1) /
// with (2) the 2-CTA weight multicast; GEO = 0 is the general loop (any geometry / precision / cluster size); GEO = 3: 2-D tiles
// (general loop, the epilogue pairs taps along T).
template <int ACT, int ELT, bool OUT32, bool F8C, int EW, int GEO = 0>
__global__ void __launch_bounds__(conv_threads(EW), 1) k_conv_tc(const ConvTcArgs a, const __grid_constant__ CUtensorMap tm_in_hi,
                                                    const __grid_constant__ CUtensorMap tm_in_lo,
                                                    const __grid_constant__ CUtensorMap tm_w_hi,
                                                    const __grid_constant__ CUtensorMap tm_w_lo) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int strip_bytes = a.strip_rows * 128;
    uint8_t* w_ring = smem;
    uint8_t* s_ring = smem + kWStages * kWTileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_ring + (size_t)a.s_stages * strip_bytes);
    uint64_t* w_full = bars;               // [kWStages]
    uint64_t* w_empty = bars + kWStages;   // [kWStages]
    uint64_t* s_full = bars + 2 * kWStages;             // [s_stages]
    uint64_t* s_empty = s_full + a.s_stages;            // [s_stages]
    uint64_t* acc_full = s_empty + a.s_stages;          // [2]
    uint64_t* acc_empty = acc_full + 2;                 // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = a.csz > 1 ? cluster_ctarank() : 0u;
    const uint16_t cmask = (uint16_t)((1u << a.csz) - 1u);
    if (threadIdx.x == 0) {
        for (int i = 0; i < kWStages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], (uint32_t)a.csz); }
        for (int i = 0; i < a.s_stages; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], EW); }
        fence_barrier_init();
    }
    if (warp == 0) {
        if (lane == 0) {
            prefetch_tensormap(&tm_in_hi); prefetch_tensormap(&tm_w_hi);
            if (a.passes == 3) { prefetch_tensormap(&tm_in_lo); prefetch_tensormap(&tm_w_lo); }
        }
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    if (a.csz > 1) cluster_sync_all();     // peers' barriers are initialised before anyone multicasts onto them
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int useful = a.tile2d ? 8 * a.tr_out : a.N - 1;   // the tap-pair shift costs one pixel (flat) / one frame (2-D)
    // 3-pass modes: per tap row both strips (hi, lo) are resident; each weight tile is fetched once:
    // W_hi[j] multiplies S_hi and S_lo, W_lo[j] multiplies S_hi  (hi*hi + lo*hi + hi*lo).
    const int n_strip_loads = a.passes == 3 ? 2 : 1;

    if (warp == 0) {
        // ===================== TMA producer =====================
        // whole warp: warp-uniform loops and barrier waits; one elected lane arms the barrier and issues the TMA loads
        if (GEO == 1 || GEO == 2) {
            // Fixed schedule of the flat 5x5 two-plane layers (see the MMA issuer): slot and parity of every weight-ring wait are
            // compile-time constants, the 10 strips of a tile walk the 4-slot strip ring.  The general loop below kept this warp
            // busy ~75 % of the time with ring / trip-count bookkeeping - barely ahead of the tensor pipe, so the weight ring was
            // rarely full when the issuer needed it.
            constexpr bool MC = GEO == 2;
            constexpr int kSliceRows = MC ? 64 : 128;
            const int w_row0 = MC ? (int)crank * kSliceRows : 0;
            uint8_t* const w_dst0 = w_ring + (size_t)w_row0 * 128;
            uint32_t scount = 0;
            for (int itn = 0; itn < a.n_iter; ++itn) {
                const int tile = blockIdx.x + itn * gridDim.x;
                const bool valid = tile < a.total_tiles;    // a padding iteration still takes part in the shared weight stream
                const int b = valid ? tile / a.tiles_per_utt : 0;
                const int q0 = (tile - b * a.tiles_per_utt) * useful;
#pragma unroll
                for (int dt = 0; dt < 5; ++dt) {
                    if (valid) {
                        const int qs = q0 - a.halo + (dt - 2) * a.dt_stride;
#pragma unroll
                        for (int sp = 0; sp < 2; ++sp) {
                            const uint32_t k = scount + 2u * dt + sp, st = k & 3u;
                            mbar_wait(&s_empty[st], ((k >> 2) & 1u) ^ 1u);
                            if (elect_one()) {
                                mbar_arrive_expect_tx(&s_full[st], (uint32_t)strip_bytes);
                                uint8_t* dst = s_ring + (size_t)st * strip_bytes;
                                for (int i = 0; i < a.n_boxes; ++i)
                                    tma_load_3d(dst + (size_t)i * a.box_rows * 128, sp == 0 ? &tm_in_hi : &tm_in_lo, &s_full[st], 0,
                                                qs + i * a.box_rows, b);
                            }
                            __syncwarp();
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
#pragma unroll
                        for (int wp = 0; wp < 2; ++wp) {
                            const int t = 2 * (dt * 3 + j) + wp;                   // constants after unrolling
                            mbar_wait(&w_empty[t % kWStages], ((t / kWStages) & 1) ^ 1);    // released by the MMA threads of all cluster members
                            if (elect_one()) {
                                mbar_arrive_expect_tx(&w_full[t % kWStages], kWTileBytes);
                                const void* tm = wp == 0 ? (const void*)&tm_w_hi : (const void*)&tm_w_lo;
                                if (MC) tma_load_2d_mc(w_dst0 + (size_t)(t % kWStages) * kWTileBytes, tm, &w_full[t % kWStages], 0,
                                                       (dt * 3 + j) * 128 + w_row0, cmask);
                                else tma_load_2d(w_dst0 + (size_t)(t % kWStages) * kWTileBytes, tm, &w_full[t % kWStages], 0, (dt * 3 + j) * 128);
                            }
                            __syncwarp();
                        }
                    }
                }
                if (valid) scount += 10u;
            }
            // drain: every arrival peers still owe this CTA's w_empty barriers has landed before the CTA may exit
            if (MC)
                for (int i = 0; i < kWStages; ++i) mbar_wait(&w_empty[i], 1);
        } else {
            int ws = 0, wph = 0, ss = 0, sph = 0;
            const int slice_rows = 128 / a.csz;
            for (int itn = 0; itn < a.n_iter; ++itn) {
                const int tile = blockIdx.x + itn * gridDim.x;
                const bool valid = tile < a.total_tiles;    // a padding iteration still takes part in the shared weight stream
                const int b = valid ? tile / a.tiles_per_utt : 0;
                const int q0 = (tile - b * a.tiles_per_utt) * useful;
                if (a.tile2d && valid) {      // one (tr + kh - 1) x 8 block per plane serves every tap of the tile
                    const int tin = tile - b * a.tiles_per_utt, tt = tin / a.n_ft, ft = tin - tt * a.n_ft;
                    for (int sp = 0; sp < n_strip_loads; ++sp) {
                        mbar_wait(&s_empty[ss], sph ^ 1);
                        if (elect_one()) {
                            mbar_arrive_expect_tx(&s_full[ss], (uint32_t)strip_bytes);
                            tma_load_4d(s_ring + (size_t)ss * strip_bytes, sp == 0 ? &tm_in_hi : &tm_in_lo, &s_full[ss], 0, ft * 8,
                                        tt * a.tr_out - a.t_halo, b);
                        }
                        __syncwarp();
                        if (++ss == a.s_stages) { ss = 0; sph ^= 1; }
                    }
                }
                // Flat tiles: only two tap rows of strips fit next to the weight ring, so a strip is requested ~2.5 MMA steps before
                // it is needed - enough for an L2 hit, not for a DRAM miss (the activation planes are the only DRAM-sourced
                // operand).  Pull the strips of this CTA's NEXT tile into L2 now, a whole tile (~15 us) ahead.
                if (!a.tile2d && a.l2_prefetch) {
                    const int ntile = tile + (int)gridDim.x;
                    if (ntile < a.total_tiles && elect_one()) {
                        const int nb = ntile / a.tiles_per_utt;
                        const int nq0 = (ntile - nb * a.tiles_per_utt) * useful;
                        for (int dt = 0; dt < a.n_dt; ++dt) {
                            const int qs = nq0 - a.halo + (dt - a.n_dt / 2) * a.dt_stride;
                            for (int i = 0; i < a.n_boxes; ++i) {
                                tma_prefetch_l2_3d(&tm_in_hi, 0, qs + i * a.box_rows, nb);
                                if (n_strip_loads == 2) tma_prefetch_l2_3d(&tm_in_lo, 0, qs + i * a.box_rows, nb);
                            }
                        }
                    }
                    __syncwarp();
                }
                for (int dt = 0; dt < a.n_dt; ++dt) {
                    const int qs = q0 - a.halo + (dt - a.n_dt / 2) * a.dt_stride;
                    for (int sp = 0; valid && !a.tile2d && sp < n_strip_loads; ++sp) {  // strip planes of this tap row: hi (, lo)
                        mbar_wait(&s_empty[ss], sph ^ 1);
                        if (elect_one()) {
                            mbar_arrive_expect_tx(&s_full[ss], (uint32_t)strip_bytes);
                            uint8_t* dst = s_ring + (size_t)ss * strip_bytes;
                            for (int i = 0; i < a.n_boxes; ++i)
                                tma_load_3d(dst + (size_t)i * a.box_rows * 128, sp == 0 ? &tm_in_hi : &tm_in_lo, &s_full[ss], 0,
                                            qs + i * a.box_rows, b);
                        }
                        __syncwarp();
                        if (++ss == a.s_stages) { ss = 0; sph ^= 1; }
                    }
                    // weight tiles of this tap row: W_hi[j] (used against both strips), then W_lo[j]
                    for (int j = 0; j < a.n_j; ++j) {
                        for (int wp = 0; wp < n_strip_loads; ++wp) {
                            mbar_wait(&w_empty[ws], wph ^ 1);        // released by the MMA threads of all cluster members
                            if (elect_one()) {
                                mbar_arrive_expect_tx(&w_full[ws], kWTileBytes);
                                const void* tm = wp == 0 ? (const void*)&tm_w_hi : (const void*)&tm_w_lo;
                                if (a.csz > 1)
                                    tma_load_2d_mc(w_ring + (size_t)ws * kWTileBytes + (size_t)crank * slice_rows * 128, tm, &w_full[ws], 0,
                                                   (dt * a.n_j + j) * 128 + (int)crank * slice_rows, cmask);
                                else
                                    tma_load_2d(w_ring + (size_t)ws * kWTileBytes, tm, &w_full[ws], 0, (dt * a.n_j + j) * 128);
                            }
                            __syncwarp();
                            if (++ws == kWStages) { ws = 0; wph ^= 1; }
                        }
                    }
                }
            }
            // drain: every arrival peers still owe this CTA's w_empty barriers has landed before the CTA may exit
            if (a.csz > 1)
                for (int i = 0; i < kWStages; ++i) {
                    mbar_wait(&w_empty[ws], wph ^ 1);
                    if (++ws == kWStages) { ws = 0; wph ^= 1; }
                }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The WHOLE warp runs the (warp-uniform) control flow and the barrier waits; one elected lane issues the tcgen05
        // instructions.  Keeping the warp converged matters: tcgen05.mma / commit take their operands from uniform registers,
        // and inside an `if (lane == 0)` region the compiler wraps every one of them in an elect-and-loop sequence that costs
        // ~100 cycles per MMA - as much as the MMA itself (ncu source view, profiles/r02_ncu_conv_f8c_b256_before_elect_summary.txt).
        {
            const uint32_t idesc = make_idesc_bf16(128, a.N, ELT);
            const uint32_t idesc8 = make_idesc_e4m3(128, a.N);
            const uint64_t w_desc0 = make_smem_desc(smem_u32(w_ring), 16, 1024, 2);     // slot 0 of the weight ring / strip ring
            const uint64_t s_desc0 = make_smem_desc(smem_u32(s_ring), 16, 1024, 2);
            if (GEO == 1 || GEO == 2) {
                constexpr bool MC = GEO == 2;             // commits are multicast to both CTAs of the cluster
                // Fixed schedule of a 5x5 fp16_f8c tile: 15 steps x (W_hi tile, e4m3 tile) = 30 weight tiles = exactly 6 turns of the
                // 5-slot weight ring, so slot AND parity of every weight wait are compile-time constants and the ring state is
                // the same at every tile start; 10 strips per tile walk the 4-slot strip ring ((count + k) & 3).  Fully unrolled:
                // the issuing warp executes the waits, 8 MMAs and 2 commits per step and nothing else - in the general loop its
                // ~100 bookkeeping instructions per step (ring wraps, runtime trip counts, constant-bank reloads; single warp,
                // ~10 cycles each) took as long as the 8 MMAs themselves (ncu source view, r02 final conv capture).
                constexpr int kSteps = 15;
                static_assert(kWStages == 5 && (2 * kSteps) % kWStages == 0 && ((2 * kSteps) / kWStages) % 2 == 0, "weight ring schedule");
                const uint32_t strip_d = (uint32_t)(strip_bytes >> 4);
                uint32_t scount = 0;                       // strips consumed so far (s_stages == 4, checked by the launcher)
                int it = 0;
                for (int itn = 0; itn < a.n_iter; ++itn) {
                    if (blockIdx.x + itn * gridDim.x >= a.total_tiles) {
                        // padding iteration: no pixels, but the cluster's shared weight stages must still be consumed and released
#pragma unroll
                        for (int t = 0; t < 2 * kSteps; ++t) {
                            mbar_wait(&w_full[t % kWStages], (t / kWStages) & 1);
                            if (elect_one()) { if (MC) umma_commit_mc(&w_empty[t % kWStages], cmask); else umma_commit(&w_empty[t % kWStages]); }
                            __syncwarp();
                        }
                        continue;
                    }
                    const int buf = it & 1, aph = (it >> 1) & 1;
                    ++it;
                    mbar_wait(&acc_empty[buf], aph ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem + (uint32_t)(buf * a.N);
#pragma unroll
                    for (int dt = 0; dt < 5; ++dt) {
                        const uint32_t k0 = scount + 2u * dt, k1 = k0 + 1u;
                        const uint32_t st0 = k0 & 3u, st1 = k1 & 3u;
                        mbar_wait(&s_full[st0], (k0 >> 2) & 1u);
                        mbar_wait(&s_full[st1], (k1 >> 2) & 1u);
                        const uint64_t sd0 = s_desc0 + (uint64_t)(st0 * strip_d), sd1 = s_desc0 + (uint64_t)(st1 * strip_d);
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            constexpr int kDummy = 0; (void)kDummy;
                            const int t0 = 2 * (dt * 3 + j), t1 = t0 + 1;          // constants after unrolling
                            const uint64_t b0 = sd0 + (uint64_t)(16 * j), b1 = sd1 + (uint64_t)(16 * j);
                            mbar_wait(&w_full[t0 % kWStages], (t0 / kWStages) & 1);
                            tc_fence_after();
                            if (elect_one()) {
                                const uint64_t a0 = w_desc0 + (uint64_t)((t0 % kWStages) * (kWTileBytes >> 4));
#pragma unroll
                                for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, a0 + 2 * k, b0 + 2 * k, idesc, (dt == 0 && j == 0 && k == 0) ? 0u : 1u);
                                if (!F8C) {                  // split operands: W_hi against both strips (hi*hi + lo*hi)
#pragma unroll
                                    for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, a0 + 2 * k, b1 + 2 * k, idesc, 1u);
                                }
                                if (MC) umma_commit_mc(&w_empty[t0 % kWStages], cmask); else umma_commit(&w_empty[t0 % kWStages]);
                            }
                            __syncwarp();
                            mbar_wait(&w_full[t1 % kWStages], (t1 / kWStages) & 1);
                            tc_fence_after();
                            if (elect_one()) {
                                const uint64_t a1 = w_desc0 + (uint64_t)((t1 % kWStages) * (kWTileBytes >> 4));
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    if (F8C) umma_f8(d_tmem, a1 + 2 * k, b1 + 2 * k, idesc8, 1u);      // e4m3 correction tile x c8 strip
                                    else umma_bf16(d_tmem, a1 + 2 * k, b0 + 2 * k, idesc, 1u);         // hi*lo: W_lo x S_hi
                                }
                                if (MC) umma_commit_mc(&w_empty[t1 % kWStages], cmask); else umma_commit(&w_empty[t1 % kWStages]);
                            }
                            __syncwarp();
                        }
                        if (elect_one()) { umma_commit(&s_empty[st0]); umma_commit(&s_empty[st1]); }
                        __syncwarp();
                    }
                    scount += 10u;
                    if (elect_one()) umma_commit(&acc_full[buf]);
                    __syncwarp();
                }
            } else {
            int ws = 0, wph = 0, ss = 0, sph = 0, it = 0;
            for (int itn = 0; itn < a.n_iter; ++itn) {
                if (blockIdx.x + itn * gridDim.x >= a.total_tiles) {
                    // padding iteration: no pixels, but the cluster's shared weight stages must still be consumed and released
                    for (int n = a.n_dt * a.n_j * n_strip_loads; n > 0; --n) {
                        mbar_wait(&w_full[ws], wph);
                        if (elect_one()) { if (a.csz > 1) umma_commit_mc(&w_empty[ws], cmask); else umma_commit(&w_empty[ws]); }
                        __syncwarp();
                        if (++ws == kWStages) { ws = 0; wph ^= 1; }
                    }
                    continue;
                }
                const int buf = it & 1, aph = (it >> 1) & 1;
                ++it;
                mbar_wait(&acc_empty[buf], aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem + (uint32_t)(buf * a.N);
                uint32_t accumulate = 0;
                // UMMA descriptors are built ONCE (ring base) and advanced by integer adds on the 14-bit (address >> 4) field:
                // the issuing warp executes ~10 dependent-latency cycles per instruction, so every instruction between two MMA
                // batches is tensor-pipe idle time (ncu source view: ~140 instructions per 8 MMAs before this, pipe 57-75 % active)
                uint64_t s_desc[2] = {0, 0};
                int s_stage[2] = {0, 0};
                for (int dt = 0; dt < a.n_dt; ++dt) {
                    // strips of this tap row: hi in stage ss (, lo in the next stage); 2-D tiles: one strip pair for all taps,
                    // tap pair s reads the window 2 s frames (16 rows) further down
                    if (!a.tile2d || dt == 0) {
                        for (int sp = 0; sp < n_strip_loads; ++sp) {
                            mbar_wait(&s_full[ss], sph);
                            s_desc[sp] = s_desc0 + (uint64_t)((uint32_t)ss * (uint32_t)(strip_bytes >> 4));
                            s_stage[sp] = ss;
                            if (++ss == a.s_stages) { ss = 0; sph ^= 1; }
                        }
                    } else {
                        for (int sp = 0; sp < n_strip_loads; ++sp) s_desc[sp] += (16 * 128) >> 4;   // next tap pair: 2 frames = 16 rows down
                    }
                    for (int j = 0; j < a.n_j; ++j) {
                        // One wait + one elected issue block PER weight tile, written out twice (not a loop: the compiler then keeps two
                        // straight-line blocks of four MMAs).  The MMAs on W_hi start while the second tile of the step - fetched by the
                        // other CTA of the cluster - may still be in flight; waiting for both tiles first cost 5 % more cycles per
                        // 5x5 layer, a predicated loop over the two tiles 38 % (same-box A/B, profiles/r02_conv_issue_loop_ab.txt).
                        const uint64_t b0 = s_desc[0] + (uint64_t)(16 * j);      // window start 2 j pixel rows further: 2 j * 128 B >> 4
                        const uint64_t b1 = s_desc[1] + (uint64_t)(16 * j);
                        mbar_wait(&w_full[ws], wph);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t a0 = w_desc0 + (uint64_t)(ws * (kWTileBytes >> 4));
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, a0 + 2 * k, b0 + 2 * k, idesc, k == 0 ? accumulate : 1u);
                            if (!F8C && n_strip_loads == 2) {      // hi*hi + lo*hi: W_hi against both strips
#pragma unroll
                                for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, a0 + 2 * k, b1 + 2 * k, idesc, 1);
                            }
                            if (a.csz > 1) umma_commit_mc(&w_empty[ws], cmask); else umma_commit(&w_empty[ws]);
                        }
                        __syncwarp();
                        if (++ws == kWStages) { ws = 0; wph ^= 1; }
                        if (n_strip_loads == 2) {
                            mbar_wait(&w_full[ws], wph);
                            tc_fence_after();
                            if (elect_one()) {
                                const uint64_t a1 = w_desc0 + (uint64_t)(ws * (kWTileBytes >> 4));
                                if (F8C) {
                                    // the e4m3 correction tile x the c8 strip as four kind::f8f6f4 MMAs (K = 32 bytes each):
                                    // x_lo*w_hi over bytes 0..63, x_hi*w_lo over bytes 64..127
#pragma unroll
                                    for (int k = 0; k < 4; ++k) umma_f8(d_tmem, a1 + 2 * k, b1 + 2 * k, idesc8, 1);
                                } else {
#pragma unroll
                                    for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, a1 + 2 * k, b0 + 2 * k, idesc, 1);     // hi*lo: W_lo x S_hi
                                }
                                if (a.csz > 1) umma_commit_mc(&w_empty[ws], cmask); else umma_commit(&w_empty[ws]);
                            }
                            __syncwarp();
                            if (++ws == kWStages) { ws = 0; wph ^= 1; }
                        }
                        __syncwarp();
                        accumulate = 1;
                    }
                    if (!a.tile2d || dt == a.n_dt - 1) {
                        if (elect_one())
                            for (int sp = 0; sp < n_strip_loads; ++sp) umma_commit(&s_empty[s_stage[sp]]);
                        __syncwarp();
                    }
                }
                if (elect_one()) umma_commit(&acc_full[buf]);
                __syncwarp();
            }
            }
        }
    } else {
        // ===================== epilogue (warps 2 ..) =====================
        // The instruction stream of these warps is what bounds the light layer (cnn2 on 2-D tiles: ~7.7 k cycles per tile against
        // 3.6 k cycles of MMA; ncu: issue + MIO bound, profiles/r02_ncu_final_cnn2_f8c_b256_summary.txt), so it is kept branch-free and
        // short: interior chunks (no row padding, no tile / plane edge: the common case) and the cut last chunk take paths without
        // per-pixel bookkeeping, store addresses are one base pointer per chunk plus compile-time offsets, tile coordinates advance
        // without divisions, the activation is nine straight-line instructions (common.cuh: mish_f), each lane converts and stores
        // a channel PAIR (one pack instruction per plane).
        constexpr bool T2D = (GEO == 3);                 // 2-D tiles (the 7x1 layer with taps paired along T)
        const int quad = warp & 3;                       // TMEM lane quadrant this warp may read
        const int cgrp = (warp - 2) >> 2;                // which 32-column chunks this warp takes
        const int co = quad * 16 + (lane >> 1), h = lane & 1;
        const float sc = a.scale[co], sh = a.shift[co];
        const bool want_lo = !OUT32 && !F8C && a.out_lo != nullptr;
        // 16-bit outputs: lanes L and L^2 hold channels co, co^1 of the same pixels; they swap every other value so that
        // each lane stores a CHANNEL PAIR (4 bytes per plane) of every second pixel instead of 2 bytes of every pixel
        const int codd = (lane >> 1) & 1;
        const size_t row_px = T2D ? (size_t)a.Fp : 0;    // 2-D tiles: pixels between consecutive frames
        int it = 0;
        // tile = (utterance b, tile tin inside it), advanced without divisions: on the 2-D layer every warp handles ONE chunk per tile,
        // so per-tile bookkeeping is not amortised (two integer divisions were 14 % of the epilogue's samples, r02 cnn2 capture)
        int b = (int)blockIdx.x / a.tiles_per_utt, tin = (int)blockIdx.x - b * a.tiles_per_utt;
        const int step_b = (int)gridDim.x / a.tiles_per_utt, step_t = (int)gridDim.x - step_b * a.tiles_per_utt;
        for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1, aph = (it >> 1) & 1;
            int q0 = 0, t0 = 0, f0 = 0;                  // flat: first pixel; 2-D: first frame / bin (column p = frame t0 + p / 8, bin f0 + p % 8)
            if (T2D) {
                const int tt = (int)(((unsigned long long)(unsigned)tin * a.n_ft_magic) >> 32);     // tin / n_ft
                t0 = tt * a.tr_out; f0 = (tin - tt * a.n_ft) * 8;
            } else q0 = tin * useful;
            const size_t plane0 = (size_t)b * a.Q;       // first pixel of the utterance's plane
            tin += step_t; b += step_b;
            if (tin >= a.tiles_per_utt) { tin -= a.tiles_per_utt; ++b; }
            mbar_wait(&acc_full[buf], aph);
            tc_fence_after();
            const uint32_t t_base = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * a.N);
            for (int c0 = cgrp * 32; c0 < a.N; c0 += 32 * (EW / 4)) {
                uint32_t r[32];
                uint32_t nxt = 0;                         // flat: column c0 + 32 (the upper tap of the chunk's last pixel)
                uint32_t rx[8] = {0, 0, 0, 0, 0, 0, 0, 0};      // 2-D: columns c0 + 32 .. c0 + 39 (the upper tap sits one frame = 8 columns on)
                tmem_ld_32x32(t_base + c0, r);
                if (T2D) {
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                 : "=r"(rx[0]), "=r"(rx[1]), "=r"(rx[2]), "=r"(rx[3]), "=r"(rx[4]), "=r"(rx[5]), "=r"(rx[6]), "=r"(rx[7])
                                 : "r"(t_base + c0 + 32) : "memory");
                } else if (c0 + 32 < a.N) {
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(nxt) : "r"(t_base + c0 + 32) : "memory");
                }
                // pixel of this lane's FIRST value of the chunk (column c0 + h) and whether the whole chunk is interior
                size_t px0;
                bool interior;
                int f = 0;
                // columns of this chunk that are output pixels of the tile (the last chunk is cut by the tap-pair shift: 31 flat, 24 2-D)
                const int ncols = useful - c0 < 32 ? useful - c0 : 32;
                if (T2D) {
                    const int tr0 = t0 + (c0 >> 3);       // the chunk covers frames tr0 .. tr0 + 3, bins f0 .. f0 + 7
                    px0 = (size_t)tr0 * a.Fp + f0 + h;
                    interior = tr0 + (ncols >> 3) <= a.T && f0 + 8 <= a.F;
                } else {
                    px0 = (size_t)q0 + c0 + h;
                    const int fc = (q0 + c0) % a.Fp;      // bin of the chunk's first column; rows are Fp pixels, bins >= F are padding
                    f = fc + h; if (f >= a.Fp) f -= a.Fp;
                    interior = fc + 33 <= a.F && q0 + c0 + 32 <= a.Q;
                }
                elt16* ohi = OUT32 ? nullptr : a.out_hi + (plane0 + px0) * 64 + (co & ~1);
                elt16* olo = want_lo ? a.out_lo + (plane0 + px0) * 64 + (co & ~1) : nullptr;
                uint8_t* oc8 = (!OUT32 && F8C) ? reinterpret_cast<uint8_t*>(a.out_lo) + (plane0 + px0) * 128 + (co & ~1) : nullptr;
                float* o32 = OUT32 ? a.out32 + (plane0 + px0) * 64 + co : nullptr;
                if (!OUT32) {        // channel-odd lanes keep the SECOND pixel of every pair (2 pixels on, same frame): fold it into the bases
                    const size_t lane_px = codd ? 2 : 0;
                    ohi += lane_px * 64;
                    if (want_lo) olo += lane_px * 64;
                    if (F8C) oc8 += lane_px * 128;
                }
                tmem_ld_wait();
                // FAST: no padding bin, plane edge or frame edge in the chunk; FULL: all 32 columns are output pixels (else the first ncols)
                auto body = [&](auto fast_tag, auto full_tag) {
                    constexpr bool FAST = decltype(fast_tag)::value, FULL = decltype(full_tag)::value;
#pragma unroll
                    for (int m = 0; m < 16; m += 2) {
                        if (FAST && !FULL && 2 * m >= ncols) break;     // warp-uniform: the cut chunk stops early
                        float y[2];
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            // even lane (h=0, lower tap) owns pixel c0+2(m+i), odd lane (upper tap) pixel c0+2(m+i)+1
                            const int mm = m + i;
                            float acc;
                            if (T2D) {
                                // out[n] = D[2co][n] + D[2co+1][n + 8]: the even lane (row 2co) needs its neighbour's column 2mm + 8,
                                // the odd lane (row 2co+1, pixel 2mm + 1) needs the even lane's column 2mm + 1 and its own 2mm + 9
                                auto col = [&](int c) { return __uint_as_float(c < 32 ? r[c & 31] : rx[(c - 32) & 7]); };
                                const float recv = __shfl_xor_sync(0xffffffffu, h ? col(2 * mm + 8) : col(2 * mm + 1), 1);
                                acc = h == 0 ? col(2 * mm) + recv : recv + col(2 * mm + 9);
                            } else {
                                const float mine_odd = __uint_as_float(r[2 * mm + 1]);
                                const float other_odd = __shfl_xor_sync(0xffffffffu, mine_odd, 1);
                                const float up_next = __uint_as_float(mm < 15 ? r[(2 * mm + 2) & 31] : nxt);
                                acc = h == 0 ? __uint_as_float(r[2 * mm]) + other_odd : other_odd + up_next;
                            }
                            y[i] = act_fast<ACT>(fmaf(acc, sc, sh));
                            if (!FAST) {     // padding bins hold zeros (select, not a branch)
                                const int fcur = T2D ? f0 + 2 * (mm & 3) + h : f;
                                y[i] *= fcur < a.F ? 1.f : 0.f;
                                f += 2;
                                if (f >= a.Fp) f -= a.Fp;
                            }
                        }
                        // pixel offset (from px0) of value mm: flat 2 mm; 2-D (mm / 4) frames + 2 (mm % 4) bins - compile-time but for Fp
                        auto off = [&](int mm) { return T2D ? (size_t)(mm >> 2) * row_px + 2 * (mm & 3) : (size_t)(2 * mm); };
                        auto valid = [&](int mm) {
                            if (FAST && FULL) return true;
                            if (FAST) return 2 * mm + (T2D ? 0 : h) < ncols;
                            if (T2D) return c0 + 2 * mm < useful && t0 + ((c0 + 2 * mm) >> 3) < a.T;   // bins: every tile covers 8 in-plane bins
                            const int p = c0 + 2 * mm + h;
                            return p < useful && q0 + p < a.Q;
                        };
                        if (OUT32) {
#pragma unroll
                            for (int i = 0; i < 2; ++i)
                                if (valid(m + i)) o32[off(m + i) * 64] = y[i];
                        } else {
                            // channel-even lane keeps the first pixel, channel-odd lane the second; each receives the partner channel
                            const float recv = __shfl_xor_sync(0xffffffffu, codd ? y[0] : y[1], 2);
                            const float v0 = codd ? recv : y[0], v1 = codd ? y[1] : recv;      // channels (co & ~1), (co | 1)
                            // the kept pixel is value m (channel-even lanes) or m + 1 (channel-odd lanes: already in the lane's base pointers)
                            const size_t o = off(m);
                            const bool ok = codd ? valid(m + 1) : valid(m);
                            if (ok) {
                                if (F8C) {
                                    uint32_t hi2;
                                    unsigned short l8, x8;
                                    split_f8c_pair(v0, v1, hi2, l8, x8);
                                    *reinterpret_cast<uint32_t*>(ohi + o * 64) = hi2;
                                    *reinterpret_cast<unsigned short*>(oc8 + o * 128) = l8;
                                    *reinterpret_cast<unsigned short*>(oc8 + o * 128 + 64) = x8;
                                } else {
                                    uint32_t hi2, lo2;
                                    split16_pair<ELT>(v0, v1, hi2, lo2);
                                    *reinterpret_cast<uint32_t*>(ohi + o * 64) = hi2;
                                    if (want_lo) *reinterpret_cast<uint32_t*>(olo + o * 64) = lo2;
                                }
                            }
                        }
                    }
                };
                if (interior) { if (ncols == 32) body(std::true_type{}, std::true_type{}); else body(std::true_type{}, std::false_type{}); }
                else body(std::false_type{}, std::false_type{});
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (a.csz > 1) cluster_sync_all();     // nobody leaves while a peer may still multicast into its shared memory
    if (warp == 1) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// cnn1 on CUDA cores, writing the bf16 hi/lo planes (K = 7, C_in = 1: not MMA-shaped)
// ---------------------------------------------------------------------------------------------
template <int ACT, int ELT, bool F8C>
__global__ void __launch_bounds__(256, 3) k_front_tc(const float* __restrict__ x, elt16* __restrict__ hi,
                                                     elt16* __restrict__ lo, const float* __restrict__ w,
                                                     const float* __restrict__ scale, const float* __restrict__ shift,
                                                     int F, int Fp, int nrows) {
    // one block per (utterance, frame) row; thread = 4 output channels x 2 adjacent pixels, its 28 filter taps
    // live in registers for the whole row (4 channels keep the register count low enough for 3 blocks per SM:
    // the kernel is bound by dependent-issue latency of the Mish / split epilogue, not by bandwidth)
    extern __shared__ float xs[];   // [Fp + 8]: x[f - 3] at xs[f]
    const int tid = threadIdx.x, cg = tid & 15, pp = tid >> 4;
    float wr[7][4], sc[4], sh[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        sc[c] = scale[cg * 4 + c]; sh[c] = shift[cg * 4 + c];
#pragma unroll
        for (int j = 0; j < 7; ++j) wr[j][c] = w[j * 64 + cg * 4 + c];
    }
    for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
        const float* xrow = x + (size_t)row * F;
        __syncthreads();
        for (int i = tid; i < Fp + 8; i += 256) {
            int f = i - 3;
            xs[i] = (f >= 0 && f < F) ? xrow[f] : 0.f;
        }
        __syncthreads();
        for (int f0 = pp * 2; f0 < Fp; f0 += 32) {
            float xv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) xv[i] = xs[f0 + i];
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const int f = f0 + p;
                float yv[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float acc = 0.f;
#pragma unroll
                    for (int j = 0; j < 7; ++j) acc = fmaf(wr[j][c], xv[p + j], acc);
                    yv[c] = act_fast<ACT>(fmaf(acc, sc[c], sh[c])) * (f < F ? 1.f : 0.f);
                }
                const size_t o = ((size_t)row * Fp + f) * 64 + cg * 4;
                if (F8C) {   // `lo` is the c8 plane: 128 bytes per pixel, [l8 x 64 | x8 x 64]
                    uint32_t h01, h23;
                    unsigned short l01, l23, x01, x23;
                    split_f8c_pair(yv[0], yv[1], h01, l01, x01);
                    split_f8c_pair(yv[2], yv[3], h23, l23, x23);
                    *reinterpret_cast<uint2*>(hi + o) = make_uint2(h01, h23);
                    uint8_t* c8 = reinterpret_cast<uint8_t*>(lo) + ((size_t)row * Fp + f) * 128 + cg * 4;
                    *reinterpret_cast<uint32_t*>(c8) = (uint32_t)l01 | ((uint32_t)l23 << 16);
                    *reinterpret_cast<uint32_t*>(c8 + 64) = (uint32_t)x01 | ((uint32_t)x23 << 16);
                } else {
                    uint32_t h01, h23, l01, l23;
                    split16_pair<ELT>(yv[0], yv[1], h01, l01);
                    split16_pair<ELT>(yv[2], yv[3], h23, l23);
                    *reinterpret_cast<uint2*>(hi + o) = make_uint2(h01, h23);
                    if (lo) *reinterpret_cast<uint2*>(lo + o) = make_uint2(l01, l23);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// cnn8 (64 -> 8, 1x1) + BN + act from the bf16 planes; output row layout [B*T][ldx] with column
// c*F+f, as fp32 and/or bf16 hi/lo (the LSTM input-projection operand)
// ---------------------------------------------------------------------------------------------
template <int ACT, bool F8C>
__global__ void __launch_bounds__(256) k_point8_tc(const elt16* __restrict__ hi, const elt16* __restrict__ lo, int elt,
                                                   const float* __restrict__ w, const float* __restrict__ scale,
                                                   const float* __restrict__ shift, float* __restrict__ x32,
                                                   elt16* __restrict__ xhi, elt16* __restrict__ xlo, int ldx,
                                                   int F, int Fp, long long nplane) {
    // 256 consecutive plane pixels per block.  The 2 x 32 KB of channel data are staged through shared
    // memory with fully coalesced 16-byte copies (rows padded to 144 B: conflict-free 16-byte reads),
    // then one thread per pixel does the 64 -> 8 contraction.
    extern __shared__ __align__(16) uint8_t stage[];
    __shared__ __align__(16) float ws[64 * 8];
    __shared__ float sc[8], sh[8];
    const int tid = threadIdx.x;
    for (int i = tid; i < 512; i += 256) ws[i] = w[i];
    if (tid < 8) { sc[tid] = scale[tid]; sh[tid] = shift[tid]; }
    const long long p0 = (long long)blockIdx.x * 256;
    uint8_t* s_hi = stage;
    uint8_t* s_lo = stage + 256 * 144;
    const int nplanes = lo ? 2 : 1;
    for (int pl = 0; pl < nplanes; ++pl) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(pl == 0 ? hi : lo) + p0 * 128;
        uint8_t* dst = pl == 0 ? s_hi : s_lo;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int chunk = i * 256 + tid;
            const int px = chunk >> 3, part = chunk & 7;
            if (F8C && pl == 1 && part >= 4) continue;      // c8 rows: only the l8 half (bytes 0..63) is read here
            if (p0 + px < nplane) {
                const uint32_t sa = (uint32_t)__cvta_generic_to_shared(dst + px * 144 + part * 16);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + (size_t)chunk * 16) : "memory");
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const long long p = p0 + tid;
    if (p >= nplane) return;
    const int f = (int)(p % Fp);
    if (f >= F) return;
    const long long bt = p / Fp;
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
#pragma unroll 2
    for (int q = 0; q < 8; ++q) {
        const uint4 vh = *reinterpret_cast<const uint4*>(s_hi + tid * 144 + q * 16);
        const elt16* ph = reinterpret_cast<const elt16*>(&vh);
        uint4 vl = make_uint4(0, 0, 0, 0);
        uint2 v8 = make_uint2(0, 0);
        if (F8C) v8 = *reinterpret_cast<const uint2*>(s_lo + tid * 144 + q * 8);   // l8 of channels 8q..8q+7 (first half of the c8 row)
        else if (lo) vl = *reinterpret_cast<const uint4*>(s_lo + tid * 144 + q * 16);
        const elt16* pl = reinterpret_cast<const elt16*>(&vl);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float v = F8C ? __half2float(__ushort_as_half(ph[k])) +
                                      e4m3_to_float((k < 4 ? v8.x : v8.y) >> (8 * (k & 3))) * (1.f / kF8cLoScale)
                                : join16(ph[k], pl[k], elt);
            const float4 w0 = *reinterpret_cast<const float4*>(ws + (q * 8 + k) * 8);
            const float4 w1 = *reinterpret_cast<const float4*>(ws + (q * 8 + k) * 8 + 4);
            acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]); acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
            acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]); acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
        }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float y = act_fast<ACT>(fmaf(acc[c], sc[c], sh[c]));
        size_t o = (size_t)bt * ldx + (size_t)c * F + f;
        if (x32) x32[o] = y;
        if (xhi) {
            elt16 yh, yl;
            split16_rt(y, elt, yh, yl);
            xhi[o] = yh;
            if (xlo) xlo[o] = yl;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// cnn8 on tensor cores.  The channels-last plane IS a K-major A operand: D[128 pixels][16 (8 used)] = A[128 px][64 ch] * W8[16][64]^T,
// so a tile is one TMA box of 128 pixel rows and 4 small MMAs per operand plane; the kernel then runs at the plane's HBM read rate
// instead of the issue rate of 512 FMAs + 128 shared-memory weight loads per pixel.
//   mode 0  one 16-bit pass              a_hi * w_hi
//   mode 1  split operands, 3 passes     a_hi * w_hi + a_lo * w_hi + a_hi * w_lo
//   mode 2  VS_PREC_FP16_F8C             a_hi * w_hi + a_hi * w_lo (both kind::f16: the weight tile is 2 KB, a second fp16 pass costs nothing
//                                        here) + l8 * (e4m3(2^-8 w_hi) + e4m3 of its rounding residual) over the 64-byte l8 half of the
//                                        c8 row (x8 is not read: 192 B / pixel)
// Warp 0: TMA producer (ring of pixel tiles), warp 1: MMA issuer (4 accumulators of 16 TMEM columns), warps 2..13: three epilogue
// groups of 4 warps (thread = one pixel x 8 channels: BN + act, 16-bit hi / lo of the LSTM operand at column c * F + f).
// ---------------------------------------------------------------------------------------------
constexpr int kP8Groups = 3, kP8Acc = 4;
constexpr int kP8Threads = 64 + 128 * kP8Groups;
struct Point8Args {
    int mode, stages, stage_bytes, n_tiles;
    long long nplane;
    int F, Fp, ldx;
    const float *scale, *shift;      // per output channel; scale already divided by the power-of-two weight scale
    float* x32;
    elt16 *xhi, *xlo;
};
template <int ACT, int ELT>
__global__ void __launch_bounds__(kP8Threads, 1) k_point8_mma(const Point8Args a, const __grid_constant__ CUtensorMap tm_hi,
                                                               const __grid_constant__ CUtensorMap tm_lo,
                                                               const __grid_constant__ CUtensorMap tm_w16,
                                                               const __grid_constant__ CUtensorMap tm_w8) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* w_hi = smem;                 // [16 rows][128 B] 128B swizzle
    uint8_t* w_lo = smem + 2048;
    uint8_t* w_8 = smem + 4096;           // [16 rows][64 B] e4m3(2^-8 w_hi), 64B swizzle
    uint8_t* w_8b = smem + 5120;          // the same for the e4m3 rounding residual of that tile (weight error 2^-8 instead of 2^-4)
    uint8_t* ring = smem + 8192;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)a.stages * a.stage_bytes);
    uint64_t* a_full = bars;
    uint64_t* a_empty = bars + a.stages;
    uint64_t* acc_full = a_empty + a.stages;
    uint64_t* acc_empty = acc_full + kP8Acc;
    uint64_t* w_full = acc_empty + kP8Acc;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < a.stages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < kP8Acc; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        mbar_init(w_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, 64); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            mbar_arrive_expect_tx(w_full, 4096u + (a.mode == 2 ? 2048u : 0u));
            tma_load_2d(w_hi, &tm_w16, w_full, 0, 0);
            tma_load_2d(w_lo, &tm_w16, w_full, 0, 16);
            if (a.mode == 2) { tma_load_2d(w_8, &tm_w8, w_full, 0, 0); tma_load_2d(w_8b, &tm_w8, w_full, 0, 16); }
        }
        __syncwarp();
        int st = 0, ph = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
            mbar_wait(&a_empty[st], ph ^ 1);
            if (elect_one()) {
                uint8_t* dst = ring + (size_t)st * a.stage_bytes;
                mbar_arrive_expect_tx(&a_full[st], (uint32_t)a.stage_bytes);
                tma_load_2d(dst, &tm_hi, &a_full[st], 0, tile * 128);
                if (a.mode != 0) tma_load_2d(dst + 16384, &tm_lo, &a_full[st], 0, tile * 128);
            }
            __syncwarp();
            if (++st == a.stages) { st = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        const uint32_t id16 = make_idesc_bf16(128, 16, ELT), id8 = make_idesc_e4m3(128, 16);
        mbar_wait(w_full, 0);
        tc_fence_after();
        const uint64_t d_wh = make_smem_desc(smem_u32(w_hi), 16, 1024, 2), d_wl = make_smem_desc(smem_u32(w_lo), 16, 1024, 2);
        const uint64_t d_w8 = make_smem_desc(smem_u32(w_8), 16, 512, 4), d_w8b = make_smem_desc(smem_u32(w_8b), 16, 512, 4);
        int st = 0, ph = 0, it = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
            const int b = it % kP8Acc;
            mbar_wait(&a_full[st], ph);
            mbar_wait(&acc_empty[b], ((it / kP8Acc) & 1) ^ 1);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(ring + (size_t)st * a.stage_bytes);
            if (elect_one()) {
                const uint64_t d_ah = make_smem_desc(a_addr, 16, 1024, 2);
                const uint32_t d_t = tmem + (uint32_t)(b * 16);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(d_t, d_ah + 2 * k, d_wh + 2 * k, id16, k ? 1u : 0u);
                if (a.mode == 1) {
                    const uint64_t d_al = make_smem_desc(a_addr + 16384, 16, 1024, 2);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(d_t, d_al + 2 * k, d_wh + 2 * k, id16, 1u);
                }
                if (a.mode != 0) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(d_t, d_ah + 2 * k, d_wl + 2 * k, id16, 1u);
                }
                if (a.mode == 2) {
                    const uint64_t d_a8 = make_smem_desc(a_addr + 16384, 16, 512, 4);
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        umma_f8(d_t, d_a8 + 2 * k, d_w8 + 2 * k, id8, 1u);
                        umma_f8(d_t, d_a8 + 2 * k, d_w8b + 2 * k, id8, 1u);
                    }
                }
                umma_commit(&a_empty[st]);
                umma_commit(&acc_full[b]);
            }
            __syncwarp();
            if (++st == a.stages) { st = 0; ph ^= 1; }
        }
    } else {
        const int grp = (warp - 2) >> 2, quad = warp & 3;
        const int row = quad * 32 + lane;
        float sc[8], sh[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) { sc[c] = a.scale[c]; sh[c] = a.shift[c]; }
        int it = grp;
        for (int tile = blockIdx.x + grp * gridDim.x; tile < a.n_tiles; tile += kP8Groups * gridDim.x, it += kP8Groups) {
            const int b = it % kP8Acc;
            mbar_wait(&acc_full[b], (it / kP8Acc) & 1);
            tc_fence_after();
            uint32_t r[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                         : "r"(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * 16)) : "memory");
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[b]);      // the values are in registers: the accumulator may be overwritten
            const long long p = (long long)tile * 128 + row;
            if (p >= a.nplane) continue;
            const int f = (int)(p % a.Fp);
            if (f >= a.F) continue;
            const size_t o0 = (size_t)(p / a.Fp) * a.ldx + f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float y = act_fast<ACT>(fmaf(__uint_as_float(r[c]), sc[c], sh[c]));
                const size_t o = o0 + (size_t)c * a.F;
                if (a.x32) a.x32[o] = y;
                if (a.xhi) {
                    elt16 yh, yl;
                    split16<ELT>(y, yh, yl);
                    a.xhi[o] = yh;
                    if (a.xlo) a.xlo[o] = yl;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 64);
}

// fp32 plane [B][T][Fp][64] <-> bf16 hi/lo planes (debug hook)
__global__ void k_plane_split(const float* __restrict__ p, elt16* __restrict__ hi, elt16* __restrict__ lo, int elt, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    elt16 h, l;
    split16_rt(p[i], elt, h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
}
__global__ void k_plane_join(const elt16* __restrict__ hi, const elt16* __restrict__ lo, int elt, float* __restrict__ p, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    p[i] = join16(hi[i], lo ? lo[i] : (elt16)0, elt);
}

// the same for the hi + c8 plane pair of VS_PREC_FP16_F8C (the join reads what a consumer of the values sees: hi + 2^-8 l8)
__global__ void k_plane_split_f8c(const float* __restrict__ p, elt16* __restrict__ hi, uint8_t* __restrict__ c8, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    elt16 h;
    float l;
    split_f8c(p[i], h, l);
    hi[i] = h;
    const unsigned short pr = e4m3x2(kF8cLoScale * l, kF8cHiScale * p[i]);
    const long long px = i >> 6;
    const int ch = (int)(i & 63);
    c8[px * 128 + ch] = (uint8_t)(pr & 0xff);
    c8[px * 128 + 64 + ch] = (uint8_t)(pr >> 8);
}
__global__ void k_plane_join_f8c(const elt16* __restrict__ hi, const uint8_t* __restrict__ c8, float* __restrict__ p, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    p[i] = __half2float(__ushort_as_half(hi[i])) + e4m3_to_float(c8[(i >> 6) * 128 + (i & 63)]) * (1.f / kF8cLoScale);
}

// max |w| of a weight tensor (bits of a non-negative float order like unsigned ints)
__global__ void k_absmax(const float* __restrict__ w, int n, unsigned int* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicMax(out, __float_as_uint(fabsf(w[i])));
}
// power of two s with max|w| * s in [2^8, 2^9): keeps the `lo` halves of fp16 weights normal
__device__ __forceinline__ float pow2_scale(unsigned int maxbits) {
    float m = __uint_as_float(maxbits);
    if (!(m > 0.f) || !isfinite(m)) return 1.f;
    int ex;
    frexpf(m, &ex);  // m = f * 2^ex, f in [0.5, 1)
    return exp2f((float)(9 - ex));
}
// weights [tap][ci][co] fp32 -> tap-pair tiles [step][row = 2co+h][ci], scaled by s, as bf16 and fp16 hi / lo
// Tap carried by weight-tile row 2co+h of step `step` (-1: padding slot).  Default: the two taps of a row pair are neighbours
// along F (df = 2j, 2j+1 of filter row dt).  pair_t (the kw = 1 layer on 2-D tiles): neighbours along T (dt = 2 step, 2 step + 1).
__device__ __forceinline__ int conv_tile_tap(int step, int h, int kh, int kw, int n_j, int pair_t) {
    if (pair_t) { const int dt = 2 * step + h; return dt < kh ? dt * kw : -1; }
    const int dt = step / n_j, df = 2 * (step % n_j) + h;
    return df < kw ? dt * kw + df : -1;
}
__host__ __device__ inline int conv_tile_steps(int kh, int kw, int pair_t) { return pair_t ? (kh + 1) / 2 : kh * ((kw + 1) / 2); }
__global__ void k_pack_conv_tc(const float* __restrict__ w32, const unsigned int* __restrict__ maxbits, elt16* __restrict__ bhi,
                               elt16* __restrict__ blo, elt16* __restrict__ hhi, elt16* __restrict__ hlo, int kh, int kw, int n_j, int pair_t) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int n = conv_tile_steps(kh, kw, pair_t) * 128 * 64;
    if (i >= n) return;
    const float s = pow2_scale(*maxbits);
    int ci = i & 63, row = (i >> 6) & 127, step = i >> 13;
    int co = row >> 1, h = row & 1;
    const int tap = conv_tile_tap(step, h, kh, kw, n_j, pair_t);
    float v = tap >= 0 ? s * w32[((size_t)tap * 64 + ci) * 64 + co] : 0.f;
    split16<0>(v, bhi[i], blo[i]);
    split16<1>(v, hhi[i], hlo[i]);
}
// VS_PREC_FP16_F8C correction tiles, same [step][row = 2co+h] order, 128 bytes per row:
// [ e4m3(2^-8 * w_hi)[ci] x 64 | e4m3(2^2 * w_lo)[ci] x 64 ]  with  w = s * weight = w_hi (fp16) + w_lo
__global__ void k_pack_conv_f8c(const float* __restrict__ w32, const unsigned int* __restrict__ maxbits, uint8_t* __restrict__ w8, int kh, int kw,
                                int n_j, int pair_t) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int n = conv_tile_steps(kh, kw, pair_t) * 128 * 64;
    if (i >= n) return;
    const float s = pow2_scale(*maxbits);
    int ci = i & 63, row = (i >> 6) & 127, step = i >> 13;
    int co = row >> 1, h = row & 1;
    const int tap = conv_tile_tap(step, h, kh, kw, n_j, pair_t);
    float v = tap >= 0 ? s * w32[((size_t)tap * 64 + ci) * 64 + co] : 0.f;
    elt16 hi;
    float lo;
    split_f8c(v, hi, lo);
    const unsigned short pr = e4m3x2(__half2float(__ushort_as_half(hi)) * (1.f / kF8cLoScale), lo * (1.f / kF8cHiScale));
    uint8_t* dst = w8 + ((size_t)step * 128 + row) * 128;
    dst[ci] = (uint8_t)(pr & 0xff);
    dst[64 + ci] = (uint8_t)(pr >> 8);
}
// W8 fp32 [ci][8] -> rows [16 (8 used)][64 ci]: 16-bit hi / lo (rows 0..15 hi, 16..31 lo; bf16 and fp16 copies) and e4m3(2^-8 w_hi), scaled by s
__global__ void k_pack_point8(const float* __restrict__ w32, const unsigned int* __restrict__ maxbits, elt16* __restrict__ wb, elt16* __restrict__ wh,
                              uint8_t* __restrict__ w8) {
    const int i = threadIdx.x + blockIdx.x * blockDim.x;     // 16 * 64
    if (i >= 1024) return;
    const int ci = i & 63, r = i >> 6;
    const float v = r < 8 ? pow2_scale(*maxbits) * w32[ci * 8 + r] : 0.f;
    split16<0>(v, wb[i], wb[1024 + i]);
    split16<1>(v, wh[i], wh[1024 + i]);
    const float t = __half2float(__ushort_as_half(wh[i])) * (1.f / kF8cLoScale);
    const unsigned int q = e4m3x2(t, 0.f) & 0xffu;
    w8[i] = (uint8_t)q;
    w8[1024 + i] = (uint8_t)(e4m3x2(t - e4m3_to_float(q), 0.f) & 0xffu);     // second tile: what the first rounding lost
}

__global__ void k_scale_tc(const float* __restrict__ scale, const unsigned int* __restrict__ maxbits, float* __restrict__ out, int n) {
    int i = threadIdx.x;
    if (i < n) out[i] = scale[i] / pow2_scale(*maxbits);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct TcState {
    void* gemm = nullptr;     // GemmState of tc_gemm.cu (must stay the first member)
    void* lstm = nullptr;     // LstmState of tc_lstm.cu
    elt16* w_hi[2][8] = {};   // [elt][layer]
    elt16* w_lo[2][8] = {};
    float* scale_tc[8] = {};  // BN scale divided by the layer's power-of-two weight scale
    uint8_t* w_c8[8] = {};    // VS_PREC_FP16_F8C: e4m3 correction tiles (k_pack_conv_f8c)
    elt16* p8_w[2] = {};      // cnn8 on tensor cores: [elt][32 rows (16 hi, 16 lo)][64]
    uint8_t* p8_w8 = nullptr; // its e4m3(2^-8 w_hi) tile [16][64]
    int point8_mma = 1;       // VOICESPLIT_POINT8_MMA = 0 selects the CUDA-core cnn8 kernel
    int fixed_schedule = 1;   // VOICESPLIT_CONV_FIXED_SCHEDULE = 0: general MMA-issue loop for the 5x5 fp16_f8c layers too
    elt16* wT_hi[2][8] = {};  // training: data-gradient weights (taps flipped, channels transposed), same tile format
    elt16* wT_lo[2][8] = {};
    float* unscale[8] = {};   // 64 copies of 1 / (power-of-two weight scale): epilogue scale of the raw-output convs
    unsigned int* wmax = nullptr;
    int max_smem = 0;
    int cluster = 2;          // CTAs per cluster sharing the conv weight fetches (VOICESPLIT_CONV_CLUSTER = 1, 2, 4 or 8)
    int tile2d = 1;           // 2-D tiles for the kw = 1 layer (VOICESPLIT_CONV_TILE2D = 0 falls back to flat tiles)
    int tile2d_dgrad = 1;     // the same for the data-gradient use of that layer (VOICESPLIT_CONV_TILE2D_DGRAD)
    int l2_prefetch = 0;      // flat tiles: L2 prefetch of the next tile's strips (VOICESPLIT_CONV_L2PREFETCH = 1 enables; measured: no gain)
};

static int tile_n_for(const vs_engine*) { return 256; }
// the kw = 1 layer (cnn2) runs on 2-D tiles with its taps paired along T
static bool conv_pairs_rows(const TcState* s, const ConvGeom& g, bool dgrad = false) {
    return g.kw == 1 && g.dil == 1 && (dgrad ? s->tile2d_dgrad : s->tile2d);
}

int tc_create(vs_engine* e) {
    TcState* s = new TcState();
    e->tc = s;
    cudaDeviceGetAttribute(&s->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, e->device);
    if (const char* c = getenv("VOICESPLIT_CONV_TILE2D")) s->tile2d = s->tile2d_dgrad = atoi(c) != 0;
    if (const char* c = getenv("VOICESPLIT_CONV_TILE2D_DGRAD")) s->tile2d_dgrad = atoi(c) != 0;
    if (const char* c = getenv("VOICESPLIT_CONV_L2PREFETCH")) s->l2_prefetch = atoi(c) != 0;
    if (const char* c = getenv("VOICESPLIT_POINT8_MMA")) s->point8_mma = atoi(c) != 0;
    if (const char* c = getenv("VOICESPLIT_CONV_FIXED_SCHEDULE")) s->fixed_schedule = atoi(c) != 0;
    if (const char* c = getenv("VOICESPLIT_CONV_CLUSTER")) {
        const int v = atoi(c);
        if (v == 1 || v == 2 || v == 4 || v == 8) s->cluster = v;
    }
    return VS_OK;
}
void tc_destroy(vs_engine* e) {
    TcState* s = (TcState*)e->tc;
    if (!s) return;
    tc_gemm_destroy(e);
    tc_lstm_destroy(s->lstm);
    for (int l = 0; l < 8; ++l) {
        for (int t = 0; t < 2; ++t) { cudaFree(s->w_hi[t][l]); cudaFree(s->w_lo[t][l]); cudaFree(s->wT_hi[t][l]); cudaFree(s->wT_lo[t][l]); }
        cudaFree(s->scale_tc[l]); cudaFree(s->unscale[l]); cudaFree(s->w_c8[l]);
    }
    cudaFree(s->p8_w[0]); cudaFree(s->p8_w[1]); cudaFree(s->p8_w8);
    cudaFree(s->wmax);
    delete s;
    e->tc = nullptr;
}

int tc_pack(vs_engine* e, cudaStream_t st) {
    TcState* s = (TcState*)e->tc;
    if (!s->wmax) VS_CUDA_TRY(cudaMalloc(&s->wmax, 8 * sizeof(unsigned int)));
    VS_CUDA_TRY(cudaMemsetAsync(s->wmax, 0, 8 * sizeof(unsigned int), st));
    for (int l = 1; l <= 6; ++l) {
        const ConvGeom g = kConv[l];
        const int n_j = (g.kw + 1) / 2;
        const size_t n = (size_t)g.kh * n_j * 128 * 64;
        if (!s->w_hi[0][l]) {
            for (int t = 0; t < 2; ++t) {
                VS_CUDA_TRY(cudaMalloc(&s->w_hi[t][l], n * sizeof(elt16)));
                VS_CUDA_TRY(cudaMalloc(&s->w_lo[t][l], n * sizeof(elt16)));
                VS_CUDA_TRY(cudaMalloc(&s->wT_hi[t][l], n * sizeof(elt16)));
                VS_CUDA_TRY(cudaMalloc(&s->wT_lo[t][l], n * sizeof(elt16)));
            }
            VS_CUDA_TRY(cudaMalloc(&s->w_c8[l], n * 2));
            VS_CUDA_TRY(cudaMalloc(&s->scale_tc[l], 64 * sizeof(float)));
            VS_CUDA_TRY(cudaMalloc(&s->unscale[l], 64 * sizeof(float)));
        }
        const int nw = g.cout * g.cin * g.kh * g.kw;
        const int pair_t = conv_pairs_rows(s, g) ? 1 : 0;      // cnn2 on 2-D tiles: tap pairs along T
        k_absmax<<<(nw + 255) / 256, 256, 0, st>>>(e->conv_w32[l], nw, s->wmax + l);
        k_pack_conv_tc<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(e->conv_w32[l], s->wmax + l, s->w_hi[0][l], s->w_lo[0][l],
                                                                    s->w_hi[1][l], s->w_lo[1][l], g.kh, g.kw, n_j, pair_t);
        k_pack_conv_f8c<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(e->conv_w32[l], s->wmax + l, s->w_c8[l], g.kh, g.kw, n_j, pair_t);
        k_scale_tc<<<1, 64, 0, st>>>(e->conv_scale[l], s->wmax + l, s->scale_tc[l], 64);
        // training: the same power-of-two scale serves the flipped/transposed data-gradient weights (same values)
        k_pack_conv_tc<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(e->conv_wT32[l], s->wmax + l, s->wT_hi[0][l], s->wT_lo[0][l],
                                                                    s->wT_hi[1][l], s->wT_lo[1][l], g.kh, g.kw, n_j, conv_pairs_rows(s, g, true) ? 1 : 0);
        k_scale_tc<<<1, 64, 0, st>>>(e->ones64, s->wmax + l, s->unscale[l], 64);
    }
    {   // cnn8 (64 -> 8, 1x1) as an MMA B operand
        if (!s->p8_w[0]) {
            VS_CUDA_TRY(cudaMalloc(&s->p8_w[0], 2048 * sizeof(elt16)));
            VS_CUDA_TRY(cudaMalloc(&s->p8_w[1], 2048 * sizeof(elt16)));
            VS_CUDA_TRY(cudaMalloc(&s->p8_w8, 2048));
            VS_CUDA_TRY(cudaMalloc(&s->scale_tc[7], 64 * sizeof(float)));
        }
        k_absmax<<<2, 256, 0, st>>>(e->conv_w32[7], 512, s->wmax + 7);
        k_pack_point8<<<4, 256, 0, st>>>(e->conv_w32[7], s->wmax + 7, s->p8_w[0], s->p8_w[1], s->p8_w8);
        k_scale_tc<<<1, 64, 0, st>>>(e->conv_scale[7], s->wmax + 7, s->scale_tc[7], 8);
    }
    VS_CUDA_TRY(cudaGetLastError());
    int rc = tc_lstm_pack(e, &s->lstm, st);
    if (rc != VS_OK) return rc;
    return tc_gemm_pack(e, st);
}
void* tc_lstm_slot(vs_engine* e) { return ((TcState*)e->tc)->lstm; }

struct ConvTcCall {            // what differs between the eval layers and the training uses of the kernel
    const elt16 *w_hi, *w_lo;
    const float *scale, *shift;
    int act;                   // VS_ACT_* or 2 (pass-through)
    float* out32;              // non-null: fp32 output plane
    int kid;
    bool f8c = false;          // w_lo = e4m3 correction tiles, in_lo / out_lo = c8 planes
    bool dgrad = false;        // data-gradient use (flipped / transposed weights)
};
static int launch_conv_tc_ex(vs_engine* e, int layer, const elt16* in_hi, const elt16* in_lo, elt16* out_hi, elt16* out_lo, int B, int T,
                             int passes, int elt, const ConvTcCall& call, cudaStream_t st);

static int launch_conv_tc(vs_engine* e, int layer, const elt16* in_hi, const elt16* in_lo,
                          elt16* out_hi, elt16* out_lo, int B, int T, int precision, cudaStream_t st) {
    TcState* s = (TcState*)e->tc;
    const int elt = tc_elt(precision);
    ConvTcCall call{s->w_hi[elt][layer], s->w_lo[elt][layer], s->scale_tc[layer], e->conv_shift[layer], e->d.activation, nullptr,
                    KID_CONV1 + layer - 1};
    if (tc_f8c(precision)) { call.w_lo = reinterpret_cast<const elt16*>(s->w_c8[layer]); call.f8c = true; }
    return launch_conv_tc_ex(e, layer, in_hi, in_lo, out_hi, out_lo, B, T, tc_passes(precision), elt, call, st);
}

// training: raw (pre-BatchNorm) forward conv, or the data gradient (flipped / transposed weights), fp32 output
int tc_train_conv(vs_engine* e, int layer, bool dgrad, const elt16* in_hi, const elt16* in_lo, const float* shift, float* out32, int B, int T,
                  int elt, int kid, cudaStream_t st) {
    TcState* s = (TcState*)e->tc;
    ConvTcCall call{dgrad ? s->wT_hi[elt][layer] : s->w_hi[elt][layer], dgrad ? s->wT_lo[elt][layer] : s->w_lo[elt][layer], s->unscale[layer],
                    shift, 2, out32, kid};
    call.dgrad = dgrad;
    return launch_conv_tc_ex(e, layer, in_hi, in_lo, nullptr, nullptr, B, T, 3, elt, call, st);
}

static int launch_conv_tc_ex(vs_engine* e, int layer, const elt16* in_hi, const elt16* in_lo, elt16* out_hi, elt16* out_lo, int B, int T,
                             int passes, int elt, const ConvTcCall& call, cudaStream_t st) {
    TcState* s = (TcState*)e->tc;
    const ConvGeom g = kConv[layer];
    const int F = e->d.num_freq, Fp = padded_freq(F);
    ConvTcArgs a{};
    a.Q = T * Fp; a.F = F; a.Fp = Fp; a.B = B;
    a.N = tile_n_for(e);
    a.tiles_per_utt = (a.Q + a.N - 2) / (a.N - 1);
    a.total_tiles = B * a.tiles_per_utt;
    a.n_dt = g.kh; a.n_j = (g.kw + 1) / 2; a.halo = g.kw / 2;
    a.dt_stride = g.dil * Fp;
    a.passes = passes;
    a.f8c = call.f8c ? 1 : 0;
    if (call.f8c && (passes != 3 || !elt || call.out32 || !in_lo || !out_lo)) { set_error("fp16_f8c conv needs the fp16 hi + c8 plane pair"); return VS_ERR_INVALID; }
    a.T = T;
    a.l2_prefetch = s->l2_prefetch;
    // kw = 1 (cnn2): 2-D tiles of tr frames x 8 bins whose (tr + kh - 1) x 8 input block is loaded once for all kh taps
    a.tile2d = conv_pairs_rows(s, g, call.dgrad) ? 1 : 0;
    if (a.tile2d) {
        // tap pairs along T: row 2co+h of step s carries tap dt = 2s + h; both halves see the window that starts 2s frames into
        // the strip, so the upper tap's products belong to the output one frame up: out[r] = D[2co][r] + D[2co+1][r + 1] (a shift
        // of 8 accumulator columns).  4 steps instead of 7; a window of tr frames yields tr - 1 output frames.
        a.tr = 28;                                  // N = 224: two strip pairs of 34 x 8 rows + the weight ring fit shared memory
        a.tr_out = a.tr - 1;
        a.t_halo = g.kh / 2;
        a.n_dt = (g.kh + 1) / 2; a.n_j = 1;
        a.N = 8 * a.tr;
        a.n_ft = Fp / 8;
        a.n_ft_magic = (unsigned int)((1ull << 32) / (unsigned)a.n_ft + 1);
        a.tiles_per_utt = a.n_ft * ((T + a.tr_out - 1) / a.tr_out);
        a.total_tiles = B * a.tiles_per_utt;
    }
    a.strip_rows = a.tile2d ? (a.tr + g.kh - 1) * 8 : a.N + 8;
    a.box_rows = a.strip_rows;
    a.n_boxes = 1;
    while (!a.tile2d && (a.box_rows > 256 || (a.box_rows % 8) != 0)) {  // split the strip into equal boxes of a multiple of 8 rows
        a.n_boxes++;
        if (a.strip_rows % a.n_boxes) { a.box_rows = 1000; continue; }
        a.box_rows = a.strip_rows / a.n_boxes;
        if (a.n_boxes > 64) { set_error("cannot split strip into TMA boxes"); return VS_ERR_INVALID; }
    }
    a.act = call.act;
    a.scale = call.scale; a.shift = call.shift;
    a.out_hi = out_hi; a.out_lo = (a.passes == 3) ? out_lo : nullptr; a.out32 = call.out32;
    const int strip_bytes = a.strip_rows * 128;
    const int fixed = 1024 + kWStages * kWTileBytes + 512;
    a.s_stages = (s->max_smem - fixed) / strip_bytes;
    if (a.s_stages > 6) a.s_stages = 6;
    if (a.s_stages < 2) { set_error("not enough shared memory for the conv strips"); return VS_ERR_UNSUPPORTED; }
    const int smem = fixed + a.s_stages * strip_bytes;

    CUtensorMap tm_in_hi, tm_in_lo, tm_w_hi, tm_w_lo;
    {
        uint64_t dims[3] = {64, (uint64_t)a.Q, (uint64_t)B};
        uint64_t str[2] = {128, (uint64_t)a.Q * 128};
        uint32_t box[3] = {64, (uint32_t)a.box_rows, 1};
        bool ok;
        if (a.tile2d) {   // [B][T][Fp][64]: box = 64 channels x 8 bins x (tr + kh - 1) frames; frames / bins outside the plane are zero-filled
            uint64_t d4[4] = {64, (uint64_t)Fp, (uint64_t)T, (uint64_t)B};
            uint64_t s4[3] = {128, (uint64_t)Fp * 128, (uint64_t)a.Q * 128};
            uint32_t b4[4] = {64, 8, (uint32_t)(a.tr + g.kh - 1), 1};
            ok = make_tmap_bf16(&tm_in_hi, (void*)in_hi, 4, d4, s4, b4, CU_TENSOR_MAP_SWIZZLE_128B);
            ok = ok && make_tmap_bf16(&tm_in_lo, (void*)(in_lo ? in_lo : in_hi), 4, d4, s4, b4, CU_TENSOR_MAP_SWIZZLE_128B);
        } else {
            ok = make_tmap_bf16(&tm_in_hi, (void*)in_hi, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
            ok = ok && make_tmap_bf16(&tm_in_lo, (void*)(in_lo ? in_lo : in_hi), 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
        }
        uint64_t wd[2] = {64, (uint64_t)a.n_dt * a.n_j * 128};
        uint64_t ws[1] = {128};
        uint32_t wb[2] = {64, 128};
        ok = ok && make_tmap_bf16(&tm_w_hi, (void*)call.w_hi, 2, wd, ws, wb, CU_TENSOR_MAP_SWIZZLE_128B);
        ok = ok && make_tmap_bf16(&tm_w_lo, (void*)call.w_lo, 2, wd, ws, wb, CU_TENSOR_MAP_SWIZZLE_128B);
        if (!ok) { set_error("cuTensorMapEncodeTiled failed"); return VS_ERR_CUDA; }
    }
    // cluster size sharing the weight fetches (see the header comment); grid = whole clusters, at most what is co-resident
    a.csz = s->cluster;
    if (a.csz > 1) {   // weight tiles arrive as 128 / csz-row slices, one per cluster member
        uint64_t wd[2] = {64, (uint64_t)a.n_dt * a.n_j * 128};
        uint64_t ws[1] = {128};
        uint32_t wb[2] = {64, (uint32_t)(128 / a.csz)};
        bool ok = make_tmap_bf16(&tm_w_hi, (void*)call.w_hi, 2, wd, ws, wb, CU_TENSOR_MAP_SWIZZLE_128B);
        ok = ok && make_tmap_bf16(&tm_w_lo, (void*)call.w_lo, 2, wd, ws, wb, CU_TENSOR_MAP_SWIZZLE_128B);
        if (!ok) { set_error("cuTensorMapEncodeTiled failed (weight slices)"); return VS_ERR_CUDA; }
    }
    cudaError_t ce;
    int grid = 0;
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)a.csz; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    const int ew = a.tile2d ? kEpiWarps2D : kEpiWarpsFlat;
    cfg.blockDim = dim3(conv_threads(ew)); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = st;
    cfg.attrs = attr; cfg.numAttrs = a.csz > 1 ? 1 : 0;
#define VS_CONV_TC_G(A, E, O, F8, W, G)                                                                          \
    do {                                                                                                         \
        ce = cudaFuncSetAttribute(k_conv_tc<A, E, O, F8, W, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);  \
        int max_ctas = e->num_sms;                                                                               \
        if (ce == cudaSuccess && a.csz > 1) {                                                                    \
            int ncl = 0;                                                                                         \
            cfg.gridDim = dim3((unsigned)(e->num_sms / a.csz * a.csz));                                          \
            ce = cudaOccupancyMaxActiveClusters(&ncl, k_conv_tc<A, E, O, F8, W, G>, &cfg);                        \
            if (ce == cudaSuccess && ncl < 1) { set_error("conv clusters do not fit the device"); return VS_ERR_UNSUPPORTED; } \
            max_ctas = ncl * a.csz < e->num_sms ? ncl * a.csz : e->num_sms / a.csz * a.csz;                       \
        }                                                                                                        \
        grid = a.total_tiles < max_ctas ? (a.total_tiles + a.csz - 1) / a.csz * a.csz : max_ctas;                \
        a.n_iter = (a.total_tiles + grid - 1) / grid;                                                            \
        cfg.gridDim = dim3((unsigned)grid);                                                                      \
        if (ce == cudaSuccess) ce = cudaLaunchKernelEx(&cfg, k_conv_tc<A, E, O, F8, W, G>, a, tm_in_hi, tm_in_lo, tm_w_hi, tm_w_lo); \
    } while (0)
    // the flat 5x5 layers in a two-plane mode take the issuer with the compile-time schedule (VOICESPLIT_CONV_FIXED_SCHEDULE=0: general loop)
    const int geo = (s->fixed_schedule && !a.tile2d && a.n_dt == 5 && a.n_j == 3 && a.s_stages == 4 && a.passes == 3 && a.N == 256 &&
                     (a.csz == 1 || a.csz == 2)) ? a.csz : 0;
#define VS_CONV_TC(A, E, O, F8)                                                                 \
    do {                                                                                        \
        if (a.tile2d) VS_CONV_TC_G(A, E, O, F8, kEpiWarps2D, 3);                                \
        else if (geo == 1) VS_CONV_TC_G(A, E, O, F8, kEpiWarpsFlat, 1);                         \
        else if (geo == 2) VS_CONV_TC_G(A, E, O, F8, kEpiWarpsFlat, 2);                         \
        else VS_CONV_TC_G(A, E, O, F8, kEpiWarpsFlat, 0);                                       \
    } while (0)
    if (call.out32) {
        if (call.act != 2) { set_error("fp32-output conv is pass-through only"); return VS_ERR_INVALID; }
        if (elt) VS_CONV_TC(2, 1, true, false); else VS_CONV_TC(2, 0, true, false);
    } else if (call.f8c) {
        if (a.act == VS_ACT_RELU) VS_CONV_TC(VS_ACT_RELU, 1, false, true); else VS_CONV_TC(VS_ACT_MISH, 1, false, true);
    } else if (a.act == VS_ACT_RELU) { if (elt) VS_CONV_TC(VS_ACT_RELU, 1, false, false); else VS_CONV_TC(VS_ACT_RELU, 0, false, false); }
    else { if (elt) VS_CONV_TC(VS_ACT_MISH, 1, false, false); else VS_CONV_TC(VS_ACT_MISH, 0, false, false); }
#undef VS_CONV_TC_G
#undef VS_CONV_TC
    if (ce == cudaSuccess) ce = cudaGetLastError();
    if (ce != cudaSuccess) { set_error(std::string("k_conv_tc launch: ") + cudaGetErrorString(ce)); return VS_ERR_CUDA; }
    e->launches++;
    if (e->profiling) prof_after(e, call.kid, st);
    return VS_OK;
}

static cudaError_t launch_front_tc(const vs_engine* e, const float* x, elt16* hi, elt16* lo, int elt, bool f8c, int B, int T, cudaStream_t st) {
    const int F = e->d.num_freq, Fp = padded_freq(F);
    const int nrows = B * T;
    const int grid = nrows < e->num_sms * 12 ? nrows : e->num_sms * 12;
    const size_t smem = (size_t)(Fp + 8) * sizeof(float);
#define VS_FRONT(A, E, F8) k_front_tc<A, E, F8><<<grid, 256, smem, st>>>(x, hi, lo, e->conv_w32[0], e->conv_scale[0], e->conv_shift[0], F, Fp, nrows)
    if (f8c) { if (e->d.activation == VS_ACT_RELU) VS_FRONT(VS_ACT_RELU, 1, true); else VS_FRONT(VS_ACT_MISH, 1, true); }
    else if (e->d.activation == VS_ACT_RELU) { if (elt) VS_FRONT(VS_ACT_RELU, 1, false); else VS_FRONT(VS_ACT_RELU, 0, false); }
    else { if (elt) VS_FRONT(VS_ACT_MISH, 1, false); else VS_FRONT(VS_ACT_MISH, 0, false); }
#undef VS_FRONT
    return cudaGetLastError();
}

template <int ACT, bool F8C>
static cudaError_t launch_point8_one(const vs_engine* e, unsigned grid, int smem, const elt16* hi, const elt16* lo, int elt, float* x32, elt16* xhi,
                                     elt16* xlo, int ldx, int F, int Fp, long long nplane, cudaStream_t st) {
    cudaError_t ce = cudaFuncSetAttribute(k_point8_tc<ACT, F8C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (ce != cudaSuccess) return ce;
    k_point8_tc<ACT, F8C><<<grid, 256, smem, st>>>(hi, lo, elt, e->conv_w32[7], e->conv_scale[7], e->conv_shift[7], x32, xhi, xlo, ldx, F, Fp, nplane);
    return cudaGetLastError();
}
// f8c: `lo` is the c8 correction plane of VS_PREC_FP16_F8C (the value is hi + 2^-8 * l8)
static cudaError_t launch_point8_tc(const vs_engine* e, const elt16* hi, const elt16* lo, int elt, bool f8c, float* x32,
                                    elt16* xhi, elt16* xlo, int ldx, int B, int T, cudaStream_t st) {
    const int F = e->d.num_freq, Fp = padded_freq(F);
    const long long nplane = (long long)B * T * Fp;
    const TcState* s = (const TcState*)e->tc;
    if (s->point8_mma && nplane < (1ll << 31) - 256) {
        Point8Args a{};
        a.mode = f8c ? 2 : (lo ? 1 : 0);
        a.stage_bytes = a.mode == 0 ? 16384 : (a.mode == 1 ? 32768 : 24576);
        a.stages = (s->max_smem - 1024 - 8192 - 512) / a.stage_bytes;
        if (a.stages > 8) a.stages = 8;
        a.n_tiles = (int)((nplane + 127) / 128);
        a.nplane = nplane; a.F = F; a.Fp = Fp; a.ldx = ldx;
        a.scale = s->scale_tc[7]; a.shift = e->conv_shift[7];
        a.x32 = x32; a.xhi = xhi; a.xlo = xlo;
        CUtensorMap tm_hi, tm_lo, tm_w16, tm_w8;
        uint64_t pd[2] = {64, (uint64_t)nplane}, ps[1] = {128};
        uint32_t pb[2] = {64, 128}, pb8[2] = {32, 128};
        uint64_t wd[2] = {64, 32}, wd8[2] = {32, 32}, ws8[1] = {64};
        uint32_t wb[2] = {64, 16}, wb8[2] = {32, 16};
        bool ok = make_tmap_bf16(&tm_hi, (void*)hi, 2, pd, ps, pb, CU_TENSOR_MAP_SWIZZLE_128B);
        // the l8 half is 64 of the 128 bytes of a c8 row: no L2 sector promotion, or the x8 half is fetched from DRAM as well
        ok = ok && (f8c ? make_tmap_bf16(&tm_lo, (void*)lo, 2, pd, ps, pb8, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE)
                        : make_tmap_bf16(&tm_lo, (void*)(lo ? lo : hi), 2, pd, ps, pb, CU_TENSOR_MAP_SWIZZLE_128B));
        ok = ok && make_tmap_bf16(&tm_w16, s->p8_w[elt ? 1 : 0], 2, wd, ps, wb, CU_TENSOR_MAP_SWIZZLE_128B);
        ok = ok && make_tmap_bf16(&tm_w8, s->p8_w8, 2, wd8, ws8, wb8, CU_TENSOR_MAP_SWIZZLE_64B);
        if (!ok) return cudaErrorInvalidValue;
        const int smem = 1024 + 8192 + a.stages * a.stage_bytes + 512;
        const unsigned grid = (unsigned)(a.n_tiles < e->num_sms ? a.n_tiles : e->num_sms);
        const bool relu = e->d.activation == VS_ACT_RELU;
#define VS_P8(A, E)                                                                                                          \
    do {                                                                                                                     \
        cudaError_t ce_ = cudaFuncSetAttribute(k_point8_mma<A, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);        \
        if (ce_ != cudaSuccess) return ce_;                                                                                  \
        k_point8_mma<A, E><<<grid, kP8Threads, smem, st>>>(a, tm_hi, tm_lo, tm_w16, tm_w8);                                   \
    } while (0)
        if (relu) { if (elt || f8c) VS_P8(VS_ACT_RELU, 1); else VS_P8(VS_ACT_RELU, 0); }
        else { if (elt || f8c) VS_P8(VS_ACT_MISH, 1); else VS_P8(VS_ACT_MISH, 0); }
#undef VS_P8
        return cudaGetLastError();
    }
    const unsigned grid = (unsigned)((nplane + 255) / 256);
    const int smem = 2 * 256 * 144;
    const bool relu = e->d.activation == VS_ACT_RELU;
    if (f8c) return relu ? launch_point8_one<VS_ACT_RELU, true>(e, grid, smem, hi, lo, elt, x32, xhi, xlo, ldx, F, Fp, nplane, st)
                         : launch_point8_one<VS_ACT_MISH, true>(e, grid, smem, hi, lo, elt, x32, xhi, xlo, ldx, F, Fp, nplane, st);
    return relu ? launch_point8_one<VS_ACT_RELU, false>(e, grid, smem, hi, lo, elt, x32, xhi, xlo, ldx, F, Fp, nplane, st)
                : launch_point8_one<VS_ACT_MISH, false>(e, grid, smem, hi, lo, elt, x32, xhi, xlo, ldx, F, Fp, nplane, st);
}

// ---- workspace of the tensor-core path ---------------------------------------------------------
struct TcWorkspace {
    elt16 *a_hi, *a_lo, *b_hi, *b_lo;  // ping-pong activation planes
    void* gemm;                         // operands of the tensor-core GEMMs
    size_t total;
};
static TcWorkspace tc_carve(const vs_engine* e, int B, int T, int precision, void* base) {
    const int F = e->d.num_freq, Fp = padded_freq(F);
    char* p = (char*)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += align_up(bytes, 1024); return r; };
    const size_t plane = (size_t)B * T * Fp * 64 * sizeof(elt16);
    TcWorkspace w{};
    w.a_hi = (elt16*)take(plane);
    w.b_hi = (elt16*)take(plane);
    if (tc_passes(precision) == 3) {
        w.a_lo = (elt16*)take(plane);
        w.b_lo = (elt16*)take(plane);
    }
    w.gemm = take(tc_gemm_workspace_bytes(e, B, T, precision));
    w.total = off;
    return w;
}
size_t tc_workspace_bytes(const vs_engine* e, int B, int T, int precision) { return tc_carve(e, B, T, precision, nullptr).total; }

// conv stack on tensor cores; result in the bf16 plane pair returned through (*res_hi, *res_lo)
static int conv_layers_tc(vs_engine* e, const float* x, const TcWorkspace& w, int B, int T, int precision, cudaStream_t st,
                          elt16** res_hi, elt16** res_lo) {
    VS_LAUNCH(e, KID_FRONT, st, launch_front_tc(e, x, w.a_hi, w.a_lo, tc_elt(precision), tc_f8c(precision), B, T, st));
    elt16 *sh = w.a_hi, *sl = w.a_lo, *dh = w.b_hi, *dl = w.b_lo;
    for (int l = 1; l <= 6; ++l) {
        int rc = launch_conv_tc(e, l, sh, sl, dh, dl, B, T, precision, st);
        if (rc != VS_OK) return rc;
        elt16* t;
        t = sh; sh = dh; dh = t;
        t = sl; sl = dl; dl = t;
    }
    *res_hi = sh; *res_lo = sl;
    return VS_OK;
}

int tc_conv_stack(vs_engine* e, const float* x, float* conv_out, int B, int T, int precision, void* ws, cudaStream_t st) {
    TcWorkspace w = tc_carve(e, B, T, precision, ws);
    elt16 *rh, *rl;
    int rc = conv_layers_tc(e, x, w, B, T, precision, st, &rh, &rl);
    if (rc != VS_OK) return rc;
    VS_LAUNCH(e, KID_POINT8, st, launch_point8_tc(e, rh, rl, tc_elt(precision), tc_f8c(precision), conv_out, nullptr, nullptr, 8 * e->d.num_freq, B, T, st));
    return VS_OK;
}

int tc_forward(vs_engine* e, const float* x, const float* emb, float* mask, float* masked, int B, int T, int precision,
               void* ws, const TcLstmBuffers& lb, cudaStream_t st) {
    TcWorkspace w = tc_carve(e, B, T, precision, ws);
    elt16 *rh, *rl;
    int rc = conv_layers_tc(e, x, w, B, T, precision, st, &rh, &rl);
    if (rc != VS_OK) return rc;
    return tc_lstm_head(e, rh, rl, tc_f8c(precision), nullptr, emb, x, mask, masked, B, T, tc_head_precision(precision), w.gemm, lb, st);
}

int tc_debug_layer(vs_engine* e, int layer, const float* x, const float* plane_in, float* plane_out, int B, int T,
                   int precision, cudaStream_t st) {
    const int F = e->d.num_freq, Fp = padded_freq(F);
    const long long n = (long long)B * T * Fp * 64;
    elt16* buf = nullptr;
    VS_CUDA_TRY(cudaMalloc(&buf, (size_t)n * 4 * sizeof(elt16)));
    elt16 *ih = buf, *il = buf + n, *oh = buf + 2 * n, *ol = buf + 3 * n;
    const bool x3 = tc_passes(precision) == 3;
    const int elt = tc_elt(precision);
    int rc = VS_OK;
    cudaError_t ce = cudaSuccess;
    const bool f8c = tc_f8c(precision);
    if (layer == 0) {
        ce = launch_front_tc(e, x, oh, x3 ? ol : nullptr, elt, f8c, B, T, st);
    } else {
        if (f8c) k_plane_split_f8c<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(plane_in, ih, reinterpret_cast<uint8_t*>(il), n);
        else k_plane_split<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(plane_in, ih, x3 ? il : nullptr, elt, n);
        rc = launch_conv_tc(e, layer, ih, x3 ? il : nullptr, oh, ol, B, T, precision, st);
    }
    if (rc == VS_OK && ce == cudaSuccess) {
        if (f8c) k_plane_join_f8c<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(oh, reinterpret_cast<const uint8_t*>(ol), plane_out, n);
        else k_plane_join<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(oh, x3 ? ol : nullptr, elt, plane_out, n);
        ce = cudaStreamSynchronize(st);
    }
    cudaFree(buf);
    if (ce != cudaSuccess) { set_error(std::string("tc_debug_layer: ") + cudaGetErrorString(ce)); return VS_ERR_CUDA; }
    return rc;
}

cudaError_t tc_launch_point8(const vs_engine* e, const elt16* hi, const elt16* lo, int elt, bool f8c, float* x32, elt16* xhi,
                             elt16* xlo, int ldx, int B, int T, cudaStream_t st) {
    return launch_point8_tc(e, hi, lo, elt, f8c, x32, xhi, xlo, ldx, B, T, st);
}

}  // namespace vs
