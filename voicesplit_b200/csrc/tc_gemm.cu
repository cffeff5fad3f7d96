// LSTM input projection and FC head on tensor cores: one warp-specialised tcgen05 GEMM
//     C[M][N] = A[M][K] * W[N][K]^T      (A = activations, rows = (utterance, frame); W = weights)
// with split 16-bit operands (hi/lo planes, 1 or 3 MMA passes into one fp32 TMEM accumulator) and
// three fused epilogues:
//   GATES : + per-utterance gate bias (W_ih[:,8F:] emb + b_ih + b_hh)  -> fp32 gates_x for the recurrence
//   FC1   : + bias, ReLU                                               -> 16-bit hi/lo operand of fc2
//   FC2   : + bias, sigmoid, * spectrogram                             -> mask (and masked) fp32
// Tile: 128 rows of A (the MMA M, one TMEM lane per row) x n_tile <= 256 rows of W (the MMA N).
// K is walked in 64-element blocks (one 128-byte swizzle atom per row); TMA zero-fills the K and
// N tails, so nothing is padded in memory.  In the 3-pass modes one pipeline stage holds all four
// operand tiles of a K block (A_hi, A_lo, W_hi, W_lo: each fetched once) and the issuer runs
// hi*hi + lo*hi + hi*lo on it.  Warps: 0 = TMA producer, 1 = MMA issuer, 2..9 = epilogue.
//
// Accumulation: tcgen05 adds into the fp32 TMEM accumulator with truncation (measured with
// tools/umma_probe.cu: ~0.3 ulp per accumulate step, always toward zero - profiles/r01_umma_probe.txt),
// which over the K = 8F = 4808 input projection (912 steps in 3-pass mode) is a -1.7e-5 relative bias.
// So the K loop is cut into chunks of kChunkKb K-blocks: each chunk accumulates in one of the two TMEM
// buffers and the epilogue warps add the chunk into fp32 registers (round-to-nearest) while the
// next chunk runs in the other buffer - the "promotion" FP8 GEMMs use, for the same reason.
#include "tc.cuh"
#include "sm100_ptx.cuh"
#include <stdlib.h>

namespace vs {
using namespace ptx;

constexpr int kGemmEpiWarps = 8;
constexpr int kChunkKb = 8;        // K blocks (of 64) accumulated in TMEM before promotion to registers
constexpr int kGemmThreads = 64 + 32 * kGemmEpiWarps;
template <int EPI, int ELT>
__global__ void __launch_bounds__(kGemmThreads, 1) k_gemm_tc(const GemmTcArgs a, const __grid_constant__ CUtensorMap tm_a_hi,
                                                             const __grid_constant__ CUtensorMap tm_a_lo,
                                                             const __grid_constant__ CUtensorMap tm_w_hi,
                                                             const __grid_constant__ CUtensorMap tm_w_lo) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int a_bytes = 128 * 128, w_bytes = a.n_tile * 128;
    const int w_bytes_al = (w_bytes + 1023) & ~1023;
    const int nplanes = a.passes == 3 ? 2 : 1;
    const int stage_bytes = nplanes * (a_bytes + w_bytes_al);   // [A_hi][A_lo][W_hi][W_lo]
    const int nst = a.stages;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)nst * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + nst;
    uint64_t* acc_full = empty + nst;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < nst; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], (uint32_t)a.csz); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kGemmEpiWarps); }
        fence_barrier_init();
        prefetch_tensormap(&tm_a_hi); prefetch_tensormap(&tm_w_hi);
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    if (a.csz > 1) cluster_sync_all();     // the peer's barriers exist before anything is multicast at them
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // Work items: (group of csz consecutive M blocks, N block); the CTAs of a cluster take the M blocks of one item, so every W tile
    // (the larger operand: n_tile x 64 against 128 x 64) is fetched from L2 once per cluster - each CTA loads 1/csz of its rows and
    // multicasts them.  The input projection is bound by L2 -> SM throughput (profiles/r02_ncu_gates_*), not by the tensor pipe.
    const int crank = a.csz > 1 ? (int)cluster_ctarank() : 0;
    const int n_items = ((a.n_tiles_m + a.csz - 1) / a.csz) * a.n_tiles_n;
    const int item0 = blockIdx.x / a.csz, item_step = gridDim.x / a.csz;
    const uint16_t cmask = (uint16_t)((1u << a.csz) - 1);

    if (warp == 0) {
        {   // producer: whole warp waits (warp-uniform), one elected lane issues the TMA loads
            int st = 0, ph = 0;
            const int w_rows = a.n_tile / a.csz;       // W rows this CTA fetches (for every CTA of the cluster)
            for (int item = item0; item < n_items; item += item_step) {
                const int mg = item / a.n_tiles_n, nb = item - mg * a.n_tiles_n;
                const int mb = mg * a.csz + crank;
                for (int kb = 0; kb < a.n_kb; ++kb) {
                    mbar_wait(&empty[st], ph ^ 1);
                    if (elect_one()) {
                    mbar_arrive_expect_tx(&full[st], (uint32_t)(nplanes * (a_bytes + w_bytes)));
                    uint8_t* dst = smem + (size_t)st * stage_bytes;
                    if (a.tn) {
                        // operands are [K rows][cols]: 64-column blocks of 64 K-rows each (8 KB, one swizzle atom wide)
                        for (int ib = 0; ib < 2; ++ib) {
                            tma_load_2d(dst + ib * 8192, &tm_a_hi, &full[st], mb * 128 + ib * 64, kb * 64);
                            if (nplanes == 2) tma_load_2d(dst + a_bytes + ib * 8192, &tm_a_lo, &full[st], mb * 128 + ib * 64, kb * 64);
                        }
                        for (int jb = 0; jb < a.n_tile / 64; ++jb) {
                            tma_load_2d(dst + nplanes * a_bytes + jb * 8192, &tm_w_hi, &full[st], nb * a.n_tile + jb * 64, kb * 64);
                            if (nplanes == 2)
                                tma_load_2d(dst + nplanes * a_bytes + w_bytes_al + jb * 8192, &tm_w_lo, &full[st], nb * a.n_tile + jb * 64, kb * 64);
                        }
                    } else if (a.csz > 1) {
                        tma_load_2d(dst, &tm_a_hi, &full[st], kb * 64, mb * 128);
                        if (nplanes == 2) tma_load_2d(dst + a_bytes, &tm_a_lo, &full[st], kb * 64, mb * 128);
                        uint8_t* wdst = dst + nplanes * a_bytes + (size_t)crank * w_rows * 128;
                        tma_load_2d_mc(wdst, &tm_w_hi, &full[st], kb * 64, nb * a.n_tile + crank * w_rows, cmask);
                        if (nplanes == 2) tma_load_2d_mc(wdst + w_bytes_al, &tm_w_lo, &full[st], kb * 64, nb * a.n_tile + crank * w_rows, cmask);
                    } else {
                        tma_load_2d(dst, &tm_a_hi, &full[st], kb * 64, mb * 128);
                        if (nplanes == 2) tma_load_2d(dst + a_bytes, &tm_a_lo, &full[st], kb * 64, mb * 128);
                        tma_load_2d(dst + nplanes * a_bytes, &tm_w_hi, &full[st], kb * 64, nb * a.n_tile);
                        if (nplanes == 2) tma_load_2d(dst + nplanes * a_bytes + w_bytes_al, &tm_w_lo, &full[st], kb * 64, nb * a.n_tile);
                    }
                    }
                    __syncwarp();
                    if (++st == nst) { st = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // MMA issuer: the whole warp runs the warp-uniform control flow and waits, one elected lane issues (an `if (lane == 0)`
        // region makes the compiler wrap every uniform-datapath instruction in an elect-and-loop sequence: ~100 cycles per MMA)
        {
            // TN: both operands MN-major; K step = 16 rows of 128 B, the 64-column blocks are 8 KB apart (LBO)
            const uint32_t idesc = make_idesc_bf16(128, a.n_tile, ELT) | (a.tn ? ((1u << 15) | (1u << 16)) : 0u);
            const uint32_t kstep16 = a.tn ? (2048u >> 4) : (32u >> 4), lbo = a.tn ? 8192u : 16u;   // K step in descriptor address units
            int st = 0, ph = 0, cc = 0;   // cc: chunk counter across tiles (TMEM buffer = cc & 1)
            for (int item = item0; item < n_items; item += item_step) {
                for (int kb0 = 0; kb0 < a.n_kb; kb0 += kChunkKb, ++cc) {
                    const int buf = cc & 1, aph = (cc >> 1) & 1;
                    mbar_wait(&acc_empty[buf], aph ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem + (uint32_t)(buf * 256);
                    uint32_t accumulate = 0;
                    const int kb1 = kb0 + kChunkKb < a.n_kb ? kb0 + kChunkKb : a.n_kb;
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(&full[st], ph);
                        tc_fence_after();
                        const uint32_t a_hi = smem_u32(smem + (size_t)st * stage_bytes), a_lo = a_hi + a_bytes;
                        const uint32_t w_hi = a_hi + nplanes * a_bytes, w_lo = w_hi + w_bytes_al;
                        if (elect_one()) {
                            const uint64_t d_ah = make_smem_desc(a_hi, lbo, 1024, 2), d_wh = make_smem_desc(w_hi, lbo, 1024, 2);
                            const uint64_t d_al = make_smem_desc(a_lo, lbo, 1024, 2), d_wl = make_smem_desc(w_lo, lbo, 1024, 2);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t o = (uint64_t)(k * kstep16);
                                umma_bf16(d_tmem, d_ah + o, d_wh + o, idesc, k == 0 ? accumulate : 1u);
                                if (nplanes == 2) {
                                    umma_bf16(d_tmem, d_al + o, d_wh + o, idesc, 1);
                                    umma_bf16(d_tmem, d_ah + o, d_wl + o, idesc, 1);
                                }
                            }
                            // the stage is free once BOTH CTAs of the cluster have consumed it: the peer multicasts into it too
                            if (a.csz > 1) umma_commit_mc(&empty[st], cmask); else umma_commit(&empty[st]);
                        }
                        __syncwarp();
                        accumulate = 1;
                        if (++st == nst) { st = 0; ph ^= 1; }
                    }
                    if (elect_one()) umma_commit(&acc_full[buf]);
                    __syncwarp();
                }
            }
        }
    } else {
        // epilogue: thread = one row of A (TMEM lane); warps of a quadrant alternate 32-column chunks
        const int quad = warp & 3, cgrp = (warp - 2) >> 2;
        constexpr int kColStep = 32 * (kGemmEpiWarps / 4);      // this warp takes columns cgrp*32 + i*kColStep .. +32
        constexpr int kMaxCols = 256 / kColStep;                // at most 4 column chunks per warp
        int cc = 0;
        for (int item = item0; item < n_items; item += item_step) {
            const int mg = item / a.n_tiles_n, nb = item - mg * a.n_tiles_n;
            const int mb = mg * a.csz + crank;
            const int m = mb * 128 + quad * 32 + lane;
            const int n0 = nb * a.n_tile;
            float acc[kMaxCols][32];
#pragma unroll
            for (int i = 0; i < kMaxCols; ++i)
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[i][j] = 0.f;
            for (int kb0 = 0; kb0 < a.n_kb; kb0 += kChunkKb, ++cc) {
                const int buf = cc & 1, aph = (cc >> 1) & 1;
                mbar_wait(&acc_full[buf], aph);
                tc_fence_after();
                const uint32_t t_base = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 256);
#pragma unroll
                for (int i = 0; i < kMaxCols; ++i) {
                    const int c0 = cgrp * 32 + i * kColStep;
                    if (c0 < a.n_tile) {
                        uint32_t r[32];
                        tmem_ld_32x32(t_base + c0, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[i][j] += __uint_as_float(r[j]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[buf]);
            }
            if (m < a.M) {
                const float* bg = (EPI == GEPI_GATES) ? a.bias_group + (size_t)(m / a.group_rows) * a.N : nullptr;
#pragma unroll
                for (int i = 0; i < kMaxCols; ++i) {
                    const int c0 = cgrp * 32 + i * kColStep;
                    if (c0 >= a.n_tile) continue;
                    const int nbase = n0 + c0;
                    if (EPI == GEPI_GATES) {
                        float* dst = a.out32 + (size_t)m * a.ld_out + nbase;
                        if (nbase + 32 <= a.N && c0 + 32 <= a.n_tile) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                float4 b4 = *reinterpret_cast<const float4*>(bg + nbase + j);
                                *reinterpret_cast<float4*>(dst + j) =
                                    make_float4(acc[i][j] + b4.x, acc[i][j + 1] + b4.y, acc[i][j + 2] + b4.z, acc[i][j + 3] + b4.w);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (nbase + j < a.N && c0 + j < a.n_tile) dst[j] = acc[i][j] + bg[nbase + j];
                        }
                    } else if (EPI == GEPI_PLAIN) {
                        float* dst = a.out32 + (size_t)m * a.ld_out + nbase;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (nbase + j < a.N && c0 + j < a.n_tile) dst[j] = acc[i][j];
                    } else if (EPI == GEPI_STFT) {
                        // columns are (re, im) pairs of DFT bin (nbase + j) / 2; row m = utterance * rows_per_utt + frame
                        const int ub = m / a.rows_per_utt, t = m - ub * a.rows_per_utt;
                        if (t < a.t_valid) {
                            const size_t row = ((size_t)ub * a.t_valid + t) * a.n_bins;
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                const int k = (nbase + j) >> 1;
                                if (k < a.n_bins && c0 + j < a.n_tile) {
                                    const float re = acc[i][j], im = acc[i][j + 1];
                                    const float mag = sqrtf(fmaf(re, re, im * im));
                                    // utils/audio_processor.py:473-474,537-544: 20 log10(max(1e-5, |D|)) - ref, clip(S / -min, -1, 0) + 1
                                    const float db = 20.f * log10f(fmaxf(1e-5f, mag)) - a.ref_db;
                                    a.out32[row + k] = fminf(fmaxf(db / -a.min_db, -1.f), 0.f) + 1.f;
                                    const float inv = mag > 0.f ? 1.f / mag : 0.f;
                                    reinterpret_cast<float2*>(a.phasor)[row + k] = mag > 0.f ? make_float2(re * inv, im * inv) : make_float2(1.f, 0.f);
                                }
                            }
                        }
                    } else if (EPI == GEPI_STFT_POWER) {
                        const int ub = m / a.rows_per_utt, t = m - ub * a.rows_per_utt;
                        if (t < a.t_valid) {
                            const size_t row = ((size_t)ub * a.t_valid + t) * a.ld16;
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                const int k = (nbase + j) >> 1;
                                if (k < a.n_bins && c0 + j < a.n_tile) {
                                    const float re = acc[i][j], im = acc[i][j + 1];
                                    elt16 vh, vl;
                                    split16<0>(fmaf(re, re, im * im), vh, vl);      // bf16: the power spans many decades
                                    a.out_hi[row + k] = vh;
                                    a.out_lo[row + k] = vl;
                                }
                            }
                        }
                    } else if (EPI == GEPI_LOGMEL) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int n = nbase + j;
                            if (n < a.N && c0 + j < a.n_tile) {
                                const float v = log10f(acc[i][j] + 1e-6f);         // utils/audio_processor.py:467
                                if (a.out32) a.out32[(size_t)m * a.ld_out + n] = v;
                                elt16 vh, vl;
                                split16<1>(v, vh, vl);
                                a.out_hi[(size_t)m * a.ld16 + n] = vh;
                                a.out_lo[(size_t)m * a.ld16 + n] = vl;
                            }
                        }
                    } else if (EPI == GEPI_ISTFT_BWD) {
                        // utils/audio_processor.py:500-509 backwards: (dRe, dIm) -> d magnitude -> d dB -> d normalised value;
                        // torch.clamp passes the gradient on the closed interval [0, 1]
                        const int ub = m / a.rows_per_utt, t = m - ub * a.rows_per_utt;
                        if (t < a.t_valid) {
                            const size_t row = ((size_t)ub * a.t_valid + t) * a.n_bins;
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                const int k = (nbase + j) >> 1;
                                if (k < a.n_bins && c0 + j < a.n_tile) {
                                    const float sv = a.g_spec[row + k];
                                    float g = 0.f;
                                    if (sv >= 0.f && sv <= 1.f) {
                                        float sn, cs;
                                        sincosf(a.g_phase[row + k], &sn, &cs);
                                        const float wr = a.q1 ? expf(cs) : cs, wi = a.q1 ? expf(sn) : sn;
                                        const float amp = exp10f(((sv - 1.f) * -a.min_db + a.ref_db) * 0.05f);
                                        g = (acc[i][j] * wr + acc[i][j + 1] * wi) * amp * (0.05f * 2.302585093f * -a.min_db);
                                    }
                                    a.out32[row + k] = g;
                                }
                            }
                        }
                    } else if (EPI == GEPI_FC1) {
                        if (c0 + 32 <= a.n_tile && nbase + 32 <= a.ld16) {
                            // a whole 32-column chunk of this row: 64 contiguous bytes per plane -> four 16-byte stores (columns
                            // N..ld16 are row padding the consumer's TMA never reads past K = N: zeros there are harmless)
                            __align__(16) elt16 vh[32], vl[32];
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const int n = nbase + j;
                                const float v = n < a.N ? fmaxf(acc[i][j] + a.bias[n], 0.f) : 0.f;
                                split16<ELT>(v, vh[j], vl[j]);
                            }
                            uint4* dh = reinterpret_cast<uint4*>(a.out_hi + (size_t)m * a.ld16 + nbase);
#pragma unroll
                            for (int q = 0; q < 4; ++q) dh[q] = reinterpret_cast<const uint4*>(vh)[q];
                            if (a.out_lo) {
                                uint4* dl = reinterpret_cast<uint4*>(a.out_lo + (size_t)m * a.ld16 + nbase);
#pragma unroll
                                for (int q = 0; q < 4; ++q) dl[q] = reinterpret_cast<const uint4*>(vl)[q];
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const int n = nbase + j;
                                if (n < a.N && c0 + j < a.n_tile) {
                                    float v = fmaxf(acc[i][j] + a.bias[n], 0.f);
                                    elt16 vh, vl;
                                    split16<ELT>(v, vh, vl);
                                    a.out_hi[(size_t)m * a.ld16 + n] = vh;
                                    if (a.out_lo) a.out_lo[(size_t)m * a.ld16 + n] = vl;
                                }
                            }
                        }
                    }
                }
            }
            if (EPI == GEPI_FC2) {
                // mask = sigmoid(acc + b), masked = x * mask.  A thread owns one ROW of the tile, but rows of the [M][F] outputs
                // are F floats apart (F odd: no vector stores): transpose each 32 x 32 chunk through shared memory so that a
                // warp instruction covers 32 consecutive columns of ONE row (128 contiguous bytes) for the x load and both stores
                float* tr = reinterpret_cast<float*>(tmem_slot + 4) + (size_t)(warp - 2) * (32 * 33);
                const int m0 = mb * 128 + quad * 32;
#pragma unroll
                for (int i = 0; i < kMaxCols; ++i) {
                    const int c0 = cgrp * 32 + i * kColStep;
                    if (c0 >= a.n_tile) continue;
                    const int nbase = n0 + c0;
                    const float bias_l = nbase + lane < a.N ? __ldg(a.bias + nbase + lane) : 0.f;   // one load per chunk, broadcast below
#pragma unroll
                    for (int j = 0; j < 32; ++j) tr[lane * 33 + j] = sigmoid_f(acc[i][j] + __shfl_sync(0xffffffffu, bias_l, j));
                    __syncwarp();
                    const int n = nbase + lane;
                    if (n < a.N && c0 + lane < a.n_tile) {
                        if (m0 + 32 <= a.M) {
                            // the spectrogram values come from HBM: 16 loads in flight per lane before the first use (a rolled loop
                            // of load -> multiply -> store waited one DRAM latency per row: 47 % of the kernel's samples on that FMUL)
#pragma unroll
                            for (int r0 = 0; r0 < 32; r0 += 16) {
                                float xv[16];
#pragma unroll
                                for (int k = 0; k < 16; ++k)
                                    xv[k] = a.masked ? __ldg(a.xmul + (size_t)(m0 + r0 + k) * a.ld_out + n) : 0.f;
#pragma unroll
                                for (int k = 0; k < 16; ++k) {
                                    const float v = tr[(r0 + k) * 33 + lane];
                                    const size_t o = (size_t)(m0 + r0 + k) * a.ld_out + n;
                                    a.out32[o] = v;
                                    if (a.masked) a.masked[o] = xv[k] * v;
                                }
                            }
                        } else {
                            for (int r = 0; r < 32 && m0 + r < a.M; ++r) {
                                const float v = tr[r * 33 + lane];
                                const size_t o = (size_t)(m0 + r) * a.ld_out + n;
                                a.out32[o] = v;
                                if (a.masked) a.masked[o] = a.xmul[o] * v;
                            }
                        }
                    }
                    __syncwarp();
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (a.csz > 1) cluster_sync_all();     // nobody leaves while the peer may still multicast into its shared memory / barriers
    if (warp == 1) tmem_dealloc(tmem, 512);
}

// fp32 [rows][cols] (row stride ld) -> 16-bit hi/lo planes [rows][cols], optionally scaled by a power of two
__global__ void k_split_matrix(const float* __restrict__ src, int ld, int rows, int cols, int ldo, const unsigned int* maxbits,
                               elt16* __restrict__ bhi, elt16* __restrict__ blo, elt16* __restrict__ hhi, elt16* __restrict__ hlo) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)rows * ldo) return;
    int c = (int)(i % ldo);
    long long r = i / ldo;
    if (c >= cols) { bhi[i] = blo[i] = hhi[i] = hlo[i] = 0; return; }
    float s = 1.f;
    if (maxbits) {
        float m = __uint_as_float(*maxbits);
        if (m > 0.f && isfinite(m)) { int ex; frexpf(m, &ex); s = exp2f((float)(9 - ex)); }
    }
    float v = s * src[r * ld + c];
    split16<0>(v, bhi[i], blo[i]);
    split16<1>(v, hhi[i], hlo[i]);
}
__global__ void k_absmax2(const float* __restrict__ w, long long n, unsigned int* out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicMax(out, __float_as_uint(fabsf(w[i])));
}
// src [rows][cols] fp32 -> transposed [cols][rows] bf16 hi/lo
__global__ void k_transpose_split_bf16(const float* __restrict__ src, int rows, int cols, elt16* __restrict__ hi, elt16* __restrict__ lo) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)rows * cols) return;
    int r = (int)(i % rows);
    long long c = i / rows;
    elt16 h, l;
    split16<0>(src[(size_t)r * cols + c], h, l);
    hi[i] = h; lo[i] = l;
}
// fp32 [rows][cols] -> hi/lo planes of one element type (activations: conv_out in the debug hook)
__global__ void k_split_rows(const float* __restrict__ src, long long n, int elt, elt16* __restrict__ hi, elt16* __restrict__ lo) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    elt16 h, l;
    split16_rt(src[i], elt, h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
}

// ---- host ---------------------------------------------------------------------------------------
struct GemmState {
    // [elt] hi / lo planes of the three weight matrices (unscaled: fp16 lo of small LSTM/FC weights stays
    // accurate enough in absolute terms, and the GATES epilogue needs unscaled sums)
    elt16 *wih_hi[2] = {}, *wih_lo[2] = {};  // [8H][8F]
    elt16 *fc1_hi[2] = {}, *fc1_lo[2] = {};  // [N1][2H]
    elt16 *fc2_hi[2] = {}, *fc2_lo[2] = {};  // [F][N1]
    elt16 *wihT_hi = nullptr, *wihT_lo = nullptr;   // training: W_ih[:, :8F]^T as [8F][8H] bf16 (operand of dX = da W_ih)
    int max_smem = 0;
    int cluster = 2;      // VOICESPLIT_GEMM_CLUSTER = 1 disables the W-tile multicast pairs
};
static GemmState* g_state(vs_engine* e);

struct TcStateHdr { void* gemm; };  // TcState (tc_conv.cu) starts with this member

int tc_gemm_pack(vs_engine* e, cudaStream_t st) {
    GemmState* g = g_state(e);
    const int F = e->d.num_freq, H = e->d.lstm_dim, N1 = e->d.fc1_dim;
    struct Item { const float* src; int rows, cols; elt16** hi; elt16** lo; };
    Item items[3] = {{e->wih_x, 8 * H, 8 * F, g->wih_hi, g->wih_lo}, {e->fc1_w, N1, 2 * H, g->fc1_hi, g->fc1_lo},
                     {e->fc2_w, F, N1, g->fc2_hi, g->fc2_lo}};
    for (Item& it : items) {
        const int ldo = (it.cols + 7) / 8 * 8;  // TMA needs 16-byte row strides
        const size_t n = (size_t)it.rows * ldo;
        for (int t = 0; t < 2; ++t) {
            if (!it.hi[t]) {
                VS_CUDA_TRY(cudaMalloc(&it.hi[t], n * sizeof(elt16)));
                VS_CUDA_TRY(cudaMalloc(&it.lo[t], n * sizeof(elt16)));
            }
        }
        k_split_matrix<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(it.src, it.cols, it.rows, it.cols, ldo, nullptr, it.hi[0], it.lo[0],
                                                                    it.hi[1], it.lo[1]);
    }
    {
        const size_t n = (size_t)8 * F * 8 * H;
        if (!g->wihT_hi) {
            VS_CUDA_TRY(cudaMalloc(&g->wihT_hi, n * sizeof(elt16)));
            VS_CUDA_TRY(cudaMalloc(&g->wihT_lo, n * sizeof(elt16)));
        }
        k_transpose_split_bf16<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(e->wih_x, 8 * H, 8 * F, g->wihT_hi, g->wihT_lo);
    }
    VS_CUDA_TRY(cudaGetLastError());
    return VS_OK;
}

void tc_gemm_destroy(vs_engine* e) {
    GemmState* g = g_state(e);
    if (!g) return;
    for (int t = 0; t < 2; ++t) {
        cudaFree(g->wih_hi[t]); cudaFree(g->wih_lo[t]); cudaFree(g->fc1_hi[t]); cudaFree(g->fc1_lo[t]);
        cudaFree(g->fc2_hi[t]); cudaFree(g->fc2_lo[t]);
    }
    cudaFree(g->wihT_hi); cudaFree(g->wihT_lo);
    delete g;
}

struct GemmWorkspace {
    elt16 *x_hi, *x_lo;    // [M][8F]   LSTM input (cnn8 output)
    elt16 *h_hi, *h_lo;    // [M][2H]   relu(lstm_out)
    elt16 *y_hi, *y_lo;    // [M][N1]   relu(fc1)
    size_t total;
};
static GemmWorkspace gemm_carve(const vs_engine* e, int B, int T, int precision, void* base) {
    const int F = e->d.num_freq, H = e->d.lstm_dim, N1 = e->d.fc1_dim;
    char* p = (char*)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += align_up(bytes, 1024); return r; };
    const size_t M = (size_t)B * T;
    const bool x3 = tc_passes(precision) == 3;
    GemmWorkspace w{};
    w.x_hi = (elt16*)take(M * 8 * F * sizeof(elt16));
    w.x_lo = x3 ? (elt16*)take(M * 8 * F * sizeof(elt16)) : nullptr;
    w.h_hi = (elt16*)take(M * 2 * H * sizeof(elt16));
    w.h_lo = x3 ? (elt16*)take(M * 2 * H * sizeof(elt16)) : nullptr;
    const size_t N1p = (size_t)(N1 + 7) / 8 * 8;
    w.y_hi = (elt16*)take(M * N1p * sizeof(elt16));
    w.y_lo = x3 ? (elt16*)take(M * N1p * sizeof(elt16)) : nullptr;
    w.total = off;
    return w;
}
size_t tc_gemm_workspace_bytes(const vs_engine* e, int B, int T, int precision) { return gemm_carve(e, B, T, precision, nullptr).total; }

int launch_gemm_tc(vs_engine* e, int epi, int kid, const elt16* a_hi, const elt16* a_lo, const elt16* w_hi, const elt16* w_lo,
                   GemmTcArgs a, int precision, cudaStream_t st) {
    GemmState* g = g_state(e);
    a.passes = tc_passes(precision);
    const int elt = tc_elt(precision);
    a.n_tiles_n = (a.N + 255) / 256;
    const int tile_q = a.tn ? 64 : 16;
    a.n_tile = (((a.N + a.n_tiles_n - 1) / a.n_tiles_n) + tile_q - 1) / tile_q * tile_q;
    a.n_tiles_m = (a.M + 127) / 128;
    a.total_tiles = a.n_tiles_m * a.n_tiles_n;
    a.n_kb = (a.K + 63) / 64;
    // pairs of M blocks share their W tiles through TMA multicast when there is enough work for every cluster (the large GEMMs)
    // (cluster == 3, VOICESPLIT_GEMM_CLUSTER=3: pairs whenever there are two M blocks - lets small test problems take this path)
    a.csz = (g->cluster > 1 && !a.tn && a.n_tile % 16 == 0 && (g->cluster == 3 ? a.n_tiles_m >= 2 : a.total_tiles >= 2 * e->num_sms)) ? 2 : 1;
    if (a.lda % 8 || a.ldw % 8) { set_error("tensor-core GEMM needs 16-byte aligned operand rows (lstm_dim % 4 == 0)"); return VS_ERR_INVALID; }
    CUtensorMap tm_a_hi, tm_a_lo, tm_w_hi, tm_w_lo;
    {
        uint64_t ad[2] = {(uint64_t)a.K, (uint64_t)a.M}, as[1] = {(uint64_t)a.lda * sizeof(elt16)};
        uint32_t ab[2] = {64, 128};
        uint64_t wd[2] = {(uint64_t)a.K, (uint64_t)a.N}, ws[1] = {(uint64_t)a.ldw * sizeof(elt16)};
        uint32_t wb[2] = {64, (uint32_t)(a.n_tile / a.csz)};
        if (a.tn) {   // [K rows][cols] operands: inner dimension = output rows / cols, boxes of 64 x 64
            ad[0] = (uint64_t)a.M; ad[1] = (uint64_t)a.K; ab[1] = 64;
            wd[0] = (uint64_t)a.N; wd[1] = (uint64_t)a.K; wb[1] = 64;
        }
        bool ok = make_tmap_bf16(&tm_a_hi, (void*)a_hi, 2, ad, as, ab, CU_TENSOR_MAP_SWIZZLE_128B);
        ok = ok && make_tmap_bf16(&tm_a_lo, (void*)(a_lo ? a_lo : a_hi), 2, ad, as, ab, CU_TENSOR_MAP_SWIZZLE_128B);
        ok = ok && make_tmap_bf16(&tm_w_hi, (void*)w_hi, 2, wd, ws, wb, CU_TENSOR_MAP_SWIZZLE_128B);
        ok = ok && make_tmap_bf16(&tm_w_lo, (void*)w_lo, 2, wd, ws, wb, CU_TENSOR_MAP_SWIZZLE_128B);
        if (!ok) { set_error("cuTensorMapEncodeTiled failed (gemm)"); return VS_ERR_CUDA; }
    }
    const int w_bytes_al = (a.n_tile * 128 + 1023) & ~1023;
    const int stage_bytes = (a.passes == 3 ? 2 : 1) * (128 * 128 + w_bytes_al);
    // FC2: 8 epilogue warps x (32 x 33) floats of transposition scratch behind the barriers
    const int epi_scratch = epi == GEPI_FC2 ? kGemmEpiWarps * 32 * 33 * 4 : 0;
    a.stages = (g->max_smem - 1024 - 512 - epi_scratch) / stage_bytes;
    if (a.stages > 6) a.stages = 6;
    if (a.stages < 2) { set_error("gemm tile does not fit shared memory"); return VS_ERR_UNSUPPORTED; }
    const int smem = 1024 + a.stages * stage_bytes + 512 + epi_scratch;
    const int n_items = (a.n_tiles_m + a.csz - 1) / a.csz * a.n_tiles_n;
    const int max_clusters = e->num_sms / a.csz;
    const int grid = (n_items < max_clusters ? n_items : max_clusters) * a.csz;
    cudaError_t ce = cudaSuccess;
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)a.csz; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kGemmThreads); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = st;
    cfg.attrs = attr; cfg.numAttrs = a.csz > 1 ? 1 : 0;
#define VS_GEMM_TC(E, L)                                                                                   \
    do {                                                                                                   \
        ce = cudaFuncSetAttribute(k_gemm_tc<E, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);      \
        if (ce == cudaSuccess) ce = cudaLaunchKernelEx(&cfg, k_gemm_tc<E, L>, a, tm_a_hi, tm_a_lo, tm_w_hi, tm_w_lo); \
    } while (0)
    if (epi == GEPI_PLAIN) { if (elt) VS_GEMM_TC(GEPI_PLAIN, 1); else VS_GEMM_TC(GEPI_PLAIN, 0); }
    else if (epi == GEPI_STFT) { if (elt) VS_GEMM_TC(GEPI_STFT, 1); else VS_GEMM_TC(GEPI_STFT, 0); }
    else if (epi == GEPI_ISTFT_BWD) { if (elt) VS_GEMM_TC(GEPI_ISTFT_BWD, 1); else VS_GEMM_TC(GEPI_ISTFT_BWD, 0); }
    else if (epi == GEPI_STFT_POWER) { if (elt) VS_GEMM_TC(GEPI_STFT_POWER, 1); else VS_GEMM_TC(GEPI_STFT_POWER, 0); }
    else if (epi == GEPI_LOGMEL) { if (elt) VS_GEMM_TC(GEPI_LOGMEL, 1); else VS_GEMM_TC(GEPI_LOGMEL, 0); }
    else if (epi == GEPI_GATES) { if (elt) VS_GEMM_TC(GEPI_GATES, 1); else VS_GEMM_TC(GEPI_GATES, 0); }
    else if (epi == GEPI_FC1) { if (elt) VS_GEMM_TC(GEPI_FC1, 1); else VS_GEMM_TC(GEPI_FC1, 0); }
    else { if (elt) VS_GEMM_TC(GEPI_FC2, 1); else VS_GEMM_TC(GEPI_FC2, 0); }
#undef VS_GEMM_TC
    if (ce == cudaSuccess) ce = cudaGetLastError();
    if (ce != cudaSuccess) { set_error(std::string("k_gemm_tc launch: ") + cudaGetErrorString(ce)); return VS_ERR_CUDA; }
    e->launches++;
    if (e->profiling) prof_after(e, kid, st);
    return VS_OK;
}

int tc_lstm_head(vs_engine* e, const elt16* plane_hi, const elt16* plane_lo, bool plane_f8c, const float* conv_out32, const float* emb,
                 const float* x, float* mask, float* masked, int B, int T, int precision, void* gemm_ws,
                 const TcLstmBuffers& lb, cudaStream_t st) {
    GemmState* g = g_state(e);
    const int F = e->d.num_freq, H = e->d.lstm_dim, E = e->d.emb_dim, N1 = e->d.fc1_dim;
    const int M = B * T, elt = tc_elt(precision);
    const bool x3 = tc_passes(precision) == 3;
    GemmWorkspace w = gemm_carve(e, B, T, precision, gemm_ws);
    if (conv_out32) {
        const long long n = (long long)M * 8 * F;
        k_split_rows<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(conv_out32, n, elt, w.x_hi, w.x_lo);
        VS_LAUNCH(e, KID_CONVERT, st, cudaGetLastError());
    } else {
        VS_LAUNCH(e, KID_POINT8, st, tc_launch_point8(e, plane_hi, plane_lo, elt, plane_f8c, nullptr, w.x_hi, w.x_lo, 8 * F, B, T, st));
    }
    // d-vector folded into a per-utterance gate bias (tiny: B x 8H x E, fp32 FFMA)
    VS_LAUNCH(e, KID_EMB_BIAS, st, launch_gemm_fp32(emb, E, e->wih_e, E, e->b_lstm, nullptr, 1, lb.bias_u, 8 * H, B, 8 * H, E,
                                                    false, EPI_NONE, nullptr, nullptr, st));
    {   // gates_x = X * W_ih[:, :8F]^T + bias_u[utterance]
        GemmTcArgs a{};
        a.M = M; a.N = 8 * H; a.K = 8 * F; a.lda = 8 * F; a.ldw = 8 * F; a.group_rows = T; a.bias_group = lb.bias_u; a.out32 = lb.gates; a.ld_out = 8 * H;
        int rc = launch_gemm_tc(e, GEPI_GATES, KID_INPROJ, w.x_hi, w.x_lo, g->wih_hi[elt], g->wih_lo[elt], a, precision, st);
        if (rc != VS_OK) return rc;
    }
    {
        int rc = tc_lstm_recurrence(e, tc_lstm_slot(e), lb.gates, lb.hout, lb.hx, w.h_hi, x3 ? w.h_lo : nullptr, B, T, precision, st);
        if (rc != VS_OK) return rc;
    }
    {   // y1 = relu(relu(h) * fc1^T + b1)
        GemmTcArgs a{};
        a.M = M; a.N = N1; a.K = 2 * H; a.lda = 2 * H; a.ldw = (2 * H + 7) / 8 * 8; a.bias = e->fc1_b;
        a.out_hi = w.y_hi; a.out_lo = x3 ? w.y_lo : nullptr; a.ld16 = (N1 + 7) / 8 * 8;
        int rc = launch_gemm_tc(e, GEPI_FC1, KID_FC1, w.h_hi, w.h_lo, g->fc1_hi[elt], g->fc1_lo[elt], a, precision, st);
        if (rc != VS_OK) return rc;
    }
    {   // mask = sigmoid(y1 * fc2^T + b2), masked = x * mask
        GemmTcArgs a{};
        a.M = M; a.N = F; a.K = N1; a.lda = (N1 + 7) / 8 * 8; a.ldw = (N1 + 7) / 8 * 8; a.bias = e->fc2_b; a.out32 = mask; a.ld_out = F; a.xmul = x; a.masked = masked;
        int rc = launch_gemm_tc(e, GEPI_FC2, KID_FC2, w.y_hi, w.y_lo, g->fc2_hi[elt], g->fc2_lo[elt], a, precision, st);
        if (rc != VS_OK) return rc;
    }
    return VS_OK;
}

// ---- training GEMMs --------------------------------------------------------------------------------
struct TrainGemmWs {
    elt16 *x_hi, *x_lo;     // [M][8F]  LSTM input, fp16 (forward) then bf16 (backward)
    elt16 *da_hi, *da_lo;   // [M][8H]  gate pre-activation gradients, bf16
    size_t total;
};
static TrainGemmWs train_gemm_carve(const vs_engine* e, int B, int T, void* base) {
    const int F = e->d.num_freq, H = e->d.lstm_dim;
    char* p = (char*)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += align_up(bytes, 1024); return r; };
    const size_t M = (size_t)B * T;
    TrainGemmWs w{};
    w.x_hi = (elt16*)take(M * 8 * F * 2); w.x_lo = (elt16*)take(M * 8 * F * 2);
    w.da_hi = (elt16*)take(M * 8 * H * 2); w.da_lo = (elt16*)take(M * 8 * H * 2);
    w.total = off;
    return w;
}
size_t tc_train_gemm_workspace_bytes(const vs_engine* e, int B, int T) { return train_gemm_carve(e, B, T, nullptr).total; }

// gates_x = X W_ih[:, :8F]^T + bias_u (fp16x3)
int tc_train_inproj(vs_engine* e, const float* xcat, const float* bias_u, float* gates, void* ws, int B, int T, cudaStream_t st) {
    GemmState* g = g_state(e);
    const int F = e->d.num_freq, H = e->d.lstm_dim;
    TrainGemmWs w = train_gemm_carve(e, B, T, ws);
    const long long n = (long long)B * T * 8 * F;
    k_split_rows<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(xcat, n, 1, w.x_hi, w.x_lo);
    VS_LAUNCH(e, KID_CONVERT, st, cudaGetLastError());
    GemmTcArgs a{};
    a.M = B * T; a.N = 8 * H; a.K = 8 * F; a.lda = 8 * F; a.ldw = 8 * F; a.group_rows = T; a.bias_group = bias_u; a.out32 = gates; a.ld_out = 8 * H;
    return launch_gemm_tc(e, GEPI_GATES, KID_INPROJ, w.x_hi, w.x_lo, g->wih_hi[1], g->wih_lo[1], a, VS_PREC_FP16X3, st);
}

// dX = da W_ih[:, :8F]  and  dW_ih[d][:, :8F] = da_d^T X  (bf16x3; the second contracts over tokens -> TN mode)
int tc_train_lstm_input_grads(vs_engine* e, const float* da, const float* xcat, float* dxcat, float* dw_ih0, float* dw_ih1, int ld_dw,
                              void* ws, int B, int T, cudaStream_t st) {
    GemmState* g = g_state(e);
    const int F = e->d.num_freq, H = e->d.lstm_dim, M = B * T;
    TrainGemmWs w = train_gemm_carve(e, B, T, ws);
    long long n = (long long)M * 8 * H;
    k_split_rows<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(da, n, 0, w.da_hi, w.da_lo);
    VS_LAUNCH(e, KID_CONVERT, st, cudaGetLastError());
    n = (long long)M * 8 * F;
    k_split_rows<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(xcat, n, 0, w.x_hi, w.x_lo);
    VS_LAUNCH(e, KID_CONVERT, st, cudaGetLastError());
    {
        GemmTcArgs a{};
        a.M = M; a.N = 8 * F; a.K = 8 * H; a.lda = 8 * H; a.ldw = 8 * H; a.out32 = dxcat; a.ld_out = 8 * F;
        int rc = launch_gemm_tc(e, GEPI_PLAIN, KID_TR_GEMM, w.da_hi, w.da_lo, g->wihT_hi, g->wihT_lo, a, VS_PREC_BF16X3, st);
        if (rc != VS_OK) return rc;
    }
    float* outs[2] = {dw_ih0, dw_ih1};
    for (int d = 0; d < 2; ++d) {
        GemmTcArgs a{};
        a.tn = 1;
        a.M = 4 * H; a.N = 8 * F; a.K = M; a.lda = 8 * H; a.ldw = 8 * F; a.out32 = outs[d]; a.ld_out = ld_dw;
        int rc = launch_gemm_tc(e, GEPI_PLAIN, KID_TR_GEMM, w.da_hi + (size_t)d * 4 * H, w.da_lo + (size_t)d * 4 * H, w.x_hi, w.x_lo, a, VS_PREC_BF16X3, st);
        if (rc != VS_OK) return rc;
    }
    return VS_OK;
}

int tc_debug_lstm_head(vs_engine* e, const float* conv_out, const float* emb, const float* x, float* mask, int B, int T,
                       int precision, const TcLstmBuffers& lb, cudaStream_t st) {
    void* ws = nullptr;
    VS_CUDA_TRY(cudaMalloc(&ws, tc_gemm_workspace_bytes(e, B, T, precision)));
    precision = tc_head_precision(precision);
    int rc = tc_lstm_head(e, nullptr, nullptr, false, conv_out, emb, x, mask, nullptr, B, T, precision, ws, lb, st);
    cudaStreamSynchronize(st);
    cudaFree(ws);
    return rc;
}

// GemmState lives behind vs_engine::tc (TcState in tc_conv.cu keeps the pointer as its first member)
static GemmState* g_state(vs_engine* e) {
    TcStateHdr* hdr = (TcStateHdr*)e->tc;
    if (!hdr->gemm) {
        GemmState* g = new GemmState();
        cudaDeviceGetAttribute(&g->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, e->device);
        if (const char* c = getenv("VOICESPLIT_GEMM_CLUSTER")) g->cluster = atoi(c) == 1 ? 1 : (atoi(c) == 3 ? 3 : 2);
        hdr->gemm = g;
    }
    return (GemmState*)hdr->gemm;
}

}  // namespace vs
