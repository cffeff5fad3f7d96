// BiLSTM recurrence on tensor cores (reference nn.LSTM, models/voicesplit/model.py:57-61,82).
//
// Persistent kernel, one CTA per (direction, slice of 16 hidden units, set of 1 or 2 groups of 128 utterances).
// The CTA keeps its 64 rows of W_hh (4 gates x 16 units, K = H) resident in shared memory as 16-bit
// hi/lo planes for the whole sequence.  Every step it
//   1. TMA-loads h_{t-1} of its 128 utterances ([128][H] hi/lo, written by the CTAs of the other
//      slices) in 64-wide K blocks,
//   2. issues  D[128 utt][64 gate cols] = h * W_slice^T  on tcgen05 (1 or 3 passes, fp32 in TMEM),
//   3. adds the precomputed input projection gates_x, applies sigmoid/tanh, updates the cell state
//      (registers: one thread = one utterance x 16 units, all four gates, for the whole sequence),
//   4. writes h_t (16-bit hi/lo exchange buffer for the next step, fp32 lstm_out, and relu(h) hi/lo
//      for the FC head) and arrives on the per-(direction, group) barrier.
// The CTAs of one (direction, group) meet at that barrier once per step; nothing else is shared.
// Optionally two groups per CTA (VOICESPLIT_LSTM_GROUPS_PER_CTA=2): independent recurrences over the same resident W slice,
// taken in turn by the producer and the MMA issuer, each with its own accumulator, cell warps and barrier.  Meant to hide one
// group's exchange (stores -> release -> the other slices' arrivals -> acquire -> TMA) under the other group's MMAs; measured
// slower than one group per CTA on twice the SMs (see tc_lstm_recurrence), kept as an experiment knob and tested.
#include "tc.cuh"
#include "sm100_ptx.cuh"
#include <stdlib.h>

namespace vs {
using namespace ptx;

constexpr int kLU = 16;         // hidden units per CTA
constexpr int kLB = 128;        // utterances per CTA (MMA M)
constexpr int kLStages = 3;     // h K-blocks in flight
constexpr int kLGroups = 2;     // utterance groups per CTA at most

struct LstmTcArgs {
    int B, T, H, nslices, nsets, gpc, group0, nkb, nk16, passes;   // nsets CTA sets of gpc (1 or 2) groups each, first group = group0
    int Bp;                       // utterance rows of the exchange buffer (multiple of 128)
    int Hp;                       // row stride of the exchange buffer (elements, multiple of 8)
    const float* gates_x;         // [B*T][8H]
    float* hout;                  // [B][T][2H]
    elt16* hx_hi;                 // [2 dir][2 parity][Bp][Hp]
    elt16* hx_lo;
    elt16* hr_hi;                 // [B*T][2H] relu(h), optional
    elt16* hr_lo;
    unsigned int* barrier;        // [2][ngroups_total]
    int ngroups_total;
    long long* timing;            // [8] cycle counters accumulated by CTA 0 (vs_debug_lstm_timing): see below
};
// timing[0] producer: spin on the group barrier     timing[1] producer: issuing TMA (incl. waiting for a free stage)
// timing[2] MMA thread: waiting for h K-blocks      timing[3] MMA thread: issuing MMAs + commits
// timing[4] cell thread 64: waiting for acc_full    timing[5] cell thread 64: TMEM load + gate math
// timing[6] cell thread 64: stores                  timing[7] cell thread 64: bar.sync + fence + atomic

// TIMING: phase timers (clock64 around every wait / issue / store phase) for vs_debug_lstm_timing - compiled out of the product
// kernel, enabled with VOICESPLIT_LSTM_TIMING=1 (tools/lstm_timing.py)
#define VS_CLK() (TIMING ? clock64() : 0ll)
template <int ELT, bool TIMING>
__global__ void __launch_bounds__(64 + 128 * kLGroups, 1) k_lstm_tc(const LstmTcArgs a, const __grid_constant__ CUtensorMap tm_w_hi,
                                                    const __grid_constant__ CUtensorMap tm_w_lo,
                                                    const __grid_constant__ CUtensorMap tm_h_hi,
                                                    const __grid_constant__ CUtensorMap tm_h_lo) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int nplanes = a.passes == 3 ? 2 : 1;
    uint8_t* w_smem = smem;                                            // [plane][kb][64 rows][128 B]
    uint8_t* a_ring = smem + (size_t)nplanes * a.nkb * 8192;           // [stage][plane][128 rows][128 B]
    const int stage_bytes = nplanes * 16384;
    uint64_t* bars = reinterpret_cast<uint64_t*>(a_ring + (size_t)kLStages * stage_bytes);
    uint64_t* a_full = bars;
    uint64_t* a_empty = bars + kLStages;
    uint64_t* w_full = a_empty + kLStages;
    uint64_t* acc_full = w_full + 1;                                   // [kLGroups]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + kLGroups);

    // CTA coordinates
    int cta = blockIdx.x;
    const int slice = cta % a.nslices; cta /= a.nslices;
    const int set = cta % a.nsets; cta /= a.nsets;
    const int d = cta;
    const int grp0 = a.group0 + set * a.gpc;                           // first group of this CTA
    const int ng = min(a.gpc, a.ngroups_total - grp0);                 // groups this CTA really has (the last set may hold one)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned int* counter0 = a.barrier + d * a.ngroups_total + grp0;   // counter of group gi: counter0 + gi

    if (threadIdx.x == 0) {
        for (int i = 0; i < kLStages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        mbar_init(w_full, 1);
        for (int i = 0; i < kLGroups; ++i) mbar_init(&acc_full[i], 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 64 * kLGroups);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ---------------- producer: W slice once, then h K-blocks every step ----------------
        // the whole warp polls / waits (warp-uniform), one elected lane issues the TMA instructions (no per-instruction
        // elect-and-loop sequences as inside an `if (lane == 0)` region)
        {
            if (elect_one()) {
                mbar_arrive_expect_tx(w_full, (uint32_t)(nplanes * a.nkb * 8192));
                for (int p = 0; p < nplanes; ++p)
                    for (int kb = 0; kb < a.nkb; ++kb)
                        tma_load_2d(w_smem + (size_t)(p * a.nkb + kb) * 8192, p == 0 ? &tm_w_hi : &tm_w_lo, w_full, kb * 64,
                                    (d * a.nslices + slice) * 64);
            }
            __syncwarp();
            int st = 0, ph = 0;
            long long tm_spin = 0, tm_issue = 0;
            for (int s = 1; s < a.T; ++s) {
                for (int gi = 0; gi < ng; ++gi) {
                    // wait until every slice of this (direction, group) has published h_{s-1}
                    const unsigned int target = (unsigned int)s * a.nslices;
                    unsigned int spins = 0;
                    const long long c0 = VS_CLK();
                    for (;;) {      // acquire loads: no separate gpu-scope fence (an extra L2 round trip) between the flag and the TMA issue
                        unsigned int seen;
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter0 + gi) : "memory");
                        if (seen >= target) break;
                        if (++spins > (1u << 28)) __trap();
                    }
                    const long long c1 = VS_CLK();
                    tm_spin += c1 - c0;
                    asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy flag read -> async-proxy (TMA) data reads
                    const int par = (s - 1) & 1;                       // buffer h_{s-1} was written to
                    const int row0 = ((d * 2 + par) * a.Bp) + (grp0 + gi) * kLB;
                    for (int kb = 0; kb < a.nkb; ++kb) {
                        mbar_wait(&a_empty[st], ph ^ 1);
                        if (elect_one()) {
                            mbar_arrive_expect_tx(&a_full[st], (uint32_t)stage_bytes);
                            uint8_t* dst = a_ring + (size_t)st * stage_bytes;
                            tma_load_2d(dst, &tm_h_hi, &a_full[st], kb * 64, row0);
                            if (nplanes == 2) tma_load_2d(dst + 16384, &tm_h_lo, &a_full[st], kb * 64, row0);
                        }
                        __syncwarp();
                        if (++st == kLStages) { st = 0; ph ^= 1; }
                    }
                    tm_issue += VS_CLK() - c1;
                }
            }
            if (TIMING && blockIdx.x == 0 && lane == 0 && a.timing) { a.timing[0] = tm_spin; a.timing[1] = tm_issue; }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer: whole warp waits, one elected lane issues ----------------
        {
            const uint32_t idesc = make_idesc_bf16(128, 64, ELT);
            mbar_wait(w_full, 0);
            tc_fence_after();
            const uint32_t w_addr = smem_u32(w_smem);
            int st = 0, ph = 0;
            long long tm_wait = 0, tm_mma = 0;
            for (int s = 1; s < a.T; ++s)
            for (int gi = 0; gi < ng; ++gi) {
                uint32_t accumulate = 0;
                const uint32_t d_t = tmem + (uint32_t)(gi * 64);
                for (int kb = 0; kb < a.nkb; ++kb) {
                    const long long c0 = VS_CLK();
                    mbar_wait(&a_full[st], ph);
                    const long long c1 = VS_CLK();
                    tm_wait += c1 - c0;
                    tc_fence_after();
                    const uint32_t h_hi = smem_u32(a_ring + (size_t)st * stage_bytes), h_lo = h_hi + 16384;
                    const uint32_t w_hi = w_addr + (uint32_t)kb * 8192, w_lo = w_addr + (uint32_t)(a.nkb + kb) * 8192;
                    if (elect_one()) {
                        const uint64_t d_hh = make_smem_desc(h_hi, 16, 1024, 2), d_hl = make_smem_desc(h_lo, 16, 1024, 2);
                        const uint64_t d_wh = make_smem_desc(w_hi, 16, 1024, 2), d_wl = make_smem_desc(w_lo, 16, 1024, 2);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (kb * 4 + k < a.nk16) {
                                umma_bf16(d_t, d_hh + 2 * k, d_wh + 2 * k, idesc, k == 0 ? accumulate : 1u);
                                if (nplanes == 2) {
                                    umma_bf16(d_t, d_hl + 2 * k, d_wh + 2 * k, idesc, 1);
                                    umma_bf16(d_t, d_hh + 2 * k, d_wl + 2 * k, idesc, 1);
                                }
                            }
                        }
                        umma_commit(&a_empty[st]);
                        if (kb == a.nkb - 1) umma_commit(&acc_full[gi]);
                    }
                    __syncwarp();
                    accumulate = 1;
                    if (++st == kLStages) { st = 0; ph ^= 1; }
                    tm_mma += VS_CLK() - c1;
                }
            }
            if (TIMING && blockIdx.x == 0 && lane == 0 && a.timing) { a.timing[2] = tm_wait; a.timing[3] = tm_mma; }
        }
    } else {
        // ---------------- cell update: thread = one utterance, 16 units ----------------
        const int gi = (warp - 2) >> 2;                      // which of the CTA's groups this warp belongs to
        if (gi < ng) {                                       // (the last set may hold a single group: its second cell team has no work)
        const int quad = warp & 3;
        const int b = (grp0 + gi) * kLB + quad * 32 + lane;  // utterance handled by this thread
        const bool valid = b < a.B;
        unsigned int* counter = counter0 + gi;
        const bool timed = gi == 0;
        const int u0 = slice * kLU;                          // first hidden unit of the slice
        const int nu = min(kLU, a.H - u0);                   // valid units in this slice
        float c[kLU];
#pragma unroll
        for (int j = 0; j < kLU; ++j) c[j] = 0.f;
        const uint32_t t_base = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(gi * 64);
        long long tm_acc = 0, tm_math = 0, tm_store = 0, tm_bar = 0;
        for (int s = 0; s < a.T; ++s) {
            const int t = d ? a.T - 1 - s : s;
            // input projection for this step (issued before waiting on the MMA)
            float gx[4][kLU];
            const float* gsrc = a.gates_x + ((size_t)(valid ? b : 0) * a.T + t) * 8 * a.H + (size_t)d * 4 * a.H + u0;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                if (nu == kLU && (a.H & 3) == 0) {
#pragma unroll
                    for (int j = 0; j < kLU; j += 4) {
                        float4 v = __ldg(reinterpret_cast<const float4*>(gsrc + (size_t)g * a.H + j));
                        gx[g][j] = v.x; gx[g][j + 1] = v.y; gx[g][j + 2] = v.z; gx[g][j + 3] = v.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < kLU; ++j) gx[g][j] = j < nu ? __ldg(gsrc + (size_t)g * a.H + j) : 0.f;
                }
            }
            long long c0 = VS_CLK();
            if (s > 0) {
                mbar_wait(&acc_full[gi], (s - 1) & 1);
                tm_acc += VS_CLK() - c0;
                c0 = VS_CLK();
                tc_fence_after();
                uint32_t r0[32], r1[32];
                tmem_ld_32x32(t_base, r0);
                tmem_ld_32x32(t_base + 32, r1);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < kLU; ++j) {
                    gx[0][j] += __uint_as_float(r0[j]);
                    gx[1][j] += __uint_as_float(r0[16 + j]);
                    gx[2][j] += __uint_as_float(r1[j]);
                    gx[3][j] += __uint_as_float(r1[16 + j]);
                }
                tc_fence_before();
            }
            const int par = s & 1;
            elt16* hx_hi = a.hx_hi + ((size_t)(d * 2 + par) * a.Bp + (valid ? b : 0)) * a.Hp + u0;
            elt16* hx_lo = a.hx_lo ? a.hx_lo + ((size_t)(d * 2 + par) * a.Bp + (valid ? b : 0)) * a.Hp + u0 : nullptr;
            const size_t oidx = ((size_t)(valid ? b : 0) * a.T + t) * 2 * a.H + (size_t)d * a.H + u0;
            __align__(16) elt16 vh[kLU], vl[kLU], rh[kLU], rl[kLU];
            float hv[kLU];
#pragma unroll
            for (int j = 0; j < kLU; ++j) {
                const float ig = sigmoid_fast(gx[0][j]), fg = sigmoid_fast(gx[1][j]);
                const float gg = tanh_fast(gx[2][j]), og = sigmoid_fast(gx[3][j]);
                c[j] = fmaf(fg, c[j], ig * gg);
                hv[j] = og * tanh_fast(c[j]);
                split16<ELT>(hv[j], vh[j], vl[j]);
                split16<ELT>(fmaxf(hv[j], 0.f), rh[j], rl[j]);
            }
            {
                const long long c1 = VS_CLK();
                tm_math += c1 - c0;
                c0 = c1;
            }
            // h_t for the next step first: only these 64 bytes per utterance are on the critical path of the exchange.  The fp32
            // lstm_out and the relu(h) planes of the FC head follow AFTER the release below - their (row-strided, half-sector)
            // stores then drain while the other slices are still being waited for, instead of delaying this CTA's fence.
            const bool fast = nu == kLU && (a.H & 7) == 0;
            if (valid) {
                if (fast) {
                    *reinterpret_cast<uint4*>(hx_hi) = *reinterpret_cast<const uint4*>(vh);
                    *reinterpret_cast<uint4*>(hx_hi + 8) = *reinterpret_cast<const uint4*>(vh + 8);
                    if (hx_lo) {
                        *reinterpret_cast<uint4*>(hx_lo) = *reinterpret_cast<const uint4*>(vl);
                        *reinterpret_cast<uint4*>(hx_lo + 8) = *reinterpret_cast<const uint4*>(vl + 8);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < kLU; ++j) {
                        if (j < nu) {
                            hx_hi[j] = vh[j];
                            if (hx_lo) hx_lo[j] = vl[j];
                        }
                    }
                }
            }
            {
                const long long c1 = VS_CLK();
                tm_store += c1 - c0;
                c0 = c1;
            }
            if (s + 1 < a.T) {
                // publish h_s: the 128 cell threads meet at a CTA barrier, then ONE thread issues the
                // gpu-scope release (cumulative over the stores ordered before the barrier) and bumps
                // the counter - the cooperative-groups grid-sync pattern, without a fence per thread
                asm volatile("bar.sync %0, 128;" ::"r"(1 + gi) : "memory");
                if (threadIdx.x == 64 + 128 * gi) {
                    __threadfence();
                    atomicAdd(counter, 1u);
                }
                tm_bar += VS_CLK() - c0;
            }
            if (valid) {
                if (fast) {
#pragma unroll
                    for (int j = 0; j < kLU; j += 4)
                        *reinterpret_cast<float4*>(a.hout + oidx + j) = make_float4(hv[j], hv[j + 1], hv[j + 2], hv[j + 3]);
                    if (a.hr_hi) {
                        *reinterpret_cast<uint4*>(a.hr_hi + oidx) = *reinterpret_cast<const uint4*>(rh);
                        *reinterpret_cast<uint4*>(a.hr_hi + oidx + 8) = *reinterpret_cast<const uint4*>(rh + 8);
                        if (a.hr_lo) {
                            *reinterpret_cast<uint4*>(a.hr_lo + oidx) = *reinterpret_cast<const uint4*>(rl);
                            *reinterpret_cast<uint4*>(a.hr_lo + oidx + 8) = *reinterpret_cast<const uint4*>(rl + 8);
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < kLU; ++j) {
                        if (j < nu) {
                            a.hout[oidx + j] = hv[j];
                            if (a.hr_hi) { a.hr_hi[oidx + j] = rh[j]; if (a.hr_lo) a.hr_lo[oidx + j] = rl[j]; }
                        }
                    }
                }
            }
        }
        if (TIMING && timed && blockIdx.x == 0 && threadIdx.x == 64 && a.timing) { a.timing[4] = tm_acc; a.timing[5] = tm_math; a.timing[6] = tm_store; a.timing[7] = tm_bar; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 64 * kLGroups);
}

// W_hh [2][4H][H] fp32 -> [2][nslices*64][Hp] 16-bit hi/lo, row = slice*64 + gate*16 + j
__global__ void k_pack_whh_tc(const float* __restrict__ whh, int H, int Hp, int nslices, elt16* __restrict__ bhi, elt16* __restrict__ blo,
                              elt16* __restrict__ hhi, elt16* __restrict__ hlo) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = (long long)2 * nslices * 64 * Hp;
    if (i >= n) return;
    int k = (int)(i % Hp);
    long long r = i / Hp;
    int row = (int)(r % 64), sl = (int)((r / 64) % nslices), d = (int)(r / (64LL * nslices));
    int g = row / kLU, j = row % kLU, u = sl * kLU + j;
    float v = (u < H && k < H) ? whh[((size_t)d * 4 * H + (size_t)g * H + u) * H + k] : 0.f;
    split16<0>(v, bhi[i], blo[i]);
    split16<1>(v, hhi[i], hlo[i]);
}

struct LstmState {
    elt16 *w_hi[2] = {}, *w_lo[2] = {};
    long long* timing = nullptr;
    int max_smem = 0;
};

int tc_lstm_pack(vs_engine* e, void** slot, cudaStream_t st) {
    if (!*slot) {
        LstmState* s = new LstmState();
        cudaDeviceGetAttribute(&s->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, e->device);
        *slot = s;
    }
    LstmState* s = (LstmState*)*slot;
    if (!s->timing) VS_CUDA_TRY(cudaMalloc(&s->timing, 8 * sizeof(long long)));
    const int H = e->d.lstm_dim, Hp = (H + 7) / 8 * 8, nslices = (H + kLU - 1) / kLU;
    const size_t n = (size_t)2 * nslices * 64 * Hp;
    for (int t = 0; t < 2; ++t) {
        if (!s->w_hi[t]) {
            VS_CUDA_TRY(cudaMalloc(&s->w_hi[t], n * sizeof(elt16)));
            VS_CUDA_TRY(cudaMalloc(&s->w_lo[t], n * sizeof(elt16)));
        }
    }
    k_pack_whh_tc<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(e->whh, H, Hp, nslices, s->w_hi[0], s->w_lo[0], s->w_hi[1], s->w_lo[1]);
    VS_CUDA_TRY(cudaGetLastError());
    return VS_OK;
}
void tc_lstm_destroy(void* slot) {
    LstmState* s = (LstmState*)slot;
    if (!s) return;
    for (int t = 0; t < 2; ++t) { cudaFree(s->w_hi[t]); cudaFree(s->w_lo[t]); }
    cudaFree(s->timing);
    delete s;
}

size_t tc_lstm_scratch_bytes(const vs_engine* e, int B) {
    const int H = e->d.lstm_dim, Hp = (H + 7) / 8 * 8;
    const size_t Bp = align_up((size_t)B, kLB);
    // hi + lo exchange buffers [2][2][Bp][Hp] + barrier counters
    return 2 * (size_t)4 * Bp * Hp * sizeof(elt16) + 4096;
}

int tc_lstm_recurrence(vs_engine* e, void* slot, const float* gates_x, float* hout, void* scratch, elt16* hr_hi, elt16* hr_lo,
                       int B, int T, int precision, cudaStream_t st) {
    LstmState* s = (LstmState*)slot;
    const int H = e->d.lstm_dim, Hp = (H + 7) / 8 * 8, nslices = (H + kLU - 1) / kLU;
    const int elt = tc_elt(precision), passes = tc_passes(precision);
    const int Bp = (int)align_up((size_t)B, kLB);
    const int ngroups_total = Bp / kLB;
    LstmTcArgs a{};
    a.B = B; a.T = T; a.H = H; a.nslices = nslices; a.nkb = (H + 63) / 64; a.nk16 = (H + 15) / 16; a.passes = passes;
    a.Bp = Bp; a.Hp = Hp; a.gates_x = gates_x; a.hout = hout;
    char* p = (char*)scratch;
    const size_t plane = (size_t)4 * Bp * Hp * sizeof(elt16);
    a.hx_hi = (elt16*)p;
    a.hx_lo = passes == 3 ? (elt16*)(p + plane) : nullptr;
    a.barrier = (unsigned int*)(p + 2 * plane);
    a.hr_hi = hr_hi; a.hr_lo = passes == 3 ? hr_lo : nullptr;
    a.ngroups_total = ngroups_total;
    a.timing = s->timing;
    if (2 * ngroups_total * (int)sizeof(unsigned int) > 4096) { set_error("batch too large for the LSTM barrier table"); return VS_ERR_INVALID; }
    const int nplanes = passes == 3 ? 2 : 1;
    const int smem = 1024 + nplanes * a.nkb * 8192 + kLStages * nplanes * 16384 + 256;
    if (smem > s->max_smem) { set_error("lstm_dim too large for the tensor-core recurrent kernel"); return VS_ERR_UNSUPPORTED; }
    // CTA sets of `gpc` groups are independent: as many as are co-resident run in one launch (2 directions x nslices CTAs per set).
    // gpc = 2 (VOICESPLIT_LSTM_GROUPS_PER_CTA=2) interleaves two groups on one CTA; measured at B = 256 (same box, ncu cycles):
    // 11.53 M cycles against 10.64 M with one group per CTA on twice the SMs - the two chains convoy through the in-order producer /
    // issuer instead of hiding each other's exchange (profiles/r02_lstm_two_groups_ab.txt) - so one group per CTA stays the default.
    static const int gpc_env = getenv("VOICESPLIT_LSTM_GROUPS_PER_CTA") ? atoi(getenv("VOICESPLIT_LSTM_GROUPS_PER_CTA")) : 1;
    a.gpc = (ngroups_total >= 2 && gpc_env >= 2) ? kLGroups : 1;
    const int max_sets = e->num_sms / (2 * nslices);
    if (max_sets < 1) { set_error("lstm_dim too large: one step of all slices must be co-resident"); return VS_ERR_UNSUPPORTED; }
    CUtensorMap tm_w_hi, tm_w_lo, tm_h_hi, tm_h_lo;
    {
        uint64_t wd[2] = {(uint64_t)H, (uint64_t)2 * nslices * 64}, ws[1] = {(uint64_t)Hp * sizeof(elt16)};
        uint32_t wb[2] = {64, 64};
        uint64_t hd[2] = {(uint64_t)H, (uint64_t)4 * Bp}, hs[1] = {(uint64_t)Hp * sizeof(elt16)};
        uint32_t hb[2] = {64, 128};
        bool ok = make_tmap_bf16(&tm_w_hi, s->w_hi[elt], 2, wd, ws, wb, CU_TENSOR_MAP_SWIZZLE_128B);
        ok = ok && make_tmap_bf16(&tm_w_lo, s->w_lo[elt], 2, wd, ws, wb, CU_TENSOR_MAP_SWIZZLE_128B);
        ok = ok && make_tmap_bf16(&tm_h_hi, a.hx_hi, 2, hd, hs, hb, CU_TENSOR_MAP_SWIZZLE_128B);
        ok = ok && make_tmap_bf16(&tm_h_lo, a.hx_lo ? a.hx_lo : a.hx_hi, 2, hd, hs, hb, CU_TENSOR_MAP_SWIZZLE_128B);
        if (!ok) { set_error("cuTensorMapEncodeTiled failed (lstm)"); return VS_ERR_CUDA; }
    }
    cudaError_t ce = cudaMemsetAsync(a.barrier, 0, 4096, st);
    if (ce != cudaSuccess) { set_error(cudaGetErrorString(ce)); return VS_ERR_CUDA; }
    static const bool timing = getenv("VOICESPLIT_LSTM_TIMING") && atoi(getenv("VOICESPLIT_LSTM_TIMING")) != 0;
    const void* fn = timing ? (elt ? (const void*)k_lstm_tc<1, true> : (const void*)k_lstm_tc<0, true>)
                            : (elt ? (const void*)k_lstm_tc<1, false> : (const void*)k_lstm_tc<0, false>);
    ce = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (ce != cudaSuccess) { set_error(cudaGetErrorString(ce)); return VS_ERR_CUDA; }
    const int nsets_total = (ngroups_total + a.gpc - 1) / a.gpc;
    for (int s0 = 0; s0 < nsets_total; s0 += max_sets) {
        a.group0 = s0 * a.gpc;
        a.nsets = nsets_total - s0 < max_sets ? nsets_total - s0 : max_sets;
        void* args[] = {(void*)&a, (void*)&tm_w_hi, (void*)&tm_w_lo, (void*)&tm_h_hi, (void*)&tm_h_lo};
        ce = cudaLaunchCooperativeKernel(fn, dim3(2 * a.nsets * nslices), dim3(64 + 128 * a.gpc), args, (size_t)smem, st);
        if (ce != cudaSuccess) { set_error(std::string("k_lstm_tc launch: ") + cudaGetErrorString(ce)); return VS_ERR_CUDA; }
        e->launches++;
    }
    if (e->profiling) prof_after(e, KID_LSTM_REC, st);
    return VS_OK;
}

int tc_lstm_read_timing(void* slot, long long* out8) {
    LstmState* s = (LstmState*)slot;
    if (!s || !s->timing) return VS_ERR_STATE;
    return cudaMemcpy(out8, s->timing, 8 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? VS_OK : VS_ERR_CUDA;
}

}  // namespace vs
