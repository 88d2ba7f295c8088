// pfb_tma.cuh -- TMA / mbarrier variant of the time-chunked SOS bank for shared-input banks.
//
// Why a second scan kernel: profiling the cp.async kernel (profiles/) showed it bound by DRAM
// traffic of scattered 128-byte rows -- every (channel, band) CTA re-read the channel's input, and
// rows of 128 bytes reach only ~3.9 TB/s on B200 where 256-byte rows reach ~5.5 TB/s
// (scripts/ubench/stream_rows.cu).  Here
//   * one CTA owns one channel, one time chunk group (32 chunks) and a GROUP of bands: a producer
//     warp streams the channel's input tiles [32 chunks x 32 samples] into a shared-memory ring with
//     TMA (cp.async.bulk.tensor, 128-byte swizzle, mbarrier complete_tx) and every band warp of the
//     CTA reads them from there, so the input is fetched once per band group instead of once per band;
//   * every band warp filters its lane's row in registers and hands its [32 x 256-byte] output tile
//     to TMA stores, so no thread spends instructions on global addressing and HBM sees 256-byte rows;
//   * the work is split into three launches whose CTAs are all independent (pre-pass A, carry, output
//     pass B), so the hardware scheduler balances the band-dependent pre-pass cost over the 148 SMs.
// The arithmetic, the chunk-transition carry and the truncated pre-pass are the ones of
// sos_scan_kernel (pfb_kernels.cuh); states and outputs are identical up to the chunk length.
#pragma once
#include <cuda.h>

#include "pfb_kernels.cuh"

namespace pfb {

constexpr int kTmaStages = 3;
constexpr int kTmaMaxBands = 11;               // most band (consumer) warps per CTA; +1 producer warp
constexpr int kTmaSmallBands = 7;              // CTA shape with 112 registers per thread (no spills for K <= 6)

struct TmaParams {
    void *zs;             // [Q][J][2K] arithmetic type: pre-pass states, then chunk start states
    const void *P;        // [B][2K][2K]
    void *states;         // [Q][Ktot][2]
    const int *warm;      // [B]
    long long L;          // chunk length, multiple of the tile time length
    int Q, B, C;
    int J;                // chunks per cascade, a multiple of JR
    int JR, CH;           // a tile's 32 rows are JR consecutive chunks of CH consecutive channels (JR * CH == 32):
                          // many channels -> few (or one) chunks per cascade and (almost) no pre-pass
    int NW, CB;           // chunk groups J / JR, channel blocks ceil(C / CH)
    int nb;               // band warps per CTA (output pass)
    int G;                // band groups = ceil(B / nb)
    int nb_a, G_a;        // the same for the pre-pass, whose CTAs are smaller: bands of one CTA should
                          // have similar pre-pass lengths, or their warps idle while the longest one runs
    int Ktot, k0;
    // Gain-normalised arithmetic (float64 only): section k runs on states scaled by 1 / (b0_0 ... b0_{k-1})
    // with numerator (1, b1/b0, b2/b0), which saves the multiply by b0 in every section (17 instead of
    // 20 FP64 operations per band sample for 4 sections; the product of the gains is applied once at
    // the output).  gains [B][K][2] = (G_k, 1 / G_k), G_k = b0_0 ... b0_{k-1}; null: plain arithmetic.
    const double *gains;
    // Chunk-impulse-response (GEMM) form of the pre-pass, float64 arithmetic: z = H x over the pre-pass window.
    // H: per band, per 4-sample step, MT tiles of 32 doubles in the m8n8k4 A-fragment order
    // (lane l <-> state 8 mt + l / 4, sample l % 4); Hoff[b]: band b's offset in doubles; null: recurrence pre-pass.
    const double *H;
    const long long *Hoff;
    // Band-level (energy) epilogue instead of the output store: sums of y^2 per `ut` tiles, written time-ordered to
    // levels[q * lev_ld + lev_off + j * (L / unit) + u]  (unit = ut tile steps; chunk j, unit u of the chunk)
    double *levels;
    long long lev_ld, lev_off;
    int ut;
    int wide;             // output pass with 512-byte fragments (sos_tma_wide_kernel); L is then a multiple of 4 half rows
    // Fused SPL-weighting prefilter (WEIGHT kernels, splweighting.py:61-80): with one chunk per cascade a tile row is
    // a whole channel, so one extra warp runs the sequential transfer-function recurrence (scipy's operation
    // order, tf_step) in place on every input tile before the band warps read it -- the weighted signal never
    // touches HBM.  wz_in [C][8]: delay values at the start of the call (a copy: band group 0 overwrites wz
    // [C][wnz] with the final ones while other groups may still have to start); taps normalised, wa[0] == 1.
    int weighted, wnz;
    const double *wz_in;
    double *wz;
    double wb[8], wa[8];
    // Decimating band (DEC kernels; the anti-alias stage of the decimating band path): band dec_band1 - 1 stores
    // only its samples of parity dec_phase, through the second output map (time L / 2); 0: none.  Its states live
    // in dec_states [C][Ktot][2]; the states array then holds Bs = B - 1 bands (Bs == 0 means B).
    int dec_band1, dec_phase, Bs;
    void *dec_states;
    // Output pass: every lane copies its own row fragment to global memory with a one-dimensional bulk copy
    // (tma_cta, kBulkRows): y base, row pitch in elements, bands stored through it (B, or B - 1 with a decimating band)
    void *y;
    long long ldy;
    int By;
    // PFB_ARITH_AUTO (float32 signals, float64 states): bands whose bit is set in f32mask run the output pass in float32
    // arithmetic with the plain coefficients cf32 [B][K][5] (float); null: every band in the kernel's arithmetic type
    const unsigned *f32mask;
    const float *cf32;
    // PFB_WARP_PROF builds only: clocks per warp role, [warp index][0 wait for input, 1 arithmetic, 2 wait for the
    // output buffer, 3 total, 4 proxy fence + warp sync, 5 barrier arrive + store issue], summed over CTAs (lane 0's clock64)
    unsigned long long *prof;
};

#ifndef PFB_WARP_PROF
#define PFB_WARP_PROF 0
#endif
#if PFB_WARP_PROF
#define PROF_DECL long long prof_t[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long prof_c = clock64(); const long long prof_0 = prof_c
#define PROF_LAP(i) do { const long long prof_n = clock64(); prof_t[i] += prof_n - prof_c; prof_c = prof_n; } while (0)
#define PROF_END(p, warp, lane) do { if ((p).prof && (lane) == 0) { prof_t[3] = clock64() - prof_0; \
    for (int prof_i = 0; prof_i < 8; ++prof_i) atomicAdd((p).prof + (warp) * 8 + prof_i, (unsigned long long)prof_t[prof_i]); } } while (0)
#else
#define PROF_DECL do {} while (0)
#define PROF_LAP(i) do {} while (0)
#define PROF_END(p, warp, lane) do {} while (0)
#endif

// States (w1, w2 per section, sections k0..) of cascade (channel, band).
template <typename A>
__device__ __forceinline__ A *cascade_states(const TmaParams &p, int chn, int band) {
    if (p.dec_band1 != 0 && band == p.dec_band1 - 1)
        return reinterpret_cast<A *>(p.dec_states) + ((size_t)chn * p.Ktot + p.k0) * 2;
    const int Bs = p.Bs ? p.Bs : p.B;
    return reinterpret_cast<A *>(p.states) + (((size_t)chn * Bs + band) * p.Ktot + p.k0) * 2;
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_addr(bar)) : "memory");
}
// Release of an input stage by a warp that only READ it (pre-pass and band-level kernels; the output pass has a proxy
// fence between its loads and the release anyway).  Every lane fences first: a lane's loads of the stage must have
// been performed before lane 0's arrive lets the producer refill it.  Without it ptxas schedules the arrive right behind
// the ISSUE of the last LDS and the refill can overtake loads still queued in the SM's shared-memory pipeline.
__device__ __forceinline__ void stage_release_after_reads(uint64_t *bar, int lane) {
    asm volatile("fence.acq_rel.cta;\n" ::: "memory");
    if (lane == 0) mbar_arrive(bar);
}
// One non-blocking phase test (no branch inside the asm: the caller decides what to do with the result).
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Wait of a single thread (the producer lane).  The loop is C++, not a branch inside the asm: ptxas must see the
// control flow, or it treats the code behind a single-lane spin as converged (see mbar_wait_warp).
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) {
#if PFB_PRODUCER_BACKOFF
        __nanosleep(PFB_PRODUCER_BACKOFF);     // the producer is never latency critical (the ring is full when it waits)
#endif
    }
}
// Wait of a whole consumer warp: every lane tests the phase (no branch inside the asm) and a warp vote ends the loop
// uniformly, so the warp stays one converged group and every lane has itself observed (acquired) the phase.
//
// History, all measured.
// (0) All 32 lanes spinning in a branch-inside-asm loop: ptxas compiles the loop as if the warp were converged behind
//     it and removes every later __syncwarp(); with 3-4 CTAs per SM the gammatone pre-pass then gave chunk states that
//     differed between runs of the same launch (scripts/ubench/fos_race.cu, profiles/r01h_fos_race.txt).
// (2) Lane 0 polls, __syncwarp(), every lane waits once more (rounds 1h-2a): reproducible, but the warp leaves that
//     sequence SPLIT -- lane 0 and lanes 1-31 run on as two groups until the next __syncwarp() and everything in
//     between is issued twice.  Clocks of lane 0 and lane 31 of the weighting warp (profiles/r02l_warp_prof.txt):
//     both spend 2160 clocks per tile in the arithmetic, one after the other.  That was the unexplained factor 2 of
//     round 1 (recurrence pre-pass 6.5 -> 14.0 ms, band levels, every fully unrolled output pass 40-100 % slower).
// (4) = (2) plus a closing __syncwarp(): reproducible and converged, but 10-40 % slower than (3) on the kernels with one
//     CTA per SM (cfg 5 18.2 / 16.1 ms against 14.4 / 9.3 ms, gammatone cfg 4 20.5 against 18.2 ms; profiles/r02o).
// (3) This form.  ptxas again removes the __syncwarp()s (the warp IS converged now), which exposed what the race of
//     (0) really was: in the pre-pass kernels nothing orders the release of a stage (lane 0's mbarrier.arrive) behind the
//     shared-memory LOADS of the tile -- ptxas schedules the arrive right behind the last LDS *issue*, the loads of 28
//     warps queue up for hundreds of clocks, and the producer's refill lands first.  Experiments (profiles/
//     r02p_wait_experiments.txt): delays, proxy fences, more try_waits and a release by every lane change nothing; the
//     pre-pass alone in form (4) removes the effect; so does a fence.acq_rel.cta of every lane in front of the arrive
//     (stage_release_after_reads), which is what every release without a proxy fence in front of it now does.
#ifndef PFB_WAIT_MODE
#define PFB_WAIT_MODE 3
#endif
__device__ __forceinline__ void mbar_wait_warp(uint64_t *bar, unsigned parity, int lane) {
#if PFB_WAIT_MODE == 0
    mbar_wait(bar, parity);
#elif PFB_WAIT_MODE == 3
    while (!__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) {
    }
#else
    if (lane == 0) mbar_wait(bar, parity);
    __syncwarp();
#if PFB_WAIT_MODE >= 2
    mbar_wait(bar, parity);
#endif
#if PFB_WAIT_MODE == 4
    __syncwarp();
#endif
#endif
}
// Tiles that are only released (not read): lane 0 waits and arrives alone, the rest of the warp does nothing.
__device__ __forceinline__ void mbar_skip_tile(uint64_t *full_bar, uint64_t *empty_bar, unsigned parity, int lane) {
#if PFB_WAIT_MODE == 3
    mbar_wait_warp(full_bar, parity, lane);          // the whole warp, converged; one lane releases
    if (lane == 0) mbar_arrive(empty_bar);
#else
    if (lane == 0) {
        mbar_wait(full_bar, parity);
        mbar_arrive(empty_bar);
    }
#endif
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::
            "r"(smem_addr(dst)),
        "l"(map), "r"(smem_addr(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, const void *src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];\n" ::"l"(map),
                 "r"(smem_addr(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, const void *src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n" ::"l"(map),
                 "r"(smem_addr(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
// one-dimensional bulk copy shared -> global of `bytes` (multiple of 16) contiguous bytes, this thread's bulk group
__device__ __forceinline__ void bulk_store_1d(void *dst, const void *src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst), "r"(smem_addr(src)), "r"(bytes)
                 : "memory");
}
// Output tile of one band warp when the lanes copy their own rows (kBulkRows): 32 rows of NH * 128 bytes, 16 bytes of
// padding per row (a lane-per-row writer with 16-byte stores then needs the minimum of four shared-memory wavefronts),
// slots 1024-byte aligned so that the decimating band can keep its swizzled tensor-store tile in the same slot.
__host__ __device__ constexpr int out_row_stride(int NH) { return NH * kRowBytes + 16; }
__host__ __device__ constexpr int out_slot_bytes(int NH) { return (32 * out_row_stride(NH) + 1023) / 1024 * 1024; }
// Which output passes copy rows per lane (see tma_cta), and the shared memory of one band warp's output tile
#ifndef PFB_BULK_ROWS
#define PFB_BULK_ROWS -1     // -1: per kernel (below); 0 / 1: never / always (A/B builds, profiles/r03d_bulk_rows.txt)
#endif
template <typename T, typename A>
__host__ __device__ constexpr bool bulk_rows(int NH, bool dec, bool weight) {
    if (dec) return false;     // the decimating band keeps its swizzled tensor-store tile in the same slot
    if (PFB_BULK_ROWS >= 0) return PFB_BULK_ROWS != 0;
    return !weight && NH == 4 && sizeof(T) == 8 && sizeof(A) == 8;      // the wide float64 output pass
}
template <typename T, typename A>
__host__ __device__ constexpr int out_tile_bytes(int NH, bool dec, bool weight) {
    return bulk_rows<T, A>(NH, dec, weight) ? out_slot_bytes(NH) : NH * kTileBytes;
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory"); }
// wait until at most `pending` of this thread's most recent bulk groups have not yet been read from shared memory
__device__ __forceinline__ void bulk_wait_read_upto(int pending) {
    switch (pending) {
        case 0: asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); break;
        case 1: asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory"); break;
        case 2: asm volatile("cp.async.bulk.wait_group.read 2;\n" ::: "memory"); break;
        default: asm volatile("cp.async.bulk.wait_group.read 3;\n" ::: "memory"); break;
    }
}
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// One section in gain-normalised form: c = {total gain (section 0 only), b1/b0 - a1, b2/b0 - a2, a1, a2}.
template <typename A>
__device__ __forceinline__ A scaled_step(A x, const A *c, A &w1, A &w2) {
    // c = {., b1/b0 - a1, b2/b0 - a2, a1, a2}: the output is formed from the input and the OLD states, not from w0 --
    // the same four FMAs, but x -> y is two dependent operations instead of four, which halves the latency of a
    // wavefront step of the cascade (a band warp that has its SM sub-partition's FP64 pipe to itself or shares it with
    // one other is bound by that chain, not by the pipe); and y no longer cancels w0 - 2 w1 + w2 at the states'
    // magnitude (1e6 times the signal in the low bands) but at that of its own terms (~300 times).
    const A t = fma(-c[4], w2, x);
    const A w0 = fma(-c[3], w1, t);
    const A y = fma(c[1], w1, fma(c[2], w2, x));
    w2 = w1;
    w1 = w0;
    return y;
}

// One section step of the arithmetic the TMA kernels run (gain-normalised or plain FMA form).
template <typename A, bool SCALED>
__device__ __forceinline__ A section_step(A x, const A *c, A &w1, A &w2) {
    return SCALED ? scaled_step<A>(x, c, w1, w2) : Biquad<A, false>::step(x, c, w1, w2);
}

// All samples of a 128-byte half row (e[0 .. NS), in registers) through the K sections, in place, in WAVEFRONT
// order: step d runs section k on sample d - k for every k.  The K updates of one step are independent of each other
// (section k reads what section k - 1 produced one step earlier), so a single warp carries K-fold instruction-level
// parallelism where the sample-by-sample order is one dependent chain of 4 K operations per sample.  Same operations,
// same results.  Measured (scripts/ubench/smsp_map.cu, profiles/r02g_smsp_map.txt): with the three band warps per
// SM sub-partition that the 11/12-warp CTAs have, 113 instead of 142 clocks per sample, i.e. the FP64 issue limit.
// `out(m, y)` receives the finished sample m.
#ifndef PFB_WAVE
#define PFB_WAVE 1
#endif
#ifndef PFB_W_ROLLED
#define PFB_W_ROLLED 0
#endif
#ifndef PFB_PRODUCER_BACKOFF
#define PFB_PRODUCER_BACKOFF 0     // nanoseconds between the producer lane's polls of an `empty` barrier (experiment)
#endif
#ifndef PFB_ONE_COMMIT
#define PFB_ONE_COMMIT -1    // output pass, bulk groups: 1 = one per tile, 0 = one per 128-byte half row, -1 = per kernel (below)
#endif
template <typename T, typename A, int K, bool SCALED, int NS, typename Out>
__device__ __forceinline__ void cascade_wave(const T *e, const A (&c)[K][5], A (&w)[K][2], Out out) {
#if PFB_WAVE
    A v[K + 1];
#pragma unroll
    for (int d = 0; d < NS + K - 1; ++d) {
#pragma unroll
        for (int k = K - 1; k >= 0; --k) {      // descending: section k + 1 has consumed v[k + 1] before k rewrites it
            const int m = d - k;
            if (m >= 0 && m < NS) {
                const A xin = k == 0 ? (A)e[m] : v[k];
                const A y = section_step<A, SCALED>(xin, c[k], w[k][0], w[k][1]);
                if (k == K - 1) out(m, SCALED ? y * c[0][0] : y);
                else v[k + 1] = y;
            }
        }
    }
#else
#pragma unroll
    for (int m = 0; m < NS; ++m) {
        A s = (A)e[m];
#pragma unroll
        for (int k = 0; k < K; ++k) s = section_step<A, SCALED>(s, c[k], w[k][0], w[k][1]);
        out(m, SCALED ? s * c[0][0] : s);
    }
#endif
}

// Filter one 128-byte half tile: x row read from the shared input tile.  MODE 1: result written to the warp's
// own output tile (same swizzled layout); MODE 0: pre-pass, nothing is written; MODE 2: the squares of the
// (output-type rounded) results are added to `acc` (band-level epilogue).
template <typename T, typename A, int K, int MODE, bool SCALED>
__device__ __forceinline__ void filter_half(const char *xrow, char *yrow, int key, const A (&c)[K][5], A (&w)[K][2],
                                            A *acc = nullptr, int okey = -1) {
    if (okey < 0) okey = key;      // swizzle key of the output row (0: linear row of the bulk-row layout)
    constexpr bool STORE = MODE == 1;
    using V = typename Vec16<T>::type;
    constexpr int EPV = Vec16<T>::N;
    // float64 arithmetic on float32 rows: two wavefronts of 16 samples (a 32-sample one holds too many doubles)
    constexpr int kBatch = (sizeof(T) == 4 && sizeof(A) == 8) ? kVecPerRow / 2 : kVecPerRow;
#pragma unroll
    for (int v0 = 0; v0 < kVecPerRow; v0 += kBatch) {
        V in[kBatch];
#pragma unroll
        for (int v = 0; v < kBatch; ++v) in[v] = *reinterpret_cast<const V *>(xrow + (((v0 + v) ^ key) << 4));
        T *e = reinterpret_cast<T *>(&in[0]);
        cascade_wave<T, A, K, SCALED, kBatch * EPV>(e, c, w, [&](int m, A y) {
            e[m] = (T)y;
            if (MODE == 2) {
                const A r = (A)e[m];
                *acc = fma(r, r, *acc);
            }
        });
        if (STORE) {
#pragma unroll
            for (int v = 0; v < kBatch; ++v) *reinterpret_cast<V *>(yrow + (((v0 + v) ^ okey) << 4)) = in[v];
        }
    }
}

// Decimating variant of filter_half (MODE 1): every sample is filtered, the samples of parity `phase` of this
// 128-byte half row (16 doubles / 32 floats -> 64 bytes) are written to 16-byte chunks c0 .. c0 + 3 of the
// output row (same swizzle).  Tile steps and half rows hold an even number of samples, so the parity of a sample
// within its vector is its parity within the chunk.
template <typename T, typename A, int K, bool SCALED>
__device__ __forceinline__ void filter_half_dec(const char *xrow, char *yrow, int c0, int key, int phase, const A (&c)[K][5],
                                                A (&w)[K][2]) {
    using V = typename Vec16<T>::type;
    constexpr int EPV = Vec16<T>::N;
    V in[kVecPerRow];
#pragma unroll
    for (int v = 0; v < kVecPerRow; ++v) in[v] = *reinterpret_cast<const V *>(xrow + ((v ^ key) << 4));
    T *e = reinterpret_cast<T *>(&in[0]);
    cascade_wave<T, A, K, SCALED, kVecPerRow * EPV>(e, c, w, [&](int m, A y) { e[m] = (T)y; });
#pragma unroll
    for (int v = 0; v < kVecPerRow; v += 2) {
        V o;
        T *oe = reinterpret_cast<T *>(&o);
        const T *e0 = reinterpret_cast<const T *>(&in[v]), *e1 = reinterpret_cast<const T *>(&in[v + 1]);
#pragma unroll
        for (int i = 0; i < EPV / 2; ++i) {
            oe[i] = phase ? e0[2 * i + 1] : e0[2 * i];
            oe[EPV / 2 + i] = phase ? e1[2 * i + 1] : e1[2 * i];
        }
        *reinterpret_cast<V *>(yrow + (((c0 + v / 2) ^ key) << 4)) = o;
    }
}

// Weighting warp: the lane's 128-byte half row of the input tile through the transfer-function recurrence, in
// place (float32 tiles: float64 arithmetic, the result rounded to float32 like the reference's float path would).
template <typename T, int NZ>
__device__ __forceinline__ void weight_half(char *row, int key, const double (&b)[8], const double (&a)[8], double (&z)[NZ]) {
    using V = typename Vec16<T>::type;
    constexpr int EPV = Vec16<T>::N;
#if PFB_W_ROLLED
    // Rolled loop over groups of four samples (100 FP64 instructions, 1.6 KB of code), the next group loaded ahead:
    // an experiment against instruction-fetch stalls of the lone weighting warp; measured slower than the unrolled
    // half row (19.9 against 17.6 ms on cfg 5, profiles/r02l_warp_prof.txt).
    constexpr int VG = 4 / EPV;                  // 16-byte vectors per group: 2 (float64), 1 (float32)
    V cur[VG], nxt[VG];
#pragma unroll
    for (int v = 0; v < VG; ++v) cur[v] = *reinterpret_cast<const V *>(row + ((v ^ key) << 4));
#pragma unroll 1
    for (int i = 0; i < kVecPerRow; i += VG) {
        const int in = i + VG < kVecPerRow ? i + VG : i;
#pragma unroll
        for (int v = 0; v < VG; ++v) nxt[v] = *reinterpret_cast<const V *>(row + (((in + v) ^ key) << 4));
#pragma unroll
        for (int v = 0; v < VG; ++v) {
            T *e = reinterpret_cast<T *>(&cur[v]);
#pragma unroll
            for (int j = 0; j < EPV; ++j) e[j] = (T)tf_step<NZ>((double)e[j], b, a, z);
        }
#pragma unroll
        for (int v = 0; v < VG; ++v) {
            *reinterpret_cast<V *>(row + (((i + v) ^ key) << 4)) = cur[v];
            cur[v] = nxt[v];
        }
    }
#else
    V v[kVecPerRow];
#pragma unroll
    for (int i = 0; i < kVecPerRow; ++i) v[i] = *reinterpret_cast<const V *>(row + ((i ^ key) << 4));
#pragma unroll
    for (int i = 0; i < kVecPerRow; ++i) {
        T *e = reinterpret_cast<T *>(&v[i]);
#pragma unroll
        for (int j = 0; j < EPV; ++j) e[j] = (T)tf_step<NZ>((double)e[j], b, a, z);
    }
#pragma unroll
    for (int i = 0; i < kVecPerRow; ++i) *reinterpret_cast<V *>(row + ((i ^ key) << 4)) = v[i];
#endif
}

// The weighting warp of a WEIGHT kernel: lane <-> tile row <-> channel (one chunk per cascade).  Waits for the
// producer's tile (full), filters it in place, hands it to the band warps (wfull).
template <typename T, int NZ, int NH, int NS>
__device__ __forceinline__ void weight_warp(const TmaParams &p, char *xs, uint64_t *full, uint64_t *wfull, long long nt,
                                            int lane, int chn, bool write_back) {
    constexpr unsigned kStageBytes = NH * kTileBytes;
    const int key = lane & 7;
    const bool ok = chn < p.C;
    double b[8], a[8], z[NZ];
#pragma unroll
    for (int k = 0; k < 8; ++k) { b[k] = p.wb[k]; a[k] = p.wa[k]; }
#pragma unroll
    for (int k = 0; k < NZ; ++k) z[k] = ok ? p.wz_in[(size_t)chn * 8 + k] : 0.0;
    PROF_DECL;
    for (long long t = 0; t < nt; ++t) {
        const int s = (int)(t % NS);
        mbar_wait_warp(&full[s], (unsigned)((t / NS) & 1), lane);
        PROF_LAP(0);
        char *xrow = xs + s * kStageBytes + lane * kRowBytes;
#pragma unroll 1
        for (int h = 0; h < NH; ++h) weight_half<T, NZ>(xrow + h * kTileBytes, key, b, a, z);
        PROF_LAP(1);
#if !PFB_W_NOFENCE
        fence_proxy_async();      // the stage is refilled through the async proxy after the band warps release it
#endif
        __syncwarp();
        PROF_LAP(4);
        if (lane == 0) mbar_arrive(&wfull[s]);
        PROF_LAP(5);
    }
    PROF_END(p, 15, lane);
    PROF_END(p, 14, lane - 31);      // lane 31's clocks: a diverged warp shows up as different laps
    if (ok && write_back) {
#pragma unroll
        for (int k = 0; k < NZ; ++k) p.wz[(size_t)chn * NZ + k] = z[k];
    }
}

// cta = ((cb * G) + g) * NW + w : channel block cb, band group g, chunk group w.
// Tile row r <-> chunk JR * w + r % JR of channel CH * cb + r / JR.
// PASS 0: pre-pass, writes zero-state chunk end states to zs; 1: output pass from zs; 2: like 1 but the band
// signals are reduced to energies per time unit instead of being stored (band-level epilogue).
// NH: 128-byte half rows per tile step (2: 256-byte row fragments; 4: 512-byte fragments, which HBM writes ~8 %
// faster but which need 16 KB output tiles per band warp); NS: input ring stages.
// WEIGHT: one more warp (index nbw + 1) runs the weighting prefilter on every input tile (weight_warp) and the
// band warps wait for ITS barrier; DEC: the band p.dec_band1 - 1 stores decimated samples through ymap2.
// YB: output tiles per band warp (2: a tile's stores drain while the next tile is filtered into the other buffer --
// for kernels with one CTA per SM, where nothing else hides the store latency: the weighted kernel.  The same for the
// plain J = 1 output pass was slower than two resident CTAs, profiles/r02c_cfg5_sweep.jsonl, and was removed).
template <typename T, typename A, int K, int PASS, bool SCALED, int NH = 2, int NS = kTmaStages, bool WEIGHT = false,
          bool DEC = false, int YB = 1>
__device__ __forceinline__ void tma_cta(const TmaParams &p, const CoefTable<A> &tab, const CUtensorMap &xmap,
                                        const CUtensorMap &ymap, const CUtensorMap &ymap2, char *smem, int cta, int ch0) {
    constexpr bool PASS_B = PASS != 0;
    constexpr bool STORE_Y = PASS == 1;
    constexpr int TLEN = kRowBytes / (int)sizeof(T);      // samples per 128-byte half row
    constexpr int TT = NH * TLEN;                         // samples per tile step (NH * 128-byte rows)
    constexpr int S = 2 * K;
    constexpr unsigned kStageBytes = NH * kTileBytes;
    // Bulk groups of the output stores.  One group per 128-byte half row lets half h of the next tile start as soon as
    // half h of this one has been read out; one group per tile saves three commits.  Measured (profiles/
    // r02v_one_commit.txt): the wide float64 output pass gains 2.4 % with one group per tile (10.18 -> 9.94 ms), the
    // float32 one loses 6 %, the 256-byte-fragment kernels are indifferent.
    constexpr bool kOneCommit = PFB_ONE_COMMIT >= 0 ? PFB_ONE_COMMIT != 0 : (NH == 4 && sizeof(T) == 8);
    const int lane = threadIdx.x & 31, key = lane & 7;
    // warp index through a warp reduction: the result lives in a uniform register, so everything
    // indexed by it (the band's coefficients) can stay in uniform registers -- DFMA R, R, UR, R reads
    // two register pairs instead of three and issues at full FP64 rate
    const int warp = __reduce_max_sync(0xffffffffu, (int)(threadIdx.x >> 5));
    const int w = cta % p.NW;
    const int cg = cta / p.NW;
    const int nbw = PASS_B ? p.nb : p.nb_a, G = PASS_B ? p.G : p.G_a;
    const int g = cg % G, cb = ch0 + cg / G;
    // Band -> (CTA, warp slot).  Contiguous groups (slot s of group g filters band g * nbw + s) everywhere but in the
    // output pass with a per-band choice of arithmetic (kMixed with p.f32mask): there group g takes the bands g, g + G,
    // g + 2 G, .. so that every CTA holds its share of the float64 bands (the low ones) -- the FP64 pipe of each SM is
    // then loaded evenly instead of one CTA in G running all-float64.  The chunk states in zs are indexed by band, so
    // the pre-pass keeps its own (contiguous, heavy-first) grouping.
    constexpr bool kMixed = STORE_Y && !WEIGHT && !DEC && sizeof(T) == 4 && sizeof(A) == 8;
    const int bstep = kMixed && p.f32mask != nullptr ? G : 1;
    const int b0 = bstep > 1 ? g : g * nbw;
    const int nb_here = bstep > 1 ? (p.B - g + G - 1) / G : min(nbw, p.B - b0);

    char *xs = smem;                                          // [stages][2][tile]
    // Output pass: the lanes copy their own row fragments (NH * 128 contiguous bytes each) to global memory with
    // one-dimensional bulk copies instead of lane 0 issuing NH tensor stores of [32 rows x 128 B].  Measured on the
    // store pattern alone (scripts/ubench/tma_store_rows.cu, profiles/r02z_ubench_store_rows.txt): 6.5-6.7 TB/s for
    // 256- and 512-byte fragments against 4.8-5.3 / 5.7 TB/s with tensor stores (contiguous writes: 6.8 TB/s), and the
    // issue is one instruction of the whole warp instead of 150-400 clocks of lane 0 per store.
    // In the kernels the copy instruction of a warp carries 32 requests and occupies the warp for 1500-2200 clocks per
    // tile (profiles/r03b_warp_prof.txt) where two tensor stores cost 270: a gain only where a tile is long in samples
    // and the pass is bound by the stores -- the wide float64 kernel (cfg 2 output pass 10.02 -> 9.67 ms).  A/B builds
    // on one box (profiles/r03d_bulk_rows.txt): the order-6 bank (11.0 -> 12.5 ms), the fused weighted bank (f32 9.3 ->
    // 12.5 ms), the float32 wide kernel (6.27 -> 6.86 ms) and float32 signals with float64 arithmetic (10.23 -> 10.54
    // ms) lose with it and keep the tensor stores.
    constexpr bool kBulkRows = STORE_Y && bulk_rows<T, A>(NH, DEC, WEIGHT);
    constexpr int kRowStride = out_row_stride(NH);
    constexpr unsigned kSlotBytes = kBulkRows ? (unsigned)out_slot_bytes(NH) : kStageBytes;
    char *ys = smem + NS * kStageBytes;               // [nb][YB][slot], output pass only
    uint64_t *full = reinterpret_cast<uint64_t *>(ys + (STORE_Y ? p.nb * YB : 0) * kSlotBytes);
    uint64_t *empty = full + NS;
    uint64_t *wfull = empty + NS;                             // WEIGHT kernels: weighting warp -> band warps
    uint64_t *in_bar = WEIGHT ? wfull : full;                 // what a band warp waits for

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], nb_here);
            if (WEIGHT) mbar_init(&wfull[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    const long long nt = p.L / TT;
    // first tile of the pre-pass: the longest pre-pass among this CTA's bands
    long long t_first = 0;
    if (!PASS_B) {
        int wmax = 0;
        for (int i = 0; i < nb_here; ++i) wmax = max(wmax, p.warm ? p.warm[b0 + i] : (int)min(p.L, 0x7fffffffLL));
        long long wt = ((long long)wmax + TT - 1) / TT;
        if (wt > nt) wt = nt;
        t_first = nt - wt;
    }

    // Warp roles.  Plain kernels: bands 0 .. nbw - 1, then the producer.
    // WEIGHT kernels spread the roles over the SM's four sub-partitions (a warp issues on sub-partition
    // warp index mod 4, scripts/ubench/smsp_map.cu): warp 3 is the weighting warp, warp 7 the producer, warp 11 idles,
    // the band warps are the others (0 1 2, 4 5 6, 8 9 10 -> at most 9 bands).  The prefilter is one sequential FP64
    // instruction stream per CTA (25 separately rounded operations per sample, 56 issue clocks alone): sharing its
    // sub-partition with band warps (and their polling lanes) made it the slowest stage of the pipeline by 2-3x
    // (profiles/r02j_warp_prof.txt: 184 clocks per sample against 77 in isolation).
    constexpr int kWeightWarp = 3, kSpreadProducer = 7;
    const int prod_warp = WEIGHT ? kSpreadProducer : nbw;
    if (WEIGHT && (warp & 3) == 3 && warp != kWeightWarp && warp != kSpreadProducer) return;
    const int slot = WEIGHT ? warp - (warp >> 2) : warp;            // band slot of a band warp
    if (warp == prod_warp) {
        // ---------------- producer: stream the channel's input tiles into the ring ----------------
        if (lane == 0) {
            for (long long t = t_first; t < nt; ++t) {
                const long long it = t - t_first;
                const int s = (int)(it % NS);
                const unsigned ph = (unsigned)((it / NS) & 1);
                mbar_wait(&empty[s], ph ^ 1u);
                mbar_expect_tx(&full[s], kStageBytes);
                char *dst = xs + s * kStageBytes;
#pragma unroll
                for (int h = 0; h < NH; ++h)
                    tma_load_3d(dst + h * kTileBytes, &xmap, &full[s], (int)(t * TT + h * TLEN), p.JR * w, p.CH * cb);
            }
        }
        return;
    }
    if (WEIGHT && warp == kWeightWarp) {
        // ---------------- weighting warp (PASS_B only, JR == 1: lane <-> channel CH * cb + lane) ----------------
        const int chn = p.CH * cb + lane;
        switch (p.wnz) {
            case 4: weight_warp<T, 4, NH, NS>(p, xs, full, wfull, nt, lane, chn, g == 0); break;
            case 5: weight_warp<T, 5, NH, NS>(p, xs, full, wfull, nt, lane, chn, g == 0); break;
            default: weight_warp<T, 6, NH, NS>(p, xs, full, wfull, nt, lane, chn, g == 0); break;
        }
        return;
    }
    if (slot >= nb_here) return;

    // ---------------- consumers: one band per warp, one (channel, chunk) row per lane ----------------
    const int band = b0 + slot * bstep;
    const int chn = p.CH * cb + lane / p.JR;
    const bool row_ok = chn < p.C;                // channel blocks may be ragged: TMA clips the tile, lanes idle
    const int q = (row_ok ? chn : 0) * p.B + band;
    const bool is_dec = DEC && p.dec_band1 != 0 && band == p.dec_band1 - 1;   // warp-uniform
    const int j = p.JR * w + lane % p.JR;
    A c[K][5], wst[K][2];
    // a second, independently reduced copy of the band index that is used for nothing but the
    // coefficient table, so its users stay on the uniform datapath
    const int warp_u = __reduce_min_sync(0xffffffffu, (int)(threadIdx.x >> 5));
    const int band_u = b0 + (WEIGHT ? warp_u - (warp_u >> 2) : warp_u) * bstep;
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int i = 0; i < 5; ++i) c[k][i] = tab.c[band_u * K + k][i];
    A *zs = reinterpret_cast<A *>(p.zs) + ((size_t)q * p.J + j) * S;
    const double *gain = SCALED ? p.gains + (size_t)band * K * 2 : nullptr;
    // per-band choice of arithmetic (float32 signals with float64 states, output pass): warp-uniform
    bool use_f32 = false;
    float cf[K][5], wf[K][2];
    if (kMixed && p.f32mask != nullptr) use_f32 = ((p.f32mask[band >> 5] >> (band & 31)) & 1u) != 0;
    if (kMixed && use_f32) {
        const A *z0 = p.J == 1 ? cascade_states<A>(p, row_ok ? chn : 0, band) : zs;
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
            for (int i = 0; i < 5; ++i) cf[k][i] = __ldg(p.cf32 + ((size_t)band * K + k) * 5 + i);
            wf[k][0] = (float)z0[2 * k];
            wf[k][1] = (float)z0[2 * k + 1];
        }
    }
    if (PASS_B) {
        // a single chunk per cascade starts from the carried-in states themselves (no pre-pass, no carry kernel)
        const A *z0 = p.J == 1 ? cascade_states<A>(p, row_ok ? chn : 0, band) : zs;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            wst[k][0] = z0[2 * k];
            wst[k][1] = z0[2 * k + 1];
            if (SCALED) { wst[k][0] *= (A)gain[2 * k + 1]; wst[k][1] *= (A)gain[2 * k + 1]; }
        }
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) wst[k][0] = wst[k][1] = A(0);
    }
    int t_mine = 0;
    if (!PASS_B) {
        // (reduced across the warp only so that the compiler sees a warp-uniform loop bound)
        const int warm = __reduce_max_sync(0xffffffffu, p.warm ? p.warm[band] : 0x7fffffff);
        long long wt = ((long long)warm + TT - 1) / TT;
        if (wt > nt) wt = nt;
        t_mine = (int)(nt - wt);
    }
    char *ybuf = ys + slot * YB * kSlotBytes;
    char *yrow = ybuf + lane * kRowBytes;               // swizzled tensor-store tile (decimating band)

    if (PASS == 2) {
        double acc_unit = 0.0;
        int tcount = 0;
        double *lev = p.levels + (size_t)q * p.lev_ld + p.lev_off + (long long)j * (nt / p.ut);
        for (long long t = 0; t < nt; ++t) {
            const int s = (int)(t % NS);
            const unsigned ph = (unsigned)((t / NS) & 1);
            mbar_wait_warp(&in_bar[s], ph, lane);
            const char *xrow = xs + s * kStageBytes + lane * kRowBytes;
            A acc = A(0);
#pragma unroll 1   // rolled like the output pass (float32 levels: 13.1 ms unrolled, see profiles/r01h_wait_variants.txt)
            for (int h = 0; h < NH; ++h) filter_half<T, A, K, 2, SCALED>(xrow + h * kTileBytes, nullptr, key, c, wst, &acc);
            __syncwarp();
            stage_release_after_reads(&empty[s], lane);
            acc_unit += (double)acc;
            if (++tcount == p.ut) {
                if (row_ok) *lev = acc_unit;
                ++lev;
                acc_unit = 0.0;
                tcount = 0;
            }
        }
    } else if (DEC && is_dec) {
        // decimating band: a tile step of NH half rows yields NH / 2 half rows of kept samples (one bulk group)
        static_assert(!DEC || NH % 2 == 0, "decimated half rows pair up");
        for (long long t = 0; t < nt; ++t) {
            const int s = (int)(t % NS);
            const unsigned ph = (unsigned)((t / NS) & 1);
            mbar_wait_warp(&in_bar[s], ph, lane);
            const char *xrow = xs + s * kStageBytes + lane * kRowBytes;
            const int yb = YB > 1 ? (int)(t % YB) * (int)kSlotBytes : 0;
            if (lane == 0) bulk_wait_read_upto(YB - 1);          // one bulk group per tile
            __syncwarp();
#pragma unroll 1
            for (int h = 0; h < NH; ++h)
                filter_half_dec<T, A, K, SCALED>(xrow + h * kTileBytes, yrow + yb + (h >> 1) * kTileBytes, (h & 1) * 4, key,
                                                 p.dec_phase, c, wst);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&empty[s]);
#pragma unroll
                for (int o = 0; o < NH / 2; ++o)
                    tma_store_4d(&ymap2, ybuf + yb + o * kTileBytes, (int)(t * (TT / 2) + o * TLEN), p.JR * w, 0, p.CH * cb);
                bulk_commit();
            }
        }
    } else if (PASS_B && kBulkRows) {
        PROF_DECL;
        // this lane's output row in global memory and in the warp's output slot
        const int By = p.By ? p.By : p.B;
        char *grow = reinterpret_cast<char *>(p.y) +
                     (((size_t)(row_ok ? chn : 0) * By + band) * (size_t)p.ldy + (size_t)j * (size_t)p.L) * sizeof(T);
        char *orow = ybuf + lane * kRowStride;
        for (long long t = 0; t < nt; ++t) {
            const int s = (int)(t % NS);
            const unsigned ph = (unsigned)((t / NS) & 1);
            mbar_wait_warp(&in_bar[s], ph, lane);
            PROF_LAP(0);
            const char *xrow = xs + s * kStageBytes + lane * kRowBytes;
            const int yb = YB > 1 ? (int)(t % YB) * (int)kSlotBytes : 0;
            // the row is overwritten once this lane's copy of tile t - YB has been read out of shared memory
            bulk_wait_read_upto(YB - 1);
            PROF_LAP(2);
            // The half rows stay a loop: fully unrolled, the wide float32 tile is ~24 KB of code per warp iteration and
            // stalls on instruction fetch.
#pragma unroll 1
            for (int h = 0; h < NH; ++h)
                filter_half<T, A, K, 1, SCALED>(xrow + h * kTileBytes, orow + yb + h * kRowBytes, key, c, wst, nullptr, 0);
            PROF_LAP(1);
            fence_proxy_async();      // this lane's row -> async proxy; also orders the tile's loads before the release
            if (row_ok) bulk_store_1d(grow + (size_t)t * (NH * kRowBytes), orow + yb, NH * kRowBytes);
            bulk_commit();
            PROF_LAP(5);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            PROF_LAP(4);
        }
        PROF_END(p, slot, lane);
    } else if (PASS_B) {
        PROF_DECL;
        for (long long t = 0; t < nt; ++t) {
            const int s = (int)(t % NS);
            const unsigned ph = (unsigned)((t / NS) & 1);
            mbar_wait_warp(&in_bar[s], ph, lane);
            PROF_LAP(0);
            const char *xrow = xs + s * kStageBytes + lane * kRowBytes;
            // The tile's NH half rows are stored together at the end of the tile: a row's fragment must reach HBM as
            // one burst (storing each 128-byte half as soon as it is filtered dropped the write efficiency from ~5.0
            // to ~4.4 TB/s, profiles/r01_ubench_tma_store_rows.txt).  Each half is its own bulk group, so half h of
            // the next tile only waits until half h of this one has been read out of shared memory and filtering
            // overlaps the draining of the other halves.
            // The half rows stay a loop.  Fully unrolled, the wide float32 tile is ~24 KB of code per warp iteration and
            // stalls on instruction fetch; and with the warp-level waits above (real WARPSYNCs in the loop) every
            // fully unrolled variant lost 40-100 % (float64 cfg 2 output pass 15.8 vs 10.0 ms, cfg 3 21.9 vs 11.2 ms,
            // profiles/r01h_wait_variants.txt), while the rolled loop costs nothing measurable.
            const int yb = YB > 1 ? (int)(t % YB) * (int)kSlotBytes : 0;
#pragma unroll 1
            for (int h = 0; h < NH; ++h) {
                // (YB buffers: the groups of the YB - 1 tiles in between may still be pending)
                if (kOneCommit) {
                    if (h == 0) {     // one bulk group per tile: the whole previous tile must have been read
                        if (lane == 0) bulk_wait_read_upto(YB - 1);
                        __syncwarp();
                    }
                } else {
                    if (lane == 0) bulk_wait_read_upto((YB - 1) * NH + NH - 1 - h);
                    __syncwarp();
                }
                PROF_LAP(2);
                if constexpr (kMixed) {
                    if (use_f32) filter_half<float, float, K, 1, false>(xrow + h * kTileBytes, yrow + yb + h * kTileBytes, key, cf, wf);
                    else filter_half<T, A, K, 1, SCALED>(xrow + h * kTileBytes, yrow + yb + h * kTileBytes, key, c, wst);
                } else {
                    filter_half<T, A, K, 1, SCALED>(xrow + h * kTileBytes, yrow + yb + h * kTileBytes, key, c, wst);
                }
                PROF_LAP(1);
            }
            fence_proxy_async();
            __syncwarp();
            PROF_LAP(4);
            if (lane == 0) {
                mbar_arrive(&empty[s]);
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    tma_store_4d(&ymap, ybuf + yb + h * kTileBytes, (int)(t * TT + h * TLEN), p.JR * w, band, p.CH * cb);
                    if (!kOneCommit) bulk_commit();
                }
                if (kOneCommit) bulk_commit();
            }
            PROF_LAP(5);
        }
        PROF_END(p, slot, lane);
    } else {
        // tiles before this band's own pre-pass start are only released, not filtered
        long long t = t_first;
        for (; t < t_mine; ++t) {
            const long long it = t - t_first;
            const int s = (int)(it % NS);
            mbar_skip_tile(&full[s], &empty[s], (unsigned)((it / NS) & 1), lane);
        }
        __syncwarp();
        for (; t < nt; ++t) {
            const long long it = t - t_first;
            const int s = (int)(it % NS);
            mbar_wait_warp(&full[s], (unsigned)((it / NS) & 1), lane);
            const char *xrow = xs + s * kStageBytes + lane * kRowBytes;
#pragma unroll 1
            for (int h = 0; h < NH; ++h) filter_half<T, A, K, 0, SCALED>(xrow + h * kTileBytes, nullptr, key, c, wst);
            __syncwarp();
            stage_release_after_reads(&empty[s], lane);
        }
    }
    if (PASS_B) {
        if (STORE_Y && (kBulkRows || lane == 0)) bulk_wait0();     // outstanding stores must complete before the CTA's smem goes away
        if (row_ok && j == p.J - 1) {    // the last chunk ends with the cascade's final state
            A *st = cascade_states<A>(p, chn, band);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const A gk = SCALED ? (A)gain[2 * k] : A(1);
                st[2 * k] = kMixed && use_f32 ? (A)wf[k][0] : wst[k][0] * gk;
                st[2 * k + 1] = kMixed && use_f32 ? (A)wf[k][1] : wst[k][1] * gk;
            }
        }
    } else if (row_ok) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const A gk = SCALED ? (A)gain[2 * k] : A(1);
            zs[2 * k] = wst[k][0] * gk;
            zs[2 * k + 1] = wst[k][1] * gk;
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Pre-pass in chunk-impulse-response form on the FP64 tensor-core path (DMMA m8n8k4).
// The zero-state end state of a chunk is linear in its samples:  z[i] = sum_t Hrev[i][t] x[t]  with
// Hrev[i][t] the response of state i at the chunk end to a unit sample at t (built on the host in long
// double).  Per band warp that is a GEMM  Z[2K x 32 chunks] = Hrev[2K x W] X[W x 32 chunks]:
// A fragments (Hrev, 8 states x 4 samples) stream from global memory / L2 and are shared by the four
// n-tiles (8 chunks each) of the warp's 32 chunks, B fragments (x) come from the shared input tile.
// 8 FMA-equivalents per chunk-sample instead of the recurrence's 14; DMMA and DFMA share one pipe on
// B200 (profiles/r01_ubench_dmma_vs_dfma.txt), so the gain is the operation count, not concurrency.
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// T: signal type, Z: type of the chunk states (the arithmetic type of the other two kernels).  The GEMM itself
// always runs in float64 -- for float32 banks that keeps the FP32 pipe free for the output pass.
template <typename T, typename Z, int K, int NBMAX>
__global__ void __launch_bounds__((NBMAX + 1) * 32, NBMAX <= 7 ? 3 : 2)   // ~80 registers: accumulators + two A-fragment sets
sos_hform_kernel(const __grid_constant__ TmaParams p, const __grid_constant__ CUtensorMap xmap) {
    extern __shared__ __align__(1024) char smem[];
    constexpr int TLEN = kRowBytes / (int)sizeof(T);
    constexpr int TT = 2 * TLEN;
    constexpr int S = 2 * K;
    constexpr int MT = (S + 7) / 8;                          // 8-state tiles
    constexpr int KS = TT / 4;                               // 4-sample steps per tile
    constexpr unsigned kStageBytes = 2 * kTileBytes;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // Band groups outermost, lowest bands first: their windows are the longest (a whole chunk), so the heavy CTAs
    // start first and the short high-band CTAs fill the tail of the grid.
    const int cta = blockIdx.x;
    const int w = cta % p.NW;
    const int cg = cta / p.NW;
    const int nbw = p.nb_a, G = p.G_a;
    const int cb = cg % p.CB, g = cg / p.CB;
    const int b0 = g * nbw;
    const int nb_here = min(nbw, p.B - b0);

    char *xs = smem;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + kTmaStages * kStageBytes);
    uint64_t *empty = full + kTmaStages;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kTmaStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], nb_here);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    const long long nt = p.L / TT;
    int wmax = 0;
    for (int i = 0; i < nb_here; ++i) wmax = max(wmax, p.warm[b0 + i]);
    long long wt_max = ((long long)wmax + TT - 1) / TT;
    if (wt_max > nt) wt_max = nt;
    const long long t_first = nt - wt_max;

    if (warp == nbw) {
        if (lane == 0) {
            for (long long t = t_first; t < nt; ++t) {
                const long long it = t - t_first;
                const int s = (int)(it % kTmaStages);
                const unsigned ph = (unsigned)((it / kTmaStages) & 1);
                mbar_wait(&empty[s], ph ^ 1u);
                mbar_expect_tx(&full[s], kStageBytes);
                char *dst = xs + s * kStageBytes;
                tma_load_3d(dst, &xmap, &full[s], (int)(t * TT), p.JR * w, p.CH * cb);
                tma_load_3d(dst + kTileBytes, &xmap, &full[s], (int)(t * TT + TLEN), p.JR * w, p.CH * cb);
            }
        }
        return;
    }
    if (warp >= nb_here) return;

    const int band = b0 + warp;
    long long wt = ((long long)p.warm[band] + TT - 1) / TT;
    if (wt > nt) wt = nt;
    const long long t_mine = nt - wt;
    // this lane's A elements: step ks of tile tl (counted from t_mine) -> Hb[((tl * KS + ks) * MT + mt) * 32 + lane]
    const double *Hb = p.H + p.Hoff[band] + lane;
    // this lane's B elements: row 8 n + lane / 4 (n-tile n), sample 4 ks + lane % 4 of the tile
    const int key = lane >> 2;
    const char *xlane = xs + key * kRowBytes;
    double acc[MT][4][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int n = 0; n < 4; ++n) acc[mt][n][0] = acc[mt][n][1] = 0.0;

    long long t = t_first;
    for (; t < t_mine; ++t) {            // tiles before this band's window are only released
        const long long it = t - t_first;
        const int s = (int)(it % kTmaStages);
        mbar_skip_tile(&full[s], &empty[s], (unsigned)((it / kTmaStages) & 1), lane);
    }
    __syncwarp();
    double a_cur[KS][MT];
    if (t < nt) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) a_cur[ks][mt] = __ldg(Hb + (size_t)(ks * MT + mt) * 32);
    }
    for (; t < nt; ++t) {
        const long long it = t - t_first;
        const int s = (int)(it % kTmaStages);
        // next tile's A fragments are requested before this tile's input is waited for
        double a_nxt[KS][MT];
        const bool more = t + 1 < nt;
        const double *Hn = Hb + (size_t)(t + 1 - t_mine) * KS * MT * 32;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) a_nxt[ks][mt] = more ? __ldg(Hn + (size_t)(ks * MT + mt) * 32) : 0.0;
        mbar_wait_warp(&full[s], (unsigned)((it / kTmaStages) & 1), lane);
        const char *xt = xlane + s * kStageBytes;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const int smp = (4 * ks) % TLEN;                          // first sample of the step within its half tile
            const char *xh = xt + ((4 * ks) / TLEN) * kTileBytes;
            const int byte = (smp + (lane & 3)) * (int)sizeof(T);
            const int off = (((byte >> 4) ^ key) << 4) + (byte & 15);
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                const double bv = (double)*reinterpret_cast<const T *>(xh + n * 8 * kRowBytes + off);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) dmma884(acc[mt][n][0], acc[mt][n][1], a_cur[ks][mt], bv);
            }
        }
        __syncwarp();
        stage_release_after_reads(&empty[s], lane);
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) a_cur[ks][mt] = a_nxt[ks][mt];
    }
    // C fragment: lane holds state 8 mt + lane / 4 of tile rows 8 n + 2 (lane % 4) + {0, 1}
    Z *zs = reinterpret_cast<Z *>(p.zs);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int i = 8 * mt + (lane >> 2);
        if (i < S) {
#pragma unroll
            for (int n = 0; n < 4; ++n) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int r = 8 * n + 2 * (lane & 3) + e;
                    const int chn = p.CH * cb + r / p.JR, j = p.JR * w + r % p.JR;
                    if (chn < p.C) zs[(((size_t)chn * p.B + band) * p.J + j) * S + i] = (Z)acc[mt][n][e];
                }
            }
        }
    }
}

// The pre-pass needs no output tiles and few registers: it is compiled for 4 resident CTAs per SM
// (which also makes ptxas keep the coefficients in uniform registers), the output pass for 2.
template <typename T, typename A, int K, int PASS, int NBMAX, bool SCALED, bool WEIGHT = false, bool DEC = false>
__global__ void __launch_bounds__((NBMAX + 1 + (WEIGHT ? 1 : 0)) * 32, PASS == 1 ? 2 : PASS == 2 ? 3 : 4)
sos_tma_kernel(const __grid_constant__ TmaParams p, const __grid_constant__ CoefTable<A> tab,
               const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap ymap,
               const __grid_constant__ CUtensorMap ymap2) {
    extern __shared__ __align__(1024) char smem[];
    tma_cta<T, A, K, PASS, SCALED, 2, kTmaStages, WEIGHT, DEC>(p, tab, xmap, ymap, ymap2, smem, blockIdx.x, 0);
}

// Fused weighting prefilter + bank (output pass or band-level epilogue): 12 warps in the spread role layout of
// tma_cta (<= 9 band warps), one CTA per SM, four input stages and two output tiles per band warp -- with a single
// resident CTA nothing else hides the latency of the TMA stores (1500-2200 clocks until a tile has been read out of
// shared memory, profiles/r02j_warp_prof.txt).
constexpr int kTmaWeightBands = 9;
constexpr int kTmaWeightStages = 4;
template <typename T, typename A, int K, int PASS, bool SCALED>
__global__ void __launch_bounds__(12 * 32, 1)
sos_tma_weighted_kernel(const __grid_constant__ TmaParams p, const __grid_constant__ CoefTable<A> tab,
                        const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap ymap) {
    extern __shared__ __align__(1024) char smem[];
    tma_cta<T, A, K, PASS, SCALED, 2, kTmaWeightStages, true, false, PASS == 1 ? 2 : 1>(p, tab, xmap, ymap, ymap, smem, blockIdx.x, 0);
}

// Carry between the two passes: per cascade, s[0] = carried-in state, s[j+1] = P s[j] + z[j];
// zs[q][j] is overwritten with chunk j's initial state.  2K lanes per cascade (lane i owns row i of
// P and component i of the state; the state vector is exchanged with shuffles), so the J
// sequential steps cost one short dot product each and the zs accesses are contiguous.
template <typename A, int K>
__global__ void __launch_bounds__(128) sos_carry_kernel(const TmaParams p) {
    constexpr int S = 2 * K;
    constexpr int LPC = S <= 2 ? 2 : S <= 4 ? 4 : S <= 8 ? 8 : 16;   // lanes per cascade (power of two)
    constexpr int CPW = 32 / LPC;                                     // cascades per warp
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPC;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long q = warp * CPW + lane / LPC;
    const bool active = q < p.Q && sub < S;
    const long long qq = q < p.Q ? q : p.Q - 1;
    const int band = (int)(qq % p.B);
    const A *P = reinterpret_cast<const A *>(p.P) + (size_t)band * S * S;
    const A *st = cascade_states<A>(p, (int)(qq / p.B), band);
    A *zs = reinterpret_cast<A *>(p.zs) + (size_t)qq * p.J * S;
    A prow[S];
#pragma unroll
    for (int m = 0; m < S; ++m) prow[m] = sub < S ? P[sub * S + m] : A(0);
    A s = sub < S ? st[sub] : A(0);
    const int base = lane - sub;
    A znext = sub < S ? zs[sub] : A(0);
    for (int j = 0; j < p.J; ++j) {
        const A z = znext;
        if (j + 1 < p.J && sub < S) znext = zs[(size_t)(j + 1) * S + sub];   // prefetch
        if (active) zs[(size_t)j * S + sub] = s;
        A acc = z;
#pragma unroll
        for (int m = 0; m < S; ++m) acc = fma(prow[m], __shfl_sync(0xffffffffu, s, base + m), acc);
        s = acc;
    }
}

// levels[q][k] = sum of r consecutive unit energies partial[q][k r .. k r + r), fixed order (deterministic)
static __global__ void levels_reduce_kernel(const double *partial, long long pld, double *levels, long long lld, long long nblocks,
                                     int r, long long Q) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Q * nblocks) return;
    const long long q = i / nblocks, k = i - q * nblocks;
    const double *src = partial + q * pld + k * r;
    double acc = 0.0;
    for (int u = 0; u < r; ++u) acc += src[u];
    levels[q * lld + k] = acc;
}

// Tensor maps of the chunked signal views (pfb_api.cu).  esz_row: bytes per element of a row (8 / 4); `L`, `ld` in elements.
bool make_x_map(CUtensorMap *map, int dtype, const void *base, long long L, int J, long long C, long long ld, int JR, int CH);
bool make_y_map(CUtensorMap *map, int dtype, void *base, long long L, int J, int B, long long C, long long ld, int JR, int CH);

// ev: null, or 4 events recorded before the pre-pass and after each of the three kernels
// phases: 1 = pre-pass, 2 = carry + output pass, 3 = all (on stream s)
// Output pass with 512-byte row fragments: up to 11 band warps + producer, 16 KB output tile per band warp,
// two 16 KB input stages, one CTA per SM.
template <typename T, typename A, int K, bool SCALED>
__global__ void __launch_bounds__((kTmaMaxBands + 1) * 32, 1)
sos_tma_wide_kernel(const __grid_constant__ TmaParams p, const __grid_constant__ CoefTable<A> tab,
                    const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap ymap) {
    extern __shared__ __align__(1024) char smem[];
    tma_cta<T, A, K, 1, SCALED, 4, 2>(p, tab, xmap, ymap, ymap, smem, blockIdx.x, 0);
}

int launch_tma_dd(const TmaParams &p, int K, const double *coef_host, const CUtensorMap &xmap, const CUtensorMap &ymap,
                  const CUtensorMap &ymap2, cudaStream_t s, int *launches, cudaEvent_t *ev, int phases);
int launch_tma_ff(const TmaParams &p, int K, const double *coef_host, const CUtensorMap &xmap, const CUtensorMap &ymap,
                  const CUtensorMap &ymap2, cudaStream_t s, int *launches, cudaEvent_t *ev, int phases);
int launch_tma_fd(const TmaParams &p, int K, const double *coef_host, const CUtensorMap &xmap, const CUtensorMap &ymap,
                  const CUtensorMap &ymap2, cudaStream_t s, int *launches, cudaEvent_t *ev, int phases);

}  // namespace pfb
