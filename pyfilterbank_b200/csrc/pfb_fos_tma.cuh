// pfb_fos_tma.cuh -- TMA / mbarrier kernels for the gammatone complex one-pole cascade
// (GammatoneFilterbank.analyze / fosfilter, pyfilterbank/gammatone.py:196-237): the structure of
// sos_tma_kernel (pfb_tma.cuh) with a different per-sample recurrence.
//   tile row = one time chunk of one signal (JR chunks x CH signals = 32 rows); a producer warp streams
//   [32 rows x 16 real samples] input tiles; each band warp turns its lane's row into 16 complex samples
//   (256 bytes) in its own output tile and hands it to 4-D TMA stores; band coefficients (g, Re c, Im c)
//   arrive in the kernel parameters (uniform registers).
//   PASS_B = false: zero-state end states of every chunk over the band's decay window (pre-pass);
//   sos_carry_kernel<double, ORDER> composes them with the 2*ORDER x 2*ORDER chunk transition; PASS_B = true
//   filters every chunk from its start state and writes the bands.  With one chunk per signal there is no
//   pre-pass and no carry.
#pragma once
#include "pfb_tma.cuh"

namespace pfb {

constexpr int kFosStages = 4;                    // input ring: 4 x [32 rows x 128 B]
constexpr int kFosMaxBands = 7;
constexpr int kFosWideBands = 11;                // wide variant: 512-byte output fragments, one CTA per SM

struct FosTmaParams {
    void *zs;             // [Q][J][2*ORDER]
    double *states;       // [Q][ORDER][2]
    const int *warm;      // [B] pre-pass samples per chunk
    long long L;
    int Q, B, C, J, JR, CH, NW, CB, nb, G;
    double *y;            // output base and row pitch in doubles (2 per complex sample): the wide output pass copies every
    long long ldy2;       // lane's own 512-byte row fragment with a one-dimensional bulk copy (like sos_tma_wide_kernel)
};

struct FosCoef {
    double c[kCoefBytes / 24][3];   // gain, Re c, Im c per band
};

template <int ORDER>
__device__ __forceinline__ void fos_sample(double x, double g, double cr, double ci, double (&zr)[ORDER], double (&zi)[ORDER],
                                           double &yr, double &yi) {
    // scipy's direct form II transposed for b = [g], a = [1, -c]:  y = z + g x ; z = c y  (g only in stage 0)
    yr = zr[0] + g * x;
    yi = zi[0];
    zr[0] = cr * yr - ci * yi;
    zi[0] = cr * yi + ci * yr;
#pragma unroll
    for (int s = 1; s < ORDER; ++s) {
        yr = zr[s] + yr;
        yi = zi[s] + yi;
        zr[s] = cr * yr - ci * yi;
        zi[s] = cr * yi + ci * yr;
    }
}

// NX: 128-byte input half rows (16 samples) per stage -> 256 * NX bytes of complex output per row and step.
// NX = 2 (with up to 11 band warps, one CTA per SM) stores 512-byte fragments, which HBM takes ~8 % faster.
template <int ORDER, bool PASS_B, int NX = 1>
__global__ void __launch_bounds__(((NX == 2 ? kFosWideBands : kFosMaxBands) + 1) * 32, NX == 2 ? 1 : (PASS_B ? 2 : 4))
fos_tma_kernel(const __grid_constant__ FosTmaParams p, const __grid_constant__ FosCoef tab,
               const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap ymap) {
    extern __shared__ __align__(1024) char smem[];
    constexpr int TLEN = 16;
    constexpr int TT = NX * TLEN;                              // samples per stage
    // every lane copies its own row fragment (256 NX bytes of complex samples) with a one-dimensional bulk copy: see
    // tma_cta (pfb_tma.cuh); cfg 4: 18.1-18.9 -> 15.9-16.9 ms (NX = 2), 19.4-21.2 -> 17.2 ms (NX = 1), profiles/r03d_bulk_rows.txt
    constexpr bool kBulkRows = PASS_B;
    constexpr int kRowStride = out_row_stride(2 * NX);
    constexpr int kXStage = NX * kTileBytes, kYTile = kBulkRows ? out_slot_bytes(2 * NX) : 2 * NX * kTileBytes;
    constexpr int S = 2 * ORDER;
    const int lane = threadIdx.x & 31, key = lane & 7;
    const int warp = __reduce_max_sync(0xffffffffu, (int)(threadIdx.x >> 5));
    const int cta = blockIdx.x;
    const int w = cta % p.NW;
    const int cg = cta / p.NW;
    const int g = cg % p.G, cb = cg / p.G;
    const int b0 = g * p.nb;
    const int nb_here = min(p.nb, p.B - b0);

    char *xs = smem;                                           // [stages][tile]
    char *ys = smem + kFosStages * kXStage;                    // [nb][2 NX][tile], output pass only
    uint64_t *full = reinterpret_cast<uint64_t *>(ys + (PASS_B ? p.nb : 0) * kYTile);
    uint64_t *empty = full + kFosStages;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kFosStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], nb_here);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    const long long nt = p.L / TT;
    long long t_first = 0;
    if (!PASS_B) {
        int wmax = 0;
        for (int i = 0; i < nb_here; ++i) wmax = max(wmax, p.warm[b0 + i]);
        long long wt = ((long long)wmax + TT - 1) / TT;
        if (wt > nt) wt = nt;
        t_first = nt - wt;
    }
    if (warp == p.nb) {
        if (lane == 0) {
            for (long long t = t_first; t < nt; ++t) {
                const long long it = t - t_first;
                const int s = (int)(it % kFosStages);
                const unsigned ph = (unsigned)((it / kFosStages) & 1);
                mbar_wait(&empty[s], ph ^ 1u);
                mbar_expect_tx(&full[s], kXStage);
#pragma unroll
                for (int h = 0; h < NX; ++h)
                    tma_load_3d(xs + s * kXStage + h * kTileBytes, &xmap, &full[s], (int)(t * TT + h * TLEN), p.JR * w, p.CH * cb);
            }
        }
        return;
    }
    if (warp >= nb_here) return;

    const int band = b0 + warp;
    const int band_u = b0 + __reduce_min_sync(0xffffffffu, (int)(threadIdx.x >> 5));
    const double gain = tab.c[band_u][0], cr = tab.c[band_u][1], ci = tab.c[band_u][2];
    const int sig = p.CH * cb + lane / p.JR;
    const bool row_ok = sig < p.C;
    const int q = (row_ok ? sig : 0) * p.B + band;
    const int j = p.JR * w + lane % p.JR;
    double zr[ORDER], zi[ORDER];
    double *zs = reinterpret_cast<double *>(p.zs) + ((size_t)q * p.J + j) * S;
    if (PASS_B) {
        const double *z0 = p.J == 1 ? p.states + (size_t)q * S : zs;
#pragma unroll
        for (int s = 0; s < ORDER; ++s) { zr[s] = z0[2 * s]; zi[s] = z0[2 * s + 1]; }
    } else {
#pragma unroll
        for (int s = 0; s < ORDER; ++s) zr[s] = zi[s] = 0.0;
    }
    long long t_mine = 0;
    if (!PASS_B) {
        const int warm = __reduce_max_sync(0xffffffffu, p.warm[band]);
        long long wt = ((long long)warm + TT - 1) / TT;
        if (wt > nt) wt = nt;
        t_mine = nt - wt;
    }
    char *ybuf = ys + warp * kYTile;
    char *yrow = ybuf + lane * kRowBytes;
    char *orow = ybuf + lane * kRowStride;                      // kBulkRows: this lane's linear row of 2 TT complex samples
    char *grow = kBulkRows ? reinterpret_cast<char *>(p.y + ((size_t)q * (size_t)p.ldy2 + (size_t)j * (size_t)(2 * p.L)))
                           : nullptr;

    long long t = t_first;
    for (; t < t_mine; ++t) {
        const long long it = t - t_first;
        const int s = (int)(it % kFosStages);
        mbar_skip_tile(&full[s], &empty[s], (unsigned)((it / kFosStages) & 1), lane);
    }
    __syncwarp();
    for (; t < nt; ++t) {
        const long long it = t - t_first;
        const int s = (int)(it % kFosStages);
        mbar_wait_warp(&full[s], (unsigned)((it / kFosStages) & 1), lane);
        const char *xrow = xs + s * kXStage + lane * kRowBytes;
        // every 128-byte output half is its own bulk group: before half k of this tile is overwritten only the
        // store of half k of the previous tile must have been read, the later halves may still be draining
        if (kBulkRows) bulk_wait_read0();          // this lane's copy of the previous stage has left its row
#pragma unroll
        for (int h = 0; h < NX; ++h) {
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                if (PASS_B && !kBulkRows) {
                    if (lane == 0) bulk_wait_read_upto(2 * NX - 1 - (2 * h + part));
                    __syncwarp();
                }
                char *yhalf = yrow + (2 * h + part) * kTileBytes;
#pragma unroll
                for (int v = 4 * part; v < 4 * part + 4; ++v) {
                    const double2 in = *reinterpret_cast<const double2 *>(xrow + h * kTileBytes + ((v ^ key) << 4));
                    double yr, yi;
                    fos_sample<ORDER>(in.x, gain, cr, ci, zr, zi, yr, yi);
                    // complex sample 2 v of input half h -> output half 2 h + part, 16-byte chunk (2 v) % 8
                    // (kBulkRows: complex sample 16 h + 2 v of the lane's linear row)
                    if (kBulkRows) *reinterpret_cast<double2 *>(orow + ((16 * h + 2 * v) << 4)) = make_double2(yr, yi);
                    else if (PASS_B) *reinterpret_cast<double2 *>(yhalf + ((((2 * v) & 7) ^ key) << 4)) = make_double2(yr, yi);
                    fos_sample<ORDER>(in.y, gain, cr, ci, zr, zi, yr, yi);
                    if (kBulkRows) *reinterpret_cast<double2 *>(orow + ((16 * h + 2 * v + 1) << 4)) = make_double2(yr, yi);
                    else if (PASS_B) *reinterpret_cast<double2 *>(yhalf + ((((2 * v + 1) & 7) ^ key) << 4)) = make_double2(yr, yi);
                }
            }
        }
        if (PASS_B) fence_proxy_async();
        else asm volatile("fence.acq_rel.cta;\n" ::: "memory");   // pre-pass: see stage_release_after_reads (pfb_tma.cuh)
        if (kBulkRows) {
            if (row_ok) bulk_store_1d(grow + (size_t)t * (2 * NX * kRowBytes), orow, 2 * NX * kRowBytes);
            bulk_commit();
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&empty[s]);
            if (PASS_B && !kBulkRows) {
#pragma unroll
                for (int h = 0; h < 2 * NX; ++h) {
                    tma_store_4d(&ymap, ybuf + h * kTileBytes, (int)(2 * t * TT + h * TLEN), p.JR * w, band, p.CH * cb);
                    bulk_commit();
                }
            }
        }
    }
    if (PASS_B) {
        if (kBulkRows || lane == 0) bulk_wait0();
        if (row_ok && j == p.J - 1) {
            double *st = p.states + (size_t)q * S;
#pragma unroll
            for (int s = 0; s < ORDER; ++s) { st[2 * s] = zr[s]; st[2 * s + 1] = zi[s]; }
        }
    } else if (row_ok) {
#pragma unroll
        for (int s = 0; s < ORDER; ++s) { zs[2 * s] = zr[s]; zs[2 * s + 1] = zi[s]; }
    }
}

}  // namespace pfb
