// pfb_kernels.cuh -- sm_100a device code for cascaded biquad (second-order-section) filtering.
//
// Replaces the inner loops of pyfilterbank/sosfilt.c:5-84.  Two kernels share one "row tile"
// engine:
//
//   sos_seq_kernel   one lane per (channel, band) cascade, sequential in time.  With EXACT the
//                    arithmetic is the reference's (separately rounded mul/add, same order), so the
//                    result is bit-identical to sosfilt.c.
//   sos_scan_kernel  one CTA per cascade, one lane per time chunk.  Pass A runs every chunk from a
//                    zero state over its last `warm` samples to get the state it hands to its
//                    successor, a carry step composes those with the chunk transition matrix
//                    P = M^L (2K x 2K, built on the host in long double), pass B re-runs every chunk
//                    from its true initial state and writes the output.  All lanes of a CTA share one
//                    coefficient set, which arrives in the kernel parameter block so the compiler
//                    keeps it in uniform registers (DFMA R, R, UR, R) instead of 10 K registers per
//                    thread.
//
// Row tile engine: each lane owns one 128-byte row of a shared-memory tile (16 doubles / 32 floats
// of ITS time range).  Rows are filled with 16-byte cp.async copies issued so that 8 consecutive
// lanes cover one row (coalesced 128-byte global segments), filtered in place by the owning lane,
// and written back the same way with streaming 16-byte stores.  The 16-byte chunk c of row r is
// stored at chunk position c ^ (r & 7), which makes both the per-lane (row = lane) and the
// cooperative (row = lane / 8 + 4 i) access patterns free of bank conflicts without padding.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pfb {

constexpr int kRowBytes = 128;
constexpr int kTileBytes = 32 * kRowBytes;    // one buffer of one warp
constexpr int kVecPerRow = kRowBytes / 16;    // 8
constexpr int kCoefBytes = 30000;             // coefficient table carried in the kernel parameters
#ifndef PFB_ROW_BATCH
#define PFB_ROW_BATCH 4                       // 16-byte vectors a lane stages in registers at a time
#endif

struct SosParams {
    const void *x;
    void *y;
    const void *coef;     // [rows][Ktot][5] = b0 b1 b2 a1 a2 in the arithmetic type (seq kernel)
    const void *P;        // [rows][2K][2K] chunk transition (scan kernel), arithmetic type
    void *states;         // [Q][Ktot][2] arithmetic type
    const int *warm;      // [rows] pre-pass samples per chunk (scan kernel) or null (= L, exact)
    long long n, ldx, ldy, L;
    int Q;                // cascades = C * B
    int B;                // bands per channel
    int xdiv;             // x row of cascade q is q / xdiv (B: shared input, 1: own row)
    int per_channel;      // coefficient row is q (1) or q % B (0)
    int Ktot, k0;         // sections per coefficient/state row, first section of this sweep
    int aligned;          // every row start is 16-byte aligned (selects the ALIGNED instantiation)
    int row0, nrows;      // scan kernel: coefficient rows covered by this launch
    int dec;              // 1: keep only every second output sample (decimate by 2 on store), y rows hold n_out samples
    int dec_phase;        // 0 / 1: parity of the kept input indices (output i <-> input 2 i + dec_phase)
};

// Coefficients of the rows one scan launch covers: c[(row - row0) * K + k][5].
template <typename A>
struct CoefTable {
    static constexpr int kMaxSections = kCoefBytes / (5 * (int)sizeof(A));
    A c[kMaxSections][5];
};

// ---------------------------------------------------------------------------------------------
template <typename T> struct Vec16;
template <> struct Vec16<double> { using type = double2; static constexpr int N = 2; };
template <> struct Vec16<float> { using type = float4; static constexpr int N = 4; };

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ void st_stream(double *p, double2 v) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};\n" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void st_stream(float *p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void st_stream1(double *p, double v) {
    asm volatile("st.global.cs.f64 [%0], %1;\n" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void st_stream1(float *p, float v) {
    asm volatile("st.global.cs.f32 [%0], %1;\n" ::"l"(p), "f"(v) : "memory");
}

// ---------------------------------------------------------------------------------------------
// One Direct-Form-II section step (sosfilt.c:17-24).  c = {b0, b1, b2, a1, a2}.
//   EXACT: w0 = (x - a1*w1) - a2*w2 ; y = (b0*w0 + b1*w1) + b2*w2, every op rounded separately.
//   fast : fused multiply-adds ordered so the loop-carried chain w1 -> w0 is a single FMA and the
//          b1*w1 + b2*w2 partial sum does not wait for w0.
template <typename A, bool EXACT> struct Biquad;

template <bool EXACT> struct Biquad<double, EXACT> {
    static __device__ __forceinline__ double step(double x, const double *c, double &w1, double &w2) {
        double w0, y;
        if (EXACT) {
            w0 = __dsub_rn(__dsub_rn(x, __dmul_rn(c[3], w1)), __dmul_rn(c[4], w2));
            y = __dadd_rn(__dadd_rn(__dmul_rn(c[0], w0), __dmul_rn(c[1], w1)), __dmul_rn(c[2], w2));
        } else {
            double t = fma(-c[4], w2, x);
            double u = fma(c[1], w1, c[2] * w2);
            w0 = fma(-c[3], w1, t);
            y = fma(c[0], w0, u);
        }
        w2 = w1;
        w1 = w0;
        return y;
    }
};
template <bool EXACT> struct Biquad<float, EXACT> {
    static __device__ __forceinline__ float step(float x, const float *c, float &w1, float &w2) {
        float w0, y;
        if (EXACT) {
            w0 = __fsub_rn(__fsub_rn(x, __fmul_rn(c[3], w1)), __fmul_rn(c[4], w2));
            y = __fadd_rn(__fadd_rn(__fmul_rn(c[0], w0), __fmul_rn(c[1], w1)), __fmul_rn(c[2], w2));
        } else {
            float t = fmaf(-c[4], w2, x);
            float u = fmaf(c[1], w1, c[2] * w2);
            w0 = fmaf(-c[3], w1, t);
            y = fmaf(c[0], w0, u);
        }
        w2 = w1;
        w1 = w0;
        return y;
    }
};

// Filter the lane's own tile row in place through K sections.  `row` is the start of the lane's
// 128-byte row, `key` = lane & 7 its swizzle key.  valid < TLEN only on the last tile of a row;
// states must not advance over the zero padding.
template <typename T, typename A, int K, bool EXACT, bool STORE>
__device__ __forceinline__ void filter_row(char *row, int key, const A (&c)[K][5], A (&w)[K][2], int valid) {
    using V = typename Vec16<T>::type;
    constexpr int EPV = Vec16<T>::N;
    constexpr int TLEN = kRowBytes / (int)sizeof(T);
    if (valid >= TLEN) {
        // stage kBatch vectors in registers, filter them, write them back: grouping the shared-memory
        // accesses lets the scheduler overlap the section chains of consecutive samples
        constexpr int kBatch = PFB_ROW_BATCH;
#pragma unroll
        for (int v0 = 0; v0 < kVecPerRow; v0 += kBatch) {
            V in[kBatch];
#pragma unroll
            for (int v = 0; v < kBatch; ++v) in[v] = *reinterpret_cast<const V *>(row + (((v0 + v) ^ key) << 4));
#pragma unroll
            for (int v = 0; v < kBatch; ++v) {
                T *e = reinterpret_cast<T *>(&in[v]);
#pragma unroll
                for (int i = 0; i < EPV; ++i) {
                    A s = (A)e[i];
#pragma unroll
                    for (int k = 0; k < K; ++k) s = Biquad<A, EXACT>::step(s, c[k], w[k][0], w[k][1]);
                    e[i] = (T)s;
                }
            }
            if (STORE) {
#pragma unroll
                for (int v = 0; v < kBatch; ++v) *reinterpret_cast<V *>(row + (((v0 + v) ^ key) << 4)) = in[v];
            }
        }
    } else if (valid > 0) {
        for (int i = 0; i < valid; ++i) {
            const int byte = i * (int)sizeof(T);
            T *e = reinterpret_cast<T *>(row + (((byte >> 4) ^ key) << 4) + (byte & 15));
            A s = (A)*e;
#pragma unroll
            for (int k = 0; k < K; ++k) s = Biquad<A, EXACT>::step(s, c[k], w[k][0], w[k][1]);
            if (STORE) *e = (T)s;
        }
    }
}

// One scipy _linear_filter step (direct form II transposed, a[0] == 1), every product and sum rounded separately
// in scipy's order: the weighting filters have a 4-fold zero at z = 1, so FMA contraction alone moves the
// result by ~1e-9 relative.   y = z0 + b0 x ; z[k] = (z[k+1] + x b[k+1]) - y a[k+1] ; z[NZ-1] = x b[NZ] - y a[NZ]
// (splweighting.weight_signal = scipy.signal.lfilter, pyfilterbank/splweighting.py:80)
template <int NZ>
__device__ __forceinline__ double tf_step(double xi, const double (&b)[8], const double (&a)[8], double (&z)[NZ]) {
    const double yi = __dadd_rn(z[0], __dmul_rn(b[0], xi));
#pragma unroll
    for (int k = 0; k < NZ - 1; ++k)
        z[k] = __dsub_rn(__dadd_rn(z[k + 1], __dmul_rn(xi, b[k + 1])), __dmul_rn(yi, a[k + 1]));
    z[NZ - 1] = __dsub_rn(__dmul_rn(xi, b[NZ]), __dmul_rn(yi, a[NZ]));
    return yi;
}

__device__ __forceinline__ int clamp_i(long long v, int hi) { return v <= 0 ? 0 : (v >= hi ? hi : (int)v); }

// ---------------------------------------------------------------------------------------------
// Cooperative tile I/O of one warp.  Tile row r <-> global row  base + r * row_stride  (elements);
// lane (rsub = lane / 8, seg = lane % 8) moves the 16-byte segment `seg` of rows rsub, rsub + 4, ...
// so every instruction covers four complete 128-byte rows.  `fast` paths assume every row holds
// the whole tile (no bounds checks, 16-byte vectors); the general paths clamp per row and fall back
// to element accesses when rows are not 16-byte aligned.
template <typename T, bool ALIGNED>
struct CoopIO {
    using V = typename Vec16<T>::type;
    static constexpr int EPV = Vec16<T>::N;
    const T *ld;          // load pointer of row rsub, segment seg, element 0
    T *st;                // store pointer, same position
    long long rstep_ld;   // 4 * row_stride (elements)
    long long rstep_st;
    uint32_t sm[2];       // smem byte offset of (row rsub, segment seg) swizzled for even / odd r
    int rsub, seg;

    __device__ __forceinline__ CoopIO(const T *xbase, T *ybase, long long ld_stride, long long st_stride, int lane)
        : rsub(lane >> 3), seg(lane & 7) {
        ld = xbase + rsub * ld_stride + seg * EPV;
        st = ybase + rsub * st_stride + seg * EPV;
        rstep_ld = 4 * ld_stride;
        rstep_st = 4 * st_stride;
        // row = 4 r + rsub  =>  row & 7 = ((r & 1) << 2) | rsub
        sm[0] = rsub * kRowBytes + ((seg ^ rsub) << 4);
        sm[1] = rsub * kRowBytes + ((seg ^ (4 | rsub)) << 4);
    }
    __device__ __forceinline__ uint32_t off(int r) const { return sm[r & 1] + r * 4 * kRowBytes; }

    __device__ __forceinline__ void load_fast(char *buf, long long e0) const {
        const uint32_t d = smem_addr(buf);
        const T *p = ld + e0;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (ALIGNED) {
                cp_async16(d + off(r), p, 16);
            } else {
#pragma unroll
                for (int i = 0; i < EPV; ++i) {
                    if (sizeof(T) == 8) cp_async8(d + off(r) + 8 * i, p + i, 8);
                    else cp_async4(d + off(r) + 4 * i, p + i, 4);
                }
            }
            p += rstep_ld;
        }
    }
    template <typename LenFn>
    __device__ __forceinline__ void load(char *buf, long long e0, LenFn len, const T *safe) const {
        const uint32_t d = smem_addr(buf);
        const long long e = e0 + seg * EPV;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int valid = clamp_i(len(r * 4 + rsub) - e, EPV);
            const T *p = valid > 0 ? ld + e0 + r * rstep_ld : safe;
            if (ALIGNED) {
                cp_async16(d + off(r), p, valid * (int)sizeof(T));
            } else {
#pragma unroll
                for (int i = 0; i < EPV; ++i) {
                    const bool ok = i < valid;
                    if (sizeof(T) == 8) cp_async8(d + off(r) + 8 * i, ok ? p + i : safe, ok ? 8 : 0);
                    else cp_async4(d + off(r) + 4 * i, ok ? p + i : safe, ok ? 4 : 0);
                }
            }
        }
    }
    __device__ __forceinline__ void store_fast(const char *buf, long long e0) const {
        T *p = st + e0;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            V v = *reinterpret_cast<const V *>(buf + off(r));
            if (ALIGNED) {
                st_stream(p, v);
            } else {
                const T *ev = reinterpret_cast<const T *>(&v);
#pragma unroll
                for (int i = 0; i < EPV; ++i) st_stream1(p + i, ev[i]);
            }
            p += rstep_st;
        }
    }
    template <typename LenFn>
    __device__ __forceinline__ void store(const char *buf, long long e0, LenFn len) const {
        const long long e = e0 + seg * EPV;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int valid = clamp_i(len(r * 4 + rsub) - e, EPV);
            if (valid > 0) {
                V v = *reinterpret_cast<const V *>(buf + off(r));
                T *p = st + e0 + r * rstep_st;
                if (ALIGNED && valid == EPV) {
                    st_stream(p, v);
                } else {
                    const T *ev = reinterpret_cast<const T *>(&v);
#pragma unroll
                    for (int i = 0; i < EPV; ++i)
                        if (i < valid) st_stream1(p + i, ev[i]);
                }
            }
        }
    }
};

// Decimating tile store: of tile elements [e0, e0 + TLEN) of every row keep those whose index within the
// row has parity `phase`; element idx goes to output position (idx - phase) / 2 of the row at
// ybase + row * out_stride.  Scalar streaming stores (the decimated stream is a small part of the traffic).
template <typename T, typename LenFn>
__device__ __forceinline__ void store_tile_dec(const char *buf, T *ybase, long long out_stride, long long e0, LenFn len,
                                               int phase, int lane) {
    using V = typename Vec16<T>::type;
    constexpr int EPV = Vec16<T>::N;
    const int rsub = lane >> 3, seg = lane & 7;
    const long long e = e0 + seg * EPV;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int row = r * 4 + rsub;
        const long long rl = len(row);
        if (e < rl) {
            const V v = *reinterpret_cast<const V *>(buf + row * kRowBytes + ((seg ^ (row & 7)) << 4));
            const T *ev = reinterpret_cast<const T *>(&v);
            T *prow = ybase + row * out_stride;
#pragma unroll
            for (int i = phase & 1; i < EPV; i += 2)
                if (e + i < rl) st_stream1(prow + ((e + i - phase) >> 1), ev[i]);
        }
    }
}

// Per-lane tile load: the lane copies elements [e0, e0 + TLEN) of ITS row into its tile row (lanes
// that share an input row hit the same global sectors, which the L1 coalesces).
template <typename T, bool ALIGNED>
__device__ __forceinline__ void load_row_own(char *row, int key, const T *row_ptr, long long len, long long e0,
                                             const T *safe) {
    constexpr int EPV = Vec16<T>::N;
    const uint32_t sb = smem_addr(row);
    const bool full = e0 + kRowBytes / (int)sizeof(T) <= len;
#pragma unroll
    for (int v = 0; v < kVecPerRow; ++v) {
        const long long e = e0 + v * EPV;
        const int valid = full ? EPV : clamp_i(len - e, EPV);
        const T *src = valid > 0 ? row_ptr + e : safe;
        const uint32_t d = sb + ((v ^ key) << 4);
        if (ALIGNED) {
            cp_async16(d, src, valid * (int)sizeof(T));
        } else {
#pragma unroll
            for (int i = 0; i < EPV; ++i) {
                const bool ok = i < valid;
                if (sizeof(T) == 8) cp_async8(d + 8 * i, ok ? src + i : safe, ok ? 8 : 0);
                else cp_async4(d + 4 * i, ok ? src + i : safe, ok ? 4 : 0);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Sequential kernel: lane <-> cascade q.  blockDim.x = 32 * warps, dynamic smem = warps * 2 tiles.
template <typename T, typename A, int K, bool EXACT, bool ALIGNED>
__global__ void __launch_bounds__(128, 4) sos_seq_kernel(const SosParams p) {
    extern __shared__ __align__(128) char smem[];
    constexpr int TLEN = kRowBytes / (int)sizeof(T);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, key = lane & 7;
    char *buf0 = smem + warp * 2 * kTileBytes;
    const long long q0 = (long long)blockIdx.x * blockDim.x + warp * 32;   // first cascade of the warp
    const long long q = q0 + lane;
    const bool active = q < p.Q;
    const int Q = p.Q;
    const long long n = p.n;
    if (q0 >= Q) return;

    A c[K][5], w[K][2];
    A *st = reinterpret_cast<A *>(p.states) + ((size_t)(active ? q : 0) * p.Ktot + p.k0) * 2;
    if (active) {
        const int srow = p.per_channel ? (int)q : (int)(q % p.B);
        const A *src = reinterpret_cast<const A *>(p.coef) + ((size_t)srow * p.Ktot + p.k0) * 5;
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
            for (int i = 0; i < 5; ++i) c[k][i] = src[k * 5 + i];
            w[k][0] = st[2 * k];
            w[k][1] = st[2 * k + 1];
        }
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
            for (int i = 0; i < 5; ++i) c[k][i] = A(0);
            w[k][0] = w[k][1] = A(0);
        }
    }
    const T *xbase = reinterpret_cast<const T *>(p.x);
    const T *xrow = xbase + (active ? (q / p.xdiv) : 0) * p.ldx;
    T *ybase = reinterpret_cast<T *>(p.y) + q0 * p.ldy;
    const long long mylen = active ? n : 0;
    const bool warp_full = q0 + 32 <= Q;
    auto rowlen = [=](int row) -> long long { return (q0 + row < Q) ? n : 0; };
    CoopIO<T, ALIGNED> io(xbase, ybase, 0, p.ldy, lane);
    char *myrow0 = buf0 + lane * kRowBytes;

    const long long nt = (n + TLEN - 1) / TLEN;
    // tiles [0, tf): tile t and t+1 complete in every row of a full warp -> unchecked, double-buffered
    const long long full_tiles = warp_full ? n / TLEN : 0;
    const long long tf = full_tiles - 1 > 0 ? full_tiles - 1 : 0;
    const bool primed = full_tiles > 0;
    if (primed) {
        load_row_own<T, ALIGNED>(myrow0, key, xrow, mylen, 0, xbase);
        cp_async_commit();
    }
    for (long long t = 0; t < tf; ++t) {
        const int cb = (int)(t & 1);
        load_row_own<T, ALIGNED>(myrow0 + (cb ^ 1) * kTileBytes, key, xrow, mylen, (t + 1) * TLEN, xbase);
        cp_async_commit();
        cp_async_wait<1>();
        filter_row<T, A, K, EXACT, true>(myrow0 + cb * kTileBytes, key, c, w, TLEN);
        __syncwarp();
        if (p.dec) store_tile_dec<T>(buf0 + cb * kTileBytes, ybase, p.ldy, t * TLEN, rowlen, p.dec_phase, lane);
        else io.store_fast(buf0 + cb * kTileBytes, t * TLEN);
        __syncwarp();
    }
    for (long long t = tf; t < nt; ++t) {
        const int cb = (int)(t & 1);
        if (!(primed && t == tf)) {
            load_row_own<T, ALIGNED>(myrow0 + cb * kTileBytes, key, xrow, mylen, t * TLEN, xbase);
            cp_async_commit();
        }
        cp_async_wait<0>();
        filter_row<T, A, K, EXACT, true>(myrow0 + cb * kTileBytes, key, c, w, clamp_i(mylen - t * TLEN, TLEN));
        __syncwarp();
        if (p.dec) store_tile_dec<T>(buf0 + cb * kTileBytes, ybase, p.ldy, t * TLEN, rowlen, p.dec_phase, lane);
        else io.store(buf0 + cb * kTileBytes, t * TLEN, rowlen);
        __syncwarp();
    }
    if (active && n > 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) { st[2 * k] = w[k][0]; st[2 * k + 1] = w[k][1]; }
    }
}

// ---------------------------------------------------------------------------------------------
// Scan kernel.  One launch covers coefficient rows [row0, row0 + nrows): for a shared bank those
// are bands and blockIdx.x = (band - row0) * C + channel (bands outer, so CTAs resident together
// carry a similar band-dependent load); for a per-channel bank they are cascades.
// Thread j <-> time chunk [j*L, (j+1)*L).  dynamic smem = warps * 2 tiles; the per-chunk states of
// the carry step live in the (then idle) tile buffers.
template <int K> struct ScanBounds { static constexpr int kMinBlocks = K <= 6 ? 3 : 2; };

template <typename T, typename A, int K, bool ALIGNED>
__global__ void __launch_bounds__(256, ScanBounds<K>::kMinBlocks)
sos_scan_kernel(const __grid_constant__ SosParams p, const __grid_constant__ CoefTable<A> tab) {
    extern __shared__ __align__(128) char smem[];
    constexpr int TLEN = kRowBytes / (int)sizeof(T);
    constexpr int S = 2 * K;
    static_assert(S * sizeof(A) <= kRowBytes, "carry slot must fit in a tile row");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, key = lane & 7;
    const int nchunks = blockDim.x;
    char *buf0 = smem + warp * 2 * kTileBytes;
    auto zslot = [&](int jj) -> A * {   // carry slot of chunk jj: row (jj & 31) of warp (jj >> 5)'s first buffer
        return reinterpret_cast<A *>(smem + (jj >> 5) * 2 * kTileBytes + (jj & 31) * kRowBytes);
    };

    int q, rl;   // cascade, coefficient row relative to row0 (both uniform over the CTA)
    if (p.per_channel) {
        rl = blockIdx.x;
        q = p.row0 + rl;
    } else {
        const int C = p.Q / p.B;
        rl = blockIdx.x / C;
        q = (blockIdx.x - rl * C) * p.B + p.row0 + rl;
    }
    const int srow = p.row0 + rl;
    const long long n = p.n, L = p.L;
    const int j = threadIdx.x, j0 = warp * 32;
    const long long myrem = n - (long long)j * L;
    const long long mylen = myrem <= 0 ? 0 : (myrem >= L ? L : myrem);

    A c[K][5], w[K][2];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int i = 0; i < 5; ++i) c[k][i] = tab.c[rl * K + k][i];

    const T *xbase = reinterpret_cast<const T *>(p.x);
    const T *xw = xbase + (long long)(q / p.xdiv) * p.ldx + (long long)j0 * L;   // warp's first chunk
    T *yw = reinterpret_cast<T *>(p.y) + (long long)q * p.ldy + (long long)j0 * L;
    // decimating store: L is even, so chunk j starts at output index j * L / 2 and parities are chunk-local
    T *yw_dec = reinterpret_cast<T *>(p.y) + (long long)q * p.ldy + (long long)j0 * (L >> 1);
    auto rowlen = [=](int row) -> long long {
        const long long r = n - (long long)(j0 + row) * L;
        return r <= 0 ? 0 : (r >= L ? L : r);
    };
    CoopIO<T, ALIGNED> io(xw, yw, L, L, lane);
    char *myrow0 = buf0 + lane * kRowBytes;
    const long long nt = L / TLEN;                            // L is a multiple of TLEN
    const bool warp_has_work = (long long)j0 * L < n;
    // rows only get shorter with the lane index, so row 31 bounds how many tiles are complete in
    // every row of the warp; those tiles take the unchecked 16-byte path
    const long long fast_tiles = rowlen(31) / TLEN;

    // ---- pass A: zero-state final state of every chunk that feeds a successor, over its last
    //      `warm` samples (warm == L: exact) ----
#pragma unroll
    for (int k = 0; k < K; ++k) w[k][0] = w[k][1] = A(0);
    const bool need_z = (long long)(j + 1) * L < n;           // chunk j+1 non-empty => chunk j is full
    long long warm = p.warm ? (long long)p.warm[srow] : L;
    if (warm > L) warm = L;
    const long long ta = nt - (warm + TLEN - 1) / TLEN;
    if ((long long)(j0 + 1) * L < n && ta < nt) {
        // tiles [ta, tf) and their successors are complete in every row: double-buffered, unchecked
        const long long tf = fast_tiles - 1 > ta ? fast_tiles - 1 : ta;
        const bool primed = ta < fast_tiles;
        if (primed) {
            io.load_fast(buf0 + (int)(ta & 1) * kTileBytes, ta * TLEN);
            cp_async_commit();
        }
        for (long long t = ta; t < tf; ++t) {
            const int cb = (int)(t & 1);
            io.load_fast(buf0 + (cb ^ 1) * kTileBytes, (t + 1) * TLEN);
            cp_async_commit();
            cp_async_wait<1>();
            __syncwarp();
            if (need_z) filter_row<T, A, K, false, false>(myrow0 + cb * kTileBytes, key, c, w, TLEN);
            __syncwarp();
        }
        // remaining tiles (at least the last one): bounds-checked, single-buffered
        for (long long t = tf; t < nt; ++t) {
            const int cb = (int)(t & 1);
            if (!(primed && t == tf)) {
                io.load(buf0 + cb * kTileBytes, t * TLEN, rowlen, xbase);
                cp_async_commit();
            }
            cp_async_wait<0>();
            __syncwarp();
            if (need_z) filter_row<T, A, K, false, false>(myrow0 + cb * kTileBytes, key, c, w, TLEN);
            __syncwarp();
        }
    }
    {
        A *z = zslot(j);
#pragma unroll
        for (int k = 0; k < K; ++k) { z[2 * k] = w[k][0]; z[2 * k + 1] = w[k][1]; }
    }
    __syncthreads();

    // ---- carry: s[j+1] = P s[j] + z[j]; slot j is overwritten with chunk j's initial state ----
    A *st = reinterpret_cast<A *>(p.states) + ((size_t)q * p.Ktot + p.k0) * 2;
    if (threadIdx.x == 0) {
        const A *P = reinterpret_cast<const A *>(p.P) + (size_t)srow * S * S;
        A s[S];
#pragma unroll
        for (int i = 0; i < S; ++i) s[i] = st[i];
        for (int jj = 0; jj < nchunks; ++jj) {
            A *zj = zslot(jj);
            A z[S];
#pragma unroll
            for (int i = 0; i < S; ++i) { z[i] = zj[i]; zj[i] = s[i]; }
            if ((long long)(jj + 1) * L >= n) break;          // chunk jj+1 is empty
#pragma unroll
            for (int i = 0; i < S; ++i) {
                A a = z[i];
#pragma unroll
                for (int m = 0; m < S; ++m) a = fma(P[i * S + m], s[m], a);
                z[i] = a;
            }
#pragma unroll
            for (int i = 0; i < S; ++i) s[i] = z[i];
        }
    }
    __syncthreads();

    // ---- pass B: filter every chunk from its true initial state, write the output ----
    if (!warp_has_work) return;
    {
        const A *z = zslot(j);
#pragma unroll
        for (int k = 0; k < K; ++k) { w[k][0] = z[2 * k]; w[k][1] = z[2 * k + 1]; }
    }
    __syncwarp();   // every lane has its state before the tile buffers are refilled
    {
        const long long tf = fast_tiles - 1 > 0 ? fast_tiles - 1 : 0;
        const bool primed = fast_tiles > 0;
        if (primed) {
            io.load_fast(buf0, 0);
            cp_async_commit();
        }
        for (long long t = 0; t < tf; ++t) {   // tiles t and t+1 are complete in every row
            const int cb = (int)(t & 1);
            io.load_fast(buf0 + (cb ^ 1) * kTileBytes, (t + 1) * TLEN);
            cp_async_commit();
            cp_async_wait<1>();
            __syncwarp();
            filter_row<T, A, K, false, true>(myrow0 + cb * kTileBytes, key, c, w, TLEN);
            __syncwarp();
            if (p.dec) store_tile_dec<T>(buf0 + cb * kTileBytes, yw_dec, L >> 1, t * TLEN, rowlen, p.dec_phase, lane);
            else io.store_fast(buf0 + cb * kTileBytes, t * TLEN);
            __syncwarp();
        }
        for (long long t = tf; t < nt; ++t) {   // bounds-checked tail, single-buffered
            const int cb = (int)(t & 1);
            if (!(primed && t == tf)) {
                io.load(buf0 + cb * kTileBytes, t * TLEN, rowlen, xbase);
                cp_async_commit();
            }
            cp_async_wait<0>();
            __syncwarp();
            filter_row<T, A, K, false, true>(myrow0 + cb * kTileBytes, key, c, w, clamp_i(mylen - t * TLEN, TLEN));
            __syncwarp();
            if (p.dec) store_tile_dec<T>(buf0 + cb * kTileBytes, yw_dec, L >> 1, t * TLEN, rowlen, p.dec_phase, lane);
            else io.store(buf0 + cb * kTileBytes, t * TLEN, rowlen);
            __syncwarp();
        }
    }
    // the thread that owns the last non-empty chunk holds the final state
    if (mylen > 0 && (long long)(j + 1) * L >= n) {
#pragma unroll
        for (int k = 0; k < K; ++k) { st[2 * k] = w[k][0]; st[2 * k + 1] = w[k][1]; }
    }
}

// launchers implemented per (T, A) in pfb_inst_*.cu; return cudaError_t as int.
// The scan launchers take the host copy of the coefficient rows ([rows][Ktot][5] doubles) and loop
// over row ranges that fit the parameter-block table; *launches counts the kernels issued.
int launch_seq_dd(const SosParams &p, int K, bool exact, cudaStream_t s);
int launch_seq_ff(const SosParams &p, int K, bool exact, cudaStream_t s);
int launch_seq_fd(const SosParams &p, int K, bool exact, cudaStream_t s);
int launch_scan_dd(SosParams p, int K, int warps, const double *coef_host, int rows, int C, cudaStream_t s, int *launches);
int launch_scan_ff(SosParams p, int K, int warps, const double *coef_host, int rows, int C, cudaStream_t s, int *launches);
int launch_scan_fd(SosParams p, int K, int warps, const double *coef_host, int rows, int C, cudaStream_t s, int *launches);

constexpr int kSeqWarps = 4;

}  // namespace pfb
