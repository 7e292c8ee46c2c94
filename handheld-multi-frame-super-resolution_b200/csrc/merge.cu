// Kernel-regression merge (Alg. 4 and Alg. 11 of the IPOL paper) for sm_100a.
//
// Replaces handheld_super_resolution/merge.py:82-233 (accumulate_ref) and :290-434 (accumulate), plus
// utils.py:62-120 (divide, add).  HBM-bound: per comp frame the accumulators num/den [Hs][Ws][3] f32 are read and
// written once (48 B per HR pixel) and raw/r/covariances are read once (12 B per LR pixel) — see DESIGN.md.
//
// Layout / mapping: one thread owns 4 consecutive HR pixels of one row, so its slice of each accumulator is
// 12 consecutive floats = 3 x 16 bytes (the [.,.,3] interleaving is part of the boundary: main() returns it).
// LR-side gathers (raw 3x3 taps, covariance quads, r, tile flow) go through the read-only L1/L2 path; they are
// reused ~36x across neighbouring HR pixels.
//
// Kernels:
//   accumulate_pow2_kernel  scales 1, 2, 4 (the benchmark path): exact float32 position split, no float64; the
//                           accumulators are updated by fire-and-forget 16-byte L2 reductions (never loaded by the SM);
//   accumulate_kernel       any scale: the reference's float64 position per pixel, float4 load/add/store;
//   accumulate_pow2_batch_kernel / accumulate_batch_kernel   several frames per pass over the accumulators, sums in registers;
//   accumulate_ref_kernel   the reference frame (+ fused divide, + the frame-sharded peer sum, hhsr_reduce_merge_ref).
// All of them share the per-pixel device functions below and the update  acc <- add.ftz(acc, r * sum)  and are
// bit-equal where their domains overlap (tests/test_gpu_parity.py).  Compiled with -fmad=false (csrc/Makefile): every
// fused multiply-add in this file is written explicitly, so the kernels round identically whatever the inlining.
#include "common.cuh"

// resident CTAs per SM the register allocator must allow (measured optimum, profiles/merge_accumulate_r01_ncu.md)
#ifndef HHSR_MERGE_MINBLOCKS
#define HHSR_MERGE_MINBLOCKS 5
#endif
#ifndef HHSR_MERGE_POW2_MINBLOCKS
#define HHSR_MERGE_POW2_MINBLOCKS 4
#endif

namespace hhsr {

struct MergeFrame {
    const float *raw, *flow, *covs, *r;
};
constexpr int kMaxBatch = 24;
struct MergeBatch {
    MergeFrame f[kMaxBatch];
    int K;
};
// CFA descriptor: packed 2x2 channel ids + whether it is a Bayer pattern (one channel twice, on a diagonal).
struct CfaInfo {
    int packed;
    int bayer;       // 1: the fast channel resolve is valid
    int dup;         // the duplicated channel (green)
    int dup_main;    // 1 if the duplicated channel sits on the main diagonal (0,0),(1,1)
    int lo;          // the smaller of the two single channels
};
struct MergeGeom {
    int H, W, nx, ts, ch, cw, Hs, Ws;
    CfaInfo cfa;
    double scale, inv_scale;
    bool pow2;      // scale is a power of two: x / scale == x * (1/scale) exactly
    int ts_shift;   // log2(ts) when ts is a power of two, else -1
    int row_begin, row_end;   // output rows processed by the batched kernels (0, Hs unless row-sharded)
};

__device__ __forceinline__ int tile_of(int i, const MergeGeom &g) { return g.ts_shift >= 0 ? (i >> g.ts_shift) : (i / g.ts); }

static CfaInfo make_cfa(const int *c) {
    CfaInfo f{pack_cfa(c), 0, 0, 0, 0};
    int dup = -1, main_diag = 0;
    if (c[0] == c[3] && c[1] != c[2]) dup = c[0], main_diag = 1;
    if (c[1] == c[2] && c[0] != c[3]) dup = c[1], main_diag = 0;
    if (dup >= 0) {
        const int s0 = main_diag ? c[1] : c[0], s1 = main_diag ? c[2] : c[3];
        if (s0 != dup && s1 != dup && s0 != s1) {
            f.bayer = 1, f.dup = dup, f.dup_main = main_diag, f.lo = s0 < s1 ? s0 : s1;
        }
    }
    return f;
}

static MergeGeom make_geom(int H, int W, int nx, int ts, int Hs, int Ws, const int *cfa, double scale) {
    int e = 0;
    const bool pow2 = std::frexp(scale, &e) == 0.5;
    int shift = -1;
    for (int k = 0; k < 16; ++k)
        if ((1 << k) == ts) shift = k;
    return MergeGeom{H, W, nx, ts, H / 2, W / 2, Hs, Ws, make_cfa(cfa), scale, 1.0 / scale, pow2, shift, 0, Hs};
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// acc + r * x with the product rounded separately and flush-to-zero addition: the arithmetic of the L2 reduction the
// fast path uses (see rmw4)
__device__ __forceinline__ float add_ftz(float acc, float r, float x) {
    float y;
    asm("add.ftz.f32 %0, %1, %2;" : "=f"(y) : "f"(acc), "f"(r * x));
    return y;
}

// (hr + 0.5) / scale exactly as the reference forms it in float64 (merge.py:319-320).  For power-of-two scales the
// multiplication by the (exact) reciprocal is the same correctly-rounded result and skips the fp64 division.
__device__ __forceinline__ double lr_coord(int hr, double scale, double inv_scale, bool pow2) {
    const double a = (double)hr + 0.5;
    return pow2 ? a * inv_scale : a / scale;
}

// Split m = lr + flow (float64, merge.py:339-340) into integer part and float32 fraction; false if m leaves
// [0, n) (merge.py:343-345).
__device__ __forceinline__ bool split_pos(double lr, float flow, int n, int &c, float &t) {
    const double m = lr + (double)flow;
    if (!(m >= 0.0 && m < (double)n)) return false;
    c = (int)m;
    t = (float)(m - (double)c);
    return true;
}

// Bilinear lookup coordinates of the covariance map at k = m/2 - 0.5 (merge.py:350-363, trunc + signed modf),
// derived exactly from m = c + t:  c odd -> k = (c-1)/2 + t/2;  c even -> k = c/2 - 1 + (0.5 + t/2);
// c == 0 -> k in [-0.5, 0): index 0 with a negative fraction (the reference extrapolates there).
__device__ __forceinline__ void cov_coord(int c, float t, int n, int &i0, int &i1, float &fr) {
    if (c & 1) {
        i0 = c >> 1;
        fr = 0.5f * t;
    } else if (c > 0) {
        i0 = (c >> 1) - 1;
        fr = fmaf(0.5f, t, 0.5f);
    } else {
        i0 = 0;
        fr = fmaf(0.5f, t, -0.5f);
    }
    i1 = min(i0 + 1, n - 1);
}

// 3x3 taps around (ci, cj).  v/a: partial sums per tap parity RELATIVE to the centre tap (row parity, col parity),
// so the accumulation indices are compile-time; the CFA channel of each partial is resolved once at the end.
// CHECK=false is the interior fast path (all 9 taps inside the frame, no per-tap tests).
template <bool CHECK>
__device__ __forceinline__ void merge_taps(const float *__restrict__ pc, int H, int W, int ci, int cj, float tx, float ty,
                                           float qxx, float qxy, float qyy, float (&v)[2][2], float (&a)[2][2]) {
#pragma unroll
    for (int di = -1; di <= 1; ++di) {
        if (CHECK && (ci + di < 0 || ci + di >= H)) continue;
        const float dy = (float)di + 0.5f - ty;                                   // i - (lr_mov_y - 0.5)
        const float qy = qyy * dy * dy, qm = qxy * dy;
        const float *row = pc + di * W;
#pragma unroll
        for (int dj = -1; dj <= 1; ++dj) {
            if (CHECK && (cj + dj < 0 || cj + dj >= W)) continue;
            const float c = __ldg(row + dj);
            const float dx = (float)dj + 0.5f - tx;
            float z = fmaf(fmaf(qxx, dx, qm), dx, qy);
            z = fminf(0.0f, z);               // == -0.5*log2e*max(0, z_ref); NaN -> 0
            const float w = ex2_approx(z);
            v[di & 1][dj & 1] = fmaf(w, c, v[di & 1][dj & 1]);
            a[di & 1][dj & 1] += w;
        }
    }
}

// Fold the four parity partials into the three colour channels (unscaled: every kernel applies the robustness as
// acc = add_ftz(acc, r * sum), so single-frame, batched and fast-path launches agree bit for bit).
// Bayer patterns take 7 selects; anything else the generic 12.
__device__ __forceinline__ void resolve_channels(const CfaInfo &cf, int ci, int cj, const float (&v)[2][2], float (&out)[3]) {
    float ch[3];
    if (cf.bayer) {
        // rel main diagonal == abs main diagonal iff ci and cj have equal parity
        const bool gm = ((cf.dup_main ^ ((ci ^ cj) & 1)) != 0);      // duplicated channel on the REL main diagonal
        const float g = (gm ? v[0][0] : v[0][1]) + (gm ? v[1][1] : v[1][0]);
        const float p = gm ? v[0][1] : v[0][0];                      // rel (0, gm) -> abs (ci, cj + gm)
        const float q = gm ? v[1][0] : v[1][1];
        const bool p_is_lo = (cfa_channel(cf.packed, ci, cj + (gm ? 1 : 0)) == cf.lo);
        const float lo = p_is_lo ? p : q, hi = p_is_lo ? q : p;
        if (cf.dup == 1) {                                           // RGGB / BGGR / GRBG / GBRG: green is channel 1
            ch[0] = lo, ch[1] = g, ch[2] = hi;
        } else {
            ch[0] = (cf.dup == 0) ? g : (cf.lo == 0 ? lo : hi);
            ch[1] = (cf.lo == 1) ? lo : hi;
            ch[2] = (cf.dup == 2) ? g : (cf.lo == 2 ? lo : hi);
        }
    } else {
        ch[0] = ch[1] = ch[2] = 0.f;
#pragma unroll
        for (int ry = 0; ry < 2; ++ry)
#pragma unroll
            for (int rx = 0; rx < 2; ++rx) {
                const int chn = cfa_channel(cf.packed, ci + ry, cj + rx);
#pragma unroll
                for (int k = 0; k < 3; ++k) ch[k] += (chn == k) ? v[ry][rx] : 0.0f;
            }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) out[k] = ch[k];
}

// Interpolated covariance -> pre-scaled inverse quadratic form.  w = exp(-z/2) = 2^(qxx dx^2 + qxy dx dy + qyy dy^2)
struct CovQuads {
    int fx0, fy0;
    float4 tr, tl, br, bl;
};

// Bilinear blend of the four covariance quads (merge.py:365-389) and its pre-scaled inverse.  The blend runs along y
// first (the two quad COLUMNS), then along x: the fast path shares the y-blended columns between the four pixels of a
// thread, and every kernel uses this order so that they stay bit-equal (the reference blends x first in float64;
// the difference is float32 rounding).
struct CovCol {
    float xx, xy, yy;
};
__device__ __forceinline__ CovCol cov_column(const float4 &top, const float4 &bot, float fry) {
    CovCol c;
    c.xx = fmaf(fry, bot.x - top.x, top.x);
    c.xy = fmaf(fry, bot.y - top.y, top.y);
    c.yy = fmaf(fry, bot.w - top.w, top.w);
    return c;
}
__device__ __forceinline__ void cov_form_cols(const CovCol &l, const CovCol &r, float frx, float &qxx, float &qxy, float &qyy) {
    const float kS = -0.72134752044448170368f;   // -0.5 * log2(e)
    const float cxx = fmaf(frx, r.xx - l.xx, l.xx);
    const float cxy = fmaf(frx, r.xy - l.xy, l.xy);
    const float cyy = fmaf(frx, r.yy - l.yy, l.yy);
    const float inv_det = __fdividef(kS, fmaf(cxx, cyy, -(cxy * cxy)));      // merge.py:391-396
    qxx = inv_det * cyy;
    qxy = -2.0f * inv_det * cxy;
    qyy = inv_det * cxx;
}
// tr/br: quads at column fx0 (rows fy0 / cy1), tl/bl: at column cx1
__device__ __forceinline__ void cov_form(const float4 &tr, const float4 &tl, const float4 &br, const float4 &bl, float frx,
                                         float fry, float &qxx, float &qxy, float &qyy) {
    cov_form_cols(cov_column(tr, br, fry), cov_column(tl, bl, fry), frx, qxx, qxy, qyy);
}

template <bool ISO>
__device__ __forceinline__ void merge_pixel(const MergeFrame &f, const MergeGeom &g, int cj, float tx, int ci, float ty,
                                            CovQuads &cq, float (&val)[3], float (&acc)[3]) {
    const float kS = -0.72134752044448170368f;   // -0.5 * log2(e)
    float qxx, qxy, qyy;
    if (ISO) {
        qxx = qyy = 2.0f * kS;   // z = 2 (dx^2 + dy^2), merge.py:419
        qxy = 0.0f;
    } else {
        int fx0, cx1, fy0, cy1;
        float frx, fry;
        cov_coord(cj, tx, g.cw, fx0, cx1, frx);
        cov_coord(ci, ty, g.ch, fy0, cy1, fry);
        {
            const float4 *c4 = reinterpret_cast<const float4 *>(f.covs);
            const float4 *r0 = c4 + (unsigned)fy0 * (unsigned)g.cw, *r1 = c4 + (unsigned)cy1 * (unsigned)g.cw;
            cq.tr = __ldg(r0 + fx0), cq.tl = __ldg(r0 + cx1), cq.br = __ldg(r1 + fx0), cq.bl = __ldg(r1 + cx1);
            cq.fx0 = fx0, cq.fy0 = fy0;
        }
        cov_form(cq.tr, cq.tl, cq.br, cq.bl, frx, fry, qxx, qxy, qyy);
    }
    float v[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, a[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    const float *pc = f.raw + ((unsigned)max(ci, 0) * (unsigned)g.W + (unsigned)max(cj, 0));
    if (ci >= 1 && ci <= g.H - 2 && cj >= 1 && cj <= g.W - 2)
        merge_taps<false>(pc, g.H, g.W, ci, cj, tx, ty, qxx, qxy, qyy, v, a);
    else
        merge_taps<true>(pc, g.H, g.W, ci, cj, tx, ty, qxx, qxy, qyy, v, a);
    // the reference multiplies every tap weight by r (merge.py:430-431); factored out (float32 rounding level) and
    // applied by the caller
    resolve_channels(g.cfa, ci, cj, v, val);
    resolve_channels(g.cfa, ci, cj, a, acc);
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// One thread = VEC consecutive HR pixels of one row.  The accumulator slice (2 x 3 float4) is prefetched into L2 at
// entry and only loaded after the gathers/weights are done, so its registers are not live during the math (more
// resident warps) while its HBM latency still overlaps the math.
// Per-pixel front end shared by both kernels: position, tile flow, robustness, then the taps.
// RowCtx caches the y-axis split for the flow tile of the previous pixel (4 consecutive pixels nearly always share it).
struct RowCtx {
    double lr_y;
    int py, i_r;
    int tile_x;     // flow tile column the cached split belongs to (-1: none)
    float2 fl;
    int ci;
    float ty;
    bool ok_y;
};

// val/acc: unscaled channel sums of this pixel; returns its robustness r (0 with val = acc = 0 when the pixel is skipped)
template <bool ISO>
__device__ __forceinline__ float merge_hr_pixel(const MergeFrame &f, const MergeGeom &g, int hr_j, RowCtx &rc, CovQuads &cq,
                                                float (&val)[3], float (&acc)[3]) {
    const double lr_x = lr_coord(hr_j, g.scale, g.inv_scale, g.pow2);
    const int ilx = (int)lr_x;
    const int tcol = tile_of(ilx, g);
    if (tcol != rc.tile_x) {
        rc.tile_x = tcol;
        rc.fl = __ldg(reinterpret_cast<const float2 *>(f.flow) + (unsigned)rc.py * (unsigned)g.nx + (unsigned)tcol);
        rc.ok_y = split_pos(rc.lr_y, rc.fl.y, g.H, rc.ci, rc.ty);
    }
    int cj;
    float tx;
    if (!rc.ok_y || !split_pos(lr_x, rc.fl.x, g.W, cj, tx)) return 0.0f;
    const float local_r = __ldg(f.r + (unsigned)rc.i_r * (unsigned)g.W + (unsigned)min(ilx, g.W - 1));
    merge_pixel<ISO>(f, g, cj, tx, rc.ci, rc.ty, cq, val, acc);
    return local_r;
}

// Single comp frame (the reference's launch granularity, merge.py:284-287).  One thread = VEC consecutive HR pixels
// of one row.  The accumulator slice (2 x 3 float4) is prefetched into L2 at entry and only loaded after the
// gathers/weights are done: its registers are not live during the math (more resident warps) while its HBM latency
// still overlaps the math.  `num += val` with val the per-frame sum, exactly as the reference.
// STORE: the accumulators are INITIALISED with this frame's contribution (0 + r * x, same flush-to-zero addition) instead
// of being read and updated — the first comp frame of a burst then needs no zero-filled accumulators.
template <bool ISO, int VEC, bool STORE = false>
__device__ __forceinline__ void accumulate_thread(const MergeFrame &f, const MergeGeom &g, float *__restrict__ num,
                                                  float *__restrict__ den, int hr_i, int j0) {
    const size_t base = ((size_t)hr_i * g.Ws + j0) * 3;
    const bool full = (VEC == 4) && (j0 + VEC <= g.Ws);
    float n[VEC][3], d[VEC][3], rr[VEC];
    RowCtx rc;
    rc.lr_y = lr_coord(hr_i, g.scale, g.inv_scale, g.pow2);                     // merge.py:319-320
    const int ily = (int)rc.lr_y;
    rc.py = tile_of(ily, g), rc.i_r = min(ily, g.H - 1);                         // merge.py:322-323, 335-336
    rc.tile_x = -1;
    CovQuads cq;
    cq.fx0 = cq.fy0 = -1;
#pragma unroll
    for (int p = 0; p < VEC; ++p) {
        n[p][0] = n[p][1] = n[p][2] = d[p][0] = d[p][1] = d[p][2] = 0.f;
        rr[p] = (j0 + p < g.Ws) ? merge_hr_pixel<ISO>(f, g, j0 + p, rc, cq, n[p], d[p]) : 0.0f;
    }
    if (full) {
        const float *nf = &n[0][0], *df = &d[0][0];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
            if (!STORE) {
                a = *reinterpret_cast<const float4 *>(num + base + 4 * q);
                c = *reinterpret_cast<const float4 *>(den + base + 4 * q);
            }
            a.x = add_ftz(a.x, rr[(4 * q) / 3], nf[4 * q]), a.y = add_ftz(a.y, rr[(4 * q + 1) / 3], nf[4 * q + 1]);
            a.z = add_ftz(a.z, rr[(4 * q + 2) / 3], nf[4 * q + 2]), a.w = add_ftz(a.w, rr[(4 * q + 3) / 3], nf[4 * q + 3]);
            c.x = add_ftz(c.x, rr[(4 * q) / 3], df[4 * q]), c.y = add_ftz(c.y, rr[(4 * q + 1) / 3], df[4 * q + 1]);
            c.z = add_ftz(c.z, rr[(4 * q + 2) / 3], df[4 * q + 2]), c.w = add_ftz(c.w, rr[(4 * q + 3) / 3], df[4 * q + 3]);
            *reinterpret_cast<float4 *>(num + base + 4 * q) = a;
            *reinterpret_cast<float4 *>(den + base + 4 * q) = c;
        }
    } else {
#pragma unroll
        for (int p = 0; p < VEC; ++p)
            if (j0 + p < g.Ws)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    num[base + p * 3 + c] = add_ftz(STORE ? 0.f : num[base + p * 3 + c], rr[p], n[p][c]);
                    den[base + p * 3 + c] = add_ftz(STORE ? 0.f : den[base + p * 3 + c], rr[p], d[p][c]);
                }
    }
}

template <bool ISO, int VEC, bool STORE>
__global__ void __launch_bounds__(256, HHSR_MERGE_MINBLOCKS) accumulate_kernel(MergeFrame f, MergeGeom g, float *__restrict__ num,
                                                                                float *__restrict__ den) {
    const int hr_i = blockIdx.y * blockDim.y + threadIdx.y;
    const int j0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (hr_i >= g.Hs || j0 >= g.Ws) return;
    if (!STORE && VEC == 4 && (threadIdx.x & 1) == 0) {   // 2 threads share 96 B = at most 2 lines per accumulator
        const size_t base = ((size_t)hr_i * g.Ws + j0) * 3;
        prefetch_l2(num + base);
        prefetch_l2(den + base);
        prefetch_l2(num + base + 23);
        prefetch_l2(den + base + 23);
    }
    accumulate_thread<ISO, VEC, STORE>(f, g, num, den, hr_i, j0);
}

// ---------------------------------------------------------------------------------------------------------
// Fast path for scale = 2^K (K = 0, 1, 2: scales 1, 2, 4), Ws % 4 == 0 and a power-of-two tile size >= 4.
// (hr + 0.5) / 2^K = ((2 hr + 1) >> (K+1)) + q with q = ((2 hr + 1) mod 2^(K+1)) / 2^(K+1) a short dyadic fraction, so
// the reference's float64 position  m = lr + flow  splits EXACTLY without float64:  flow = fi + ff (trunc / signed
// fraction, exact in float32),  floor(m) = base + fi + l  with  l = floor(q + ff) in {-1, 0, 1} decided by exact
// comparisons of ff against 1 - q and -q,  and the fraction  t = (q - l) + ff  is one correctly rounded float32 add of
// exactly representable terms — the same value as the reference's float32(m - trunc(m)).  The four pixels of a thread
// share the row, the flow tile and (for K = 1) two x-splits; row pointers, covariance rows and the y-split are
// formed once per thread.  Threads whose 3x3 windows touch the frame border (or leave the frame) take the generic
// per-pixel code, so results are identical to accumulate_kernel.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_q(float q, float ff, int &l, float &t) {
    l = (ff >= 1.0f - q) ? 1 : ((ff < -q) ? -1 : 0);
    t = (q - (float)l) + ff;
}

template <bool ISO, bool STORE>
__device__ __noinline__ void accumulate_thread_border(const MergeFrame *f, const MergeGeom *g, float *num, float *den, int hr_i,
                                                      int j0) {
    accumulate_thread<ISO, 4, STORE>(*f, *g, num, den, hr_i, j0);
}

// p[0..3] += r_k * x_k as ONE fire-and-forget 16-byte reduction performed by the L2 (REDG.E.ADD.F32x4): the SM never
// loads the accumulators, so their HBM/L2 latency is off the warps' critical path.  Each address receives exactly one
// reduction per launch (deterministic).  Like every f32 atomic the L2 adder flushes subnormals (add.ftz); the generic
// kernels use add_ftz() below so that all merge kernels still agree bit for bit.
template <bool STORE>
__device__ __forceinline__ void rmw4(float *__restrict__ p, float r0, float a, float r1, float b, float r2, float c, float r3,
                                     float d) {
    if (STORE) {   // first frame of a burst: initialise instead of accumulate (0 + r * x with the same flush-to-zero add)
        *reinterpret_cast<float4 *>(p) = make_float4(add_ftz(0.f, r0, a), add_ftz(0.f, r1, b), add_ftz(0.f, r2, c), add_ftz(0.f, r3, d));
        return;
    }
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(r0 * a), "f"(r1 * b), "f"(r2 * c), "f"(r3 * d)
                 : "memory");
}

// Interior 3x3 taps addressed by a 32-bit element offset `o` of the centre tap (one IMAD.WIDE per row, immediates
// for the columns); same arithmetic and order as merge_taps<false>.
__device__ __forceinline__ void merge_taps_off(const float *__restrict__ raw, int W, int o, float tx, float ty, float qxx,
                                               float qxy, float qyy, float (&v)[2][2], float (&a)[2][2]) {
#pragma unroll
    for (int di = -1; di <= 1; ++di) {
        const float dy = (float)di + 0.5f - ty;
        const float qy = qyy * dy * dy, qm = qxy * dy;
        const float *row = raw + (o + di * W);
#pragma unroll
        for (int dj = -1; dj <= 1; ++dj) {
            const float c = __ldg(row + dj);
            const float dx = (float)dj + 0.5f - tx;
            float z = fmaf(fmaf(qxx, dx, qm), dx, qy);
            z = fminf(0.0f, z);
            const float w = ex2_approx(z);
            v[di & 1][dj & 1] = fmaf(w, c, v[di & 1][dj & 1]);
            a[di & 1][dj & 1] += w;
        }
    }
}

// Channel resolve for Bayer patterns with green = channel 1: the pattern is RGGB read at phase (py0, px0), so with
// sy = (ci + py0) & 1, sx = (cj + px0) & 1 the relative-parity partial v[ry][rx] sits at RGGB position (ry^sy, rx^sx):
// R = v[sy][sx], B = v[!sy][!sx], G = v[sy][!sx] + v[!sy][sx] (first-row + second-row partial, like resolve_channels).
__device__ __forceinline__ void resolve_rggb(bool sy, bool sx, const float (&v)[2][2], float (&out)[3]) {
    const float t0 = sy ? v[1][0] : v[0][0], t1 = sy ? v[1][1] : v[0][1];   // RGGB row 0 (R G)
    const float b0 = sy ? v[0][0] : v[1][0], b1 = sy ? v[0][1] : v[1][1];   // RGGB row 1 (G B)
    const float R = sx ? t1 : t0, Gt = sx ? t0 : t1, Gb = sx ? b1 : b0, B = sx ? b0 : b1;
    const float G = sy ? (Gb + Gt) : (Gt + Gb);                             // rel row 0 partial first
    out[0] = R, out[1] = G, out[2] = B;
}

// One comp frame, the four pixels of a thread.  `sink.pixel(p, r, val, acc)` receives the unscaled channel sums of pixel p
// and its robustness; returns false (nothing emitted) when the thread has to take the generic per-pixel code: a 3x3
// window touching the frame border or leaving the frame, or a CFA that is not a green-is-1 Bayer pattern.
template <bool ISO, int K, class Sink>
__device__ __forceinline__ bool pow2_frame(const MergeFrame &f, const MergeGeom &g, int by, float qy, int bx0, int tile, Sink &sink) {
    constexpr int SH = K + 1, MASK = (1 << SH) - 1;
    constexpr float INV = 1.0f / (float)(1 << SH);
    const int W = g.W, cw = g.cw;
    const float2 fl = __ldg(reinterpret_cast<const float2 *>(f.flow) + tile);
    const float fiy = truncf(fl.y), fix = truncf(fl.x);
    const float ffy = fl.y - fiy, ffx = fl.x - fix;
    int ly;
    float ty;
    split_q(qy, ffy, ly, ty);
    const int ci = by + (int)fiy + ly;
    const int bx = bx0 + (int)fix;
    int cj[4];
    float tx[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int bp = (2 * p + 1) >> SH;
        const float qp = (float)((2 * p + 1) & MASK) * INV;
        int lx;
        split_q(qp, ffx, lx, tx[p]);
        cj[p] = bx + bp + lx;
    }
    // every 3x3 window strictly inside the frame (cj is non-decreasing in p) and a green-is-1 Bayer pattern?
    if (!(ci >= 1 && ci <= g.H - 2 && cj[0] >= 1 && cj[3] <= W - 2 && g.cfa.bayer && g.cfa.dup == 1)) return false;
    // RGGB phase of the pattern: (py0, px0) = position of channel 0 ... pattern(y, x) = RGGB(y + py0, x + px0)
    const int red_at = ((g.cfa.packed & 3) == 0) ? 0 : (((g.cfa.packed >> 2) & 3) == 0) ? 1 : (((g.cfa.packed >> 4) & 3) == 0) ? 2 : 3;
    const bool sy = ((ci + (red_at >> 1)) & 1) != 0;
    const int px0 = red_at & 1;
    const int orow = ci * W;
    const float *rrow = f.r + (by * W + bx0);
    // covariance rows: inside the frame interior k = m/2 - 0.5 has i0 = (c - 1) >> 1, i1 = i0 + 1 (no clamping) and
    // fraction t/2 (+ 1/2 for even c) — cov_coord() without its border cases
    const int oqy = ((ci - 1) >> 1) * cw;
    const float fry = fmaf(0.5f, ty, (ci & 1) ? 0.0f : 0.5f);
    // the four pixels span at most 3 (K = 0: 4) LR pixels, i.e. quad columns i0 .. i0 + NCOL - 1: blend them along y once
    // (2 NCOL loads instead of 16) and pick the column pair of each pixel
    constexpr int NCOL = (K == 0) ? 4 : 3;
    CovCol col[NCOL];
    const int i0 = (cj[0] - 1) >> 1;
    if (!ISO) {
        const float4 *q0 = reinterpret_cast<const float4 *>(f.covs) + (oqy + i0);
        const float4 *q1 = q0 + cw;
        const int dmax = ((cj[3] - 1) >> 1) - i0;             // column dmax + 1 is the last one any pixel needs (it exists)
        col[0] = cov_column(__ldg(q0), __ldg(q1), fry);
        col[1] = cov_column(__ldg(q0 + 1), __ldg(q1 + 1), fry);
#pragma unroll
        for (int c = 2; c < NCOL; ++c) col[c] = (c <= dmax + 1) ? cov_column(__ldg(q0 + c), __ldg(q1 + c), fry) : col[c - 1];
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int bp = (2 * p + 1) >> SH;
        float qxx, qxy, qyy;
        if (ISO) {
            qxx = qyy = 2.0f * -0.72134752044448170368f, qxy = 0.0f;
        } else {
            const int dcol = ((cj[p] - 1) >> 1) - i0;         // this pixel blends columns (dcol, dcol + 1)
            const float frx = fmaf(0.5f, tx[p], (cj[p] & 1) ? 0.0f : 0.5f);
            CovCol l = col[0], r = col[1];
#pragma unroll
            for (int c = 1; c < NCOL - 1; ++c)
                if (dcol == c) l = col[c], r = col[c + 1];
            cov_form_cols(l, r, frx, qxx, qxy, qyy);
        }
        float v[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, a[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
        merge_taps_off(f.raw, W, orow + cj[p], tx[p], ty, qxx, qxy, qyy, v, a);
        const float rp = __ldg(rrow + bp);
        const bool sx = ((cj[p] + px0) & 1) != 0;
        float val[3], acc[3];
        resolve_rggb(sy, sx, v, val);
        resolve_rggb(sy, sx, a, acc);
        sink.pixel(p, rp, val, acc);
    }
    return true;
}

// Sink of the single-frame kernel: float4 number p-1 of the 12-float slice is complete after pixel p and leaves as one
// 16-byte L2 reduction (or store) per accumulator.
template <bool STORE>
struct RedSink {
    float *num, *den;      // this thread's slices
    float n[12], d[12], rr[4];
    __device__ __forceinline__ void pixel(int p, float r, const float (&val)[3], const float (&acc)[3]) {
        rr[p] = r;
#pragma unroll
        for (int c = 0; c < 3; ++c) n[3 * p + c] = val[c], d[3 * p + c] = acc[c];
        if (p >= 1) {
            const int q = p - 1;
            rmw4<STORE>(num + 4 * q, rr[(4 * q) / 3], n[4 * q], rr[(4 * q + 1) / 3], n[4 * q + 1], rr[(4 * q + 2) / 3], n[4 * q + 2],
                        rr[(4 * q + 3) / 3], n[4 * q + 3]);
            rmw4<STORE>(den + 4 * q, rr[(4 * q) / 3], d[4 * q], rr[(4 * q + 1) / 3], d[4 * q + 1], rr[(4 * q + 2) / 3], d[4 * q + 2],
                        rr[(4 * q + 3) / 3], d[4 * q + 3]);
        }
    }
};

template <bool ISO, int K, bool STORE>
__global__ void __launch_bounds__(256, HHSR_MERGE_POW2_MINBLOCKS) accumulate_pow2_kernel(const __grid_constant__ MergeFrame f,
                                                                                     const __grid_constant__ MergeGeom g,
                                                                                     float *__restrict__ num,
                                                                                     float *__restrict__ den) {
    constexpr int SH = K + 1, MASK = (1 << SH) - 1;
    constexpr float INV = 1.0f / (float)(1 << SH);
    const int hr_i = blockIdx.y * blockDim.y + threadIdx.y;
    const int j0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (hr_i >= g.Hs || j0 >= g.Ws) return;
    const size_t base = ((size_t)hr_i * g.Ws + j0) * 3;
    if (!STORE && (threadIdx.x & 1) == 0) {   // pre-touch the lines the L2 reductions will hit
        prefetch_l2(num + base);
        prefetch_l2(den + base);
        prefetch_l2(num + base + 23);
        prefetch_l2(den + base + 23);
    }
    // y split (shared by the four pixels)
    const int n2 = 2 * hr_i + 1;
    const int by = n2 >> SH;                                   // int(lr_y) <= H - 1
    const float qy = (float)(n2 & MASK) * INV;
    const int bx0 = j0 >> K;                                   // int(lr_x) of pixel 0; all four pixels lie in one tile
    RedSink<STORE> sink;
    sink.num = num + base, sink.den = den + base;
    if (!pow2_frame<ISO, K>(f, g, by, qy, bx0, (by >> g.ts_shift) * g.nx + (bx0 >> g.ts_shift), sink))
        accumulate_thread_border<ISO, STORE>(&f, &g, num, den, hr_i, j0);
}

// ---------------------------------------------------------------------------------------------------------
// Frame-batched fast path: B comp frames in ONE pass over the accumulators.  The thread keeps its 2 x 12 accumulator
// floats in registers (loaded once, or zero for the initialising batch), adds the frames in list order with the same
// flush-to-zero addition as the L2 reduction of the single-frame kernel (bit-identical to B single-frame launches),
// and stores the slice once: accumulator traffic per frame falls from 48 to 48/B (24/B) bytes per HR pixel, which
// moves the merge from the HBM roofline to the instruction-issue limit of the tap arithmetic.
// ---------------------------------------------------------------------------------------------------------
#ifndef HHSR_MERGE_BATCH_MINBLOCKS
#define HHSR_MERGE_BATCH_MINBLOCKS 3
#endif
struct AddSink {
    float n[12], d[12];
    __device__ __forceinline__ void pixel(int p, float r, const float (&val)[3], const float (&acc)[3]) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            n[3 * p + c] = add_ftz(n[3 * p + c], r, val[c]);
            d[3 * p + c] = add_ftz(d[3 * p + c], r, acc[c]);
        }
    }
};

// border threads of the batched kernel: unscaled sums and robustness of the four pixels of one frame through the generic
// per-pixel code, into a scratch array (kept out of line so that the register accumulators never have their address taken)
template <bool ISO>
__device__ __noinline__ void border_frame_sums(const MergeFrame *f, const MergeGeom *g, int hr_i, int j0, float *out) {
    RowCtx rc;
    rc.lr_y = lr_coord(hr_i, g->scale, g->inv_scale, g->pow2);
    const int ily = (int)rc.lr_y;
    rc.py = tile_of(ily, *g), rc.i_r = min(ily, g->H - 1);
    rc.tile_x = -1;
    CovQuads cq;
    cq.fx0 = cq.fy0 = -1;
    for (int p = 0; p < 4; ++p) {
        float val[3] = {0.f, 0.f, 0.f}, acc[3] = {0.f, 0.f, 0.f};
        out[24 + p] = (j0 + p < g->Ws) ? merge_hr_pixel<ISO>(*f, *g, j0 + p, rc, cq, val, acc) : 0.0f;
        for (int c = 0; c < 3; ++c) out[3 * p + c] = val[c], out[12 + 3 * p + c] = acc[c];
    }
}

// FIN: the batch is the LAST one of the burst and the reference frame follows in the same pass — its window sums
// (accumulate_ref, merge.py:82-233) are added to the register accumulators, the quotient num / den (utils.divide) is
// formed and ONLY the finished image is written: num / den never return to HBM (48 B per HR pixel less than a separate
// merge_ref pass, which is also one launch less).  Same per-pixel code and operation order as accumulate_ref_kernel with
// fuse_divide, hence bit-identical to merge + merge_ref.
struct RefFinish {
    const float *raw, *covs;
    float *out;       // [Hs][Ws][3], indexed like num
};
template <bool ISO>
__device__ __forceinline__ bool ref_pixel(const float *__restrict__ raw, const float *__restrict__ covs, const MergeGeom &g, int ox,
                                          int oy, const double *__restrict__ acc_rob, int max_frame_count, int rad_max,
                                          float max_multiplier, float (&val)[3], float (&acc)[3]);

template <bool ISO, int K, bool STORE, bool FIN>
__global__ void __launch_bounds__(256, HHSR_MERGE_BATCH_MINBLOCKS) accumulate_pow2_batch_kernel(const __grid_constant__ MergeBatch b,
                                                                                            const __grid_constant__ MergeGeom g,
                                                                                            float *__restrict__ num,
                                                                                            float *__restrict__ den,
                                                                                            const __grid_constant__ RefFinish fin) {
    constexpr int SH = K + 1, MASK = (1 << SH) - 1;
    constexpr float INV = 1.0f / (float)(1 << SH);
    const int hr_i = g.row_begin + blockIdx.y * blockDim.y + threadIdx.y;
    const int j0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (hr_i >= g.row_end || j0 >= g.Ws) return;
    const size_t base = ((size_t)hr_i * g.Ws + j0) * 3;
    AddSink sink;
    if (STORE) {
#pragma unroll
        for (int q = 0; q < 12; ++q) sink.n[q] = sink.d[q] = 0.f;
    } else {   // issued first: in flight during the first frame's window arithmetic
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const float4 a = *reinterpret_cast<const float4 *>(num + base + 4 * q);
            const float4 c = *reinterpret_cast<const float4 *>(den + base + 4 * q);
            sink.n[4 * q] = a.x, sink.n[4 * q + 1] = a.y, sink.n[4 * q + 2] = a.z, sink.n[4 * q + 3] = a.w;
            sink.d[4 * q] = c.x, sink.d[4 * q + 1] = c.y, sink.d[4 * q + 2] = c.z, sink.d[4 * q + 3] = c.w;
        }
    }
    const int n2 = 2 * hr_i + 1;
    const int by = n2 >> SH;
    const float qy = (float)(n2 & MASK) * INV;
    const int bx0 = j0 >> K;
    const int tile = (by >> g.ts_shift) * g.nx + (bx0 >> g.ts_shift);
#pragma unroll 1
    for (int k = 0; k < b.K; ++k) {
        if (!pow2_frame<ISO, K>(b.f[k], g, by, qy, bx0, tile, sink)) {
            float tmp[28];
            border_frame_sums<ISO>(&b.f[k], &g, hr_i, j0, tmp);
#pragma unroll
            for (int q = 0; q < 12; ++q) {
                sink.n[q] = add_ftz(sink.n[q], tmp[24 + q / 3], tmp[q]);
                sink.d[q] = add_ftz(sink.d[q], tmp[24 + q / 3], tmp[12 + q]);
            }
        }
    }
    if (FIN) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            float val[3], acc[3];
            ref_pixel<ISO>(fin.raw, fin.covs, g, j0 + p, hr_i, nullptr, 0, 0, 0.f, val, acc);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float n_ = sink.n[3 * p + c] + val[c], d_ = sink.d[3 * p + c] + acc[c];     // accumulate_ref_kernel, fuse_divide
                sink.n[3 * p + c] = n_ / d_;
            }
        }
#pragma unroll
        for (int q = 0; q < 3; ++q)
            *reinterpret_cast<float4 *>(fin.out + base + 4 * q) = make_float4(sink.n[4 * q], sink.n[4 * q + 1], sink.n[4 * q + 2], sink.n[4 * q + 3]);
        return;
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        *reinterpret_cast<float4 *>(num + base + 4 * q) = make_float4(sink.n[4 * q], sink.n[4 * q + 1], sink.n[4 * q + 2], sink.n[4 * q + 3]);
        *reinterpret_cast<float4 *>(den + base + 4 * q) = make_float4(sink.d[4 * q], sink.d[4 * q + 1], sink.d[4 * q + 2], sink.d[4 * q + 3]);
    }
}

// K comp frames in one pass over the accumulators (B200 addition).  The slice is loaded first and the frames are
// added in list order, so the result is bit-identical to K single-frame launches.
template <bool ISO, int VEC, bool STORE>
__global__ void __launch_bounds__(256, 2) accumulate_batch_kernel(const __grid_constant__ MergeBatch b, const __grid_constant__ MergeGeom g,
                                                                  float *__restrict__ num, float *__restrict__ den) {
    const int hr_i = g.row_begin + blockIdx.y * blockDim.y + threadIdx.y;
    const int j0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (hr_i >= g.row_end || j0 >= g.Ws) return;
    const size_t base = ((size_t)hr_i * g.Ws + j0) * 3;
    const bool full = (VEC == 4) && (j0 + VEC <= g.Ws);
    float n[VEC * 3], d[VEC * 3];
    if (STORE) {
#pragma unroll
        for (int q = 0; q < VEC * 3; ++q) n[q] = d[q] = 0.f;
    } else if (full) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const float4 a = *reinterpret_cast<const float4 *>(num + base + 4 * q);
            const float4 c = *reinterpret_cast<const float4 *>(den + base + 4 * q);
            n[4 * q] = a.x, n[4 * q + 1] = a.y, n[4 * q + 2] = a.z, n[4 * q + 3] = a.w;
            d[4 * q] = c.x, d[4 * q + 1] = c.y, d[4 * q + 2] = c.z, d[4 * q + 3] = c.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < VEC * 3; ++q) {
            const bool ok = j0 + q / 3 < g.Ws;
            n[q] = ok ? num[base + q] : 0.f;
            d[q] = ok ? den[base + q] : 0.f;
        }
    }
    RowCtx rc;
    rc.lr_y = lr_coord(hr_i, g.scale, g.inv_scale, g.pow2);
    const int ily = (int)rc.lr_y;
    rc.py = tile_of(ily, g), rc.i_r = min(ily, g.H - 1);
    for (int k = 0; k < b.K; ++k) {
        CovQuads cq;
        cq.fx0 = cq.fy0 = -1;
        rc.tile_x = -1;
#pragma unroll
        for (int p = 0; p < VEC; ++p) {
            if (j0 + p >= g.Ws) break;
            float val[3] = {0.f, 0.f, 0.f}, acc[3] = {0.f, 0.f, 0.f};
            const float rl = merge_hr_pixel<ISO>(b.f[k], g, j0 + p, rc, cq, val, acc);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                n[p * 3 + c] = add_ftz(n[p * 3 + c], rl, val[c]);
                d[p * 3 + c] = add_ftz(d[p * 3 + c], rl, acc[c]);
            }
        }
    }
    if (full) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            *reinterpret_cast<float4 *>(num + base + 4 * q) = make_float4(n[4 * q], n[4 * q + 1], n[4 * q + 2], n[4 * q + 3]);
            *reinterpret_cast<float4 *>(den + base + 4 * q) = make_float4(d[4 * q], d[4 * q + 1], d[4 * q + 2], d[4 * q + 3]);
        }
    } else {
#pragma unroll
        for (int q = 0; q < VEC * 3; ++q)
            if (j0 + q / 3 < g.Ws) {
                num[base + q] = n[q];
                den[base + q] = d[q];
            }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Reference frame (merge.py:82-233).  The reference rounds the HR->LR position to float32 (`coarse_ref_sub_pos`
// is a float32 local array), so tap offsets `tap - pos` and the covariance-map coordinate (pos - 0.5)/2 are exact
// in float32; the remaining arithmetic (bilinear Omega, inverse, quadratic form, exp) is float32 here where the
// reference promotes to float64 — float32-rounding-level differences, pinned by the goldens.
// rows [row_begin, row_end) of the output are processed (frame-sharded runs normalise row slices).
// ---------------------------------------------------------------------------------------------------------
// Up to 16 peer accumulator pairs (frame-sharded runs): the kernel sums them in rank order while it normalises its
// row slice, reading the peers' HBM over NVLink (peer-mapped pointers) — the reduction, merge_ref and divide in one pass.
constexpr int kMaxPeers = 16;
struct RefPeers {
    const float *num[kMaxPeers], *den[kMaxPeers];
    int n;             // 0: accumulate into (num, den) in place
};

// weight of one tap.  The reference's float32 accumulators keep weights down to the subnormal range (1e-45) and a
// channel fed only by far taps of a narrow kernel is normalised from exactly those: evaluate them in float64.
__device__ __noinline__ float tiny_weight(float z) { return (float)exp2((double)z); }
__device__ __forceinline__ float ref_weight(float z) { return (z > -120.f) ? ex2_approx(z) : tiny_weight(z); }

// One HR pixel of the reference frame: per-channel sums val/acc of its window; returns whether the accumulators are
// to be overwritten (accumulated-robustness denoiser, merge.py:223-233).
template <bool ISO>
__device__ __forceinline__ bool ref_pixel(const float *__restrict__ raw, const float *__restrict__ covs, const MergeGeom &g, int ox,
                                          int oy, const double *__restrict__ acc_rob, int max_frame_count, int rad_max,
                                          float max_multiplier, float (&val)[3], float (&acc)[3]) {
    // :113-114; for power-of-two scales the multiplication by the exact reciprocal is the same correctly-rounded value
    const float pos_y = g.pow2 ? (float)((double)oy * g.inv_scale) : (float)((double)oy / g.scale);
    const float pos_x = g.pow2 ? (float)((double)ox * g.inv_scale) : (float)((double)ox / g.scale);
    const float kS = -0.72134752044448170368f;   // -0.5 * log2(e)
    float qxx, qxy, qyy;
    if (ISO) {
        qxx = qyy = 2.0f * kS, qxy = 0.f;                                                        // :211
    } else {
        const float gy = (pos_y - 0.5f) * 0.5f, gx = (pos_x - 0.5f) * 0.5f;                      // :132-133 (exact)
        const int fx0 = (int)fmaxf(floorf(gx), 0.f), fy0 = (int)fmaxf(floorf(gy), 0.f);
        const int cx1 = min(fx0 + 1, g.cw - 1), cy1 = min(fy0 + 1, g.ch - 1);
        const float rx = gx - truncf(gx), ry = gy - truncf(gy);                                  // linalg.py:190-191
        const float4 *c4 = reinterpret_cast<const float4 *>(covs);
        const float4 c00 = __ldg(c4 + (fy0 * g.cw + fx0)), c01 = __ldg(c4 + (fy0 * g.cw + cx1));
        const float4 c10 = __ldg(c4 + (cy1 * g.cw + fx0)), c11 = __ldg(c4 + (cy1 * g.cw + cx1));
        const float w00 = (1.f - rx) * (1.f - ry), w01 = rx * (1.f - ry), w10 = (1.f - rx) * ry, w11 = rx * ry;
        const float m00 = fmaf(c11.x, w11, fmaf(c10.x, w10, fmaf(c01.x, w01, c00.x * w00)));
        const float m01 = fmaf(c11.y, w11, fmaf(c10.y, w10, fmaf(c01.y, w01, c00.y * w00)));
        const float m10 = fmaf(c11.z, w11, fmaf(c10.z, w10, fmaf(c01.z, w01, c00.z * w00)));
        const float m11 = fmaf(c11.w, w11, fmaf(c10.w, w10, fmaf(c01.w, w01, c00.w * w00)));
        const float det = __fmaf_rn(m00, m11, -__fmul_rn(m01, m10));                             // linalg.py:53
        float i00 = 1.f, i0110 = 0.f, i11 = 1.f;
        if (fabsf(det) > 1e-10f) {                                                               // EPSILON_DIV
            const float det_i = 1.0f / det;
            i00 = m11 * det_i, i0110 = -(m01 + m10) * det_i, i11 = m00 * det_i;
        }
        qxx = kS * i00, qxy = kS * i0110, qyy = kS * i11;                                        // linalg.py:83
    }
    int rad = 1;
    bool overwrite = false;
    if (acc_rob != nullptr) {                                                                    // :167-176
        const int ay = min((int)rintf(pos_y), g.H - 1), ax = min((int)rintf(pos_x), g.W - 1);
        const double la = __ldg(acc_rob + (size_t)ay * g.W + ax);
        if (la <= (double)max_frame_count) {
            const float inv_p = 1.0f / max_multiplier;                                           // y /= power, :218
            qxx *= inv_p, qxy *= inv_p, qyy *= inv_p;
            rad = rad_max;
        }
        overwrite = la < (double)max_frame_count;
    }
    const int cx = (int)rintf(pos_x), cy = (int)rintf(pos_y);                                    // round half even
    val[0] = val[1] = val[2] = acc[0] = acc[1] = acc[2] = 0.f;
    if (rad == 1 && cy >= 1 && cy <= g.H - 2 && cx >= 1 && cx <= g.W - 2) {
        // interior 3x3 window: no bounds tests, partial sums per tap parity relative to the centre (compile-time
        // indices), CFA channel of each partial resolved once
        float v[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, a[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
        const int o = cy * g.W + cx;
#pragma unroll
        for (int i = -1; i <= 1; ++i) {
            const float dy = (float)(cy + i) - pos_y;
            const float qy = qyy * dy * dy, qm = qxy * dy;
            const float *row = raw + (o + i * g.W);
#pragma unroll
            for (int j = -1; j <= 1; ++j) {
                const float c = __ldg(row + j);
                const float dx = (float)(cx + j) - pos_x;
                const float z = fminf(0.f, fmaf(fmaf(qxx, dx, qm), dx, qy));                     // max(0, y): NaN -> 0
                const float w = ref_weight(z);
                v[i & 1][j & 1] = fmaf(c, w, v[i & 1][j & 1]);
                a[i & 1][j & 1] += w;
            }
        }
#pragma unroll
        for (int ry_ = 0; ry_ < 2; ++ry_)
#pragma unroll
            for (int rx_ = 0; rx_ < 2; ++rx_) {
                const int chn = cfa_channel(g.cfa.packed, cy + ry_, cx + rx_);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    val[k] += (chn == k) ? v[ry_][rx_] : 0.f;
                    acc[k] += (chn == k) ? a[ry_][rx_] : 0.f;
                }
            }
        return overwrite;
    }
    for (int i = -rad; i <= rad; ++i) {
        const int yy = cy + i;
        if (yy < 0 || yy >= g.H) continue;
        const float dy = (float)yy - pos_y;
        const float qy = qyy * dy * dy, qm = qxy * dy;
        for (int j = -rad; j <= rad; ++j) {
            const int xx = cx + j;
            if (xx < 0 || xx >= g.W) continue;
            const int chn = cfa_channel(g.cfa.packed, yy, xx);
            const float c = __ldg(raw + (yy * g.W + xx));
            const float dx = (float)xx - pos_x;
            const float z = fminf(0.f, fmaf(fmaf(qxx, dx, qm), dx, qy));                         // max(0, y): NaN -> 0
            const float w = ref_weight(z);
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (chn == k) {
                    val[k] = fmaf(c, w, val[k]);
                    acc[k] += w;
                }
        }
    }
    return overwrite;
}

// VEC consecutive HR pixels of a row per thread (VEC = 4: float4 accesses of the 12-float slice).  Output:
//   peers.n == 0 : num <- (num + val) [/ (den + acc) when fuse_divide], den <- den + acc, in place (merge.py:223-233);
//   peers.n  > 0 : out_num <- (sum_p peers.num[p] + val) [/ ...]; den is not written (nobody reads it afterwards).
template <bool ISO, int VEC>
__global__ void __launch_bounds__(256) accumulate_ref_kernel(const float *__restrict__ raw, const float *__restrict__ covs,
                                                             const __grid_constant__ MergeGeom g, float *num, float *den,
                                                             const double *__restrict__ acc_rob, int max_frame_count,
                                                             int rad_max, float max_multiplier, int fuse_divide, int row_begin,
                                                             int row_end, const __grid_constant__ RefPeers peers, float *out_num) {
    const int ox0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC, oy = row_begin + blockIdx.y * blockDim.y + threadIdx.y;
    if (ox0 >= g.Ws || oy >= row_end) return;
    const size_t o = ((size_t)oy * g.Ws + ox0) * 3;
    float nn[VEC * 3], dd[VEC * 3];
    // running sums of the accumulators first: the (remote) loads are in flight while the window math runs
    if (peers.n == 0) {
        if (VEC == 4) {
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const float4 a = *reinterpret_cast<const float4 *>(num + o + 4 * q), c = *reinterpret_cast<const float4 *>(den + o + 4 * q);
                nn[4 * q] = a.x, nn[4 * q + 1] = a.y, nn[4 * q + 2] = a.z, nn[4 * q + 3] = a.w;
                dd[4 * q] = c.x, dd[4 * q + 1] = c.y, dd[4 * q + 2] = c.z, dd[4 * q + 3] = c.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) nn[k] = num[o + k], dd[k] = den[o + k];
        }
    } else {
#pragma unroll
        for (int k = 0; k < VEC * 3; ++k) nn[k] = dd[k] = 0.f;
        for (int p = 0; p < peers.n; ++p) {
            const float *pn = peers.num[p] + o, *pd = peers.den[p] + o;
            if (VEC == 4) {
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const float4 a = *reinterpret_cast<const float4 *>(pn + 4 * q), c = *reinterpret_cast<const float4 *>(pd + 4 * q);
                    nn[4 * q] += a.x, nn[4 * q + 1] += a.y, nn[4 * q + 2] += a.z, nn[4 * q + 3] += a.w;
                    dd[4 * q] += c.x, dd[4 * q + 1] += c.y, dd[4 * q + 2] += c.z, dd[4 * q + 3] += c.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 3; ++k) nn[k] += pn[k], dd[k] += pd[k];
            }
        }
    }
#pragma unroll
    for (int p = 0; p < VEC; ++p) {
        float val[3], acc[3];
        const bool overwrite = ref_pixel<ISO>(raw, covs, g, ox0 + p, oy, acc_rob, max_frame_count, rad_max, max_multiplier, val, acc);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float n_ = overwrite ? val[k] : nn[3 * p + k] + val[k];
            const float d_ = overwrite ? acc[k] : dd[3 * p + k] + acc[k];
            if (fuse_divide) n_ = n_ / d_;
            nn[3 * p + k] = n_, dd[3 * p + k] = d_;
        }
    }
    float *on = peers.n == 0 ? num : out_num;
    if (VEC == 4) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            *reinterpret_cast<float4 *>(on + o + 4 * q) = make_float4(nn[4 * q], nn[4 * q + 1], nn[4 * q + 2], nn[4 * q + 3]);
            if (peers.n == 0)
                *reinterpret_cast<float4 *>(den + o + 4 * q) = make_float4(dd[4 * q], dd[4 * q + 1], dd[4 * q + 2], dd[4 * q + 3]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            on[o + k] = nn[k];
            if (peers.n == 0) den[o + k] = dd[k];
        }
    }
}

__global__ void divide_kernel(float *__restrict__ num, const float *__restrict__ den, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 4 <= n) {
            float4 a = *reinterpret_cast<float4 *>(num + i);
            const float4 b = *reinterpret_cast<const float4 *>(den + i);
            a.x /= b.x, a.y /= b.y, a.z /= b.z, a.w /= b.w;
            *reinterpret_cast<float4 *>(num + i) = a;
        } else {
            for (size_t k = i; k < n; ++k) num[k] /= den[k];
        }
    }
}

// A += B_0 + B_1 + ... in list order (float64 accumulation): K x utils.add (utils.py:93-120) in one pass over A
struct AddList {
    const float *b[kMaxBatch];
    int K;
};
__global__ void add_many_f64_f32_kernel(double *__restrict__ A, AddList l, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double a = A[i];
        for (int k = 0; k < l.K; ++k) a += (double)__ldg(l.b[k] + i);
        A[i] = a;
    }
}

__global__ void add_f64_f32_kernel(double *__restrict__ A, const float *__restrict__ B, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) A[i] += (double)B[i];
}

static int check_merge_args(const void *raw, const void *num, const void *den, int H, int W, int Hs, int Ws,
                            double scale, const int *cfa) {
    HHSR_REQUIRE(raw && num && den && cfa, "null pointer");
    HHSR_REQUIRE(H > 1 && W > 1 && Hs > 0 && Ws > 0, "non-positive size");
    HHSR_REQUIRE(scale >= 1.0, "scale must be >= 1 (params.py:8)");
    HHSR_REQUIRE(((uintptr_t)num % 16 == 0) && ((uintptr_t)den % 16 == 0), "num/den must be 16-byte aligned");
    return 0;
}

template <int VEC, bool STORE>
static void launch_accumulate_vec(const MergeBatch &b, const MergeGeom &g, float *num, float *den, int iso, dim3 grid,
                                  dim3 block, cudaStream_t st) {
    if (b.K == 1 && g.row_begin == 0 && g.row_end == g.Hs) {
        if (iso)
            accumulate_kernel<true, VEC, STORE><<<grid, block, 0, st>>>(b.f[0], g, num, den);
        else
            accumulate_kernel<false, VEC, STORE><<<grid, block, 0, st>>>(b.f[0], g, num, den);
    } else {
        if (iso)
            accumulate_batch_kernel<true, VEC, STORE><<<grid, block, 0, st>>>(b, g, num, den);
        else
            accumulate_batch_kernel<false, VEC, STORE><<<grid, block, 0, st>>>(b, g, num, den);
    }
}

// scale 1, 2 or 4 with the geometry the fast path assumes (-1: the any-scale kernel has to run)
static int pow2_fast_shift(const MergeGeom &g) {
    if (!g.pow2 || g.ts_shift < 2 || g.Ws % 4 != 0) return -1;
    for (int k = 0; k <= 2; ++k)
        if (g.scale == (double)(1 << k) && g.Hs == (g.H << k) && g.Ws == (g.W << k)) return k;
    return -1;
}

template <int K, bool STORE>
static void launch_pow2(const MergeBatch &b, const MergeGeom &g, float *num, float *den, int iso, dim3 grid, dim3 block,
                        cudaStream_t st, const RefFinish *fin) {
    if (fin != nullptr) {
        if (iso)
            accumulate_pow2_batch_kernel<true, K, STORE, true><<<grid, block, 0, st>>>(b, g, num, den, *fin);
        else
            accumulate_pow2_batch_kernel<false, K, STORE, true><<<grid, block, 0, st>>>(b, g, num, den, *fin);
        return;
    }
    const RefFinish none{nullptr, nullptr, nullptr};
    if (b.K == 1 && g.row_begin == 0 && g.row_end == g.Hs) {
        if (iso)
            accumulate_pow2_kernel<true, K, STORE><<<grid, block, 0, st>>>(b.f[0], g, num, den);
        else
            accumulate_pow2_kernel<false, K, STORE><<<grid, block, 0, st>>>(b.f[0], g, num, den);
    } else {
        if (iso)
            accumulate_pow2_batch_kernel<true, K, STORE, false><<<grid, block, 0, st>>>(b, g, num, den, none);
        else
            accumulate_pow2_batch_kernel<false, K, STORE, false><<<grid, block, 0, st>>>(b, g, num, den, none);
    }
}

template <bool STORE>
static int launch_accumulate(const MergeBatch &b, const MergeGeom &g, float *num, float *den, int iso, bool generic,
                             cudaStream_t st, const RefFinish *fin = nullptr) {
    dim3 block(32, 8);
    const int k = generic ? -1 : pow2_fast_shift(g);
    const int rows = g.row_end - g.row_begin;
    if (k >= 0) {
        dim3 grid(ceil_div(g.Ws, 32 * 4), ceil_div(rows, 8));
        if (k == 0) launch_pow2<0, STORE>(b, g, num, den, iso, grid, block, st, fin);
        if (k == 1) launch_pow2<1, STORE>(b, g, num, den, iso, grid, block, st, fin);
        if (k == 2) launch_pow2<2, STORE>(b, g, num, den, iso, grid, block, st, fin);
        return launch_status("merge_accumulate");
    }
    if (fin != nullptr) return unsupported("the fused finish needs the power-of-two fast path (scale 1, 2 or 4, output width a multiple of 4)");
    if (g.Ws % 4 == 0)
        launch_accumulate_vec<4, STORE>(b, g, num, den, iso, dim3(ceil_div(g.Ws, 32 * 4), ceil_div(rows, 8)), block, st);
    else
        launch_accumulate_vec<1, STORE>(b, g, num, den, iso, dim3(ceil_div(g.Ws, 32), ceil_div(rows, 8)), block, st);
    return launch_status("merge_accumulate");
}

}  // namespace hhsr

using namespace hhsr;

static int merge_frames(const float *const *raws, const float *const *flows, const float *const *covs, const float *const *rs,
                        int K, int H, int W, int ny, int nx, int ts, float *num, float *den, int Hs, int Ws, double scale,
                        const int *cfa_host, int iso, int flags, int row_begin, int row_end, hhsr_stream_t stream,
                        const RefFinish *fin = nullptr) {
    HHSR_REQUIRE((flags & ~(HHSR_MERGE_INIT | HHSR_MERGE_GENERIC)) == 0, "unknown merge flag");
    HHSR_REQUIRE(fin == nullptr || !(flags & HHSR_MERGE_GENERIC), "the fused finish runs on the fast path only");
    HHSR_REQUIRE(0 <= row_begin && row_begin < row_end && row_end <= Hs, "row range must satisfy 0 <= begin < end <= Hs");
    const bool generic = (flags & HHSR_MERGE_GENERIC) != 0;
    HHSR_REQUIRE(raws && flows && rs && K > 0, "null frame list");
    HHSR_REQUIRE(iso || covs, "covs required for the steerable kernel");
    if (int e = check_merge_args(raws[0], num, den, H, W, Hs, Ws, scale, cfa_host)) return e;
    HHSR_REQUIRE(ts > 0 && ny * ts >= H && nx * ts >= W, "flow grid does not cover the frame");
    MergeGeom g = make_geom(H, W, nx, ts, Hs, Ws, cfa_host, scale);
    g.row_begin = row_begin, g.row_end = row_end;
    // num / den point at row `row_begin` (a caller may own only that slice); the kernels index rows absolutely
    num -= (size_t)row_begin * Ws * 3, den -= (size_t)row_begin * Ws * 3;
    RefFinish finish{nullptr, nullptr, nullptr};
    if (fin != nullptr) {
        HHSR_REQUIRE(fin->raw && fin->out && (iso || fin->covs), "fused finish: null reference frame / covariances / output");
        HHSR_REQUIRE((uintptr_t)fin->out % 16 == 0 && (iso || (uintptr_t)fin->covs % 16 == 0), "fused finish: misaligned buffers");
        HHSR_REQUIRE(pow2_fast_shift(g) >= 0, "the fused finish needs the power-of-two fast path (scale 1, 2 or 4, output width % 4 == 0)");
        finish = RefFinish{fin->raw, iso ? nullptr : fin->covs, fin->out - (size_t)row_begin * Ws * 3};
    }
    for (int k0 = 0; k0 < K; k0 += kMaxBatch) {
        MergeBatch b;
        b.K = (K - k0 < kMaxBatch) ? K - k0 : kMaxBatch;
        for (int k = 0; k < b.K; ++k) {
            HHSR_REQUIRE(raws[k0 + k] && flows[k0 + k] && rs[k0 + k], "null frame pointer");
            HHSR_REQUIRE(iso || ((uintptr_t)covs[k0 + k] % 16 == 0 && covs[k0 + k]), "covs must be 16-byte aligned");
            b.f[k] = MergeFrame{raws[k0 + k], flows[k0 + k], iso ? nullptr : covs[k0 + k], rs[k0 + k]};
        }
        // only the first chunk of an initialising call stores; later chunks accumulate onto it
        const bool store = (flags & HHSR_MERGE_INIT) != 0 && k0 == 0;
        const RefFinish *f = (fin != nullptr && k0 + kMaxBatch >= K) ? &finish : nullptr;     // the last chunk finishes
        const int e = store ? launch_accumulate<true>(b, g, num, den, iso, generic, (cudaStream_t)stream, f)
                            : launch_accumulate<false>(b, g, num, den, iso, generic, (cudaStream_t)stream, f);
        if (e) return e;
    }
    return 0;
}

extern "C" int hhsr_merge_accumulate_batch(const float *const *raws, const float *const *flows,
                                           const float *const *covs, const float *const *rs, int K, int H, int W,
                                           int ny, int nx, int ts, float *num, float *den, int Hs, int Ws,
                                           double scale, const int *cfa_host, int iso, int flags, hhsr_stream_t stream) {
    return merge_frames(raws, flows, covs, rs, K, H, W, ny, nx, ts, num, den, Hs, Ws, scale, cfa_host, iso, flags, 0, Hs, stream);
}

extern "C" int hhsr_merge_accumulate_rows(const float *const *raws, const float *const *flows, const float *const *covs,
                                          const float *const *rs, int K, int H, int W, int ny, int nx, int ts,
                                          float *num_rows, float *den_rows, int Hs, int Ws, double scale, const int *cfa_host,
                                          int iso, int flags, int row_begin, int row_end, hhsr_stream_t stream) {
    return merge_frames(raws, flows, covs, rs, K, H, W, ny, nx, ts, num_rows, den_rows, Hs, Ws, scale, cfa_host, iso, flags,
                        row_begin, row_end, stream);
}

extern "C" int hhsr_merge_finish_rows(const float *const *raws, const float *const *flows, const float *const *covs,
                                      const float *const *rs, int K, int H, int W, int ny, int nx, int ts, float *num_rows,
                                      float *den_rows, int Hs, int Ws, double scale, const int *cfa_host, int iso, int flags,
                                      int row_begin, int row_end, const float *ref_raw, const float *ref_covs, float *out_rows,
                                      hhsr_stream_t stream) {
    RefFinish fin{ref_raw, ref_covs, out_rows};
    return merge_frames(raws, flows, covs, rs, K, H, W, ny, nx, ts, num_rows, den_rows, Hs, Ws, scale, cfa_host, iso, flags,
                        row_begin, row_end, stream, &fin);
}

extern "C" int hhsr_merge_accumulate(const float *raw, int H, int W, const float *flow, int ny, int nx, int ts,
                                     const float *covs, const float *r, float *num, float *den, int Hs, int Ws,
                                     double scale, const int *cfa_host, int iso, hhsr_stream_t stream) {
    return merge_frames(&raw, &flow, &covs, &r, 1, H, W, ny, nx, ts, num, den, Hs, Ws, scale, cfa_host, iso, 0, 0, Hs, stream);
}

extern "C" int hhsr_merge_init_accumulate(const float *raw, int H, int W, const float *flow, int ny, int nx, int ts,
                                          const float *covs, const float *r, float *num, float *den, int Hs, int Ws,
                                          double scale, const int *cfa_host, int iso, hhsr_stream_t stream) {
    return merge_frames(&raw, &flow, &covs, &r, 1, H, W, ny, nx, ts, num, den, Hs, Ws, scale, cfa_host, iso, HHSR_MERGE_INIT, 0, Hs, stream);
}

static int launch_merge_ref(const float *raw, const float *covs, const MergeGeom &g, float *num, float *den, int iso,
                            const double *acc_rob, int max_frame_count, int rad_max, double max_multiplier, int fuse_divide,
                            int row_begin, int row_end, const RefPeers &peers, float *out_num, cudaStream_t st) {
    dim3 block(32, 8);
    const float mm = (float)max_multiplier;
    if (g.Ws % 4 == 0) {
        dim3 grid(ceil_div(g.Ws, 32 * 4), ceil_div(row_end - row_begin, 8));
        if (iso)
            accumulate_ref_kernel<true, 4><<<grid, block, 0, st>>>(raw, covs, g, num, den, acc_rob, max_frame_count, rad_max, mm,
                                                                   fuse_divide, row_begin, row_end, peers, out_num);
        else
            accumulate_ref_kernel<false, 4><<<grid, block, 0, st>>>(raw, covs, g, num, den, acc_rob, max_frame_count, rad_max, mm,
                                                                    fuse_divide, row_begin, row_end, peers, out_num);
    } else {
        dim3 grid(ceil_div(g.Ws, 32), ceil_div(row_end - row_begin, 8));
        if (iso)
            accumulate_ref_kernel<true, 1><<<grid, block, 0, st>>>(raw, covs, g, num, den, acc_rob, max_frame_count, rad_max, mm,
                                                                   fuse_divide, row_begin, row_end, peers, out_num);
        else
            accumulate_ref_kernel<false, 1><<<grid, block, 0, st>>>(raw, covs, g, num, den, acc_rob, max_frame_count, rad_max, mm,
                                                                    fuse_divide, row_begin, row_end, peers, out_num);
    }
    return launch_status("merge_ref");
}

extern "C" int hhsr_merge_ref(const float *raw, int H, int W, const float *covs, float *num, float *den, int Hs,
                              int Ws, double scale, const int *cfa_host, int iso, const double *acc_rob,
                              int max_frame_count, int rad_max, double max_multiplier, int fuse_divide, int row_begin,
                              int row_end, hhsr_stream_t stream) {
    if (int e = check_merge_args(raw, num, den, H, W, Hs, Ws, scale, cfa_host)) return e;
    HHSR_REQUIRE(iso || (covs && (uintptr_t)covs % 16 == 0), "covs required (16-byte aligned) for the steerable kernel");
    HHSR_REQUIRE(acc_rob == nullptr || (rad_max >= 0 && max_multiplier > 0.0), "rad_max >= 0 and max_multiplier > 0 required");
    HHSR_REQUIRE(0 <= row_begin && row_begin < row_end && row_end <= Hs, "row range must satisfy 0 <= begin < end <= Hs");
    MergeGeom g = make_geom(H, W, 0, 1, Hs, Ws, cfa_host, scale);
    RefPeers peers;
    peers.n = 0;
    return launch_merge_ref(raw, covs, g, num, den, iso, acc_rob, max_frame_count, rad_max, max_multiplier, fuse_divide,
                            row_begin, row_end, peers, nullptr, (cudaStream_t)stream);
}

extern "C" int hhsr_merge_ref_rows(const float *raw, int H, int W, const float *covs, float *num_rows, float *den_rows, int Hs, int Ws,
                                   double scale, const int *cfa_host, int iso, const double *acc_rob, int max_frame_count,
                                   int rad_max, double max_multiplier, int fuse_divide, int row_begin, int row_end,
                                   hhsr_stream_t stream) {
    HHSR_REQUIRE(num_rows && den_rows, "null pointer");
    HHSR_REQUIRE(0 <= row_begin && row_begin < row_end && row_end <= Hs && Ws > 0, "row range must satisfy 0 <= begin < end <= Hs");
    // the slices start at row `row_begin`; the kernel indexes rows absolutely
    const size_t off = (size_t)row_begin * Ws * 3;
    return hhsr_merge_ref(raw, H, W, covs, num_rows - off, den_rows - off, Hs, Ws, scale, cfa_host, iso, acc_rob, max_frame_count,
                          rad_max, max_multiplier, fuse_divide, row_begin, row_end, stream);
}

extern "C" int hhsr_reduce_merge_ref(const float *const *peer_nums, const float *const *peer_dens, int n_peers,
                                     const float *raw, int H, int W, const float *covs, float *out_num, int Hs, int Ws,
                                     double scale, const int *cfa_host, int iso, const double *acc_rob,
                                     int max_frame_count, int rad_max, double max_multiplier, int fuse_divide,
                                     int row_begin, int row_end, hhsr_stream_t stream) {
    HHSR_REQUIRE(peer_nums && peer_dens && n_peers >= 1 && n_peers <= kMaxPeers, "1 to 16 peer accumulator pairs required");
    if (int e = check_merge_args(raw, out_num, out_num, H, W, Hs, Ws, scale, cfa_host)) return e;
    HHSR_REQUIRE(iso || (covs && (uintptr_t)covs % 16 == 0), "covs required (16-byte aligned) for the steerable kernel");
    HHSR_REQUIRE(acc_rob == nullptr || (rad_max >= 0 && max_multiplier > 0.0), "rad_max >= 0 and max_multiplier > 0 required");
    HHSR_REQUIRE(0 <= row_begin && row_begin < row_end && row_end <= Hs, "row range must satisfy 0 <= begin < end <= Hs");
    MergeGeom g = make_geom(H, W, 0, 1, Hs, Ws, cfa_host, scale);
    RefPeers peers;
    peers.n = n_peers;
    for (int p = 0; p < n_peers; ++p) {
        HHSR_REQUIRE(peer_nums[p] && peer_dens[p], "null peer pointer");
        HHSR_REQUIRE((uintptr_t)peer_nums[p] % 16 == 0 && (uintptr_t)peer_dens[p] % 16 == 0, "peer accumulators must be 16-byte aligned");
        peers.num[p] = peer_nums[p], peers.den[p] = peer_dens[p];
    }
    return launch_merge_ref(raw, covs, g, nullptr, nullptr, iso, acc_rob, max_frame_count, rad_max, max_multiplier, fuse_divide,
                            row_begin, row_end, peers, out_num, (cudaStream_t)stream);
}

extern "C" int hhsr_divide(float *num, const float *den, size_t n, hhsr_stream_t stream) {
    HHSR_REQUIRE(num && den && n > 0, "null pointer or empty array");
    HHSR_REQUIRE(((uintptr_t)num % 16 == 0) && ((uintptr_t)den % 16 == 0), "num/den must be 16-byte aligned");
    const int block = 256;
    size_t blocks = (n / 4 + block - 1) / block + 1;
    if (blocks > 148 * 32) blocks = 148 * 32;
    divide_kernel<<<(unsigned)blocks, block, 0, (cudaStream_t)stream>>>(num, den, n);
    return launch_status("divide");
}

extern "C" int hhsr_add_many_f64_f32(double *A, const float *const *Bs, int K, size_t n, hhsr_stream_t stream) {
    HHSR_REQUIRE(A && Bs && K > 0 && n > 0, "null pointer or empty list");
    const int block = 256;
    size_t blocks = (n + block - 1) / block;
    if (blocks > 148 * 16) blocks = 148 * 16;
    for (int k0 = 0; k0 < K; k0 += kMaxBatch) {
        AddList l;
        l.K = (K - k0 < kMaxBatch) ? K - k0 : kMaxBatch;
        for (int k = 0; k < l.K; ++k) {
            HHSR_REQUIRE(Bs[k0 + k], "null array in list");
            l.b[k] = Bs[k0 + k];
        }
        add_many_f64_f32_kernel<<<(unsigned)blocks, block, 0, (cudaStream_t)stream>>>(A, l, n);
    }
    return launch_status("add_many");
}

extern "C" int hhsr_add_f64_f32(double *A, const float *B, size_t n, hhsr_stream_t stream) {
    HHSR_REQUIRE(A && B && n > 0, "null pointer or empty array");
    const int block = 256;
    size_t blocks = (n + block - 1) / block;
    if (blocks > 148 * 32) blocks = 148 * 32;
    add_f64_f32_kernel<<<(unsigned)blocks, block, 0, (cudaStream_t)stream>>>(A, B, n);
    return launch_status("add");
}
