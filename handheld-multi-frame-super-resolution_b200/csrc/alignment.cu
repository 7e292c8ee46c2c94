// Coarse-to-fine alignment kernels for sm_100a: gradients + tile Hessians, flow upscaling, L2 block matching,
// the "L1" level as the compiled reference executes it, and inverse-compositional Lucas-Kanade (ICA).
//
// Replaces handheld_super_resolution/ICA.py:15-76 (init_ica: two F.conv2d + compute_hessian, one thread per
// tile), alignment.py:150-172 (upscale_lvl), block_matching.py:20-76,348-378 (gather + batched rFFT/irFFT +
// box-filter conv2d + argmin), block_matching.py:78-345 (cuda_L1_local_search*) and ICA.py:78-481
// (ica_kernel_{8,16,32,64}).  One CTA per tile; tiles and search windows are staged in shared memory, SSD sums
// are reduced with warp shuffles.
#include "common.cuh"
#include <atomic>

namespace hhsr {

// Opt-in to > 48 KB of dynamic shared memory: a per-device function attribute, set once per (kernel, device) instead of
// on every launch.  The only process-wide state of the library besides the last-error string: a bit per device.
template <class Kernel>
static void ensure_dynamic_smem(Kernel kernel, std::atomic<unsigned long long> &done, size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_relaxed) & bit) return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 > bytes ? (int)(200 * 1024) : (int)bytes);
    done.fetch_or(bit, std::memory_order_relaxed);
}

// ---------------------------------------------------------------------------------------------------------
// gradients + Hessian: one CTA per ts x ts cell of ceil(h/ts) x ceil(w/ts); complete tiles also emit H.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grad_hessian_kernel(const float *__restrict__ img, int h, int w, int ts, int ny,
                                                           int nx, float *__restrict__ gradx, float *__restrict__ grady,
                                                           float *__restrict__ hessian) {
    const int tx = blockIdx.x, ty = blockIdx.y;
    float h00 = 0.f, h01 = 0.f, h11 = 0.f;
    for (int p = threadIdx.x; p < ts * ts; p += blockDim.x) {
        const int y = ty * ts + p / ts, x = tx * ts + p % ts;
        if (y >= h || x >= w) continue;
        const float l = x > 0 ? __ldg(img + (size_t)y * w + x - 1) : 0.f;       // zero 'same' padding, ICA.py:20-21
        const float r = x < w - 1 ? __ldg(img + (size_t)y * w + x + 1) : 0.f;
        const float u = y > 0 ? __ldg(img + (size_t)(y - 1) * w + x) : 0.f;
        const float d = y < h - 1 ? __ldg(img + (size_t)(y + 1) * w + x) : 0.f;
        const float gx = r - l, gy = d - u;
        gradx[(size_t)y * w + x] = gx;
        grady[(size_t)y * w + x] = gy;
        h00 = fmaf(gx, gx, h00), h01 = fmaf(gx, gy, h01), h11 = fmaf(gy, gy, h11);
    }
    if (ty >= ny || tx >= nx) return;   // uniform per CTA
    __shared__ float red[3][8];
    h00 = warp_sum(h00), h01 = warp_sum(h01), h11 = warp_sum(h11);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) red[0][warp] = h00, red[1][warp] = h01, red[2][warp] = h11;
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f, c = 0.f;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) a += red[0][k], b += red[1][k], c += red[2][k];
        reinterpret_cast<float4 *>(hessian)[(size_t)ty * nx + tx] = make_float4(a, b, b, c);
    }
}

// ---------------------------------------------------------------------------------------------------------
// flow upscaling (alignment.py:150-172); bilinear / bicubic follow torch F.interpolate(align_corners=False)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float cubic1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

__global__ void upscale_flow_kernel(const float2 *__restrict__ in, int ny_in, int nx_in, float2 *__restrict__ out, int ny_out,
                                    int nx_out, int rep, float factor, int mode) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= nx_out || y >= ny_out) return;
    float2 v = make_float2(0.f, 0.f);
    if (y < ny_in * rep && x < nx_in * rep) {
        if (mode == 0) {
            v = in[(size_t)(y / rep) * nx_in + x / rep];
        } else {
            const float sc = 1.0f / (float)rep;
            float sy = sc * ((float)y + 0.5f) - 0.5f, sx = sc * ((float)x + 0.5f) - 0.5f;
            if (mode == 1) {
                sy = fmaxf(sy, 0.f), sx = fmaxf(sx, 0.f);
                const int y0 = (int)sy, x0 = (int)sx;
                const int y1 = min(y0 + 1, ny_in - 1), x1 = min(x0 + 1, nx_in - 1);
                const float ly = sy - (float)y0, lx = sx - (float)x0;
                const float2 a = in[(size_t)y0 * nx_in + x0], b = in[(size_t)y0 * nx_in + x1];
                const float2 c = in[(size_t)y1 * nx_in + x0], d = in[(size_t)y1 * nx_in + x1];
                v.x = (1.f - ly) * ((1.f - lx) * a.x + lx * b.x) + ly * ((1.f - lx) * c.x + lx * d.x);
                v.y = (1.f - ly) * ((1.f - lx) * a.y + lx * b.y) + ly * ((1.f - lx) * c.y + lx * d.y);
            } else {
                const float A = -0.75f;
                const float fy = floorf(sy), fx = floorf(sx);
                const float ty = sy - fy, tx = sx - fx;
                const float wy[4] = {cubic2(ty + 1.f, A), cubic1(ty, A), cubic1(1.f - ty, A), cubic2(2.f - ty, A)};
                const float wx[4] = {cubic2(tx + 1.f, A), cubic1(tx, A), cubic1(1.f - tx, A), cubic2(2.f - tx, A)};
                float ax = 0.f, ay = 0.f;
                for (int i = 0; i < 4; ++i) {
                    const int yy = min(max((int)fy - 1 + i, 0), ny_in - 1);
                    float rx = 0.f, ry = 0.f;
                    for (int j = 0; j < 4; ++j) {
                        const int xx = min(max((int)fx - 1 + j, 0), nx_in - 1);
                        const float2 q = in[(size_t)yy * nx_in + xx];
                        rx += wx[j] * q.x, ry += wx[j] * q.y;
                    }
                    ax += wy[i] * rx, ay += wy[i] * ry;
                }
                v = make_float2(ax, ay);
            }
        }
        v.x *= factor, v.y *= factor;
    }
    out[(size_t)y * nx_out + x] = v;
}

// ---------------------------------------------------------------------------------------------------------
// L2 block matching.  E(v,u) = sum m^2 - 2 sum ref*m, accumulated in float64 so the argmin is the exact one
// (the reference obtains the same quantity through float32 FFTs; only exact/near ties can differ).
// ---------------------------------------------------------------------------------------------------------
// METRIC 1 is the L1 level as the reference INTENDS it (block_matching.py:78-345: sum |ref - m| over the tile, moving
// samples outside the frame read as zero, first minimum, flow <- rint(flow) + shift); the compiled reference never
// reaches it (bm_l1_compat_kernel below), so it is offered behind an explicit switch only.
template <int TS, int METRIC = 0>
__global__ void __launch_bounds__(256) bm_l2_kernel(const float *__restrict__ ref, int ref_w, const float *__restrict__ mov,
                                                    int mov_h, int mov_w, float2 *__restrict__ flow, int nx, int r) {
    extern __shared__ double bsm[];   // tile and window are staged as float64: the inner loop is LDS.64 + DFMA
    constexpr int NT = TS >= 16 ? 256 : 64;
    constexpr int CPL = TS > 32 ? TS / 32 : 1;          // columns per lane
    constexpr int LPR = TS >= 32 ? 32 : TS;             // lanes per row
    constexpr int RPP = 32 / LPR;                       // rows per pass of a warp
    const int tx = blockIdx.x, ty = blockIdx.y;
    const int sw = TS + 2 * r, n = 2 * r + 1;
    double *s_ref = bsm, *s_win = bsm + TS * TS, *s_err = bsm + TS * TS + sw * sw;
    const float2 f = flow[(size_t)ty * nx + tx];
    const int fx = (int)rintf(f.x), fy = (int)rintf(f.y);                       // flow.round(), :352
    for (int p = threadIdx.x; p < TS * TS; p += NT)
        s_ref[p] = (METRIC ? 1.0 : -2.0) * (double)__ldg(ref + (size_t)(ty * TS + p / TS) * ref_w + tx * TS + p % TS);
    for (int yy0 = threadIdx.x / 32; yy0 < sw; yy0 += NT / 32) {               // one warp per window row
        const int yr = ty * TS + fy - r + yy0;
        const int yy = min(max(yr, 0), mov_h - 1);                              // clamp, :368-369
        for (int xx0 = threadIdx.x & 31; xx0 < sw; xx0 += 32) {
            const int xr = tx * TS + fx - r + xx0;
            const int xx = min(max(xr, 0), mov_w - 1);
            const bool inside = yr == yy && xr == xx;
            s_win[yy0 * sw + xx0] = (METRIC && !inside) ? 0.0 : (double)__ldg(mov + (size_t)yy * mov_w + xx);   // L1: zero fill, :112-121
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lx = (lane % LPR) * CPL, ly0 = lane / LPR;
    for (int s = warp; s < n * n; s += NT / 32) {
        const int v = s / n, u = s % n;
        double e = 0.0;
        const double *wp = s_win + (ly0 + v) * sw + lx + u;
        const double *rp = s_ref + ly0 * TS + lx;
#pragma unroll 4
        for (int y = 0; y < TS; y += RPP) {
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const double m = wp[c];
                e = METRIC ? e + fabs(rp[c] - m) : fma(m, m + rp[c], e);         // |ref - m|  or  m^2 - 2 ref m
            }
            wp += RPP * sw;
            rp += RPP * TS;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
        if (lane == 0) s_err[s] = e;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int best = 0;
        double be = s_err[0];
        for (int s = 1; s < n * n; ++s)
            if (s_err[s] < be) be = s_err[s], best = s;                          // first minimum, torch.argmin
        if (METRIC)
            flow[(size_t)ty * nx + tx] = make_float2((float)(fx + best % n - r), (float)(fy + best / n - r));   // rint(flow) + shift
        else
            flow[(size_t)ty * nx + tx] = make_float2(f.x + (float)(best % n - r), f.y + (float)(best / n - r));
    }
}

// Register-tiled variant for 32x32 tiles (the default tile size), persistent over tiles.
//   E(v,u) = S(v,u) + C(v,u),  S = sum of m^2 over the shifted window,  C = sum of m * (-2 ref).
// C: each warp owns 4 tile rows, each lane one column and keeps -2*ref of its 4 pixels in registers; one window value read
// from shared memory feeds up to 4 (row, vertical shift) pairs and every term is ONE float64 FMA.  Per-lane partials of a
// group of 3 horizontal shifts x all vertical shifts are transposed through shared memory and summed in a fixed order
// (rows, then lanes, then warps).  S: column sums of m^2 per vertical shift (32 FMAs each, fixed order), then 32 column
// sums per shift — 14 % of the operations of C instead of one extra float64 add per term.  Both parts depend only on the
// CONTENT of a shifted window, so equal windows (clamped borders) give bit-equal energies and the first-minimum rule is
// preserved.  A CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...: while it searches one tile, the window and the
// reference pixels of its next tile are already in flight (registers), so the global-memory latency that bounded the
// one-tile-per-CTA version (long-scoreboard stalls, FP64 pipe 40 % busy) is off the critical path.
constexpr int BM_RW = 8, BM_NW = 32 / BM_RW, BM_NT = 32 * BM_NW;     // tile rows per warp, warps and threads per CTA
template <int R>
__global__ void __launch_bounds__(BM_NT, 3) bm_l2_tiled32_kernel(const float *__restrict__ ref, int ref_w, const float *__restrict__ mov,
                                                                 int mov_h, int mov_w, float2 *__restrict__ flow, int nx, int ntiles) {
    constexpr int TS = 32, N = 2 * R + 1, SW = TS + 2 * R, RW = BM_RW, NW = BM_NW, NT = BM_NT, UG = 3, NG = (N + UG - 1) / UG, NV = N * UG;
    constexpr int NWIN = SW * SW, PF = (NWIN + NT - 1) / NT, VG = (N + 2) / 3;
    static_assert(NV <= 32, "search radius too large for the tiled kernel");
    extern __shared__ double bsm[];
    double *s_win0 = bsm;                       // [2][SW][SW]
    double *s_part = s_win0 + 2 * NWIN;         // [NW][N*N]
    double *s_col = s_part + NW * N * N;        // [N][SW]
    double *s_err = s_col + N * SW;             // [N*N]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const int yb = warp * RW;

    // window sample `i` (row-major in the SW x SW window) of tile t with rounded flow (fx, fy): clamped read (:368-369)
    auto win_load = [&](int t, int fx, int fy, int i) -> float {
        const int yy0 = i / SW, xx0 = i - yy0 * SW;
        const int ty = t / nx, tx = t - ty * nx;
        const int yy = min(max(ty * TS + fy - R + yy0, 0), mov_h - 1), xx = min(max(tx * TS + fx - R + xx0, 0), mov_w - 1);
        return __ldg(mov + (size_t)yy * mov_w + xx);
    };
    auto ref_load = [&](int t, int k) -> float {
        const int ty = t / nx, tx = t - ty * nx;
        return __ldg(ref + (size_t)(ty * TS + yb + k) * ref_w + tx * TS + lane);
    };

    int t = blockIdx.x, buf = 0;
    float2 f = make_float2(0.f, 0.f);
    float rfn[RW];
    if (t < ntiles) {
        f = flow[t];
        const int fx = (int)rintf(f.x), fy = (int)rintf(f.y);                    // flow.round(), :352
        for (int i = tid; i < NWIN; i += NT) s_win0[i] = (double)win_load(t, fx, fy, i);
#pragma unroll
        for (int k = 0; k < RW; ++k) rfn[k] = ref_load(t, k);
    }
    __syncthreads();
    for (; t < ntiles; t += gridDim.x, buf ^= 1) {
        double *s_win = s_win0 + buf * NWIN;
        double rf[RW];
#pragma unroll
        for (int k = 0; k < RW; ++k) rf[k] = -2.0 * (double)rfn[k];
        // next tile of this CTA: flow, window and reference pixels go in flight now and land in registers
        const int tn = t + gridDim.x;
        float2 fn = make_float2(0.f, 0.f);
        float pf[PF];
        if (tn < ntiles) {
            fn = flow[tn];                                                        // tile tn is only ever written by this CTA
            const int fx = (int)rintf(fn.x), fy = (int)rintf(fn.y);
#pragma unroll
            for (int q = 0; q < PF; ++q) {
                const int i = tid + q * NT;
                pf[q] = (i < NWIN) ? win_load(tn, fx, fy, i) : 0.f;
            }
#pragma unroll
            for (int k = 0; k < RW; ++k) rfn[k] = ref_load(tn, k);
        }
        // S, step 1: column sums of m^2 for every vertical shift.  Task = (column X, group of 3 shifts): the TS + 2 rows it
        // needs are read once and feed up to 3 sums, each accumulated over its 32 rows in ascending order
        for (int c = tid; c < VG * SW; c += NT) {
            const int vg = c / SW, X = c - vg * SW, v0 = vg * 3;
            const double *q = s_win + v0 * SW + X;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll 2
            for (int y = 0; y < TS + 2; ++y) {
                const double m = (v0 * SW + X + y * SW < NWIN) ? q[y * SW] : 0.0;
                const double mm = __dmul_rn(m, m);      // explicit roundings: the three sums must round alike (equal windows)
                if (y < TS) a0 = __dadd_rn(a0, mm);
                if (y >= 1 && y < TS + 1) a1 = __dadd_rn(a1, mm);
                if (y >= 2) a2 = __dadd_rn(a2, mm);
            }
            s_col[v0 * SW + X] = a0;
            if (v0 + 1 < N) s_col[(v0 + 1) * SW + X] = a1;
            if (v0 + 2 < N) s_col[(v0 + 2) * SW + X] = a2;
        }
        // C: register-tiled cross term
        for (int g = 0; g < NG; ++g) {
            double acc[N][UG];
#pragma unroll
            for (int v = 0; v < N; ++v)
#pragma unroll
                for (int uu = 0; uu < UG; ++uu) acc[v][uu] = 0.0;
#pragma unroll
            for (int rr = 0; rr < RW + N - 1; ++rr) {
                double m[UG];
#pragma unroll
                for (int uu = 0; uu < UG; ++uu) {
                    const int u = g * UG + uu;
                    m[uu] = (u < N) ? s_win[(yb + rr) * SW + lane + u] : 0.0;
                }
#pragma unroll
                for (int k = 0; k < RW; ++k) {
                    const int v = rr - k;        // window row yb+rr is row (yb+k) displaced by v
                    if (v >= 0 && v < N) {
#pragma unroll
                        for (int uu = 0; uu < UG; ++uu) acc[v][uu] = fma(m[uu], rf[k], acc[v][uu]);
                    }
                }
            }
            // sum over the 32 lanes without shared memory: transpose-reduction by butterflies (16 + 8 + 4 + 2 + 1 exchanges);
            // accumulator i ends in lane i.  Every accumulator goes through the same pairing tree (lane l with l^16, then
            // ^8, ...), and IEEE addition commutes, so equal windows still give bit-equal totals
            double a[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = (i < NV) ? acc[i / UG][i % UG] : 0.0;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                const bool hi = (lane & o) != 0;
#pragma unroll
                for (int j = 0; j < o; ++j) {
                    const double send = hi ? a[j] : a[j + o], keep = hi ? a[j + o] : a[j];
                    a[j] = __dadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, o));
                }
            }
            if (lane < NV) {
                const int v = lane / UG, u = g * UG + lane % UG;
                if (u < N) s_part[warp * N * N + v * N + u] = a[0];
            }
        }
        __syncthreads();                        // s_col and s_part complete
        if (tid < N * N) {
            const int v = tid / N, u = tid - v * N;
            const double *q = s_col + v * SW + u;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;      // S, step 2: the 32 column sums of this shift, fixed order
#pragma unroll
            for (int x = 0; x < TS; x += 4) s0 += q[x], s1 += q[x + 1], s2 += q[x + 2], s3 += q[x + 3];
            double e = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) e += s_part[w * N * N + tid];
            s_err[tid] = ((s0 + s1) + (s2 + s3)) + e;
        }
        // the prefetched window of the next tile goes into the other buffer (nobody reads it before the next barrier)
        if (tn < ntiles) {
            double *nw = s_win0 + (buf ^ 1) * NWIN;
#pragma unroll
            for (int q = 0; q < PF; ++q) {
                const int i = tid + q * NT;
                if (i < NWIN) nw[i] = (double)pf[q];
            }
        }
        __syncthreads();
        if (warp == 0) {
            // first minimum in raster order (torch.argmin): lane l scans shifts l, l + 32, l + 64 in that order, then the
            // lanes combine (energy, index) pairs — smaller energy wins, equal energies keep the smaller index
            // (a NaN energy never wins a `<` test: it counts as +inf, and a NaN at shift 0 keeps shift 0 like the serial scan)
            int best = lane;
            double be = (lane < N * N) ? s_err[lane] : INFINITY;
            be = (be != be) ? INFINITY : be;
#pragma unroll
            for (int sft = lane + 32; sft < N * N; sft += 32)
                if (s_err[sft] < be) be = s_err[sft], best = sft;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                const double oe = __shfl_xor_sync(0xffffffffu, be, o);
                const int oi = __shfl_xor_sync(0xffffffffu, best, o);
                if (oe < be || (oe == be && oi < best)) be = oe, best = oi;
            }
            if (lane == 0) {
                if (s_err[0] != s_err[0]) best = 0;
                flow[t] = make_float2(f.x + (float)(best % N - R), f.y + (float)(best / N - R));
            }
        }
        f = fn;
    }
}

__global__ void bm_l1_compat_kernel(float *__restrict__ flow, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flow[i] = rintf(flow[i]);
}

// ---------------------------------------------------------------------------------------------------------
// ICA.  MODE 0: ts = 8 (clamped sampling, float64 1/det); MODE 1: ts = 16, 32 (zero fill);
// MODE 2: ts = 64 (zero fill, rows as read by the reference's sliding window — SURVEY Q4).
// PPT = pixels per thread, NT threads per tile.
// ---------------------------------------------------------------------------------------------------------
template <int TS, int MODE, int NT>
__global__ void __launch_bounds__(NT) ica_kernel(const float *__restrict__ ref, const float *__restrict__ gradx,
                                                 const float *__restrict__ grady, int ref_w, const float4 *__restrict__ hessian,
                                                 const float *__restrict__ mov, int h, int w, float2 *__restrict__ flow, int nx,
                                                 int n_iter) {
    constexpr int PPT = TS * TS / NT;
    const int px = blockIdx.x, py = blockIdx.y, tid = threadIdx.x;
    const float4 Hm = __ldg(hessian + (size_t)py * nx + px);
    const float A00 = Hm.x, A01 = Hm.y, A10 = Hm.z, A11 = Hm.w;
    const float det = A00 * A11 - A01 * A10;
    float det_inv_f = 0.f;
    double det_inv_d = 0.0;
    if (MODE == 0) {
        if (fabs((double)det) < 1e-10) return;                                  // ICA.py:124-126
        det_inv_d = 1.0 / (double)det;
    } else {
        if (fabsf(det) < 1e-10f) return;
        det_inv_f = 1.0f / det;
    }
    __shared__ float s_flow[2];
    __shared__ float s_red[2][NT / 32];
    if (tid == 0) {
        const float2 f = flow[(size_t)py * nx + px];
        s_flow[0] = f.x, s_flow[1] = f.y;
    }
    float rc[PPT], gx[PPT], gy[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        const int p = tid + k * NT;
        const size_t o = (size_t)(py * TS + p / TS) * ref_w + px * TS + p % TS;
        rc[k] = __ldg(ref + o), gx[k] = __ldg(gradx + o), gy[k] = __ldg(grady + o);
    }
    for (int it = 0; it < n_iter; ++it) {
        __syncthreads();
        const float ax = s_flow[0], ay = s_flow[1];
        const int ix = (int)ax, iy = (int)ay;                                   // trunc toward zero
        const float frx = ax - truncf(ax), fry = ay - truncf(ay);              // signed modf
        float B0 = 0.f, B1 = 0.f;
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const int p = tid + k * NT;
            const int ly = p / TS, lx = p % TS;
            const int X = px * TS + lx + ix;
            int Yt = py * TS + ly + iy, Yb = 0;
            float m00, m01, m10, m11;
            if (MODE == 0) {
                const int Xf = min(max(X, 0), w - 1), Yf = min(max(Yt, 0), h - 1);
                const int Xc = min(max(Xf + 1, 0), w - 1), Yc = min(max(Yf + 1, 0), h - 1);
                m00 = __ldg(mov + (size_t)Yf * w + Xf), m01 = __ldg(mov + (size_t)Yf * w + Xc);
                m10 = __ldg(mov + (size_t)Yc * w + Xf), m11 = __ldg(mov + (size_t)Yc * w + Xc);
            } else {
                if (MODE == 2) {
                    const int i_in = ly & 3;                 // ICA.py:441-445
                    const int Ybase = Yt - i_in;
                    Yt = (i_in == 0) ? Ybase : Ybase + i_in + 1;
                    Yb = Ybase + i_in + 2;
                } else {
                    Yb = Yt + 1;
                }
                const bool x0 = X >= 0 && X < w, x1 = X + 1 >= 0 && X + 1 < w;
                const bool yt = Yt >= 0 && Yt < h, yb = Yb >= 0 && Yb < h;
                m00 = (yt && x0) ? __ldg(mov + (size_t)Yt * w + X) : 0.f;
                m01 = (yt && x1) ? __ldg(mov + (size_t)Yt * w + X + 1) : 0.f;
                m10 = (yb && x0) ? __ldg(mov + (size_t)Yb * w + X) : 0.f;
                m11 = (yb && x1) ? __ldg(mov + (size_t)Yb * w + X + 1) : 0.f;
            }
            const float top = m00 + (m01 - m00) * frx;
            const float bot = m10 + (m11 - m10) * frx;
            const float gt = (top + (bot - top) * fry) - rc[k];
            B0 += -gx[k] * gt;
            B1 += -gy[k] * gt;
        }
        B0 = warp_sum(B0), B1 = warp_sum(B1);
        if ((tid & 31) == 0) s_red[0][tid >> 5] = B0, s_red[1][tid >> 5] = B1;
        __syncthreads();
        if (tid == 0) {
            float b0 = s_red[0][0], b1 = s_red[1][0];
            for (int k = 1; k < NT / 32; ++k) b0 += s_red[0][k], b1 += s_red[1][k];
            if (MODE == 0) {
                s_flow[0] = (float)((double)s_flow[0] + det_inv_d * (double)(A11 * b0 - A01 * b1));
                s_flow[1] = (float)((double)s_flow[1] + det_inv_d * (double)(-A10 * b0 + A00 * b1));
            } else {
                s_flow[0] += det_inv_f * (A11 * b0 - A01 * b1);
                s_flow[1] += det_inv_f * (-A10 * b0 + A00 * b1);
            }
        }
    }
    __syncthreads();
    if (tid == 0) flow[(size_t)py * nx + px] = make_float2(s_flow[0], s_flow[1]);
}

// ts = 32 (the default tile size, zero-fill sampling): one thread = 4 x 2 pixels of a tile (128 threads per tile), so the
// bilinear taps of its pixels overlap (3 x 5 moving samples instead of 8 x 4) and ref / gradients load as float4;
// the two block-wide sums of an iteration are finished redundantly by every thread from double-buffered per-warp
// partials (one barrier per iteration instead of two, no serial solve by thread 0).  Same per-pixel arithmetic as
// ica_kernel<32, 1, 256>; only the summation order of B differs (float32 rounding level).
// GRAD: gradx / grady are not read — they are the central differences of `ref` (what hhsr_grad_hessian writes, ICA.py:20-21:
// right - left, down - up, zero outside) and are re-formed from four rows of ref: 96 MB less to read per 12 MP frame for
// identical values.
template <bool VEC4, bool GRAD = false>
__global__ void __launch_bounds__(128) ica32_kernel(const float *__restrict__ ref, const float *__restrict__ gradx,
                                                    const float *__restrict__ grady, int ref_w, const float4 *__restrict__ hessian,
                                                    const float *__restrict__ mov, int h, int w, float2 *__restrict__ flow, int nx,
                                                    int n_iter) {
    constexpr int TS = 32;
    const int px = blockIdx.x, py = blockIdx.y, tid = threadIdx.x;
    const float4 Hm = __ldg(hessian + (size_t)py * nx + px);
    const float A00 = Hm.x, A01 = Hm.y, A10 = Hm.z, A11 = Hm.w;
    const float det = A00 * A11 - A01 * A10;
    if (fabsf(det) < 1e-10f) return;                                            // ICA.py:219-221 (block-uniform)
    const float det_inv = 1.0f / det;
    __shared__ float s_red[2][2][4];
    const int lx = (tid & 7) * 4, ly = (tid >> 3) * 2;
    const int gx0 = px * TS + lx, gy0 = py * TS + ly;
    float rc[2][4], gx[2][4], gy[2][4];
    if (GRAD) {
        const int ref_h = (int)gridDim.y * TS;      // the launcher requires ny * ts == ref_h and nx * ts == ref_w in this mode
        float rows[4][6];                           // ref rows gy0-1 .. gy0+2, columns gx0-1 .. gx0+4 (zero outside)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int y = gy0 - 1 + r;
            const bool yin = y >= 0 && y < ref_h;
            const size_t o = (size_t)(yin ? y : gy0) * ref_w + gx0;
            const float4 a = __ldg(reinterpret_cast<const float4 *>(ref + o));
            rows[r][1] = yin ? a.x : 0.f, rows[r][2] = yin ? a.y : 0.f, rows[r][3] = yin ? a.z : 0.f, rows[r][4] = yin ? a.w : 0.f;
            rows[r][0] = rows[r][5] = 0.f;
            if (r == 1 || r == 2) {                 // the side columns only feed gradx of this thread's own two rows
                rows[r][0] = gx0 > 0 ? __ldg(ref + o - 1) : 0.f;
                rows[r][5] = gx0 + 4 < ref_w ? __ldg(ref + o + 4) : 0.f;
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                rc[r][k] = rows[r + 1][k + 1];
                gx[r][k] = rows[r + 1][k + 2] - rows[r + 1][k];
                gy[r][k] = rows[r + 2][k + 1] - rows[r][k + 1];
            }
    } else
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const size_t o = (size_t)(gy0 + r) * ref_w + gx0;
        if (VEC4) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(ref + o));
            const float4 b = __ldg(reinterpret_cast<const float4 *>(gradx + o));
            const float4 c = __ldg(reinterpret_cast<const float4 *>(grady + o));
            rc[r][0] = a.x, rc[r][1] = a.y, rc[r][2] = a.z, rc[r][3] = a.w;
            gx[r][0] = b.x, gx[r][1] = b.y, gx[r][2] = b.z, gx[r][3] = b.w;
            gy[r][0] = c.x, gy[r][1] = c.y, gy[r][2] = c.z, gy[r][3] = c.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) rc[r][k] = __ldg(ref + o + k), gx[r][k] = __ldg(gradx + o + k), gy[r][k] = __ldg(grady + o + k);
        }
    }
    const float2 f0 = flow[(size_t)py * nx + px];
    float ax = f0.x, ay = f0.y;                                                 // identical in every thread
    for (int it = 0; it < n_iter; ++it) {
        const int ix = (int)ax, iy = (int)ay;                                   // trunc toward zero
        const float frx = ax - truncf(ax), fry = ay - truncf(ay);              // signed modf
        const int X = gx0 + ix, Y = gy0 + iy;
        float m[3][5];
        if (X >= 0 && X + 4 < w && Y >= 0 && Y + 2 < h) {
            const float *q = mov + (size_t)Y * w + X;
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int j = 0; j < 5; ++j) m[r][j] = __ldg(q + r * w + j);
        } else {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const bool yin = Y + r >= 0 && Y + r < h;
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const bool xin = X + j >= 0 && X + j < w;
                    m[r][j] = (yin && xin) ? __ldg(mov + (size_t)(Y + r) * w + X + j) : 0.f;   // zero fill, ICA.py:240-243
                }
            }
        }
        float B0 = 0.f, B1 = 0.f;
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float top = m[r][k] + (m[r][k + 1] - m[r][k]) * frx;
                const float bot = m[r + 1][k] + (m[r + 1][k + 1] - m[r + 1][k]) * frx;
                const float gt = (top + (bot - top) * fry) - rc[r][k];
                B0 += -gx[r][k] * gt;
                B1 += -gy[r][k] * gt;
            }
        B0 = warp_sum(B0), B1 = warp_sum(B1);
        float(*red)[4] = s_red[it & 1];
        if ((tid & 31) == 0) red[0][tid >> 5] = B0, red[1][tid >> 5] = B1;
        __syncthreads();
        float b0 = red[0][0], b1 = red[1][0];
#pragma unroll
        for (int k = 1; k < 4; ++k) b0 += red[0][k], b1 += red[1][k];
        ax += det_inv * (A11 * b0 - A01 * b1);                                  // ICA.py:268-272
        ay += det_inv * (-A10 * b0 + A00 * b1);
    }
    if (tid == 0) flow[(size_t)py * nx + px] = make_float2(ax, ay);
}

}  // namespace hhsr

using namespace hhsr;

extern "C" int hhsr_grad_hessian(const float *img, int h, int w, int ts, float *gradx, float *grady, float *hessian,
                                 hhsr_stream_t stream) {
    HHSR_REQUIRE(img && gradx && grady && hessian, "null pointer");
    HHSR_REQUIRE(h > 0 && w > 0 && ts > 0, "non-positive size");
    HHSR_REQUIRE((uintptr_t)hessian % 16 == 0, "hessian must be 16-byte aligned");
    const int ny = h / ts, nx = w / ts;
    dim3 grid(ceil_div(w, ts), ceil_div(h, ts));
    const int nt = ts * ts >= 256 ? 256 : 64;
    grad_hessian_kernel<<<grid, nt, 0, (cudaStream_t)stream>>>(img, h, w, ts, ny, nx, gradx, grady, hessian);
    return launch_status("grad_hessian");
}

extern "C" int hhsr_upscale_flow(const float *flow_in, int ny_in, int nx_in, float *flow_out, int ny_out, int nx_out,
                                 int repeat, float factor, int mode, hhsr_stream_t stream) {
    HHSR_REQUIRE(flow_in && flow_out, "null pointer");
    HHSR_REQUIRE(ny_in > 0 && nx_in > 0 && ny_out > 0 && nx_out > 0 && repeat >= 1, "non-positive size");
    HHSR_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0 (nearest), 1 (bilinear) or 2 (bicubic)");
    dim3 block(16, 16), grid(ceil_div(nx_out, 16), ceil_div(ny_out, 16));
    upscale_flow_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2 *>(flow_in), ny_in, nx_in,
                                                                 reinterpret_cast<float2 *>(flow_out), ny_out, nx_out, repeat,
                                                                 factor, mode);
    return launch_status("upscale_flow");
}

extern "C" int hhsr_bm_l2_search(const float *ref, int ref_h, int ref_w, const float *mov, int mov_h, int mov_w,
                                 float *flow, int ny, int nx, int ts, int radius, hhsr_stream_t stream) {
    HHSR_REQUIRE(ref && mov && flow, "null pointer");
    HHSR_REQUIRE(ny > 0 && nx > 0 && mov_h > 0 && mov_w > 0, "non-positive size");
    if (!(ts == 8 || ts == 16 || ts == 32 || ts == 64))
        return unsupported("L2 block matching tile size must be 8, 16, 32 or 64 (block_matching.py:48-57)");
    HHSR_REQUIRE(radius >= 0 && radius <= 8, "search radius must be in [0, 8]");
    HHSR_REQUIRE(ny * ts <= ref_h && nx * ts <= ref_w, "tile grid exceeds the reference level");
    const int sw = ts + 2 * radius, n = 2 * radius + 1;
    const size_t smem = (size_t)(ts * ts + sw * sw + n * n) * sizeof(double);
    dim3 grid(nx, ny);
    cudaStream_t st = (cudaStream_t)stream;
    float2 *F2 = reinterpret_cast<float2 *>(flow);
#define HHSR_BM(TS, NT)                                                                                      \
    do {                                                                                                     \
        static std::atomic<unsigned long long> done{0};                                                      \
        if (smem > 48 * 1024) ensure_dynamic_smem(bm_l2_kernel<TS>, done, smem);                              \
        bm_l2_kernel<TS><<<grid, NT, smem, st>>>(ref, ref_w, mov, mov_h, mov_w, F2, nx, radius);              \
    } while (0)
    if (ts == 32 && radius >= 1 && radius <= 4) {
#define HHSR_BMT(R)                                                                                                   \
    do {                                                                                                              \
        constexpr int N = 2 * R + 1, SW = 32 + 2 * R;                                                                 \
        const size_t sm = (size_t)(2 * SW * SW + BM_NW * N * N + N * SW + N * N) * sizeof(double);                    \
        static std::atomic<unsigned long long> done{0};                                                               \
        ensure_dynamic_smem(bm_l2_tiled32_kernel<R>, done, sm);                                                       \
        const int ntiles = nx * ny;                                                                                   \
        const int ctas = ntiles < 148 * 3 ? ntiles : 148 * 3;       /* persistent: 3 CTAs per SM walk the tiles */    \
        bm_l2_tiled32_kernel<R><<<ctas, BM_NT, sm, st>>>(ref, ref_w, mov, mov_h, mov_w, F2, nx, ntiles);               \
    } while (0)
        switch (radius) {
            case 1: HHSR_BMT(1); break;
            case 2: HHSR_BMT(2); break;
            case 3: HHSR_BMT(3); break;
            default: HHSR_BMT(4); break;
        }
#undef HHSR_BMT
        return launch_status("bm_l2_search");
    }
    switch (ts) {
        case 8: HHSR_BM(8, 64); break;
        case 16: HHSR_BM(16, 256); break;
        case 32: HHSR_BM(32, 256); break;
        default: HHSR_BM(64, 256); break;
    }
#undef HHSR_BM
    return launch_status("bm_l2_search");
}

extern "C" int hhsr_bm_l1_search(const float *ref, int ref_h, int ref_w, const float *mov, int mov_h, int mov_w,
                                 float *flow, int ny, int nx, int ts, int radius, hhsr_stream_t stream) {
    HHSR_REQUIRE(ref && mov && flow, "null pointer");
    HHSR_REQUIRE(ny > 0 && nx > 0 && mov_h > 0 && mov_w > 0, "non-positive size");
    if (!(ts == 16 || ts == 32 || ts == 64))
        return unsupported("L1 local search tile size must be 16, 32 or 64 (block_matching.py:86-93)");
    HHSR_REQUIRE(radius >= 0 && radius <= 8, "search radius must be in [0, 8]");
    HHSR_REQUIRE(ny * ts <= ref_h && nx * ts <= ref_w, "tile grid exceeds the reference level");
    const int sw = ts + 2 * radius, n = 2 * radius + 1;
    const size_t smem = (size_t)(ts * ts + sw * sw + n * n) * sizeof(double);
    dim3 grid(nx, ny);
    cudaStream_t st = (cudaStream_t)stream;
    float2 *F2 = reinterpret_cast<float2 *>(flow);
#define HHSR_BM1(TS)                                                                                          \
    do {                                                                                                     \
        static std::atomic<unsigned long long> done{0};                                                      \
        if (smem > 48 * 1024) ensure_dynamic_smem(bm_l2_kernel<TS, 1>, done, smem);                           \
        bm_l2_kernel<TS, 1><<<grid, 256, smem, st>>>(ref, ref_w, mov, mov_h, mov_w, F2, nx, radius);          \
    } while (0)
    switch (ts) {
        case 16: HHSR_BM1(16); break;
        case 32: HHSR_BM1(32); break;
        default: HHSR_BM1(64); break;
    }
#undef HHSR_BM1
    return launch_status("bm_l1_search");
}

extern "C" int hhsr_bm_l1_compat(float *flow, int n, hhsr_stream_t stream) {
    HHSR_REQUIRE(flow && n > 0, "null pointer or empty flow");
    bm_l1_compat_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(flow, n);
    return launch_status("bm_l1_compat");
}

extern "C" int hhsr_ica(const float *ref, const float *gradx, const float *grady, int ref_h, int ref_w,
                        const float *hessian, const float *mov, int mov_h, int mov_w, float *flow, int ny, int nx, int ts,
                        int n_iter, hhsr_stream_t stream) {
    HHSR_REQUIRE(ref && hessian && mov && flow, "null pointer");
    // gradx == grady == NULL: the gradients are the central differences of ref (hhsr_grad_hessian) and are re-formed on the fly
    const bool on_the_fly = !gradx && !grady;
    HHSR_REQUIRE(on_the_fly || (gradx && grady), "gradx and grady must both be given or both be null");
    if (on_the_fly && !(ts == 32 && ny * ts == ref_h && nx * ts == ref_w && ref_w % 4 == 0 && (uintptr_t)ref % 16 == 0))
        return unsupported("on-the-fly gradients need tile size 32, a reference level that is a whole number of tiles and 16-byte aligned rows");
    HHSR_REQUIRE(ny > 0 && nx > 0 && mov_h > 0 && mov_w > 0 && n_iter > 0, "non-positive size");
    HHSR_REQUIRE(ny * ts <= ref_h && nx * ts <= ref_w, "tile grid exceeds the reference level");
    HHSR_REQUIRE((uintptr_t)hessian % 16 == 0 && (uintptr_t)flow % 8 == 0, "hessian/flow misaligned");
    dim3 grid(nx, ny);
    cudaStream_t st = (cudaStream_t)stream;
    const float4 *H4 = reinterpret_cast<const float4 *>(hessian);
    float2 *F2 = reinterpret_cast<float2 *>(flow);
    switch (ts) {
        case 8: ica_kernel<8, 0, 64><<<grid, 64, 0, st>>>(ref, gradx, grady, ref_w, H4, mov, mov_h, mov_w, F2, nx, n_iter); break;
        case 16: ica_kernel<16, 1, 256><<<grid, 256, 0, st>>>(ref, gradx, grady, ref_w, H4, mov, mov_h, mov_w, F2, nx, n_iter); break;
        case 32:
            if (!gradx && !grady)
                ica32_kernel<true, true><<<grid, 128, 0, st>>>(ref, gradx, grady, ref_w, H4, mov, mov_h, mov_w, F2, nx, n_iter);
            else if (ref_w % 4 == 0 && (uintptr_t)ref % 16 == 0 && (uintptr_t)gradx % 16 == 0 && (uintptr_t)grady % 16 == 0)
                ica32_kernel<true><<<grid, 128, 0, st>>>(ref, gradx, grady, ref_w, H4, mov, mov_h, mov_w, F2, nx, n_iter);
            else
                ica32_kernel<false><<<grid, 128, 0, st>>>(ref, gradx, grady, ref_w, H4, mov, mov_h, mov_w, F2, nx, n_iter);
            break;
        case 64: ica_kernel<64, 2, 256><<<grid, 256, 0, st>>>(ref, gradx, grady, ref_w, H4, mov, mov_h, mov_w, F2, nx, n_iter); break;
        default: return unsupported("ICA kernel for this tile size not implemented (ICA.py:100)");
    }
    return launch_status("ica");
}
