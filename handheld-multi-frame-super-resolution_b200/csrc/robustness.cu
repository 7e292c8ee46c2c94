// Robustness estimation (Alg. 6-9) for sm_100a.
//
// Replaces handheld_super_resolution/robustness.py: cuda_compute_guide_image (:206-226) + cuda_compute_local_stats
// (:268-294) -> guide_stats_kernel; cuda_uspcale_dogson (:358-418) -> upscale_warp_kernel; and the chain
// upscale_warp(comp) -> cuda_compute_dist (:452-462) -> cuda_apply_noise_model (:504-533) -> cuda_compute_s
// (:569-611) -> cuda_robustness_threshold (:626-639) -> robustness_kernel (one pass, no full-size temporaries;
// the reference materialises 9 of them); cuda_compute_local_min (:669-687) + utils.cuda_add (:116-120) ->
// local_min5_kernel.
//
// Arithmetic follows the compiled reference: float64 wherever Numba promotes (white-balance division, Dodgson
// weights, noise-model shrinkage, the final S*exp - t), float32 sums where the reference keeps float32 arrays.
#include "common.cuh"

namespace hhsr {

constexpr int RBX = 32, RBY = 8;

struct GuideParams {
    int cfa;         // packed 2x2 channel ids
    double inv_wb[3];   // 1 / white balance gain: x * (1/wb) rounds to the same float32 as the reference's float64 x / wb
                        // (the float64 results differ by <= 1 ulp, 2^-29 of a float32 ulp) at a third of the instructions
};

// guide value of channel c at guide pixel (y, x): robustness.py:206-226
__device__ __forceinline__ void guide_rgb(const float *__restrict__ raw, int W, int y, int x, const GuideParams &p,
                                          float (&out)[3]) {
    const float2 a = __ldg(reinterpret_cast<const float2 *>(raw + (size_t)(2 * y) * W + 2 * x));
    const float2 b = __ldg(reinterpret_cast<const float2 *>(raw + (size_t)(2 * y + 1) * W + 2 * x));
    const float q[4] = {a.x, a.y, b.x, b.y};
    double g = 0.0;
    out[0] = out[1] = out[2] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = (p.cfa >> (2 * k)) & 3;
        const double v = (double)q[k] * p.inv_wb[c];
        if (c == 1) g += v;
        if (c == 0) out[0] = (float)v;     // selects, not out[c]: a run-time index would put the array in local memory
        if (c == 2) out[2] = (float)v;
    }
    out[1] = (float)(g * 0.5);
}

__global__ void __launch_bounds__(RBX *RBY) guide_stats_kernel(const float *__restrict__ raw, int W, int h, int w,
                                                               GuideParams p, float *__restrict__ means,
                                                               float *__restrict__ vars) {
    __shared__ float s[3][RBY + 2][RBX + 2];
    const int x0 = blockIdx.x * RBX, y0 = blockIdx.y * RBY;
    for (int t = threadIdx.y * RBX + threadIdx.x; t < (RBY + 2) * (RBX + 2); t += RBX * RBY) {
        const int ly = t / (RBX + 2), lx = t % (RBX + 2);
        const int gy = min(max(y0 + ly - 1, 0), h - 1), gx = min(max(x0 + lx - 1, 0), w - 1);   // clamp, :287-288
        float rgb[3];
        guide_rgb(raw, W, gy, gx, p, rgb);
        s[0][ly][lx] = rgb[0], s[1][ly][lx] = rgb[1], s[2][ly][lx] = rgb[2];
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= w || y >= h) return;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const float v = s[c][threadIdx.y + i][threadIdx.x + j];
                s1 += v;
                s2 = __fmaf_rn(v, v, s2);
            }
        const double ninth = 1.0 / 9.0;           // (double)s / 9.0 of the reference, as a multiplication (see inv_wb)
        const double m = (double)s1 * ninth;
        const size_t o = ((size_t)c * h + y) * w + x;
        means[o] = (float)m;
        if (vars) vars[o] = (float)((double)s2 * ninth - m * m);
    }
}

// The two halves of guide_stats_kernel as stand-alone stages of the reference API (robustness.py:173-226, 228-294);
// same arithmetic in the same order, so their composition is bit-equal to the fused kernel.
__global__ void __launch_bounds__(RBX *RBY) guide_image_kernel(const float *__restrict__ raw, int W, int h, int w, GuideParams p,
                                                               float *__restrict__ guide) {
    const int x = blockIdx.x * RBX + threadIdx.x, y = blockIdx.y * RBY + threadIdx.y;
    if (x >= w || y >= h) return;
    float rgb[3];
    guide_rgb(raw, W, y, x, p, rgb);
#pragma unroll
    for (int c = 0; c < 3; ++c) guide[((size_t)c * h + y) * w + x] = rgb[c];
}

__global__ void __launch_bounds__(RBX *RBY) local_stats_kernel(const float *__restrict__ guide, int h, int w,
                                                               float *__restrict__ means, float *__restrict__ vars) {
    const int x = blockIdx.x * RBX + threadIdx.x, y = blockIdx.y * RBY + threadIdx.y, c = blockIdx.z;
    if (x >= w || y >= h) return;
    const float *g = guide + (size_t)c * h * w;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = -1; i <= 1; ++i)
#pragma unroll
        for (int j = -1; j <= 1; ++j) {
            const float v = __ldg(g + (size_t)min(max(y + i, 0), h - 1) * w + min(max(x + j, 0), w - 1));   // clamp, :287-288
            s1 += v;
            s2 = __fmaf_rn(v, v, s2);
        }
    const double ninth = 1.0 / 9.0;
    const double m = (double)s1 * ninth;
    const size_t o = ((size_t)c * h + y) * w + x;
    means[o] = (float)m;
    vars[o] = (float)((double)s2 * ninth - m * m);
}

__device__ __forceinline__ double dodgson(double t) {   // utils_image.py:398-406
    const double a = fabs(t);
    if (a <= 0.5) return -2.0 * a * a + 1.0;
    if (a <= 1.5) return a * a - 5.0 / 2.0 * a + 1.5;
    return 0.0;
}

// x2 Dodgson upsampling of a 3-channel statistic at raw pixel (y, x) displaced by (fx, fy); false if the source
// position leaves the guide image (the reference then writes +inf).
__device__ __forceinline__ bool dodgson_sample(const float *__restrict__ lr, int h, int w, int y, int x, float fx,
                                               float fy, float (&out)[3]) {
    const double ly = ((double)y + (double)fy + 0.5) / 2.0 - 0.5;    // robustness.py:380-381 (s = 2)
    const double lx = ((double)x + (double)fx + 0.5) / 2.0 - 0.5;
    if (!(0.0 <= ly && ly < (double)h && 0.0 <= lx && lx < (double)w)) return false;
    const int cy = (int)llrint(ly), cx = (int)llrint(lx);
    float buf[3] = {0.f, 0.f, 0.f};
    double wacc = 0.0;
    const size_t plane = (size_t)h * w;
#pragma unroll
    for (int i = -1; i <= 1; ++i) {
        const int y_ = min(max(cy + i, 0), h - 1);
        const double wy = dodgson((double)y_ - ly);
#pragma unroll
        for (int j = -1; j <= 1; ++j) {
            const int x_ = min(max(cx + j, 0), w - 1);
            const double wgt = wy * dodgson((double)x_ - lx);
            const float *q = lr + (size_t)y_ * w + x_;
#pragma unroll
            for (int c = 0; c < 3; ++c) buf[c] = (float)((double)buf[c] + (double)__ldg(q + c * plane) * wgt);
            wacc += wgt;
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) out[c] = (float)((double)buf[c] / wacc);
    return true;
}

// Same sampler for the per-frame hot kernel, in float32 with an exact integer/fraction split of the position:
// y + 0.5 + f = T + u with T = y + trunc(f) (integer), u = 0.5 + frac(f) (float32, exact to 1 ulp), hence
// ly = (T + u)/2 - 0.5 = (T >> 1) + v, v = 0.5 (T & 1) + u/2 - 0.5 in (-1, 1): tap offsets and Dodgson weights are
// formed from v, so the float32 result agrees with the reference's float64 weights to float32 rounding (positions
// of several thousand pixels would otherwise cost 1e-4 px).
struct Axis {
    int i[3];      // clamped tap indices
    float w[3];    // Dodgson weights
    bool ok;
};
__device__ __forceinline__ float dodgson_f(float t) {
    const float a = fabsf(t);
    if (a <= 0.5f) return -2.0f * a * a + 1.0f;
    if (a <= 1.5f) return a * a - 2.5f * a + 1.5f;
    return 0.0f;
}
__device__ __forceinline__ Axis dodgson_axis(int y, float f, int n) {
    Axis A;
    const float tf = truncf(f);
    const int T = y + (int)tf;
    const float u = 0.5f + (f - tf);
    const int base = T >> 1;                                  // arithmetic shift == floor(T/2)
    const float v = 0.5f * (float)(T & 1) + 0.5f * u - 0.5f;   // ly = base + v
    const int fl = base + (int)floorf(v);
    A.ok = (fl >= 0) && (fl < n);                             // 0 <= ly < n  (robustness.py:383-384)
    float rv = rintf(v);
    int c = base + (int)rv;
    if (fabsf(v - rv) == 0.5f && (c & 1)) c += (v > rv) ? 1 : -1;   // round half to even on ly, not on v
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int q = min(max(c + k - 1, 0), n - 1);
        A.i[k] = q;
        A.w[k] = dodgson_f((float)(q - base) - v);
    }
    return A;
}

__global__ void __launch_bounds__(RBX *RBY) upscale_warp_kernel(const float *__restrict__ lr, int h, int w,
                                                                const float *__restrict__ flow, int nx, int ts,
                                                                float *__restrict__ hr) {
    const int x = blockIdx.x * RBX + threadIdx.x, y = blockIdx.y * RBY + threadIdx.y;
    const int H = 2 * h, W = 2 * w;
    if (x >= W || y >= H) return;
    float fx = 0.f, fy = 0.f;
    if (flow) {
        const float2 f = __ldg(reinterpret_cast<const float2 *>(flow) + (size_t)(y / ts) * nx + x / ts);
        fx = f.x, fy = f.y;
    }
    // float32 taps from the exact integer/fraction position split (see dodgson_axis); the float64 sampler above is
    // kept as the literal restatement and used when HHSR_DODGSON_F64 is defined
#ifdef HHSR_DODGSON_F64
    float v[3];
    const bool ok = dodgson_sample(lr, h, w, y, x, fx, fy, v);
#else
    const Axis ay = dodgson_axis(y, fy, h), ax = dodgson_axis(x, fx, w);
    const bool ok = ay.ok && ax.ok;
    float v[3] = {0.f, 0.f, 0.f};
    if (ok) {
        const unsigned lplane = (unsigned)h * (unsigned)w;
        float wacc = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float *row = lr + (unsigned)ay.i[i] * (unsigned)w;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const float wgt = ay.w[i] * ax.w[j];
#pragma unroll
                for (int c = 0; c < 3; ++c) v[c] = fmaf(__ldg(row + ax.i[j] + c * lplane), wgt, v[c]);
                wacc += wgt;
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = v[c] / wacc;
    }
#endif
    const size_t plane = (size_t)H * W, o = (size_t)y * W + x;
#pragma unroll
    for (int c = 0; c < 3; ++c) hr[o + c * plane] = ok ? v[c] : INFINITY;
}

struct RobParams {
    double t, s1, s2, Mt;
    int force_generic;
};

// flow irregularity of tile (py, px): s1 if the range of the flow over the 3x3 tile neighbourhood exceeds Mt
// (robustness.py:569-611)
__device__ __forceinline__ float tile_S(const float2 *__restrict__ fl, int py, int px, int ny, int nx, const RobParams &p) {
    float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
#pragma unroll
    for (int i = -1; i <= 1; ++i)
#pragma unroll
        for (int j = -1; j <= 1; ++j) {
            const int yy = py + i, xx = px + j;
            if (yy < 0 || yy >= ny || xx < 0 || xx >= nx) continue;
            const float2 q = __ldg(fl + (size_t)yy * nx + xx);
            mxx = fmaxf(mxx, q.x), mxy = fmaxf(mxy, q.y), mnx = fminf(mnx, q.x), mny = fminf(mny, q.y);
        }
    const float d0 = mxx - mnx, d1 = mxy - mny;
    return ((double)(d0 * d0 + d1 * d1) > p.Mt * p.Mt) ? (float)p.s1 : (float)p.s2;
}

// (sigma_t^2, d_t^2) per brightness level as float32 pairs, so the per-pixel lookup is one 8-byte load
__global__ void noise_table_kernel(const double *__restrict__ std_curve, const double *__restrict__ diff_curve, int n,
                                   float2 *__restrict__ table) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) table[i] = make_float2((float)(std_curve[i] * std_curve[i]), (float)(diff_curve[i] * diff_curve[i]));
}

// Reference-side noise terms, once per burst: everything of cuda_apply_noise_model (robustness.py:504-533) that
// depends on the reference statistics only.  terms[c] = d_t^2 of the brightness level of channel c (c = 0..2),
// terms[3] = sum_c max(sigma_p^2, sigma_t^2) accumulated in channel order in float32.
__global__ void ref_noise_terms_kernel(const float *__restrict__ ref_means, const float *__restrict__ ref_vars, size_t plane,
                                       const float2 *__restrict__ table, int n_curve, float *__restrict__ terms) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < plane; o += stride) {
        float sigma_sq = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float brightness = ref_means[o + c * plane];
            int id = 0;
            if (isfinite(brightness)) id = (int)llrint(1000.0 * (double)brightness);   // robustness.py:519 (float64 product)
            id = min(max(id, 0), n_curve - 1);
            const float2 t = __ldg(table + id);
            sigma_sq += fmaxf(ref_vars[o + c * plane], t.x);                           // max(sigma_p^2, sigma_t^2), :524
            terms[o + c * plane] = t.y;
        }
        terms[o + 3 * plane] = sigma_sq;
    }
}

// Reference-side statistics in ONE pass (init_robustness + the reference part of the noise model): x2 Dodgson
// upsampling (no flow) of the guide means and variances [3][h][w] and, from them, the noise terms — the upsampled
// variances are consumed on the fly and only written when the caller wants the reference's (means, stds) pair.
// One thread = the 2x2 raw pixels of one guide pixel.  Without flow the sampling position is (y + 0.5)/2 - 0.5, i.e.
// guide row y >> 1 -/+ 0.25 for even/odd y: three taps around the nearest guide pixel with the constant Dodgson weights
// q(-0.75), q(0.25), q(1.25) = 0.1875, 0.875, -0.0625 (sum 1; reversed for odd pixels).  Guide pixels on the border ring
// (clamped taps, the +inf rule for y = 0 / x = 0) take the per-pixel sampler of upscale_warp_kernel.
__global__ void __launch_bounds__(RBX *RBY) ref_stats_terms_kernel(const float *__restrict__ gm, const float *__restrict__ gv, int h,
                                                                   int w, const float2 *__restrict__ table, int n_curve,
                                                                   float *__restrict__ ref_means, float *__restrict__ ref_vars,
                                                                   float *__restrict__ terms) {
    const int gx = blockIdx.x * RBX + threadIdx.x, gy = blockIdx.y * RBY + threadIdx.y;
    if (gx >= w || gy >= h) return;
    const int H = 2 * h, W = 2 * w;
    const size_t plane = (size_t)H * W, lplane = (size_t)h * w;
    float m[2][2][3], v[2][2][3];      // [row parity][col parity][channel]
    bool ok[2][2] = {{true, true}, {true, true}};
    if (gy >= 1 && gy <= h - 2 && gx >= 1 && gx <= w - 2) {
        const float wk[2][3] = {{0.1875f, 0.875f, -0.0625f}, {-0.0625f, 0.875f, 0.1875f}};
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int s_ = 0; s_ < 2; ++s_) {
                const float *src = (s_ ? gv : gm) + c * lplane + (size_t)(gy - 1) * w + (gx - 1);
                float t[3][3];
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) t[i][j] = __ldg(src + i * w + j);
#pragma unroll
                for (int py = 0; py < 2; ++py) {
                    float col[3];
#pragma unroll
                    for (int j = 0; j < 3; ++j) col[j] = fmaf(t[2][j], wk[py][2], fmaf(t[1][j], wk[py][1], t[0][j] * wk[py][0]));
#pragma unroll
                    for (int px = 0; px < 2; ++px) {
                        const float r = fmaf(col[2], wk[px][2], fmaf(col[1], wk[px][1], col[0] * wk[px][0]));
                        if (s_) v[py][px][c] = r; else m[py][px][c] = r;
                    }
                }
            }
    } else {
#pragma unroll 1
        for (int py = 0; py < 2; ++py)
#pragma unroll 1
            for (int px = 0; px < 2; ++px) {
                const Axis ay = dodgson_axis(2 * gy + py, 0.f, h), ax = dodgson_axis(2 * gx + px, 0.f, w);
                ok[py][px] = ay.ok && ax.ok;
                float am[3] = {0.f, 0.f, 0.f}, av[3] = {0.f, 0.f, 0.f}, wacc = 0.f;
                if (ok[py][px]) {
                    for (int i = 0; i < 3; ++i)
                        for (int j = 0; j < 3; ++j) {
                            const float wgt = ay.w[i] * ax.w[j];
                            const size_t q = (size_t)ay.i[i] * w + ax.i[j];
                            for (int c = 0; c < 3; ++c) {
                                am[c] = fmaf(__ldg(gm + q + c * lplane), wgt, am[c]);
                                av[c] = fmaf(__ldg(gv + q + c * lplane), wgt, av[c]);
                            }
                            wacc += wgt;
                        }
                }
                for (int c = 0; c < 3; ++c) {
                    m[py][px][c] = ok[py][px] ? am[c] / wacc : INFINITY;      // out of the guide image: +inf (robustness.py:383-391)
                    v[py][px][c] = ok[py][px] ? av[c] / wacc : INFINITY;
                }
            }
    }
#pragma unroll
    for (int py = 0; py < 2; ++py) {
        const size_t o = (size_t)(2 * gy + py) * W + 2 * gx;
        float sig[2] = {0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float dt[2];
#pragma unroll
            for (int px = 0; px < 2; ++px) {
                const float brightness = m[py][px][c];
                int id = 0;
                if (isfinite(brightness)) id = (int)llrint(1000.0 * (double)brightness);   // robustness.py:519
                id = min(max(id, 0), n_curve - 1);
                const float2 t = __ldg(table + id);
                sig[px] += fmaxf(v[py][px][c], t.x);                                       // :524
                dt[px] = t.y;
            }
            *reinterpret_cast<float2 *>(ref_means + o + c * plane) = make_float2(m[py][0][c], m[py][1][c]);
            *reinterpret_cast<float2 *>(terms + o + c * plane) = make_float2(dt[0], dt[1]);
            if (ref_vars) *reinterpret_cast<float2 *>(ref_vars + o + c * plane) = make_float2(v[py][0][c], v[py][1][c]);
        }
        *reinterpret_cast<float2 *>(terms + o + 3 * plane) = make_float2(sig[0], sig[1]);
    }
}

// Noise model + threshold for one pixel given its warped comp means (robustness.py:452-462, 504-533, 626-639)
__device__ __forceinline__ float robustness_finish(float d_sq, float sigma_sq, float S, double t) {
    const float e = expf(-__fdividef(d_sq, sigma_sq));                      // math.exp(float32), :638
    const float v = (float)((double)(S * e) - t);
    return fminf(1.f, fmaxf(0.f, v));                                       // NaN -> 0
}
__device__ __forceinline__ float shrunk_dist(float brightness, float comp, float dt_sq) {
    const float d_p = fabsf(brightness - comp);                             // :462
    const float d_p_sq = d_p * d_p;
    const float shrink = __fdividef(d_p_sq, d_p_sq + dt_sq);
    return d_p_sq * shrink * shrink;
}

// Generic per-pixel path (any tile size, frame borders): Dodgson axes per pixel.
__device__ __forceinline__ float robustness_pixel(const float *__restrict__ comp_lr, const float *__restrict__ ref_means,
                                                  const float *__restrict__ terms, int H, int W, int x, int y, const Axis &ay,
                                                  const Axis &ax, float S, double t) {
    const int h = H / 2, w = W / 2;
    const unsigned plane = (unsigned)H * (unsigned)W, lplane = (unsigned)h * (unsigned)w, o = (unsigned)y * (unsigned)W + x;
    const float rm0 = __ldg(ref_means + o);
    if (!(ay.ok && ax.ok && isfinite(rm0))) return 0.f;   // any non-finite statistic ends as clamp(NaN) = 0
    float buf[3] = {0.f, 0.f, 0.f}, wacc = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float *row = comp_lr + (unsigned)ay.i[i] * (unsigned)w;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float wgt = ay.w[i] * ax.w[j];
            const float *q = row + ax.i[j];
#pragma unroll
            for (int c = 0; c < 3; ++c) buf[c] = fmaf(__ldg(q + c * lplane), wgt, buf[c]);
            wacc += wgt;
        }
    }
    const float inv_w = __fdividef(1.0f, wacc);
    float d_sq = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float brightness = c == 0 ? rm0 : __ldg(ref_means + o + c * plane);
        d_sq += shrunk_dist(brightness, buf[c] * inv_w, __ldg(terms + o + c * plane));
    }
    return robustness_finish(d_sq, __ldg(terms + o + 3 * plane), S, t);
}

// One thread = 4 pixels x 2 rows; a 8x16-thread block covers 32x32 raw pixels.  When the tile size is a multiple of 32
// the block lies inside one flow tile, and because the guide image is at half resolution the Dodgson tap weights
// depend on the pixel only through the PARITY of (x + trunc(flow)): two weight sets per axis for the whole block.
// The block prologue expands them to zero-padded per-pixel weight rows (5 columns for the 4 pixels, 4 rows for the
// 2 rows of a thread), so a thread reads a 4x5 window of the comp guide means per channel (instead of 9 taps per
// pixel) and blends it separably.  Blocks whose windows touch the guide border (clamped taps, +inf band) and
// tile sizes that are not a multiple of 32 use the per-pixel path.
#ifndef HHSR_ROB_MINBLOCKS
#define HHSR_ROB_MINBLOCKS 8
#endif
constexpr int RTX = 8, RTY = 16;
struct RobTile {
    float wx[4][5], wy[2][4];
    int dx0, dy0, fast;
    float S;
    float2 f;
};

// Window of pixel (or row) k along one axis: T = coordinate + trunc(flow) has parity e = (it + k) & 1, the guide
// position is (T >> 1) + v with v = 0.5 e + 0.5 u - 0.5, u = 0.5 + frac(flow) (see dodgson_axis); the 3 taps start at
// (T >> 1) + rint(v) - 1.  Returns that start relative to (coordinate0 >> 1) (coordinate0 even) and the normalised
// weights.  Ties of rint(): the extra tap has weight q(1.5) = 0 either way.
__device__ __forceinline__ int axis_window(float fl, int k, float (&wgt)[3]) {
    const float tf = truncf(fl);
    const int it = (int)tf;
    const float u = 0.5f + (fl - tf);
    const float v = 0.5f * (float)((it + k) & 1) + 0.5f * u - 0.5f;
    const int off = (int)rintf(v);
    float sum = 0.f;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        wgt[q] = dodgson_f((float)(off - 1 + q) - v);
        sum += wgt[q];
    }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int q = 0; q < 3; ++q) wgt[q] *= inv;
    return ((it + k) >> 1) + off - 1;
}

__global__ void __launch_bounds__(RTX *RTY, HHSR_ROB_MINBLOCKS) robustness_kernel(const float *__restrict__ comp_lr, const float *__restrict__ ref_means,
                                                                 const float *__restrict__ terms, int H, int W,
                                                                 const float *__restrict__ flow, int ny, int nx, int ts,
                                                                 RobParams p, float *__restrict__ R) {
    const float2 *fl = reinterpret_cast<const float2 *>(flow);
    const int h = H / 2, w = W / 2;
    const int xb = blockIdx.x * (RTX * 4), yb = blockIdx.y * (RTY * 2);
    const int x0 = xb + threadIdx.x * 4, y0 = yb + threadIdx.y * 2;
    __shared__ RobTile s;
    const bool uniform = (ts % 32) == 0 && (W % 4) == 0 && !p.force_generic;
    if (uniform) {
        // prologue spread over the first warp: lanes 0-3 the x rows of the weight table, 4-5 the y rows, 6 the
        // flow-irregularity factor, 7 the border test
        const int tid = threadIdx.y * RTX + threadIdx.x;
        if (tid < 8) {
            const int py = yb / ts, px = xb / ts;
            const float2 f = __ldg(fl + (size_t)py * nx + px);
            float wgt[3], w0[3];
            if (tid < 4) {
                const int d = axis_window(f.x, tid, wgt) - axis_window(f.x, 0, w0);
#pragma unroll
                for (int m = 0; m < 5; ++m) {
                    const int k = m - d;
                    s.wx[tid][m] = (k == 0) ? wgt[0] : (k == 1) ? wgt[1] : (k == 2) ? wgt[2] : 0.f;
                }
            } else if (tid < 6) {
                const int d = axis_window(f.y, tid - 4, wgt) - axis_window(f.y, 0, w0);
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const int k = m - d;
                    s.wy[tid - 4][m] = (k == 0) ? wgt[0] : (k == 1) ? wgt[1] : (k == 2) ? wgt[2] : 0.f;
                }
            } else if (tid == 6) {
                s.S = tile_S(fl, py, px, ny, nx, p);
                s.f = f;
            } else {
                const int dj0 = axis_window(f.x, 0, wgt), di0 = axis_window(f.y, 0, w0);
                s.dx0 = dj0, s.dy0 = di0;
                // every window of the block inside the guide image?  first window start >= 0, last window end <= n - 1
                const int xl = min(xb + RTX * 4, W) - 4, yl = min(yb + RTY * 2, H) - 2;
                s.fast = ((xb >> 1) + dj0 >= 0) && ((xl >> 1) + dj0 + 4 <= w - 1) && ((yb >> 1) + di0 >= 0) &&
                         ((yl >> 1) + di0 + 3 <= h - 1);
            }
        }
        __syncthreads();
    }
    if (x0 >= W || y0 >= H) return;
    const unsigned plane = (unsigned)H * (unsigned)W, lplane = (unsigned)h * (unsigned)w;
    if (!uniform || !s.fast) {
#pragma unroll 1
        for (int i = 0; i < 2; ++i)
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                const int x = x0 + j, y = y0 + i;
                if (x >= W || y >= H) continue;
                const int py = y / ts, px = x / ts;
                const float2 f = uniform ? s.f : __ldg(fl + (size_t)py * nx + px);
                const float S = uniform ? s.S : tile_S(fl, py, px, ny, nx, p);
                const Axis ay = dodgson_axis(y, f.y, h), ax = dodgson_axis(x, f.x, w);
                R[(size_t)y * W + x] = robustness_pixel(comp_lr, ref_means, terms, H, W, x, y, ay, ax, S, p.t);
            }
        return;
    }
    // 32-bit element offsets (one IMAD.WIDE per row pointer, immediates for the columns); 7 planes of a 50 MP frame fit
    const float *win = comp_lr + (((y0 >> 1) + s.dy0) * w + ((x0 >> 1) + s.dx0));
    const int o = y0 * W + x0, iplane = H * W, ilplane = h * w;
    float d_sq[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    unsigned finite = 0;   // bit i*4+j: reference mean of channel 0 is finite (not in the +inf band)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float col[2][5] = {{0.f, 0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f, 0.f}};
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const float *row = win + (c * ilplane + m * w);
            const float w0 = s.wy[0][m], w1 = s.wy[1][m];
#pragma unroll
            for (int n = 0; n < 5; ++n) {
                const float v = __ldg(row + n);
                col[0][n] = fmaf(v, w0, col[0][n]);
                col[1][n] = fmaf(v, w1, col[1][n]);
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float4 rm = __ldg(reinterpret_cast<const float4 *>(ref_means + (o + c * iplane + i * W)));
            const float4 dt = __ldg(reinterpret_cast<const float4 *>(terms + (o + c * iplane + i * W)));
            const float rmv[4] = {rm.x, rm.y, rm.z, rm.w}, dtv[4] = {dt.x, dt.y, dt.z, dt.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float cm = 0.f;
#pragma unroll
                for (int n = 0; n < 5; ++n) cm = fmaf(col[i][n], s.wx[j][n], cm);
                d_sq[i][j] += shrunk_dist(rmv[j], cm, dtv[j]);
                if (c == 0 && isfinite(rmv[j])) finite |= 1u << (i * 4 + j);
            }
        }
    }
    const float S = s.S;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float4 sg = __ldg(reinterpret_cast<const float4 *>(terms + (o + 3 * iplane + i * W)));
        const float sgv[4] = {sg.x, sg.y, sg.z, sg.w};
        float out[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) out[j] = ((finite >> (i * 4 + j)) & 1u) ? robustness_finish(d_sq[i][j], sgv[j], S, p.t) : 0.f;
        *reinterpret_cast<float4 *>(R + (o + i * W)) = make_float4(out[0], out[1], out[2], out[3]);
    }
}

__global__ void __launch_bounds__(RBX *RBY) local_min5_kernel(const float *__restrict__ R, int H, int W, float *__restrict__ r,
                                                              double *__restrict__ acc_rob) {
    __shared__ float s[RBY + 4][RBX + 4];
    const int x0 = blockIdx.x * RBX, y0 = blockIdx.y * RBY;
    for (int ly = threadIdx.y; ly < RBY + 4; ly += RBY) {
        const int gy = min(max(y0 + ly - 2, 0), H - 1);
        for (int lx = threadIdx.x; lx < RBX + 4; lx += RBX) {
            const int gx = min(max(x0 + lx - 2, 0), W - 1);
            s[ly][lx] = __ldg(R + (size_t)gy * W + gx);
        }
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= W || y >= H) return;
    float m = INFINITY;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const float *q = &s[threadIdx.y + i][threadIdx.x];
        m = fminf(m, fminf(fminf(fminf(q[0], q[1]), fminf(q[2], q[3])), q[4]));
    }
    const size_t o = (size_t)y * W + x;
    r[o] = m;
    if (acc_rob) acc_rob[o] += (double)m;
}

// Vectorised variant (W % 4 == 0, 16-byte aligned rows): one thread = 4 consecutive pixels x LMR rows.  Each input row
// is read once as three float4 (columns x-4 .. x+7, of which x-2 .. x+5 are used), reduced horizontally to the four
// 5-wide minima, and a rolling window of the last five row results gives the vertical minimum: 3 LDG.128 per 4
// output pixels and row instead of 25 shared-memory reads per pixel.  Minimum is order-independent, so the result is
// bit-identical to local_min5_kernel.
constexpr int LMR = 8;
__global__ void __launch_bounds__(RBX *RBY) local_min5_vec4_kernel(const float *__restrict__ R, int H, int W, float *__restrict__ r,
                                                                   double *__restrict__ acc_rob) {
    const int x = (blockIdx.x * RBX + threadIdx.x) * 4, y0 = (blockIdx.y * RBY + threadIdx.y) * LMR;
    if (x >= W || y0 >= H) return;
    const bool inner = x >= 4 && x + 8 <= W;
    float4 win[4];   // horizontal minima of the previous four rows
#pragma unroll
    for (int k = 0; k < LMR + 4; ++k) {
        const int yy = min(max(y0 + k - 2, 0), H - 1);                     // clamp, robustness.py:680-681
        const float *row = R + (size_t)yy * W;
        float v[8];
        if (inner) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(row + x - 4));
            const float4 b = __ldg(reinterpret_cast<const float4 *>(row + x));
            const float4 c = __ldg(reinterpret_cast<const float4 *>(row + x + 4));
            v[0] = a.z, v[1] = a.w, v[2] = b.x, v[3] = b.y, v[4] = b.z, v[5] = b.w, v[6] = c.x, v[7] = c.y;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __ldg(row + min(max(x + j - 2, 0), W - 1));
        }
        const float m12 = fminf(v[1], v[2]), m34 = fminf(v[3], v[4]), m56 = fminf(v[5], v[6]);
        const float m1234 = fminf(m12, m34), m3456 = fminf(m34, m56);
        float4 h;
        h.x = fminf(v[0], m1234), h.y = fminf(m1234, v[5]), h.z = fminf(v[2], m3456), h.w = fminf(m3456, v[7]);
        if (k >= 4) {
            const int y = y0 + k - 4;
            if (y < H) {
                float4 m;
                m.x = fminf(fminf(fminf(win[0].x, win[1].x), fminf(win[2].x, win[3].x)), h.x);
                m.y = fminf(fminf(fminf(win[0].y, win[1].y), fminf(win[2].y, win[3].y)), h.y);
                m.z = fminf(fminf(fminf(win[0].z, win[1].z), fminf(win[2].z, win[3].z)), h.z);
                m.w = fminf(fminf(fminf(win[0].w, win[1].w), fminf(win[2].w, win[3].w)), h.w);
                const size_t o = (size_t)y * W + x;
                *reinterpret_cast<float4 *>(r + o) = m;
                if (acc_rob) {
                    double2 *a = reinterpret_cast<double2 *>(acc_rob + o);
                    double2 a0 = a[0], a1 = a[1];
                    a0.x += (double)m.x, a0.y += (double)m.y, a1.x += (double)m.z, a1.y += (double)m.w;
                    a[0] = a0, a[1] = a1;
                }
            }
        }
        win[0] = win[1], win[1] = win[2], win[2] = win[3], win[3] = h;
    }
}

}  // namespace hhsr

using namespace hhsr;

extern "C" int hhsr_guide_stats(const float *raw, int H, int W, const int *cfa_host, const double *wb_host,
                                float *means, float *vars, hhsr_stream_t stream) {
    HHSR_REQUIRE(raw && cfa_host && wb_host && means, "null pointer");
    HHSR_REQUIRE(H >= 2 && W >= 2 && W % 2 == 0, "frame must be at least 2x2 with an even width");
    HHSR_REQUIRE((uintptr_t)raw % 8 == 0, "raw must be 8-byte aligned");
    GuideParams p;
    p.cfa = pack_cfa(cfa_host);
    for (int k = 0; k < 4; ++k) HHSR_REQUIRE(cfa_host[k] >= 0 && cfa_host[k] <= 2, "cfa entries must be 0, 1 or 2");
    for (int c = 0; c < 3; ++c) {
        HHSR_REQUIRE(wb_host[c] != 0.0, "white balance gains of the three colour channels must be non-zero");
        p.inv_wb[c] = 1.0 / wb_host[c];
    }
    const int h = H / 2, w = W / 2;
    dim3 block(RBX, RBY), grid(ceil_div(w, RBX), ceil_div(h, RBY));
    guide_stats_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(raw, W, h, w, p, means, vars);
    return launch_status("guide_stats");
}

extern "C" int hhsr_guide_image(const float *raw, int H, int W, const int *cfa_host, const double *wb_host, float *guide,
                                hhsr_stream_t stream) {
    HHSR_REQUIRE(raw && cfa_host && wb_host && guide, "null pointer");
    HHSR_REQUIRE(H >= 2 && W >= 2 && W % 2 == 0, "frame must be at least 2x2 with an even width");
    HHSR_REQUIRE((uintptr_t)raw % 8 == 0, "raw must be 8-byte aligned");
    GuideParams p;
    p.cfa = pack_cfa(cfa_host);
    for (int k = 0; k < 4; ++k) HHSR_REQUIRE(cfa_host[k] >= 0 && cfa_host[k] <= 2, "cfa entries must be 0, 1 or 2");
    for (int c = 0; c < 3; ++c) {
        HHSR_REQUIRE(wb_host[c] != 0.0, "white balance gains of the three colour channels must be non-zero");
        p.inv_wb[c] = 1.0 / wb_host[c];
    }
    const int h = H / 2, w = W / 2;
    dim3 block(RBX, RBY), grid(ceil_div(w, RBX), ceil_div(h, RBY));
    guide_image_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(raw, W, h, w, p, guide);
    return launch_status("guide_image");
}

extern "C" int hhsr_local_stats(const float *guide, int channels, int h, int w, float *means, float *vars,
                                hhsr_stream_t stream) {
    HHSR_REQUIRE(guide && means && vars, "null pointer");
    HHSR_REQUIRE(h > 0 && w > 0 && channels > 0 && channels <= 65535, "non-positive size");
    dim3 block(RBX, RBY), grid(ceil_div(w, RBX), ceil_div(h, RBY), channels);
    local_stats_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(guide, h, w, means, vars);
    return launch_status("local_stats");
}

extern "C" int hhsr_upscale_warp_stats(const float *lr, int h, int w, const float *flow, int ny, int nx, int ts,
                                       float *hr, hhsr_stream_t stream) {
    HHSR_REQUIRE(lr && hr, "null pointer");
    HHSR_REQUIRE(h > 0 && w > 0, "non-positive size");
    HHSR_REQUIRE(flow == nullptr || (ts > 0 && ny * ts >= 2 * h && nx * ts >= 2 * w), "flow grid does not cover the frame");
    dim3 block(RBX, RBY), grid(ceil_div(2 * w, RBX), ceil_div(2 * h, RBY));
    upscale_warp_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(lr, h, w, flow, nx, ts > 0 ? ts : 1, hr);
    return launch_status("upscale_warp_stats");
}

extern "C" int hhsr_noise_table(const double *std_curve, const double *diff_curve, int n_curve, float *table,
                                hhsr_stream_t stream) {
    HHSR_REQUIRE(std_curve && diff_curve && table, "null pointer");
    HHSR_REQUIRE(n_curve > 0, "empty noise curves");
    HHSR_REQUIRE((uintptr_t)table % 8 == 0, "table must be 8-byte aligned");
    noise_table_kernel<<<ceil_div(n_curve, 256), 256, 0, (cudaStream_t)stream>>>(std_curve, diff_curve, n_curve,
                                                                                reinterpret_cast<float2 *>(table));
    return launch_status("noise_table");
}

extern "C" int hhsr_robustness_ref_terms(const float *ref_means, const float *ref_vars, int H, int W,
                                         const float *noise_table, int n_curve, float *terms, hhsr_stream_t stream) {
    HHSR_REQUIRE(ref_means && ref_vars && noise_table && terms, "null pointer");
    HHSR_REQUIRE(H > 0 && W > 0 && n_curve > 0, "non-positive size");
    HHSR_REQUIRE((uintptr_t)noise_table % 8 == 0, "noise table must be 8-byte aligned");
    const size_t plane = (size_t)H * W;
    size_t blocks = (plane + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    ref_noise_terms_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(ref_means, ref_vars, plane,
                                                                            reinterpret_cast<const float2 *>(noise_table),
                                                                            n_curve, terms);
    return launch_status("robustness_ref_terms");
}

extern "C" int hhsr_ref_stats_terms(const float *guide_means, const float *guide_vars, int h, int w, const float *noise_table,
                                    int n_curve, float *ref_means, float *ref_vars, float *terms, hhsr_stream_t stream) {
    HHSR_REQUIRE(guide_means && guide_vars && noise_table && ref_means && terms, "null pointer");
    HHSR_REQUIRE(h >= 1 && w >= 1 && n_curve > 0, "non-positive size");
    HHSR_REQUIRE((uintptr_t)noise_table % 8 == 0 && (uintptr_t)ref_means % 8 == 0 && (uintptr_t)terms % 8 == 0 &&
                     (uintptr_t)ref_vars % 8 == 0, "outputs and noise table must be 8-byte aligned");
    dim3 block(RBX, RBY), grid(ceil_div(w, RBX), ceil_div(h, RBY));
    ref_stats_terms_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(guide_means, guide_vars, h, w,
                                                                    reinterpret_cast<const float2 *>(noise_table), n_curve,
                                                                    ref_means, ref_vars, terms);
    return launch_status("ref_stats_terms");
}

extern "C" int hhsr_robustness(const float *comp_means_lr, const float *ref_means, const float *ref_terms, int H, int W,
                               const float *flow, int ny, int nx, int ts, double t, double s1, double s2, double Mt,
                               float *R, int flags, hhsr_stream_t stream) {
    HHSR_REQUIRE(comp_means_lr && ref_means && ref_terms && flow && R, "null pointer");
    HHSR_REQUIRE((flags & ~HHSR_ROBUSTNESS_GENERIC) == 0, "unknown robustness flag");
    HHSR_REQUIRE(H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, "frame sides must be even");
    HHSR_REQUIRE(ts > 0 && ny * ts >= H && nx * ts >= W, "flow grid does not cover the frame");
    HHSR_REQUIRE((uintptr_t)ref_means % 16 == 0 && (uintptr_t)ref_terms % 16 == 0 && (uintptr_t)R % 16 == 0,
                 "ref_means / ref_terms / R must be 16-byte aligned");
    RobParams p{t, s1, s2, Mt, (flags & HHSR_ROBUSTNESS_GENERIC) ? 1 : 0};
    dim3 block(RTX, RTY), grid(ceil_div(W, RTX * 4), ceil_div(H, RTY * 2));
    robustness_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(comp_means_lr, ref_means, ref_terms, H, W, flow, ny, nx, ts, p, R);
    return launch_status("robustness");
}

extern "C" int hhsr_local_min5(const float *R, int H, int W, float *r, double *acc_rob, hhsr_stream_t stream) {
    HHSR_REQUIRE(R && r, "null pointer");
    HHSR_REQUIRE(H > 0 && W > 0, "non-positive size");
    dim3 block(RBX, RBY);
    if (W % 4 == 0 && (uintptr_t)R % 16 == 0 && (uintptr_t)r % 16 == 0 && (uintptr_t)acc_rob % 16 == 0) {
        dim3 grid(ceil_div(W, RBX * 4), ceil_div(H, RBY * LMR));
        local_min5_vec4_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(R, H, W, r, acc_rob);
    } else {
        dim3 grid(ceil_div(W, RBX), ceil_div(H, RBY));
        local_min5_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(R, H, W, r, acc_rob);
    }
    return launch_status("local_min5");
}
