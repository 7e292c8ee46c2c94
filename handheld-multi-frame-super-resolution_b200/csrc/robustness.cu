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
    double wb[3];
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
        const double v = (double)q[k] / p.wb[c];
        if (c == 1)
            g += v;
        else
            out[c] = (float)v;
    }
    out[1] = (float)(g / 2.0);
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
        const double m = (double)s1 / 9.0;
        const size_t o = ((size_t)c * h + y) * w + x;
        means[o] = (float)m;
        if (vars) vars[o] = (float)((double)s2 / 9.0 - m * m);
    }
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
    int n_curve;
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

// Noise model + threshold for one pixel given its warped comp means (robustness.py:452-462, 504-533, 626-639)
__device__ __forceinline__ float robustness_value(const float (&cm)[3], float rm0, const float *__restrict__ ref_means,
                                                  const float *__restrict__ ref_vars, unsigned o, unsigned plane,
                                                  const float2 *__restrict__ table, const RobParams &p, float S) {
    float sigma_sq = 0.f, d_sq = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float brightness = c == 0 ? rm0 : __ldg(ref_means + o + c * plane);
        int id = (int)llrint(1000.0 * (double)brightness);                  // robustness.py:519 (float64 product)
        id = min(max(id, 0), p.n_curve - 1);
        const float2 t = __ldg(table + id);
        const float sigma_p_sq = __ldg(ref_vars + o + c * plane);
        sigma_sq += fmaxf(sigma_p_sq, t.x);                                 // max(sigma_p^2, sigma_t^2), :524
        const float d_p = fabsf(brightness - cm[c]);                        // :462
        const float d_p_sq = d_p * d_p;
        const float shrink = __fdividef(d_p_sq, d_p_sq + t.y);
        d_sq += d_p_sq * shrink * shrink;
    }
    const float e = expf(-__fdividef(d_sq, sigma_sq));                      // math.exp(float32), :638
    const float v = (float)((double)(S * e) - p.t);
    return fminf(1.f, fmaxf(0.f, v));                                       // NaN -> 0
}

// One thread per raw pixel.  When the tile size is a multiple of 32 a 32x8 block lies inside one flow tile: flow,
// S and the Dodgson axes (8 row axes + 32 column axes instead of 2 per thread) are then computed once per block.
__global__ void __launch_bounds__(RBX *RBY) robustness_kernel(const float *__restrict__ comp_lr, const float *__restrict__ ref_means,
                                                              const float *__restrict__ ref_vars, int H, int W,
                                                              const float *__restrict__ flow, int ny, int nx, int ts,
                                                              const float2 *__restrict__ table, RobParams p,
                                                              float *__restrict__ R) {
    const int x = blockIdx.x * RBX + threadIdx.x, y = blockIdx.y * RBY + threadIdx.y;
    const float2 *fl = reinterpret_cast<const float2 *>(flow);
    const int h = H / 2, w = W / 2;
    __shared__ float s_S;
    __shared__ Axis s_ay[RBY], s_ax[RBX];
    const bool uniform = (ts % RBX) == 0;
    Axis ay, ax;
    float S;
    if (uniform) {
        const int tid = threadIdx.y * RBX + threadIdx.x;
        const int py = (blockIdx.y * RBY) / ts, px = (blockIdx.x * RBX) / ts;
        const float2 f = __ldg(fl + (size_t)py * nx + px);
        if (tid < RBX)
            s_ax[tid] = dodgson_axis(min(blockIdx.x * RBX + tid, W - 1), f.x, w);
        else if (tid < RBX + RBY)
            s_ay[tid - RBX] = dodgson_axis(min(blockIdx.y * RBY + tid - RBX, H - 1), f.y, h);
        else if (tid == RBX + RBY)
            s_S = tile_S(fl, py, px, ny, nx, p);
        __syncthreads();
        if (x >= W || y >= H) return;
        ay = s_ay[threadIdx.y], ax = s_ax[threadIdx.x], S = s_S;
    } else {
        if (x >= W || y >= H) return;
        const int py = y / ts, px = x / ts;
        const float2 f = __ldg(fl + (size_t)py * nx + px);
        S = tile_S(fl, py, px, ny, nx, p);
        ay = dodgson_axis(y, f.y, h), ax = dodgson_axis(x, f.x, w);
    }
    const unsigned plane = (unsigned)H * (unsigned)W, lplane = (unsigned)h * (unsigned)w, o = (unsigned)y * (unsigned)W + x;
    float out = 0.f;   // any non-finite statistic ends as clamp(NaN) = 0 in the reference
    const float rm0 = __ldg(ref_means + o);
    if (ay.ok && ax.ok && isfinite(rm0)) {
        float buf[3] = {0.f, 0.f, 0.f}, wacc;
        if (ay.i[2] == ay.i[0] + 2 && ax.i[2] == ax.i[0] + 2) {
            // interior: the 3x3 taps are contiguous -> one base pointer per (plane,row), immediate column offsets
            const float *q0 = comp_lr + ((unsigned)ay.i[0] * (unsigned)w + (unsigned)ax.i[0]);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float *qc = q0 + c * lplane;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const float *row = qc + i * w;
                    const float r3 = fmaf(__ldg(row), ax.w[0], fmaf(__ldg(row + 1), ax.w[1], __ldg(row + 2) * ax.w[2]));
                    buf[c] = fmaf(r3, ay.w[i], buf[c]);
                }
            }
            wacc = (ay.w[0] + ay.w[1] + ay.w[2]) * (ax.w[0] + ax.w[1] + ax.w[2]);
        } else {
            wacc = 0.f;
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
        }
        const float inv_w = __fdividef(1.0f, wacc);
        const float cm[3] = {buf[0] * inv_w, buf[1] * inv_w, buf[2] * inv_w};
        out = robustness_value(cm, rm0, ref_means, ref_vars, o, plane, table, p, S);
    }
    R[o] = out;
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
    for (int c = 0; c < 3; ++c) p.wb[c] = wb_host[c];
    const int h = H / 2, w = W / 2;
    dim3 block(RBX, RBY), grid(ceil_div(w, RBX), ceil_div(h, RBY));
    guide_stats_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(raw, W, h, w, p, means, vars);
    return launch_status("guide_stats");
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

extern "C" int hhsr_robustness(const float *comp_means_lr, const float *ref_means, const float *ref_vars, int H,
                               int W, const float *flow, int ny, int nx, int ts, const float *noise_table, int n_curve,
                               double t, double s1, double s2, double Mt, float *R, hhsr_stream_t stream) {
    HHSR_REQUIRE(comp_means_lr && ref_means && ref_vars && flow && noise_table && R, "null pointer");
    HHSR_REQUIRE(H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, "frame sides must be even");
    HHSR_REQUIRE(ts > 0 && ny * ts >= H && nx * ts >= W, "flow grid does not cover the frame");
    HHSR_REQUIRE(n_curve > 0, "empty noise curves");
    HHSR_REQUIRE((uintptr_t)noise_table % 8 == 0, "noise table must be 8-byte aligned");
    RobParams p{t, s1, s2, Mt, n_curve};
    dim3 block(RBX, RBY), grid(ceil_div(W, RBX), ceil_div(H, RBY));
    robustness_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(comp_means_lr, ref_means, ref_vars, H, W, flow, ny, nx, ts,
                                                               reinterpret_cast<const float2 *>(noise_table), p, R);
    return launch_status("robustness");
}

extern "C" int hhsr_local_min5(const float *R, int H, int W, float *r, double *acc_rob, hhsr_stream_t stream) {
    HHSR_REQUIRE(R && r, "null pointer");
    HHSR_REQUIRE(H > 0 && W > 0, "non-positive size");
    dim3 block(RBX, RBY), grid(ceil_div(W, RBX), ceil_div(H, RBY));
    local_min5_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(R, H, W, r, acc_rob);
    return launch_status("local_min5");
}
