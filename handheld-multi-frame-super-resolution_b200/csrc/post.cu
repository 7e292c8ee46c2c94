// Output side of the pipeline for sm_100a: everything the reference does to the merged
// image on the HOST before it is saved, moved onto the device so that only the finished (optionally 8/16-bit) image
// crosses PCIe.
//
// Replaces handheld_super_resolution/raw2rgb.py:212-250 (postprocess: colour matrix, unsharp mask, devignetting,
// gamma), the quantisation of run_handheld.py:132-150 (nan_to_num, clip, img_as_ubyte) and
// handheld_super_resolution/utils_image.py:174-309 (frame_count_denoising_gauss / _median).
//
// The unsharp mask is skimage.filters.unsharp_mask(img, radius, amount, channel_axis=2, preserve_range=True), i.e.
// img + (img - G_sigma * img) * amount with G the separable scipy.ndimage.gaussian_filter(sigma=radius, truncate=4,
// mode='reflect'): axis 0 then axis 1, float64 accumulation, the result of each pass rounded to float32 (scipy keeps
// the input dtype between the passes).  Both passes follow that arithmetic; NaN pixels spread through the blur exactly
// like in the reference (0 * NaN = NaN).
#include "common.cuh"

namespace hhsr {

constexpr int kPostMaxTaps = 129;      // radius <= 64 (sigma <= 16)
struct PostTaps {
    double w[kPostMaxTaps];
    int radius;
};

__device__ __forceinline__ int reflect_idx(int i, int n) {   // scipy 'reflect': (d c b a | a b c d | d c b a)
    while (i < 0 || i >= n) i = (i < 0) ? (-i - 1) : (2 * n - i - 1);
    return i;
}

__device__ __forceinline__ float clip01_keep_nan(float v) { return (v != v) ? v : fminf(fmaxf(v, 0.f), 1.f); }   // np.clip

// colour matrix (raw2rgb.py:139-146 apply_ccm + the clip of :226): float32 dot products in channel order
__global__ void post_ccm_clip_kernel(float *__restrict__ img, size_t n_px, float m00, float m01, float m02, float m10, float m11,
                                     float m12, float m20, float m21, float m22) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += stride) {
        const float r = img[3 * i], g = img[3 * i + 1], b = img[3 * i + 2];
        const float o0 = __fmaf_rn(m02, b, __fmaf_rn(m01, g, m00 * r));
        const float o1 = __fmaf_rn(m12, b, __fmaf_rn(m11, g, m10 * r));
        const float o2 = __fmaf_rn(m22, b, __fmaf_rn(m21, g, m20 * r));
        img[3 * i] = clip01_keep_nan(o0), img[3 * i + 1] = clip01_keep_nan(o1), img[3 * i + 2] = clip01_keep_nan(o2);
    }
}

// first pass of the Gaussian (axis 0): the interleaved [H][W][3] image is [H][3W] floats and every column is filtered
// independently.  One thread = 4 consecutive floats of one output row.
__global__ void __launch_bounds__(256) post_blur_cols_kernel(const float *__restrict__ src, int H, int Wc, const __grid_constant__ PostTaps t,
                                                             float *__restrict__ dst) {
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= Wc || y >= H) return;
    const int R = t.radius;
    if (x + 4 <= Wc && (Wc & 3) == 0) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        for (int k = 0; k <= 2 * R; ++k) {
            const int yy = reflect_idx(y + k - R, H);
            const float4 v = __ldg(reinterpret_cast<const float4 *>(src + (size_t)yy * Wc + x));
            const double w = t.w[k];
            a0 = fma(w, (double)v.x, a0), a1 = fma(w, (double)v.y, a1), a2 = fma(w, (double)v.z, a2), a3 = fma(w, (double)v.w, a3);
        }
        *reinterpret_cast<float4 *>(dst + (size_t)y * Wc + x) = make_float4((float)a0, (float)a1, (float)a2, (float)a3);
    } else {
        for (int j = 0; j < 4 && x + j < Wc; ++j) {
            double a = 0.0;
            for (int k = 0; k <= 2 * R; ++k) a = fma(t.w[k], (double)__ldg(src + (size_t)reflect_idx(y + k - R, H) * Wc + x + j), a);
            dst[(size_t)y * Wc + x + j] = (float)a;
        }
    }
}

struct PostParams {
    int sharpen;        // 1: tmp holds the column-blurred image; finish the blur along x and apply the unsharp mask
    float amount;
    int devignette;     // raw2rgb.py:203-210
    int gamma;          // 1: x ** inv_gamma after clipping (raw2rgb.py:143-146)
    float inv_gamma;
    int out_kind;       // 0 float32 (clip only, NaN kept: what process() returns), 1 uint8, 2 uint16 (nan_to_num + clip + rint)
};

// second pass of the Gaussian (axis 1) fused with everything that follows: unsharp mask, devignetting, clip, gamma, clip,
// quantisation.  One thread = one pixel (3 channels).
__global__ void __launch_bounds__(256) post_finish_kernel(const float *__restrict__ img, const float *__restrict__ tmp, int H, int W,
                                                          const __grid_constant__ PostTaps t, const __grid_constant__ PostParams p,
                                                          void *__restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const size_t o = ((size_t)y * W + x) * 3;
    float v[3] = {img[o], img[o + 1], img[o + 2]};
    if (p.sharpen) {
        const int R = t.radius;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        const float *row = tmp + (size_t)y * W * 3;
        for (int k = 0; k <= 2 * R; ++k) {
            const float *q = row + 3 * reflect_idx(x + k - R, W);
            const double w = t.w[k];
            a0 = fma(w, (double)__ldg(q), a0), a1 = fma(w, (double)__ldg(q + 1), a1), a2 = fma(w, (double)__ldg(q + 2), a2);
        }
        const float b[3] = {(float)a0, (float)a1, (float)a2};
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = __fadd_rn(v[c], __fmul_rn(__fsub_rn(v[c], b[c]), p.amount));   // image + (image - blurred) * amount
    }
    double d[3] = {(double)v[0], (double)v[1], (double)v[2]};
    if (p.devignette) {   // float64 like the reference (np.linspace / np.outer are float64): (2 - cos(f)^4) * image
        const double hw = (double)H / (double)W * 1.5707963267948966;
        const double fy = (H > 1) ? fabs(-hw + (2.0 * hw) * (double)y / (double)(H - 1)) : fabs(-hw);
        const double fx = (W > 1) ? fabs(-1.5707963267948966 + 3.141592653589793 * (double)x / (double)(W - 1)) : 1.5707963267948966;
        const double cs = cos(fy * fx);
        const double gain = 2.0 - (cs * cs) * (cs * cs);
#pragma unroll
        for (int c = 0; c < 3; ++c) d[c] = gain * d[c];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float f = (float)d[c];                      // the reference keeps float64 after devignetting; the final image is
        f = clip01_keep_nan(f);                     // clipped to [0,1] where float32 rounding of a float64 value is harmless
        if (p.gamma) f = clip01_keep_nan(powf(f, p.inv_gamma));
        v[c] = f;
    }
    if (p.out_kind == 0) {
        float *q = reinterpret_cast<float *>(out) + o;
        q[0] = v[0], q[1] = v[1], q[2] = v[2];
    } else {
        // run_handheld.py:132-133,150: nan_to_num, clip to [0,1], img_as_ubyte = rint(x * 255) in float32
        const float top = (p.out_kind == 1) ? 255.f : 65535.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float f = (v[c] != v[c]) ? 0.f : fminf(fmaxf(v[c], 0.f), 1.f);
            const float qv = rintf(__fmul_rn(f, top));
            if (p.out_kind == 1)
                reinterpret_cast<unsigned char *>(out)[o + c] = (unsigned char)qv;
            else
                reinterpret_cast<unsigned short *>(out)[o + c] = (unsigned short)qv;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Frame-count-aware denoisers (utils_image.py:174-309), applied to the merged image where few frames were accumulated.
// Upstream they cannot run: the host wrappers read `config.mode` / `config.scale` from the denoiser's own sub-config
// (utils_image.py:177-178, 243-244), and the Gaussian kernel iterates `range(-t, t+1)` over a float t = 3*sigma
// (:210-215).  Semantics kept here: mode/scale come from the main configuration, t = ceil(3*sigma); everything else
// — the (y - 0.5)/(2*scale) lookup of the accumulated robustness, sigma / radius laws, float64 weights and sums, the
// literal bubble sort and the upper median buffer[k//2] — follows the kernels line by line.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int py_round_to_int(double v) { return (int)llrint(v); }   // round half to even, like Python / Numba

__global__ void __launch_bounds__(256) frame_count_gauss_kernel(const float *__restrict__ noisy, int Hs, int Ws, const double *__restrict__ r_acc,
                                                                int H, int W, double scale, double sigma_max, double max_frame_count,
                                                                float *__restrict__ denoised) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= Ws || y >= Hs) return;
    const int yg = min(max(py_round_to_int(((double)y - 0.5) / (2.0 * scale)), 0), H - 1);     // utils_image.py:204-205
    const int xg = min(max(py_round_to_int(((double)x - 0.5) / (2.0 * scale)), 0), W - 1);
    const double r = fmin(r_acc[(size_t)yg * W + xg], max_frame_count);
    const double sigma = sigma_max * (max_frame_count - r) / max_frame_count;                    // denoise_power_gauss, :228-231
    const int t = (int)ceil(3.0 * sigma);
    const size_t o = ((size_t)y * Ws + x) * 3;
    if (t <= 0) {
        denoised[o] = noisy[o], denoised[o + 1] = noisy[o + 1], denoised[o + 2] = noisy[o + 2];
        return;
    }
    const double inv = 1.0 / (2.0 * sigma * sigma);
    double num[3] = {0.0, 0.0, 0.0}, den = 0.0;
    for (int i = -t; i <= t; ++i) {
        const int yy = y + i;
        if (yy < 0 || yy >= Hs) continue;
        for (int j = -t; j <= t; ++j) {
            const int xx = x + j;
            if (xx < 0 || xx >= Ws) continue;
            const double w = exp(-(double)(j * j + i * i) * inv);
            const float *q = noisy + ((size_t)yy * Ws + xx) * 3;
            num[0] += w * (double)__ldg(q), num[1] += w * (double)__ldg(q + 1), num[2] += w * (double)__ldg(q + 2);
            den += w;
        }
    }
    denoised[o] = (float)(num[0] / den), denoised[o + 1] = (float)(num[1] / den), denoised[o + 2] = (float)(num[2] / den);
}

constexpr int kMedianMaxRadius = 7;       // (2*7+1)^2 = 225 <= the reference's 256-entry buffer (:281); larger radii overflow it upstream
__global__ void __launch_bounds__(128) frame_count_median_kernel(const float *__restrict__ noisy, int Hs, int Ws, const double *__restrict__ r_acc,
                                                                 int H, int W, double scale, double radius_max, double max_frame_count,
                                                                 float *__restrict__ denoised) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, c = blockIdx.z;
    if (x >= Ws || y >= Hs) return;
    const int yg = min(max(py_round_to_int(((double)y - 0.5) / (2.0 * scale)), 0), H - 1);
    const int xg = min(max(py_round_to_int(((double)x - 0.5) / (2.0 * scale)), 0), W - 1);
    const double r = fmin(r_acc[(size_t)yg * W + xg], max_frame_count);
    int radius = py_round_to_int(radius_max * (max_frame_count - r) / max_frame_count);         // denoise_power_median, :303-306
    radius = min(min(14, radius), kMedianMaxRadius);
    const size_t o = ((size_t)y * Ws + x) * 3 + c;
    if (radius <= 0) {
        denoised[o] = noisy[o];
        return;
    }
    float buf[(2 * kMedianMaxRadius + 1) * (2 * kMedianMaxRadius + 1)];
    int k = 0;
    for (int i = -radius; i <= radius; ++i)
        for (int j = -radius; j <= radius; ++j) {
            const int xx = x + j, yy = y + i;
            if (yy >= 0 && yy < Hs && xx >= 0 && xx < Ws) buf[k++] = __ldg(noisy + ((size_t)yy * Ws + xx) * 3 + c);
        }
    for (int i = 0; i < k - 1; ++i)                      // bubble_sort, :308-315 (NaN compares false: never swapped, like upstream)
        for (int j = 0; j < k - i - 1; ++j)
            if (buf[j] > buf[j + 1]) {
                const float s = buf[j];
                buf[j] = buf[j + 1], buf[j + 1] = s;
            }
    denoised[o] = buf[k / 2];
}

static int fill_taps(PostTaps &t, const double *taps_host, int radius) {
    HHSR_REQUIRE(taps_host && radius >= 0 && 2 * radius + 1 <= kPostMaxTaps, "Gaussian radius out of range (<= 64)");
    t.radius = radius;
    for (int i = 0; i < 2 * radius + 1; ++i) t.w[i] = taps_host[i];
    return 0;
}

}  // namespace hhsr

using namespace hhsr;

extern "C" int hhsr_post_ccm_clip(float *img, size_t n_px, const float *ccm_host, hhsr_stream_t stream) {
    HHSR_REQUIRE(img && ccm_host && n_px > 0, "null pointer or empty image");
    size_t blocks = (n_px + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    const float *m = ccm_host;
    post_ccm_clip_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(img, n_px, m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8]);
    return launch_status("post_ccm_clip");
}

extern "C" int hhsr_post_blur_cols(const float *img, int H, int W, const double *taps_host, int radius, float *tmp,
                                   hhsr_stream_t stream) {
    HHSR_REQUIRE(img && tmp, "null pointer");
    HHSR_REQUIRE(H > 0 && W > 0, "non-positive size");
    HHSR_REQUIRE((uintptr_t)img % 16 == 0 && (uintptr_t)tmp % 16 == 0, "images must be 16-byte aligned");
    PostTaps t;
    if (int e = fill_taps(t, taps_host, radius)) return e;
    const int Wc = 3 * W;
    dim3 block(32, 8), grid(ceil_div(ceil_div(Wc, 4), 32), ceil_div(H, 8));
    post_blur_cols_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(img, H, Wc, t, tmp);
    return launch_status("post_blur_cols");
}

extern "C" int hhsr_post_finish(const float *img, const float *tmp, int H, int W, const double *taps_host, int radius, float amount,
                                int devignette, float inv_gamma, int out_kind, void *out, hhsr_stream_t stream) {
    HHSR_REQUIRE(img && out, "null pointer");
    HHSR_REQUIRE(H > 0 && W > 0, "non-positive size");
    HHSR_REQUIRE(out_kind >= 0 && out_kind <= 2, "out_kind must be 0 (float32), 1 (uint8) or 2 (uint16)");
    HHSR_REQUIRE(inv_gamma >= 0.f, "inv_gamma must be >= 0 (0: no gamma compression)");
    PostTaps t;
    t.radius = 0;
    if (tmp != nullptr)
        if (int e = fill_taps(t, taps_host, radius)) return e;
    PostParams p{tmp != nullptr ? 1 : 0, amount, devignette ? 1 : 0, inv_gamma > 0.f ? 1 : 0, inv_gamma, out_kind};
    dim3 block(32, 8), grid(ceil_div(W, 32), ceil_div(H, 8));
    post_finish_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(img, tmp, H, W, t, p, out);
    return launch_status("post_finish");
}

extern "C" int hhsr_frame_count_denoise_gauss(const float *img, int Hs, int Ws, const double *acc_rob, int H, int W, double scale,
                                              double sigma_max, double max_frame_count, float *out, hhsr_stream_t stream) {
    HHSR_REQUIRE(img && acc_rob && out && img != out, "null pointer, or in-place call");
    HHSR_REQUIRE(Hs > 0 && Ws > 0 && H > 0 && W > 0, "non-positive size");
    HHSR_REQUIRE(scale >= 1.0 && sigma_max >= 0.0 && sigma_max <= 16.0 && max_frame_count > 0.0, "scale >= 1, 0 <= sigma_max <= 16, max_frame_count > 0 required");
    dim3 block(32, 8), grid(ceil_div(Ws, 32), ceil_div(Hs, 8));
    frame_count_gauss_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(img, Hs, Ws, acc_rob, H, W, scale, sigma_max, max_frame_count, out);
    return launch_status("frame_count_denoise_gauss");
}

extern "C" int hhsr_frame_count_denoise_median(const float *img, int Hs, int Ws, const double *acc_rob, int H, int W, double scale,
                                               double radius_max, double max_frame_count, float *out, hhsr_stream_t stream) {
    HHSR_REQUIRE(img && acc_rob && out && img != out, "null pointer, or in-place call");
    HHSR_REQUIRE(Hs > 0 && Ws > 0 && H > 0 && W > 0, "non-positive size");
    HHSR_REQUIRE(scale >= 1.0 && radius_max >= 0.0 && max_frame_count > 0.0, "scale >= 1, radius_max >= 0, max_frame_count > 0 required");
    if (radius_max > (double)kMedianMaxRadius + 0.5)
        return unsupported("median radius above 7 overflows the reference's 256-entry window buffer (utils_image.py:281)");
    dim3 block(32, 4), grid(ceil_div(Ws, 32), ceil_div(Hs, 4), 3);
    frame_count_median_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(img, Hs, Ws, acc_rob, H, W, scale, radius_max, max_frame_count, out);
    return launch_status("frame_count_denoise_median");
}
