// Steering-kernel estimation (Alg. 5) for sm_100a: one fused kernel raw -> covariance.
//
// Replaces handheld_super_resolution/kernels.py:29-243 (estimate_kernels, cuda_estimate_kernel, compute_k),
// utils_image.py:117-170 (GAT), :346-357 (cuda_decimate_to_grey), the two F.conv2d gradient filters
// (kernels.py:97-112) and linalg.py:86-185 (closed-form 2x2 eigen-decomposition) — 3 kernels, 2 convolutions
// and 4 full-size temporaries in the reference.  HBM traffic: raw read once (4 B/px), covariance written once
// (16 B per Bayer quad).
//
// Arithmetic mirrors the compiled reference operation by operation (float64 where Numba promotes: GAT, the
// discriminant, A, D, k; float32 elsewhere) with the fused multiply-adds NVVM emits for it made explicit
// (structure-tensor accumulation, determinant, eigenvector norm) — matched to the B200 goldens to 1 ulp.
#include "common.cuh"

namespace hhsr {

constexpr int KBX = 32, KBY = 8;

struct KernelParams {
    double alpha, beta, k_detail, k_denoise, D_th, D_tr, k_stretch, k_shrink;
    double inv_D_tr, inv_k_shrink;   // host-side reciprocals (one float64 division per launch instead of per thread)
    int law;  // 0 hard threshold, 1 linear
};

// float64 square root to < 2 ulp from a float32 rsqrt seed and two Newton steps (10 instructions against ~30 for the
// IEEE sqrt sequence); every use below rounds the result to float32, which absorbs the difference.
__device__ __forceinline__ double sqrt_fast(double v) {
    if (!(v > 1e-30 && v < 1e30)) return sqrt(v);      // zero, tiny, huge, negative, NaN: the IEEE sequence
    double r = (double)rsqrtf((float)v);
    r = fma(0.5 * r, fma(-v * r, r, 1.0), r);          // rsqrt to ~46 bits
    const double s = v * r;
    return fma(fma(-s, s, v), 0.5 * r, s);
}

__device__ __forceinline__ float gat_px(float x, double alpha, double c0, double two_over_alpha) {
    double v = alpha * (double)x + c0;           // alpha*I + 3/8*alpha^2 + beta   (utils_image.py:165)
    v = (v > 0.0) ? v : 0.0;
    return (float)(two_over_alpha * sqrt_fast(v));
}

__global__ void __launch_bounds__(KBX *KBY) estimate_kernels_kernel(const float *__restrict__ raw, int H, int W, int h,
                                                                    int w, KernelParams p, float *__restrict__ covs) {
    __shared__ float g[KBY + 2][KBX + 2];   // decimated variance-stabilised grey, 1-quad halo
    const int x0 = blockIdx.x * KBX, y0 = blockIdx.y * KBY;
    const double c0 = 3.0 / 8.0 * p.alpha * p.alpha + p.beta, toa = 2.0 / p.alpha;
    for (int t = threadIdx.y * KBX + threadIdx.x; t < (KBY + 2) * (KBX + 2); t += KBX * KBY) {
        const int ly = t / (KBX + 2), lx = t % (KBX + 2);
        const int qy = y0 + ly - 1, qx = x0 + lx - 1;
        float v = 0.f;
        if (qy >= 0 && qy < h && qx >= 0 && qx < w) {
            const float2 a = __ldg(reinterpret_cast<const float2 *>(raw + (size_t)(2 * qy) * W + 2 * qx));
            const float2 b = __ldg(reinterpret_cast<const float2 *>(raw + (size_t)(2 * qy + 1) * W + 2 * qx));
            // utils_image.py:353-357: c (float64) accumulates the four float32 GAT values in raster order
            double c = (double)gat_px(a.x, p.alpha, c0, toa);
            c += (double)gat_px(a.y, p.alpha, c0, toa);
            c += (double)gat_px(b.x, p.alpha, c0, toa);
            c += (double)gat_px(b.y, p.alpha, c0, toa);
            v = (float)(c * 0.25);
        }
        g[ly][lx] = v;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= w || y >= h) return;
    float T00 = 0.f, T01 = 0.f, T11 = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int gy_ = y - 1 + i, gx_ = x - 1 + j;   // point of the (h-1)x(w-1) gradient grid
            if (gy_ < 0 || gy_ >= h - 1 || gx_ < 0 || gx_ >= w - 1) continue;
            const int ly = threadIdx.y + i, lx = threadIdx.x + j;   // g[ly][lx] == grey(gy_, gx_)
            const float g00 = g[ly][lx], g01 = g[ly][lx + 1], g10 = g[ly + 1][lx], g11 = g[ly + 1][lx + 1];
            // kernels.py:97-112: horizontal [-.5,.5]/[.5,.5] then vertical [.5,.5]/[-.5,.5] (x0.5 is exact)
            const float dx = 0.5f * (0.5f * (g01 - g00)) + 0.5f * (0.5f * (g11 - g10));
            const float dy = 0.5f * (0.5f * (g10 + g11)) - 0.5f * (0.5f * (g00 + g01));
            T00 = __fmaf_rn(dx, dx, T00);
            T01 = __fmaf_rn(dx, dy, T01);
            T11 = __fmaf_rn(dy, dy, T11);
        }
    // eigenvalues (linalg.py:86-130): roots of l^2 + b l + c, discriminant in float64
    const float bq = -(T00 + T11);
    const float cq = __fmaf_rn(T00, T11, -__fmul_rn(T01, T01));
    double delta = (double)__fmul_rn(bq, bq) - 4.0 * (double)cq;
    delta = (0.0 > delta) ? 0.0 : delta;
    const double sq = sqrt_fast(delta);
    const double r1 = (-(double)bq + sq) * 0.5, r2 = (-(double)bq - sq) * 0.5;
    float l1, l2;
    if (fabs(r1) >= fabs(r2)) {
        l1 = (float)r1, l2 = (float)r2;
    } else {
        l1 = (float)r2, l2 = (float)r1;
    }
    // eigenvectors (linalg.py:132-179)
    float e1x, e1y, e2x, e2y;
    if (T01 == 0.f && T00 == T11) {
        e1x = 1.f, e1y = 0.f, e2x = 0.f, e2y = 1.f;
    } else {
        e1x = __fadd_rn(T00, T01) - l2;
        e1y = __fadd_rn(T01, T11) - l2;
        if (e1x == 0.f) {
            e1y = 1.f, e2x = 1.f, e2y = 0.f;
        } else if (e1y == 0.f) {
            e1x = 1.f, e2x = 0.f, e2y = 1.f;
        } else {
            const float nrm = __fsqrt_rn(__fmaf_rn(e1x, e1x, __fmul_rn(e1y, e1y)));
            e1x = __fdiv_rn(e1x, nrm);
            e1y = __fdiv_rn(e1y, nrm);
            e2y = fabsf(e1x);
            e2x = -e1y * copysignf(1.f, e1x);
        }
    }
    // compute_k (kernels.py:194-243)
    const double A = 1.0 + (double)__fsqrt_rn(__fdiv_rn(l1 - l2, l1 + l2));
    double D = 1.0 - (double)__fsqrt_rn(l1) * p.inv_D_tr + p.D_th;
    D = (D > 0.0) ? D : 0.0;      // Numba max(0, x): NaN -> 0
    D = (D < 1.0) ? D : 1.0;
    double k1, k2;
    if (p.law == 0) {
        if (A > 1.95) {
            k1 = p.inv_k_shrink, k2 = p.k_stretch;
        } else {
            k1 = 1.0, k2 = 1.0;
        }
    } else {
        k1 = 1.0 + A * 0.5 * (p.inv_k_shrink - 1.0);
        k2 = 1.0 + A * 0.5 * (p.k_stretch - 1.0);
    }
    const float kk1 = (float)(p.k_detail * ((1.0 - D) * k1 + D * p.k_denoise));
    const float kk2 = (float)(p.k_detail * ((1.0 - D) * k2 + D * p.k_denoise));
    const float k1s = kk1 * kk1, k2s = kk2 * kk2;
    const float cxx = k1s * e1x * e1x + k2s * e2x * e2x;
    const float cxy = k1s * e1x * e1y + k2s * e2x * e2y;
    const float cyy = k1s * e1y * e1y + k2s * e2y * e2y;
    reinterpret_cast<float4 *>(covs)[(size_t)y * w + x] = make_float4(cxx, cxy, cxy, cyy);
}

// stand-alone stages of the reference API (utils_image.py:117-170, 346-357); the fused kernel above does not use them
__global__ void gat_kernel(const float *__restrict__ img, size_t n, double alpha, double beta, float *__restrict__ out) {
    const double c0 = 3.0 / 8.0 * alpha * alpha + beta, toa = 2.0 / alpha;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = gat_px(img[i], alpha, c0, toa);
}

__global__ void decimate_kernel(const float *__restrict__ img, int W, int h, int w, float *__restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    double c = (double)img[(size_t)(2 * y) * W + 2 * x];
    c += (double)img[(size_t)(2 * y) * W + 2 * x + 1];
    c += (double)img[(size_t)(2 * y + 1) * W + 2 * x];
    c += (double)img[(size_t)(2 * y + 1) * W + 2 * x + 1];
    out[(size_t)y * w + x] = (float)(c / 4.0);
}

}  // namespace hhsr

using namespace hhsr;

extern "C" int hhsr_gat(const float *img, size_t n, double alpha, double beta, float *out, hhsr_stream_t stream) {
    HHSR_REQUIRE(img && out && n > 0, "null pointer or empty array");
    HHSR_REQUIRE(alpha > 0.0, "alpha should be positive (utils_image.py:139)");
    size_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    gat_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(img, n, alpha, beta, out);
    return launch_status("gat");
}

extern "C" int hhsr_decimate_to_grey(const float *img, int H, int W, float *out, hhsr_stream_t stream) {
    HHSR_REQUIRE(img && out, "null pointer");
    HHSR_REQUIRE(H >= 2 && W >= 2, "frame must be at least 2x2");
    const int h = H / 2, w = W / 2;
    dim3 block(32, 8), grid(ceil_div(w, 32), ceil_div(h, 8));
    decimate_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(img, W, h, w, out);
    return launch_status("decimate_to_grey");
}

extern "C" int hhsr_estimate_kernels(const float *raw, int H, int W, double alpha, double beta, double k_detail,
                                     double k_denoise, double D_th, double D_tr, double k_stretch, double k_shrink,
                                     int law, float *covs, hhsr_stream_t stream) {
    HHSR_REQUIRE(raw && covs, "null pointer");
    HHSR_REQUIRE(H >= 2 && W >= 2 && W % 2 == 0, "frame must be at least 2x2 with an even width");
    HHSR_REQUIRE(alpha > 0.0, "alpha should be positive (utils_image.py:139)");
    HHSR_REQUIRE(law == 0 || law == 1, "selection law must be 0 (hard_threshold) or 1 (linear)");
    HHSR_REQUIRE(((uintptr_t)raw % 8 == 0) && ((uintptr_t)covs % 16 == 0), "raw must be 8-byte, covs 16-byte aligned");
    const int h = H / 2, w = W / 2;
    HHSR_REQUIRE(D_tr != 0.0 && k_shrink != 0.0, "D_tr and k_shrink must be non-zero");
    KernelParams p{alpha, beta, k_detail, k_denoise, D_th, D_tr, k_stretch, k_shrink, 1.0 / D_tr, 1.0 / k_shrink, law};
    dim3 block(KBX, KBY), grid(ceil_div(w, KBX), ceil_div(h, KBY));
    estimate_kernels_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(raw, H, W, h, w, p, covs);
    return launch_status("estimate_kernels");
}
