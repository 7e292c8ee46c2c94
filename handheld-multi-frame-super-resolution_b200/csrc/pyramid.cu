// Grey-image band mask and Gaussian pyramid for sm_100a.
//
// Replaces handheld_super_resolution/utils_image.py:82-100 (the four masked fills + two fftshift copies of
// compute_grey_images; the FFTs themselves stay cuFFT), alignment.py:26-37 (circular padding) and
// utils_image.py:360-391 (cuda_downsample: two full-resolution F.conv2d followed by a strided slice).  The
// downsample kernel computes only the kept outputs: y pass then x pass inside one CTA, input tile staged in
// shared memory, so HBM traffic is one read of the level plus one write of the next (the reference reads and
// writes the full-resolution image three times).
#include "common.cuh"

namespace hhsr {

// keep flag of the reference's mask on the UNSHIFTED frequency index k of an axis of length n:
// shifted index s = (k + n/2) mod n is kept iff n/4 <= s < n - ceil(n/4)   (utils_image.py:92-95)
__device__ __forceinline__ bool band_keep(int k, int n) {   // 0 <= k <= n
    int s = k + n / 2;
    s -= (s >= n) ? n : 0;
    s -= (s >= n) ? n : 0;   // k == n (the partner of k == 0)
    return s >= n / 4 && s < n - (n + 3) / 4;
}

// sy/sx: strides of spec in complex elements.  torch.fft.rfft2 on CUDA returns a column-major [H][W/2+1] tensor
// (strides (1, H)); xfast selects which index runs along threadIdx.x so that accesses stay coalesced either way.
__global__ void grey_band_mask_kernel(float2 *__restrict__ spec, int H, int W, int Wc, long long sy, long long sx,
                                      int xfast, float scale) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y;
    const int kx = xfast ? f : s, ky = xfast ? s : f;
    if (kx >= Wc || ky >= H) return;
    const float a = (band_keep(ky, H) && band_keep(kx, W)) ? 0.5f : 0.f;
    const float b = (band_keep(H - ky, H) && band_keep(W - kx, W)) ? 0.5f : 0.f;   // k = n wraps to 0 inside
    const float m = (a + b) * scale;   // Re(ifft2(M F)) == ifft2(0.5 (M(k) + M(-k)) F) for a real image
    float2 *p = spec + (long long)ky * sy + (long long)kx * sx;
    if (m == 0.f)
        *p = make_float2(0.f, 0.f);
    else if (m != 1.f) {
        float2 v = *p;
        v.x *= m, v.y *= m;
        *p = v;
    }
}

// Same mask for the layout torch.fft.rfft2 actually returns on CUDA (ky contiguous: sy == 1): one CTA row per kx, so the
// two kx keep-flags are CTA-uniform (about half of the columns are zero as a whole), two ky per thread and one 16-byte
// store when both vanish.
__global__ void __launch_bounds__(128) grey_band_mask_cols_kernel(float2 *__restrict__ spec, int H, int W, long long sx,
                                                                  float scale) {
    const int kx = blockIdx.y, ky0 = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (ky0 >= H) return;
    const bool kxa = band_keep(kx, W), kxb = band_keep(W - kx, W);
    float2 *p = spec + (long long)kx * sx + ky0;
    float m[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int ky = ky0 + j;
        const float a = (kxa && ky < H && band_keep(ky, H)) ? 0.5f : 0.f;
        const float b = (kxb && ky < H && band_keep(H - ky, H)) ? 0.5f : 0.f;
        m[j] = (a + b) * scale;
    }
    const bool pair = ky0 + 1 < H;
    if (pair && m[0] == 0.f && m[1] == 0.f && (((uintptr_t)p) & 15) == 0) {
        *reinterpret_cast<float4 *>(p) = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        if (j == 1 && !pair) break;
        if (m[j] == 0.f)
            p[j] = make_float2(0.f, 0.f);
        else if (m[j] != 1.f) {
            float2 v = p[j];
            v.x *= m[j], v.y *= m[j];
            p[j] = v;
        }
    }
}

__global__ void pad_circular_kernel(const float *__restrict__ src, int h, int w, float *__restrict__ dst, int hp, int wp) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= wp || y >= hp) return;
    dst[(size_t)y * wp + x] = __ldg(src + (size_t)(y % h) * w + (x % w));
}

constexpr int kMaxTaps = 33;
struct Taps {
    float g[kMaxTaps];
};
constexpr int DBX = 32, DBY = 8;

// smem: in[(DBY*F + 2R)][(DBX*F + 2R)] then tmp[DBY][(DBX*F + 2R)].  F = 0 selects the run-time (f, R) version;
// the compile-time versions (F = 2, R = 4 and F = 4, R = 8: the factors of the default pyramid) unroll the taps and
// keep all index arithmetic in shifts.
template <int F, int RR>
__global__ void __launch_bounds__(DBX *DBY) gauss_downsample_kernel(const float *__restrict__ src, int h, int w, int f_rt, int R_rt,
                                                                    Taps taps, float *__restrict__ dst, int h2, int w2) {
    extern __shared__ float sm[];
    const int f = F ? F : f_rt, R = F ? RR : R_rt;
    const int K = 2 * R + 1;
    const int tin_w = DBX * f + 2 * R, tin_h = DBY * f + 2 * R;
    float *tin = sm, *tmp = sm + tin_w * tin_h;
    const int ox0 = blockIdx.x * DBX, oy0 = blockIdx.y * DBY;
    const int ix0 = ox0 * f, iy0 = oy0 * f;
    const int tid = threadIdx.y * DBX + threadIdx.x;
    for (int ly = threadIdx.y; ly < tin_h; ly += DBY) {           // one warp per input row: coalesced, no div/mod
        const int gy = iy0 + ly;
        for (int lx = threadIdx.x; lx < tin_w; lx += DBX) {
            const int gx = ix0 + lx;
            tin[ly * tin_w + lx] = (gy < h && gx < w) ? __ldg(src + (size_t)gy * w + gx) : 0.f;
        }
    }
    __syncthreads();
    // y pass (first F.conv2d, utils_image.py:383): only the DBY kept rows; thread (r = threadIdx.y) sweeps columns
    {
        const int r = threadIdx.y;
        for (int c = threadIdx.x; c < tin_w; c += DBX) {
            float acc = 0.f;
            const float *col = tin + (r * f) * tin_w + c;
            if (F) {
#pragma unroll
                for (int a = 0; a < 2 * RR + 1; ++a) acc = fmaf(taps.g[a], col[a * tin_w], acc);
            } else {
                for (int a = 0; a < K; ++a) acc = fmaf(taps.g[a], col[a * tin_w], acc);
            }
            tmp[r * tin_w + c] = acc;
        }
    }
    __syncthreads();
    (void)tid;
    const int ox = ox0 + threadIdx.x, oy = oy0 + threadIdx.y;
    if (ox >= w2 || oy >= h2) return;
    float acc = 0.f;   // x pass (second F.conv2d, :384)
    const float *rowp = tmp + threadIdx.y * tin_w + threadIdx.x * f;
    if (F) {
#pragma unroll
        for (int b = 0; b < 2 * RR + 1; ++b) acc = fmaf(taps.g[b], rowp[b], acc);
    } else {
        for (int b = 0; b < K; ++b) acc = fmaf(taps.g[b], rowp[b], acc);
    }
    dst[(size_t)oy * w2 + ox] = acc;
}


// The two factors of the default pyramid (2 with 9 taps, 4 with 17 taps) as a column-streaming kernel: a thread owns one
// INPUT column of the CTA's strip and walks down it with the 2R + 1 rows of the y pass in registers (every input sample
// is loaded exactly once, coalesced across the warp, all loads of a strip independent of each other); the RY rows of
// y-pass results go to shared memory once, then the x pass reads its 2R + 1 neighbours as aligned 8- / 16-byte vectors
// (conflict-free).  Same arithmetic, same order as the generic kernel: acc = fma(g[a], v[a], acc), a ascending, y pass
// then x pass — the outputs are bit-identical.
template <int F, int RR, int RY>
__global__ void __launch_bounds__(256) gauss_downsample_stream_kernel(const float *__restrict__ src, int h, int w, Taps taps,
                                                                      float *__restrict__ dst, int h2, int w2) {
    constexpr int K = 2 * RR + 1, NT = 256, OX = (NT - 2 * RR) / F;
    __shared__ __align__(16) float tmp[RY][NT];
    const int tid = threadIdx.x;
    const int ox0 = blockIdx.x * OX, oy0 = blockIdx.y * RY;
    const int gx = ox0 * F + tid, iy0 = oy0 * F;
    const float *p = src + (size_t)iy0 * w + gx;
    const bool colok = gx < w;
    float win[K];
#pragma unroll
    for (int a = 0; a < K - F; ++a) win[a + F] = (colok && iy0 + a < h) ? __ldg(p + (size_t)a * w) : 0.f;
#pragma unroll
    for (int r = 0; r < RY; ++r) {
#pragma unroll
        for (int a = 0; a < K - F; ++a) win[a] = win[a + F];
#pragma unroll
        for (int j = 0; j < F; ++j) {
            const int row = F * r + K - F + j;
            win[K - F + j] = (colok && iy0 + row < h) ? __ldg(p + (size_t)row * w) : 0.f;
        }
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < K; ++a) acc = fmaf(taps.g[a], win[a], acc);
        tmp[r][tid] = acc;
    }
    __syncthreads();
    for (int idx = tid; idx < RY * OX; idx += NT) {
        const int r = idx / OX, oxl = idx - r * OX;
        const int ox = ox0 + oxl, oy = oy0 + r;
        if (ox >= w2 || oy >= h2) continue;
        const float *rowp = &tmp[r][F * oxl];
        float v[K];
        if (F == 2) {
#pragma unroll
            for (int j = 0; j < K / 2; ++j) {
                const float2 t = *reinterpret_cast<const float2 *>(rowp + 2 * j);
                v[2 * j] = t.x, v[2 * j + 1] = t.y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < K / 4; ++j) {
                const float4 t = *reinterpret_cast<const float4 *>(rowp + 4 * j);
                v[4 * j] = t.x, v[4 * j + 1] = t.y, v[4 * j + 2] = t.z, v[4 * j + 3] = t.w;
            }
        }
        v[K - 1] = rowp[K - 1];
        float acc = 0.f;
#pragma unroll
        for (int b = 0; b < K; ++b) acc = fmaf(taps.g[b], v[b], acc);
        dst[(size_t)oy * w2 + ox] = acc;
    }
}

template <int F, int RR, int RY>
static void launch_stream(const float *src, int h, int w, const Taps &t, float *dst, int h2, int w2, cudaStream_t st) {
    constexpr int OX = (256 - 2 * RR) / F;
    dim3 grid(ceil_div(w2, OX), ceil_div(h2, RY));
    gauss_downsample_stream_kernel<F, RR, RY><<<grid, 256, 0, st>>>(src, h, w, t, dst, h2, w2);
}

}  // namespace hhsr

using namespace hhsr;

extern "C" int hhsr_grey_band_mask(float *spec, int H, int W, long long stride_y, long long stride_x, float scale,
                                   hhsr_stream_t stream) {
    HHSR_REQUIRE(spec, "null pointer");
    HHSR_REQUIRE(H > 0 && W > 0, "non-positive size");
    HHSR_REQUIRE((uintptr_t)spec % 8 == 0, "spectrum must be 8-byte aligned");
    HHSR_REQUIRE(stride_y > 0 && stride_x > 0, "strides must be positive");
    HHSR_REQUIRE(scale > 0.0f, "scale must be positive");
    const int Wc = W / 2 + 1;
    if (stride_y == 1) {
        dim3 block(128), grid(ceil_div(ceil_div(H, 2), 128), Wc);
        grey_band_mask_cols_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(reinterpret_cast<float2 *>(spec), H, W, stride_x, scale);
        return launch_status("grey_band_mask");
    }
    const int xfast = stride_x <= stride_y;
    const int nfast = xfast ? Wc : H, nslow = xfast ? H : Wc;
    dim3 block(128), grid(ceil_div(nfast, 128), nslow);
    grey_band_mask_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(reinterpret_cast<float2 *>(spec), H, W, Wc, stride_y,
                                                                   stride_x, xfast, scale);
    return launch_status("grey_band_mask");
}

extern "C" int hhsr_pad_circular(const float *src, int h, int w, float *dst, int hp, int wp, hhsr_stream_t stream) {
    HHSR_REQUIRE(src && dst, "null pointer");
    HHSR_REQUIRE(h > 0 && w > 0 && hp >= h && wp >= w, "padded size must be >= source size");
    dim3 block(32, 8), grid(ceil_div(wp, 32), ceil_div(hp, 8));
    pad_circular_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, h, w, dst, hp, wp);
    return launch_status("pad_circular");
}

extern "C" int hhsr_gauss_downsample(const float *src, int h, int w, int factor, const float *taps_host, int radius,
                                     float *dst, int h2, int w2, hhsr_stream_t stream) {
    HHSR_REQUIRE(src && dst && taps_host, "null pointer");
    HHSR_REQUIRE(factor >= 1 && radius >= 0 && 2 * radius + 1 <= kMaxTaps, "factor/radius out of range");
    HHSR_REQUIRE(h > 2 * radius && w > 2 * radius, "level smaller than the filter");
    HHSR_REQUIRE(h2 == (h - 2 * radius) / factor && w2 == (w - 2 * radius) / factor && h2 > 0 && w2 > 0,
                 "output shape must be ((h-2r)/f, (w-2r)/f)");
    Taps t;
    for (int i = 0; i < 2 * radius + 1; ++i) t.g[i] = taps_host[i];
    const int tin_w = DBX * factor + 2 * radius, tin_h = DBY * factor + 2 * radius;
    const size_t smem = (size_t)(tin_w * tin_h + DBY * tin_w) * sizeof(float);
    if (smem > 200 * 1024) return unsupported("downsampling factor too large for one CTA tile");
    dim3 block(DBX, DBY), grid(ceil_div(w2, DBX), ceil_div(h2, DBY));
    cudaStream_t st = (cudaStream_t)stream;
    if (factor == 2 && radius == 4) {
        // strips of 32 output rows; short strips when the level is too small to fill the GPU otherwise
        if ((long long)ceil_div(w2, 124) * ceil_div(h2, 32) >= 296) launch_stream<2, 4, 32>(src, h, w, t, dst, h2, w2, st);
        else launch_stream<2, 4, 8>(src, h, w, t, dst, h2, w2, st);
    } else if (factor == 4 && radius == 8) {
        if ((long long)ceil_div(w2, 60) * ceil_div(h2, 16) >= 148) launch_stream<4, 8, 16>(src, h, w, t, dst, h2, w2, st);
        else launch_stream<4, 8, 4>(src, h, w, t, dst, h2, w2, st);
    } else {
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(gauss_downsample_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        gauss_downsample_kernel<0, 0><<<grid, block, smem, st>>>(src, h, w, factor, radius, t, dst, h2, w2);
    }
    return launch_status("gauss_downsample");
}
