// Noise curves sigma(b), d(b) by Monte-Carlo on the device, sm_100a.
//
// Replaces handheld_super_resolution/fast_monte_carlo.py:31-66 (unitary_MC) and :68-101 (regular_MC: a multiprocessing
// pool over brightness levels, ~10 s on the host, unseeded): for every requested brightness level b, n_patches pairs of
// 3x3 patches  clip(b + sqrt(alpha b + beta) N(0,1), 0, 1)  are drawn and
//     diff_mean = mean |mean(patch1) - mean(patch2)|,   std_mean = 0.5 mean(std(patch1) + std(patch2))
// (population standard deviation, float64 like NumPy) are returned.  One CTA per level; the normals come from a
// counter-based Philox4x32-10 generator keyed by (seed, level) and indexed by (patch, draw), so the result depends only
// on (seed, n_patches, level list) — process() becomes reproducible, which the reference is not.  The interpolation of
// the two curves between the clipped ends (run_fast_MC, :157-230) stays on the host (1001 values).
#include "common.cuh"

namespace hhsr {

__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1,
                                              unsigned (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0, c1 = lo1, c2 = hi0 ^ c3 ^ k1, c3 = lo0;
        k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
    }
    out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}

// two uint32 -> two independent N(0,1) (Box-Muller in float64; u1 in (0, 1])
__device__ __forceinline__ void box_muller(unsigned a, unsigned b, double &n0, double &n1) {
    const double u1 = ((double)a + 1.0) * (1.0 / 4294967296.0), u2 = (double)b * (1.0 / 4294967296.0);
    const double r = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    n0 = r * c, n1 = r * s;
}

constexpr int kMcThreads = 256;

__global__ void __launch_bounds__(kMcThreads) noise_mc_kernel(const double *__restrict__ brightness, double alpha, double beta,
                                                              int n_patches, unsigned long long seed, double *__restrict__ diff_mean,
                                                              double *__restrict__ std_mean) {
    const int level = blockIdx.x;
    const double b = brightness[level];
    const double sd = sqrt(b * alpha + beta);
    const unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32) ^ (0x9E3779B9u * (unsigned)(level + 1));
    double acc_diff = 0.0, acc_std = 0.0;
    for (int p = threadIdx.x; p < n_patches; p += kMcThreads) {
        double mean[2], sdev[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {            // the two patches of a pair
            double v[12];
#pragma unroll
            for (int q = 0; q < 3; ++q) {        // 3 Philox calls -> 12 normals, 9 used
                unsigned r[4];
                philox4x32_10((unsigned)p, (unsigned)(h * 3 + q), (unsigned)level, 0x48485352u, k0, k1, r);
                box_muller(r[0], r[1], v[4 * q], v[4 * q + 1]);
                box_muller(r[2], r[3], v[4 * q + 2], v[4 * q + 3]);
            }
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                v[i] = fmin(fmax(b + sd * v[i], 0.0), 1.0);
                s += v[i];
            }
            const double m = s / 9.0;
            double ss = 0.0;
#pragma unroll
            for (int i = 0; i < 9; ++i) ss += (v[i] - m) * (v[i] - m);
            mean[h] = m, sdev[h] = sqrt(ss / 9.0);      // np.std: population standard deviation
        }
        acc_diff += fabs(mean[0] - mean[1]);
        acc_std += sdev[0] + sdev[1];
    }
    __shared__ double red[2][kMcThreads];
    red[0][threadIdx.x] = acc_diff, red[1][threadIdx.x] = acc_std;
    __syncthreads();
    for (int o = kMcThreads / 2; o > 0; o >>= 1) {      // fixed tree: deterministic
        if (threadIdx.x < o) red[0][threadIdx.x] += red[0][threadIdx.x + o], red[1][threadIdx.x] += red[1][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        diff_mean[level] = red[0][0] / (double)n_patches;
        std_mean[level] = 0.5 * red[1][0] / (double)n_patches;
    }
}

}  // namespace hhsr

using namespace hhsr;

extern "C" int hhsr_noise_mc(const double *brightness, int n_levels, double alpha, double beta, int n_patches,
                             unsigned long long seed, double *diff_mean, double *std_mean, hhsr_stream_t stream) {
    HHSR_REQUIRE(brightness && diff_mean && std_mean, "null pointer");
    HHSR_REQUIRE(n_levels > 0 && n_patches > 0, "non-positive size");
    HHSR_REQUIRE(alpha >= 0.0 && beta >= 0.0, "alpha and beta must be non-negative");
    noise_mc_kernel<<<n_levels, kMcThreads, 0, (cudaStream_t)stream>>>(brightness, alpha, beta, n_patches, seed, diff_mean, std_mean);
    return launch_status("noise_mc");
}
