// Row-sharded merge across GPUs (a B200 addition): the exchange step, sm_100a.
//
// Frames are sharded over ranks for alignment / robustness / kernel estimation.  Instead of summing full-size partial
// accumulators (24 B per HR pixel and rank over NVLink), every rank then merges ALL frames into ITS OWN slice of output
// rows: it only needs, from each frame's owner, the band of LR rows its slice can touch — raw, robustness and covariance
// rows (12 B per LR pixel of the band) and the tile flow.  These two kernels form that one exchange point:
//   band_extents_kernel  per frame: read the owner's flow over NVLink (peer-mapped pointer), keep a local copy, and derive
//                        from its vertical range the LR row band [lo, hi) the slice needs — on the device, no host sync;
//   band_copy_kernel     pull exactly those rows of the three planes from the owners' HBM into local full-size planes
//                        (16-byte loads over NVLink, grid-stride), so that the merge kernels run on local memory.
#include "common.cuh"

namespace hhsr {

constexpr int kMaxGather = 24;
struct GatherList {
    const float *raw[kMaxGather], *r[kMaxGather], *covs[kMaxGather], *flow[kMaxGather];   // sources (peer-mapped or local)
    float *raw_l[kMaxGather], *r_l[kMaxGather], *covs_l[kMaxGather], *flow_l[kMaxGather];  // local destinations
    int n;
};

// ext[f] = (lo, hi, covs_lo, covs_hi)
constexpr int kExtThreads = 1024;
__global__ void __launch_bounds__(kExtThreads) band_extents_kernel(const __grid_constant__ GatherList l, int ny, int nx, int ts, int H, int lr0,
                                                                   int lr1, int4 *__restrict__ ext) {
    const int f = blockIdx.x;
    const float2 *src = reinterpret_cast<const float2 *>(l.flow[f]);
    float2 *dst = reinterpret_cast<float2 *>(l.flow_l[f]);
    const int py0 = lr0 / ts, py1 = min((lr1 - 1) / ts, ny - 1);
    float mn = INFINITY, mx = -INFINITY;
    const int n = ny * nx;
    // four independent (remote) loads in flight per thread: the NVLink round trip is paid ~n / 4096 times, not n / 256
    for (int i0 = threadIdx.x; i0 < n; i0 += 4 * kExtThreads) {
        float2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * kExtThreads;
            v[u] = (i < n) ? src[i] : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * kExtThreads;
            if (i >= n) continue;
            if (dst != src) dst[i] = v[u];
            const int py = i / nx;
            if (py >= py0 && py <= py1) mn = fminf(mn, v[u].y), mx = fmaxf(mx, v[u].y);
        }
    }
    __shared__ float smn[kExtThreads / 32], smx[kExtThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)), mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) smn[threadIdx.x >> 5] = mn, smx[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < kExtThreads / 32; ++k) mn = fminf(mn, smn[k]), mx = fmaxf(mx, smx[k]);
        // centre rows floor(lr + flow_y) for lr in [lr0, lr1): 3x3 taps around them, one more row of slack; the rows
        // [lr0, lr1) themselves are read for the robustness
        const float flo = fmaxf((float)lr0 + mn - 3.0f, 0.0f), fhi = fminf((float)lr1 + mx + 3.0f, (float)H);
        int lo = (mn <= mx) ? (int)floorf(flo) : lr0, hi = (mn <= mx) ? (int)ceilf(fhi) : lr1;
        lo = max(min(lo, lr0), 0), hi = min(max(hi, lr1), H);
        const int ch = H / 2;
        ext[f] = make_int4(lo, hi, max(lo / 2 - 1, 0), min(hi / 2 + 2, ch));
    }
}

__global__ void __launch_bounds__(256) band_copy_kernel(const __grid_constant__ GatherList l, int W, const int4 *__restrict__ ext) {
    const int f = blockIdx.z, plane = blockIdx.y;
    const float *src = plane == 0 ? l.raw[f] : (plane == 1 ? l.r[f] : l.covs[f]);
    float *dst = plane == 0 ? l.raw_l[f] : (plane == 1 ? l.r_l[f] : l.covs_l[f]);
    if (src == nullptr || src == dst) return;                 // iso kernel (no covariances) / a frame this rank owns
    const int4 e = ext[f];
    const size_t row = (plane == 2) ? (size_t)(W / 2) * 4 : (size_t)W;     // floats per row of the plane
    const size_t begin = (size_t)(plane == 2 ? e.z : e.x) * row, end = (size_t)(plane == 2 ? e.w : e.y) * row;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    if ((row & 3) == 0) {
        // 4 x 16 bytes in flight per thread (the loads of one iteration are independent): NVLink latency is several
        // microseconds under load and one outstanding load per thread leaves the links half idle
        const float4 *s4 = reinterpret_cast<const float4 *>(src);
        float4 *d4 = reinterpret_cast<float4 *>(dst);
        const size_t e4 = end / 4;
        for (size_t i0 = begin / 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < e4; i0 += 4 * stride) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i0 + u * stride < e4) v[u] = s4[i0 + u * stride];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i0 + u * stride < e4) d4[i0 + u * stride] = v[u];
        }
    } else {
        for (size_t i = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += stride) dst[i] = src[i];
    }
}

}  // namespace hhsr

using namespace hhsr;

extern "C" int hhsr_gather_bands(const float *const *raws, const float *const *rs, const float *const *covs,
                                 const float *const *flows, float *const *raws_local, float *const *rs_local,
                                 float *const *covs_local, float *const *flows_local, int n_frames, int H, int W, int ny, int nx,
                                 int ts, int lr_begin, int lr_end, int *extents, hhsr_stream_t stream) {
    HHSR_REQUIRE(raws && rs && flows && raws_local && rs_local && flows_local && extents, "null pointer");
    HHSR_REQUIRE(n_frames > 0 && n_frames <= kMaxGather, "1 to 24 frames per call");
    HHSR_REQUIRE(H > 1 && W > 1 && ts > 0 && ny * ts >= H && nx * ts >= W, "flow grid does not cover the frame");
    HHSR_REQUIRE(0 <= lr_begin && lr_begin < lr_end && lr_end <= H, "LR row range must satisfy 0 <= begin < end <= H");
    HHSR_REQUIRE((uintptr_t)extents % 16 == 0, "extents must be 16-byte aligned");
    GatherList l;
    l.n = n_frames;
    for (int f = 0; f < n_frames; ++f) {
        HHSR_REQUIRE(raws[f] && rs[f] && flows[f] && raws_local[f] && rs_local[f] && flows_local[f], "null frame pointer");
        const bool has_covs = covs && covs[f];
        HHSR_REQUIRE(!has_covs || (covs_local && covs_local[f]), "covs source without a local destination");
        HHSR_REQUIRE((uintptr_t)raws[f] % 16 == 0 && (uintptr_t)rs[f] % 16 == 0 && (uintptr_t)raws_local[f] % 16 == 0 &&
                         (uintptr_t)rs_local[f] % 16 == 0 && (uintptr_t)flows[f] % 8 == 0 && (uintptr_t)flows_local[f] % 8 == 0,
                     "planes must be 16-byte aligned, flows 8-byte aligned");
        l.raw[f] = raws[f], l.r[f] = rs[f], l.covs[f] = has_covs ? covs[f] : nullptr, l.flow[f] = flows[f];
        l.raw_l[f] = raws_local[f], l.r_l[f] = rs_local[f], l.covs_l[f] = has_covs ? covs_local[f] : nullptr, l.flow_l[f] = flows_local[f];
    }
    cudaStream_t st = (cudaStream_t)stream;
    band_extents_kernel<<<n_frames, kExtThreads, 0, st>>>(l, ny, nx, ts, H, lr_begin, lr_end, reinterpret_cast<int4 *>(extents));
    band_copy_kernel<<<dim3(148, 3, n_frames), 256, 0, st>>>(l, W, reinterpret_cast<const int4 *>(extents));
    return launch_status("gather_bands");
}
