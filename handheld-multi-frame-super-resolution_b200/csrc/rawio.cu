// RAW input side of the hot path for sm_100a: sensor counts (uint16, as decoded from a DNG) -> normalised float32.
//
// Replaces the host loop of handheld_super_resolution/utils_dng.py:146-160: the
// reference converts the whole burst to float32 on the CPU and uploads 4 B per pixel; here the burst crosses PCIe as
// 2 B per pixel and is normalised on the device, fused with the widening.  Arithmetic is the reference's float32
// chain operation by operation, with its Python-scalar operands rounded to float32 as NumPy does:
//     x = (float32(raw) - black[c]) / (white - black[c]);  x *= wb[c] / wb[1]        (c = CFA channel of the pixel)
// so the result is bit-identical to the NumPy code.  HBM traffic: 2 B in + 4 B out per pixel.
#include "common.cuh"

namespace hhsr {

struct RawNorm {
    float black[4], den[4], gain[4];   // per CFA position (row & 1) * 2 + (col & 1)
};

__device__ __forceinline__ float norm_px(unsigned v, float black, float den, float gain) {
    return __fmul_rn(__fdiv_rn(__fsub_rn((float)v, black), den), gain);
}

// one thread = 8 consecutive pixels of a row (one 16-byte load, two 16-byte stores); W % 8 == 0
__global__ void __launch_bounds__(256) normalize_u16_vec8_kernel(const uint4 *__restrict__ in, int H, int W8, RawNorm p,
                                                                 float4 *__restrict__ out) {
    const int x8 = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x8 >= W8 || y >= H) return;
    const int r = (y & 1) * 2;
    const float b0 = p.black[r], b1 = p.black[r + 1], d0 = p.den[r], d1 = p.den[r + 1], g0 = p.gain[r], g1 = p.gain[r + 1];
    const size_t i = (size_t)y * W8 + x8;
    const uint4 v = __ldg(in + i);
    float4 lo, hi;
    lo.x = norm_px(v.x & 0xffffu, b0, d0, g0), lo.y = norm_px(v.x >> 16, b1, d1, g1);
    lo.z = norm_px(v.y & 0xffffu, b0, d0, g0), lo.w = norm_px(v.y >> 16, b1, d1, g1);
    hi.x = norm_px(v.z & 0xffffu, b0, d0, g0), hi.y = norm_px(v.z >> 16, b1, d1, g1);
    hi.z = norm_px(v.w & 0xffffu, b0, d0, g0), hi.w = norm_px(v.w >> 16, b1, d1, g1);
    out[2 * i] = lo;
    out[2 * i + 1] = hi;
}

__global__ void normalize_u16_kernel(const unsigned short *__restrict__ in, int H, int W, RawNorm p, float *__restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const int c = (y & 1) * 2 + (x & 1);
    const size_t i = (size_t)y * W + x;
    out[i] = norm_px(in[i], p.black[c], p.den[c], p.gain[c]);
}

}  // namespace hhsr

using namespace hhsr;

extern "C" int hhsr_normalize_raw_u16(const unsigned short *raw, int H, int W, const float *black4, const float *den4,
                                      const float *gain4, float *out, hhsr_stream_t stream) {
    HHSR_REQUIRE(raw && black4 && den4 && gain4 && out, "null pointer");
    HHSR_REQUIRE(H > 0 && W > 0, "non-positive size");
    RawNorm p;
    for (int k = 0; k < 4; ++k) {
        HHSR_REQUIRE(den4[k] != 0.0f, "white level equals black level");
        p.black[k] = black4[k], p.den[k] = den4[k], p.gain[k] = gain4[k];
    }
    if (W % 8 == 0 && (uintptr_t)raw % 16 == 0 && (uintptr_t)out % 16 == 0) {
        dim3 block(64, 4), grid(ceil_div(W / 8, 64), ceil_div(H, 4));
        normalize_u16_vec8_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4 *>(raw), H, W / 8, p,
                                                                           reinterpret_cast<float4 *>(out));
    } else {
        dim3 block(32, 8), grid(ceil_div(W, 32), ceil_div(H, 8));
        normalize_u16_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(raw, H, W, p, out);
    }
    return launch_status("normalize_raw_u16");
}
