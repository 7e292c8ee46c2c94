// Shared helpers for libhhsr.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdarg>
#include <cmath>
#include <cstdint>
#include "hhsr.h"

namespace hhsr {

void set_error(const char *fmt, ...);

inline int launch_status(const char *what) {
    cudaError_t e = cudaPeekAtLastError();  // never synchronises
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

inline int bad(const char *what) {
    set_error("bad argument: %s", what);
    return HHSR_E_BADARG;
}
inline int unsupported(const char *what) {
    set_error("unsupported: %s", what);
    return HHSR_E_UNSUPPORTED;
}

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

inline int pack_cfa(const int *cfa) {  // 2 bits per entry, index = (row&1)*2 + (col&1)
    return (cfa[0] & 3) | ((cfa[1] & 3) << 2) | ((cfa[2] & 3) << 4) | ((cfa[3] & 3) << 6);
}
__device__ __forceinline__ int cfa_channel(int packed, int i, int j) {
    return (packed >> ((((i & 1) << 1) | (j & 1)) << 1)) & 3;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace hhsr

#define HHSR_REQUIRE(cond, what) \
    do {                         \
        if (!(cond)) return hhsr::bad(what); \
    } while (0)
