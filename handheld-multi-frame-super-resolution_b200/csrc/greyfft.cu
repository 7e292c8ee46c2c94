// Grey image of Alg. 3 (ComputeGrayscaleImage) as three shared-memory FFT passes for sm_100a.
//
// Replaces handheld_super_resolution/utils_image.py:82-100 as a whole: fft2, fftshift, the four masked fills, ifftshift,
// ifft2 and .real.  The mask is an ideal half-band low-pass, so only W/4 + 1 of the W/2 + 1 half-spectrum columns
// survive it; everything that would be multiplied by zero is never computed or stored:
//
//   1. rows forward   one CTA per PAIR of image rows: z = row_a + i row_b, one W-point complex transform in shared
//                     memory, the two half spectra are separated on the way out and only the KX kept columns are written
//                     (48 MB read, 24 MB written at 12 MP — a full-width R2C would write 48 MB);
//   2. columns        one CTA per tile of CW kept columns, whole columns in shared memory: forward H-point transform,
//                     band mask (with 1/(2HW) folded in), inverse transform, written back in place — the forward result
//                     stays in digit-reversed order because the inverse wants exactly that (no permutation pass, and the
//                     spectrum crosses HBM once instead of three times);
//   3. rows inverse   one CTA per pair of rows: the two Hermitian half spectra are recombined into one complex
//                     spectrum (zeros above the band), one inverse transform yields both real rows.
//
// HBM traffic 192 MB per 12 MP frame against ~600 MB for rfft2 + mask + irfft2 through cuFFT (7 kernels).  The
// arithmetic (few large register-resident radices, float32, table twiddles rounded from float64) is in fft_core.cuh, shared with the
// CPU emulation that tests/ checks against numpy.fft.
#include <atomic>

#include "common.cuh"
#include "fft_core.cuh"

namespace hhsr {

using fft::c32;
using fft::Plan;

template <class Kernel>
static void ensure_smem(Kernel kernel, std::atomic<unsigned long long> &done) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_relaxed) & bit) return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    done.fetch_or(bit, std::memory_order_relaxed);
}

#ifndef HHSR_GREY_ROW_CTAS
#define HHSR_GREY_ROW_CTAS 4
#endif
#ifndef HHSR_GREY_ROW_THREADS
#define HHSR_GREY_ROW_THREADS 128
#endif
constexpr int kRowThreads = HHSR_GREY_ROW_THREADS, kRowCtas = HHSR_GREY_ROW_CTAS, kColThreads = 512;
constexpr size_t kMaxSmem = 227 * 1024;

// tw[k] = e^{-2 pi i k / n} rounded from float64; ppos_of_k: physical (padded) shared-memory position of frequency k in the
// digit-reversed order of fft_core.cuh; k_of_pos: frequency held at logical position p
__global__ void fft_tables_kernel(Plan pl, c32 *tw, int *ppos_of_k, int *k_of_pos) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= pl.n) return;
    double s, c;
    sincospi(2.0 * (double)k / (double)pl.n, &s, &c);
    tw[k] = c32{(float)c, (float)-s};
    const int p = fft::digit_reversed(pl, k);
    ppos_of_k[k] = pl.pad ? fft::phys<true>(p) : p;
    k_of_pos[p] = k;
}

template <bool PAD>
__global__ void __launch_bounds__(kRowThreads, kRowCtas) grey_rows_forward_kernel(const float *__restrict__ img, int W, Plan pl,
                                                                        const c32 *__restrict__ tw,
                                                                        const int *__restrict__ ppos_of_k, c32 *__restrict__ spec,
                                                                        long long pitch, int KX, int KXp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c32 *s = reinterpret_cast<c32 *>(smem_raw);
    const int tid = threadIdx.x, nt = blockDim.x;
    const size_t r = 2 * (size_t)blockIdx.x;
    fft::rows_load_pair<PAD>(s, img + r * W, img + (r + 1) * W, W, tid, nt);
    __syncthreads();
    for (int i = 0; i < pl.count; ++i) {
        fft::run_stage<false, PAD>(s, tw, pl, i, tid, nt);
        __syncthreads();
    }
    fft::rows_store_half_spectra(s, ppos_of_k, spec + r * pitch, spec + (r + 1) * pitch, W, KX, KXp, tid, nt);
}

template <bool PAD>
__global__ void __launch_bounds__(kColThreads) grey_cols_kernel(c32 *__restrict__ spec, long long pitch, int H, int Hp, int W,
                                                                int CW, Plan pl, const c32 *__restrict__ tw,
                                                                const int *__restrict__ k_of_pos, float scale) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c32 *s = reinterpret_cast<c32 *>(smem_raw);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int c0 = blockIdx.x * CW;
    fft::cols_load_tile<PAD>(s, spec, pitch, H, Hp, c0, CW, tid, nt);
    __syncthreads();
    // a contiguous group of nt / CW threads per column
    const int per = nt / CW, c = tid / per, ctid = tid - c * per;
    const bool active = c < CW;
    c32 *col = s + c * Hp;
    for (int i = 0; i < pl.count; ++i) {
        if (active) fft::run_stage<false, PAD>(col, tw, pl, i, ctid, per);
        __syncthreads();
    }
    if (active) fft::cols_mask_column<PAD>(col, k_of_pos, H, W, c0 + c, scale, ctid, per);
    __syncthreads();
    for (int i = pl.count - 1; i >= 0; --i) {
        if (active) fft::run_stage<true, PAD>(col, tw, pl, i, ctid, per);
        __syncthreads();
    }
    fft::cols_store_tile<PAD>(s, spec, pitch, H, Hp, c0, CW, tid, nt);
}

template <bool PAD>
__global__ void __launch_bounds__(kRowThreads, kRowCtas) grey_rows_inverse_kernel(const c32 *__restrict__ spec, long long pitch, int W,
                                                                        Plan pl, const c32 *__restrict__ tw,
                                                                        const int *__restrict__ ppos_of_k, int KX,
                                                                        float *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c32 *s = reinterpret_cast<c32 *>(smem_raw);
    const int tid = threadIdx.x, nt = blockDim.x;
    const size_t r = 2 * (size_t)blockIdx.x;
    fft::rows_zero(s, fft::phys_len(W, PAD), tid, nt);
    __syncthreads();
    fft::rows_scatter_half_spectra(s, ppos_of_k, spec + r * pitch, spec + (r + 1) * pitch, W, KX, tid, nt);
    __syncthreads();
    for (int i = pl.count - 1; i >= 0; --i) {
        fft::run_stage<true, PAD>(s, tw, pl, i, tid, nt);
        __syncthreads();
    }
    fft::rows_store_pair<PAD>(s, out + r * W, out + (r + 1) * W, W, tid, nt);
}

// ---- host side: sizes, layout of the plan and work buffers
struct GreyLayout {
    Plan pw, ph;
    int KX, KXp, CW, Hp;
    size_t off_twW, off_twH, off_posW, off_kposW, off_posH, off_kposH, plan_bytes, work_bytes, smem_rows, smem_cols;
};

static int grey_layout(int H, int W, GreyLayout &g) {
    if (H < 8 || W < 8 || (H & 1)) return unsupported("grey FFT: needs an even number of rows and at least 8 x 8 pixels");
    if (!fft::make_plan(W, g.pw) || !fft::make_plan(H, g.ph))
        return unsupported("grey FFT: image sizes must factor into primes up to 19");
    g.KX = fft::kept_columns(W);
    g.smem_rows = (size_t)fft::phys_len(W, g.pw.pad) * sizeof(c32);
    if (g.smem_rows > kMaxSmem) return unsupported("grey FFT: row too long for shared memory");
    // columns per CTA: whole columns in shared memory; the fewest waves over the SMs, then the smallest tile
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
        cudaGetLastError();
        sms = 148;
    }
    const int hlen = fft::phys_len(H, g.ph.pad);
    g.CW = 0;
    long long best = 0;
    for (int cw = 8; cw >= 1; --cw) {
        int hp = hlen;
        const int want = (16 / cw) % 16;     // Hp = 16/CW (mod 16): the transposing tile copies are conflict-free
        while (cw > 1 && hp % 16 != want) ++hp;
        if ((size_t)cw * hp * sizeof(c32) > kMaxSmem) continue;
        const long long cost = (long long)ceil_div(ceil_div(g.KX, cw), sms) * cw;
        if (!g.CW || cost < best) g.CW = cw, g.Hp = hp, best = cost;
    }
    if (!g.CW) return unsupported("grey FFT: column too long for shared memory");
    g.smem_cols = (size_t)g.CW * g.Hp * sizeof(c32);
    g.KXp = ceil_div(g.KX, g.CW) * g.CW;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    size_t o = 0;
    g.off_twW = o, o += up((size_t)W * sizeof(c32));
    g.off_twH = o, o += up((size_t)H * sizeof(c32));
    g.off_posW = o, o += up((size_t)W * sizeof(int));
    g.off_kposW = o, o += up((size_t)W * sizeof(int));
    g.off_posH = o, o += up((size_t)H * sizeof(int));
    g.off_kposH = o, o += up((size_t)H * sizeof(int));
    g.plan_bytes = o;
    g.work_bytes = (size_t)H * g.KXp * sizeof(c32);
    return 0;
}

static std::atomic<unsigned long long> g_attr[6];

template <bool PW, bool PH>
static void launch_grey(const GreyLayout &g, const float *img, int H, int W, const char *p, c32 *spec, float *out, cudaStream_t st) {
    ensure_smem(grey_rows_forward_kernel<PW>, g_attr[PW ? 1 : 0]);
    ensure_smem(grey_rows_inverse_kernel<PW>, g_attr[PW ? 3 : 2]);
    ensure_smem(grey_cols_kernel<PH>, g_attr[PH ? 5 : 4]);
    grey_rows_forward_kernel<PW><<<H / 2, kRowThreads, g.smem_rows, st>>>(img, W, g.pw, (const c32 *)(p + g.off_twW),
                                                                         (const int *)(p + g.off_posW), spec, g.KXp, g.KX, g.KXp);
    grey_cols_kernel<PH><<<g.KXp / g.CW, kColThreads, g.smem_cols, st>>>(spec, g.KXp, H, g.Hp, W, g.CW, g.ph, (const c32 *)(p + g.off_twH),
                                                                        (const int *)(p + g.off_kposH), 0.5f / ((float)H * (float)W));
    grey_rows_inverse_kernel<PW><<<H / 2, kRowThreads, g.smem_rows, st>>>(spec, g.KXp, W, g.pw, (const c32 *)(p + g.off_twW),
                                                                         (const int *)(p + g.off_posW), g.KX, out);
}

}  // namespace hhsr

using namespace hhsr;

extern "C" int hhsr_grey_fft_sizes(int H, int W, size_t *plan_bytes, size_t *work_bytes) {
    HHSR_REQUIRE(plan_bytes && work_bytes, "null pointer");
    GreyLayout g;
    if (int e = grey_layout(H, W, g)) return e;
    *plan_bytes = g.plan_bytes, *work_bytes = g.work_bytes;
    return 0;
}

extern "C" int hhsr_grey_fft_plan(void *plan, int H, int W, hhsr_stream_t stream) {
    HHSR_REQUIRE(plan, "null pointer");
    HHSR_REQUIRE((uintptr_t)plan % 256 == 0, "plan must be 256-byte aligned");
    GreyLayout g;
    if (int e = grey_layout(H, W, g)) return e;
    char *p = static_cast<char *>(plan);
    cudaStream_t st = (cudaStream_t)stream;
    fft_tables_kernel<<<ceil_div(W, 256), 256, 0, st>>>(g.pw, (c32 *)(p + g.off_twW), (int *)(p + g.off_posW), (int *)(p + g.off_kposW));
    fft_tables_kernel<<<ceil_div(H, 256), 256, 0, st>>>(g.ph, (c32 *)(p + g.off_twH), (int *)(p + g.off_posH), (int *)(p + g.off_kposH));
    return launch_status("grey_fft_plan");
}

extern "C" int hhsr_grey_fft(const float *img, int H, int W, const void *plan, void *work, float *out, hhsr_stream_t stream) {
    HHSR_REQUIRE(img && plan && work && out, "null pointer");
    HHSR_REQUIRE((uintptr_t)plan % 256 == 0 && (uintptr_t)work % 16 == 0, "plan / work buffers misaligned");
    GreyLayout g;
    if (int e = grey_layout(H, W, g)) return e;
    const char *p = static_cast<const char *>(plan);
    c32 *spec = static_cast<c32 *>(work);
    cudaStream_t st = (cudaStream_t)stream;
    if (g.pw.pad && g.ph.pad) launch_grey<true, true>(g, img, H, W, p, spec, out, st);
    else if (g.pw.pad) launch_grey<true, false>(g, img, H, W, p, spec, out, st);
    else if (g.ph.pad) launch_grey<false, true>(g, img, H, W, p, spec, out, st);
    else launch_grey<false, false>(g, img, H, W, p, spec, out, st);
    return launch_status("grey_fft");
}
