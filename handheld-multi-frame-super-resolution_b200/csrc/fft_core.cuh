// Mixed-radix in-place FFT building blocks of the grey-image kernels (greyfft.cu).
//
// Everything here is __host__ __device__ and written per "thread" (tid, nthreads) with the barriers left to the caller:
// the kernels call a phase, then __syncthreads(); tests/native/greyfft_emul.cpp compiles the very same phases with g++
// and runs the threads of a CTA one after the other, so the index arithmetic (digit-reversed order, row pairing,
// band mask, pruning, padding) is checked against numpy.fft on the CPU before any GPU time is spent.
//
// Transform layout.  A length-n transform, n = r_0 r_1 ... r_{S-1}, runs IN PLACE as S decimation-in-frequency stages:
// stage s works on sub-blocks of length L_s = n / (r_0 ... r_{s-1}); the butterfly (b, j), j < m = L_s / r_s, reads
// x[b L + j + t m], t < r_s, takes their DFT_r and multiplies output t' by w_L^{j t'} — each butterfly reads and writes
// the same r locations, so a stage needs no second buffer and no ordering inside it.  Frequency k ends at the
// digit-reversed position pos(k) = sum_s t_s n / (r_0 ... r_s) with k = t_0 + r_0 (t_1 + r_1 (...)).  The inverse runs
// the same graph backwards (conjugate twiddle, then conjugate DFT_r, stages in reverse order): it takes digit-reversed
// input and delivers natural order, unnormalised.  Forward followed by inverse therefore needs no permutation at all
// (the column pass), and the row passes fold the permutation into their global <-> shared copies.
//
// Radices.  Few, large stages: a thread holds a whole butterfly of up to 32 points in registers — composite radices
// (16 = 4x4, 25 = 5x5, 20 = 4x5, ...) are two layers of small DFTs with compile-time twiddles in between — so a
// 4000-point row is 3 passes over shared memory (16 x 10 x 25) and a 3000-point column 3 (20 x 10 x 15).  Even radices
// run first (long strides), the odd one last: neighbouring butterflies of the last stage are then an odd number of
// complex words apart, which is conflict-free.  Sizes without an odd factor (8192, 6144) pad shared memory by one word
// per 16 instead (`pad`): logical index i lives at i + (i >> 4).
#pragma once

#include <utility>

#include "fft_consts.cuh"

#if defined(__CUDACC__)
#define FFT_HD __host__ __device__ __forceinline__
#else
#define FFT_HD inline
#endif

namespace hhsr {
namespace fft {

struct alignas(8) c32 {
    float x, y;
};
struct alignas(16) f4 {
    float x, y, z, w;
};

constexpr int kMaxStages = 10;
struct StageDesc {
    int radix, m, nol;     // sub-block length L = radix * m, nol = n / L
    unsigned magic;        // floor(2^32 / m) + 1: q / m == umulhi(q, magic) for q * m < 2^32 (m > 1)
};
struct Plan {
    int n, count, pad;     // pad: 1 -> logical index i is stored at i + (i >> 4)
    StageDesc st[kMaxStages];
};

FFT_HD int phys_len(int n, int pad) { return pad ? n + (n >> 4) + 1 : n; }
template <bool PAD>
FFT_HD int phys(int i) {
    return PAD ? i + (i >> 4) : i;
}

// ---- host: choose the radices.  Fewest stages; among those prefer a solution with an odd radix (it runs last and
// needs no padding), then the smallest largest radix (register pressure), then the largest smallest radix.
namespace detail {
constexpr int kNumRadices = 22;
constexpr int kRadices[kNumRadices] = {32, 25, 24, 21, 20, 19, 17, 16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2};
struct Search {
    int best[kMaxStages], nbest;
    long long best_score;
    int cur[kMaxStages];
    void go(int rem, int depth, int start) {
        if (rem == 1) {
            int has_odd = 0, mx = 0, mn = 1 << 30;
            for (int i = 0; i < depth; ++i) has_odd |= cur[i] & 1, mx = cur[i] > mx ? cur[i] : mx, mn = cur[i] < mn ? cur[i] : mn;
            const long long score = ((long long)depth << 24) | ((long long)(has_odd ? 0 : 1) << 16) | ((long long)mx << 8) | (63 - mn);
            if (nbest == 0 || score < best_score) {
                best_score = score, nbest = depth;
                for (int i = 0; i < depth; ++i) best[i] = cur[i];
            }
            return;
        }
        if (depth >= kMaxStages || (nbest && depth + 1 > nbest)) return;
        for (int i = start; i < kNumRadices; ++i)
            if (rem % kRadices[i] == 0) cur[depth] = kRadices[i], go(rem / kRadices[i], depth + 1, i);
    }
};
}  // namespace detail

inline bool make_plan(int n, Plan &p) {
    p.n = n, p.count = 0, p.pad = 0;
    if (n < 2 || n > 65536) return false;
    int rem = n;
    const int primes[8] = {2, 3, 5, 7, 11, 13, 17, 19};
    for (int q : primes)
        while (rem % q == 0) rem /= q;
    if (rem != 1) return false;
    detail::Search s;
    s.nbest = 0, s.best_score = 0;
    s.go(n, 0, 0);
    if (s.nbest == 0) return false;
    // even radices first (descending), odd ones last (ascending)
    int order[kMaxStages], cnt = 0;
    for (int pass = 0; pass < 2; ++pass)
        for (int v = (pass == 0 ? 32 : 2); pass == 0 ? v >= 2 : v <= 32; v += (pass == 0 ? -1 : 1))
            for (int i = 0; i < s.nbest; ++i)
                if (s.best[i] == v && ((v & 1) == pass)) order[cnt++] = v;
    p.count = cnt;
    int L = n;
    for (int i = 0; i < cnt; ++i) {
        StageDesc &d = p.st[i];
        d.radix = order[i], d.m = L / order[i], d.nol = n / L;
        d.magic = d.m > 1 ? (unsigned)((1ull << 32) / (unsigned)d.m + 1ull) : 0u;
        L = d.m;
    }
    p.pad = (order[cnt - 1] & 1) ? 0 : 1;
    return true;
}

FFT_HD int div_magic(int q, unsigned magic) {
#if defined(__CUDA_ARCH__)
    return (int)__umulhi((unsigned)q, magic);
#else
    return (int)(((unsigned long long)(unsigned)q * magic) >> 32);
#endif
}

// ---- compile-time loops (indices usable as template arguments and constexpr table subscripts)
template <class F, int... I>
FFT_HD void static_for_impl(F &&f, std::integer_sequence<int, I...>) {
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
FFT_HD void static_for(F &&f) {
    static_for_impl(f, std::make_integer_sequence<int, N>{});
}

FFT_HD c32 cmul(c32 a, c32 b) { return c32{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
FFT_HD c32 cadd(c32 a, c32 b) { return c32{a.x + b.x, a.y + b.y}; }
FFT_HD c32 csub(c32 a, c32 b) { return c32{a.x - b.x, a.y - b.y}; }
// multiplication by -i (forward) or +i (inverse)
template <bool INV>
FFT_HD c32 rot90(c32 a) {
    return INV ? c32{-a.y, a.x} : c32{a.y, -a.x};
}
FFT_HD c32 ldc(const c32 *p) {
#if defined(__CUDA_ARCH__)
    const float2 v = __ldg(reinterpret_cast<const float2 *>(p));
    return c32{v.x, v.y};
#else
    return *p;
#endif
}
// Bulk data (image rows, spectra) is read once: streaming loads (evict-first) keep it from pushing the twiddle and
// digit-reversal tables, which every butterfly gathers from, out of L1.
FFT_HD f4 ld4(const float *p) {   // 16-byte aligned
#if defined(__CUDA_ARCH__)
    const float4 v = __ldcs(reinterpret_cast<const float4 *>(p));
    return f4{v.x, v.y, v.z, v.w};
#else
    return f4{p[0], p[1], p[2], p[3]};
#endif
}
FFT_HD c32 ldc_stream(const c32 *p) {
#if defined(__CUDA_ARCH__)
    const float2 v = __ldcs(reinterpret_cast<const float2 *>(p));
    return c32{v.x, v.y};
#else
    return *p;
#endif
}
FFT_HD int ldi(const int *p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
FFT_HD bool aligned16(const void *p) { return (reinterpret_cast<unsigned long long>(p) & 15ull) == 0; }

// DFT of R points, e^{-2 pi i t t'/R} (INV: conjugate), in place: the prime and power-of-two kernels
template <int R, bool INV>
FFT_HD void dft_small(c32 (&v)[R]) {
    if constexpr (R == 2) {
        const c32 a = v[0], b = v[1];
        v[0] = cadd(a, b), v[1] = csub(a, b);
    } else if constexpr (R == 3) {
        const c32 t = cadd(v[1], v[2]), d = csub(v[1], v[2]);
        const c32 a = c32{v[0].x - 0.5f * t.x, v[0].y - 0.5f * t.y};
        const c32 b = rot90<INV>(c32{0.86602540378443865f * d.x, 0.86602540378443865f * d.y});
        v[0] = cadd(v[0], t), v[1] = cadd(a, b), v[2] = csub(a, b);
    } else if constexpr (R == 4) {
        const c32 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]), t2 = cadd(v[1], v[3]), t3 = rot90<INV>(csub(v[1], v[3]));
        v[0] = cadd(t0, t2), v[2] = csub(t0, t2), v[1] = cadd(t1, t3), v[3] = csub(t1, t3);
    } else if constexpr (R == 5) {
        const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f, s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
        const c32 t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]), t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
        const c32 a1 = c32{v[0].x + c1 * t1.x + c2 * t2.x, v[0].y + c1 * t1.y + c2 * t2.y};
        const c32 a2 = c32{v[0].x + c2 * t1.x + c1 * t2.x, v[0].y + c2 * t1.y + c1 * t2.y};
        const c32 b1 = rot90<INV>(c32{s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y});
        const c32 b2 = rot90<INV>(c32{s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y});
        v[0] = c32{v[0].x + t1.x + t2.x, v[0].y + t1.y + t2.y};
        v[1] = cadd(a1, b1), v[4] = csub(a1, b1), v[2] = cadd(a2, b2), v[3] = csub(a2, b2);
    } else if constexpr (R == 7) {
        const float c1 = 0.62348980185873353f, c2 = -0.22252093395631440f, c3 = -0.90096886790241913f;
        const float s1 = 0.78183148246802981f, s2 = 0.97492791218182361f, s3 = 0.43388373911755812f;
        const c32 p1 = cadd(v[1], v[6]), p2 = cadd(v[2], v[5]), p3 = cadd(v[3], v[4]);
        const c32 m1 = csub(v[1], v[6]), m2 = csub(v[2], v[5]), m3 = csub(v[3], v[4]);
        const c32 a1 = c32{v[0].x + c1 * p1.x + c2 * p2.x + c3 * p3.x, v[0].y + c1 * p1.y + c2 * p2.y + c3 * p3.y};
        const c32 a2 = c32{v[0].x + c2 * p1.x + c3 * p2.x + c1 * p3.x, v[0].y + c2 * p1.y + c3 * p2.y + c1 * p3.y};
        const c32 a3 = c32{v[0].x + c3 * p1.x + c1 * p2.x + c2 * p3.x, v[0].y + c3 * p1.y + c1 * p2.y + c2 * p3.y};
        const c32 b1 = rot90<INV>(c32{s1 * m1.x + s2 * m2.x + s3 * m3.x, s1 * m1.y + s2 * m2.y + s3 * m3.y});
        const c32 b2 = rot90<INV>(c32{s2 * m1.x - s3 * m2.x - s1 * m3.x, s2 * m1.y - s3 * m2.y - s1 * m3.y});
        const c32 b3 = rot90<INV>(c32{s3 * m1.x - s1 * m2.x + s2 * m3.x, s3 * m1.y - s1 * m2.y + s2 * m3.y});
        v[0] = c32{v[0].x + p1.x + p2.x + p3.x, v[0].y + p1.y + p2.y + p3.y};
        v[1] = cadd(a1, b1), v[6] = csub(a1, b1), v[2] = cadd(a2, b2), v[5] = csub(a2, b2), v[3] = cadd(a3, b3), v[4] = csub(a3, b3);
    } else {
        static_assert(R == 8, "unsupported radix");
        const float h = 0.70710678118654752f;
        c32 a[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = cadd(v[i], v[i + 4]), a[i + 4] = csub(v[i], v[i + 4]);
        // w8^1, w8^2, w8^3 on the difference terms (w8 = e^{-+ i pi/4})
        a[5] = INV ? c32{h * (a[5].x - a[5].y), h * (a[5].x + a[5].y)} : c32{h * (a[5].x + a[5].y), h * (a[5].y - a[5].x)};
        a[6] = rot90<INV>(a[6]);
        a[7] = INV ? c32{-h * (a[7].x + a[7].y), h * (a[7].x - a[7].y)} : c32{h * (a[7].y - a[7].x), -h * (a[7].x + a[7].y)};
        c32 e[4] = {a[0], a[1], a[2], a[3]}, o[4] = {a[4], a[5], a[6], a[7]};
        dft_small<4, INV>(e), dft_small<4, INV>(o);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[2 * i] = e[i], v[2 * i + 1] = o[i];
    }
}

// Prime radices 11, 13, 17, 19 (frame sizes such as 6240 x 4160 or 5472 x 3648): the direct transform on the symmetric
// and antisymmetric pairs, X_k = v_0 + sum_n cos(2 pi k n / R) (v_n + v_{R-n}) -+ i sum_n sin(2 pi k n / R) (v_n - v_{R-n})
// and X_{R-k} its mirror — (R-1)^2 real multiply-adds, about what a radix-16 butterfly costs for R = 13.
template <int R, bool INV>
FFT_HD void dft_prime(c32 (&v)[R]) {
    constexpr int HALF = (R - 1) / 2;
    c32 p[HALF], m[HALF];
    c32 x0 = v[0];
    static_for<HALF>([&](auto N) {
        constexpr int n = decltype(N)::value;
        p[n] = cadd(v[n + 1], v[R - 1 - n]), m[n] = csub(v[n + 1], v[R - 1 - n]);
        x0 = cadd(x0, p[n]);
    });
    c32 out[R];
    out[0] = x0;
    static_for<HALF>([&](auto K) {
        constexpr int k = decltype(K)::value + 1;
        c32 a = v[0], b = c32{0.f, 0.f};
        static_for<HALF>([&](auto N) {
            constexpr int n = decltype(N)::value + 1;
            constexpr float c = Wtab<R>::c[(k * n) % R], s = Wtab<R>::s[(k * n) % R];
            a = c32{a.x + c * p[n - 1].x, a.y + c * p[n - 1].y};
            b = c32{b.x + s * m[n - 1].x, b.y + s * m[n - 1].y};
        });
        b = rot90<INV>(b);
        out[k] = cadd(a, b), out[R - k] = csub(a, b);
    });
    static_for<R>([&](auto T) { v[decltype(T)::value] = out[decltype(T)::value]; });
}

// multiplication by the compile-time twiddle e^{-+ 2 pi i N / R}
template <int R, int N, bool INV>
FFT_HD c32 mul_const(c32 a) {
    constexpr float c = Wtab<R>::c[N % R], s = INV ? Wtab<R>::s[N % R] : -Wtab<R>::s[N % R];
    if constexpr (s == 0.f && c == 1.f) return a;
    else if constexpr (s == 0.f && c == -1.f) return c32{-a.x, -a.y};
    else if constexpr (c == 0.f && s == 1.f) return c32{-a.y, a.x};
    else if constexpr (c == 0.f && s == -1.f) return c32{a.y, -a.x};
    else return c32{a.x * c - a.y * s, a.x * s + a.y * c};
}

// Composite radix R = A * B in registers: B transforms of A points (over t_A, t = t_B + B t_A), the twiddles
// w_R^{t_B k_A}, then A transforms of B points; output k = k_A + A k_B.
template <int A, int B, bool INV>
FFT_HD void dft_composite(c32 (&v)[A * B]) {
    constexpr int R = A * B;
    c32 u[R];
    static_for<B>([&](auto TB) {
        constexpr int tb = decltype(TB)::value;
        c32 tmp[A];
        static_for<A>([&](auto TA) { tmp[decltype(TA)::value] = v[tb + B * decltype(TA)::value]; });
        dft_small<A, INV>(tmp);
        static_for<A>([&](auto KA) {
            constexpr int ka = decltype(KA)::value;
            u[tb + B * ka] = mul_const<R, tb * ka, INV>(tmp[ka]);
        });
    });
    static_for<A>([&](auto KA) {
        constexpr int ka = decltype(KA)::value;
        c32 tmp[B];
        static_for<B>([&](auto TB) { tmp[decltype(TB)::value] = u[decltype(TB)::value + B * ka]; });
        dft_small<B, INV>(tmp);
        static_for<B>([&](auto KB) { v[ka + A * decltype(KB)::value] = tmp[decltype(KB)::value]; });
    });
}

template <int R, bool INV>
FFT_HD void dft(c32 (&v)[R]) {
    if constexpr (R == 6) dft_composite<2, 3, INV>(v);
    else if constexpr (R == 9) dft_composite<3, 3, INV>(v);
    else if constexpr (R == 10) dft_composite<2, 5, INV>(v);
    else if constexpr (R == 12) dft_composite<4, 3, INV>(v);
    else if constexpr (R == 14) dft_composite<2, 7, INV>(v);
    else if constexpr (R == 15) dft_composite<3, 5, INV>(v);
    else if constexpr (R == 16) dft_composite<4, 4, INV>(v);
    else if constexpr (R == 20) dft_composite<4, 5, INV>(v);
    else if constexpr (R == 21) dft_composite<3, 7, INV>(v);
    else if constexpr (R == 24) dft_composite<8, 3, INV>(v);
    else if constexpr (R == 25) dft_composite<5, 5, INV>(v);
    else if constexpr (R == 32) dft_composite<8, 4, INV>(v);
    else if constexpr (R == 11 || R == 13 || R == 17 || R == 19) dft_prime<R, INV>(v);
    else dft_small<R, INV>(v);
}

// w^1 ... w^{R-1} for w = tw[s] (tw[k] = e^{-2 pi i k / n}): the powers of two are table entries, every other power
// the product of its highest power of two and the (already formed) rest — at most 4 roundings deep for R = 32, while
// R - 1 gathers would cost more than the butterfly and a chain of R - 1 products would lose a digit.
template <int R, bool CONJ>
FFT_HD void twiddle_powers(const c32 *tw, int s, c32 (&w)[R]) {
    static_for<R>([&](auto T) {
        constexpr int t = decltype(T)::value;
        if constexpr (t >= 1 && (t & (t - 1)) == 0) {
            w[t] = ldc(tw + t * s);
            if (CONJ) w[t].y = -w[t].y;
        }
    });
    static_for<R>([&](auto T) {
        constexpr int t = decltype(T)::value;
        if constexpr (t >= 3 && (t & (t - 1)) != 0) {
            constexpr int hb = t >= 16 ? 16 : t >= 8 ? 8 : t >= 4 ? 4 : 2;
            w[t] = cmul(w[hb], w[t - hb]);
        }
    });
}

// One butterfly q (0 <= q < n / R) of a stage, forward (DIF) or inverse (DIT), on x (physical addressing per PAD).
template <int R, bool INV, bool PAD>
FFT_HD void butterfly(c32 *x, const c32 *tw, const StageDesc &d, int q) {
    const int m = d.m;
    c32 v[R];
    if (m == 1) {
        const int base = q * R;
        static_for<R>([&](auto T) { v[decltype(T)::value] = x[phys<PAD>(base + decltype(T)::value)]; });
        dft<R, INV>(v);
        static_for<R>([&](auto T) { x[phys<PAD>(base + decltype(T)::value)] = v[decltype(T)::value]; });
        return;
    }
    const int b = div_magic(q, d.magic), j = q - b * m;
    const int base = b * (m * R) + j;
    static_for<R>([&](auto T) { v[decltype(T)::value] = x[phys<PAD>(base + decltype(T)::value * m)]; });
    c32 w[R];
    twiddle_powers<R, INV>(tw, d.nol * j, w);
    if (INV) {
        static_for<R>([&](auto T) {
            if constexpr (decltype(T)::value >= 1) v[decltype(T)::value] = cmul(v[decltype(T)::value], w[decltype(T)::value]);
        });
        dft<R, true>(v);
    } else {
        dft<R, false>(v);
        static_for<R>([&](auto T) {
            if constexpr (decltype(T)::value >= 1) v[decltype(T)::value] = cmul(v[decltype(T)::value], w[decltype(T)::value]);
        });
    }
    static_for<R>([&](auto T) { x[phys<PAD>(base + decltype(T)::value * m)] = v[decltype(T)::value]; });
}

template <int R, bool INV, bool PAD>
FFT_HD void stage_loop(c32 *x, const c32 *tw, const StageDesc &d, int nb, int tid, int nt) {
    for (int q = tid; q < nb; q += nt) butterfly<R, INV, PAD>(x, tw, d, q);
}

// All butterflies of stage s that thread `tid` of `nt` owns.  Caller puts a barrier between stages.
template <bool INV, bool PAD>
FFT_HD void run_stage(c32 *x, const c32 *tw, const Plan &p, int s, int tid, int nt) {
    const StageDesc d = p.st[s];
    const int nb = p.n / d.radix;
    switch (d.radix) {
    case 2: stage_loop<2, INV, PAD>(x, tw, d, nb, tid, nt); break;
    case 3: stage_loop<3, INV, PAD>(x, tw, d, nb, tid, nt); break;
    case 4: stage_loop<4, INV, PAD>(x, tw, d, nb, tid, nt); break;
    case 5: stage_loop<5, INV, PAD>(x, tw, d, nb, tid, nt); break;
    case 6: stage_loop<6, INV, PAD>(x, tw, d, nb, tid, nt); break;
    case 7: stage_loop<7, INV, PAD>(x, tw, d, nb, tid, nt); break;
    case 8: stage_loop<8, INV, PAD>(x, tw, d, nb, tid, nt); break;
    case 9: stage_loop<9, INV, PAD>(x, tw, d, nb, tid, nt); break;
    case 10: stage_loop<10, INV, PAD>(x, tw, d, nb, tid, nt); break;
    case 11: stage_loop<11, INV, PAD>(x, tw, d, nb, tid, nt); break;
    case 12: stage_loop<12, INV, PAD>(x, tw, d, nb, tid, nt); break;
    case 13: stage_loop<13, INV, PAD>(x, tw, d, nb, tid, nt); break;
    case 14: stage_loop<14, INV, PAD>(x, tw, d, nb, tid, nt); break;
    case 15: stage_loop<15, INV, PAD>(x, tw, d, nb, tid, nt); break;
    case 16: stage_loop<16, INV, PAD>(x, tw, d, nb, tid, nt); break;
    case 17: stage_loop<17, INV, PAD>(x, tw, d, nb, tid, nt); break;
    case 19: stage_loop<19, INV, PAD>(x, tw, d, nb, tid, nt); break;
    case 20: stage_loop<20, INV, PAD>(x, tw, d, nb, tid, nt); break;
    case 21: stage_loop<21, INV, PAD>(x, tw, d, nb, tid, nt); break;
    case 24: stage_loop<24, INV, PAD>(x, tw, d, nb, tid, nt); break;
    case 25: stage_loop<25, INV, PAD>(x, tw, d, nb, tid, nt); break;
    default: stage_loop<32, INV, PAD>(x, tw, d, nb, tid, nt); break;
    }
}

// logical position of frequency k after the forward stages (and the position the inverse stages expect it at)
FFT_HD int digit_reversed(const Plan &p, int k) {
    int rem = k, pos = 0;
    for (int s = 0; s < p.count; ++s) {
        const int r = p.st[s].radix;
        pos += (rem % r) * p.st[s].m;
        rem /= r;
    }
    return pos;
}

// ---- the reference's band mask (utils_image.py:92-95) on the UNSHIFTED frequency index k of an axis of length n:
// shifted index s = (k + n/2) mod n is kept iff n/4 <= s < n - ceil(n/4)
FFT_HD bool band_keep(int k, int n) {   // 0 <= k <= n
    int s = k + n / 2;
    s -= (s >= n) ? n : 0;
    s -= (s >= n) ? n : 0;   // k == n (the partner of k == 0)
    return s >= n / 4 && s < n - (n + 3) / 4;
}
// Re(ifft2(M F)) == ifft2(0.5 (M(k) + M(-k)) F) for a real image: the weight of half-spectrum entry (ky, kx)
FFT_HD float mask_weight(int ky, int H, int kx, int W) {
    const float a = (band_keep(ky, H) && band_keep(kx, W)) ? 0.5f : 0.f;
    const float b = (band_keep(H - ky, H) && band_keep(W - kx, W)) ? 0.5f : 0.f;
    return a + b;
}
// number of leading half-spectrum columns that can be non-zero after the mask (the rest is never computed)
inline int kept_columns(int W) {
    int last = 0;
    for (int kx = 0; kx <= W / 2; ++kx)
        if (band_keep(kx, W) || band_keep(W - kx, W)) last = kx;
    return last + 1;
}

// ---- row passes.  Two real rows a, b ride one complex transform z = a + i b (W points):
//   A[k] = (Z[k] + conj Z[W-k]) / 2,  B[k] = (Z[k] - conj Z[W-k]) / (2i);  the 1/2 is folded into the mask scale.
// ppos_of_k[k]: PHYSICAL shared-memory position of frequency k (digit reversal and padding folded into one table).
// Global reads of every copy phase are issued in batches BEFORE the first dependent shared-memory store, so a thread has
// kBatch (vector) loads in flight instead of one: these phases are pure latency otherwise.
constexpr int kBatch = 4;

template <bool PAD>
FFT_HD void rows_load_pair(c32 *s, const float *a, const float *b, int W, int tid, int nt) {
    if ((W & 3) == 0 && aligned16(a) && aligned16(b)) {
        const int n4 = W >> 2;
        for (int c0 = tid; c0 < n4; c0 += kBatch * nt) {
            f4 va[kBatch], vb[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u)
                if (c0 + u * nt < n4) va[u] = ld4(a + 4 * (c0 + u * nt)), vb[u] = ld4(b + 4 * (c0 + u * nt));
#pragma unroll
            for (int u = 0; u < kBatch; ++u)
                if (c0 + u * nt < n4) {
                    c32 *d = s + phys<PAD>(4 * (c0 + u * nt));      // 4 consecutive logical indices never straddle a pad slot
                    d[0] = c32{va[u].x, vb[u].x}, d[1] = c32{va[u].y, vb[u].y}, d[2] = c32{va[u].z, vb[u].z}, d[3] = c32{va[u].w, vb[u].w};
                }
        }
        return;
    }
    for (int x = tid; x < W; x += nt) s[phys<PAD>(x)] = c32{a[x], b[x]};
}
// writes 2A[k] -> specA[k], 2B[k] -> specB[k] for k < KX, zeros up to KXp
FFT_HD void rows_store_half_spectra(const c32 *s, const int *ppos_of_k, c32 *specA, c32 *specB, int W, int KX, int KXp, int tid,
                                    int nt) {
    for (int k0 = tid; k0 < KXp; k0 += kBatch * nt) {
        int p[kBatch], pn[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int k = k0 + u * nt;
            if (k < KX) p[u] = ldi(ppos_of_k + k), pn[u] = ldi(ppos_of_k + (k == 0 ? 0 : W - k));
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int k = k0 + u * nt;
            if (k >= KXp) continue;
            c32 A = c32{0.f, 0.f}, B = A;
            if (k < KX) {
                const c32 z = s[p[u]], zn = s[pn[u]];
                A = c32{z.x + zn.x, z.y - zn.y};
                B = c32{z.y + zn.y, zn.x - z.x};
            }
            specA[k] = A, specB[k] = B;
        }
    }
}
// inverse: Z[k] = A[k] + i B[k], Z[W-k] = conj A[k] + i conj B[k] at their digit-reversed positions (s zeroed before);
// the imaginary parts of A[0], B[0] are dropped like a C2R transform does
FFT_HD void rows_zero(c32 *s, int len, int tid, int nt) {
    for (int x = tid; x < len; x += nt) s[x] = c32{0.f, 0.f};
}
FFT_HD void rows_scatter_half_spectra(c32 *s, const int *ppos_of_k, const c32 *specA, const c32 *specB, int W, int KX, int tid,
                                      int nt) {
    for (int k0 = tid; k0 < KX; k0 += kBatch * nt) {
        c32 A[kBatch], B[kBatch];
        int p[kBatch], pn[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int k = k0 + u * nt;
            if (k < KX) A[u] = ldc_stream(specA + k), B[u] = ldc_stream(specB + k), p[u] = ldi(ppos_of_k + k), pn[u] = ldi(ppos_of_k + (k == 0 ? 0 : W - k));
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int k = k0 + u * nt;
            if (k >= KX) continue;
            if (k == 0) {
                s[p[u]] = c32{A[u].x, B[u].x};
            } else {
                s[p[u]] = c32{A[u].x - B[u].y, A[u].y + B[u].x};
                if (2 * k != W) s[pn[u]] = c32{A[u].x + B[u].y, B[u].x - A[u].y};
            }
        }
    }
}
template <bool PAD>
FFT_HD void rows_store_pair(const c32 *s, float *a, float *b, int W, int tid, int nt) {
    if ((W & 3) == 0 && aligned16(a) && aligned16(b)) {
        for (int c = tid; c < (W >> 2); c += nt) {
            const c32 *z = s + phys<PAD>(4 * c);
            *reinterpret_cast<f4 *>(a + 4 * c) = f4{z[0].x, z[1].x, z[2].x, z[3].x};
            *reinterpret_cast<f4 *>(b + 4 * c) = f4{z[0].y, z[1].y, z[2].y, z[3].y};
        }
        return;
    }
    for (int x = tid; x < W; x += nt) {
        const c32 z = s[phys<PAD>(x)];
        a[x] = z.x, b[x] = z.y;
    }
}

// ---- column pass on a tile of CW half-spectrum columns held column-major in shared memory, s[c * Hp + phys(y)].
// Copies: thread -> (column c = tid % CW, row y0 = tid / CW), rows advance by nt / CW (threads beyond that sit out).
template <bool PAD>
FFT_HD void cols_load_tile(c32 *s, const c32 *spec, long long pitch, int H, int Hp, int c0, int CW, int tid, int nt) {
    const int step = nt / CW, y0 = tid / CW, c = tid - y0 * CW;
    if (y0 >= step) return;
    constexpr int kB = 2 * kBatch;
    for (int yb = y0; yb < H; yb += kB * step) {
        c32 v[kB];
#pragma unroll
        for (int u = 0; u < kB; ++u)
            if (yb + u * step < H) v[u] = ldc_stream(spec + (long long)(yb + u * step) * pitch + c0 + c);
#pragma unroll
        for (int u = 0; u < kB; ++u)
            if (yb + u * step < H) s[c * Hp + phys<PAD>(yb + u * step)] = v[u];
    }
}
template <bool PAD>
FFT_HD void cols_store_tile(const c32 *s, c32 *spec, long long pitch, int H, int Hp, int c0, int CW, int tid, int nt) {
    const int step = nt / CW, y0 = tid / CW, c = tid - y0 * CW;
    if (y0 >= step) return;
    for (int y = y0; y < H; y += step) spec[(long long)y * pitch + c0 + c] = s[c * Hp + phys<PAD>(y)];
}
// band mask (times `scale`) on one digit-reversed column spectrum (column kx), by the `per` threads that own the column
template <bool PAD>
FFT_HD void cols_mask_column(c32 *col, const int *k_of_pos, int H, int W, int kx, float scale, int ctid, int per) {
    for (int p0 = ctid; p0 < H; p0 += kBatch * per) {
        int ky[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u)
            if (p0 + u * per < H) ky[u] = ldi(k_of_pos + p0 + u * per);
#pragma unroll
        for (int u = 0; u < kBatch; ++u)
            if (p0 + u * per < H) {
                const float m = mask_weight(ky[u], H, kx, W) * scale;
                c32 &z = col[phys<PAD>(p0 + u * per)];
                z = c32{z.x * m, z.y * m};
            }
    }
}

}  // namespace fft
}  // namespace hhsr
