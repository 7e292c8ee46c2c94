// CPU emulation of the three grey-image FFT kernels (csrc/greyfft.cu): the SAME per-thread phases (csrc/fft_core.cuh,
// compiled here with g++), the threads of a CTA run one after the other, a barrier being the end of each loop.
// Test infrastructure only (tests/test_greyfft_emul.py compares it with numpy.fft); never part of the product.
#include <cmath>
#include <cstdlib>
#include <vector>

#include "fft_core.cuh"

using namespace hhsr::fft;

static void tables(const Plan &pl, std::vector<c32> &tw, std::vector<int> &ppos_of_k, std::vector<int> &k_of_pos) {
    tw.resize(pl.n), ppos_of_k.resize(pl.n), k_of_pos.resize(pl.n);
    const double pi = 3.14159265358979323846;
    for (int k = 0; k < pl.n; ++k) {
        const double a = 2.0 * pi * (double)k / (double)pl.n;
        tw[k] = c32{(float)std::cos(a), (float)-std::sin(a)};
        const int p = digit_reversed(pl, k);
        ppos_of_k[k] = pl.pad ? phys<true>(p) : p, k_of_pos[p] = k;
    }
}

template <bool INV>
static void stage(c32 *x, const c32 *tw, const Plan &pl, int s, int t, int nt) {
    if (pl.pad) run_stage<INV, true>(x, tw, pl, s, t, nt);
    else run_stage<INV, false>(x, tw, pl, s, t, nt);
}

extern "C" int emul_factorize(int n, int *radix, int *pad) {
    Plan pl;
    if (!make_plan(n, pl)) return -1;
    for (int i = 0; i < pl.count; ++i) radix[i] = pl.st[i].radix;
    *pad = pl.pad;
    return pl.count;
}

// plain transform of one complex vector (interleaved), natural order in and out: checks stages + digit reversal
extern "C" int emul_fft(float *data, int n, int inverse, int nt) {
    Plan pl;
    if (!make_plan(n, pl)) return -1;
    std::vector<c32> tw;
    std::vector<int> ppos, kpos;
    tables(pl, tw, ppos, kpos);
    std::vector<c32> s(phys_len(n, pl.pad), c32{NAN, NAN});
    c32 *x = reinterpret_cast<c32 *>(data);
    if (!inverse) {
        for (int i = 0; i < n; ++i) s[pl.pad ? phys<true>(i) : i] = x[i];
        for (int i = 0; i < pl.count; ++i)
            for (int t = 0; t < nt; ++t) stage<false>(s.data(), tw.data(), pl, i, t, nt);
        for (int k = 0; k < n; ++k) x[k] = s[ppos[k]];
    } else {
        for (int k = 0; k < n; ++k) s[ppos[k]] = x[k];
        for (int i = pl.count - 1; i >= 0; --i)
            for (int t = 0; t < nt; ++t) stage<true>(s.data(), tw.data(), pl, i, t, nt);
        for (int i = 0; i < n; ++i) x[i] = s[pl.pad ? phys<true>(i) : i];
    }
    return 0;
}

extern "C" int emul_kept_columns(int W) { return kept_columns(W); }

// the whole grey image, kernel by kernel
extern "C" int emul_grey(const float *img, int H, int W, float *out, int CW, int row_threads, int col_threads) {
    Plan pw, ph;
    if ((H & 1) || !make_plan(W, pw) || !make_plan(H, ph)) return -1;
    std::vector<c32> twW, twH;
    std::vector<int> posW, kposW, posH, kposH;
    tables(pw, twW, posW, kposW), tables(ph, twH, posH, kposH);
    const int KX = kept_columns(W), KXp = (KX + CW - 1) / CW * CW;
    int Hp = phys_len(H, ph.pad);
    while (CW > 1 && Hp % 16 != (16 / CW) % 16) ++Hp;
    std::vector<c32> spec((size_t)H * KXp, c32{NAN, NAN});
    const int wlen = phys_len(W, pw.pad);
    // 1. rows forward: one CTA per row pair
    for (int cta = 0; cta < H / 2; ++cta) {
        std::vector<c32> s(wlen, c32{NAN, NAN});
        const size_t r = 2 * (size_t)cta;
        const int nt = row_threads;
        for (int t = 0; t < nt; ++t) {
            if (pw.pad) rows_load_pair<true>(s.data(), img + r * W, img + (r + 1) * W, W, t, nt);
            else rows_load_pair<false>(s.data(), img + r * W, img + (r + 1) * W, W, t, nt);
        }
        for (int i = 0; i < pw.count; ++i)
            for (int t = 0; t < nt; ++t) stage<false>(s.data(), twW.data(), pw, i, t, nt);
        for (int t = 0; t < nt; ++t)
            rows_store_half_spectra(s.data(), posW.data(), spec.data() + r * KXp, spec.data() + (r + 1) * KXp, W, KX, KXp, t, nt);
    }
    // 2. columns: one CTA per tile of CW columns
    const float scale = 0.5f / ((float)H * (float)W);
    for (int cta = 0; cta < KXp / CW; ++cta) {
        std::vector<c32> s((size_t)CW * Hp, c32{NAN, NAN});
        const int nt = col_threads, c0 = cta * CW, per = nt / CW;
        for (int t = 0; t < nt; ++t) {
            if (ph.pad) cols_load_tile<true>(s.data(), spec.data(), KXp, H, Hp, c0, CW, t, nt);
            else cols_load_tile<false>(s.data(), spec.data(), KXp, H, Hp, c0, CW, t, nt);
        }
        for (int i = 0; i < ph.count; ++i)
            for (int t = 0; t < nt; ++t)
                if (t / per < CW) stage<false>(s.data() + (t / per) * Hp, twH.data(), ph, i, t % per, per);
        for (int t = 0; t < nt; ++t) {
            if (t / per >= CW) continue;
            if (ph.pad) cols_mask_column<true>(s.data() + (t / per) * Hp, kposH.data(), H, W, c0 + t / per, scale, t % per, per);
            else cols_mask_column<false>(s.data() + (t / per) * Hp, kposH.data(), H, W, c0 + t / per, scale, t % per, per);
        }
        for (int i = ph.count - 1; i >= 0; --i)
            for (int t = 0; t < nt; ++t)
                if (t / per < CW) stage<true>(s.data() + (t / per) * Hp, twH.data(), ph, i, t % per, per);
        for (int t = 0; t < nt; ++t) {
            if (ph.pad) cols_store_tile<true>(s.data(), spec.data(), KXp, H, Hp, c0, CW, t, nt);
            else cols_store_tile<false>(s.data(), spec.data(), KXp, H, Hp, c0, CW, t, nt);
        }
    }
    // 3. rows inverse
    for (int cta = 0; cta < H / 2; ++cta) {
        std::vector<c32> s(wlen, c32{NAN, NAN});
        const size_t r = 2 * (size_t)cta;
        const int nt = row_threads;
        for (int t = 0; t < nt; ++t) rows_zero(s.data(), wlen, t, nt);
        for (int t = 0; t < nt; ++t)
            rows_scatter_half_spectra(s.data(), posW.data(), spec.data() + r * KXp, spec.data() + (r + 1) * KXp, W, KX, t, nt);
        for (int i = pw.count - 1; i >= 0; --i)
            for (int t = 0; t < nt; ++t) stage<true>(s.data(), twW.data(), pw, i, t, nt);
        for (int t = 0; t < nt; ++t) {
            if (pw.pad) rows_store_pair<true>(s.data(), out + r * W, out + (r + 1) * W, W, t, nt);
            else rows_store_pair<false>(s.data(), out + r * W, out + (r + 1) * W, W, t, nt);
        }
    }
    return 0;
}
