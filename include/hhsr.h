/* hhsr.h — C ABI of libhhsr.so: the B200-native (sm_100a) handheld burst super-resolution hot path.
 *
 * The reference (Jamy-L/Handheld-Multi-Frame-Super-Resolution) has no FFI layer: its hot path is a set of Python
 * stage functions that launch Numba-CUDA kernels.  Each entry point below replaces the
 * device work of one of those stage functions; the reference interface it replaces is cited as
 * `handheld_super_resolution/<file>:<line>`.  INTEGRATION.md shows the ctypes stub a maintainer of the
 * reference would add to call them.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to a C-contiguous array unless the parameter is documented "host";
 *   - images are [rows][cols] float32; flow fields are [ny][nx][2] float32 holding (dx, dy);
 *   - the caller owns all memory (inputs, outputs, workspaces); the library keeps no device state;
 *   - every call only enqueues work on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream)
 *     and never synchronises;
 *   - return value: 0 on success, a negative HHSR_E_* code for rejected arguments, a positive cudaError_t when
 *     a launch failed; hhsr_last_error_string() (thread-local) describes the last failure.
 */
#ifndef HHSR_H
#define HHSR_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HHSR_VERSION 100 /* 0.1.0 */

#define HHSR_E_BADARG (-1)      /* null pointer, non-positive size, misaligned buffer */
#define HHSR_E_UNSUPPORTED (-2) /* tile size / radius / mode outside what the reference supports */

typedef void *hhsr_stream_t;

int hhsr_version(void);
const char *hhsr_last_error_string(void);

/* ---- grey image, Alg. 3 (utils_image.py:82-100).  The forward/inverse FFTs stay cuFFT calls made by the host
 * (torch.fft.rfft2 / irfft2); this applies the reference's band mask to the half spectrum in place:
 * spec[ky][kx] *= 0.5*(My(ky)Mx(kx) + My(-ky)Mx(-kx)), the Hermitian-symmetrised form of the mask the reference
 * applies to the full shifted spectrum before taking .real.  spec: [H][W/2+1] complex64 (interleaved re,im) with
 * strides (stride_y, stride_x) in complex elements — cuFFT-through-torch hands back a column-major spectrum.
 * `scale` multiplies the kept entries as well: pass 1/(H*W) with an UNNORMALISED inverse transform
 * (irfft2(norm="forward")) and the separate normalisation pass over the image disappears; pass 1 otherwise. */
int hhsr_grey_band_mask(float *spec, int H, int W, long long stride_y, long long stride_x, float scale,
                        hhsr_stream_t stream);

/* ---- grey image, Alg. 3, as a whole (utils_image.py:82-100: fft2, fftshift, four masked fills, ifftshift, ifft2,
 * .real) with the library's own shared-memory FFT passes: rows forward (two real rows per complex transform, only the
 * W/4 + 1 half-spectrum columns the mask keeps are stored), columns (forward, band mask, inverse in one pass), rows
 * inverse.  Sizes must factor into primes up to 19 with H even (HHSR_E_UNSUPPORTED otherwise: the caller keeps the cuFFT
 * route with hhsr_grey_band_mask for those).
 *   hhsr_grey_fft_sizes: bytes of the two caller-owned device buffers for an H x W image (host pointers out);
 *   hhsr_grey_fft_plan:  fills `plan` (twiddle tables rounded from float64, digit-reversal tables) once per (H, W);
 *   hhsr_grey_fft:       out[H][W] = grey(img[H][W]); `work` is scratch (the pruned spectrum), `plan` read-only. */
int hhsr_grey_fft_sizes(int H, int W, size_t *plan_bytes, size_t *work_bytes);
int hhsr_grey_fft_plan(void *plan, int H, int W, hhsr_stream_t stream);
int hhsr_grey_fft(const float *img, int H, int W, const void *plan, void *work, float *out, hhsr_stream_t stream);

/* ---- Gaussian pyramid (alignment.py:74-82, utils_image.py:360-391) */
/* circular padding of the reference grey image to a multiple of the tile size (alignment.py:26-37) */
int hhsr_pad_circular(const float *src, int h, int w, float *dst, int hp, int wp, hhsr_stream_t stream);
/* one pyramid level: separable valid correlation with `taps` (host pointer, 2*radius+1 floats, radius <= 16),
 * y pass then x pass, keeping only outputs (i*factor, j*factor): dst[h2][w2], h2 = (h-2r)/factor. */
int hhsr_gauss_downsample(const float *src, int h, int w, int factor, const float *taps_host, int radius,
                          float *dst, int h2, int w2, hhsr_stream_t stream);

/* ---- ICA initialisation (ICA.py:15-76): central-difference gradients (zero outside) and the per-tile 2x2
 * Hessian [ny][nx][2][2], ny = h/ts, nx = w/ts. */
int hhsr_grad_hessian(const float *img, int h, int w, int ts, float *gradx, float *grady, float *hessian,
                      hhsr_stream_t stream);

/* ---- flow upscaling between pyramid levels (alignment.py:150-172): out = factor * upsample(in, repeat),
 * zero-padded (or cropped) to [ny_out][nx_out].  mode: 0 nearest, 1 bilinear, 2 bicubic (torch
 * F.interpolate semantics, align_corners=False). */
int hhsr_upscale_flow(const float *flow_in, int ny_in, int nx_in, float *flow_out, int ny_out, int nx_out,
                      int repeat, float factor, int mode, hhsr_stream_t stream);

/* ---- L2 block matching (block_matching.py:20-76, 348-378): per tile, exhaustive search of
 * sum(m^2) - 2 sum(ref*m) over (2r+1)^2 integer shifts around rint(flow), moving image read with clamped
 * coordinates, first minimum in v-major order; flow += (u*, v*) in place.  ts in {8,16,32,64}, radius <= 8. */
int hhsr_bm_l2_search(const float *ref, int ref_h, int ref_w, const float *mov, int mov_h, int mov_w,
                      float *flow, int ny, int nx, int ts, int radius, hhsr_stream_t stream);
/* ---- "L1" level as the compiled reference executes it (block_matching.py:78-345): the SAD search
 * result is discarded and flow <- rint(flow) (half to even).  n = ny*nx*2 floats. */
int hhsr_bm_l1_compat(float *flow, int n, hhsr_stream_t stream);
/* ---- the L1 level as the reference INTENDS it (block_matching.py:78-345, never reached by the compiled reference —
 * SURVEY Q1; selected by block_matching.tuning.l1_compat = false only): per tile, exhaustive search of sum |ref - m|
 * over (2r+1)^2 integer shifts around rint(flow), moving samples outside the frame read as 0, sums in float64, first
 * minimum in v-major order; flow <- rint(flow) + (u*, v*) in place.  ts in {16,32,64}, radius <= 8. */
int hhsr_bm_l1_search(const float *ref, int ref_h, int ref_w, const float *mov, int mov_h, int mov_w,
                      float *flow, int ny, int nx, int ts, int radius, hhsr_stream_t stream);

/* ---- ICA / Lucas-Kanade refinement (ICA.py:78-481): n_iter Gauss-Newton steps per tile, in place on flow.
 * ts selects the reference kernel being reproduced: 8 (clamped sampling, fp64 1/det), 16 and 32 (zero fill),
 * 64 (zero fill + sliding-window row quirk).
 * gradx == grady == NULL (ts 32, reference level a whole number of tiles, ref_w % 4 == 0): the gradients are taken to be the
 * central differences of `ref` that hhsr_grad_hessian writes and are re-formed inside the kernel (same values, two full
 * planes less to read). */
int hhsr_ica(const float *ref, const float *gradx, const float *grady, int ref_h, int ref_w,
             const float *hessian, const float *mov, int mov_h, int mov_w, float *flow, int ny, int nx, int ts,
             int n_iter, hhsr_stream_t stream);

/* ---- steering kernels, Alg. 5 (kernels.py:29-243, linalg.py:86-185, utils_image.py:117-170,346-357): fused
 * GAT -> 2x2 decimation -> gradients -> structure tensor -> eigen-decomposition -> covariance.
 * covs: [H/2][W/2][2][2].  law: 0 hard_threshold, 1 linear. */
int hhsr_estimate_kernels(const float *raw, int H, int W, double alpha, double beta, double k_detail,
                          double k_denoise, double D_th, double D_tr, double k_stretch, double k_shrink, int law,
                          float *covs, hhsr_stream_t stream);

/* stand-alone pieces of the same stage, kept because the reference exposes them (utils_image.py:117-170 GAT,
 * utils_image.py:346-357 compute_grey_images(method="decimating")); hhsr_estimate_kernels does not need them. */
int hhsr_gat(const float *img, size_t n, double alpha, double beta, float *out, hhsr_stream_t stream);
int hhsr_decimate_to_grey(const float *img, int H, int W, float *out, hhsr_stream_t stream);

/* ---- robustness, Alg. 6-9 (robustness.py).  cfa_host: 4 ints (row-major 2x2 channel ids), wb_host: 3 doubles. */
/* guide image + 3x3 local statistics at half resolution (robustness.py:173-294): means/vars [3][H/2][W/2];
 * vars may be NULL. */
int hhsr_guide_stats(const float *raw, int H, int W, const int *cfa_host, const double *wb_host, float *means,
                     float *vars, hhsr_stream_t stream);
/* the two halves of hhsr_guide_stats as stand-alone stages of the reference API: guide image [3][H/2][W/2]
 * (robustness.py:173-226) and 3x3 edge-replicated mean / variance of a [channels][h][w] image (robustness.py:228-294);
 * their composition is bit-equal to hhsr_guide_stats. */
int hhsr_guide_image(const float *raw, int H, int W, const int *cfa_host, const double *wb_host, float *guide,
                     hhsr_stream_t stream);
int hhsr_local_stats(const float *guide, int channels, int h, int w, float *means, float *vars, hhsr_stream_t stream);
/* x2 Dodgson upsampling (+ tile-flow warp when flow != NULL) of a [3][h][w] statistic to [3][2h][2w]
 * (robustness.py:296-418); +inf where the source position leaves the guide image. */
int hhsr_upscale_warp_stats(const float *lr, int h, int w, const float *flow, int ny, int nx, int ts, float *hr,
                            hhsr_stream_t stream);
/* noise curves (float64, n_curve entries, as the reference uploads them, super_resolution.py:98-99) -> float32 table
 * [n_curve][2] = (sigma_t^2, d_t^2) read by hhsr_robustness; build once per burst. */
int hhsr_noise_table(const double *std_curve, const double *diff_curve, int n_curve, float *table, hhsr_stream_t stream);
/* reference-side part of the noise model (robustness.py:504-533), once per burst: terms [4][H][W] =
 * (d_t^2 of channel 0, 1, 2 at the brightness level round(1000 * ref_mean), sum_c max(ref_var_c, sigma_t^2)).
 * ref_means/ref_vars: [3][H][W] from hhsr_upscale_warp_stats; noise_table from hhsr_noise_table. */
int hhsr_robustness_ref_terms(const float *ref_means, const float *ref_vars, int H, int W, const float *noise_table,
                              int n_curve, float *terms, hhsr_stream_t stream);
/* the three calls above in one pass for the reference frame (init_robustness, robustness.py:23-76, + the reference
 * part of the noise model): guide_means/guide_vars [3][h][w] from hhsr_guide_stats -> ref_means [3][2h][2w],
 * terms [4][2h][2w]; ref_vars [3][2h][2w] is only written when non-NULL. */
int hhsr_ref_stats_terms(const float *guide_means, const float *guide_vars, int h, int w, const float *noise_table,
                         int n_curve, float *ref_means, float *ref_vars, float *terms, hhsr_stream_t stream);
/* fused per-pixel robustness (robustness.py:421-639): warped Dodgson upsampling of the comp guide means,
 * |mean difference|, noise-model shrinkage, flow-irregularity factor S and threshold -> R [H][W].
 * ref_means: [3][H][W] from hhsr_upscale_warp_stats; ref_terms: [4][H][W] from hhsr_robustness_ref_terms.
 * flags: HHSR_ROBUSTNESS_GENERIC forces the per-pixel path where the block-uniform fast path applies (A/B parity tests). */
#define HHSR_ROBUSTNESS_GENERIC 1
int hhsr_robustness(const float *comp_means_lr, const float *ref_means, const float *ref_terms, int H, int W,
                    const float *flow, int ny, int nx, int ts, double t, double s1, double s2, double Mt, float *R,
                    int flags, hhsr_stream_t stream);
/* 5x5 edge-replicated local minimum (robustness.py:641-687); when acc_rob != NULL also acc_rob += r
 * (utils.py:93-120, float64 accumulator). */
int hhsr_local_min5(const float *R, int H, int W, float *r, double *acc_rob, hhsr_stream_t stream);

/* ---- RAW input (utils_dng.py:146-160): sensor counts uint16 [H][W] -> normalised float32
 * out = (float(raw) - black4[p]) / den4[p] * gain4[p], p = (row & 1) * 2 + (col & 1) the CFA position, every
 * operation rounded to float32 like the reference's NumPy code (den = white - black, gain = wb[c] / wb[1]). */
int hhsr_normalize_raw_u16(const unsigned short *raw, int H, int W, const float *black4, const float *den4,
                           const float *gain4, float *out, hhsr_stream_t stream);

/* ---- merge, Alg. 4 (merge.py:236-434): accumulate one aligned comp frame into num/den [Hs][Ws][3].
 * iso: 0 steerable kernel, 1 isotropic. */
int hhsr_merge_accumulate(const float *raw, int H, int W, const float *flow, int ny, int nx, int ts,
                          const float *covs, const float *r, float *num, float *den, int Hs, int Ws, double scale,
                          const int *cfa_host, int iso, hhsr_stream_t stream);
/* same arguments; num/den are INITIALISED with this frame's contribution instead of being updated (their previous
 * contents are ignored): the first comp frame of a burst needs no zero-filled accumulators (B200 addition). */
int hhsr_merge_init_accumulate(const float *raw, int H, int W, const float *flow, int ny, int nx, int ts,
                               const float *covs, const float *r, float *num, float *den, int Hs, int Ws,
                               double scale, const int *cfa_host, int iso, hhsr_stream_t stream);
/* Same arithmetic for K comp frames in ONE pass over the accumulators (B200 addition; frame order preserved per pixel,
 * bit-identical to K single-frame calls): the accumulators are read and written once per call instead of once per
 * frame.  raws/flows/covs/rs: HOST arrays of K device pointers.  flags: HHSR_MERGE_INIT — num/den are initialised by
 * the batch (previous contents ignored) instead of updated; HHSR_MERGE_GENERIC — run the any-scale kernel even where
 * the power-of-two fast path applies (A/B parity tests). */
#define HHSR_MERGE_INIT 1
#define HHSR_MERGE_GENERIC 2
int hhsr_merge_accumulate_batch(const float *const *raws, const float *const *flows, const float *const *covs,
                                const float *const *rs, int K, int H, int W, int ny, int nx, int ts, float *num,
                                float *den, int Hs, int Ws, double scale, const int *cfa_host, int iso, int flags,
                                hhsr_stream_t stream);
/* ---- merge of the reference frame, Alg. 11 (merge.py:22-233).  acc_rob (float64 [H][W]) may be NULL; when given,
 * the accumulated-robustness denoiser rules apply (widened window / overwrite).  fuse_divide != 0 additionally
 * performs utils.divide (num <- num/den) in the same pass.  Only output rows [row_begin, row_end) are processed
 * (0, Hs for the whole image; frame-sharded runs normalise one row slice per GPU). */
int hhsr_merge_ref(const float *raw, int H, int W, const float *covs, float *num, float *den, int Hs, int Ws,
                   double scale, const int *cfa_host, int iso, const double *acc_rob, int max_frame_count,
                   int rad_max, double max_multiplier, int fuse_divide, int row_begin, int row_end,
                   hhsr_stream_t stream);

/* ---- row-sharded merge across GPUs (a B200 addition, the reference is single-GPU).  Frames are
 * sharded over ranks up to the merge; each rank then merges ALL frames into its own slice of output rows and needs, from
 * every frame's owner, only the LR row band that slice can touch.
 * hhsr_gather_bands: for n_frames <= 24 frames, sources raws/rs/covs/flows (HOST arrays of DEVICE pointers, peer-mapped
 * over NVLink or local; covs or covs[f] may be NULL for the isotropic kernel) -> local full-size planes raws_local /
 * rs_local / covs_local and flow copies flows_local.  The band of frame f is derived ON THE DEVICE from the vertical
 * range of its flow over the tile rows covering LR rows [lr_begin, lr_end) and stored in extents[4 f .. 4 f + 3] =
 * (row_lo, row_hi, covs_row_lo, covs_row_hi) (device int array, 16-byte aligned); only those rows are copied.  A frame
 * whose source equals its destination (owned by this rank) is left in place. */
int hhsr_gather_bands(const float *const *raws, const float *const *rs, const float *const *covs,
                      const float *const *flows, float *const *raws_local, float *const *rs_local,
                      float *const *covs_local, float *const *flows_local, int n_frames, int H, int W, int ny, int nx, int ts,
                      int lr_begin, int lr_end, int *extents, hhsr_stream_t stream);
/* hhsr_merge_accumulate_batch restricted to output rows [row_begin, row_end): num_rows / den_rows point at row
 * `row_begin` (the caller may own only that slice: (row_end - row_begin) * Ws * 3 floats each). */
int hhsr_merge_accumulate_rows(const float *const *raws, const float *const *flows, const float *const *covs,
                               const float *const *rs, int K, int H, int W, int ny, int nx, int ts, float *num_rows,
                               float *den_rows, int Hs, int Ws, double scale, const int *cfa_host, int iso, int flags,
                               int row_begin, int row_end, hhsr_stream_t stream);
/* The LAST batch of a burst fused with merge_ref and divide (B200 addition): as hhsr_merge_accumulate_rows, then — in
 * the same pass, on the register accumulators — the reference frame's window sums (merge.py:82-233, plain mode: no
 * accumulated-robustness denoiser) and the quotient num / den (utils.py:62-90); only the finished image rows are
 * written to out_rows (pointing at row `row_begin`), num_rows / den_rows are read (unless HHSR_MERGE_INIT) but NOT
 * written.  Bit-identical to hhsr_merge_accumulate_rows + hhsr_merge_ref_rows(fuse_divide); 48 B per HR pixel less HBM
 * traffic and one launch less.  Needs the power-of-two fast path (scale 1, 2 or 4, Ws % 4 == 0): HHSR_E_UNSUPPORTED
 * otherwise.  With HHSR_MERGE_INIT num_rows / den_rows are only used as scratch when K > 24. */
int hhsr_merge_finish_rows(const float *const *raws, const float *const *flows, const float *const *covs,
                           const float *const *rs, int K, int H, int W, int ny, int nx, int ts, float *num_rows,
                           float *den_rows, int Hs, int Ws, double scale, const int *cfa_host, int iso, int flags,
                           int row_begin, int row_end, const float *ref_raw, const float *ref_covs, float *out_rows,
                           hhsr_stream_t stream);
/* hhsr_merge_ref on slice buffers: num_rows / den_rows point at row `row_begin`. */
int hhsr_merge_ref_rows(const float *raw, int H, int W, const float *covs, float *num_rows, float *den_rows, int Hs, int Ws,
                        double scale, const int *cfa_host, int iso, const double *acc_rob, int max_frame_count, int rad_max,
                        double max_multiplier, int fuse_divide, int row_begin, int row_end, hhsr_stream_t stream);

/* ---- frame-sharded runs (a B200 addition, the reference is single-GPU): the one reduction point of
 * the pipeline fused with merge_ref and divide.  peer_nums/peer_dens: HOST arrays of n_peers DEVICE pointers to the
 * ranks' private accumulators [Hs][Ws][3] (peer-mapped over NVLink, e.g. torch symmetric memory), summed in array
 * order; rows [row_begin, row_end) of  (sum_p num_p + ref) / (sum_p den_p + ref)  are written to out_num (which may
 * itself be a peer pointer, e.g. rank 0's image).  Remaining arguments as hhsr_merge_ref. */
int hhsr_reduce_merge_ref(const float *const *peer_nums, const float *const *peer_dens, int n_peers, const float *raw,
                          int H, int W, const float *covs, float *out_num, int Hs, int Ws, double scale,
                          const int *cfa_host, int iso, const double *acc_rob, int max_frame_count, int rad_max,
                          double max_multiplier, int fuse_divide, int row_begin, int row_end, hhsr_stream_t stream);

/* ---- output side: the host post-process of the reference on the device, so that only
 * the finished image crosses PCIe.  img: the merged image [H][W][3] float32 (H, W here are the OUTPUT sizes). */
/* colour matrix + clip (raw2rgb.py:139-146, 224-226): img[p] <- clip(ccm @ img[p], 0, 1) in place; ccm_host: 9 floats,
 * row-major cam2rgb. */
int hhsr_post_ccm_clip(float *img, size_t n_px, const float *ccm_host, hhsr_stream_t stream);
/* unsharp mask (raw2rgb.py:228-238 -> skimage.filters.unsharp_mask -> scipy.ndimage.gaussian_filter(sigma, truncate=4,
 * mode="reflect")), first pass: Gaussian along the rows axis (axis 0), float64 accumulation, float32 result in tmp.
 * taps_host: 2*radius+1 float64 weights (radius <= 64). */
int hhsr_post_blur_cols(const float *img, int H, int W, const double *taps_host, int radius, float *tmp,
                        hhsr_stream_t stream);
/* second pass (axis 1) fused with everything that follows (raw2rgb.py:228-250 and run_handheld.py:132-150):
 * img + (img - blurred) * amount when tmp != NULL (tmp from hhsr_post_blur_cols; NULL: sharpening off), devignetting
 * (raw2rgb.py:203-210) when devignette != 0, clip to [0,1], x ** inv_gamma when inv_gamma > 0, clip.
 * out_kind 0: float32 [H][W][3] (NaN kept, what process() returns); 1: uint8, 2: uint16 — nan_to_num, clip,
 * rint(x * 255 | 65535) (skimage img_as_ubyte / img_as_uint on a float32 image). */
int hhsr_post_finish(const float *img, const float *tmp, int H, int W, const double *taps_host, int radius, float amount,
                     int devignette, float inv_gamma, int out_kind, void *out, hhsr_stream_t stream);
/* frame-count-aware denoisers of the merged image (utils_image.py:174-231 gauss, :233-315 median): acc_rob is the
 * accumulated robustness float64 [H][W] (raw resolution), img/out [Hs][Ws][3] (out != img).  Upstream these stages
 * cannot run (see csrc/post.cu); mode/scale come from the main configuration, Gaussian window half-width = ceil(3 sigma),
 * median radius <= 7. */
int hhsr_frame_count_denoise_gauss(const float *img, int Hs, int Ws, const double *acc_rob, int H, int W, double scale,
                                   double sigma_max, double max_frame_count, float *out, hhsr_stream_t stream);
int hhsr_frame_count_denoise_median(const float *img, int Hs, int Ws, const double *acc_rob, int H, int W, double scale,
                                    double radius_max, double max_frame_count, float *out, hhsr_stream_t stream);

/* ---- noise curves (fast_monte_carlo.py:31-101 unitary_MC / regular_MC): for each of the n_levels
 * brightness values (device float64 array) the Monte-Carlo means over n_patches pairs of noisy clipped 3x3 patches,
 * diff_mean[l] = mean |mean(p1) - mean(p2)| and std_mean[l] = 0.5 mean(std(p1) + std(p2)) (device float64 arrays).
 * Counter-based Philox generator: the result is a function of (seed, n_patches, levels) only. */
int hhsr_noise_mc(const double *brightness, int n_levels, double alpha, double beta, int n_patches,
                  unsigned long long seed, double *diff_mean, double *std_mean, hhsr_stream_t stream);

/* ---- element-wise helpers (utils.py:62-120) */
int hhsr_divide(float *num, const float *den, size_t n, hhsr_stream_t stream);
int hhsr_add_f64_f32(double *A, const float *B, size_t n, hhsr_stream_t stream);
/* A += B_0 + ... + B_{K-1} in list order, one pass over A (Bs: HOST array of K device pointers). */
int hhsr_add_many_f64_f32(double *A, const float *const *Bs, int K, size_t n, hhsr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HHSR_H */
