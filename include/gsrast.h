/*
 * gsrast.h — C ABI of libgsrast.so: the sm_100a rasterizer hot path of GaussianSplatting.jl.
 *
 * Drop-in boundary.  The reference has no FFI: its operator surface is the Julia pair
 *     rasterize(means_3d, shs, opacities, scales, rotations, R_w2c, t_w2c; rast, camera, sh_degree,
 *               background, covisibilities, uncertainties)          src/rasterization/rasterizer.jl:255-408
 *     ChainRulesCore.rrule(::typeof(rasterize), ...) -> ∇rasterize   src/rasterization/rasterizer.jl:416-573
 * plus the state other components read (rast.gstate.radii, rast.gstate.∇means_2d — src/strategy.jl:85-86)
 * and `_update_stats!` (src/strategy.jl:118-136).  ext/GaussianSplattingCUDAExt would `ccall` the entry
 * points below in place of the KernelAbstractions kernel launches (INTEGRATION.md shows the binding).
 * Widened per SURVEY.md §8(f): the activation-fused functor (gsr_forward_raw / gsr_backward_raw), the fused SSIM
 * operator and photometric loss (gsr_ssim_*, gsr_photometric_loss), the densification kernels (gsr_densify_masks ...
 * gsr_split_children), 3DGS .ply files (gsr_ply_*), and the multi-GPU
 * split of the backward (gsr_set_accumulator / gsr_backward_render / gsr_backward_gaussians_peers).
 *
 * Conventions
 *   - every pointer named *_dev / documented "device" is a CUDA device pointer owned by the caller;
 *   - arrays use the reference's memory layout unchanged (Julia column-major): means (3,N) = N packed
 *     float[3]; rotations (4,N) wxyz, 16-byte aligned; shs (3,K,N); image (C,W,H) = H rows × W × C;
 *   - every function returns 0 on success or a negative GsrStatus; the message is at gsr_last_error();
 *   - all work is enqueued on the caller's `stream` (a cudaStream_t passed as void*); a forward performs
 *     exactly one stream synchronisation (the 4-byte n_rendered read, as rasterizer.jl:337 does);
 *   - a handle is re-entrant but not thread-safe; handles are independent (two rasterizers coexist when
 *     a sky dome is used, src/sky_dome.jl:143-145);
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with GSR_ECUDA.
 */
#ifndef GSRAST_H
#define GSRAST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GSR_API __attribute__((visibility("default")))
#else
#define GSR_API
#endif

typedef enum {
    GSR_OK = 0,
    GSR_EINVAL = -1, /* bad argument (e.g. width % 16 != 0 — rasterizer.jl:66,281,435) */
    GSR_ECUDA = -2,  /* CUDA runtime / launch failure */
    GSR_ENOMEM = -3, /* workspace allocation failed */
    GSR_ESTATE = -4  /* backward without a matching forward on this handle */
} GsrStatus;

/* math_mode values */
#define GSR_MATH_REFERENCE 0 /* every compositing operation in the reference's op order: no FMA contraction, libdevice
                                expf, IEEE division (render.jl:90-106, 237-263).  The cross-check build.            */
#define GSR_MATH_FAST 1      /* log2-domain exponent, contracted FMAs, one ex2.approx.  Fastest; image error grows with
                                the conditioning of a pixel (elongated / near-opaque Gaussians), so it is NOT held to
                                the flat 1e-5.                                                                       */
#define GSR_MATH_STRICT 2    /* DEFAULT.  sigma bit-exact in the reference's order, alpha = min(0.99, o expf(-sigma)) with
                                libdevice's expf (what the reference's CUDA extension compiles exp to), depth-channel
                                sums in the reference's order; cheaper forms only where the outputs are insensitive.  Meets 1e-5 absolute (image, depth) / 1e-4 relative
                                (gradients) against the fp32 restatement of the reference.                           */
#define GSR_MATH_EXPERIMENT 256 /* + policy word: A/B builds only (-DGSR_POLICY_AB), not a supported surface        */

/* Constructor arguments of `GaussianRasterizer(kab; width, height, mode, near_plane, far_plane)`
 * (rasterizer.jl:60-90) plus the constants `rasterize` hard-codes (rasterizer.jl:294-295). */
typedef struct {
    int32_t width;       /* multiple of 16 */
    int32_t height;      /* multiple of 16 */
    int32_t channels;    /* 3 = :rgb, 5 = :rgbd, 8 = :rgbdn (rasterizer.jl:47-51) */
    float near_plane;    /* 0.2f */
    float far_plane;     /* 1000.f */
    int32_t radius_clip; /* 3 */
    float blur_eps;      /* 0.3f */
    int32_t math_mode;   /* GSR_MATH_* */
} GsrConfig;

/* The fields of `Camera` the path reads (src/camera.jl:2-45; rasterizer.jl:285-291, 310, 321). */
typedef struct {
    float R[9];           /* w2c[1:3,1:3], column-major */
    float t[3];           /* w2c[1:3,4] */
    float focal[2];
    float principal[2];   /* in [0,1] */
    float cam_center[3];  /* c2w[1:3,4] */
    const float *R_dev;   /* optional device (3,3) column-major / (3,) arrays: the positional R_w2c, t_w2c */
    const float *t_dev;   /* of `rasterize` (pose optimisation, projection.jl:71-75); NULL = use R, t    */
} GsrCamera;

/* Device views of the handle-owned state (GeometryState / BinningState / ImageState, states.jl).
 * Valid until the next gsr_forward / gsr_release_scene_buffers / gsr_destroy on the handle. */
typedef struct {
    int64_t n;                      /* Gaussians of the last forward */
    int64_t n_rendered;             /* tile instances M of the last forward */
    const int32_t *radii;           /* [n]   rast.gstate.radii */
    float *grad_means2d;            /* [n,2] rast.gstate.∇means_2d (pixel units), written by gsr_backward */
    const float *means2d;           /* [n,2] */
    const float *depths;            /* [n] */
    const float *conics;            /* [n,3] conic_opacities */
    const float *rgbs;              /* [n,3] */
    const uint8_t *clamped;         /* [n,3] */
    const int32_t *tiles_touched;   /* [n] */
    const int32_t *points_offset;   /* [n] inclusive scan */
    const float *normals;           /* [n,3] (channels == 8) or NULL */
    const uint64_t *keys_unsorted;  /* [M] */
    const uint32_t *values_unsorted;/* [M] 1-based ids */
    const uint64_t *keys_sorted;    /* [M] */
    const uint32_t *values_sorted;  /* [M] */
    const uint32_t *ranges;         /* [T,2] start (0-based), end (exclusive) */
    const uint32_t *n_contrib;      /* [H,W] */
    const float *accum_alpha;       /* [H,W] */
} GsrStateViews;

typedef struct GsrHandle GsrHandle;

GSR_API const char *gsr_version(void);
GSR_API const char *gsr_last_error(const GsrHandle *h); /* h may be NULL: last error of a failed gsr_create */

/* GaussianRasterizer(kab; ...) — rasterizer.jl:60-90 */
GSR_API int gsr_create(const GsrConfig *cfg, GsrHandle **out);
/* KA.unsafe_free!(rast) — rasterizer.jl:136-145 */
GSR_API int gsr_destroy(GsrHandle *h);
/* release_scene_buffers!(rast) — rasterizer.jl:111-123 */
GSR_API int gsr_release_scene_buffers(GsrHandle *h);
/* memory_usage(rast) — rasterizer.jl:127-134 (workspace bytes owned by the handle) */
GSR_API int gsr_memory_usage(const GsrHandle *h, size_t *bytes);
/* Views of the state of the last forward.  points_offset / keys_unsorted / values_unsorted — intermediates the hot path
 * does not need in the reference's layout (it scans and emits in depth order) — are written, in the reference's order,
 * by this call when they are first asked for (two launches + a stream synchronisation on the forward's stream). */
GSR_API int gsr_get_state(GsrHandle *h, GsrStateViews *views);
/* Number of forwards this handle has run.  gsr_backward differentiates the LAST forward (its per-pixel and binning
 * state live in the handle, as rast.{g,b,i}state do in the reference): a caller that interleaves forwards keeps the
 * value returned after its forward and compares before its backward (the Python mirror and the Julia shim raise). */
GSR_API int64_t gsr_forward_generation(const GsrHandle *h);

/* rasterize(...) — rasterizer.jl:255-408.
 *   n, K          Gaussians, stored SH coefficients per Gaussian ((max_sh_degree+1)^2); sh_degree in [0,3]
 *   means/scales  device (3,n); rotations device (4,n) 16-byte aligned; opacities device (1,n) activated;
 *                 shs device (3,K,n)
 *   background    host float[3]
 *   image_out     device (channels,W,H), fully overwritten (zero image when n_rendered == 0, rasterizer.jl:338)
 *   covis         optional device bool[n]  (render.jl:112)   uncert: optional device float (W,H) (render.jl:109,128)
 *   n_rendered    optional host out
 */
GSR_API int gsr_forward(GsrHandle *h, const GsrCamera *cam, int64_t n, int32_t sh_degree, int32_t K, const float *means,
                const float *shs, const float *opacities, const float *scales, const float *rotations,
                const float background[3], float *image_out, uint8_t *covis, float *uncert, int64_t *n_rendered,
                void *stream);

/* ∇rasterize(...) — rasterizer.jl:416-550, for the forward last run on this handle with the same inputs.
 *   vpixels       device (channels,W,H)
 *   vmeans (3,n), vshs (3,K,n), vopacities (1,n), vscales (3,n), vrot (4,n): device outputs, fully written
 *   (zeros for culled Gaussians) when accumulate == 0, added to when accumulate != 0 (view batches);
 *   vR (3,3 column-major) / vt (3): optional device outputs of the pose path (projection.jl:243-256),
 *   always accumulated into (caller zero-fills, as rasterizer.jl:500-501 does).
 */
GSR_API int gsr_backward(GsrHandle *h, const GsrCamera *cam, int64_t n, int32_t sh_degree, int32_t K, const float *means,
                 const float *shs, const float *opacities, const float *scales, const float *rotations,
                 const float background[3], const float *vpixels, float *vmeans, float *vshs, float *vopacities,
                 float *vscales, float *vrot, float *vR, float *vt, int32_t accumulate, void *stream);

/* ---- multi-GPU: per-Gaussian backward fused with the cross-GPU gradient reduction over peer memory -----------
 * (not in the reference, which is single-GPU; semantics = sum over the ranks' views of ∇rasterize)
 * gsr_set_accumulator: make the handle keep its per-Gaussian accumulator ([capacity][16 or 20] floats: 64 B per
 *   Gaussian for :rgb / :rgbd, 80 B for :rgbdn; second moments as doubles) in caller-provided memory — one accumulator
 *   per view of a batch.  A pointer swap when the capacity covers the current scene (the state of the last forward
 *   stays valid); NULL, 0 restores the private buffer.
 * gsr_export_accumulator: after gsr_backward_render, the accumulator (+ the view's visibility / clamp flags) as fp32
 *   EXCHANGE rows ([n][12 or 16] floats) — the form to publish in peer-mapped memory.
 * gsr_backward_render: first half of ∇rasterize — zero-fill + ∇render! into the accumulator (+ ∇means_2d).
 * gsr_backward_gaussians_peers: second half for `world` ranks at once.  Rank `rank` owns the Gaussian slice
 *   [rank*chunk, (rank+1)*chunk): it loads every rank's accumulator rows for that slice (peer_gacc[v], P2P loads),
 *   applies view v's ∇project / ∇spherical_harmonics (cams[v]), sums over v and stores the reduced rows into every
 *   rank's table (peer_tables[p], P2P stores; layout [vrot 4n | vmeans 3n | vscales 3n | vopacities n | vshs 3Kn]).
 *   The caller separates the two halves, and the kernel from the table's consumers, with cross-rank barriers. */
GSR_API int gsr_set_accumulator(GsrHandle *h, float *gacc_dev, int64_t capacity_gaussians);
GSR_API int gsr_backward_render(GsrHandle *h, int64_t n, const float background[3], const float *vpixels, void *stream);
GSR_API int gsr_backward_gaussians_peers(GsrHandle *h, int32_t world, int32_t rank, const GsrCamera *cams,
                                 const float *const *peer_gacc, float *const *peer_tables, int64_t n, int32_t sh_degree,
                                 int32_t K, const float *means, const float *shs, const float *opacities,
                                 const float *scales, const float *rotations, void *stream);
/* The same kernel for a batch of n_views (<= 16) views spread over `world` ranks in any way (several views per rank, or
 * all on one GPU with world == 1): view v was rendered with cams[v] and its accumulator is view_gacc[v] (one accumulator
 * per view: gsr_set_accumulator before that view's gsr_backward_render); rank `rank` reduces its slice of the Gaussians
 * over all n_views and stores the rows into the `world` tables.  On one GPU this replaces n_views accumulating
 * gsr_backward calls — one pass over the parameters and ONE write of the gradient table per batch instead of n_views
 * read-modify-write passes.  gsr_backward_gaussians_peers(world, ...) == this with n_views = world.
 * exchange_rows != 0: view_gacc[v] holds EXCHANGE rows written by gsr_export_accumulator ([n][12 or 16] plain floats) —
 * the form to put in peer-mapped memory: a quarter smaller than the handle's own rows (whose second moments are doubles),
 * and the peers' remote loads of these rows are what bounds the kernel at 8 ranks.  0: the handle's own accumulator layout.
 * peer_tables[p] == NULL for p != rank: rank p does not receive this rank's rows — with only its own pointer set, every
 * rank ends with the reduced rows of ITS slice only (reduce-scatter; a Gaussian-sharded optimizer needs no more). */
GSR_API int gsr_export_accumulator(GsrHandle *h, int64_t n, float *rows_dev, void *stream);
GSR_API int gsr_backward_gaussians_views(GsrHandle *h, int32_t n_views, const GsrCamera *cams, const float *const *view_gacc,
                                 int32_t exchange_rows, int32_t world, int32_t rank, float *const *peer_tables, int64_t n, int32_t sh_degree,
                                 int32_t K, const float *means, const float *shs, const float *opacities,
                                 const float *scales, const float *rotations, void *stream);

/* update_stats!(strategy, rast.gstate.radii, rast.gstate.∇means_2d, resolution) — strategy.jl:107-136 */
GSR_API int gsr_update_stats(GsrHandle *h, int64_t n, int32_t *max_radii, float *accum_grad_means2d, float *denom,
                     void *stream);

/* Host-buffer entry points (callers whose parameters live in host memory; bench.py's end-to-end number).
 * gsr_forward_backward_host: copies the five parameter arrays and vpixels from (pinned) host memory, runs forward +
 * backward, copies image and gradients back, returns when the host outputs are complete.  Any host output may be NULL.
 * gsr_forward_backward_host_async: same work, but returns once everything is enqueued (uploads on an internal
 * H2D stream, downloads on an internal D2H stream, two staging slots), so that consecutive submissions overlap
 * step k's compute and download with step k+1's upload; host outputs of a submission are valid after gsr_host_wait
 * (or after two further submissions).  Inputs must stay untouched until the call returns; n_rendered is final on return. */
GSR_API int gsr_forward_backward_host(GsrHandle *h, const GsrCamera *cam, int64_t n, int32_t sh_degree, int32_t K,
                              const float *means_h, const float *shs_h, const float *opacities_h,
                              const float *scales_h, const float *rotations_h, const float background[3],
                              const float *vpixels_h, float *image_h, float *vmeans_h, float *vshs_h,
                              float *vopacities_h, float *vscales_h, float *vrot_h, int64_t *n_rendered,
                              void *stream);
GSR_API int gsr_forward_backward_host_async(GsrHandle *h, const GsrCamera *cam, int64_t n, int32_t sh_degree, int32_t K,
                              const float *means_h, const float *shs_h, const float *opacities_h,
                              const float *scales_h, const float *rotations_h, const float background[3],
                              const float *vpixels_h, float *image_h, float *vmeans_h, float *vshs_h,
                              float *vopacities_h, float *vscales_h, float *vrot_h, int64_t *n_rendered,
                              void *stream);
GSR_API int gsr_host_wait(GsrHandle *h);
/* debug timeline of the last two async submissions (needs gsr_profile_enable): per slot {h2d start, h2d end,
 * compute start, compute end, d2h start, d2h end} in ms since the first submission */
GSR_API int gsr_host_timeline(GsrHandle *h, float out[12]);

/* Stand-alone stages (known-answer tests of the reference: runtests.jl:486-494; sort yardstick). */
/* identify_tile_range!(ranges, keys) — utils.jl:56-78; ranges_dev (2,T) must be pre-zeroed by the caller. */
GSR_API int gsr_identify_tile_range(const uint64_t *keys_dev, int64_t m, uint32_t *ranges_dev, void *stream);
/* stable ascending sort of (key,value) pairs on the `tile_bits + depth bits` that can differ
 * (sortperm! + 2×_permute!, rasterizer.jl:357-372).  depth bits are taken relative to bits(near) when near > 0. */
GSR_API int gsr_sort_pairs(GsrHandle *h, const uint64_t *keys_in_dev, const uint32_t *vals_in_dev, int64_t m,
                   uint64_t *keys_out_dev, uint32_t *vals_out_dev, void *stream);

/* ---- the activation pre-pass folded into the path ------------------------------------
 * The functor `rast(means_3d, opacities, scales, rotations, sh_color, sh_remainder; ...)` — rasterizer.jl:200-253 —
 * concatenates features_dc | features_rest into (3,K,N), applies NU.sigmoid to the opacities and exp to the
 * (3,N) or isotropic (1,N) log-scales, then calls `rasterize`; Zygote differentiates those broadcasts.  These two
 * entry points take the RAW parameters and return RAW-parameter cotangents: the activations are evaluated on load
 * inside the per-Gaussian kernels (sigmoid(x) = 1/(1+exp(-x)), exp = expf) and their pullbacks in the per-Gaussian
 * backward, so no activated copy, concatenated SH array or split/sliced gradient ever touches HBM.
 *   features_dc (3,1,N); features_rest (3,K-1,N) (NULL when K == 1); scales_raw (3,N), or (1,N) when isotropic != 0;
 *   vscales_raw has the shape of scales_raw.  Everything else as gsr_forward / gsr_backward. */
GSR_API int gsr_forward_raw(GsrHandle *h, const GsrCamera *cam, int64_t n, int32_t sh_degree, int32_t K,
                            const float *means_dev, const float *features_dc_dev, const float *features_rest_dev,
                            const float *opacities_raw_dev, const float *scales_raw_dev, int32_t isotropic,
                            const float *rotations_dev, const float background[3], float *image_dev,
                            uint8_t *covisibilities_dev, float *uncertainties_dev, int64_t *n_rendered, void *stream);
GSR_API int gsr_backward_raw(GsrHandle *h, const GsrCamera *cam, int64_t n, int32_t sh_degree, int32_t K,
                             const float *means_dev, const float *features_dc_dev, const float *features_rest_dev,
                             const float *opacities_raw_dev, const float *scales_raw_dev, int32_t isotropic,
                             const float *rotations_dev, const float background[3], const float *vpixels_dev,
                             float *vmeans_dev, float *vfeatures_dc_dev, float *vfeatures_rest_dev,
                             float *vopacities_raw_dev, float *vscales_raw_dev, float *vrot_dev, float *vR_dev,
                             float *vt_dev, int32_t accumulate, void *stream);

/* ---- the densification consumer of radii / ∇means_2d -----------------------------------
 * The device steps of `densify_and_prune!` (src/densification.jl) as stateless launches over plain device arrays; the
 * host sequence (clone, split, prune; Adam moments and statistics follow their parameters) is the caller's, as in the
 * reference (gsrast/densify.py mirrors it).  Arrays are the raw model arrays: scales (3,N) log-scales or (1,N) when
 * isotropic != 0, opacities pre-sigmoid.  Masks are N bytes (Bool).  Errors via gsr_last_error(NULL). */
/* ∇ = accum ./ denom with NaN -> 0 for the first n_grad Gaussians (0 for the rest: `padded_grad`, :74-75);
 * clone_mask = ∇ > threshold && max(exp(scales)) < gamma (:35-38); split_mask = ∇ >= threshold && max(exp(scales)) > gamma
 * (:78-81); gamma = extent * dense_percent.  Either mask may be NULL. */
GSR_API int gsr_densify_masks(int64_t n, int64_t n_grad, const float *accum_dev, const float *denom_dev,
                              const float *scales_dev, int32_t isotropic, float grad_threshold, float gamma,
                              uint8_t *clone_mask_dev, uint8_t *split_mask_dev, void *stream);
/* valid = sigmoid(opacity) > min_opacity [&& max_radii < max_screen_size && max(exp(scales)) < gamma when
 * max_screen_size > 0; gamma = 0.1 * pruning_extent] (:19-25). */
GSR_API int gsr_prune_mask(int64_t n, const float *opacities_dev, const float *scales_dev, int32_t isotropic,
                           const int32_t *max_radii_dev, float min_opacity, int32_t max_screen_size, float gamma,
                           uint8_t *valid_mask_dev, void *stream);
/* Exclusive prefix of a byte mask (the compaction index every array of the model shares) and the number of set
 * entries (*count_dev, int64 on the device).  scratch_dev: gsr_mask_offsets_scratch_words(n) int32 words. */
GSR_API size_t gsr_mask_offsets_scratch_words(int64_t n);
GSR_API int gsr_mask_offsets(int64_t n, const uint8_t *mask_dev, int32_t *offsets_dev, int64_t *count_dev,
                             int32_t *scratch_dev, void *stream);
/* dst[:, j + c*count] = src[:, i] for every i with mask[i], j = offsets[i], c < repeat: the reference's `x[:, mask]`,
 * `x[:, :, mask]` (prune_points!, densify_clone!, _prune_optimizer!) and `repeat(x[:, mask], 1, n_split)`
 * (densify_split!, :83-94).  A row is row_bytes (multiple of 4) contiguous bytes: parameters, Adam moments,
 * statistics and ids alike. */
GSR_API int gsr_gather_rows(int64_t n, int32_t row_bytes, const void *src_dev, const uint8_t *mask_dev,
                            const int32_t *offsets_dev, void *dst_dev, int32_t repeat, int64_t count, void *stream);
/* The children of a split, in place on the gathered + repeated arrays of m new Gaussians: stds = exp(scales);
 * points += R(q) * (stds .* noise) (`_add_split_noise!`, :123-136); scales = log(stds / (0.8 * n_split)) (:91).
 * noise (3,m): N(0,1) samples supplied by the caller (the reference draws them from the device RNG in the kernel). */
GSR_API int gsr_split_children(int64_t m, float *points_dev, float *scales_dev, int32_t isotropic,
                               const float *rotations_dev, const float *noise_dev, int32_t n_split, void *stream);

/* ---- the loss either side of the path ------------------------------------------------
 * Fused SSIM.  Arrays are the reference's (W,H,CH,B) column-major Float32, i.e. planar [b][c][y][x]; any W, H
 * (zero padding outside the image, 11-tap sigma=1.5 window).  No handle: stateless, errors via the return code
 * and gsr_last_error(NULL). */
/* `_fused_ssim(img; ref, C1, C2, train)` — fused_ssim.jl:354-372 (kernel :34-258).  Writes ssim_map and, when
 * train != 0, the three partial-derivative maps the pullback needs (pass NULL for them otherwise). */
GSR_API int gsr_ssim_forward(int32_t width, int32_t height, int32_t channels, int32_t batch, const float *img_dev,
                             const float *ref_dev, float C1, float C2, int32_t train, float *ssim_map_dev,
                             float *dm_dmu1_dev, float *dm_dsigma1_sq_dev, float *dm_dsigma12_dev, void *stream);
/* `fused_ssim_bwd(img, ref, dL_dmap, dm_dmu1, dm_dsigma1_sq, dm_dsigma12)` — fused_ssim.jl:374-389 (kernel
 * :261-352), the pullback of the rrule at :397-407.  Every element of dL_dimg is written. */
GSR_API int gsr_ssim_backward(int32_t width, int32_t height, int32_t channels, int32_t batch, const float *img_dev,
                              const float *ref_dev, const float *dL_dmap_dev, const float *dm_dmu1_dev,
                              const float *dm_dsigma1_sq_dev, const float *dm_dsigma12_dev, float *dL_dimg_dev,
                              void *stream);
/* The photometric loss of Trainer.step! (training.jl:684-699) and its pullback to the raster image in one call:
 *   x = image[1:3,:,:] ; total = (1-lambda)*mean|x - target| + lambda*(1 - mean(fused_ssim(x; ref=target)))
 * image_dev / vpixels_dev: the handle's (C,W,H) raster image / its cotangent (channels >= 3 get zeros), consumed
 * and produced in place of the slice / permutedims / reshape passes and their pullbacks; target_dev: (W,H,3).
 * loss_dev: 3 floats {total, mean|x - t|, mean ssim} written on the stream (read them when convenient).
 * vpixels_dev feeds gsr_backward directly.  Scratch (three W*H*3 maps) lives in the handle. */
GSR_API int gsr_photometric_loss(GsrHandle *h, const float *image_dev, const float *target_dev, float lambda_dssim,
                                 float *vpixels_dev, float *loss_dev, void *stream);

/* ---- getting real scenes in and out (host side, no GPU involved) ------------------------
 * 3DGS `.ply` files as read / written by import_ply / export_ply (gaussians.jl:157-247, via PlyIO.jl there).
 * Properties are matched by NAME (order and storage type free; ascii, little- and big-endian binary); f_rest_* is
 * channel-major in the file and (3,R,N) in memory.  Arrays are HOST buffers in the reference's model layout and hold
 * the RAW parameters gsr_forward_raw takes: points (3,N), features_dc (3,1,N), features_rest (3,R,N), opacities (1,N)
 * pre-sigmoid, scales (3,N) log, rotations (4,N) wxyz.  Errors: GSR_EINVAL + gsr_ply_last_error(). */
GSR_API int gsr_ply_open(const char *path, int64_t *n, int32_t *n_rest_coeffs /* R = K-1 per channel */, void **reader);
GSR_API int gsr_ply_read(void *reader, float *points, float *features_dc, float *features_rest, float *opacities,
                         float *scales, float *rotations);
GSR_API void gsr_ply_close(void *reader);
GSR_API int gsr_ply_write(const char *path, int64_t n, int32_t n_rest_coeffs, const float *points,
                          const float *features_dc, const float *features_rest, const float *opacities,
                          const float *scales, const float *rotations);
/* Same, with the row count of `scales` stated: 3 = (3,N); 1 = an isotropic model's (1,N), for which export_ply emits
 * only `scale_0` (gaussians.jl:176 iterates axes(scales,1)).  gsr_ply_write is the n_scale_rows = 3 case. */
GSR_API int gsr_ply_write_scales(const char *path, int64_t n, int32_t n_rest_coeffs, int32_t n_scale_rows,
                                 const float *points, const float *features_dc, const float *features_rest,
                                 const float *opacities, const float *scales, const float *rotations);
GSR_API const char *gsr_ply_last_error(void);

/* Per-stage device timing (CUDA events on the caller's stream; SURVEY.md §5 "tracing / profiling").
 * When enabled, gsr_forward / gsr_backward bracket each stage with events; gsr_profile_get synchronises on
 * them and returns the durations of the last forward + backward in milliseconds (0 for stages that did not run). */
enum {
    GSR_STAGE_PREPROCESS = 0, /* project! + spherical_harmonics! + count_tiles + feature packing */
    GSR_STAGE_SCAN,           /* cumsum! + n_rendered read-back */
    GSR_STAGE_DUPLICATE,      /* duplicate_with_keys! */
    GSR_STAGE_SORT,           /* sortperm! + 2 x _permute! */
    GSR_STAGE_RANGES,         /* fill!(ranges) + identify_tile_range! */
    GSR_STAGE_RENDER_FWD,     /* render! */
    GSR_STAGE_ZERO_GRADS,     /* zero-fill of the per-Gaussian accumulators */
    GSR_STAGE_RENDER_BWD,     /* ∇render! */
    GSR_STAGE_GAUSS_BWD,      /* ∇project! + ∇spherical_harmonics! */
    GSR_STAGE_PRESORT,        /* depth sort of the Gaussians ahead of cumsum! / duplicate_with_keys! (part of the sort) */
    GSR_NUM_STAGES
};
GSR_API int gsr_profile_enable(GsrHandle *h, int32_t enable);
GSR_API int gsr_profile_get(GsrHandle *h, float ms[GSR_NUM_STAGES]);

/* FP32 FMA micro-benchmark on the current device (dependent FFMA chains, all SMs): the measured FP32 peak
 * that the compositing kernels' roofline fraction is quoted against. */
GSR_API int gsr_measure_fp32_peak(double *tflops, void *stream);

/* Test hook: exp(-sigma) three ways, element-wise over n device floats — the A/B split policy (one ex2.approx after a
 * Cody-Waite split), libdevice's expf, and the inlined instruction sequence GSR_MATH_STRICT uses (may be NULL), which
 * tests/ require to equal libdevice's expf bit for bit. */
GSR_API int gsr_debug_exp_neg(const float *sigma_dev, float *split_dev, float *libdevice_dev, float *inlined_dev, int64_t n,
                              void *stream);

/* Environment knobs, read once (per handle / per process).  Measurement aids, not part of the supported surface; every
 * setting produces the same results bit for bit.
 *   GSR_PRESORT=0          one 5-pass radix sort of the 64-bit instance keys instead of the depth pre-sort of the Gaussians +
 *                          the tile-digit sort of 32-bit instance keys (default 1)
 *   GSR_TILE_ORDER=1       the compositing kernels take their tiles heaviest first (default 0: raster order)
 *   GSR_SORT_IPT=8|16, GSR_PRESORT_IPT=8|16    keys per thread of the radix passes over the instances / the Gaussians
 *   GSR_SORT_RANK=match|ballot|auto            ranking primitive of the radix passes
 *   GSR_DUP_CTAS_PER_SM=n  CTAs per SM of the cooperative duplicate kernel (default 4) */

/* Kernels launched by this library since process start (bench.py's gpu_launches). */
GSR_API int64_t gsr_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* GSRAST_H */
