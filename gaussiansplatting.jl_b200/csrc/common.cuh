// common.cuh — shared declarations of libgsrast (sm_100a).  Internal; the public surface is include/gsrast.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gsrast.h"

#define GSR_TILE 16          // BLOCK = (16,16)            GaussianSplatting.jl:55-56
#define GSR_TILE_PIXELS 256  // BLOCK_SIZE                 GaussianSplatting.jl:57

// Kernel-side camera + config (by value in the parameter space; R/t optionally re-read from device memory).
struct DevCamera {
    float R[9];  // column-major w2c rotation
    float t[3];
    float focal[2];
    float principal[2];
    float cam_center[3];
    const float *R_dev;
    const float *t_dev;
    int32_t width, height;
    int32_t grid_x, grid_y;
    float near_plane, far_plane, blur_eps;
    int32_t radius_clip;
};

// Per-Gaussian packed record streamed by the compositing kernels (written by preprocess):
//   float4 q0 = {mean2d.x, mean2d.y, conic.a, conic.b}
//   float4 q1 = {conic.c, opacity, f0, f1}
//   float4 q2 = {f2, f3, f4, f5}            (channels <= 6 -> 3 quads, 48 B)
//   float4 q3 = {f6, f7, 0, 0}              (channels == 8 -> 4 quads, 64 B)
__host__ __device__ constexpr int rec_quads(int channels) { return channels <= 6 ? 3 : 4; }
// Per-Gaussian gradient accumulator filled by the backward compositing kernel, moment form.  In 4-byte slots:
//   0-1   floats        S v_sigma*dx, S v_sigma*dy                       (-> v_mean2d = conic * (Sx, Sy))
//   2-7   three DOUBLES S v_sigma*dx^2, S v_sigma*dx*dy, S v_sigma*dy^2  (-> v_conic)
//   8     float         S e*v_alpha         9  flags word (multi-GPU: visible | clamped rgb)        10-11 pad
//   12-15 floats        v_f0 .. v_f3 (rgb, depth)           16-18 v_f5 .. v_f7 (normal; channels == 8 only), 19 pad
// The second moments feed the ill-conditioned part of the chain (the rotation / scale gradients of elongated Gaussians
// are differences of large terms): summed with fp32 atomics over the thousands of rows a large splat receives, their
// order-dependent rounding alone moved `vrot` by up to 1e-4 of its maximum between runs at C2 (60 runs: median 4e-5).
// fp64 REDs make those sums order-independent to ~1e-13; every per-row partial sum stays fp32.
__host__ __device__ constexpr int acc_floats(int channels) { return channels <= 6 ? 16 : 20; }
#define GSR_ACC_FLAGS_SLOT 9

struct AccRow {
    float sx, sy, sxx, sxy, syy, se;
    uint32_t flags;
    float f[8];  // per-channel feature cotangents; f[4] (the constant-1 alpha feature) is always 0
};
// EXCHANGE form of an accumulator row — what crosses NVLink in the fused multi-GPU backward: the same quantities as plain
// floats, 12 (channels <= 6) / 16 floats per row:  {Sx, Sy, Sxx, Sxy, Syy, Se, f0, f1, f2, f3, -, flags | f5, f6, f7, -}
// (flags in the last slot of the row).  The per-Gaussian chain converts the fp64 sums to fp32 anyway, so nothing is
// lost; the rows are a quarter smaller and the peers pull them with 3 instead of 4 128-bit loads per view and Gaussian
// (those remote loads set the fused kernel's time at 8 ranks: 0.56 ms with 64-byte rows against 0.38 ms with 48-byte ones).
__host__ __device__ constexpr int exchange_floats(int channels) { return channels <= 6 ? 12 : 16; }
__device__ __forceinline__ AccRow load_exchange_row(const float *row, const int channels) {
    AccRow r;
    const float4 a0 = *reinterpret_cast<const float4 *>(row);
    const float4 a1 = *reinterpret_cast<const float4 *>(row + 4);
    const float4 a2 = *reinterpret_cast<const float4 *>(row + 8);
    r.sx = a0.x; r.sy = a0.y; r.sxx = a0.z; r.sxy = a0.w; r.syy = a1.x; r.se = a1.y;
    r.f[0] = a1.z; r.f[1] = a1.w; r.f[2] = a2.x; r.f[3] = a2.y;
    r.f[4] = r.f[5] = r.f[6] = r.f[7] = 0.f;
    r.flags = __float_as_uint(a2.w);
    if (channels > 5) {
        const float4 a3 = *reinterpret_cast<const float4 *>(row + 12);
        r.f[5] = a3.x; r.f[6] = a3.y; r.f[7] = a3.z;
        r.flags = __float_as_uint(a3.w);
    }
    if (channels == 3) r.f[3] = 0.f;
    return r;
}

__device__ __forceinline__ AccRow load_acc_row(const float *acc, const int channels) {
    AccRow r;
    const float4 q0 = *reinterpret_cast<const float4 *>(acc);       // 128-bit loads (rows are 16-byte aligned)
    const double2 d1 = *reinterpret_cast<const double2 *>(acc + 4);
    const float4 q2 = *reinterpret_cast<const float4 *>(acc + 8);
    const float4 q3 = *reinterpret_cast<const float4 *>(acc + 12);
    r.sx = q0.x; r.sy = q0.y;
    r.sxx = (float)__hiloint2double(__float_as_int(q0.w), __float_as_int(q0.z));
    r.sxy = (float)d1.x; r.syy = (float)d1.y;
    r.se = q2.x;
    r.flags = __float_as_uint(q2.y);
    r.f[0] = q3.x; r.f[1] = q3.y; r.f[2] = q3.z; r.f[3] = q3.w;
    r.f[4] = r.f[5] = r.f[6] = r.f[7] = 0.f;
    if (channels > 5) {
        const float4 q4 = *reinterpret_cast<const float4 *>(acc + 16);
        r.f[5] = q4.x; r.f[6] = q4.y; r.f[7] = q4.z;
    }
    if (channels == 3) r.f[3] = 0.f;
    return r;
}

struct GeomPtrs {  // GeometryState (states.jl:2-47), SoA
    float *depths;
    float2 *means2d;
    float2 *grad_means2d;
    float *rgbs;            // [n,3]
    uint8_t *clamped;       // [n,3]
    int32_t *tiles_touched;
    int32_t *points_offset;
    float *conics;          // [n,3]
    int32_t *radii;
    float *normals;         // [n,3] or null
    float4 *rec;            // [n, rec_quads]
    float *gacc;            // [n, acc_floats]
};

// ---- launchers (one per translation unit) ---------------------------------------------------------------
// How the caller's parameter arrays are to be read (the activation pre-pass of the functor,
// rasterizer.jl:217-248, folded into the kernels).  All-zero = activated inputs, one (3,K,N) SH array.
struct ParamSpec {
    int raw_opacity = 0;            // opacities are pre-sigmoid (NU.sigmoid, rasterizer.jl:229)
    int raw_scale = 0;              // scales are log-scales (exp, rasterizer.jl:237)
    int isotropic = 0;              // scales is (1,N), broadcast to the three axes (rasterizer.jl:236,240-243)
    const float *sh_rest = nullptr; // non-null: `shs` is features_dc (3,1,N) and this features_rest (3,K-1,N) (hcat, :218)
    float *vsh_rest = nullptr;      // backward: cotangent of sh_rest
};
__device__ __forceinline__ float act_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

void launch_preprocess(const DevCamera &cam, int64_t n, int sh_degree, int K, int channels, const float *means,
                       const float *shs, const float *opac, const float *scales, const float *rots,
                       const GeomPtrs &g, cudaStream_t s, const ParamSpec &ps = ParamSpec());

// perm != nullptr: scan of tiles_touched[perm[j]] (emission in depth order)
void launch_scan_tiles(int64_t n, const int32_t *tiles_touched, const uint32_t *perm, int32_t *offsets,
                       uint32_t *scan_state, int64_t *total_dev, cudaStream_t s);
size_t scan_state_words(int64_t n);

struct SortPlan {
    int tile_bits, depth_bits, passes;
    uint32_t depth_base;
    int bit_lo;  // first bit of the compact key (tile << depth_bits | depth - depth_base) the passes sort on
    int small;   // 1: a sort over the N Gaussians (latency-bound partial wave): smaller pass tiles
};
SortPlan presort_plan(const SortPlan &plan);    // the Gaussians' own depth sort (flag bit + depth bits)
SortPlan tile_only_plan(const SortPlan &plan);  // the instance sort once emission is in depth order
void launch_presort_keys(int64_t n, const GeomPtrs &g, const SortPlan &plan, uint64_t *keys, uint32_t *vals, uint32_t *ghist,
                         cudaStream_t s);
// ghist != nullptr: also accumulate the radix-sort digit histograms [passes][256] (zeroed by sort_prepare)
// offsets: inclusive scan of tiles touched in EMISSION order; perm != nullptr: emission slot j is Gaussian perm[j]
void launch_duplicate(const DevCamera &cam, int64_t n, const GeomPtrs &g, const int32_t *offsets, const uint32_t *perm,
                      uint64_t *keys, uint32_t *vals, const SortPlan &plan, uint32_t *ghist, cudaStream_t s);
SortPlan make_sort_plan(int64_t n_tiles, float near_plane, float far_plane);
size_t sort_temp_words(int64_t m, const SortPlan &plan);
// keys_in/vals_in are left intact; result lands in keys_out/vals_out; (keys_tmp, vals_tmp) is scratch of size m.
uint32_t *sort_prepare(const SortPlan &plan, int64_t m, uint32_t *temp_words, cudaStream_t s);
// hist_ready: the digit histograms were already accumulated into sort_prepare()'s buffer (by launch_duplicate)
void launch_sort_pairs(const SortPlan &plan, int64_t m, const uint64_t *keys_in, const uint32_t *vals_in,
                       uint64_t *keys_out, uint32_t *vals_out, uint64_t *keys_tmp, uint32_t *vals_tmp,
                       uint32_t *temp_words, bool hist_ready, cudaStream_t s);

void launch_tile_ranges(int64_t m, const uint64_t *keys_sorted, uint32_t *ranges, cudaStream_t s);
// instance binning on bare 32-bit tile ids (emission in depth order: the depth pre-sort made the depth bits redundant)
void launch_duplicate_tiles(const DevCamera &cam, int64_t n, const GeomPtrs &g, const int32_t *offsets, const uint32_t *perm,
                            uint32_t *tiles, uint32_t *vals, const SortPlan &plan, uint32_t *ghist, cudaStream_t s);
void launch_sort_tiles(const SortPlan &plan, int64_t m, const uint32_t *tiles_in, const uint32_t *vals_in, uint32_t *tiles_out,
                       uint32_t *vals_out, uint32_t *tiles_tmp, uint32_t *vals_tmp, uint32_t *temp_words, cudaStream_t s);
void launch_tile_ranges32(int64_t m, const uint32_t *tiles_sorted, uint32_t *ranges, cudaStream_t s);
// order[0..n_tiles): the tiles sorted by falling instance count (19 % resolution), for the compositing kernels' CTA order
void launch_tile_order(int64_t n_tiles, const uint32_t *ranges, uint32_t *order, cudaStream_t s);
void launch_materialize_keys(int64_t m, const uint32_t *tiles_sorted, const uint32_t *vals_sorted, const float *depths,
                             uint64_t *keys_sorted, cudaStream_t s);

// both return 0, or -1 when no kernel was compiled for `math_mode`
// order: tile ids in the order the CTAs should take them (launch_tile_order), or nullptr for raster order
int launch_render_forward(int channels, int math_mode, int width, int height, const uint32_t *ranges, const uint32_t *order,
                          const uint32_t *vals_sorted, const float4 *rec, const float *bg, float *image,
                          uint32_t *n_contrib, float *accum_alpha, uint8_t *covis, float *uncert, cudaStream_t s);

int launch_render_backward(int channels, int math_mode, int width, int height, const uint32_t *ranges, const uint32_t *order,
                           const uint32_t *vals_sorted, const float4 *rec, const float *bg, const float *vpixels,
                           const uint32_t *n_contrib, const float *accum_alpha, float *gacc, cudaStream_t s);
int render_math_mode_supported(int math_mode);
int launch_exp_neg_probe(const float *sigma, float *split, float *libdev, float *inlined, int64_t n, cudaStream_t s);

void launch_backward_gaussians(const DevCamera &cam, int64_t n, int sh_degree, int K, int channels,
                               const float *means, const float *shs, const float *opac, const float *scales,
                               const float *rots,
                               const GeomPtrs &g, float *vmeans, float *vshs, float *vopac, float *vscales,
                               float *vrot, float *vR, float *vt, int accumulate, cudaStream_t s,
                               const ParamSpec &ps = ParamSpec());

void launch_update_stats(int64_t n, const int32_t *radii, const float2 *grad_means2d, uint32_t width,
                         uint32_t height, int32_t *max_radii, float *accum, float *denom, cudaStream_t s);

int launch_fp32_peak(cudaStream_t s, double *ms, double *flops);

// ---- peer-fused per-Gaussian backward (backward_peers.cu) --------------------------------------------------
#define GSR_MAX_PEERS 8    // ranks of one node
#define GSR_MAX_VIEWS 16   // views of one batch summed by gsr_backward_gaussians_views
struct PeerCamera {
    float R[9], t[3], focal[2], principal[2], cam_center[3];
    int32_t width, height;
    float blur_eps;
};
struct PeerArgs {
    PeerCamera cams[GSR_MAX_VIEWS];    // one camera per view of the batch
    const float *gacc[GSR_MAX_VIEWS];  // that view's accumulator [n][AF], in the memory of the rank that rendered it
    float *table[GSR_MAX_PEERS];       // every rank's gradient table: [vrot 4n | vmeans 3n | vscales 3n | vopac n | vshs 3Kn]
    int32_t n_views, world, rank;
    int32_t exchange_rows;  // 1: gacc[v] holds EXCHANGE rows (exchange_floats per row), 0: the handle's own accumulator layout
    int64_t n, lo, hi;
    int32_t sh_degree, K, channels, sh_stride;
    int32_t vsh_aligned;  // every table's SH segment (offset 11n floats) is 16-byte aligned
    const float *means, *shs, *opac, *scales, *rots;
};
void launch_pack_flags(int64_t n, int channels, const int32_t *radii, const uint8_t *clamped, float *gacc, cudaStream_t s);
// accumulator rows (+ visibility / clamp flags) -> exchange rows
void launch_export_rows(int64_t n, int channels, const float *gacc, float *rows, cudaStream_t s);
void launch_grad_means2d(int64_t n, int channels, const int32_t *radii, const float *conics, const float *gacc,
                         float2 *out, cudaStream_t s);
int launch_backward_gaussians_peers(const PeerArgs &args, cudaStream_t s);

void count_launch(int n = 1);

// ---- densification kernels (densify.cu) ----------------------------------------------------------------------
int launch_densify_masks(int64_t n, int64_t n_grad, const float *accum, const float *denom, const float *scales,
                         int isotropic, float grad_threshold, float gamma, uint8_t *clone_mask, uint8_t *split_mask,
                         cudaStream_t s);
int launch_prune_mask(int64_t n, const float *opacities, const float *scales, int isotropic, const int32_t *max_radii,
                      float min_opacity, int32_t max_screen_size, float gamma, uint8_t *valid, cudaStream_t s);
size_t mask_offsets_scratch_words(int64_t n);
int launch_mask_offsets(int64_t n, const uint8_t *mask, int32_t *offsets, int64_t *count_dev, int32_t *scratch,
                        cudaStream_t s);
int launch_gather_rows(int64_t n, int row_words, const void *src, const uint8_t *mask, const int32_t *offsets, void *dst,
                       int repeat, int64_t count, cudaStream_t s);
int launch_split_children(int64_t m, float *points, float *scales, int isotropic, const float *rotations,
                          const float *noise, int n_split, cudaStream_t s);

// ---- fused SSIM / photometric loss (ssim.cu) -----------------------------------------------------------------
int launch_ssim_forward(int W, int H, int CH, int B, const float *img, const float *ref, float C1, float C2, int train,
                        float *ssim_map, float *dm_dmu1, float *dm_dsigma1_sq, float *dm_dsigma12, cudaStream_t s);
int launch_ssim_backward(int W, int H, int CH, int B, const float *img, const float *ref, const float *dL_dmap,
                         const float *dm_dmu1, const float *dm_dsigma1_sq, const float *dm_dsigma12, float *dL_dimg,
                         cudaStream_t s);
int launch_photometric_loss(int W, int H, int C, const float *image, const float *target, float lambda, float C1,
                            float C2, float *scratch, double *acc, float *vpixels, float *loss, cudaStream_t s);

// ---- coalesced copy between a contiguous span of `nb` rows of `row` floats in global memory and per-thread padded
// rows in shared memory (row g at s + g*stride).  128-bit global accesses when `aligned16`; the (row, column)
// of each quad is advanced incrementally (two integer divisions per thread in total, none in the loop).
__device__ __forceinline__ void rows_global_to_shared(const float *__restrict__ src, float *s, const int nb,
                                                      const int row, const int stride, const int tid,
                                                      const int nthreads, const bool aligned16) {
    const int total = nb * row;
    const int nq = aligned16 ? (total >> 2) : 0;
    if (nq > 0) {
        const int step = 4 * nthreads, step_g = step / row, step_r = step - step_g * row;
        int e = 4 * tid, g = e / row, r = e - g * row;
        const bool whole = (row & 3) == 0;  // a quad never straddles two rows
        for (int q0 = tid; q0 < nq; q0 += 4 * nthreads) {  // 4 independent 128-bit loads in flight per thread
            float4 v4[4];
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (q0 + j * nthreads < nq) v4[j] = __ldg(reinterpret_cast<const float4 *>(src) + q0 + j * nthreads);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (q0 + j * nthreads < nq) {
                    const float4 v = v4[j];
                    float *d = s + g * stride + r;
                    if (whole) {
                        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
                    } else {
                        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int u = 0; u < 4; u++) d[r + u >= row ? u + stride - row : u] = vv[u];
                    }
                    r += step_r; g += step_g;
                    if (r >= row) { r -= row; g += 1; }
                }
            }
        }
    }
    for (int e = (nq << 2) + tid; e < total; e += nthreads) {
        const int g = e / row, r = e - g * row;
        s[g * stride + r] = __ldg(src + e);
    }
}
template <bool ACC>
__device__ __forceinline__ void rows_shared_to_global(float *__restrict__ dst, const float *s, const int nb,
                                                      const int row, const int stride, const int tid,
                                                      const int nthreads, const bool aligned16) {
    const int total = nb * row;
    const int nq = aligned16 ? (total >> 2) : 0;
    if (nq > 0) {
        const int step = 4 * nthreads, step_g = step / row, step_r = step - step_g * row;
        int e = 4 * tid, g = e / row, r = e - g * row;
        const bool whole = (row & 3) == 0;
        for (int q = tid; q < nq; q += nthreads) {
            const float *p = s + g * stride + r;
            float4 v;
            if (whole) {
                v = make_float4(p[0], p[1], p[2], p[3]);
            } else {
                float vv[4];
#pragma unroll
                for (int u = 0; u < 4; u++) vv[u] = p[r + u >= row ? u + stride - row : u];
                v = make_float4(vv[0], vv[1], vv[2], vv[3]);
            }
            float4 *d4 = reinterpret_cast<float4 *>(dst) + q;
            if (ACC) {
                const float4 o = *d4;
                v = make_float4(o.x + v.x, o.y + v.y, o.z + v.z, o.w + v.w);
            }
            *d4 = v;
            r += step_r; g += step_g;
            if (r >= row) { r -= row; g += 1; }
        }
    }
    for (int e = (nq << 2) + tid; e < total; e += nthreads) {
        const int g = e / row, r = e - g * row;
        if (ACC) dst[e] += s[g * stride + r]; else dst[e] = s[g * stride + r];
    }
}

// get_rect — utils.jl:18-29, fp32 op order preserved (callers compile with -fmad=false or use no FMA-able form).
__device__ __forceinline__ void get_rect(float px, float py, int32_t radius, int32_t gx, int32_t gy, int32_t &x0,
                                         int32_t &y0, int32_t &x1, int32_t &y1) {
    const float r = (float)radius;
    // (p - r) / 16 ; floor ; trunc ; clamp.  Division by 16 is exact scaling, written as the reference does.
    x0 = min(max(__float2int_rz(floorf(__fdiv_rn(__fsub_rn(px, r), 16.0f))), 0), gx);
    y0 = min(max(__float2int_rz(floorf(__fdiv_rn(__fsub_rn(py, r), 16.0f))), 0), gy);
    // gpu_cld(p + r, 16) = trunc(floor(((p + r) + 16 - 1) / 16))
    x1 = min(max(__float2int_rz(floorf(__fdiv_rn(__fsub_rn(__fadd_rn(__fadd_rn(px, r), 16.0f), 1.0f), 16.0f))), 0), gx);
    y1 = min(max(__float2int_rz(floorf(__fdiv_rn(__fsub_rn(__fadd_rn(__fadd_rn(py, r), 16.0f), 1.0f), 16.0f))), 0), gy);
}
