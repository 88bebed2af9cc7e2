// gsrast.cu — C ABI (include/gsrast.h), handle / workspace lifecycle and stage orchestration.
//
// Mirrors the control flow of `rasterize` (src/rasterization/rasterizer.jl:255-408) and `∇rasterize`
// (:416-550): the handle plays the role of `rast.{g,b,i}state` (states.jl), grown monotonically like
// rasterizer.jl:275-278,340-343 and released by gsr_release_scene_buffers (rasterizer.jl:111-123).
#include <atomic>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges are no-ops unless a profiler injects itself

#include "common.cuh"

namespace {

std::atomic<int64_t> g_launches{0};
const char *const kStageNames[GSR_NUM_STAGES] = {"gsr:preprocess", "gsr:scan",       "gsr:duplicate",
                                                 "gsr:sort",       "gsr:ranges",     "gsr:render_fwd",
                                                 "gsr:zero_grads", "gsr:render_bwd", "gsr:gauss_bwd",
                                                 "gsr:presort"};
thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

}  // namespace

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

struct GsrHandle {
    GsrConfig cfg;
    int grid_x = 0, grid_y = 0;
    int64_t n_tiles = 0;
    SortPlan plan;
    std::string err;
    size_t bytes = 0;  // device bytes owned

    // GeometryState
    int64_t cap_n = 0;
    GeomPtrs g{};
    float *gacc_external = nullptr;  // caller-provided accumulator storage (symmetric / peer-mapped memory)
    int64_t gacc_external_cap = 0;
    // BinningState
    int64_t cap_m = 0;
    uint64_t *keys_unsorted = nullptr, *keys_sorted = nullptr, *keys_tmp = nullptr;
    uint32_t *vals_unsorted = nullptr, *vals_sorted = nullptr, *vals_tmp = nullptr;
    uint32_t *sort_temp = nullptr;
    size_t sort_temp_cap = 0;  // words
    // ImageState
    uint32_t *ranges = nullptr, *n_contrib = nullptr;
    uint32_t *tile_order = nullptr;  // tiles by falling instance count (CTA order of the compositing kernels)
    bool use_tile_order = false;     // GSR_TILE_ORDER=1: heaviest tiles first (pays on clustered scenes)
    float *accum_alpha = nullptr;
    // depth pre-sort of the Gaussians (binning.cu): ping-pong pairs, the permutation, tiles touched scanned in that order
    bool presort = true;               // GSR_PRESORT=0 restores the single 5-pass instance sort (A/B measurements)
    SortPlan pplan, tplan;             // the Gaussians' sort / the tile-digit instance sort that follows it
    uint64_t *pk[3] = {nullptr, nullptr, nullptr};
    uint32_t *pv[3] = {nullptr, nullptr, nullptr};
    uint32_t *psort_temp = nullptr;
    size_t psort_temp_cap = 0;
    int32_t *offsets_sorted = nullptr;
    bool ref_binning_valid = false;    // points_offset / keys_unsorted / values_unsorted hold the reference's content
    cudaStream_t last_stream = nullptr;
    DevCamera last_cam{};
    // scan
    uint32_t *scan_state = nullptr;
    size_t scan_cap = 0;  // 64-bit words
    int64_t *total_dev = nullptr;   // device alias of total_host (zero-copy)
    int64_t *total_host = nullptr;  // pinned + mapped: the scan kernel stores n_rendered straight into host memory, so
                                    // the read-back never queues behind bulk D2H traffic on the copy engine

    // state of the last forward
    int64_t last_n = 0, last_m = 0;
    bool fwd_valid = false;
    int64_t generation = 0;  // bumped by every forward: a backward belongs to exactly one of them

    // per-stage timing
    bool profile = false;
    cudaEvent_t ev[2 * GSR_NUM_STAGES] = {};
    bool ev_used[GSR_NUM_STAGES] = {};

    // staging for the host-buffer entry points: two slots so that step k+1's H2D overlaps step k's compute / D2H
    struct HostSlot {
        DevBuf means, shs, opac, scales, rots, vpix, image, vmeans, vshs, vopac, vscales, vrot;
        cudaEvent_t h2d_done = nullptr, compute_done = nullptr, d2h_done = nullptr;
        cudaEvent_t t_h2d0 = nullptr, t_h2d1 = nullptr, t_c0 = nullptr, t_c1 = nullptr, t_d0 = nullptr, t_d1 = nullptr;  // timeline
        bool d2h_pending = false;
    } slot[2];
    int next_slot = 0;
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t t_base = nullptr;

    // photometric loss scratch: three (W,H,3) derivative maps + two double accumulators
    float *loss_maps = nullptr;
    double *loss_acc = nullptr;
};

namespace {

int fail(GsrHandle *h, int code, const std::string &msg) {
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}
int cuda_fail(GsrHandle *h, cudaError_t e, const char *what) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    return fail(h, e == cudaErrorMemoryAllocation ? GSR_ENOMEM : GSR_ECUDA, buf);
}
#define CK(call)                                                    \
    do {                                                            \
        cudaError_t e_ = (call);                                    \
        if (e_ != cudaSuccess) return cuda_fail(h, e_, #call);      \
    } while (0)

template <typename T>
cudaError_t dev_alloc(GsrHandle *h, T **p, size_t count) {
    *p = nullptr;
    if (count == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(p), count * sizeof(T));
    if (e == cudaSuccess) h->bytes += count * sizeof(T);
    return e;
}
template <typename T>
void dev_free(GsrHandle *h, T *&p, size_t count) {
    if (p) {
        cudaFree(p);
        h->bytes -= count * sizeof(T);
        p = nullptr;
    }
}

void free_geometry(GsrHandle *h) {
    const size_t n = (size_t)h->cap_n;
    const int ch = h->cfg.channels;
    dev_free(h, h->g.depths, n);
    dev_free(h, h->g.means2d, n);
    dev_free(h, h->g.grad_means2d, n);
    dev_free(h, h->g.rgbs, 3 * n);
    dev_free(h, h->g.clamped, 3 * n);
    dev_free(h, h->g.tiles_touched, n);
    dev_free(h, h->g.points_offset, n);
    dev_free(h, h->g.conics, 3 * n);
    dev_free(h, h->g.radii, n);
    if (ch > 5) dev_free(h, h->g.normals, 3 * n);
    dev_free(h, h->g.rec, (size_t)rec_quads(ch) * n);
    if (h->gacc_external) h->g.gacc = nullptr;
    else dev_free(h, h->g.gacc, (size_t)acc_floats(ch) * n);
    dev_free(h, h->scan_state, 2 * h->scan_cap);
    h->scan_cap = 0;
    for (int k = 0; k < 3; k++) {
        dev_free(h, h->pk[k], n);
        dev_free(h, h->pv[k], n);
    }
    dev_free(h, h->psort_temp, h->psort_temp_cap);
    h->psort_temp_cap = 0;
    dev_free(h, h->offsets_sorted, n);
    h->cap_n = 0;
}

int ensure_geometry_alloc(GsrHandle *h, int64_t n, cudaStream_t s);

// Grow-only GeometryState.  A failed growth releases whatever it had allocated (the byte accounting of
// gsr_memory_usage stays exact) and leaves the handle without geometry buffers.
int ensure_geometry(GsrHandle *h, int64_t n, cudaStream_t s) {
    if (n <= h->cap_n) return GSR_OK;
    free_geometry(h);  // rasterizer.jl:275-278: a larger scene replaces the whole GeometryState
    if (h->gacc_external && h->gacc_external_cap < n)
        return fail(h, GSR_EINVAL, "external accumulator smaller than the scene");
    h->cap_n = n;  // free_geometry() releases by capacity: set it before the first allocation
    const int rc = ensure_geometry_alloc(h, n, s);
    if (rc != GSR_OK) {
        const std::string msg = h->err;
        free_geometry(h);
        h->err = msg;
    }
    return rc;
}

int ensure_geometry_alloc(GsrHandle *h, int64_t n, cudaStream_t s) {
    const size_t c = (size_t)n;
    const int ch = h->cfg.channels;
    CK(dev_alloc(h, &h->g.depths, c));
    CK(dev_alloc(h, &h->g.means2d, c));
    CK(dev_alloc(h, &h->g.grad_means2d, c));
    CK(dev_alloc(h, &h->g.rgbs, 3 * c));
    CK(dev_alloc(h, &h->g.clamped, 3 * c));
    CK(dev_alloc(h, &h->g.tiles_touched, c));
    CK(dev_alloc(h, &h->g.points_offset, c));
    CK(dev_alloc(h, &h->g.conics, 3 * c));
    CK(dev_alloc(h, &h->g.radii, c));
    if (ch > 5) CK(dev_alloc(h, &h->g.normals, 3 * c));
    CK(dev_alloc(h, &h->g.rec, (size_t)rec_quads(ch) * c));
    if (h->gacc_external) {
        h->g.gacc = h->gacc_external;
    } else {
        CK(dev_alloc(h, &h->g.gacc, (size_t)acc_floats(ch) * c));
    }
    h->scan_cap = scan_state_words(n);
    CK(dev_alloc(h, &h->scan_state, 2 * h->scan_cap));
    if (h->presort) {
        for (int k = 0; k < 3; k++) {
            CK(dev_alloc(h, &h->pk[k], c));
            CK(dev_alloc(h, &h->pv[k], c));
        }
        h->psort_temp_cap = sort_temp_words(n, h->pplan);
        CK(dev_alloc(h, &h->psort_temp, h->psort_temp_cap));
        CK(dev_alloc(h, &h->offsets_sorted, c));
    }
    // KA.zeros in the reference (states.jl:30-47): stale-state reads of never-visible rows see zeros.  Enqueued on
    // the caller's stream: a legacy-stream cudaMemset is not ordered against a cudaStreamNonBlocking stream (torch
    // side streams) and could land after this forward's preprocess kernel.
    CK(cudaMemsetAsync(h->g.depths, 0, c * 4, s));
    CK(cudaMemsetAsync(h->g.means2d, 0, c * 8, s));
    CK(cudaMemsetAsync(h->g.grad_means2d, 0, c * 8, s));
    CK(cudaMemsetAsync(h->g.rgbs, 0, c * 12, s));
    CK(cudaMemsetAsync(h->g.clamped, 0, c * 3, s));
    CK(cudaMemsetAsync(h->g.conics, 0, c * 12, s));
    CK(cudaMemsetAsync(h->g.radii, 0, c * 4, s));
    if (ch > 5) CK(cudaMemsetAsync(h->g.normals, 0, c * 12, s));
    return GSR_OK;
}

void free_binning(GsrHandle *h) {
    const size_t m = (size_t)h->cap_m;
    dev_free(h, h->keys_unsorted, m);
    dev_free(h, h->keys_sorted, m);
    dev_free(h, h->keys_tmp, m);
    dev_free(h, h->vals_unsorted, m);
    dev_free(h, h->vals_sorted, m);
    dev_free(h, h->vals_tmp, m);
    dev_free(h, h->sort_temp, h->sort_temp_cap);
    h->sort_temp_cap = 0;
    h->cap_m = 0;
}

int ensure_binning(GsrHandle *h, int64_t m) {
    if (m <= h->cap_m) return GSR_OK;
    free_binning(h);  // rasterizer.jl:340-343
    const int64_t cap = m + m / 4 + 1024;  // slack: M drifts upward during training
    const size_t c = (size_t)cap;
    CK(dev_alloc(h, &h->keys_unsorted, c));
    CK(dev_alloc(h, &h->keys_sorted, c));
    CK(dev_alloc(h, &h->keys_tmp, c));
    CK(dev_alloc(h, &h->vals_unsorted, c));
    CK(dev_alloc(h, &h->vals_sorted, c));
    CK(dev_alloc(h, &h->vals_tmp, c));
    h->sort_temp_cap = sort_temp_words(cap, h->plan);
    CK(dev_alloc(h, &h->sort_temp, h->sort_temp_cap));
    h->cap_m = cap;
    return GSR_OK;
}

void make_dev_camera(const GsrHandle *h, const GsrCamera *cam, DevCamera *d) {
    memcpy(d->R, cam->R, sizeof d->R);
    memcpy(d->t, cam->t, sizeof d->t);
    memcpy(d->focal, cam->focal, sizeof d->focal);
    memcpy(d->principal, cam->principal, sizeof d->principal);
    memcpy(d->cam_center, cam->cam_center, sizeof d->cam_center);
    d->R_dev = (cam->R_dev && cam->t_dev) ? cam->R_dev : nullptr;
    d->t_dev = (cam->R_dev && cam->t_dev) ? cam->t_dev : nullptr;
    d->width = h->cfg.width;
    d->height = h->cfg.height;
    d->grid_x = h->grid_x;
    d->grid_y = h->grid_y;
    d->near_plane = h->cfg.near_plane;
    d->far_plane = h->cfg.far_plane;
    d->blur_eps = h->cfg.blur_eps;
    d->radius_clip = h->cfg.radius_clip;
}

struct StageTimer {  // records an event pair around a stage when profiling is on
    GsrHandle *h;
    cudaStream_t s;
    int stage;
    StageTimer(GsrHandle *h_, cudaStream_t s_, int stage_) : h(h_), s(s_), stage(stage_) {
        nvtxRangePushA(kStageNames[stage]);  // one NVTX range per stage (host-side launch span)
        if (h->profile) cudaEventRecord(h->ev[2 * stage], s);
    }
    ~StageTimer() {
        nvtxRangePop();
        if (h->profile) {
            cudaEventRecord(h->ev[2 * stage + 1], s);
            h->ev_used[stage] = true;
        }
    }
};

int ensure_stage(GsrHandle *h, DevBuf &b, size_t bytes) {
    if (bytes <= b.bytes) return GSR_OK;
    if (b.p) { cudaFree(b.p); h->bytes -= b.bytes; b.p = nullptr; b.bytes = 0; }
    CK(cudaMalloc(&b.p, bytes));
    b.bytes = bytes;
    h->bytes += bytes;
    return GSR_OK;
}

}  // namespace

extern "C" {

const char *gsr_version(void) { return "gsrast 0.1.0 (sm_100a)"; }

const char *gsr_last_error(const GsrHandle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int64_t gsr_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int64_t gsr_forward_generation(const GsrHandle *h) { return h ? h->generation : -1; }

int gsr_create(const GsrConfig *cfg, GsrHandle **out) {
    GsrHandle *h = nullptr;
    if (!cfg || !out) return fail(h, GSR_EINVAL, "gsr_create: null argument");
    *out = nullptr;
    if (cfg->width <= 0 || cfg->height <= 0 || cfg->width % 16 != 0 || cfg->height % 16 != 0)
        return fail(h, GSR_EINVAL, "width and height must be positive multiples of 16 (rasterizer.jl:66)");
    if (cfg->channels != 3 && cfg->channels != 5 && cfg->channels != 8)
        return fail(h, GSR_EINVAL, "Invalid render mode: channels must be 3 (:rgb), 5 (:rgbd) or 8 (:rgbdn)");
    if (!render_math_mode_supported(cfg->math_mode))
        return fail(h, GSR_EINVAL, "math_mode must be GSR_MATH_STRICT, GSR_MATH_REFERENCE or GSR_MATH_FAST");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(h, GSR_ECUDA, "no CUDA device available: libgsrast has no CPU fallback");
    }
    h = new GsrHandle();
    h->cfg = *cfg;
    h->grid_x = cfg->width / GSR_TILE;
    h->grid_y = cfg->height / GSR_TILE;
    h->n_tiles = (int64_t)h->grid_x * h->grid_y;
    h->plan = make_sort_plan(h->n_tiles, cfg->near_plane, cfg->far_plane);
    h->pplan = presort_plan(h->plan);
    h->tplan = tile_only_plan(h->plan);
    {
        const char *e = getenv("GSR_TILE_ORDER");
        h->use_tile_order = e && atoi(e) != 0;
    }
    {
        const char *e = getenv("GSR_PRESORT");
        // (the cooperative duplicate packs a rectangle's origin and width into 10 + 12 + 10 bits)
        h->presort = !(e && atoi(e) == 0) && h->grid_x <= 1023 && h->grid_y <= 4095;
    }
    const size_t px = (size_t)cfg->width * cfg->height;
    int rc = GSR_OK;
    do {
        if ((e = dev_alloc(h, &h->ranges, 2 * (size_t)h->n_tiles)) != cudaSuccess) break;
        if ((e = dev_alloc(h, &h->tile_order, (size_t)h->n_tiles)) != cudaSuccess) break;
        if ((e = dev_alloc(h, &h->n_contrib, px)) != cudaSuccess) break;
        if ((e = dev_alloc(h, &h->accum_alpha, px)) != cudaSuccess) break;
        if ((e = cudaHostAlloc(reinterpret_cast<void **>(&h->total_host), sizeof(int64_t), cudaHostAllocMapped)) != cudaSuccess) break;
        if ((e = cudaHostGetDevicePointer(reinterpret_cast<void **>(&h->total_dev), h->total_host, 0)) != cudaSuccess) break;
        if ((e = cudaMemset(h->ranges, 0, 2 * (size_t)h->n_tiles * 4)) != cudaSuccess) break;
        if ((e = cudaMemset(h->n_contrib, 0, px * 4)) != cudaSuccess) break;
        if ((e = cudaMemset(h->accum_alpha, 0, px * 4)) != cudaSuccess) break;
        // the fills ran on the legacy stream; later work arrives on caller streams that may be non-blocking
        if ((e = cudaDeviceSynchronize()) != cudaSuccess) break;
    } while (0);
    if (e != cudaSuccess) {
        rc = cuda_fail(nullptr, e, "gsr_create allocation");
        gsr_destroy(h);
        return rc;
    }
    *out = h;
    return GSR_OK;
}

int gsr_release_scene_buffers(GsrHandle *h) {
    if (!h) return GSR_EINVAL;
    free_geometry(h);
    free_binning(h);
    for (auto &sl : h->slot) {
        if (sl.d2h_pending && sl.d2h_done) cudaEventSynchronize(sl.d2h_done);
        sl.d2h_pending = false;
        DevBuf *st[] = {&sl.means, &sl.shs, &sl.opac, &sl.scales, &sl.rots, &sl.vpix,
                        &sl.image, &sl.vmeans, &sl.vshs, &sl.vopac, &sl.vscales, &sl.vrot};
        for (DevBuf *b : st)
            if (b->p) { cudaFree(b->p); h->bytes -= b->bytes; b->p = nullptr; b->bytes = 0; }
    }
    h->fwd_valid = false;
    h->last_n = h->last_m = 0;
    return GSR_OK;
}

int gsr_destroy(GsrHandle *h) {
    if (!h) return GSR_OK;
    gsr_release_scene_buffers(h);
    const size_t px = (size_t)h->cfg.width * h->cfg.height;
    dev_free(h, h->ranges, 2 * (size_t)h->n_tiles);
    dev_free(h, h->tile_order, (size_t)h->n_tiles);
    dev_free(h, h->n_contrib, px);
    dev_free(h, h->accum_alpha, px);
    dev_free(h, h->loss_maps, 9 * px);
    dev_free(h, h->loss_acc, 2);
    if (h->total_host) cudaFreeHost(h->total_host);
    for (cudaEvent_t e : h->ev)
        if (e) cudaEventDestroy(e);
    for (auto &sl : h->slot)
        for (cudaEvent_t e : {sl.h2d_done, sl.compute_done, sl.d2h_done, sl.t_h2d0, sl.t_h2d1, sl.t_c0, sl.t_c1, sl.t_d0,
                              sl.t_d1})
            if (e) cudaEventDestroy(e);
    if (h->t_base) cudaEventDestroy(h->t_base);
    if (h->h2d_stream) cudaStreamDestroy(h->h2d_stream);
    if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
    delete h;
    return GSR_OK;
}

int gsr_memory_usage(const GsrHandle *h, size_t *bytes) {
    if (!h || !bytes) return GSR_EINVAL;
    *bytes = h->bytes;
    return GSR_OK;
}

// Writes the buffers of the last forward that the hot path does not keep in the reference's form (see gsr_get_state).
static int materialize_reference_binning(GsrHandle *h) {
    if (h->ref_binning_valid || !h->fwd_valid || h->last_n <= 0) return GSR_OK;
    // With the depth pre-sort the hot path scans and emits in depth order and sorts bare tile ids; the reference's
    // buffers it does not need in that form — cumsum(tiles_touched) in index order (rasterizer.jl:333), the unsorted keys
    // / values of duplicate_with_keys! (utils.jl:85-120) and the 64-bit sorted keys (rasterizer.jl:357-372) — are
    // written here, on demand, for whoever reads the state.  The sort has consumed the depth-order emission by now,
    // so its buffers are free to hold the reference-order one.
    cudaStream_t s = h->last_stream;
    launch_scan_tiles(h->last_n, h->g.tiles_touched, nullptr, h->g.points_offset, h->scan_state, h->total_dev, s);
    if (h->last_m > 0) {
        launch_materialize_keys(h->last_m, reinterpret_cast<const uint32_t *>(h->keys_tmp), h->vals_sorted, h->g.depths,
                                h->keys_sorted, s);
        launch_duplicate(h->last_cam, h->last_n, h->g, h->g.points_offset, nullptr, h->keys_unsorted, h->vals_unsorted,
                         h->plan, nullptr, s);
    }
    CK(cudaStreamSynchronize(s));
    h->ref_binning_valid = true;
    return GSR_OK;
}

int gsr_get_state(GsrHandle *h, GsrStateViews *v) {
    if (!h || !v) return GSR_EINVAL;
    memset(v, 0, sizeof *v);
    {
        const int rc = materialize_reference_binning(h);
        if (rc) return rc;
    }
    v->n = h->last_n;
    v->n_rendered = h->last_m;
    v->radii = h->g.radii;
    v->grad_means2d = reinterpret_cast<float *>(h->g.grad_means2d);
    v->means2d = reinterpret_cast<const float *>(h->g.means2d);
    v->depths = h->g.depths;
    v->conics = h->g.conics;
    v->rgbs = h->g.rgbs;
    v->clamped = h->g.clamped;
    v->tiles_touched = h->g.tiles_touched;
    v->points_offset = h->g.points_offset;
    v->normals = h->g.normals;
    v->keys_unsorted = h->keys_unsorted;
    v->values_unsorted = h->vals_unsorted;
    v->keys_sorted = h->keys_sorted;
    v->values_sorted = h->vals_sorted;
    v->ranges = h->ranges;
    v->n_contrib = h->n_contrib;
    v->accum_alpha = h->accum_alpha;
    return GSR_OK;
}

static int forward_impl(GsrHandle *h, const GsrCamera *cam, int64_t n, int32_t sh_degree, int32_t K, const float *means,
                        const float *shs, const float *opacities, const float *scales, const float *rotations,
                        const float background[3], float *image_out, uint8_t *covis, float *uncert, int64_t *n_rendered,
                        void *stream, const ParamSpec &ps) {
    if (!h) return GSR_EINVAL;
    if (!cam || !image_out || !background) return fail(h, GSR_EINVAL, "gsr_forward: null argument");
    if (n < 0 || sh_degree < 0 || sh_degree > 3 || K < (sh_degree + 1) * (sh_degree + 1))
        return fail(h, GSR_EINVAL, "gsr_forward: need 0 <= sh_degree <= 3 and K >= (sh_degree+1)^2");
    if (n > 0 && (!means || !shs || !opacities || !scales || !rotations))
        return fail(h, GSR_EINVAL, "gsr_forward: null parameter array");
    if (reinterpret_cast<uintptr_t>(rotations) & 15)
        return fail(h, GSR_EINVAL, "gsr_forward: rotations must be 16-byte aligned (128-bit loads, projection.jl:85)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int ch = h->cfg.channels;
    const size_t image_bytes = (size_t)ch * h->cfg.width * h->cfg.height * sizeof(float);
    h->fwd_valid = false;
    h->generation++;
    if (n_rendered) *n_rendered = 0;

    int rc = ensure_geometry(h, n, s);
    if (rc) return rc;
    DevCamera dc;
    make_dev_camera(h, cam, &dc);

    int64_t m = 0;
    if (h->profile)
        for (int k = 0; k < GSR_NUM_STAGES; k++) h->ev_used[k] = false;
    if (n > 0) {
        {
            StageTimer tm(h, s, GSR_STAGE_PREPROCESS);
            launch_preprocess(dc, n, sh_degree, K, ch, means, shs, opacities, scales, rotations, h->g, s, ps);
        }
        const uint32_t *perm = nullptr;
        if (h->presort) {  // Gaussians in depth order: the instance sort below then only has the tile digits left
            StageTimer tm(h, s, GSR_STAGE_PRESORT);
            uint32_t *phist = sort_prepare(h->pplan, n, h->psort_temp, s);
            launch_presort_keys(n, h->g, h->pplan, h->pk[0], h->pv[0], phist, s);
            launch_sort_pairs(h->pplan, n, h->pk[0], h->pv[0], h->pk[1], h->pv[1], h->pk[2], h->pv[2], h->psort_temp,
                              /*hist_ready=*/true, s);
            perm = h->pv[1];
        }
        {
            StageTimer tm(h, s, GSR_STAGE_SCAN);
            launch_scan_tiles(n, h->g.tiles_touched, perm, perm ? h->offsets_sorted : h->g.points_offset, h->scan_state,
                              h->total_dev, s);
        }
        CK(cudaStreamSynchronize(s));  // the one host sync of the forward (rasterizer.jl:337)
        m = *h->total_host;
    }
    h->ref_binning_valid = !h->presort;
    h->last_stream = s;
    h->last_cam = dc;

    h->last_n = n;
    h->last_m = m;
    if (m == 0) {  // rasterizer.jl:283,338: zero image, NOT background
        CK(cudaMemsetAsync(image_out, 0, image_bytes, s));
        h->fwd_valid = true;
        return GSR_OK;
    }
    if (m >= (1ll << 30)) return fail(h, GSR_EINVAL, "gsr_forward: more than 2^30 tile instances");
    rc = ensure_binning(h, m);
    if (rc) return rc;

    if (h->presort) {
        // Emission is in depth order, so an instance's sort key is its bare 32-bit tile id: 8 bytes per instance through
        // duplicate / the tile-digit passes / the range scan instead of 12.  The three 64-bit key buffers serve as 32-bit
        // scratch (unsorted | pass scratch in the halves of keys_unsorted, sorted tile ids in keys_tmp); keys_sorted holds
        // the canonical 64-bit keys once gsr_get_state asks for them.
        uint32_t *t_in = reinterpret_cast<uint32_t *>(h->keys_unsorted), *t_tmp = t_in + h->cap_m;
        uint32_t *t_out = reinterpret_cast<uint32_t *>(h->keys_tmp);
        {
            StageTimer tm(h, s, GSR_STAGE_DUPLICATE);
            uint32_t *ghist = sort_prepare(h->tplan, m, h->sort_temp, s);
            launch_duplicate_tiles(dc, n, h->g, h->offsets_sorted, h->pv[1], t_in, h->vals_unsorted, h->tplan, ghist, s);
        }
        {
            StageTimer tm(h, s, GSR_STAGE_SORT);
            launch_sort_tiles(h->tplan, m, t_in, h->vals_unsorted, t_out, h->vals_sorted, t_tmp, h->vals_tmp, h->sort_temp, s);
        }
        {
            StageTimer tm(h, s, GSR_STAGE_RANGES);
            CK(cudaMemsetAsync(h->ranges, 0, 2 * (size_t)h->n_tiles * sizeof(uint32_t), s));  // rasterizer.jl:375
            launch_tile_ranges32(m, t_out, h->ranges, s);
            if (h->use_tile_order) launch_tile_order(h->n_tiles, h->ranges, h->tile_order, s);
        }
    } else {
        {
            StageTimer tm(h, s, GSR_STAGE_DUPLICATE);
            uint32_t *ghist = sort_prepare(h->plan, m, h->sort_temp, s);
            launch_duplicate(dc, n, h->g, h->g.points_offset, nullptr, h->keys_unsorted, h->vals_unsorted, h->plan, ghist, s);
        }
        {
            StageTimer tm(h, s, GSR_STAGE_SORT);
            launch_sort_pairs(h->plan, m, h->keys_unsorted, h->vals_unsorted, h->keys_sorted, h->vals_sorted, h->keys_tmp,
                              h->vals_tmp, h->sort_temp, /*hist_ready=*/true, s);
        }
        {
            StageTimer tm(h, s, GSR_STAGE_RANGES);
            CK(cudaMemsetAsync(h->ranges, 0, 2 * (size_t)h->n_tiles * sizeof(uint32_t), s));  // rasterizer.jl:375
            launch_tile_ranges(m, h->keys_sorted, h->ranges, s);
            if (h->use_tile_order) launch_tile_order(h->n_tiles, h->ranges, h->tile_order, s);
        }
    }
    const uint32_t *order = h->use_tile_order ? h->tile_order : nullptr;
    {
        StageTimer tm(h, s, GSR_STAGE_RENDER_FWD);
        float bg[8] = {background[0], background[1], background[2], 0.f, 0.f, 0.f, 0.f, 0.f};  // rasterizer.jl:411-414
        if (launch_render_forward(ch, h->cfg.math_mode, h->cfg.width, h->cfg.height, h->ranges, order, h->vals_sorted, h->g.rec,
                                  bg, image_out, h->n_contrib, h->accum_alpha, covis, uncert, s))
            return fail(h, GSR_EINVAL, "no compositing kernel for this math_mode");
    }
    CK(cudaGetLastError());
    if (n_rendered) *n_rendered = m;
    h->fwd_valid = true;
    return GSR_OK;
}

int gsr_forward(GsrHandle *h, const GsrCamera *cam, int64_t n, int32_t sh_degree, int32_t K, const float *means,
                const float *shs, const float *opacities, const float *scales, const float *rotations,
                const float background[3], float *image_out, uint8_t *covis, float *uncert, int64_t *n_rendered,
                void *stream) {
    return forward_impl(h, cam, n, sh_degree, K, means, shs, opacities, scales, rotations, background, image_out, covis,
                        uncert, n_rendered, stream, ParamSpec());
}

static int raw_spec(GsrHandle *h, int32_t K, const float *features_rest, float *vfeatures_rest, bool backward,
                    int32_t isotropic, ParamSpec *ps) {
    static const float dummy_rest = 0.f;
    static float dummy_vrest = 0.f;
    if (K > 1 && (!features_rest || (backward && !vfeatures_rest)))
        return fail(h, GSR_EINVAL, "raw parameters: features_rest (and its cotangent) required when K > 1");
    ps->raw_opacity = 1;
    ps->raw_scale = 1;
    ps->isotropic = isotropic ? 1 : 0;
    // K == 1: no remainder coefficients; a non-null tag still selects the split layout (it is never dereferenced)
    ps->sh_rest = features_rest ? features_rest : &dummy_rest;
    ps->vsh_rest = vfeatures_rest ? vfeatures_rest : &dummy_vrest;
    return GSR_OK;
}

int gsr_forward_raw(GsrHandle *h, const GsrCamera *cam, int64_t n, int32_t sh_degree, int32_t K, const float *means,
                    const float *features_dc, const float *features_rest, const float *opacities_raw,
                    const float *scales_raw, int32_t isotropic, const float *rotations, const float background[3],
                    float *image_out, uint8_t *covis, float *uncert, int64_t *n_rendered, void *stream) {
    if (!h) return GSR_EINVAL;
    ParamSpec ps;
    int rc = raw_spec(h, K, features_rest, nullptr, false, isotropic, &ps);
    if (rc) return rc;
    return forward_impl(h, cam, n, sh_degree, K, means, features_dc, opacities_raw, scales_raw, rotations, background,
                        image_out, covis, uncert, n_rendered, stream, ps);
}

static int backward_impl(GsrHandle *h, const GsrCamera *cam, int64_t n, int32_t sh_degree, int32_t K, const float *means,
                         const float *shs, const float *opacities, const float *scales, const float *rotations,
                         const float background[3], const float *vpixels, float *vmeans, float *vshs, float *vopacities,
                         float *vscales, float *vrot, float *vR, float *vt, int32_t accumulate, void *stream,
                         const ParamSpec &ps) {
    if (!h) return GSR_EINVAL;
    if (!cam || !background || !vpixels || !vmeans || !vshs || !vopacities || !vscales || !vrot || !opacities)
        return fail(h, GSR_EINVAL, "gsr_backward: null argument");
    if (!h->fwd_valid || n != h->last_n)
        return fail(h, GSR_ESTATE, "gsr_backward: no matching gsr_forward on this handle");
    if ((vR == nullptr) != (vt == nullptr)) return fail(h, GSR_EINVAL, "gsr_backward: pass both vR and vt or neither");
    if (sh_degree < 0 || sh_degree > 3 || K < (sh_degree + 1) * (sh_degree + 1))
        return fail(h, GSR_EINVAL, "gsr_backward: need 0 <= sh_degree <= 3 and K >= (sh_degree+1)^2");
    if ((reinterpret_cast<uintptr_t>(rotations) & 15) || (reinterpret_cast<uintptr_t>(vrot) & 15))
        return fail(h, GSR_EINVAL, "gsr_backward: rotations / vrot must be 16-byte aligned");
    if (n == 0) return GSR_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int ch = h->cfg.channels;
    DevCamera dc;
    make_dev_camera(h, cam, &dc);
    {   // always: a Gaussian can have radius > 0 and no tile (get_rect yields x1 == x0), so the per-Gaussian backward
        // reads accumulator rows even when no instance was emitted
        StageTimer tm(h, s, GSR_STAGE_ZERO_GRADS);
        CK(cudaMemsetAsync(h->g.gacc, 0, (size_t)n * acc_floats(ch) * sizeof(float), s));
    }
    if (h->last_m > 0) {
        StageTimer tm(h, s, GSR_STAGE_RENDER_BWD);
        float bg[8] = {background[0], background[1], background[2], 0.f, 0.f, 0.f, 0.f, 0.f};
        if (launch_render_backward(ch, h->cfg.math_mode, h->cfg.width, h->cfg.height, h->ranges,
                                   h->use_tile_order ? h->tile_order : nullptr, h->vals_sorted, h->g.rec,
                                   bg, vpixels, h->n_contrib, h->accum_alpha, h->g.gacc, s))
            return fail(h, GSR_EINVAL, "no compositing kernel for this math_mode");
    }
    {
        StageTimer tm(h, s, GSR_STAGE_GAUSS_BWD);
        launch_backward_gaussians(dc, n, sh_degree, K, ch, means, shs, opacities, scales, rotations, h->g, vmeans,
                                  vshs, vopacities, vscales, vrot, vR, vt, accumulate, s, ps);
    }
    CK(cudaGetLastError());
    return GSR_OK;
}

int gsr_backward(GsrHandle *h, const GsrCamera *cam, int64_t n, int32_t sh_degree, int32_t K, const float *means,
                 const float *shs, const float *opacities, const float *scales, const float *rotations,
                 const float background[3], const float *vpixels, float *vmeans, float *vshs, float *vopacities,
                 float *vscales, float *vrot, float *vR, float *vt, int32_t accumulate, void *stream) {
    return backward_impl(h, cam, n, sh_degree, K, means, shs, opacities, scales, rotations, background, vpixels, vmeans,
                         vshs, vopacities, vscales, vrot, vR, vt, accumulate, stream, ParamSpec());
}

int gsr_backward_raw(GsrHandle *h, const GsrCamera *cam, int64_t n, int32_t sh_degree, int32_t K, const float *means,
                     const float *features_dc, const float *features_rest, const float *opacities_raw,
                     const float *scales_raw, int32_t isotropic, const float *rotations, const float background[3],
                     const float *vpixels, float *vmeans, float *vfeatures_dc, float *vfeatures_rest,
                     float *vopacities_raw, float *vscales_raw, float *vrot, float *vR, float *vt, int32_t accumulate,
                     void *stream) {
    if (!h) return GSR_EINVAL;
    ParamSpec ps;
    int rc = raw_spec(h, K, features_rest, vfeatures_rest, true, isotropic, &ps);
    if (rc) return rc;
    return backward_impl(h, cam, n, sh_degree, K, means, features_dc, opacities_raw, scales_raw, rotations, background,
                         vpixels, vmeans, vfeatures_dc, vopacities_raw, vscales_raw, vrot, vR, vt, accumulate, stream, ps);
}

int gsr_set_accumulator(GsrHandle *h, float *gacc_dev, int64_t capacity_gaussians) {
    if (!h) return GSR_EINVAL;
    if ((gacc_dev == nullptr) != (capacity_gaussians == 0) || (reinterpret_cast<uintptr_t>(gacc_dev) & 15))
        return fail(h, GSR_EINVAL, "gsr_set_accumulator: need a 16-byte aligned buffer and its capacity (or NULL, 0)");
    if (gacc_dev == h->gacc_external && capacity_gaussians == h->gacc_external_cap) return GSR_OK;
    const size_t af = (size_t)acc_floats(h->cfg.channels);
    if (h->cap_n > 0 && gacc_dev && capacity_gaussians >= h->cap_n) {
        // Only the pointer changes (one accumulator per view of a batch): the forward never touches the accumulator, so
        // the state of the last forward stays valid and nothing else is reallocated.
        if (!h->gacc_external) dev_free(h, h->g.gacc, af * (size_t)h->cap_n);  // the private buffer is no longer needed
        h->g.gacc = gacc_dev;
    } else {
        free_geometry(h);  // the geometry state is rebuilt around the new accumulator on the next forward
        h->fwd_valid = false;
    }
    h->gacc_external = gacc_dev;
    h->gacc_external_cap = capacity_gaussians;
    return GSR_OK;
}

int gsr_backward_render(GsrHandle *h, int64_t n, const float background[3], const float *vpixels, void *stream) {
    if (!h) return GSR_EINVAL;
    if (!background || !vpixels) return fail(h, GSR_EINVAL, "gsr_backward_render: null argument");
    if (!h->fwd_valid || n != h->last_n) return fail(h, GSR_ESTATE, "gsr_backward_render: no matching gsr_forward on this handle");
    if (n == 0) return GSR_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int ch = h->cfg.channels;
    {
        StageTimer tm(h, s, GSR_STAGE_ZERO_GRADS);
        CK(cudaMemsetAsync(h->g.gacc, 0, (size_t)n * acc_floats(ch) * sizeof(float), s));
        launch_pack_flags(n, ch, h->g.radii, h->g.clamped, h->g.gacc, s);
    }
    if (h->last_m > 0) {
        StageTimer tm(h, s, GSR_STAGE_RENDER_BWD);
        float bg[8] = {background[0], background[1], background[2], 0.f, 0.f, 0.f, 0.f, 0.f};
        if (launch_render_backward(ch, h->cfg.math_mode, h->cfg.width, h->cfg.height, h->ranges,
                                   h->use_tile_order ? h->tile_order : nullptr, h->vals_sorted, h->g.rec,
                                   bg, vpixels, h->n_contrib, h->accum_alpha, h->g.gacc, s))
            return fail(h, GSR_EINVAL, "no compositing kernel for this math_mode");
    }
    launch_grad_means2d(n, ch, h->g.radii, h->g.conics, h->g.gacc, h->g.grad_means2d, s);
    CK(cudaGetLastError());
    return GSR_OK;
}

int gsr_export_accumulator(GsrHandle *h, int64_t n, float *rows_dev, void *stream) {
    if (!h) return GSR_EINVAL;
    if (!rows_dev || (reinterpret_cast<uintptr_t>(rows_dev) & 15))
        return fail(h, GSR_EINVAL, "gsr_export_accumulator: need a 16-byte aligned destination");
    if (!h->fwd_valid || n != h->last_n) return fail(h, GSR_ESTATE, "gsr_export_accumulator: no matching gsr_forward / gsr_backward_render");
    launch_export_rows(n, h->cfg.channels, h->g.gacc, rows_dev, static_cast<cudaStream_t>(stream));
    CK(cudaGetLastError());
    return GSR_OK;
}

int gsr_backward_gaussians_views(GsrHandle *h, int32_t n_views, const GsrCamera *cams, const float *const *view_gacc,
                                 int32_t exchange_rows, int32_t world, int32_t rank, float *const *peer_tables, int64_t n, int32_t sh_degree,
                                 int32_t K, const float *means, const float *shs, const float *opacities,
                                 const float *scales, const float *rotations, void *stream) {
    if (!h) return GSR_EINVAL;
    if (n_views < 1 || n_views > GSR_MAX_VIEWS || !cams || !view_gacc)
        return fail(h, GSR_EINVAL, "gsr_backward_gaussians_views: need 1 <= n_views <= 16 and per-view camera / accumulator arrays");
    if (world < 1 || world > GSR_MAX_PEERS || rank < 0 || rank >= world || !peer_tables)
        return fail(h, GSR_EINVAL, "gsr_backward_gaussians_views: need 1 <= world <= 8, 0 <= rank < world and per-rank tables");
    if (!means || !shs || !opacities || !scales || !rotations || n < 0)
        return fail(h, GSR_EINVAL, "gsr_backward_gaussians_views: null parameter array");
    if (sh_degree < 0 || sh_degree > 3 || K < (sh_degree + 1) * (sh_degree + 1))
        return fail(h, GSR_EINVAL, "gsr_backward_gaussians_views: need 0 <= sh_degree <= 3 and K >= (sh_degree+1)^2");
    if (n == 0) return GSR_OK;
    PeerArgs a;
    memset(&a, 0, sizeof a);
    a.vsh_aligned = 1;
    for (int v = 0; v < n_views; v++) {
        if (!view_gacc[v]) return fail(h, GSR_EINVAL, "gsr_backward_gaussians_views: null accumulator pointer");
        if (reinterpret_cast<uintptr_t>(view_gacc[v]) & 15)
            return fail(h, GSR_EINVAL, "gsr_backward_gaussians_views: accumulators must be 16-byte aligned");
        memcpy(a.cams[v].R, cams[v].R, sizeof a.cams[v].R);
        memcpy(a.cams[v].t, cams[v].t, sizeof a.cams[v].t);
        memcpy(a.cams[v].focal, cams[v].focal, sizeof a.cams[v].focal);
        memcpy(a.cams[v].principal, cams[v].principal, sizeof a.cams[v].principal);
        memcpy(a.cams[v].cam_center, cams[v].cam_center, sizeof a.cams[v].cam_center);
        a.cams[v].width = h->cfg.width;
        a.cams[v].height = h->cfg.height;
        a.cams[v].blur_eps = h->cfg.blur_eps;
        a.gacc[v] = view_gacc[v];
    }
    if (!peer_tables[rank]) return fail(h, GSR_EINVAL, "gsr_backward_gaussians_views: the calling rank's table pointer is null");
    for (int p = 0; p < world; p++) {
        if (!peer_tables[p]) continue;  // that rank keeps only its own slice (reduce-scatter semantics)
        if (reinterpret_cast<uintptr_t>(peer_tables[p]) & 15)
            return fail(h, GSR_EINVAL, "gsr_backward_gaussians_views: tables must be 16-byte aligned");
        a.table[p] = peer_tables[p];
        if (reinterpret_cast<uintptr_t>(peer_tables[p] + 11 * n) & 15) a.vsh_aligned = 0;
    }
    a.n_views = n_views;
    a.exchange_rows = exchange_rows ? 1 : 0;
    a.world = world;
    a.rank = rank;
    a.n = n;
    int64_t chunk = (n + world - 1) / world;
    chunk = (chunk + 63) / 64 * 64;  // slices start on a CTA boundary (16-byte aligned SH spans)
    a.lo = rank * chunk < n ? rank * chunk : n;
    a.hi = a.lo + chunk < n ? a.lo + chunk : n;
    a.sh_degree = sh_degree;
    a.K = K;
    a.channels = h->cfg.channels;
    a.sh_stride = (3 * K) | 1;
    a.means = means; a.shs = shs; a.opac = opacities; a.scales = scales; a.rots = rotations;
    StageTimer tm(h, static_cast<cudaStream_t>(stream), GSR_STAGE_GAUSS_BWD);
    if (launch_backward_gaussians_peers(a, static_cast<cudaStream_t>(stream)) != 0)
        return fail(h, GSR_ECUDA, "gsr_backward_gaussians_views: launch failed");
    return GSR_OK;
}

int gsr_backward_gaussians_peers(GsrHandle *h, int32_t world, int32_t rank, const GsrCamera *cams,
                                 const float *const *peer_gacc, float *const *peer_tables, int64_t n, int32_t sh_degree,
                                 int32_t K, const float *means, const float *shs, const float *opacities,
                                 const float *scales, const float *rotations, void *stream) {
    // one view per rank: view v's accumulator is rank v's, in the handle's own row layout
    return gsr_backward_gaussians_views(h, world, cams, peer_gacc, 0, world, rank, peer_tables, n, sh_degree, K, means, shs,
                                        opacities, scales, rotations, stream);
}

int gsr_update_stats(GsrHandle *h, int64_t n, int32_t *max_radii, float *accum_grad_means2d, float *denom,
                     void *stream) {
    if (!h) return GSR_EINVAL;
    if (!max_radii || !accum_grad_means2d || !denom) return fail(h, GSR_EINVAL, "gsr_update_stats: null argument");
    if (n > h->last_n) return fail(h, GSR_ESTATE, "gsr_update_stats: n exceeds the last forward");
    launch_update_stats(n, h->g.radii, h->g.grad_means2d, (uint32_t)h->cfg.width, (uint32_t)h->cfg.height, max_radii,
                        accum_grad_means2d, denom, static_cast<cudaStream_t>(stream));
    CK(cudaGetLastError());
    return GSR_OK;
}

int gsr_forward_backward_host_async(GsrHandle *h, const GsrCamera *cam, int64_t n, int32_t sh_degree, int32_t K,
                                    const float *means_h, const float *shs_h, const float *opacities_h,
                                    const float *scales_h, const float *rotations_h, const float background[3],
                                    const float *vpixels_h, float *image_h, float *vmeans_h, float *vshs_h,
                                    float *vopacities_h, float *vscales_h, float *vrot_h, int64_t *n_rendered,
                                    void *stream) {
    if (!h) return GSR_EINVAL;
    if (!means_h || !shs_h || !opacities_h || !scales_h || !rotations_h || !vpixels_h)
        return fail(h, GSR_EINVAL, "gsr_forward_backward_host: null input");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t f = sizeof(float), N = (size_t)n;
    const size_t img = (size_t)h->cfg.channels * h->cfg.width * h->cfg.height * f;
    if (!h->h2d_stream) {
        CK(cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
        for (auto &sl : h->slot) {
            CK(cudaEventCreateWithFlags(&sl.h2d_done, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&sl.compute_done, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&sl.d2h_done, cudaEventDisableTiming));
            for (cudaEvent_t *e : {&sl.t_h2d0, &sl.t_h2d1, &sl.t_c0, &sl.t_c1, &sl.t_d0, &sl.t_d1}) CK(cudaEventCreate(e));
        }
        CK(cudaEventCreate(&h->t_base));
        CK(cudaEventRecord(h->t_base, h->h2d_stream));
    }
    GsrHandle::HostSlot &sl = h->slot[h->next_slot];
    h->next_slot ^= 1;
    const bool reused = sl.d2h_pending;  // the slot carried the submission before the previous one
    int rc;
    if ((rc = ensure_stage(h, sl.means, 3 * N * f)) || (rc = ensure_stage(h, sl.shs, 3 * (size_t)K * N * f)) ||
        (rc = ensure_stage(h, sl.opac, N * f)) || (rc = ensure_stage(h, sl.scales, 3 * N * f)) ||
        (rc = ensure_stage(h, sl.rots, 4 * N * f)) || (rc = ensure_stage(h, sl.vpix, img)) ||
        (rc = ensure_stage(h, sl.image, img)) || (rc = ensure_stage(h, sl.vmeans, 3 * N * f)) ||
        (rc = ensure_stage(h, sl.vshs, 3 * (size_t)K * N * f)) || (rc = ensure_stage(h, sl.vopac, N * f)) ||
        (rc = ensure_stage(h, sl.vscales, 3 * N * f)) || (rc = ensure_stage(h, sl.vrot, 4 * N * f)))
        return rc;
    // ---- H2D on the upload stream (overlaps the previous step's compute and download) ----------------------
    cudaStream_t up = h->h2d_stream;
    // slot reuse is ordered on the device, never on the host: the upload waits for the compute that last read this
    // slot's inputs; the compute waits for the download that last read this slot's outputs
    // (the compute that last read this slot's inputs, two submissions ago, is already complete: the previous
    //  submission's forward synchronised the caller's stream)
    if (reused) {
        CK(cudaStreamWaitEvent(up, sl.compute_done, 0));
        CK(cudaStreamWaitEvent(s, sl.d2h_done, 0));
    }
    if (h->profile) CK(cudaEventRecord(sl.t_h2d0, up));
    CK(cudaMemcpyAsync(sl.means.p, means_h, 3 * N * f, cudaMemcpyHostToDevice, up));
    CK(cudaMemcpyAsync(sl.scales.p, scales_h, 3 * N * f, cudaMemcpyHostToDevice, up));
    CK(cudaMemcpyAsync(sl.rots.p, rotations_h, 4 * N * f, cudaMemcpyHostToDevice, up));
    CK(cudaMemcpyAsync(sl.opac.p, opacities_h, N * f, cudaMemcpyHostToDevice, up));
    CK(cudaMemcpyAsync(sl.shs.p, shs_h, 3 * (size_t)K * N * f, cudaMemcpyHostToDevice, up));
    CK(cudaMemcpyAsync(sl.vpix.p, vpixels_h, img, cudaMemcpyHostToDevice, up));
    if (h->profile) CK(cudaEventRecord(sl.t_h2d1, up));
    CK(cudaEventRecord(sl.h2d_done, up));
    CK(cudaStreamWaitEvent(s, sl.h2d_done, 0));
    if (h->profile) CK(cudaEventRecord(sl.t_c0, s));
    // ---- compute on the caller's stream ---------------------------------------------------------------------
    auto F = [](DevBuf &b) { return static_cast<float *>(b.p); };
    rc = gsr_forward(h, cam, n, sh_degree, K, F(sl.means), F(sl.shs), F(sl.opac), F(sl.scales), F(sl.rots), background,
                     F(sl.image), nullptr, nullptr, n_rendered, stream);
    if (rc) return rc;
    cudaStream_t down = h->d2h_stream;
    if (image_h) {  // the image can leave while the backward runs
        CK(cudaEventRecord(sl.compute_done, s));
        CK(cudaStreamWaitEvent(down, sl.compute_done, 0));
        CK(cudaMemcpyAsync(image_h, sl.image.p, img, cudaMemcpyDeviceToHost, down));
    }
    rc = gsr_backward(h, cam, n, sh_degree, K, F(sl.means), F(sl.shs), F(sl.opac), F(sl.scales), F(sl.rots), background,
                      F(sl.vpix), F(sl.vmeans), F(sl.vshs), F(sl.vopac), F(sl.vscales), F(sl.vrot), nullptr, nullptr, 0,
                      stream);
    if (rc) return rc;
    // ---- D2H on the download stream ---------------------------------------------------------------------------
    if (h->profile) CK(cudaEventRecord(sl.t_c1, s));
    CK(cudaEventRecord(sl.compute_done, s));
    CK(cudaStreamWaitEvent(down, sl.compute_done, 0));
    if (h->profile) CK(cudaEventRecord(sl.t_d0, down));
    if (vmeans_h) CK(cudaMemcpyAsync(vmeans_h, sl.vmeans.p, 3 * N * f, cudaMemcpyDeviceToHost, down));
    if (vopacities_h) CK(cudaMemcpyAsync(vopacities_h, sl.vopac.p, N * f, cudaMemcpyDeviceToHost, down));
    if (vscales_h) CK(cudaMemcpyAsync(vscales_h, sl.vscales.p, 3 * N * f, cudaMemcpyDeviceToHost, down));
    if (vrot_h) CK(cudaMemcpyAsync(vrot_h, sl.vrot.p, 4 * N * f, cudaMemcpyDeviceToHost, down));
    if (vshs_h) CK(cudaMemcpyAsync(vshs_h, sl.vshs.p, 3 * (size_t)K * N * f, cudaMemcpyDeviceToHost, down));
    if (h->profile) CK(cudaEventRecord(sl.t_d1, down));
    CK(cudaEventRecord(sl.d2h_done, down));
    sl.d2h_pending = true;
    return GSR_OK;
}

// debug: {h2d start, h2d end, compute start, compute end, d2h start, d2h end} of both slots, ms since the first
// submission (valid after gsr_host_wait with profiling enabled)
int gsr_host_timeline(GsrHandle *h, float out[12]) {
    if (!h || !out || !h->t_base) return GSR_EINVAL;
    for (int k = 0; k < 2; k++) {
        cudaEvent_t ev[6] = {h->slot[k].t_h2d0, h->slot[k].t_h2d1, h->slot[k].t_c0, h->slot[k].t_c1, h->slot[k].t_d0, h->slot[k].t_d1};
        for (int j = 0; j < 6; j++) {
            out[6 * k + j] = 0.f;
            if (ev[j] && cudaEventQuery(ev[j]) == cudaSuccess) cudaEventElapsedTime(&out[6 * k + j], h->t_base, ev[j]);
        }
    }
    cudaGetLastError();
    return GSR_OK;
}

int gsr_host_wait(GsrHandle *h) {
    if (!h) return GSR_EINVAL;
    for (auto &sl : h->slot)
        if (sl.d2h_pending) { CK(cudaEventSynchronize(sl.d2h_done)); sl.d2h_pending = false; }
    return GSR_OK;
}

int gsr_forward_backward_host(GsrHandle *h, const GsrCamera *cam, int64_t n, int32_t sh_degree, int32_t K,
                              const float *means_h, const float *shs_h, const float *opacities_h,
                              const float *scales_h, const float *rotations_h, const float background[3],
                              const float *vpixels_h, float *image_h, float *vmeans_h, float *vshs_h,
                              float *vopacities_h, float *vscales_h, float *vrot_h, int64_t *n_rendered,
                              void *stream) {
    int rc = gsr_forward_backward_host_async(h, cam, n, sh_degree, K, means_h, shs_h, opacities_h, scales_h, rotations_h,
                                             background, vpixels_h, image_h, vmeans_h, vshs_h, vopacities_h, vscales_h,
                                             vrot_h, n_rendered, stream);
    if (rc) return rc;
    return gsr_host_wait(h);
}

int gsr_profile_enable(GsrHandle *h, int32_t enable) {
    if (!h) return GSR_EINVAL;
    if (enable && !h->ev[0])
        for (cudaEvent_t &e : h->ev) CK(cudaEventCreate(&e));
    h->profile = enable != 0;
    for (bool &u : h->ev_used) u = false;
    return GSR_OK;
}

int gsr_profile_get(GsrHandle *h, float ms[GSR_NUM_STAGES]) {
    if (!h || !ms) return GSR_EINVAL;
    for (int k = 0; k < GSR_NUM_STAGES; k++) {
        ms[k] = 0.f;
        if (h->profile && h->ev_used[k]) {
            CK(cudaEventSynchronize(h->ev[2 * k + 1]));
            CK(cudaEventElapsedTime(&ms[k], h->ev[2 * k], h->ev[2 * k + 1]));
        }
    }
    return GSR_OK;
}

int gsr_debug_exp_neg(const float *sigma_dev, float *split_dev, float *libdevice_dev, float *inlined_dev, int64_t n,
                      void *stream) {
    if (n < 0 || (n > 0 && (!sigma_dev || !split_dev || !libdevice_dev))) return GSR_EINVAL;
    return launch_exp_neg_probe(sigma_dev, split_dev, libdevice_dev, inlined_dev, n, static_cast<cudaStream_t>(stream)) ? GSR_ECUDA : GSR_OK;
}

int gsr_measure_fp32_peak(double *tflops, void *stream) {
    if (!tflops) return GSR_EINVAL;
    double ms = 0.0, flops = 0.0;
    if (launch_fp32_peak(static_cast<cudaStream_t>(stream), &ms, &flops) != 0) return GSR_ECUDA;
    *tflops = flops / (ms * 1e-3) / 1e12;
    return GSR_OK;
}

int gsr_densify_masks(int64_t n, int64_t n_grad, const float *accum_dev, const float *denom_dev, const float *scales_dev,
                      int32_t isotropic, float grad_threshold, float gamma, uint8_t *clone_mask_dev,
                      uint8_t *split_mask_dev, void *stream) {
    GsrHandle *h = nullptr;
    if (n < 0 || n_grad < 0 || n_grad > n) return fail(h, GSR_EINVAL, "gsr_densify_masks: need 0 <= n_grad <= n");
    if (n > 0 && (!scales_dev || (n_grad > 0 && (!accum_dev || !denom_dev)))) return fail(h, GSR_EINVAL, "gsr_densify_masks: null argument");
    if (launch_densify_masks(n, n_grad, accum_dev, denom_dev, scales_dev, isotropic, grad_threshold, gamma, clone_mask_dev,
                             split_mask_dev, static_cast<cudaStream_t>(stream)))
        return cuda_fail(h, cudaGetLastError(), "densify_masks_kernel");
    return GSR_OK;
}

int gsr_prune_mask(int64_t n, const float *opacities_dev, const float *scales_dev, int32_t isotropic,
                   const int32_t *max_radii_dev, float min_opacity, int32_t max_screen_size, float gamma,
                   uint8_t *valid_mask_dev, void *stream) {
    GsrHandle *h = nullptr;
    if (n < 0) return fail(h, GSR_EINVAL, "gsr_prune_mask: n < 0");
    if (n > 0 && (!opacities_dev || !valid_mask_dev || (max_screen_size > 0 && (!max_radii_dev || !scales_dev))))
        return fail(h, GSR_EINVAL, "gsr_prune_mask: null argument");
    if (launch_prune_mask(n, opacities_dev, scales_dev, isotropic, max_radii_dev, min_opacity, max_screen_size, gamma,
                          valid_mask_dev, static_cast<cudaStream_t>(stream)))
        return cuda_fail(h, cudaGetLastError(), "prune_mask_kernel");
    return GSR_OK;
}

size_t gsr_mask_offsets_scratch_words(int64_t n) { return n > 0 ? mask_offsets_scratch_words(n) : 1; }

int gsr_mask_offsets(int64_t n, const uint8_t *mask_dev, int32_t *offsets_dev, int64_t *count_dev, int32_t *scratch_dev,
                     void *stream) {
    GsrHandle *h = nullptr;
    if (n < 0 || n >= (1ll << 31) || !count_dev) return fail(h, GSR_EINVAL, "gsr_mask_offsets: bad argument");
    if (n > 0 && (!mask_dev || !offsets_dev || !scratch_dev)) return fail(h, GSR_EINVAL, "gsr_mask_offsets: null argument");
    if (launch_mask_offsets(n, mask_dev, offsets_dev, count_dev, scratch_dev, static_cast<cudaStream_t>(stream)))
        return cuda_fail(h, cudaGetLastError(), "mask_offsets");
    return GSR_OK;
}

int gsr_gather_rows(int64_t n, int32_t row_bytes, const void *src_dev, const uint8_t *mask_dev, const int32_t *offsets_dev,
                    void *dst_dev, int32_t repeat, int64_t count, void *stream) {
    GsrHandle *h = nullptr;
    if (n < 0 || row_bytes < 0 || (row_bytes & 3) || repeat < 1 || count < 0)
        return fail(h, GSR_EINVAL, "gsr_gather_rows: need row_bytes % 4 == 0, repeat >= 1");
    if (n == 0 || row_bytes == 0 || count == 0) return GSR_OK;
    if (!src_dev || !mask_dev || !offsets_dev || !dst_dev) return fail(h, GSR_EINVAL, "gsr_gather_rows: null argument");
    if (launch_gather_rows(n, row_bytes / 4, src_dev, mask_dev, offsets_dev, dst_dev, repeat, count,
                           static_cast<cudaStream_t>(stream)))
        return cuda_fail(h, cudaGetLastError(), "gather_rows_kernel");
    return GSR_OK;
}

int gsr_split_children(int64_t m, float *points_dev, float *scales_dev, int32_t isotropic, const float *rotations_dev,
                       const float *noise_dev, int32_t n_split, void *stream) {
    GsrHandle *h = nullptr;
    if (m < 0 || n_split < 1) return fail(h, GSR_EINVAL, "gsr_split_children: bad argument");
    if (m == 0) return GSR_OK;
    if (!points_dev || !scales_dev || !rotations_dev || !noise_dev || (reinterpret_cast<uintptr_t>(rotations_dev) & 15))
        return fail(h, GSR_EINVAL, "gsr_split_children: null argument or rotations not 16-byte aligned");
    if (launch_split_children(m, points_dev, scales_dev, isotropic, rotations_dev, noise_dev, n_split,
                              static_cast<cudaStream_t>(stream)))
        return cuda_fail(h, cudaGetLastError(), "split_children_kernel");
    return GSR_OK;
}

int gsr_ssim_forward(int32_t width, int32_t height, int32_t channels, int32_t batch, const float *img_dev,
                     const float *ref_dev, float C1, float C2, int32_t train, float *ssim_map_dev, float *dm_dmu1_dev,
                     float *dm_dsigma1_sq_dev, float *dm_dsigma12_dev, void *stream) {
    GsrHandle *h = nullptr;
    if (width < 0 || height < 0 || channels < 0 || batch < 0 || (int64_t)channels * batch > 65535)
        return fail(h, GSR_EINVAL, "gsr_ssim_forward: bad shape");
    if ((int64_t)width * height * channels * batch == 0) return GSR_OK;
    if (!img_dev || !ref_dev || !ssim_map_dev || (train && (!dm_dmu1_dev || !dm_dsigma1_sq_dev || !dm_dsigma12_dev)))
        return fail(h, GSR_EINVAL, "gsr_ssim_forward: null argument");
    if (launch_ssim_forward(width, height, channels, batch, img_dev, ref_dev, C1, C2, train, ssim_map_dev, dm_dmu1_dev,
                            dm_dsigma1_sq_dev, dm_dsigma12_dev, static_cast<cudaStream_t>(stream)))
        return cuda_fail(h, cudaGetLastError(), "ssim_fwd_kernel");
    return GSR_OK;
}

int gsr_ssim_backward(int32_t width, int32_t height, int32_t channels, int32_t batch, const float *img_dev,
                      const float *ref_dev, const float *dL_dmap_dev, const float *dm_dmu1_dev,
                      const float *dm_dsigma1_sq_dev, const float *dm_dsigma12_dev, float *dL_dimg_dev, void *stream) {
    GsrHandle *h = nullptr;
    if (width < 0 || height < 0 || channels < 0 || batch < 0 || (int64_t)channels * batch > 65535)
        return fail(h, GSR_EINVAL, "gsr_ssim_backward: bad shape");
    if ((int64_t)width * height * channels * batch == 0) return GSR_OK;
    if (!img_dev || !ref_dev || !dL_dmap_dev || !dm_dmu1_dev || !dm_dsigma1_sq_dev || !dm_dsigma12_dev || !dL_dimg_dev)
        return fail(h, GSR_EINVAL, "gsr_ssim_backward: null argument");
    if (launch_ssim_backward(width, height, channels, batch, img_dev, ref_dev, dL_dmap_dev, dm_dmu1_dev,
                             dm_dsigma1_sq_dev, dm_dsigma12_dev, dL_dimg_dev, static_cast<cudaStream_t>(stream)))
        return cuda_fail(h, cudaGetLastError(), "ssim_bwd_kernel");
    return GSR_OK;
}

int gsr_photometric_loss(GsrHandle *h, const float *image_dev, const float *target_dev, float lambda_dssim,
                         float *vpixels_dev, float *loss_dev, void *stream) {
    if (!h) return GSR_EINVAL;
    if (!image_dev || !target_dev || !vpixels_dev || !loss_dev)
        return fail(h, GSR_EINVAL, "gsr_photometric_loss: null argument");
    if (!(lambda_dssim >= 0.f && lambda_dssim <= 1.f)) return fail(h, GSR_EINVAL, "gsr_photometric_loss: lambda outside [0,1]");
    const size_t px = (size_t)h->cfg.width * h->cfg.height;
    if (!h->loss_maps) {
        CK(dev_alloc(h, &h->loss_maps, 9 * px));
        CK(dev_alloc(h, &h->loss_acc, 2));
    }
    // C1, C2: the Float32 keyword defaults of fused_ssim (fused_ssim.jl:391)
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    if (launch_photometric_loss(h->cfg.width, h->cfg.height, h->cfg.channels, image_dev, target_dev, lambda_dssim, C1, C2,
                                h->loss_maps, h->loss_acc, vpixels_dev, loss_dev, static_cast<cudaStream_t>(stream)))
        return cuda_fail(h, cudaGetLastError(), "photometric loss");
    return GSR_OK;
}

int gsr_identify_tile_range(const uint64_t *keys_dev, int64_t m, uint32_t *ranges_dev, void *stream) {
    if (m < 0 || (m > 0 && (!keys_dev || !ranges_dev))) return GSR_EINVAL;
    launch_tile_ranges(m, keys_dev, ranges_dev, static_cast<cudaStream_t>(stream));
    return cudaGetLastError() == cudaSuccess ? GSR_OK : GSR_ECUDA;
}

int gsr_sort_pairs(GsrHandle *h, const uint64_t *keys_in_dev, const uint32_t *vals_in_dev, int64_t m,
                   uint64_t *keys_out_dev, uint32_t *vals_out_dev, void *stream) {
    if (!h) return GSR_EINVAL;
    if (m < 0 || m >= (1ll << 30)) return fail(h, GSR_EINVAL, "gsr_sort_pairs: m out of range");
    if (m == 0) return GSR_OK;
    if (!keys_in_dev || !vals_in_dev || !keys_out_dev || !vals_out_dev)
        return fail(h, GSR_EINVAL, "gsr_sort_pairs: null argument");
    int rc = materialize_reference_binning(h);  // this call reuses the pass scratch that still holds the sorted tile ids
    if (rc) return rc;
    rc = ensure_binning(h, m);
    if (rc) return rc;
    sort_prepare(h->plan, m, h->sort_temp, static_cast<cudaStream_t>(stream));
    launch_sort_pairs(h->plan, m, keys_in_dev, vals_in_dev, keys_out_dev, vals_out_dev, h->keys_tmp, h->vals_tmp,
                      h->sort_temp, /*hist_ready=*/false, static_cast<cudaStream_t>(stream));
    CK(cudaGetLastError());
    return GSR_OK;
}

}  // extern "C"
