// backward_peers.cu — per-Gaussian backward FUSED with the cross-GPU gradient reduction (NVLink peer memory).
//
// Multi-GPU semantics: batch gradient = sum over views of ∇rasterize; views are sharded over
// ranks, parameters replicated.  The baseline runs backward_gaussians_kernel on every rank over all N Gaussians
// and then ncclAllReduce()s the 59-float-per-Gaussian table (236 B/Gaussian, 2(G-1)/G x 236 MB per link at 1M).
//
// This kernel does both steps at once over peer-mapped (symmetric) memory:
//   * rank r owns the Gaussian slice S_r = [lo, hi);
//   * for every Gaussian of its slice it LOADS the 64/80-byte moment accumulator of EVERY view straight
//     from that rank's HBM over NVLink (P2P loads: (G-1)/G x 64 B/Gaussian instead of 236 B), applies that view's
//     ∇project / ∇spherical_harmonics chain (cameras of all views are kernel parameters; Gaussian parameters are
//     replicated) and sums the G per-view gradients in registers / shared memory in a fixed order (deterministic,
//     unlike an all-reduce);
//   * the reduced row is STORED into every rank's gradient table (P2P stores), so all ranks end up with the full
//     replicated sum exactly as after an all-reduce.
// Link traffic per rank: (G-1)/G x (48 + 236) B/Gaussian vs 2(G-1)/G x 236 for the ring all-reduce, and the
// transfers overlap the math of the other Gaussians in flight.  The caller brackets the kernel with two
// symmetric-memory barriers (accumulators complete / tables complete).
//
// The per-view math is the body of backward_gaussians_kernel (same citations); differences: the conic is recomputed
// from the parameters (the forward's conic lives in the other rank's private state) and the visible / clamped
// flags travel in the accumulator row's spare slot (pack_flags_kernel).
#include "common.cuh"
#include "grad_chain.cuh"

#define BP_THREADS 64

namespace {

using namespace gchain;  // grad_chain.cuh: the pullbacks shared with backward_gaussians.cu

// flags word in accumulator slot GSR_ACC_FLAGS_SLOT: bit 3 = visible (radii > 0), bits 0..2 = clamped rgb
__global__ void __launch_bounds__(256)
pack_flags_kernel(const int64_t n, const int AF, const int32_t *__restrict__ radii, const uint8_t *__restrict__ clamped,
                  float *__restrict__ gacc) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t f = 0;
    if (radii[i] > 0) f = 8u | (clamped[3 * i] ? 1u : 0u) | (clamped[3 * i + 1] ? 2u : 0u) | (clamped[3 * i + 2] ? 4u : 0u);
    gacc[i * (int64_t)AF + GSR_ACC_FLAGS_SLOT] = __uint_as_float(f);
}

// accumulator rows -> exchange rows (common.cuh): one thread per Gaussian, coalesced enough for a 100-MB one-off pass
__global__ void __launch_bounds__(256)
export_rows_kernel(const int64_t n, const int channels, const float *__restrict__ gacc, float *__restrict__ rows) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const AccRow a = load_acc_row(gacc + i * (int64_t)acc_floats(channels), channels);
    float4 *dst = reinterpret_cast<float4 *>(rows + i * (int64_t)exchange_floats(channels));
    dst[0] = make_float4(a.sx, a.sy, a.sxx, a.sxy);
    dst[1] = make_float4(a.syy, a.se, a.f[0], a.f[1]);
    if (channels > 5) {
        dst[2] = make_float4(a.f[2], a.f[3], 0.f, 0.f);
        dst[3] = make_float4(a.f[5], a.f[6], a.f[7], __uint_as_float(a.flags));
    } else {
        dst[2] = make_float4(a.f[2], a.f[3], 0.f, __uint_as_float(a.flags));
    }
}

// rast.gstate.∇means_2d of the LOCAL view for all Gaussians (strategy.jl:85-86 reads it): conic * (Sx, Sy)
__global__ void __launch_bounds__(256)
grad_means2d_kernel(const int64_t n, const int AF, const int32_t *__restrict__ radii, const float *__restrict__ conics,
                    const float *__restrict__ gacc, float2 *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float2 v = make_float2(0.f, 0.f);
    if (radii[i] > 0) {
        const float2 sxy = *reinterpret_cast<const float2 *>(gacc + i * (int64_t)AF);
        const float sx = sxy.x, sy = sxy.y;
        const float ca = conics[3 * i], cb = conics[3 * i + 1], cc = conics[3 * i + 2];
        v = make_float2(ca * sx + cb * sy, cb * sx + cc * sy);
    }
    out[i] = v;
}

__global__ void __launch_bounds__(BP_THREADS)
backward_gaussians_peers_kernel(const PeerArgs A) {
    extern __shared__ float smem[];  // [BP_THREADS][stride] SH coefficients | [BP_THREADS][stride] SH gradient sums
    const int tid = threadIdx.x;
    const int stride = A.sh_stride;  // (A.world = ranks holding a table; A.n_views = accumulators to sum)
    float *s_sh = smem, *s_vsh = smem + BP_THREADS * stride;
    const int64_t block0 = A.lo + (int64_t)blockIdx.x * BP_THREADS;
    const int64_t i = block0 + tid;
    const int row = 3 * A.K;
    const int k_used = (A.sh_degree + 1) * (A.sh_degree + 1);
    const int64_t nb = (A.hi - block0) < BP_THREADS ? (A.hi - block0) : BP_THREADS;
    const int64_t span = nb * row;
    const bool in = i < A.hi;
    const int AF = acc_floats(A.channels);
    const int64_t n = A.n;

    // SH coefficients of the CTA's Gaussians (replicated parameter): coalesced span copy into padded rows
    if (A.sh_degree > 0) {
        const float *src = A.shs + block0 * row;
        if (k_used == A.K) {
            rows_global_to_shared(src, s_sh, (int)nb, row, stride, tid, BP_THREADS, (reinterpret_cast<uintptr_t>(src) & 15) == 0);
        } else if (in) {
            const float *s1 = A.shs + i * (int64_t)row;
            for (int e = 0; e < 3 * k_used; e++) s_sh[tid * stride + e] = s1[e];
        }
    }
    for (int e = 0; e < row; e++) s_vsh[tid * stride + e] = 0.f;
    __syncthreads();

    float vmean[3] = {0.f, 0.f, 0.f}, vscale[3] = {0.f, 0.f, 0.f}, vq[4] = {0.f, 0.f, 0.f, 0.f}, vop = 0.f;
    if (in) {
        const float mean[3] = {A.means[3 * i], A.means[3 * i + 1], A.means[3 * i + 2]};
        const float sc[3] = {A.scales[3 * i], A.scales[3 * i + 1], A.scales[3 * i + 2]};
        const float4 q4 = *reinterpret_cast<const float4 *>(A.rots + 4 * i);
        const float op = A.opac[i];
        // view-independent: unnorm_quat2rot (render.jl:322-333), M = R_g diag(s), Σ = M M' (render.jl:291-294)
        float Rg[9], Mm[9], Sg[9];
        const Quat q = quat_to_rot(q4, Rg);
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int r = 0; r < 3; r++) M3(Mm, r, j) = M3(Rg, r, j) * sc[j];
        mul33_nt(Mm, Mm, Sg);
        const int kk = thinnest_axis(sc);
        float vRr[9];  // Σ_views cotangent of R_g (linear), converted to the quaternion gradient once
#pragma unroll
        for (int k = 0; k < 9; k++) vRr[k] = 0.f;
        const float *sh = s_sh + tid * stride;
        float *vsh = s_vsh + tid * stride;

        for (int v = 0; v < A.n_views; v++) {
            // ---- this view's accumulator row, straight from the memory of the rank that rendered it (P2P load) ----
            const AccRow a = A.exchange_rows ? load_exchange_row(A.gacc[v] + i * (int64_t)exchange_floats(A.channels), A.channels)
                                             : load_acc_row(A.gacc[v] + i * (int64_t)AF, A.channels);
            const uint32_t flags = a.flags;
            if (!(flags & 8u)) continue;  // culled in view v (projection.jl:172-176)
            const PeerCamera &cam = A.cams[v];
            const float *R = cam.R, *t = cam.t;
            const float *vcol = a.f;

            float mc[3], Sc[9], T1[9];
#pragma unroll
            for (int r = 0; r < 3; r++) mc[r] = R[r] * mean[0] + R[r + 3] * mean[1] + R[r + 6] * mean[2] + t[r];
            mul33(R, Sg, T1);
            mul33_nt(T1, R, Sc);
            const Persp P = persp_setup(cam.focal, cam.principal, cam.width, cam.height, mc);
            // conic of view v, recomputed (the forward's conic lives in the rendering rank's private state)
            float ca, cb, cc;
            conic_from_cov(P, Sc, cam.blur_eps, ca, cb, cc);
            // moments -> cotangents (render.jl:264-273)
            const float vm2[2] = {ca * a.sx + cb * a.sy, cb * a.sx + cc * a.sy};
            const float vcn[3] = {0.5f * a.sxx, 0.5f * a.sxy, 0.5f * a.syy};
            vop += op > 0.0f ? a.se / op : 0.0f;
            float vS2[4], vSc[9], vmc[3];
            grad_inverse2(ca, cb, cc, vcn, vS2);
            grad_perspective(P, mc, Sc, vS2, vm2, vSc, vmc);
            if (A.channels > 3) vmc[2] += vcol[3];  // vdepth (projection.jl:218-222)
            // ∇pos_world_to_cam, ∇covar_world_to_cam (projection.jl:363-393)
#pragma unroll
            for (int r = 0; r < 3; r++) vmean[r] += M3(R, 0, r) * vmc[0] + M3(R, 1, r) * vmc[1] + M3(R, 2, r) * vmc[2];
            float vSg[9];
            mul33_tn(R, vSc, T1);
            mul33(T1, R, vSg);
            // ∇quat_scale_to_cov (render.jl:302-320): accumulate the cotangent of R_g and of the scales
            float vM[9], sym[9];
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int r = 0; r < 3; r++) M3(sym, r, j) = M3(vSg, r, j) + M3(vSg, j, r);
            mul33(sym, Mm, vM);
#pragma unroll
            for (int j = 0; j < 3; j++) {
#pragma unroll
                for (int r = 0; r < 3; r++) M3(vRr, r, j) += M3(vM, r, j) * sc[j];
                vscale[j] += M3(Rg, 0, j) * M3(vM, 0, j) + M3(Rg, 1, j) * M3(vM, 1, j) + M3(Rg, 2, j) * M3(vM, 2, j);
            }
            if (A.channels > 5) {  // normal channel (projection.jl:227-236)
                const float vn[3] = {vcol[5], vcol[6], vcol[7]};
                float gr[3];
                grad_normal(R, Rg, mc, kk, vn, gr);
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    if (kk == 0) M3(vRr, r, 0) += gr[r];
                    else if (kk == 1) M3(vRr, r, 1) += gr[r];
                    else M3(vRr, r, 2) += gr[r];
                }
            }
            // ∇color_from_sh! + ∇normalize (spherical_harmonics.jl:76-181)
            {
                float vc[3], basis[16];
#pragma unroll
                for (int c = 0; c < 3; c++) vc[c] = (flags >> c) & 1u ? 0.f : vcol[c];
                grad_sh(A.sh_degree, mean, cam.cam_center, sh, vc, basis, vmean);
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    if (k < k_used) {
#pragma unroll
                        for (int c = 0; c < 3; c++) vsh[3 * k + c] += basis[k] * vc[c];
                    }
                }
            }
        }
        // ∇unnorm_quat2rot (render.jl:335-366), once on the summed cotangent of R_g
        grad_quat(vRr, q, vq);

        // ---- store the reduced row into EVERY rank's table (P2P stores) ---------------------------------------
        for (int p = 0; p < A.world; p++) {
            float *tb = A.table[p];
            if (!tb) continue;  // a rank without a table pointer does not receive the rows (reduce-scatter only)
            *reinterpret_cast<float4 *>(tb + 4 * i) = make_float4(vq[0], vq[1], vq[2], vq[3]);
            float *pm = tb + 4 * n + 3 * i, *ps = tb + 7 * n + 3 * i;
            pm[0] = vmean[0]; pm[1] = vmean[1]; pm[2] = vmean[2];
            ps[0] = vscale[0]; ps[1] = vscale[1]; ps[2] = vscale[2];
            tb[10 * n + i] = vop;
        }
    }
    __syncthreads();
    // SH-gradient span of the CTA: coalesced 128-bit stores to every rank
    {
        const int64_t nq = A.vsh_aligned ? (span >> 2) : 0;  // block0 * row * 4 B is 16-byte aligned (lo % BP_THREADS == 0)
        for (int64_t q = tid; q < nq; q += BP_THREADS) {
            const int e = (int)(q << 2);
            const int gq = e / row, rq = e - gq * row;
            float vv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                int gg = gq, rr = rq + u;
                if (rr >= row) { rr -= row; gg += 1; }
                vv[u] = s_vsh[gg * stride + rr];
            }
            const float4 val = make_float4(vv[0], vv[1], vv[2], vv[3]);
            for (int p = 0; p < A.world; p++)
                if (A.table[p]) *(reinterpret_cast<float4 *>(A.table[p] + 11 * n + block0 * row) + q) = val;
        }
        for (int64_t e = (nq << 2) + tid; e < span; e += BP_THREADS) {
            const int gg = (int)(e / row), rr = (int)(e - (int64_t)gg * row);
            for (int p = 0; p < A.world; p++)
                if (A.table[p]) A.table[p][11 * n + block0 * row + e] = s_vsh[gg * stride + rr];
        }
    }
}

}  // namespace

void launch_pack_flags(int64_t n, int channels, const int32_t *radii, const uint8_t *clamped, float *gacc, cudaStream_t s) {
    if (n <= 0) return;
    pack_flags_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, acc_floats(channels), radii, clamped, gacc);
    count_launch();
}

void launch_export_rows(int64_t n, int channels, const float *gacc, float *rows, cudaStream_t s) {
    if (n <= 0) return;
    export_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, channels, gacc, rows);
    count_launch();
}

void launch_grad_means2d(int64_t n, int channels, const int32_t *radii, const float *conics, const float *gacc,
                         float2 *out, cudaStream_t s) {
    if (n <= 0) return;
    grad_means2d_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, acc_floats(channels), radii, conics, gacc, out);
    count_launch();
}

int launch_backward_gaussians_peers(const PeerArgs &args, cudaStream_t s) {
    const int64_t cnt = args.hi - args.lo;
    if (cnt <= 0) return 0;
    const size_t smem = 2 * (size_t)BP_THREADS * args.sh_stride * sizeof(float);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(backward_gaussians_peers_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        attr = true;
    }
    backward_gaussians_peers_kernel<<<(unsigned)((cnt + BP_THREADS - 1) / BP_THREADS), BP_THREADS, smem, s>>>(args);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
