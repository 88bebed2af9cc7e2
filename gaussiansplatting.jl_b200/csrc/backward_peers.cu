// backward_peers.cu — per-Gaussian backward FUSED with the cross-GPU gradient reduction (NVLink peer memory).
//
// Multi-GPU semantics: batch gradient = sum over views of ∇rasterize; views are sharded over
// ranks, parameters replicated.  The baseline runs backward_gaussians_kernel on every rank over all N Gaussians
// and then ncclAllReduce()s the 59-float-per-Gaussian table (236 B/Gaussian, 2(G-1)/G x 236 MB per link at 1M).
//
// This kernel does both steps at once over peer-mapped (symmetric) memory:
//   * rank r owns the Gaussian slice S_r = [lo, hi);
//   * for every Gaussian of its slice it LOADS the 48/64-byte moment accumulator of EVERY rank's view straight
//     from that rank's HBM over NVLink (P2P loads: (G-1)/G x 48 B/Gaussian instead of 236 B), applies that view's
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

#define BP_THREADS 64

namespace {

#define M3(m, i, j) ((m)[(i) + 3 * (j)])
__device__ __forceinline__ void mul33(const float *A, const float *B, float *C) {
    float T[9];
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++) T[i + 3 * j] = M3(A, i, 0) * M3(B, 0, j) + M3(A, i, 1) * M3(B, 1, j) + M3(A, i, 2) * M3(B, 2, j);
#pragma unroll
    for (int k = 0; k < 9; k++) C[k] = T[k];
}
__device__ __forceinline__ void mul33_tn(const float *A, const float *B, float *C) {
    float T[9];
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++) T[i + 3 * j] = M3(A, 0, i) * M3(B, 0, j) + M3(A, 1, i) * M3(B, 1, j) + M3(A, 2, i) * M3(B, 2, j);
#pragma unroll
    for (int k = 0; k < 9; k++) C[k] = T[k];
}
__device__ __forceinline__ void mul33_nt(const float *A, const float *B, float *C) {
    float T[9];
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++) T[i + 3 * j] = M3(A, i, 0) * M3(B, j, 0) + M3(A, i, 1) * M3(B, j, 1) + M3(A, i, 2) * M3(B, j, 2);
#pragma unroll
    for (int k = 0; k < 9; k++) C[k] = T[k];
}

#define SH0 0.28209479177387814f
#define SH1 0.4886025119029199f
#define SH2C1 1.0925484305920792f
#define SH2C2 -1.0925484305920792f
#define SH2C3 0.31539156525252005f
#define SH2C4 -1.0925484305920792f
#define SH2C5 0.5462742152960396f
#define SH3C1 -0.5900435899266435f
#define SH3C2 2.890611442640554f
#define SH3C3 -0.4570457994644658f
#define SH3C4 0.3731763325901154f
#define SH3C5 -0.4570457994644658f
#define SH3C6 1.445305721320277f
#define SH3C7 -0.5900435899266435f

// flags word in accumulator slot AF-1: bit 3 = visible (radii > 0), bits 0..2 = clamped rgb
__global__ void __launch_bounds__(256)
pack_flags_kernel(const int64_t n, const int AF, const int32_t *__restrict__ radii, const uint8_t *__restrict__ clamped,
                  float *__restrict__ gacc) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t f = 0;
    if (radii[i] > 0) f = 8u | (clamped[3 * i] ? 1u : 0u) | (clamped[3 * i + 1] ? 2u : 0u) | (clamped[3 * i + 2] ? 4u : 0u);
    gacc[i * (int64_t)AF + AF - 1] = __uint_as_float(f);
}

// rast.gstate.∇means_2d of the LOCAL view for all Gaussians (strategy.jl:85-86 reads it): conic * (Sx, Sy)
__global__ void __launch_bounds__(256)
grad_means2d_kernel(const int64_t n, const int AF, const int32_t *__restrict__ radii, const float *__restrict__ conics,
                    const float *__restrict__ gacc, float2 *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float2 v = make_float2(0.f, 0.f);
    if (radii[i] > 0) {
        const float sx = gacc[i * (int64_t)AF], sy = gacc[i * (int64_t)AF + 1];
        const float ca = conics[3 * i], cb = conics[3 * i + 1], cc = conics[3 * i + 2];
        v = make_float2(ca * sx + cb * sy, cb * sx + cc * sy);
    }
    out[i] = v;
}

__global__ void __launch_bounds__(BP_THREADS)
backward_gaussians_peers_kernel(const PeerArgs A) {
    extern __shared__ float smem[];  // [BP_THREADS][stride] SH coefficients | [BP_THREADS][stride] SH gradient sums
    const int tid = threadIdx.x;
    const int stride = A.sh_stride;
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
        const float qn = sqrtf(q4.x * q4.x + q4.y * q4.y + q4.z * q4.z + q4.w * q4.w);
        const float qi = 1.0f / qn;
        const float w = qi * q4.x, x = qi * q4.y, y = qi * q4.z, z = qi * q4.w;
        float Rg[9], Mm[9], Sg[9];
        Rg[0] = 1.0f - 2.0f * (y * y + z * z); Rg[1] = 2.0f * (x * y + w * z); Rg[2] = 2.0f * (x * z - w * y);
        Rg[3] = 2.0f * (x * y - w * z); Rg[4] = 1.0f - 2.0f * (x * x + z * z); Rg[5] = 2.0f * (y * z + w * x);
        Rg[6] = 2.0f * (x * z + w * y); Rg[7] = 2.0f * (y * z - w * x); Rg[8] = 1.0f - 2.0f * (x * x + y * y);
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int r = 0; r < 3; r++) M3(Mm, r, j) = M3(Rg, r, j) * sc[j];
        mul33_nt(Mm, Mm, Sg);
        const int kk = (sc[0] <= sc[1] && sc[0] <= sc[2]) ? 0 : ((sc[1] <= sc[2]) ? 1 : 2);
        float vRr[9];  // Σ_views cotangent of R_g (linear), converted to the quaternion gradient once
#pragma unroll
        for (int k = 0; k < 9; k++) vRr[k] = 0.f;
        const float *sh = s_sh + tid * stride;
        float *vsh = s_vsh + tid * stride;

        for (int v = 0; v < A.world; v++) {
            // ---- this view's accumulator row, straight from rank v's memory (P2P load when v != rank) ----------
            const float *acc = A.gacc[v] + i * (int64_t)AF;
            const float4 a0 = *reinterpret_cast<const float4 *>(acc);
            const float4 a1 = *reinterpret_cast<const float4 *>(acc + 4);
            const float4 a2 = *reinterpret_cast<const float4 *>(acc + 8);
            float4 a3 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (AF > 12) a3 = *reinterpret_cast<const float4 *>(acc + 12);
            const uint32_t flags = __float_as_uint(AF > 12 ? a3.w : a2.w);
            if (!(flags & 8u)) continue;  // culled in view v (projection.jl:172-176)
            const PeerCamera &cam = A.cams[v];
            const float *R = cam.R, *t = cam.t;
            float vcol[8] = {a1.z, a1.w, a2.x, a2.y, 0.f, a2.w, a3.x, a3.y};
            if (A.channels <= 5) vcol[5] = 0.f;  // slot 11 holds the flags when AF == 12

            float mc[3], Sc[9], T1[9];
#pragma unroll
            for (int r = 0; r < 3; r++) mc[r] = R[r] * mean[0] + R[r + 3] * mean[1] + R[r + 6] * mean[2] + t[r];
            mul33(R, Sg, T1);
            mul33_nt(T1, R, Sc);
            const float res[2] = {(float)cam.width, (float)cam.height};
            float lim[2], limn[2], txy[2];
            const float rz = 1.0f / mc[2];
            const float rz2 = rz * rz, rz3 = rz2 * rz;
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const float stf = 0.3f * ((0.5f * res[k]) / cam.focal[k]);
                const float pp = cam.principal[k] * res[k];
                lim[k] = (res[k] - pp) / cam.focal[k] + stf;
                limn[k] = pp / cam.focal[k] + stf;
                txy[k] = mc[2] * fminf(lim[k], fmaxf(-limn[k], mc[k] * rz));
            }
            const float fx = cam.focal[0], fy = cam.focal[1];
            const float J[6] = {fx * rz, 0.f, 0.f, fy * rz, -fx * txy[0] * rz2, -fy * txy[1] * rz2};
#define J_(r, c) J[(r) + 2 * (c)]
            // conic of view v, recomputed: Σ2D = J Σcam J' + blur, inverse (projection.jl:259-287, render.jl:368-396)
            float ca, cb, cc;
            {
                float TJ[6];
#pragma unroll
                for (int j = 0; j < 3; j++)
#pragma unroll
                    for (int r = 0; r < 2; r++) TJ[r + 2 * j] = J_(r, 0) * M3(Sc, 0, j) + J_(r, 1) * M3(Sc, 1, j) + J_(r, 2) * M3(Sc, 2, j);
                const float s00 = TJ[0] * J_(0, 0) + TJ[2] * J_(0, 1) + TJ[4] * J_(0, 2) + cam.blur_eps;
                const float s10 = TJ[1] * J_(0, 0) + TJ[3] * J_(0, 1) + TJ[5] * J_(0, 2);
                const float s01 = TJ[0] * J_(1, 0) + TJ[2] * J_(1, 1) + TJ[4] * J_(1, 2);
                const float s11 = TJ[1] * J_(1, 0) + TJ[3] * J_(1, 1) + TJ[5] * J_(1, 2) + cam.blur_eps;
                const float det_inv = 1.0f / (s00 * s11 - s01 * s10);
                ca = s11 * det_inv; cb = -s01 * det_inv; cc = s00 * det_inv;
            }
            // moments -> cotangents (render.jl:264-273)
            const float vm2[2] = {ca * a0.x + cb * a0.y, cb * a0.x + cc * a0.y};
            const float vcn[3] = {0.5f * a0.z, 0.5f * a0.w, 0.5f * a1.x};
            vop += op > 0.0f ? a1.y / op : 0.0f;
            // ∇inverse (render.jl:383-385)
            float vS2[4];
            {
                const float X[4] = {ca, cb, cb, cc}, V[4] = {vcn[0], vcn[1], vcn[1], vcn[2]};
                float Tm[4];
#pragma unroll
                for (int j = 0; j < 2; j++)
#pragma unroll
                    for (int r = 0; r < 2; r++) Tm[r + 2 * j] = -(X[r] * V[2 * j] + X[r + 2] * V[1 + 2 * j]);
#pragma unroll
                for (int j = 0; j < 2; j++)
#pragma unroll
                    for (int r = 0; r < 2; r++) vS2[r + 2 * j] = Tm[r] * X[2 * j] + Tm[r + 2] * X[1 + 2 * j];
            }
            // ∇perspective_projection (projection.jl:289-353)
            float vSc[9], vmc[3];
            {
#define V2(r, c) vS2[(r) + 2 * (c)]
                float Aq[6];
#pragma unroll
                for (int j = 0; j < 2; j++)
#pragma unroll
                    for (int r = 0; r < 3; r++) Aq[r + 3 * j] = J_(0, r) * V2(0, j) + J_(1, r) * V2(1, j);
#pragma unroll
                for (int j = 0; j < 3; j++)
#pragma unroll
                    for (int r = 0; r < 3; r++) M3(vSc, r, j) = Aq[r] * J_(0, j) + Aq[r + 3] * J_(1, j);
                float B1[6], B2[6], vJ[6];
#pragma unroll
                for (int j = 0; j < 3; j++)
#pragma unroll
                    for (int r = 0; r < 2; r++) {
                        B1[r + 2 * j] = V2(r, 0) * J_(0, j) + V2(r, 1) * J_(1, j);
                        B2[r + 2 * j] = V2(0, r) * J_(0, j) + V2(1, r) * J_(1, j);
                    }
#pragma unroll
                for (int j = 0; j < 3; j++)
#pragma unroll
                    for (int r = 0; r < 2; r++)
                        vJ[r + 2 * j] = (B1[r] * M3(Sc, j, 0) + B1[r + 2] * M3(Sc, j, 1) + B1[r + 4] * M3(Sc, j, 2)) +
                                        (B2[r] * M3(Sc, 0, j) + B2[r + 2] * M3(Sc, 1, j) + B2[r + 4] * M3(Sc, 2, j));
#undef V2
#define VJ(r, c) vJ[((r) - 1) + 2 * ((c) - 1)]
                float vx = fx * rz * vm2[0];
                float vy = fy * rz * vm2[1];
                float vz = -rz2 * (fx * mc[0] * vm2[0] + fy * mc[1] * vm2[1]);
                const float ax = mc[0] * rz, ay = mc[1] * rz;
                if (-limn[0] <= ax && ax <= lim[0]) vx += -fx * rz2 * VJ(1, 3);
                else vz += -fx * rz3 * VJ(1, 3) * txy[0];
                if (-limn[1] <= ay && ay <= lim[1]) vy += -fy * rz2 * VJ(2, 3);
                else vz += -fy * rz3 * VJ(2, 3) * txy[1];
                vz += -fx * rz2 * VJ(1, 1) - fy * rz2 * VJ(2, 2) + 2.0f * fx * txy[0] * rz3 * VJ(1, 3) +
                      2.0f * fy * txy[1] * rz3 * VJ(2, 3);
#undef VJ
                vmc[0] = vx; vmc[1] = vy; vmc[2] = vz;
            }
#undef J_
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
                const float axs[3] = {kk == 0 ? Rg[0] : (kk == 1 ? Rg[3] : Rg[6]), kk == 0 ? Rg[1] : (kk == 1 ? Rg[4] : Rg[7]),
                                      kk == 0 ? Rg[2] : (kk == 1 ? Rg[5] : Rg[8])};
                float nc[3];
#pragma unroll
                for (int r = 0; r < 3; r++) nc[r] = R[r] * axs[0] + R[r + 3] * axs[1] + R[r + 6] * axs[2];
                const float sign = (nc[0] * mc[0] + nc[1] * mc[1] + nc[2] * mc[2]) > 0.f ? -1.f : 1.f;
                const float vn[3] = {vcol[5], vcol[6], vcol[7]};
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    const float gr = sign * (M3(R, 0, r) * vn[0] + M3(R, 1, r) * vn[1] + M3(R, 2, r) * vn[2]);
                    if (kk == 0) M3(vRr, r, 0) += gr;
                    else if (kk == 1) M3(vRr, r, 1) += gr;
                    else M3(vRr, r, 2) += gr;
                }
            }
            // ∇color_from_sh! + ∇normalize (spherical_harmonics.jl:76-181)
            {
                const float d0 = mean[0] - cam.cam_center[0], d1 = mean[1] - cam.cam_center[1], d2 = mean[2] - cam.cam_center[2];
                const float s2 = d0 * d0 + d1 * d1 + d2 * d2;
                const float inv = 1.0f / sqrtf(s2);
                const float X = inv * d0, Y = inv * d1, Z = inv * d2;
                float vc[3];
#pragma unroll
                for (int c = 0; c < 3; c++) vc[c] = (flags >> c) & 1u ? 0.f : vcol[c];
                const float x2 = X * X, y2 = Y * Y, z2 = Z * Z, xy = X * Y, xz = X * Z, yz = Y * Z;
                float vdir[3] = {0.f, 0.f, 0.f};
                float basis[16];
                basis[0] = SH0;
                if (A.sh_degree > 0) {
                    basis[1] = -SH1 * Y; basis[2] = SH1 * Z; basis[3] = -SH1 * X;
                    if (A.sh_degree > 1) {
                        basis[4] = SH2C1 * xy; basis[5] = SH2C2 * yz; basis[6] = SH2C3 * (2.0f * z2 - x2 - y2);
                        basis[7] = SH2C4 * xz; basis[8] = SH2C5 * (x2 - y2);
                        if (A.sh_degree > 2) {
                            basis[9] = SH3C1 * Y * (3.0f * x2 - y2); basis[10] = SH3C2 * xy * Z;
                            basis[11] = SH3C3 * Y * (4.0f * z2 - x2 - y2);
                            basis[12] = SH3C4 * Z * (2.0f * z2 - 3.0f * x2 - 3.0f * y2);
                            basis[13] = SH3C5 * X * (4.0f * z2 - x2 - y2); basis[14] = SH3C6 * Z * (x2 - y2);
                            basis[15] = SH3C7 * X * (x2 - 3.0f * y2);
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < 3; c++) {
#define S(k) sh[3 * ((k) - 1) + c]
                    float gx = 0.f, gy = 0.f, gz = 0.f;
                    if (A.sh_degree > 0) {
                        gx = -SH1 * S(4); gy = -SH1 * S(2); gz = SH1 * S(3);
                        if (A.sh_degree > 1) {
                            gx += SH2C1 * Y * S(5) + SH2C3 * 2.0f * -X * S(7) + SH2C4 * Z * S(8) + SH2C5 * 2.0f * X * S(9);
                            gy += SH2C1 * X * S(5) + SH2C2 * Z * S(6) + SH2C3 * 2.0f * -Y * S(7) + SH2C5 * 2.0f * -Y * S(9);
                            gz += SH2C2 * Y * S(6) + SH2C3 * 4.0f * Z * S(7) + SH2C4 * X * S(8);
                            if (A.sh_degree > 2) {
                                gx += SH3C1 * S(10) * 6.0f * xy + SH3C2 * S(11) * yz + SH3C3 * S(12) * -2.0f * xy +
                                      SH3C4 * S(13) * -6.0f * xz + SH3C5 * S(14) * (-3.0f * x2 + 4.0f * z2 - y2) +
                                      SH3C6 * S(15) * 2.0f * xz + SH3C7 * S(16) * 3.0f * (x2 - y2);
                                gy += SH3C1 * S(10) * 3.0f * (x2 - y2) + SH3C2 * S(11) * xz +
                                      SH3C3 * S(12) * (-3.0f * y2 + 4.0f * z2 - x2) + SH3C4 * S(13) * -6.0f * yz +
                                      SH3C5 * S(14) * -2.0f * xy + SH3C6 * S(15) * -2.0f * yz + SH3C7 * S(16) * -6.0f * xy;
                                gz += SH3C2 * S(11) * xy + SH3C3 * S(12) * 8.0f * yz +
                                      SH3C4 * S(13) * 3.0f * (2.0f * z2 - x2 - y2) + SH3C5 * S(14) * 8.0f * xz +
                                      SH3C6 * S(15) * (x2 - y2);
                            }
                        }
                    }
#undef S
                    vdir[0] += gx * vc[c]; vdir[1] += gy * vc[c]; vdir[2] += gz * vc[c];
                }
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    if (k < k_used) {
#pragma unroll
                        for (int c = 0; c < 3; c++) vsh[3 * k + c] += basis[k] * vc[c];
                    }
                }
                const float inv_s = 1.0f / sqrtf(s2 * s2 * s2);
                vmean[0] += ((s2 - d0 * d0) * vdir[0] - d1 * d0 * vdir[1] - d2 * d0 * vdir[2]) * inv_s;
                vmean[1] += (-d0 * d1 * vdir[0] + (s2 - d1 * d1) * vdir[1] - d2 * d1 * vdir[2]) * inv_s;
                vmean[2] += (-d0 * d2 * vdir[0] - d1 * d2 * vdir[1] + (s2 - d2 * d2) * vdir[2]) * inv_s;
            }
        }
        // ∇unnorm_quat2rot (render.jl:335-366), once on the summed cotangent of R_g
#define V(r, c) M3(vRr, (r) - 1, (c) - 1)
        float vqn[4];
        vqn[0] = 2.0f * (x * (V(3, 2) - V(2, 3)) + y * (V(1, 3) - V(3, 1)) + z * (V(2, 1) - V(1, 2)));
        vqn[1] = 2.0f * (-2.0f * x * (V(2, 2) + V(3, 3)) + y * (V(2, 1) + V(1, 2)) + z * (V(3, 1) + V(1, 3)) + w * (V(3, 2) - V(2, 3)));
        vqn[2] = 2.0f * (x * (V(2, 1) + V(1, 2)) - 2.0f * y * (V(1, 1) + V(3, 3)) + z * (V(3, 2) + V(2, 3)) + w * (V(1, 3) - V(3, 1)));
        vqn[3] = 2.0f * (x * (V(3, 1) + V(1, 3)) + y * (V(3, 2) + V(2, 3)) - 2.0f * z * (V(1, 1) + V(2, 2)) + w * (V(2, 1) - V(1, 2)));
#undef V
        const float qd = vqn[0] * w + vqn[1] * x + vqn[2] * y + vqn[3] * z;
        vq[0] = (vqn[0] - qd * w) * qi; vq[1] = (vqn[1] - qd * x) * qi; vq[2] = (vqn[2] - qd * y) * qi; vq[3] = (vqn[3] - qd * z) * qi;

        // ---- store the reduced row into EVERY rank's table (P2P stores) ---------------------------------------
        for (int p = 0; p < A.world; p++) {
            float *tb = A.table[p];
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
                *(reinterpret_cast<float4 *>(A.table[p] + 11 * n + block0 * row) + q) = val;
        }
        for (int64_t e = (nq << 2) + tid; e < span; e += BP_THREADS) {
            const int gg = (int)(e / row), rr = (int)(e - (int64_t)gg * row);
            for (int p = 0; p < A.world; p++) A.table[p][11 * n + block0 * row + e] = s_vsh[gg * stride + rr];
        }
    }
}

}  // namespace

void launch_pack_flags(int64_t n, int channels, const int32_t *radii, const uint8_t *clamped, float *gacc, cudaStream_t s) {
    if (n <= 0) return;
    pack_flags_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, acc_floats(channels), radii, clamped, gacc);
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
