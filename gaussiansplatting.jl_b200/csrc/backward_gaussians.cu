// backward_gaussians.cu — fused per-Gaussian backward stage for sm_100a.
//
// One kernel replaces ∇project! (src/rasterization/projection.jl:132-257), ∇spherical_harmonics!
// (spherical_harmonics.jl:20-38), the zero-fills of ∇rasterize (rasterizer.jl:437-446) and the slicing of
// vcolor_features into vrgbs / vdepths / vnormals (rasterizer.jl:486-493): every output row is written exactly
// once (zeros for culled Gaussians, which the reference obtains from KA.zeros + early return), so no memset
// pass over the 59-float-per-Gaussian gradient table is needed.  `accumulate` adds instead (view batches).
// Also publishes rast.gstate.∇means_2d (pixel units) and vopacities from the packed accumulator.
//
// Gradients are tolerance-checked (1e-4 relative), so FMA contraction is allowed here.
#include "common.cuh"

#define BG_THREADS 128

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
__device__ __forceinline__ void mul33_tn(const float *A, const float *B, float *C) {  // C = A' * B
    float T[9];
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++) T[i + 3 * j] = M3(A, 0, i) * M3(B, 0, j) + M3(A, 1, i) * M3(B, 1, j) + M3(A, 2, i) * M3(B, 2, j);
#pragma unroll
    for (int k = 0; k < 9; k++) C[k] = T[k];
}
__device__ __forceinline__ void mul33_nt(const float *A, const float *B, float *C) {  // C = A * B'
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

template <bool ACC>
__device__ __forceinline__ void put(float *p, float v) {
    if (ACC) *p += v; else *p = v;
}

template <bool ACC, bool RAW>
__global__ void __launch_bounds__(BG_THREADS, 6)
backward_gaussians_kernel(const DevCamera cam, const int64_t n, const int sh_degree, const int K, const int channels,
                          const float *__restrict__ means, const float *__restrict__ shs,
                          const float *__restrict__ opac, const float *__restrict__ scales, const float *__restrict__ rots, const GeomPtrs g,
                          float *__restrict__ vmeans, float *__restrict__ vshs, float *__restrict__ vopac,
                          float *__restrict__ vscales, float *__restrict__ vrot, float *vR_out, float *vt_out,
                          const int sh_stride, const ParamSpec ps) {
    // SH coefficients in / SH gradients out are staged through shared memory: the (3,K,n) rows of a CTA's
    // 128 Gaussians form one contiguous span that is read and written with coalesced 128-bit accesses, while each
    // thread works on its own padded (odd stride -> conflict-free) row.
    extern __shared__ float s_sh[];  // [BG_THREADS][sh_stride]
    const int tid = threadIdx.x;
    const int64_t block0 = (int64_t)blockIdx.x * BG_THREADS;
    const int64_t i = block0 + tid;
    const bool pose = vR_out != nullptr;
    float pose_acc[12];
#pragma unroll
    for (int k = 0; k < 12; k++) pose_acc[k] = 0.f;
    const int row = 3 * K;
    const int k_used = (sh_degree + 1) * (sh_degree + 1);
    const int64_t nb = (n - block0) < BG_THREADS ? (n - block0) : BG_THREADS;
    const int64_t span = nb * row;
    const bool visible_t = (i < n) && g.radii[i] > 0;
    const bool aligned16 = (((uintptr_t)shs | (uintptr_t)vshs) & 15) == 0;
    if (__syncthreads_or(visible_t) && sh_degree > 0) {
        if (k_used == K) {
            if (RAW && ps.sh_rest) {
                rows_global_to_shared(shs + block0 * 3, s_sh, (int)nb, 3, sh_stride, tid, BG_THREADS, false);
                if (K > 1)
                    rows_global_to_shared(ps.sh_rest + block0 * (int64_t)(row - 3), s_sh + 3, (int)nb, row - 3, sh_stride, tid,
                                          BG_THREADS, (reinterpret_cast<uintptr_t>(ps.sh_rest) & 15) == 0);
            } else {
                rows_global_to_shared(shs + block0 * row, s_sh, (int)nb, row, sh_stride, tid, BG_THREADS, aligned16);
            }
        } else if (visible_t) {
            if (RAW && ps.sh_rest) {
                for (int e = 0; e < 3; e++) s_sh[tid * sh_stride + e] = shs[3 * i + e];
                const float *src = ps.sh_rest + i * (int64_t)(row - 3);
                for (int e = 3; e < 3 * k_used; e++) s_sh[tid * sh_stride + e] = src[e - 3];
            } else {
                const float *src = shs + i * (int64_t)row;
                for (int e = 0; e < 3 * k_used; e++) s_sh[tid * sh_stride + e] = src[e];
            }
        }
    }
    __syncthreads();

    if (i < n) {
        const int AF = acc_floats(channels);
        const bool visible = visible_t;
        float *vsh = s_sh + tid * sh_stride;  // this thread's row: coefficients in, gradients out (in place)
        if (!visible) {
            // projection.jl:172-176: culled rows keep the zero gradient; ∇spherical_harmonics! runs with a zero
            // colour cotangent and also yields zero.
            g.grad_means2d[i] = make_float2(0.f, 0.f);
            if (!ACC) {
#pragma unroll
                for (int k = 0; k < 3; k++) vmeans[3 * i + k] = 0.f;
                if (RAW && ps.isotropic) vscales[i] = 0.f;
                else
#pragma unroll
                    for (int k = 0; k < 3; k++) vscales[3 * i + k] = 0.f;
                *reinterpret_cast<float4 *>(vrot + 4 * i) = make_float4(0.f, 0.f, 0.f, 0.f);
                vopac[i] = 0.f;
            }
            for (int e = 0; e < row; e++) vsh[e] = 0.f;
        } else {
            const float *acc = g.gacc + i * (int64_t)AF;
            const float4 a0 = *reinterpret_cast<const float4 *>(acc);
            const float4 a1 = *reinterpret_cast<const float4 *>(acc + 4);
            const float4 a2 = *reinterpret_cast<const float4 *>(acc + 8);
            // the compositing backward accumulates moments (render.cu): convert to the reference's cotangents
            //   v_mean2d = conic * (Sx, Sy)             render.jl:269-272
            //   v_conic  = 0.5 * (Sxx, Sxy, Syy)        render.jl:264-268
            //   v_opacity = (sum e*v_alpha) / opacity   render.jl:273  (e = opacity*G)
            const float ca = g.conics[3 * i], cb = g.conics[3 * i + 1], cc = g.conics[3 * i + 2];
            const float vm2[2] = {ca * a0.x + cb * a0.y, cb * a0.x + cc * a0.y};
            const float vcn[3] = {0.5f * a0.z, 0.5f * a0.w, 0.5f * a1.x};
            const float op = (RAW && ps.raw_opacity) ? act_sigmoid(opac[i]) : opac[i];
            float vop = op > 0.0f ? a1.y / op : 0.0f;
            if (RAW && ps.raw_opacity) vop *= op * (1.0f - op);  // pullback of sigmoid (rasterizer.jl:229)
            float vcol[8] = {a1.z, a1.w, a2.x, a2.y, a2.z, a2.w, 0.f, 0.f};
            if (channels > 5) {
                const float4 a3 = *reinterpret_cast<const float4 *>(acc + 12);
                vcol[6] = a3.x; vcol[7] = a3.y;
            }
            g.grad_means2d[i] = make_float2(vm2[0], vm2[1]);
            put<ACC>(vopac + i, vop);

            float R[9], t[3];
            if (cam.R_dev) {
#pragma unroll
                for (int k = 0; k < 9; k++) R[k] = cam.R_dev[k];
#pragma unroll
                for (int k = 0; k < 3; k++) t[k] = cam.t_dev[k];
            } else {
#pragma unroll
                for (int k = 0; k < 9; k++) R[k] = cam.R[k];
#pragma unroll
                for (int k = 0; k < 3; k++) t[k] = cam.t[k];
            }
            const float mean[3] = {means[3 * i], means[3 * i + 1], means[3 * i + 2]};
            float sc[3];
            if (RAW && ps.isotropic) {
                sc[0] = sc[1] = sc[2] = expf(scales[i]);
            } else {
                sc[0] = scales[3 * i]; sc[1] = scales[3 * i + 1]; sc[2] = scales[3 * i + 2];
                if (RAW && ps.raw_scale) { sc[0] = expf(sc[0]); sc[1] = expf(sc[1]); sc[2] = expf(sc[2]); }
            }
            const float4 q4 = *reinterpret_cast<const float4 *>(rots + 4 * i);

            // ∇inverse (render.jl:383-385): vΣ2D = -Σ⁻¹ vΣ⁻¹ Σ⁻¹ with symmetric 2x2 operands (projection.jl:178-188)
            float vS2[4];  // column-major 2x2
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

            // recompute the forward intermediates (projection.jl:202-209)
            float mc[3];
#pragma unroll
            for (int r = 0; r < 3; r++) mc[r] = R[r] * mean[0] + R[r + 3] * mean[1] + R[r + 6] * mean[2] + t[r];
            const float qn = sqrtf(q4.x * q4.x + q4.y * q4.y + q4.z * q4.z + q4.w * q4.w);
            const float qi = 1.0f / qn;
            const float w = qi * q4.x, x = qi * q4.y, y = qi * q4.z, z = qi * q4.w;
            float Rg[9];
            Rg[0] = 1.0f - 2.0f * (y * y + z * z); Rg[1] = 2.0f * (x * y + w * z); Rg[2] = 2.0f * (x * z - w * y);
            Rg[3] = 2.0f * (x * y - w * z); Rg[4] = 1.0f - 2.0f * (x * x + z * z); Rg[5] = 2.0f * (y * z + w * x);
            Rg[6] = 2.0f * (x * z + w * y); Rg[7] = 2.0f * (y * z - w * x); Rg[8] = 1.0f - 2.0f * (x * x + y * y);
            float Mm[9], Sg[9], Sc[9], T1[9];
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int r = 0; r < 3; r++) M3(Mm, r, j) = M3(Rg, r, j) * sc[j];
            mul33_nt(Mm, Mm, Sg);
            mul33(R, Sg, T1);
            mul33_nt(T1, R, Sc);

            // ∇perspective_projection (projection.jl:289-353)
            float vSc[9], vmc[3];
            {
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
#define V2(r, c) vS2[(r) + 2 * (c)]
                float A[6];  // J' * vΣ2D  (3x2)
#pragma unroll
                for (int j = 0; j < 2; j++)
#pragma unroll
                    for (int r = 0; r < 3; r++) A[r + 3 * j] = J_(0, r) * V2(0, j) + J_(1, r) * V2(1, j);
#pragma unroll
                for (int j = 0; j < 3; j++)
#pragma unroll
                    for (int r = 0; r < 3; r++) M3(vSc, r, j) = A[r] * J_(0, j) + A[r + 3] * J_(1, j);
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
#undef J_
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
            if (channels > 3) vmc[2] += vcol[3];  // vdepth (projection.jl:218-222)

            // ∇pos_world_to_cam (projection.jl:363-373), ∇covar_world_to_cam (:382-393)
            float vmean[3];
#pragma unroll
            for (int r = 0; r < 3; r++) vmean[r] = M3(R, 0, r) * vmc[0] + M3(R, 1, r) * vmc[1] + M3(R, 2, r) * vmc[2];
            float vSg[9];
            mul33_tn(R, vSc, T1);
            mul33(T1, R, vSg);
            if (pose) {
                float vRl[9], A1[9], A2[9];
#pragma unroll
                for (int j = 0; j < 3; j++)
#pragma unroll
                    for (int r = 0; r < 3; r++) M3(vRl, r, j) = vmc[r] * mean[j];
                mul33(vSc, R, A1);
                mul33_nt(A1, Sg, A1);   // (vΣcam R) Σ'
                mul33_tn(vSc, R, A2);
                mul33(A2, Sg, A2);      // (vΣcam' R) Σ
#pragma unroll
                for (int k = 0; k < 9; k++) {
                    const float v = vRl[k] + A1[k] + A2[k];
                    pose_acc[k] = fabsf(v) > 1e-7f ? v : 0.f;  // projection.jl:243-256
                }
#pragma unroll
                for (int k = 0; k < 3; k++) pose_acc[9 + k] = fabsf(vmc[k]) > 1e-7f ? vmc[k] : 0.f;
            }

            // normal channel: cotangent lands on one column of R_g (projection.jl:227-236)
            float vRg[9];
#pragma unroll
            for (int k = 0; k < 9; k++) vRg[k] = 0.f;
            if (channels > 5) {
                const int kk = (sc[0] <= sc[1] && sc[0] <= sc[2]) ? 0 : ((sc[1] <= sc[2]) ? 1 : 2);
                const float ax[3] = {kk == 0 ? Rg[0] : (kk == 1 ? Rg[3] : Rg[6]),
                                     kk == 0 ? Rg[1] : (kk == 1 ? Rg[4] : Rg[7]),
                                     kk == 0 ? Rg[2] : (kk == 1 ? Rg[5] : Rg[8])};
                float nc[3];
#pragma unroll
                for (int r = 0; r < 3; r++) nc[r] = R[r] * ax[0] + R[r + 3] * ax[1] + R[r + 6] * ax[2];
                const float sign = (nc[0] * mc[0] + nc[1] * mc[1] + nc[2] * mc[2]) > 0.f ? -1.f : 1.f;
                const float vn[3] = {vcol[5], vcol[6], vcol[7]};
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    const float gr = M3(R, 0, r) * vn[0] + M3(R, 1, r) * vn[1] + M3(R, 2, r) * vn[2];
                    if (kk == 0) M3(vRg, r, 0) = sign * gr;
                    else if (kk == 1) M3(vRg, r, 1) = sign * gr;
                    else M3(vRg, r, 2) = sign * gr;
                }
            }

            // ∇quat_scale_to_cov (render.jl:302-320)
            float vM[9], vRr[9], sym[9];
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int r = 0; r < 3; r++) M3(sym, r, j) = M3(vSg, r, j) + M3(vSg, j, r);
            mul33(sym, Mm, vM);
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int r = 0; r < 3; r++) M3(vRr, r, j) = M3(vM, r, j) * sc[j] + M3(vRg, r, j);
            float vscale[3];
#pragma unroll
            for (int j = 0; j < 3; j++) vscale[j] = M3(Rg, 0, j) * M3(vM, 0, j) + M3(Rg, 1, j) * M3(vM, 1, j) + M3(Rg, 2, j) * M3(vM, 2, j);
            // ∇unnorm_quat2rot (render.jl:335-366)
#define V(r, c) M3(vRr, (r) - 1, (c) - 1)
            float vqn[4];
            vqn[0] = 2.0f * (x * (V(3, 2) - V(2, 3)) + y * (V(1, 3) - V(3, 1)) + z * (V(2, 1) - V(1, 2)));
            vqn[1] = 2.0f * (-2.0f * x * (V(2, 2) + V(3, 3)) + y * (V(2, 1) + V(1, 2)) + z * (V(3, 1) + V(1, 3)) + w * (V(3, 2) - V(2, 3)));
            vqn[2] = 2.0f * (x * (V(2, 1) + V(1, 2)) - 2.0f * y * (V(1, 1) + V(3, 3)) + z * (V(3, 2) + V(2, 3)) + w * (V(1, 3) - V(3, 1)));
            vqn[3] = 2.0f * (x * (V(3, 1) + V(1, 3)) + y * (V(3, 2) + V(2, 3)) - 2.0f * z * (V(1, 1) + V(2, 2)) + w * (V(2, 1) - V(1, 2)));
#undef V
            const float qd = vqn[0] * w + vqn[1] * x + vqn[2] * y + vqn[3] * z;
            const float vq[4] = {(vqn[0] - qd * w) * qi, (vqn[1] - qd * x) * qi, (vqn[2] - qd * y) * qi, (vqn[3] - qd * z) * qi};

            // ∇color_from_sh! (spherical_harmonics.jl:76-171) + ∇normalize (:174-181)
            {
                const float d0 = mean[0] - cam.cam_center[0], d1 = mean[1] - cam.cam_center[1], d2 = mean[2] - cam.cam_center[2];
                const float s2 = d0 * d0 + d1 * d1 + d2 * d2;
                const float inv = 1.0f / sqrtf(s2);
                const float dxn = inv * d0, dyn = inv * d1, dzn = inv * d2;
                float vc[3];
#pragma unroll
                for (int c = 0; c < 3; c++) vc[c] = vcol[c] * (1.0f - (float)g.clamped[3 * i + c]);
                const float X = dxn, Y = dyn, Z = dzn;
                const float x2 = X * X, y2 = Y * Y, z2 = Z * Z, xy = X * Y, xz = X * Z, yz = Y * Z;
                const float *sh = vsh;  // staged coefficients (read fully before the row is overwritten below)
                float vdir[3] = {0.f, 0.f, 0.f};
                float basis[16];
                basis[0] = SH0;
                if (sh_degree > 0) {
                    basis[1] = -SH1 * Y; basis[2] = SH1 * Z; basis[3] = -SH1 * X;
                    if (sh_degree > 1) {
                        basis[4] = SH2C1 * xy; basis[5] = SH2C2 * yz; basis[6] = SH2C3 * (2.0f * z2 - x2 - y2);
                        basis[7] = SH2C4 * xz; basis[8] = SH2C5 * (x2 - y2);
                        if (sh_degree > 2) {
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
                    if (sh_degree > 0) {
                        gx = -SH1 * S(4); gy = -SH1 * S(2); gz = SH1 * S(3);
                        if (sh_degree > 1) {
                            gx += SH2C1 * Y * S(5) + SH2C3 * 2.0f * -X * S(7) + SH2C4 * Z * S(8) + SH2C5 * 2.0f * X * S(9);
                            gy += SH2C1 * X * S(5) + SH2C2 * Z * S(6) + SH2C3 * 2.0f * -Y * S(7) + SH2C5 * 2.0f * -Y * S(9);
                            gz += SH2C2 * Y * S(6) + SH2C3 * 4.0f * Z * S(7) + SH2C4 * X * S(8);
                            if (sh_degree > 2) {
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
                for (int k = 0; k < 16; k++) {  // constant trip count keeps basis[] in registers
                    if (k < k_used) {
#pragma unroll
                        for (int c = 0; c < 3; c++) vsh[3 * k + c] = basis[k] * vc[c];
                    }
                }
                for (int e = 3 * k_used; e < row; e++) vsh[e] = 0.f;  // rows k+1..K stay zero (Appendix A.5)
                const float inv_s = 1.0f / sqrtf(s2 * s2 * s2);
                vmean[0] += ((s2 - d0 * d0) * vdir[0] - d1 * d0 * vdir[1] - d2 * d0 * vdir[2]) * inv_s;
                vmean[1] += (-d0 * d1 * vdir[0] + (s2 - d1 * d1) * vdir[1] - d2 * d1 * vdir[2]) * inv_s;
                vmean[2] += (-d0 * d2 * vdir[0] - d1 * d2 * vdir[1] + (s2 - d2 * d2) * vdir[2]) * inv_s;
            }

#pragma unroll
            for (int k = 0; k < 3; k++) put<ACC>(vmeans + 3 * i + k, vmean[k]);
            if (RAW && (ps.raw_scale || ps.isotropic)) {  // pullback of exp: d exp(s) = exp(s) ds (rasterizer.jl:237)
#pragma unroll
                for (int k = 0; k < 3; k++) vscale[k] *= sc[k];
            }
            if (RAW && ps.isotropic) {
                put<ACC>(vscales + i, (vscale[0] + vscale[1]) + vscale[2]);  // pullback of vcat(s, s, s)
            } else {
#pragma unroll
                for (int k = 0; k < 3; k++) put<ACC>(vscales + 3 * i + k, vscale[k]);
            }
            if (ACC) {
                float4 o = *reinterpret_cast<float4 *>(vrot + 4 * i);
                o.x += vq[0]; o.y += vq[1]; o.z += vq[2]; o.w += vq[3];
                *reinterpret_cast<float4 *>(vrot + 4 * i) = o;
            } else {
                *reinterpret_cast<float4 *>(vrot + 4 * i) = make_float4(vq[0], vq[1], vq[2], vq[3]);  // vstore! projection.jl:240
            }
        }
    }

    // coalesced write-out of the CTA's SH-gradient span
    __syncthreads();
    {
        if (RAW && ps.sh_rest) {
            rows_shared_to_global<ACC>(vshs + block0 * 3, s_sh, (int)nb, 3, sh_stride, tid, BG_THREADS, false);
            if (K > 1)
                rows_shared_to_global<ACC>(ps.vsh_rest + block0 * (int64_t)(row - 3), s_sh + 3, (int)nb, row - 3, sh_stride, tid,
                                           BG_THREADS, (reinterpret_cast<uintptr_t>(ps.vsh_rest) & 15) == 0);
        } else {
            rows_shared_to_global<ACC>(vshs + block0 * row, s_sh, (int)nb, row, sh_stride, tid, BG_THREADS, aligned16);
        }
    }

    if (pose) {  // CTA-level reduction of the 12 pose cotangents before one atomic each (TODO at projection.jl:242)
        __shared__ float s_pose[BG_THREADS / 32][12];
        const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
        for (int k = 0; k < 12; k++) {
            float v = pose_acc[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) s_pose[warp][k] = v;
        }
        __syncthreads();
        if (threadIdx.x < 12) {
            float v = 0.f;
#pragma unroll
            for (int w2 = 0; w2 < BG_THREADS / 32; w2++) v += s_pose[w2][threadIdx.x];
            if (v != 0.f) atomicAdd(threadIdx.x < 9 ? vR_out + threadIdx.x : vt_out + (threadIdx.x - 9), v);
        }
    }
}

__global__ void __launch_bounds__(256)
update_stats_kernel(const int64_t n, const int32_t *__restrict__ radii, const float2 *__restrict__ gm2, const float w,
                    const float h, int32_t *__restrict__ max_radii, float *__restrict__ accum,
                    float *__restrict__ denom) {
    // _update_stats! — strategy.jl:118-136 (op order kept; this TU may contract x*x+y*y, tolerance-checked)
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t r = radii[i];
    if (!(r > 0)) return;
    max_radii[i] = max(max_radii[i], r);
    const float2 gv = gm2[i];
    const float gx = __fmul_rn(__fmul_rn(gv.x, w), 0.5f), gy = __fmul_rn(__fmul_rn(gv.y, h), 0.5f);
    accum[i] += sqrtf(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)));
    denom[i] += 1.0f;
}

// FP32 FMA micro-benchmark: 8 independent dependent-FFMA chains per thread, 1024 threads x 2 CTAs per SM.
__global__ void __launch_bounds__(1024) fp32_peak_kernel(float *out, const int iters, const float a, const float b) {
    float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f,
          x7 = x0 + 7.f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    const float r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (r == 123.456f) out[0] = r;  // never true: keeps the chains alive
}

}  // namespace

int launch_fp32_peak(cudaStream_t s, double *ms, double *flops) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    float *out = nullptr;
    if (cudaMalloc(&out, 4) != cudaSuccess) return -1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 4096, blocks = sms * 2;
    fp32_peak_kernel<<<blocks, 1024, 0, s>>>(out, 64, 0.999f, 0.001f);  // warm-up
    cudaEventRecord(e0, s);
    fp32_peak_kernel<<<blocks, 1024, 0, s>>>(out, iters, 0.999f, 0.001f);
    cudaEventRecord(e1, s);
    cudaEventSynchronize(e1);
    float t = 0.f;
    cudaEventElapsedTime(&t, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    count_launch(2);
    *ms = t;
    *flops = 2.0 * 64.0 * (double)iters * 1024.0 * (double)blocks;  // 8 chains x 8 unroll FMAs per iteration
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

void launch_backward_gaussians(const DevCamera &cam, int64_t n, int sh_degree, int K, int channels,
                               const float *means, const float *shs, const float *opac, const float *scales,
                               const float *rots,
                               const GeomPtrs &g, float *vmeans, float *vshs, float *vopac, float *vscales,
                               float *vrot, float *vR, float *vt, int accumulate, cudaStream_t s, const ParamSpec &ps) {
    if (n <= 0) return;
    const unsigned blocks = (unsigned)((n + BG_THREADS - 1) / BG_THREADS);
    int stride = 3 * K;
    if ((stride & 1) == 0) stride += 1;
    const size_t smem = (size_t)BG_THREADS * stride * sizeof(float);
    const bool raw = ps.raw_opacity || ps.raw_scale || ps.isotropic || ps.sh_rest;
#define GSR_BG(AC, RW)                                                                                                    \
    backward_gaussians_kernel<AC, RW><<<blocks, BG_THREADS, smem, s>>>(cam, n, sh_degree, K, channels, means, shs, opac,    \
                                                                     scales, rots, g, vmeans, vshs, vopac, vscales, vrot, \
                                                                     vR, vt, stride, ps)
    if (accumulate) {
        if (raw) GSR_BG(true, true); else GSR_BG(true, false);
    } else {
        if (raw) GSR_BG(false, true); else GSR_BG(false, false);
    }
#undef GSR_BG
    count_launch();
}

void launch_update_stats(int64_t n, const int32_t *radii, const float2 *grad_means2d, uint32_t width,
                         uint32_t height, int32_t *max_radii, float *accum, float *denom, cudaStream_t s) {
    if (n <= 0) return;
    update_stats_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, radii, grad_means2d, (float)width, (float)height,
                                                                   max_radii, accum, denom);
    count_launch();
}
