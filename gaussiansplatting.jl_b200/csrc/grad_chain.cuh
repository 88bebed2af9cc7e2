// grad_chain.cuh — the per-(view, Gaussian) pullbacks of ∇project! and ∇spherical_harmonics!, written once.
//
// Used by backward_gaussians.cu (one view, the forward's own state) and backward_peers.cu (all views of a batch, over
// peer memory): both kernels evaluate exactly these expressions, in this order, so the single-GPU and the multi-GPU
// gradients cannot drift apart.  Each function cites the reference lines it differentiates.
//
// Gradients are tolerance-checked (1e-4 relative), so FMA contraction is allowed here.
#pragma once

#define M3(m, i, j) ((m)[(i) + 3 * (j)])

namespace gchain {

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

// real spherical-harmonics constants (utils.jl:33-48)
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

// unnorm_quat2rot (render.jl:322-333): normalised quaternion (w, x, y, z), 1/|q| and the rotation matrix R_g
struct Quat {
    float qi, w, x, y, z;
};
__device__ __forceinline__ Quat quat_to_rot(const float4 q4, float *Rg) {
    Quat q;
    const float qn = sqrtf(q4.x * q4.x + q4.y * q4.y + q4.z * q4.z + q4.w * q4.w);
    q.qi = 1.0f / qn;
    q.w = q.qi * q4.x; q.x = q.qi * q4.y; q.y = q.qi * q4.z; q.z = q.qi * q4.w;
    const float w = q.w, x = q.x, y = q.y, z = q.z;
    Rg[0] = 1.0f - 2.0f * (y * y + z * z); Rg[1] = 2.0f * (x * y + w * z); Rg[2] = 2.0f * (x * z - w * y);
    Rg[3] = 2.0f * (x * y - w * z); Rg[4] = 1.0f - 2.0f * (x * x + z * z); Rg[5] = 2.0f * (y * z + w * x);
    Rg[6] = 2.0f * (x * z + w * y); Rg[7] = 2.0f * (y * z - w * x); Rg[8] = 1.0f - 2.0f * (x * x + y * y);
    return q;
}

// ∇unnorm_quat2rot (render.jl:335-366): cotangent of R_g -> cotangent of the un-normalised quaternion
__device__ __forceinline__ void grad_quat(const float *vRr, const Quat &q, float *vq) {
    const float w = q.w, x = q.x, y = q.y, z = q.z, qi = q.qi;
#define V(r, c) M3(vRr, (r) - 1, (c) - 1)
    float vqn[4];
    vqn[0] = 2.0f * (x * (V(3, 2) - V(2, 3)) + y * (V(1, 3) - V(3, 1)) + z * (V(2, 1) - V(1, 2)));
    vqn[1] = 2.0f * (-2.0f * x * (V(2, 2) + V(3, 3)) + y * (V(2, 1) + V(1, 2)) + z * (V(3, 1) + V(1, 3)) + w * (V(3, 2) - V(2, 3)));
    vqn[2] = 2.0f * (x * (V(2, 1) + V(1, 2)) - 2.0f * y * (V(1, 1) + V(3, 3)) + z * (V(3, 2) + V(2, 3)) + w * (V(1, 3) - V(3, 1)));
    vqn[3] = 2.0f * (x * (V(3, 1) + V(1, 3)) + y * (V(3, 2) + V(2, 3)) - 2.0f * z * (V(1, 1) + V(2, 2)) + w * (V(2, 1) - V(1, 2)));
#undef V
    const float qd = vqn[0] * w + vqn[1] * x + vqn[2] * y + vqn[3] * z;
    vq[0] = (vqn[0] - qd * w) * qi; vq[1] = (vqn[1] - qd * x) * qi; vq[2] = (vqn[2] - qd * y) * qi; vq[3] = (vqn[3] - qd * z) * qi;
}

// ∇inverse (render.jl:383-385): vΣ2D = -Σ⁻¹ vΣ⁻¹ Σ⁻¹ with symmetric 2x2 operands (projection.jl:178-188);
// conic = (ca, cb, cc), vcn = cotangent of the conic's three entries, vS2 column-major 2x2
__device__ __forceinline__ void grad_inverse2(const float ca, const float cb, const float cc, const float *vcn, float *vS2) {
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

// The forward intermediates of perspective_projection for one (camera, Gaussian) (projection.jl:259-287): clamped
// tangents and the 2x3 Jacobian J (column-major)
struct Persp {
    float lim[2], limn[2], txy[2], rz, rz2, rz3, fx, fy, J[6];
};
__device__ __forceinline__ Persp persp_setup(const float *focal, const float *principal, const int width, const int height,
                                             const float *mc) {
    Persp P;
    const float res[2] = {(float)width, (float)height};
    P.rz = 1.0f / mc[2];
    P.rz2 = P.rz * P.rz;
    P.rz3 = P.rz2 * P.rz;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const float stf = 0.3f * ((0.5f * res[k]) / focal[k]);
        const float pp = principal[k] * res[k];
        P.lim[k] = (res[k] - pp) / focal[k] + stf;
        P.limn[k] = pp / focal[k] + stf;
        P.txy[k] = mc[2] * fminf(P.lim[k], fmaxf(-P.limn[k], mc[k] * P.rz));
    }
    P.fx = focal[0];
    P.fy = focal[1];
    P.J[0] = P.fx * P.rz; P.J[1] = 0.f; P.J[2] = 0.f; P.J[3] = P.fy * P.rz;
    P.J[4] = -P.fx * P.txy[0] * P.rz2; P.J[5] = -P.fy * P.txy[1] * P.rz2;
    return P;
}
#define GC_J(r, c) P.J[(r) + 2 * (c)]

// conic of a view recomputed from Σcam: Σ2D = J Σcam J' + blur, then its inverse (projection.jl:259-287, render.jl:368-396)
__device__ __forceinline__ void conic_from_cov(const Persp &P, const float *Sc, const float blur_eps, float &ca, float &cb,
                                               float &cc) {
    float TJ[6];
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int r = 0; r < 2; r++) TJ[r + 2 * j] = GC_J(r, 0) * M3(Sc, 0, j) + GC_J(r, 1) * M3(Sc, 1, j) + GC_J(r, 2) * M3(Sc, 2, j);
    const float s00 = TJ[0] * GC_J(0, 0) + TJ[2] * GC_J(0, 1) + TJ[4] * GC_J(0, 2) + blur_eps;
    const float s10 = TJ[1] * GC_J(0, 0) + TJ[3] * GC_J(0, 1) + TJ[5] * GC_J(0, 2);
    const float s01 = TJ[0] * GC_J(1, 0) + TJ[2] * GC_J(1, 1) + TJ[4] * GC_J(1, 2);
    const float s11 = TJ[1] * GC_J(1, 0) + TJ[3] * GC_J(1, 1) + TJ[5] * GC_J(1, 2) + blur_eps;
    const float det_inv = 1.0f / (s00 * s11 - s01 * s10);
    ca = s11 * det_inv; cb = -s01 * det_inv; cc = s00 * det_inv;
}

// ∇perspective_projection (projection.jl:289-353): (vΣ2D, v_mean2d) -> (vΣcam, v_mean_cam)
__device__ __forceinline__ void grad_perspective(const Persp &P, const float *mc, const float *Sc, const float *vS2,
                                                 const float *vm2, float *vSc, float *vmc) {
#define V2(r, c) vS2[(r) + 2 * (c)]
    float A[6];  // J' * vΣ2D  (3x2)
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
        for (int r = 0; r < 3; r++) A[r + 3 * j] = GC_J(0, r) * V2(0, j) + GC_J(1, r) * V2(1, j);
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int r = 0; r < 3; r++) M3(vSc, r, j) = A[r] * GC_J(0, j) + A[r + 3] * GC_J(1, j);
    float B1[6], B2[6], vJ[6];
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int r = 0; r < 2; r++) {
            B1[r + 2 * j] = V2(r, 0) * GC_J(0, j) + V2(r, 1) * GC_J(1, j);
            B2[r + 2 * j] = V2(0, r) * GC_J(0, j) + V2(1, r) * GC_J(1, j);
        }
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int r = 0; r < 2; r++)
            vJ[r + 2 * j] = (B1[r] * M3(Sc, j, 0) + B1[r + 2] * M3(Sc, j, 1) + B1[r + 4] * M3(Sc, j, 2)) +
                            (B2[r] * M3(Sc, 0, j) + B2[r + 2] * M3(Sc, 1, j) + B2[r + 4] * M3(Sc, 2, j));
#undef V2
#define VJ(r, c) vJ[((r) - 1) + 2 * ((c) - 1)]
    const float fx = P.fx, fy = P.fy, rz = P.rz, rz2 = P.rz2, rz3 = P.rz3;
    float vx = fx * rz * vm2[0];
    float vy = fy * rz * vm2[1];
    float vz = -rz2 * (fx * mc[0] * vm2[0] + fy * mc[1] * vm2[1]);
    const float ax = mc[0] * rz, ay = mc[1] * rz;
    if (-P.limn[0] <= ax && ax <= P.lim[0]) vx += -fx * rz2 * VJ(1, 3);
    else vz += -fx * rz3 * VJ(1, 3) * P.txy[0];
    if (-P.limn[1] <= ay && ay <= P.lim[1]) vy += -fy * rz2 * VJ(2, 3);
    else vz += -fy * rz3 * VJ(2, 3) * P.txy[1];
    vz += -fx * rz2 * VJ(1, 1) - fy * rz2 * VJ(2, 2) + 2.0f * fx * P.txy[0] * rz3 * VJ(1, 3) +
          2.0f * fy * P.txy[1] * rz3 * VJ(2, 3);
#undef VJ
    vmc[0] = vx; vmc[1] = vy; vmc[2] = vz;
}

// normal channel (projection.jl:227-236): the cotangent of the normal lands on column kk (the thinnest axis) of R_g.
// Returns the three entries of that column's cotangent in gr[].
__device__ __forceinline__ void grad_normal(const float *R, const float *Rg, const float *mc, const int kk, const float *vn,
                                            float *gr) {
    const float ax[3] = {kk == 0 ? Rg[0] : (kk == 1 ? Rg[3] : Rg[6]), kk == 0 ? Rg[1] : (kk == 1 ? Rg[4] : Rg[7]),
                         kk == 0 ? Rg[2] : (kk == 1 ? Rg[5] : Rg[8])};
    float nc[3];
#pragma unroll
    for (int r = 0; r < 3; r++) nc[r] = R[r] * ax[0] + R[r + 3] * ax[1] + R[r + 6] * ax[2];
    const float sign = (nc[0] * mc[0] + nc[1] * mc[1] + nc[2] * mc[2]) > 0.f ? -1.f : 1.f;
#pragma unroll
    for (int r = 0; r < 3; r++) gr[r] = sign * (M3(R, 0, r) * vn[0] + M3(R, 1, r) * vn[1] + M3(R, 2, r) * vn[2]);
}
__device__ __forceinline__ int thinnest_axis(const float *sc) {
    return (sc[0] <= sc[1] && sc[0] <= sc[2]) ? 0 : ((sc[1] <= sc[2]) ? 1 : 2);
}

// ∇color_from_sh! (spherical_harmonics.jl:76-171) + ∇normalize (:174-181).  sh = the Gaussian's coefficients [k][rgb],
// vc = colour cotangent (zeroed where the colour was clamped).  Fills basis[0..k) (the SH gradient of coefficient k is
// basis[k] * vc) and ADDS the direction pullback to vmean.
__device__ __forceinline__ void grad_sh(const int sh_degree, const float *mean, const float *cam_center, const float *sh,
                                        const float *vc, float *basis, float *vmean) {
    const float d0 = mean[0] - cam_center[0], d1 = mean[1] - cam_center[1], d2 = mean[2] - cam_center[2];
    const float s2 = d0 * d0 + d1 * d1 + d2 * d2;
    const float inv = 1.0f / sqrtf(s2);
    const float X = inv * d0, Y = inv * d1, Z = inv * d2;
    const float x2 = X * X, y2 = Y * Y, z2 = Z * Z, xy = X * Y, xz = X * Z, yz = Y * Z;
    float vdir[3] = {0.f, 0.f, 0.f};
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
    const float inv_s = 1.0f / sqrtf(s2 * s2 * s2);
    vmean[0] += ((s2 - d0 * d0) * vdir[0] - d1 * d0 * vdir[1] - d2 * d0 * vdir[2]) * inv_s;
    vmean[1] += (-d0 * d1 * vdir[0] + (s2 - d1 * d1) * vdir[1] - d2 * d1 * vdir[2]) * inv_s;
    vmean[2] += (-d0 * d2 * vdir[0] - d1 * d2 * vdir[1] + (s2 - d2 * d2) * vdir[2]) * inv_s;
}

}  // namespace gchain
