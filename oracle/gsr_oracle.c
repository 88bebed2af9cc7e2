/*
 * gsr_oracle.c — CPU restatement of GaussianSplatting.jl's rasterizer hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (gaussiansplatting.jl_b200/)
 * may include, link or call this file.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker / baseline.
 *
 * PARITY STATUS: the reference (100 % Julia, KernelAbstractions `@kernel cpu=false`)
 * cannot execute in this container (no Julia toolchain, no GPU) and ships no golden
 * arrays.  This restatement is pinned against every known-answer test the reference's
 * own test-suite holds for the path (get_rect `test/runtests.jl:308-324`, tile ranges
 * `:486-494`) and against its finite-difference / end-to-end property tests
 * (`:86-306`, `:555-611`, `:697-853`), re-expressed in tests/test_oracle_*.py.
 * Bit-level parity with the reference binary is UNPINNED: the op order below is the
 * documented normative choice: Julia never contracts a*b+c,
 * StaticArrays' unrolled products sum left-to-right, `normalize(v) = inv(norm(v))*v`,
 * `A*B*C = (A*B)*C` for the shapes used (LinearAlgebra._tri_matmul tie → left).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC  (see oracle/Makefile)
 *   default            → real = float   (faithful fp32 restatement; liboracle_f32.so)
 *   -DORC_F64          → real = double  (same formulas in fp64 for FD / gradient tie-breaks)
 *
 * Layout conventions (identical to the reference, Julia column-major):
 *   (3,N) matrix  == N packed real[3];  image (C,W,H) == H rows × W cols × C channels.
 *   3x3 matrices are column-major: M[i + 3*j] is row i, column j (0-based).
 *   Gaussian ids in `values` are 1-based, exactly as `duplicate_with_keys!` emits them.
 */
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef ORC_F64
typedef double real;
#define R_SQRT sqrt
#define R_EXP exp
#define R_FLOOR floor
#define R_CEIL ceil
#define R_FABS fabs
#else
typedef float real;
#define R_SQRT sqrtf
#define R_EXP expf
#define R_FLOOR floorf
#define R_CEIL ceilf
#define R_FABS fabsf
#endif

#define RC(x) ((real)(x##f)) /* fp32 literal widened to `real` */
#define EXPORT __attribute__((visibility("default")))

#define BLOCK 16 /* GaussianSplatting.jl:55-56 BLOCK = (16,16) */

/* OpenMP team size of every stage below; returns the size in effect.  n <= 0 only queries.  (A launcher such as
 * torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU baseline must not inherit that silently.) */
EXPORT int orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

static inline real rmin_(real a, real b) { return a < b ? a : b; }
static inline real rmax_(real a, real b) { return a > b ? a : b; }
static inline int32_t iclamp(int32_t x, int32_t lo, int32_t hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* unsafe_trunc(Int32, x): CUDA lowers fptosi to cvt.rzi.s32.f32, which saturates. */
static inline int32_t trunc_i32(real x) {
    if (!(x == x)) return 0;
    if (x >= (real)2147483647.0) return INT32_MAX;
    if (x <= (real)-2147483648.0) return INT32_MIN;
    return (int32_t)x;
}

/* ------------------------------------------------------------------ */
/* 3x3 / 2x2 helpers in StaticArrays' unrolled left-to-right order     */
/* ------------------------------------------------------------------ */
#define M3(m, i, j) ((m)[(i) + 3 * (j)])

static void mul33(const real *A, const real *B, real *C) { /* C = A*B */
    real T[9];
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++)
            T[i + 3 * j] = (M3(A, i, 0) * M3(B, 0, j) + M3(A, i, 1) * M3(B, 1, j)) + M3(A, i, 2) * M3(B, 2, j);
    memcpy(C, T, sizeof T);
}
static void transpose33(const real *A, real *T) {
    real t[9];
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++) t[i + 3 * j] = M3(A, j, i);
    memcpy(T, t, sizeof t);
}
static void mulvec3(const real *A, const real *v, real *o) { /* o = A*v */
    real t[3];
    for (int i = 0; i < 3; i++) t[i] = (M3(A, i, 0) * v[0] + M3(A, i, 1) * v[1]) + M3(A, i, 2) * v[2];
    o[0] = t[0]; o[1] = t[1]; o[2] = t[2];
}

/* ------------------------------------------------------------------ */
/* math helpers (render.jl:288-420, projection.jl:259-393)             */
/* ------------------------------------------------------------------ */

/* unnorm_quat2rot — render.jl:322-333.  q = (w,x,y,z). */
EXPORT void orc_unnorm_quat2rot(const real *q_in, real *R) {
    real n = R_SQRT(((q_in[0] * q_in[0] + q_in[1] * q_in[1]) + q_in[2] * q_in[2]) + q_in[3] * q_in[3]);
    real inv = RC(1.0) / n; /* normalize(q) = inv(norm(q)) * q */
    real w = inv * q_in[0], x = inv * q_in[1], y = inv * q_in[2], z = inv * q_in[3];
    real x2 = x * x, y2 = y * y, z2 = z * z;
    real xy = x * y, xz = x * z, yz = y * z;
    real wx = w * x, wy = w * y, wz = w * z;
    R[0] = RC(1.0) - RC(2.0) * (y2 + z2); R[1] = RC(2.0) * (xy + wz); R[2] = RC(2.0) * (xz - wy);
    R[3] = RC(2.0) * (xy - wz); R[4] = RC(1.0) - RC(2.0) * (x2 + z2); R[5] = RC(2.0) * (yz + wx);
    R[6] = RC(2.0) * (xz + wy); R[7] = RC(2.0) * (yz - wx); R[8] = RC(1.0) - RC(2.0) * (x2 + y2);
}

/* ∇unnorm_quat2rot — render.jl:335-366. */
EXPORT void orc_grad_unnorm_quat2rot(const real *q_in, const real *vR, real *vq) {
    real n = R_SQRT(((q_in[0] * q_in[0] + q_in[1] * q_in[1]) + q_in[2] * q_in[2]) + q_in[3] * q_in[3]);
    real inv_norm = RC(1.0) / n;
    real q[4] = {q_in[0] * inv_norm, q_in[1] * inv_norm, q_in[2] * inv_norm, q_in[3] * inv_norm};
    real w = q[0], x = q[1], y = q[2], z = q[3];
#define V(i, j) M3(vR, (i) - 1, (j) - 1)
    real vqn[4];
    vqn[0] = RC(2.0) * ((x * (V(3, 2) - V(2, 3)) + y * (V(1, 3) - V(3, 1))) + z * (V(2, 1) - V(1, 2)));
    vqn[1] = RC(2.0) * (((RC(-2.0) * x * (V(2, 2) + V(3, 3)) + y * (V(2, 1) + V(1, 2))) + z * (V(3, 1) + V(1, 3))) +
                        w * (V(3, 2) - V(2, 3)));
    vqn[2] = RC(2.0) * (((x * (V(2, 1) + V(1, 2)) - RC(2.0) * y * (V(1, 1) + V(3, 3))) + z * (V(3, 2) + V(2, 3))) +
                        w * (V(1, 3) - V(3, 1)));
    vqn[3] = RC(2.0) * (((x * (V(3, 1) + V(1, 3)) + y * (V(3, 2) + V(2, 3))) - RC(2.0) * z * (V(1, 1) + V(2, 2))) +
                        w * (V(2, 1) - V(1, 2)));
#undef V
    real d = ((vqn[0] * q[0] + vqn[1] * q[1]) + vqn[2] * q[2]) + vqn[3] * q[3];
    for (int k = 0; k < 4; k++) vq[k] = (vqn[k] - d * q[k]) * inv_norm;
}

/* quat_scale_to_cov — render.jl:291-294: M = R*diag(s); Σ = M*M'. */
EXPORT void orc_quat_scale_to_cov(const real *R, const real *s, real *Sigma) {
    real S[9] = {s[0], 0, 0, 0, s[1], 0, 0, 0, s[2]};
    real M[9], Mt[9];
    mul33(R, S, M);
    transpose33(M, Mt);
    mul33(M, Mt, Sigma);
}

/* ∇quat_scale_to_cov — render.jl:302-320. */
EXPORT void orc_grad_quat_scale_to_cov(const real *q, const real *s, const real *R, const real *vSigma,
                                       const real *vR_extra, real *vq, real *vscale) {
    real S[9] = {s[0], 0, 0, 0, s[1], 0, 0, 0, s[2]};
    real M[9], vSt[9], sym[9], vM[9], vR[9];
    mul33(R, S, M);
    transpose33(vSigma, vSt);
    for (int k = 0; k < 9; k++) sym[k] = vSigma[k] + vSt[k];
    mul33(sym, M, vM);
    mul33(vM, S, vR);
    for (int k = 0; k < 9; k++) vR[k] = vR[k] + (vR_extra ? vR_extra[k] : (real)0);
    orc_grad_unnorm_quat2rot(q, vR, vq);
    for (int j = 0; j < 3; j++)
        vscale[j] = (M3(R, 0, j) * M3(vM, 0, j) + M3(R, 1, j) * M3(vM, 1, j)) + M3(R, 2, j) * M3(vM, 2, j);
}

/* pos_world_to_cam — projection.jl:355-361. */
EXPORT void orc_pos_world_to_cam(const real *R, const real *t, const real *p, real *out) {
    real r[3];
    mulvec3(R, p, r);
    out[0] = r[0] + t[0]; out[1] = r[1] + t[1]; out[2] = r[2] + t[2];
}

/* ∇pos_world_to_cam — projection.jl:363-373. */
EXPORT void orc_grad_pos_world_to_cam(const real *R, const real *t, const real *p, const real *v, real *vR, real *vt,
                                      real *vp) {
    (void)t;
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++) M3(vR, i, j) = v[i] * p[j];
    vt[0] = v[0]; vt[1] = v[1]; vt[2] = v[2];
    real Rt[9];
    transpose33(R, Rt);
    mulvec3(Rt, v, vp);
}

/* covar_world_to_cam — projection.jl:375-380: (R*Σ)*R'. */
EXPORT void orc_covar_world_to_cam(const real *R, const real *Sigma, real *out) {
    real T[9], Rt[9];
    mul33(R, Sigma, T);
    transpose33(R, Rt);
    mul33(T, Rt, out);
}

/* ∇covar_world_to_cam — projection.jl:382-393.  vR_io: grad-in on entry, grad-out on exit. */
EXPORT void orc_grad_covar_world_to_cam(const real *R, const real *Sigma, const real *vSc, real *vR_io, real *vSigma) {
    real St[9], vSct[9], A[9], B[9], Rt[9];
    transpose33(Sigma, St);
    transpose33(vSc, vSct);
    mul33(vSc, R, A); mul33(A, St, A);      /* (vΣcam*R)*Σ' */
    mul33(vSct, R, B); mul33(B, Sigma, B);  /* (vΣcam'*R)*Σ */
    for (int k = 0; k < 9; k++) vR_io[k] = (vR_io[k] + A[k]) + B[k];
    transpose33(R, Rt);
    mul33(Rt, vSc, A);
    mul33(A, R, vSigma);
}

typedef struct {
    real tan_fov[2], stf[2], pp[2], lim[2], lim_neg[2];
} ProjConsts;

static void proj_consts(const real *focal, const int32_t *res, const real *principal, ProjConsts *c) {
    for (int k = 0; k < 2; k++) {
        real r = (real)res[k];
        c->tan_fov[k] = (RC(0.5) * r) / focal[k];
        c->stf[k] = RC(0.3) * c->tan_fov[k];
        c->pp[k] = principal[k] * r;
        c->lim[k] = (r - c->pp[k]) / focal[k] + c->stf[k];
        c->lim_neg[k] = c->pp[k] / focal[k] + c->stf[k];
    }
}

/* perspective_projection — projection.jl:259-287.  Σ2D is 2x2 column-major. */
EXPORT void orc_perspective_projection(const real *mean, const real *Sigma, const real *focal, const int32_t *res,
                                       const real *principal, real *S2, real *mean2d) {
    ProjConsts c;
    proj_consts(focal, res, principal, &c);
    real rz = RC(1.0) / mean[2];
    real rz2 = rz * rz;
    real txy[2];
    for (int k = 0; k < 2; k++) {
        mean2d[k] = (rz * focal[k]) * mean[k] + c.pp[k];
        txy[k] = mean[2] * rmin_(c.lim[k], rmax_(-c.lim_neg[k], mean[k] * rz));
    }
    /* J (2x3 column-major): [fx*rz 0 -fx*tx*rz²; 0 fy*rz -fy*ty*rz²] */
    real J[6] = {focal[0] * rz, 0, 0, focal[1] * rz, ((-focal[0]) * txy[0]) * rz2, ((-focal[1]) * txy[1]) * rz2};
#define J_(i, j) J[(i) + 2 * (j)]
    real T[6]; /* T = J*Σ (2x3) */
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 2; i++)
            T[i + 2 * j] = (J_(i, 0) * M3(Sigma, 0, j) + J_(i, 1) * M3(Sigma, 1, j)) + J_(i, 2) * M3(Sigma, 2, j);
    /* Σ2D = T*J' (2x2): [i,j] = Σ_k T[i,k]*J[j,k] */
    for (int j = 0; j < 2; j++)
        for (int i = 0; i < 2; i++)
            S2[i + 2 * j] = (T[i + 0] * J_(j, 0) + T[i + 2] * J_(j, 1)) + T[i + 4] * J_(j, 2);
#undef J_
}

/* ∇perspective_projection — projection.jl:289-353. */
EXPORT void orc_grad_perspective_projection(const real *mean, const real *Sigma, const real *focal, const int32_t *res,
                                            const real *principal, const real *vS2, const real *vmean2d, real *vSigma,
                                            real *vmean) {
    ProjConsts c;
    proj_consts(focal, res, principal, &c);
    real rz = RC(1.0) / mean[2];
    real rz2 = rz * rz, rz3 = rz2 * rz;
    real txy[2];
    for (int k = 0; k < 2; k++) txy[k] = mean[2] * rmin_(c.lim[k], rmax_(-c.lim_neg[k], mean[k] * rz));
    real J[6] = {focal[0] * rz, 0, 0, focal[1] * rz, ((-focal[0]) * txy[0]) * rz2, ((-focal[1]) * txy[1]) * rz2};
#define J_(i, j) J[(i) + 2 * (j)]
#define V2(i, j) vS2[(i) + 2 * (j)]
    /* vΣ = (J'*vΣ2D)*J : A = J' * vΣ2D (3x2), vΣ = A*J (3x3) */
    real A[6];
    for (int j = 0; j < 2; j++)
        for (int i = 0; i < 3; i++) A[i + 3 * j] = J_(0, i) * V2(0, j) + J_(1, i) * V2(1, j);
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++) M3(vSigma, i, j) = A[i + 0] * J_(0, j) + A[i + 3] * J_(1, j);
    /* vJ = (vΣ2D*J)*Σ' + (vΣ2D'*J)*Σ  (2x3) */
    real B1[6], B2[6], vJ[6];
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 2; i++) {
            B1[i + 2 * j] = V2(i, 0) * J_(0, j) + V2(i, 1) * J_(1, j);
            B2[i + 2 * j] = V2(0, i) * J_(0, j) + V2(1, i) * J_(1, j);
        }
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 2; i++) {
            real a = (B1[i + 0] * M3(Sigma, j, 0) + B1[i + 2] * M3(Sigma, j, 1)) + B1[i + 4] * M3(Sigma, j, 2);
            real b = (B2[i + 0] * M3(Sigma, 0, j) + B2[i + 2] * M3(Sigma, 1, j)) + B2[i + 4] * M3(Sigma, 2, j);
            vJ[i + 2 * j] = a + b;
        }
#define VJ(i, j) vJ[((i) - 1) + 2 * ((j) - 1)]
    real vx = (focal[0] * rz) * vmean2d[0];
    real vy = (focal[1] * rz) * vmean2d[1];
    real vz = (-rz2) * ((focal[0] * mean[0]) * vmean2d[0] + (focal[1] * mean[1]) * vmean2d[1]);
    real ax = mean[0] * rz, ay = mean[1] * rz;
    if (-c.lim_neg[0] <= ax && ax <= c.lim[0]) vx += ((-focal[0]) * rz2) * VJ(1, 3);
    else vz += (((-focal[0]) * rz3) * VJ(1, 3)) * txy[0];
    if (-c.lim_neg[1] <= ay && ay <= c.lim[1]) vy += ((-focal[1]) * rz2) * VJ(2, 3);
    else vz += (((-focal[1]) * rz3) * VJ(2, 3)) * txy[1];
    vz += ((((-focal[0]) * rz2) * VJ(1, 1) - (focal[1] * rz2) * VJ(2, 2)) +
           (((RC(2.0) * focal[0]) * txy[0]) * rz3) * VJ(1, 3)) +
          (((RC(2.0) * focal[1]) * txy[1]) * rz3) * VJ(2, 3);
#undef VJ
#undef V2
#undef J_
    vmean[0] = vx; vmean[1] = vy; vmean[2] = vz;
}

/* add_blur — render.jl:387-396.  Returns det of the blurred matrix. */
EXPORT real orc_add_blur(const real *S2, real eps, real *S2b, real *compensation) {
    real det_orig = S2[0] * S2[3] - S2[2] * S2[1];
    S2b[0] = S2[0] + eps; S2b[1] = S2[1]; S2b[2] = S2[2]; S2b[3] = S2[3] + eps;
    real det_blur = S2b[0] * S2b[3] - S2b[2] * S2b[1];
    if (compensation) *compensation = R_SQRT(rmax_((real)0, det_orig / det_blur));
    return det_blur;
}

/* ∇add_blur — render.jl:398-413 (dead on the path: compensations = nothing; kept for the FD test). */
EXPORT void orc_grad_add_blur(real comp, real vcomp, const real *conic2x2, real eps, real *out) {
    real det = conic2x2[0] * conic2x2[3] - conic2x2[2] * conic2x2[1];
    real vs = RC(0.5) * vcomp / (comp + RC(1e-6));
    real ct = RC(1.0) - comp * comp;
    out[0] = vs * (ct * conic2x2[0] - eps * det);
    out[1] = vs * ct * conic2x2[1];
    out[2] = vs * ct * conic2x2[2];
    out[3] = vs * (ct * conic2x2[3] - eps * det);
}

/* inverse — render.jl:368-381.  Returns det. */
EXPORT real orc_inverse(const real *x, real *xi) {
    real det = x[0] * x[3] - x[2] * x[1];
    if (det == (real)0) { /* `det ≈ 0f0` with default rtol is `det == 0` */
        xi[0] = xi[1] = xi[2] = xi[3] = 0;
        return det;
    }
    real det_inv = RC(1.0) / det;
    real tmp = (-x[2]) * det_inv;
    xi[0] = x[3] * det_inv; xi[1] = tmp; xi[2] = tmp; xi[3] = x[0] * det_inv;
    return det;
}

/* ∇inverse — render.jl:383-385: (-x*vx)*x, 2x2 column-major. */
EXPORT void orc_grad_inverse(const real *x, const real *vx, real *out) {
    real nx[4] = {-x[0], -x[1], -x[2], -x[3]}, T[4];
    for (int j = 0; j < 2; j++)
        for (int i = 0; i < 2; i++) T[i + 2 * j] = nx[i] * vx[2 * j] + nx[i + 2] * vx[1 + 2 * j];
    for (int j = 0; j < 2; j++)
        for (int i = 0; i < 2; i++) out[i + 2 * j] = T[i] * x[2 * j] + T[i + 2] * x[1 + 2 * j];
}

/* max_eigval_2D — render.jl:415-420. */
static real max_eigval_2d(const real *S2b, real det) {
    real mid = RC(0.5) * (S2b[0] + S2b[3]);
    return mid + R_SQRT(rmax_(RC(0.1), mid * mid - det));
}

/* ∇normalize — spherical_harmonics.jl:174-181. */
EXPORT void orc_grad_normalize(const real *d, const real *vd, real *out) {
    real s2 = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2];
    real inv_s = RC(1.0) / R_SQRT((s2 * s2) * s2);
    out[0] = (((s2 - d[0] * d[0]) * vd[0] - (d[1] * d[0]) * vd[1]) - (d[2] * d[0]) * vd[2]) * inv_s;
    out[1] = ((((-d[0]) * d[1]) * vd[0] + (s2 - d[1] * d[1]) * vd[1]) - (d[2] * d[1]) * vd[2]) * inv_s;
    out[2] = ((((-d[0]) * d[2]) * vd[0] - (d[1] * d[2]) * vd[1]) + (s2 - d[2] * d[2]) * vd[2]) * inv_s;
}

/* gaussian_normal — projection.jl:14-27.  Returns k (1-based); writes n_cam (signed) and sign. */
EXPORT int32_t orc_gaussian_normal(const real *Rw2c, const real *Rg, const real *scale, const real *mean_cam, real *n_out,
                                   real *sign_out) {
    int32_t k = (scale[0] <= scale[1] && scale[0] <= scale[2]) ? 1 : ((scale[1] <= scale[2]) ? 2 : 3);
    real axis[3] = {M3(Rg, 0, k - 1), M3(Rg, 1, k - 1), M3(Rg, 2, k - 1)};
    real n[3];
    mulvec3(Rw2c, axis, n);
    real d = (n[0] * mean_cam[0] + n[1] * mean_cam[1]) + n[2] * mean_cam[2];
    real sign = d > (real)0 ? RC(-1.0) : RC(1.0);
    n_out[0] = sign * n[0]; n_out[1] = sign * n[1]; n_out[2] = sign * n[2];
    if (sign_out) *sign_out = sign;
    return k;
}

/* get_rect — utils.jl:18-29.  rect = {xmin, ymin, xmax, ymax}. */
EXPORT void orc_get_rect(const real *pixel, int32_t radius, const int32_t *grid, int32_t *rect) {
    real r = (real)radius, b = (real)BLOCK;
    for (int k = 0; k < 2; k++) {
        rect[k] = iclamp(trunc_i32(R_FLOOR((pixel[k] - r) / b)), 0, grid[k]);
        /* gpu_cld(x, y) = trunc(floor((x + y - 1) / y)) */
        int32_t c = trunc_i32(R_FLOOR((((pixel[k] + r) + b) - RC(1.0)) / b));
        rect[2 + k] = iclamp(c, 0, grid[k]);
    }
}

/* ------------------------------------------------------------------ */
/* spherical harmonics (spherical_harmonics.jl, constants utils.jl:33-48) */
/* ------------------------------------------------------------------ */
#define SH0 RC(0.28209479177387814)
#define SH1 RC(0.4886025119029199)
#define SH2C1 RC(1.0925484305920792)
#define SH2C2 RC(-1.0925484305920792)
#define SH2C3 RC(0.31539156525252005)
#define SH2C4 RC(-1.0925484305920792)
#define SH2C5 RC(0.5462742152960396)
#define SH3C1 RC(-0.5900435899266435)
#define SH3C2 RC(2.890611442640554)
#define SH3C3 RC(-0.4570457994644658)
#define SH3C4 RC(0.3731763325901154)
#define SH3C5 RC(-0.4570457994644658)
#define SH3C6 RC(1.445305721320277)
#define SH3C7 RC(-0.5900435899266435)
#define EPS32 RC(1.1920929e-07) /* eps(Float32) */

static void unit_dir(const real *p, const real *cam, real *dir_orig, real *dir) {
    for (int k = 0; k < 3; k++) dir_orig[k] = p[k] - cam[k];
    real n = R_SQRT((dir_orig[0] * dir_orig[0] + dir_orig[1] * dir_orig[1]) + dir_orig[2] * dir_orig[2]);
    real inv = RC(1.0) / n;
    for (int k = 0; k < 3; k++) dir[k] = inv * dir_orig[k];
}

/* compute_colors_from_sh — spherical_harmonics.jl:41-74.  shs: K packed real[3]. */
EXPORT void orc_compute_colors_from_sh(const real *point, const real *cam, const real *shs, int degree, real *rgb,
                                       uint8_t *clamped) {
    real dir_o[3], d[3] = {0, 0, 0};
    if (degree > 0) unit_dir(point, cam, dir_o, d);
    real x = d[0], y = d[1], z = d[2];
    real x2 = x * x, y2 = y * y, z2 = z * z, xy = x * y, xz = x * z, yz = y * z;
    for (int c = 0; c < 3; c++) {
#define S(k) shs[3 * ((k) - 1) + c]
        real res = SH0 * S(1);
        if (degree > 0) {
            res = ((res - (SH1 * y) * S(2)) + (SH1 * z) * S(3)) - (SH1 * x) * S(4);
            if (degree > 1) {
                res = ((((res + (SH2C1 * xy) * S(5)) + (SH2C2 * yz) * S(6)) +
                        (SH2C3 * ((RC(2.0) * z2 - x2) - y2)) * S(7)) +
                       (SH2C4 * xz) * S(8)) +
                      (SH2C5 * (x2 - y2)) * S(9);
                if (degree > 2) {
                    res = ((((((res + ((SH3C1 * y) * (RC(3.0) * x2 - y2)) * S(10)) + ((SH3C2 * xy) * z) * S(11)) +
                              ((SH3C3 * y) * ((RC(4.0) * z2 - x2) - y2)) * S(12)) +
                             ((SH3C4 * z) * ((RC(2.0) * z2 - RC(3.0) * x2) - RC(3.0) * y2)) * S(13)) +
                            ((SH3C5 * x) * ((RC(4.0) * z2 - x2) - y2)) * S(14)) +
                           ((SH3C6 * z) * (x2 - y2)) * S(15)) +
                          ((SH3C7 * x) * (x2 - RC(3.0) * y2)) * S(16);
                }
            }
        }
#undef S
        res = (res + RC(0.5)) + EPS32;
        rgb[c] = rmax_((real)0, res);
        clamped[c] = res < (real)0;
    }
}

/* ∇color_from_sh! — spherical_harmonics.jl:76-171.  Writes vshs[0..k) rows; returns vmean (to be ADDED). */
EXPORT void orc_grad_color_from_sh(const real *point, const real *cam, const real *shs, int degree,
                                   const uint8_t *clamped, const real *vcolor_in, real *vshs, real *vmean) {
    real dir_o[3], d[3];
    unit_dir(point, cam, dir_o, d);
    real vc[3];
    for (int c = 0; c < 3; c++) vc[c] = vcolor_in[c] * (RC(1.0) - (real)clamped[c]);
    real x = d[0], y = d[1], z = d[2];
    real x2 = x * x, y2 = y * y, z2 = z * z, xy = x * y, xz = x * z, yz = y * z;
    real dx[3] = {0, 0, 0}, dy[3] = {0, 0, 0}, dz[3] = {0, 0, 0};
#define S(k) shs[3 * ((k) - 1) + c]
#define VS(k) vshs[3 * ((k) - 1) + c]
    for (int c = 0; c < 3; c++) {
        VS(1) = SH0 * vc[c];
        if (degree > 0) {
            VS(2) = ((-SH1) * y) * vc[c];
            VS(3) = (SH1 * z) * vc[c];
            VS(4) = ((-SH1) * x) * vc[c];
            dx[c] = (-SH1) * S(4);
            dy[c] = (-SH1) * S(2);
            dz[c] = SH1 * S(3);
            if (degree > 1) {
                VS(5) = (SH2C1 * xy) * vc[c];
                VS(6) = (SH2C2 * yz) * vc[c];
                VS(7) = (SH2C3 * ((RC(2.0) * z2 - x2) - y2)) * vc[c];
                VS(8) = (SH2C4 * xz) * vc[c];
                VS(9) = (SH2C5 * (x2 - y2)) * vc[c];
                dx[c] = (((dx[c] + (SH2C1 * y) * S(5)) + ((SH2C3 * RC(2.0)) * (-x)) * S(7)) + (SH2C4 * z) * S(8)) +
                        ((SH2C5 * RC(2.0)) * x) * S(9);
                dy[c] = (((dy[c] + (SH2C1 * x) * S(5)) + (SH2C2 * z) * S(6)) + ((SH2C3 * RC(2.0)) * (-y)) * S(7)) +
                        ((SH2C5 * RC(2.0)) * (-y)) * S(9);
                dz[c] = ((dz[c] + (SH2C2 * y) * S(6)) + ((SH2C3 * RC(4.0)) * z) * S(7)) + (SH2C4 * x) * S(8);
                if (degree > 2) {
                    VS(10) = ((SH3C1 * y) * (RC(3.0) * x2 - y2)) * vc[c];
                    VS(11) = ((SH3C2 * xy) * z) * vc[c];
                    VS(12) = ((SH3C3 * y) * ((RC(4.0) * z2 - x2) - y2)) * vc[c];
                    VS(13) = ((SH3C4 * z) * ((RC(2.0) * z2 - RC(3.0) * x2) - RC(3.0) * y2)) * vc[c];
                    VS(14) = ((SH3C5 * x) * ((RC(4.0) * z2 - x2) - y2)) * vc[c];
                    VS(15) = ((SH3C6 * z) * (x2 - y2)) * vc[c];
                    VS(16) = ((SH3C7 * x) * (x2 - RC(3.0) * y2)) * vc[c];
                    dx[c] = ((((((dx[c] + (((SH3C1 * S(10)) * RC(3.0)) * RC(2.0)) * xy) + (SH3C2 * S(11)) * yz) +
                                ((SH3C3 * S(12)) * RC(-2.0)) * xy) +
                               (((SH3C4 * S(13)) * RC(-3.0)) * RC(2.0)) * xz) +
                              (SH3C5 * S(14)) * ((RC(-3.0) * x2 + RC(4.0) * z2) - y2)) +
                             ((SH3C6 * S(15)) * RC(2.0)) * xz) +
                            ((SH3C7 * S(16)) * RC(3.0)) * (x2 - y2);
                    dy[c] = ((((((dy[c] + ((SH3C1 * S(10)) * RC(3.0)) * (x2 - y2)) + (SH3C2 * S(11)) * xz) +
                                (SH3C3 * S(12)) * ((RC(-3.0) * y2 + RC(4.0) * z2) - x2)) +
                               (((SH3C4 * S(13)) * RC(-3.0)) * RC(2.0)) * yz) +
                              ((SH3C5 * S(14)) * RC(-2.0)) * xy) +
                             ((SH3C6 * S(15)) * RC(-2.0)) * yz) +
                            (((SH3C7 * S(16)) * RC(-3.0)) * RC(2.0)) * xy;
                    dz[c] = ((((dz[c] + (SH3C2 * S(11)) * xy) + (((SH3C3 * S(12)) * RC(4.0)) * RC(2.0)) * yz) +
                              ((SH3C4 * S(13)) * RC(3.0)) * ((RC(2.0) * z2 - x2) - y2)) +
                             (((SH3C5 * S(14)) * RC(4.0)) * RC(2.0)) * xz) +
                            (SH3C6 * S(15)) * (x2 - y2);
                }
            }
        }
    }
#undef S
#undef VS
    real vdir[3] = {(dx[0] * vc[0] + dx[1] * vc[1]) + dx[2] * vc[2], (dy[0] * vc[0] + dy[1] * vc[1]) + dy[2] * vc[2],
                    (dz[0] * vc[0] + dz[1] * vc[1]) + dz[2] * vc[2]};
    orc_grad_normalize(dir_o, vdir, vmean);
}

/* ------------------------------------------------------------------ */
/* stage kernels                                                        */
/* ------------------------------------------------------------------ */
typedef struct {
    real R[9];         /* w2c rotation, column-major */
    real t[3];
    real focal[2];
    real principal[2]; /* in [0,1] */
    real cam_center[3];
    int32_t width, height;
} OrcCamera;

typedef struct {
    real near_plane, far_plane, blur_eps;
    int32_t radius_clip;
} OrcConfig;

/* project! — projection.jl:39-130.  normals may be NULL. Culled rows: only radii[i]=0 is written. */
EXPORT void orc_project(int64_t n, const real *means, const real *scales, const real *rots, const OrcCamera *cam,
                        const OrcConfig *cfg, real *depths, int32_t *radii, real *means2d, real *conics,
                        real *normals) {
    int32_t res[2] = {cam->width, cam->height};
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        real mc[3];
        orc_pos_world_to_cam(cam->R, cam->t, means + 3 * i, mc);
        if (!(cfg->near_plane < mc[2] && mc[2] < cfg->far_plane)) { radii[i] = 0; continue; }
        real Rg[9], Sg[9], Sc[9], S2[4], m2[2], S2b[4], S2i[4];
        orc_unnorm_quat2rot(rots + 4 * i, Rg);
        orc_quat_scale_to_cov(Rg, scales + 3 * i, Sg);
        orc_covar_world_to_cam(cam->R, Sg, Sc);
        orc_perspective_projection(mc, Sc, cam->focal, res, cam->principal, S2, m2);
        real det = orc_add_blur(S2, cfg->blur_eps, S2b, NULL);
        if (!(det > (real)0)) { radii[i] = 0; continue; }
        orc_inverse(S2b, S2i);
        real lam = max_eigval_2d(S2b, det);
        int32_t radius = trunc_i32(R_CEIL(RC(3.0) * R_SQRT(lam)));
        if (radius <= cfg->radius_clip) { radii[i] = 0; continue; }
        real rf = (real)radius;
        if ((m2[0] + rf) <= (real)0 || (m2[0] - rf) >= (real)res[0] || (m2[1] + rf) <= (real)0 ||
            (m2[1] - rf) >= (real)res[1]) {
            radii[i] = 0;
            continue;
        }
        radii[i] = radius;
        means2d[2 * i] = m2[0]; means2d[2 * i + 1] = m2[1];
        depths[i] = mc[2];
        conics[3 * i] = S2i[0]; conics[3 * i + 1] = S2i[1]; conics[3 * i + 2] = S2i[3];
        if (normals) orc_gaussian_normal(cam->R, Rg, scales + 3 * i, mc, normals + 3 * i, NULL);
    }
}

/* spherical_harmonics! — spherical_harmonics.jl:1-18.  shs: (3,K,N). */
EXPORT void orc_spherical_harmonics(int64_t n, int K, int degree, const int32_t *radii, const real *means,
                                    const real *cam_center, const real *shs, real *rgbs, uint8_t *clamped) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        if (!(radii[i] > 0)) continue;
        orc_compute_colors_from_sh(means + 3 * i, cam_center, shs + (int64_t)3 * K * i, degree, rgbs + 3 * i,
                                   clamped + 3 * i);
    }
}

/* count_tiles_per_gaussian! — utils.jl:122-142. */
EXPORT void orc_count_tiles(int64_t n, const real *means2d, const int32_t *radii, const int32_t *grid,
                            int32_t *tiles_touched) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        if (!(radii[i] > 0)) { tiles_touched[i] = 0; continue; }
        int32_t rect[4];
        orc_get_rect(means2d + 2 * i, radii[i], grid, rect);
        tiles_touched[i] = (rect[2] - rect[0]) * (rect[3] - rect[1]);
    }
}

/* cumsum! — rasterizer.jl:333-335 (inclusive, Int32).  Returns the last element (n_rendered). */
EXPORT int64_t orc_cumsum(int64_t n, const int32_t *in, int32_t *out) {
    int32_t acc = 0;
    for (int64_t i = 0; i < n; i++) { acc += in[i]; out[i] = acc; }
    return n ? (int64_t)acc : 0;
}

static inline uint32_t depth_bits(real d) {
    float f = (float)d;
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}

/* duplicate_with_keys! — utils.jl:85-120. values are 1-based ids. */
EXPORT void orc_duplicate_with_keys(int64_t n, const real *means2d, const real *depths, const int32_t *offsets,
                                    const int32_t *radii, const int32_t *grid, uint64_t *keys, uint32_t *values) {
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t i = 0; i < n; i++) {
        if (!(radii[i] > 0)) continue;
        int32_t rect[4];
        orc_get_rect(means2d + 2 * i, radii[i], grid, rect);
        uint64_t depth = depth_bits(depths[i]);
        int64_t off = i == 0 ? 0 : offsets[i - 1];
        for (int32_t y = rect[1]; y < rect[3]; y++)
            for (int32_t x = rect[0]; x < rect[2]; x++) {
                uint64_t key = (uint64_t)y * (uint64_t)grid[0] + (uint64_t)x;
                key <<= 32;
                key |= depth;
                keys[off] = key;
                values[off] = (uint32_t)(i + 1);
                off++;
            }
    }
}

/* sortperm! + 2×_permute! — rasterizer.jl:357-372: ascending, ties keep emission order (stable LSD radix). */
EXPORT void orc_sort_pairs(int64_t m, const uint64_t *keys_in, const uint32_t *vals_in, uint64_t *keys_out,
                           uint32_t *vals_out) {
    uint64_t *ka = (uint64_t *)malloc(sizeof(uint64_t) * (m ? m : 1)), *kb = (uint64_t *)malloc(sizeof(uint64_t) * (m ? m : 1));
    uint32_t *va = (uint32_t *)malloc(sizeof(uint32_t) * (m ? m : 1)), *vb = (uint32_t *)malloc(sizeof(uint32_t) * (m ? m : 1));
    memcpy(ka, keys_in, sizeof(uint64_t) * m);
    memcpy(va, vals_in, sizeof(uint32_t) * m);
    for (int pass = 0; pass < 8; pass++) {
        int64_t hist[257] = {0};
        int sh = 8 * pass;
        for (int64_t i = 0; i < m; i++) hist[((ka[i] >> sh) & 0xFF) + 1]++;
        if (hist[((m ? ka[0] : 0) >> sh & 0xFF) + 1] == m) continue; /* all same digit: pass is identity */
        for (int d = 0; d < 256; d++) hist[d + 1] += hist[d];
        for (int64_t i = 0; i < m; i++) {
            int64_t p = hist[(ka[i] >> sh) & 0xFF]++;
            kb[p] = ka[i];
            vb[p] = va[i];
        }
        uint64_t *tk = ka; ka = kb; kb = tk;
        uint32_t *tv = va; va = vb; vb = tv;
    }
    memcpy(keys_out, ka, sizeof(uint64_t) * m);
    memcpy(vals_out, va, sizeof(uint32_t) * m);
    free(ka); free(kb); free(va); free(vb);
}

/* identify_tile_range! — utils.jl:56-78.  ranges: (2,T), caller pre-zeroes (rasterizer.jl:375). */
EXPORT void orc_identify_tile_range(int64_t m, const uint64_t *keys, uint32_t *ranges) {
    for (int64_t i = 1; i <= m; i++) { /* 1-based like the kernel */
        uint32_t tile = (uint32_t)(keys[i - 1] >> 32);
        if (i == 1) {
            ranges[2 * tile] = 0;
        } else {
            uint32_t prev = (uint32_t)(keys[i - 2] >> 32);
            if (tile != prev) {
                ranges[2 * prev + 1] = (uint32_t)(i - 1);
                ranges[2 * tile] = (uint32_t)(i - 1);
            }
        }
        if (i == m) ranges[2 * tile + 1] = (uint32_t)m;
    }
}

/* feature packing — rasterizer.jl:380-391.  channels ∈ {5,8}; runs over ALL rows (stale rows included). */
EXPORT void orc_pack_features(int64_t n, int channels, const real *rgbs, const real *depths, const real *normals,
                              real *features) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        real *f = features + (int64_t)channels * i;
        f[0] = rgbs[3 * i]; f[1] = rgbs[3 * i + 1]; f[2] = rgbs[3 * i + 2];
        f[3] = depths[i];
        f[4] = RC(1.0);
        if (channels > 5) { f[5] = normals[3 * i]; f[6] = normals[3 * i + 1]; f[7] = normals[3 * i + 2]; }
    }
}

/* render! — render.jl:1-130.  One "workgroup" per 16x16 tile; per-pixel sequential loop is equivalent
 * to the rounds-of-256 structure (the position counter `contributor` counts every entry).
 * counts (optional, int64[2]) accumulates evaluated / blended pair counts for the roofline formulas.
 * ambig (optional, per pixel) flags pixels where some pair sits within `ambig_rel` (relative) of one of the
 * kernel's discontinuities (σ<0, α<1/255, T'<1e-4): there a 1-ulp difference in exp() legitimately flips a
 * branch, so parity tests hold those pixels to a looser bound (see tests/parity.py).  ambig_g (optional, per
 * Gaussian) flags the Gaussian of such a pair: its gradient gains or loses that pair's whole contribution.
 * ambig_cond widens the sigma / alpha windows by ambig_cond * (|cb dx dy| + |ca dx^2|/2 + |cc dy^2|/2): a
 * contracted (FMA) evaluation of sigma differs from this one by a few ulps of its largest term, which for
 * elongated Gaussians far exceeds ulps of sigma itself (cancellation).
 * cond (optional, per pixel) = sum over blended pairs of (alpha/(1-alpha) + alpha*T) * (1 + S/2), S = the sum of the
 * magnitudes of sigma's three terms: the first-order bound, in units of the relative error eps of one exp(), of the
 * pixel's colour error when every alpha carries a relative error eps*(1 + S/2) — eps from exp() itself plus half
 * an eps per unit of sigma's largest term (a differently ordered / contracted evaluation of the quadratic form;
 * cancellation makes that dominant for elongated Gaussians).  alpha/(1-alpha) is the conditioning of the
 * transmittance product behind the pair (near-opaque Gaussians, alpha -> 0.99, contribute up to 99 each),
 * alpha*T the pair's own weight. */
EXPORT void orc_render(int channels, int32_t width, int32_t height, const uint32_t *ranges, const uint32_t *values,
                       const real *means2d, const real *opacities, const real *conics, const real *features,
                       const real *background, real *out_color, uint32_t *n_contrib, real *accum_alpha,
                       uint8_t *covis, real *uncert, int64_t *counts, int32_t tile_y0, int32_t tile_y1,
                       uint8_t *ambig, real ambig_rel, uint8_t *ambig_g, real ambig_cond, real *cond) {
    int32_t gx = (width + BLOCK - 1) / BLOCK, gy = (height + BLOCK - 1) / BLOCK;
    if (tile_y1 <= 0 || tile_y1 > gy) tile_y1 = gy;
    if (tile_y0 < 0) tile_y0 = 0;
    int64_t ev_total = 0, bl_total = 0;
#pragma omp parallel for schedule(dynamic, 1) collapse(2) reduction(+ : ev_total, bl_total)
    for (int32_t ty = tile_y0; ty < tile_y1; ty++)
        for (int32_t tx = 0; tx < gx; tx++) {
            uint32_t r0 = ranges[2 * (ty * gx + tx)], r1 = ranges[2 * (ty * gx + tx) + 1];
            for (int32_t ly = 0; ly < BLOCK; ly++)
                for (int32_t lx = 0; lx < BLOCK; lx++) {
                    int32_t px = tx * BLOCK + lx, py = ty * BLOCK + ly;
                    if (!(px < width && py < height)) continue;
                    real T = RC(1.0);
                    uint32_t contributor = 0, last = 0;
                    real color[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    real unc = 0, kappa = 0;
                    for (uint32_t p = r0; p < r1; p++) {
                        contributor++;
                        uint32_t g = values[p] - 1;
                        real dx = means2d[2 * g] - (real)px, dy = means2d[2 * g + 1] - (real)py;
                        const real *cn = conics + 3 * (int64_t)g;
                        real sigma = (cn[1] * dx) * dy + RC(0.5) * (cn[0] * (dx * dx) + cn[2] * (dy * dy));
                        ev_total++;
                        real win = ambig_rel, S = 0;
                        if (ambig || cond) S = R_FABS((cn[1] * dx) * dy) + RC(0.5) * (R_FABS(cn[0] * (dx * dx)) + R_FABS(cn[2] * (dy * dy)));
                        if (ambig) win += ambig_cond * S;
                        if (ambig && R_FABS(sigma) <= win) { ambig[(int64_t)py * width + px] = 1; if (ambig_g) ambig_g[g] = 1; }
                        if (sigma < (real)0) continue;
                        real alpha = rmin_(RC(0.99), opacities[g] * R_EXP(-sigma));
                        if (ambig && R_FABS(alpha * RC(255.0) - RC(1.0)) <= win) { ambig[(int64_t)py * width + px] = 1; if (ambig_g) ambig_g[g] = 1; }
                        if (alpha < RC(1.0) / RC(255.0)) continue;
                        real Tt = T * (RC(1.0) - alpha);
                        if (ambig && R_FABS(Tt * RC(1e4) - RC(1.0)) <= ambig_rel) { ambig[(int64_t)py * width + px] = 1; if (ambig_g) ambig_g[g] = 1; }
                        if (Tt < RC(1e-4)) break;
                        const real *f = features + (int64_t)channels * g;
                        for (int c = 0; c < channels; c++) color[c] += (f[c] * alpha) * T;
                        bl_total++;
                        kappa += (alpha / (RC(1.0) - alpha) + alpha * T) * (RC(1.0) + RC(0.5) * S);
                        if (uncert) unc += alpha * T;
                        if (covis && T > RC(0.5)) covis[g] = 1;
                        T = Tt;
                        last = contributor;
                    }
                    int64_t pi = (int64_t)py * width + px;
                    accum_alpha[pi] = T;
                    n_contrib[pi] = last;
                    for (int c = 0; c < channels; c++) out_color[pi * channels + c] = color[c] + T * background[c];
                    if (uncert) uncert[pi] = unc;
                    if (cond) cond[pi] = kappa;
                }
        }
    if (counts) { counts[0] += ev_total; counts[1] += bl_total; }
}

/* ∇render! — render.jl:132-286.  Atomic targets are double accumulators here (documented oracle choice:
 * the reference's fp32 atomic order is arbitrary; fp64 accumulation is the tie-breaker), then rounded.
 * vcolors (C,N), vopac (N), vconics (3,N), vmeans2d (2,N) — all `double`, caller pre-zeroes. */
EXPORT void orc_grad_render(int channels, int32_t width, int32_t height, const uint32_t *ranges, const uint32_t *values,
                            const real *means2d, const real *opacities, const real *conics, const real *features,
                            const real *background, const real *vpixels, const uint32_t *n_contrib,
                            const real *accum_alpha, double *vcolors, double *vopac, double *vconics,
                            double *vmeans2d, int64_t *counts, int32_t tile_y0, int32_t tile_y1) {
    int32_t gx = (width + BLOCK - 1) / BLOCK, gy = (height + BLOCK - 1) / BLOCK;
    if (tile_y1 <= 0 || tile_y1 > gy) tile_y1 = gy;
    if (tile_y0 < 0) tile_y0 = 0;
    int64_t ev_total = 0, bl_total = 0;
#pragma omp parallel for schedule(dynamic, 1) collapse(2) reduction(+ : ev_total, bl_total)
    for (int32_t ty = tile_y0; ty < tile_y1; ty++)
        for (int32_t tx = 0; tx < gx; tx++) {
            uint32_t r0 = ranges[2 * (ty * gx + tx)], r1 = ranges[2 * (ty * gx + tx) + 1];
            int32_t to_do = (int32_t)(r1 - r0);
            for (int32_t ly = 0; ly < BLOCK; ly++)
                for (int32_t lx = 0; lx < BLOCK; lx++) {
                    int32_t px = tx * BLOCK + lx, py = ty * BLOCK + ly;
                    if (!(px < width && py < height)) continue;
                    int64_t pi = (int64_t)py * width + px;
                    real T_final = accum_alpha[pi], T = T_final;
                    int32_t contributor = to_do, last_contributor = (int32_t)n_contrib[pi];
                    real accum_rec[8] = {0}, last_color[8] = {0}, last_alpha = 0;
                    const real *vpix = vpixels + pi * channels;
                    real bgdot = 0;
                    for (int c = 0; c < channels; c++) bgdot = c == 0 ? background[0] * vpix[0] : bgdot + background[c] * vpix[c];
                    for (int32_t j = 0; j < to_do; j++) {
                        contributor--;
                        if (contributor >= last_contributor) continue;
                        uint32_t g = values[r1 - 1 - j] - 1;
                        real dx = means2d[2 * g] - (real)px, dy = means2d[2 * g + 1] - (real)py;
                        const real *cn = conics + 3 * (int64_t)g;
                        real opacity = opacities[g];
                        real sigma = (cn[1] * dx) * dy + RC(0.5) * (cn[0] * (dx * dx) + cn[2] * (dy * dy));
                        ev_total++;
                        if (sigma < (real)0) continue;
                        real G = R_EXP(-sigma);
                        real alpha = rmin_(RC(0.99), opacity * G);
                        if (alpha < RC(1.0) / RC(255.0)) continue;
                        bl_total++;
                        T = T / (RC(1.0) - alpha);
                        real fac = alpha * T;
                        const real *col = features + (int64_t)channels * g;
                        real va = 0;
                        for (int c = 0; c < channels; c++) {
                            real vc = fac * vpix[c];
#pragma omp atomic
                            vcolors[(int64_t)channels * g + c] += (double)vc;
                            accum_rec[c] = last_alpha * last_color[c] + (RC(1.0) - last_alpha) * accum_rec[c];
                            last_color[c] = col[c];
                            va += (col[c] - accum_rec[c]) * vpix[c];
                        }
                        va *= T;
                        va += ((-T_final) / (RC(1.0) - alpha)) * bgdot;
                        last_alpha = alpha;
                        real vs = ((-opacity) * G) * va;
                        real vcn0 = (RC(0.5) * vs) * (dx * dx), vcn1 = ((RC(0.5) * vs) * dx) * dy,
                             vcn2 = (RC(0.5) * vs) * (dy * dy);
                        real vx = vs * (cn[0] * dx + cn[1] * dy), vy = vs * (cn[1] * dx + cn[2] * dy);
                        real vo = G * va;
#pragma omp atomic
                        vmeans2d[2 * (int64_t)g] += (double)vx;
#pragma omp atomic
                        vmeans2d[2 * (int64_t)g + 1] += (double)vy;
#pragma omp atomic
                        vconics[3 * (int64_t)g] += (double)vcn0;
#pragma omp atomic
                        vconics[3 * (int64_t)g + 1] += (double)vcn1;
#pragma omp atomic
                        vconics[3 * (int64_t)g + 2] += (double)vcn2;
#pragma omp atomic
                        vopac[g] += (double)vo;
                    }
                }
        }
    if (counts) { counts[0] += ev_total; counts[1] += bl_total; }
}

/* ∇project! — projection.jl:132-257.  vdepths / vnormals / vR_out / vt_out may be NULL.
 * Rows with radii<=0 are left untouched (outputs rely on zero-init, projection.jl:172-176).
 * Pose gradients use double accumulators (atomics in the reference), filtered at |v|>1e-7. */
EXPORT void orc_grad_project(int64_t n, const real *vmeans2d, const real *vconics, const real *vdepths,
                             const real *vnormals, const real *conics, const int32_t *radii, const real *means,
                             const real *scales, const real *rots, const OrcCamera *cam, real *vmeans, real *vscales,
                             real *vrots, double *vR_out, double *vt_out) {
    int32_t res[2] = {cam->width, cam->height};
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        if (!(radii[i] > 0)) continue;
        const real *cn = conics + 3 * i, *vc = vconics + 3 * i;
        real S2i[4] = {cn[0], cn[1], cn[1], cn[2]};
        real vS2i[4] = {vc[0], vc[1], vc[1], vc[2]};
        real vS2[4];
        orc_grad_inverse(S2i, vS2i, vS2);
        real mc[3], Rg[9], Sg[9], Sc[9];
        orc_pos_world_to_cam(cam->R, cam->t, means + 3 * i, mc);
        orc_unnorm_quat2rot(rots + 4 * i, Rg);
        orc_quat_scale_to_cov(Rg, scales + 3 * i, Sg);
        orc_covar_world_to_cam(cam->R, Sg, Sc);
        real vSc[9], vmc[3];
        orc_grad_perspective_projection(mc, Sc, cam->focal, res, cam->principal, vS2, vmeans2d + 2 * i, vSc, vmc);
        if (vdepths) vmc[2] = vmc[2] + vdepths[i];
        real vR[9], vt[3], vmean[3], vSg[9];
        orc_grad_pos_world_to_cam(cam->R, cam->t, means + 3 * i, vmc, vR, vt, vmean);
        orc_grad_covar_world_to_cam(cam->R, Sg, vSc, vR, vSg);
        real vRg[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (vnormals) {
            real nrm[3], sign;
            int32_t k = orc_gaussian_normal(cam->R, Rg, scales + 3 * i, mc, nrm, &sign);
            real Rt[9], g[3];
            transpose33(cam->R, Rt);
            mulvec3(Rt, vnormals + 3 * i, g);
            for (int r = 0; r < 3; r++) M3(vRg, r, k - 1) = sign * g[r];
        }
        real vq[4], vs[3];
        orc_grad_quat_scale_to_cov(rots + 4 * i, scales + 3 * i, Rg, vSg, vRg, vq, vs);
        for (int k = 0; k < 3; k++) { vmeans[3 * i + k] = vmean[k]; vscales[3 * i + k] = vs[k]; }
        for (int k = 0; k < 4; k++) vrots[4 * i + k] = vq[k];
        if (vR_out) {
            for (int rr = 0; rr < 3; rr++) {
                for (int rc = 0; rc < 3; rc++) {
                    real v = M3(vR, rr, rc);
                    if (R_FABS(v) > RC(1e-7)) {
#pragma omp atomic
                        vR_out[rr + 3 * rc] += (double)v;
                    }
                }
                real v = vt[rr];
                if (R_FABS(v) > RC(1e-7)) {
#pragma omp atomic
                    vt_out[rr] += (double)v;
                }
            }
        }
    }
}

/* ∇spherical_harmonics! — spherical_harmonics.jl:20-38.  No radii guard; vmeans[i] += vmean. vshs (3,K,N). */
EXPORT void orc_grad_spherical_harmonics(int64_t n, int K, int degree, const real *means, const real *cam_center,
                                         const real *shs, const uint8_t *clamped, const real *vcolors, int vc_stride,
                                         real *vshs, real *vmeans) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        real vm[3];
        orc_grad_color_from_sh(means + 3 * i, cam_center, shs + (int64_t)3 * K * i, degree, clamped + 3 * i,
                               vcolors + (int64_t)vc_stride * i, vshs + (int64_t)3 * K * i, vm);
        for (int k = 0; k < 3; k++) vmeans[3 * i + k] += vm[k];
    }
}

/* _update_stats! — strategy.jl:118-136. */
EXPORT void orc_update_stats(int64_t n, const int32_t *radii, const real *vmeans2d, uint32_t width, uint32_t height,
                             int32_t *max_radii, real *accum, real *denom) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        int32_t r = radii[i];
        if (!(r > 0)) continue;
        max_radii[i] = max_radii[i] > r ? max_radii[i] : r;
        real gx = (vmeans2d[2 * i] * (real)width) * RC(0.5), gy = (vmeans2d[2 * i + 1] * (real)height) * RC(0.5);
        accum[i] += R_SQRT(gx * gx + gy * gy);
        denom[i] += RC(1.0);
    }
}

/* ---------------------------------------------------------------------------------------------------------
 * Fused SSIM — fused_ssim.jl.  Arrays are the reference's (W,H,CH,B) column-major, i.e.
 * planar [b][c][y][x].  11-tap separable window (fused_ssim.jl:11-24: sigma = 1.5, normalised; the literals
 * are the reference's float32 constants — entry 4 is one ulp below the correctly rounded formula value, so
 * they are data, not derivable), zero padding outside the image (get_pix_value, :27-31).
 * Op order of the two passes (:83-112, :170-205): pairs (left + right) * w for d = 1..5 with w = G[5 - d],
 * then the centre tap; squares and products are formed per tap before the pair sum.
 * ------------------------------------------------------------------------------------------------------- */
static const float SSIM_G[11] = {0.001028380123898387f, 0.0075987582094967365f, 0.036000773310661316f,
                                 0.10936068743467331f,  0.21300552785396576f,  0.26601171493530273f,
                                 0.21300552785396576f,  0.10936068743467331f,  0.036000773310661316f,
                                 0.0075987582094967365f, 0.001028380123898387f};

static inline real ssim_pix(const real *plane, int32_t W, int32_t H, int32_t y, int32_t x) {
    return (x < 0 || x >= W || y < 0 || y >= H) ? (real)0 : plane[(int64_t)y * W + x];
}

/* _fused_ssim! — fused_ssim.jl:34-258.  ssim_map always; the three partial-derivative maps when train != 0. */
EXPORT void orc_fused_ssim(int32_t W, int32_t H, int32_t CH, int32_t B, const real *img, const real *ref, real C1,
                           real C2, int train, real *ssim_map, real *dm_dmu1, real *dm_dsigma1_sq, real *dm_dsigma12) {
    const int64_t plane = (int64_t)W * H;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t pc = 0; pc < (int64_t)CH * B; pc++) {
        const real *X = img + pc * plane, *Y = ref + pc * plane;
        /* horizontal pass on rows -5 .. H+4 (rows outside the image are all-zero inputs -> exact zeros) */
        real *xc = (real *)malloc(sizeof(real) * 5 * (size_t)(H + 10) * W);
        for (int32_t yy = -5; yy < H + 5; yy++)
            for (int32_t x = 0; x < W; x++) {
                real sX = 0, sX2 = 0, sY = 0, sY2 = 0, sXY = 0;
                for (int d = 1; d <= 5; d++) {
                    real w = (real)SSIM_G[5 - d];
                    real Xl = ssim_pix(X, W, H, yy, x - d), Yl = ssim_pix(Y, W, H, yy, x - d);
                    real Xr = ssim_pix(X, W, H, yy, x + d), Yr = ssim_pix(Y, W, H, yy, x + d);
                    sX += (Xl + Xr) * w;
                    sX2 += (Xl * Xl + Xr * Xr) * w;
                    sY += (Yl + Yr) * w;
                    sY2 += (Yl * Yl + Yr * Yr) * w;
                    sXY += (Xl * Yl + Xr * Yr) * w;
                }
                real cx = ssim_pix(X, W, H, yy, x), cy = ssim_pix(Y, W, H, yy, x), wc = (real)SSIM_G[5];
                sX += cx * wc;
                sX2 += cx * cx * wc;
                sY += cy * wc;
                sY2 += cy * cy * wc;
                sXY += cx * cy * wc;
                real *o = xc + 5 * ((int64_t)(yy + 5) * W + x);
                o[0] = sX; o[1] = sX2; o[2] = sY; o[3] = sY2; o[4] = sXY;
            }
        for (int32_t y = 0; y < H; y++)
            for (int32_t x = 0; x < W; x++) {
                real out[5] = {0, 0, 0, 0, 0};
                for (int d = 1; d <= 5; d++) {
                    real w = (real)SSIM_G[5 - d];
                    const real *top = xc + 5 * ((int64_t)(y + 5 - d) * W + x), *bot = xc + 5 * ((int64_t)(y + 5 + d) * W + x);
                    for (int k = 0; k < 5; k++) out[k] += (top[k] + bot[k]) * w;
                }
                const real *ctr = xc + 5 * ((int64_t)(y + 5) * W + x);
                for (int k = 0; k < 5; k++) out[k] += ctr[k] * (real)SSIM_G[5];
                real mu1 = out[0], mu2 = out[2];
                real mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2;
                real sigma1_sq = out[1] - mu1_sq, sigma2_sq = out[3] - mu2_sq, sigma12 = out[4] - mu1 * mu2;
                real A = mu1_sq + mu2_sq + C1, Bv = sigma1_sq + sigma2_sq + C2;
                real Cv = RC(2.0) * mu1 * mu2 + C1, Dv = RC(2.0) * sigma12 + C2;
                int64_t o = pc * plane + (int64_t)y * W + x;
                ssim_map[o] = (Cv * Dv) / (A * Bv);
                if (train) {
                    dm_dmu1[o] = (mu2 * RC(2.0) * Dv) / (A * Bv) - (mu2 * RC(2.0) * Cv) / (A * Bv) -
                                 (mu1 * RC(2.0) * Cv * Dv) / (A * A * Bv) + (mu1 * RC(2.0) * Cv * Dv) / (A * Bv * Bv);
                    dm_dsigma1_sq[o] = (-Cv * Dv) / (A * Bv * Bv);
                    dm_dsigma12[o] = (RC(2.0) * Cv) / (A * Bv);
                }
            }
        free(xc);
    }
}

/* _fused_ssim_bwd! — fused_ssim.jl:261-352.  dL_dimg = conv(dm_dmu1*dL) + 2*img*conv(dm_dsigma1_sq*dL) + ref*conv(dm_dsigma12*dL). */
EXPORT void orc_fused_ssim_bwd(int32_t W, int32_t H, int32_t CH, int32_t B, const real *img, const real *ref,
                               const real *dL_dmap, const real *dm_dmu1, const real *dm_dsigma1_sq,
                               const real *dm_dsigma12, real *dL_dimg) {
    const int64_t plane = (int64_t)W * H;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t pc = 0; pc < (int64_t)CH * B; pc++) {
        const real *L = dL_dmap + pc * plane, *M0 = dm_dmu1 + pc * plane, *M1 = dm_dsigma1_sq + pc * plane,
                   *M2 = dm_dsigma12 + pc * plane;
        real *sc = (real *)malloc(sizeof(real) * 3 * (size_t)(H + 10) * W);
        for (int32_t yy = -5; yy < H + 5; yy++)
            for (int32_t x = 0; x < W; x++) {
                real a[3] = {0, 0, 0};
                for (int d = 1; d <= 5; d++) {
                    real w = (real)SSIM_G[5 - d];
                    real cl = ssim_pix(L, W, H, yy, x - d), cr = ssim_pix(L, W, H, yy, x + d);
                    real l0 = ssim_pix(M0, W, H, yy, x - d) * cl, r0 = ssim_pix(M0, W, H, yy, x + d) * cr;
                    real l1 = ssim_pix(M1, W, H, yy, x - d) * cl, r1 = ssim_pix(M1, W, H, yy, x + d) * cr;
                    real l2 = ssim_pix(M2, W, H, yy, x - d) * cl, r2 = ssim_pix(M2, W, H, yy, x + d) * cr;
                    a[0] += (l0 + r0) * w;
                    a[1] += (l1 + r1) * w;
                    a[2] += (l2 + r2) * w;
                }
                real cc = ssim_pix(L, W, H, yy, x), wc = (real)SSIM_G[5];
                a[0] += (ssim_pix(M0, W, H, yy, x) * cc) * wc;
                a[1] += (ssim_pix(M1, W, H, yy, x) * cc) * wc;
                a[2] += (ssim_pix(M2, W, H, yy, x) * cc) * wc;
                real *o = sc + 3 * ((int64_t)(yy + 5) * W + x);
                o[0] = a[0]; o[1] = a[1]; o[2] = a[2];
            }
        for (int32_t y = 0; y < H; y++)
            for (int32_t x = 0; x < W; x++) {
                real s[3] = {0, 0, 0};
                for (int d = 1; d <= 5; d++) {
                    real w = (real)SSIM_G[5 - d];
                    const real *top = sc + 3 * ((int64_t)(y + 5 - d) * W + x), *bot = sc + 3 * ((int64_t)(y + 5 + d) * W + x);
                    for (int k = 0; k < 3; k++) s[k] += (top[k] + bot[k]) * w;
                }
                const real *ctr = sc + 3 * ((int64_t)(y + 5) * W + x);
                for (int k = 0; k < 3; k++) s[k] += ctr[k] * (real)SSIM_G[5];
                int64_t o = pc * plane + (int64_t)y * W + x;
                dL_dimg[o] = s[0] + RC(2.0) * img[o] * s[1] + ref[o] * s[2];
            }
        free(sc);
    }
}


EXPORT int orc_real_size(void) { return (int)sizeof(real); }
