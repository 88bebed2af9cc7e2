// preprocess.cu — fused per-Gaussian forward stage for sm_100a.
//
// One kernel replaces the reference's project! (src/rasterization/projection.jl:39-130),
// spherical_harmonics! (spherical_harmonics.jl:1-18), count_tiles_per_gaussian! (utils.jl:122-142) and the
// feature-packing broadcasts (rasterizer.jl:380-391): one pass over means/scales/rotations/SH instead of
// three launches that each re-read `means`, plus 3-4 broadcast kernels.
//
// THIS TRANSLATION UNIT IS COMPILED WITH -fmad=false: radii, means_2d, depths, conics, rgbs, clamped and
// tiles_touched must be bit-exact with the reference op order, which never
// contracts a*b+c.  Products are written out in StaticArrays' unrolled left-to-right order, including the
// structural zeros of diag(s) and J (x + 0*y is only an identity up to the sign of zero).
//
// Memory behaviour (HBM-bound stage): rotations are 128-bit loads; the (3,K,n) SH block of a 128-Gaussian
// CTA is one contiguous span, copied with coalesced 128-bit loads into padded shared memory (row stride odd
// -> conflict-free per-thread reads) only when the CTA has a visible Gaussian; the packed 48/64-byte record
// the compositing kernels stream is written with 128-bit stores.
#include "common.cuh"

#define PP_THREADS 128

namespace {

__device__ __forceinline__ void mul33(const float *A, const float *B, float *C) {  // column-major, C = A*B
    float T[9];
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++)
            T[i + 3 * j] = (A[i] * B[3 * j] + A[i + 3] * B[1 + 3 * j]) + A[i + 6] * B[2 + 3 * j];
#pragma unroll
    for (int k = 0; k < 9; k++) C[k] = T[k];
}
__device__ __forceinline__ void transpose33(const float *A, float *T) {
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++) T[i + 3 * j] = A[j + 3 * i];
}

// SH constants — utils.jl:33-48
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
#define EPS32 1.1920929e-07f

template <bool ALIGNED16, bool RAW>
__global__ void __launch_bounds__(PP_THREADS, 6)
preprocess_kernel(const DevCamera cam, const int64_t n, const int sh_degree, const int K, const int channels,
                  const float *__restrict__ means, const float *__restrict__ shs, const float *__restrict__ opac,
                  const float *__restrict__ scales, const float *__restrict__ rots, const GeomPtrs g,
                  const int sh_stride, const ParamSpec ps) {
    extern __shared__ float s_sh[];  // [PP_THREADS][sh_stride]
    const int tid = threadIdx.x;
    const int64_t block0 = (int64_t)blockIdx.x * PP_THREADS;
    const int64_t i = block0 + tid;
    const bool in_range = i < n;

    float R[9], t[3];
    if (cam.R_dev) {  // device-resident pose (pose optimisation path, projection.jl:71-75 / utils.jl:7-12)
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

    int32_t radius = 0;
    float mean[3] = {0.f, 0.f, 0.f}, mc[3] = {0.f, 0.f, 1.f}, m2[2] = {0.f, 0.f}, conic[3] = {0.f, 0.f, 0.f};
    float nrm[3] = {0.f, 0.f, 0.f};
    if (in_range) {
        mean[0] = means[3 * i]; mean[1] = means[3 * i + 1]; mean[2] = means[3 * i + 2];
        // pos_world_to_cam: R*p + t  (projection.jl:355-361)
#pragma unroll
        for (int r = 0; r < 3; r++) mc[r] = ((R[r] * mean[0] + R[r + 3] * mean[1]) + R[r + 6] * mean[2]) + t[r];
        if (cam.near_plane < mc[2] && mc[2] < cam.far_plane) {  // projection.jl:79
            const float4 q4 = *reinterpret_cast<const float4 *>(rots + 4 * i);  // 128-bit load (simd.jl:1-11)
            float sc[3];
            if (RAW && ps.isotropic) {
                sc[0] = sc[1] = sc[2] = expf(scales[i]);  // rasterizer.jl:240-243
            } else {
                sc[0] = scales[3 * i]; sc[1] = scales[3 * i + 1]; sc[2] = scales[3 * i + 2];
                if (RAW && ps.raw_scale) { sc[0] = expf(sc[0]); sc[1] = expf(sc[1]); sc[2] = expf(sc[2]); }
            }
            // unnorm_quat2rot (render.jl:322-333): normalize(q) = inv(norm(q)) * q
            const float qn = sqrtf(((q4.x * q4.x + q4.y * q4.y) + q4.z * q4.z) + q4.w * q4.w);
            const float qi = 1.0f / qn;
            const float w = qi * q4.x, x = qi * q4.y, y = qi * q4.z, z = qi * q4.w;
            const float x2 = x * x, y2 = y * y, z2 = z * z, xy = x * y, xz = x * z, yz = y * z;
            const float wx = w * x, wy = w * y, wz = w * z;
            float Rg[9];
            Rg[0] = 1.0f - 2.0f * (y2 + z2); Rg[1] = 2.0f * (xy + wz); Rg[2] = 2.0f * (xz - wy);
            Rg[3] = 2.0f * (xy - wz); Rg[4] = 1.0f - 2.0f * (x2 + z2); Rg[5] = 2.0f * (yz + wx);
            Rg[6] = 2.0f * (xz + wy); Rg[7] = 2.0f * (yz - wx); Rg[8] = 1.0f - 2.0f * (x2 + y2);
            // quat_scale_to_cov (render.jl:291-294): M = R*diag(s); Σ = M*M'
            const float S[9] = {sc[0], 0.f, 0.f, 0.f, sc[1], 0.f, 0.f, 0.f, sc[2]};
            float M[9], Mt[9], Sg[9], Sc[9], Rt[9], T1[9];
            mul33(Rg, S, M);
            transpose33(M, Mt);
            mul33(M, Mt, Sg);
            // covar_world_to_cam (projection.jl:375-380): (R*Σ)*R'
            mul33(R, Sg, T1);
            transpose33(R, Rt);
            mul33(T1, Rt, Sc);
            // perspective_projection (projection.jl:259-287)
            const float res[2] = {(float)cam.width, (float)cam.height};
            float pp[2], lim[2], limn[2], txy[2];
            const float rz = 1.0f / mc[2];
            const float rz2 = rz * rz;
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const float tan_fov = (0.5f * res[k]) / cam.focal[k];
                const float stf = 0.3f * tan_fov;
                pp[k] = cam.principal[k] * res[k];
                lim[k] = (res[k] - pp[k]) / cam.focal[k] + stf;
                limn[k] = pp[k] / cam.focal[k] + stf;
                m2[k] = (rz * cam.focal[k]) * mc[k] + pp[k];
                txy[k] = mc[2] * fminf(lim[k], fmaxf(-limn[k], mc[k] * rz));
            }
            const float J[6] = {cam.focal[0] * rz, 0.f, 0.f, cam.focal[1] * rz,
                                ((-cam.focal[0]) * txy[0]) * rz2, ((-cam.focal[1]) * txy[1]) * rz2};
            float TJ[6], S2[4];
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int r = 0; r < 2; r++)
                    TJ[r + 2 * j] = (J[r] * Sc[3 * j] + J[r + 2] * Sc[1 + 3 * j]) + J[r + 4] * Sc[2 + 3 * j];
#pragma unroll
            for (int j = 0; j < 2; j++)
#pragma unroll
                for (int r = 0; r < 2; r++) S2[r + 2 * j] = (TJ[r] * J[j] + TJ[r + 2] * J[j + 2]) + TJ[r + 4] * J[j + 4];
            // add_blur (render.jl:387-396)
            const float a = S2[0] + cam.blur_eps, d = S2[3] + cam.blur_eps, b21 = S2[1], b12 = S2[2];
            const float det = a * d - b12 * b21;
            if (det > 0.0f) {  // projection.jl:94
                // inverse (render.jl:368-381)
                const float det_inv = 1.0f / det;
                const float tmp = (-b12) * det_inv;
                // max_eigval_2D (render.jl:415-420), radius (projection.jl:102-103)
                const float mid = 0.5f * (a + d);
                const float lam = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
                const int32_t rad = __float2int_rz(ceilf(3.0f * sqrtf(lam)));
                if (rad > cam.radius_clip) {
                    const float rf = (float)rad;
                    const bool off = (m2[0] + rf) <= 0.0f || (m2[0] - rf) >= res[0] || (m2[1] + rf) <= 0.0f ||
                                     (m2[1] - rf) >= res[1];  // projection.jl:110-118
                    if (!off) {
                        radius = rad;
                        conic[0] = d * det_inv; conic[1] = tmp; conic[2] = a * det_inv;
                        if (channels > 5) {  // gaussian_normal (projection.jl:14-27)
                            const int k = (sc[0] <= sc[1] && sc[0] <= sc[2]) ? 0 : ((sc[1] <= sc[2]) ? 1 : 2);
                            const float ax[3] = {k == 0 ? Rg[0] : (k == 1 ? Rg[3] : Rg[6]),
                                                 k == 0 ? Rg[1] : (k == 1 ? Rg[4] : Rg[7]),
                                                 k == 0 ? Rg[2] : (k == 1 ? Rg[5] : Rg[8])};
                            float nc[3];
#pragma unroll
                            for (int r = 0; r < 3; r++) nc[r] = (R[r] * ax[0] + R[r + 3] * ax[1]) + R[r + 6] * ax[2];
                            const float dd = (nc[0] * mc[0] + nc[1] * mc[1]) + nc[2] * mc[2];
                            const float sign = dd > 0.0f ? -1.0f : 1.0f;
                            nrm[0] = sign * nc[0]; nrm[1] = sign * nc[1]; nrm[2] = sign * nc[2];
                        }
                    }
                }
            }
        }
    }
    const bool visible = radius > 0;

    // ---- SH coefficients into shared memory (only if this CTA renders something) -------------------------
    const int k_used = (sh_degree + 1) * (sh_degree + 1);
    const bool any_visible = __syncthreads_or(visible);
    if (any_visible) {
        if (k_used == K) {
            const int nb = (int)((n - block0) < PP_THREADS ? (n - block0) : PP_THREADS);
            if (RAW && ps.sh_rest) {  // features_dc | features_rest kept apart by the caller: same rows, two spans
                rows_global_to_shared(shs + block0 * 3, s_sh, nb, 3, sh_stride, tid, PP_THREADS, false);
                if (K > 1)
                    rows_global_to_shared(ps.sh_rest + block0 * (int64_t)(3 * (K - 1)), s_sh + 3, nb, 3 * (K - 1), sh_stride, tid,
                                          PP_THREADS, (reinterpret_cast<uintptr_t>(ps.sh_rest) & 15) == 0);
            } else {
                rows_global_to_shared(shs + block0 * (int64_t)(3 * K), s_sh, nb, 3 * K, sh_stride, tid, PP_THREADS, ALIGNED16);
            }
        } else if (visible) {  // k_used < K: only the leading coefficients are needed — direct strided loads
            if (RAW && ps.sh_rest) {
                for (int e = 0; e < 3; e++) s_sh[tid * sh_stride + e] = shs[3 * i + e];
                const float *src = ps.sh_rest + i * (int64_t)(3 * (K - 1));
                for (int e = 3; e < 3 * k_used; e++) s_sh[tid * sh_stride + e] = src[e - 3];
            } else {
                const float *src = shs + i * (int64_t)(3 * K);
                for (int e = 0; e < 3 * k_used; e++) s_sh[tid * sh_stride + e] = src[e];
            }
        }
    }
    __syncthreads();

    if (!in_range) return;
    g.radii[i] = radius;  // always written (projection.jl:80,95,105,116,120)
    if (!visible) {
        g.tiles_touched[i] = 0;  // utils.jl:132-135
        return;
    }

    // compute_colors_from_sh (spherical_harmonics.jl:41-74)
    float rgb[3];
    uint8_t cl[3];
    {
        float x = 0.f, y = 0.f, z = 0.f;
        if (sh_degree > 0) {
            const float d0 = mean[0] - cam.cam_center[0], d1 = mean[1] - cam.cam_center[1],
                        d2 = mean[2] - cam.cam_center[2];
            const float inv = 1.0f / sqrtf((d0 * d0 + d1 * d1) + d2 * d2);
            x = inv * d0; y = inv * d1; z = inv * d2;
        }
        const float x2 = x * x, y2 = y * y, z2 = z * z, xy = x * y, xz = x * z, yz = y * z;
        const float *sh = s_sh + tid * sh_stride;
#pragma unroll
        for (int c = 0; c < 3; c++) {
#define S(k) sh[3 * ((k) - 1) + c]
            float res = SH0 * S(1);
            if (sh_degree > 0) {
                res = ((res - (SH1 * y) * S(2)) + (SH1 * z) * S(3)) - (SH1 * x) * S(4);
                if (sh_degree > 1) {
                    res = ((((res + (SH2C1 * xy) * S(5)) + (SH2C2 * yz) * S(6)) +
                            (SH2C3 * ((2.0f * z2 - x2) - y2)) * S(7)) +
                           (SH2C4 * xz) * S(8)) +
                          (SH2C5 * (x2 - y2)) * S(9);
                    if (sh_degree > 2) {
                        res = ((((((res + ((SH3C1 * y) * (3.0f * x2 - y2)) * S(10)) + ((SH3C2 * xy) * z) * S(11)) +
                                  ((SH3C3 * y) * ((4.0f * z2 - x2) - y2)) * S(12)) +
                                 ((SH3C4 * z) * ((2.0f * z2 - 3.0f * x2) - 3.0f * y2)) * S(13)) +
                                ((SH3C5 * x) * ((4.0f * z2 - x2) - y2)) * S(14)) +
                               ((SH3C6 * z) * (x2 - y2)) * S(15)) +
                              ((SH3C7 * x) * (x2 - 3.0f * y2)) * S(16);
                    }
                }
            }
#undef S
            res = (res + 0.5f) + EPS32;
            rgb[c] = fmaxf(0.0f, res);
            cl[c] = res < 0.0f;
        }
    }

    // tiles touched (utils.jl:122-142)
    int32_t x0, y0, x1, y1;
    get_rect(m2[0], m2[1], radius, cam.grid_x, cam.grid_y, x0, y0, x1, y1);
    g.tiles_touched[i] = (x1 - x0) * (y1 - y0);

    g.means2d[i] = make_float2(m2[0], m2[1]);
    g.depths[i] = mc[2];
    g.conics[3 * i] = conic[0]; g.conics[3 * i + 1] = conic[1]; g.conics[3 * i + 2] = conic[2];
    g.rgbs[3 * i] = rgb[0]; g.rgbs[3 * i + 1] = rgb[1]; g.rgbs[3 * i + 2] = rgb[2];
    g.clamped[3 * i] = cl[0]; g.clamped[3 * i + 1] = cl[1]; g.clamped[3 * i + 2] = cl[2];
    if (channels > 5) { g.normals[3 * i] = nrm[0]; g.normals[3 * i + 1] = nrm[1]; g.normals[3 * i + 2] = nrm[2]; }

    // packed record for the compositing kernels; features = rgb, depth, 1, normal (rasterizer.jl:380-386)
    const float o = (RAW && ps.raw_opacity) ? act_sigmoid(opac[i]) : opac[i];
    const int RQ = rec_quads(channels);
    float4 *rec = g.rec + i * RQ;
    rec[0] = make_float4(m2[0], m2[1], conic[0], conic[1]);
    rec[1] = make_float4(conic[2], o, rgb[0], rgb[1]);
    if (channels == 3) {
        rec[2] = make_float4(rgb[2], 0.f, 0.f, 0.f);
    } else {
        rec[2] = make_float4(rgb[2], mc[2], 1.0f, channels > 5 ? nrm[0] : 0.f);
        if (channels > 5) rec[3] = make_float4(nrm[1], nrm[2], 0.f, 0.f);
    }
}

}  // namespace

void launch_preprocess(const DevCamera &cam, int64_t n, int sh_degree, int K, int channels, const float *means,
                       const float *shs, const float *opac, const float *scales, const float *rots,
                       const GeomPtrs &g, cudaStream_t s, const ParamSpec &ps) {
    if (n <= 0) return;
    int stride = 3 * K;
    if ((stride & 1) == 0) stride += 1;  // odd row stride: conflict-free per-thread reads
    const size_t smem = (size_t)PP_THREADS * stride * sizeof(float);
    const int64_t blocks = (n + PP_THREADS - 1) / PP_THREADS;
    const bool aligned = (reinterpret_cast<uintptr_t>(shs) & 15) == 0;
    const bool raw = ps.raw_opacity || ps.raw_scale || ps.isotropic || ps.sh_rest;
#define GSR_PP(AL, RW)                                                                                                   \
    preprocess_kernel<AL, RW><<<(unsigned)blocks, PP_THREADS, smem, s>>>(cam, n, sh_degree, K, channels, means, shs, opac, \
                                                                       scales, rots, g, stride, ps)
    if (aligned) {
        if (raw) GSR_PP(true, true); else GSR_PP(true, false);
    } else {
        if (raw) GSR_PP(false, true); else GSR_PP(false, false);
    }
#undef GSR_PP
    count_launch();
}
