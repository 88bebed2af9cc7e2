// ssim.cu — fused SSIM forward / backward and the L1 + D-SSIM photometric loss for sm_100a.
//
// Replaces _fused_ssim! (src/fused_ssim.jl:34-258), _fused_ssim_bwd! (:261-352) and, for the fused loss entry
// point, the slice / permutedims / mean|x-t| / fused_ssim / mean chain of Trainer.step! (src/training.jl:684-694)
// together with its Zygote pullback down to the cotangent of the raster image.
//
// One CTA = one 32 x 32 pixel tile of one (channel, batch) plane, 256 threads.  The 11-tap separable window runs
// out of shared memory in two passes exactly as the reference's (pairs (left + right) * w for d = 1..5, centre
// tap last), zero padding outside the image.  The kernels are FP32-issue bound (~270 flops per pixel-channel
// against 24 bytes), so the shape minimises issued instructions per output:
//   * 32 x 32 tiles: every global row access of a warp is one 128-byte line; halo overhead 1.7x instead of the
//     2.6x of 16 x 16 tiles, and only 1.31x of the horizontal pass is spent on halo rows;
//   * horizontal pass: a thread produces two ADJACENT columns from 12 inputs per array (64-bit shared loads), so
//     the squares / products x*x, y*y, x*y are formed once per input instead of once per tap;
//   * vertical pass: a thread produces four ADJACENT rows, which share 10 of their 11 input rows (17.5 instead
//     of 55 shared-memory loads per output);
//   * inputs and outputs are addressed through element strides, so the same kernels read the rasterizer's
//     interleaved (C,W,H) image and write its (C,W,H) cotangent directly — the reference permutes to (W,H,C,1)
//     and back with separate passes;
//   * the loss variant folds in what surrounds the SSIM in the trainer: the sums of |x - t| and of the SSIM map
//     (block-reduced, one double atomic each per CTA — the map itself is never written), the constant
//     dL/dmap = -lambda/n, and the L1 pullback (1 - lambda)/n * sign(x - t) in the backward's epilogue.
// Floating point, tolerance-checked (1e-5 on the map, 1e-4 relative on gradients): FMA contraction allowed.
#include "common.cuh"

namespace {

constexpr int SS_X = 32, SS_Y = 32, SS_HALO = 5;
constexpr int SS_TW = SS_X + 2 * SS_HALO;  // 42 columns staged
constexpr int SS_TP = 44;                   // row pitch (floats): 8-byte aligned pairs
constexpr int SS_TH = SS_Y + 2 * SS_HALO;  // 42 rows staged
constexpr int SS_VR = 4;                    // adjacent output rows per thread in the vertical pass
constexpr int SS_THREADS = 256;

// fused_ssim.jl:11-24 — the reference's float32 taps (sigma = 1.5; entry 4 is one ulp below the formula value)
__constant__ float c_gauss[11] = {0.001028380123898387f, 0.0075987582094967365f, 0.036000773310661316f,
                                  0.10936068743467331f,  0.21300552785396576f,  0.26601171493530273f,
                                  0.21300552785396576f,  0.10936068743467331f,  0.036000773310661316f,
                                  0.0075987582094967365f, 0.001028380123898387f};

struct Strides {  // element strides of a (x, y, channel, batch) indexed array
    int64_t x, y, c, b;
};
__host__ __device__ inline Strides planar(int W, int H, int CH) {
    return Strides{1, W, (int64_t)W * H, (int64_t)W * H * CH};
}

__device__ __forceinline__ float block_sum(float v, float *s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < SS_THREADS / 32) t = s_red[threadIdx.x];
    if (warp == 0) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    return t;  // valid in thread 0
}

// TRAIN: write the three partial-derivative maps.  LOSS: accumulate sum(ssim) and sum|x - t| into acc[0], acc[1]
// (ssim_map may be null).
template <bool TRAIN, bool LOSS>
__global__ void __launch_bounds__(SS_THREADS)
ssim_fwd_kernel(const int W, const int H, const int CH, const float *__restrict__ img, const Strides si,
                const float *__restrict__ ref, const Strides sr, const float C1, const float C2,
                float *__restrict__ ssim_map, float *__restrict__ dm_dmu1, float *__restrict__ dm_dsigma1_sq,
                float *__restrict__ dm_dsigma12, double *__restrict__ acc) {
    __shared__ __align__(16) float s_x[SS_TH][SS_TP], s_y[SS_TH][SS_TP];
    __shared__ __align__(16) float s_c[5][SS_TH][SS_X];
    __shared__ float s_red[SS_THREADS / 32];
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int c = blockIdx.z % CH, b = blockIdx.z / CH;
    const int x0 = blockIdx.x * SS_X, y0 = blockIdx.y * SS_Y;
    const float *ip = img + c * si.c + b * si.b, *rp = ref + c * sr.c + b * sr.b;

    // 1) tile + halo, zero padded (get_pix_value, fused_ssim.jl:27-31): one warp per staged row, lane -> columns
    //    lane and lane + 32; column offsets hoisted, all of a thread's (up to 24) loads issued before the stores
    float l1_local = 0.f;
    {
        const int gxa = x0 + tx - SS_HALO, gxb = gxa + 32;
        const bool oka = gxa >= 0 && gxa < W, okb = tx + 32 < SS_TW && gxb < W;
        const int64_t ia = (int64_t)gxa * si.x, ib = (int64_t)gxb * si.x, ra = (int64_t)gxa * sr.x, rb = (int64_t)gxb * sr.x;
        constexpr int NR = (SS_TH + SS_THREADS / 32 - 1) / (SS_THREADS / 32);  // 6 rows per warp
        float Xa[NR], Ya[NR], Xb[NR], Yb[NR];
#pragma unroll
        for (int i = 0; i < NR; i++) {
            const int ly = ty + i * (SS_THREADS / 32);
            const int gy = y0 + ly - SS_HALO;
            const bool row_ok = ly < SS_TH && gy >= 0 && gy < H;
            const float *irow = ip + (int64_t)gy * si.y, *rrow = rp + (int64_t)gy * sr.y;
            Xa[i] = (row_ok && oka) ? __ldg(irow + ia) : 0.f;
            Ya[i] = (row_ok && oka) ? __ldg(rrow + ra) : 0.f;
            Xb[i] = (row_ok && okb) ? __ldg(irow + ib) : 0.f;
            Yb[i] = (row_ok && okb) ? __ldg(rrow + rb) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < NR; i++) {
            const int ly = ty + i * (SS_THREADS / 32);
            if (ly < SS_TH) {
                s_x[ly][tx] = Xa[i];
                s_y[ly][tx] = Ya[i];
                if (tx + 32 < SS_TP) {  // columns 42, 43 are padding
                    s_x[ly][tx + 32] = Xb[i];
                    s_y[ly][tx + 32] = Yb[i];
                }
                // centre pixels of the tile (each exactly once per CTA): the L1 term
                if (LOSS && ly >= SS_HALO && ly < SS_HALO + SS_Y) {
                    if (tx >= SS_HALO && oka && y0 + ly - SS_HALO < H) l1_local += fabsf(Xa[i] - Ya[i]);
                    if (tx + 32 < SS_HALO + SS_X && okb && y0 + ly - SS_HALO < H) l1_local += fabsf(Xb[i] - Yb[i]);
                }
            }
        }
    }
    __syncthreads();

    // 2) horizontal 11-tap pass over the 42 staged rows (fused_ssim.jl:83-160): task = (row, pair of columns)
    for (int task = tid; task < SS_TH * (SS_X / 2); task += SS_THREADS) {
        const int r = task >> 4, cp = (task & 15) * 2;  // outputs at columns cp, cp + 1 read inputs cp .. cp + 11
        float X[12], Y[12], XX[12], YY[12], XY[12];
#pragma unroll
        for (int j = 0; j < 6; j++) {
            const float2 a = *reinterpret_cast<const float2 *>(&s_x[r][cp + 2 * j]);
            const float2 bb = *reinterpret_cast<const float2 *>(&s_y[r][cp + 2 * j]);
            X[2 * j] = a.x; X[2 * j + 1] = a.y;
            Y[2 * j] = bb.x; Y[2 * j + 1] = bb.y;
        }
#pragma unroll
        for (int j = 0; j < 12; j++) {
            XX[j] = X[j] * X[j];
            YY[j] = Y[j] * Y[j];
            XY[j] = X[j] * Y[j];
        }
        float res[5][2];
#pragma unroll
        for (int o = 0; o < 2; o++) {
            float sX = 0.f, sX2 = 0.f, sY = 0.f, sY2 = 0.f, sXY = 0.f;
#pragma unroll
            for (int d = 1; d <= SS_HALO; d++) {
                const float w = c_gauss[SS_HALO - d];
                const int l = o + SS_HALO - d, rr = o + SS_HALO + d;
                sX += (X[l] + X[rr]) * w;
                sX2 += (XX[l] + XX[rr]) * w;
                sY += (Y[l] + Y[rr]) * w;
                sY2 += (YY[l] + YY[rr]) * w;
                sXY += (XY[l] + XY[rr]) * w;
            }
            const float wc = c_gauss[SS_HALO];
            const int m = o + SS_HALO;
            res[0][o] = sX + X[m] * wc;
            res[1][o] = sX2 + XX[m] * wc;
            res[2][o] = sY + Y[m] * wc;
            res[3][o] = sY2 + YY[m] * wc;
            res[4][o] = sXY + XY[m] * wc;
        }
#pragma unroll
        for (int k = 0; k < 5; k++) *reinterpret_cast<float2 *>(&s_c[k][r][cp]) = make_float2(res[k][0], res[k][1]);
    }
    __syncthreads();

    // 3) vertical pass: this thread's four adjacent output rows 4*ty .. 4*ty + 3 read staged rows 4*ty .. 4*ty + 13
    float out[SS_VR][5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        float v[SS_VR + 10];
#pragma unroll
        for (int j = 0; j < SS_VR + 10; j++) v[j] = s_c[k][SS_VR * ty + j][tx];
#pragma unroll
        for (int o = 0; o < SS_VR; o++) {
            float a = 0.f;
#pragma unroll
            for (int d = 1; d <= SS_HALO; d++) a += (v[o + SS_HALO - d] + v[o + SS_HALO + d]) * c_gauss[SS_HALO - d];
            out[o][k] = a + v[o + SS_HALO] * c_gauss[SS_HALO];
        }
    }
    float ssim_local = 0.f;
    const int px = x0 + tx;
#pragma unroll
    for (int o = 0; o < SS_VR; o++) {
        const int py = y0 + SS_VR * ty + o;
        if (px < W && py < H) {
            const float mu1 = out[o][0], mu2 = out[o][2];
            const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2;
            const float sigma1_sq = out[o][1] - mu1_sq, sigma2_sq = out[o][3] - mu2_sq, sigma12 = out[o][4] - mu1 * mu2;
            const float A = mu1_sq + mu2_sq + C1, Bv = sigma1_sq + sigma2_sq + C2;
            const float Cv = 2.f * mu1 * mu2 + C1, Dv = 2.f * sigma12 + C2;
            const float val = (Cv * Dv) / (A * Bv);  // fused_ssim.jl:233
            const int64_t op = (int64_t)px + (int64_t)py * W + ((int64_t)c + (int64_t)b * CH) * ((int64_t)W * H);
            if (ssim_map) ssim_map[op] = val;
            ssim_local += val;
            if (TRAIN) {  // fused_ssim.jl:237-250
                const float AB = A * Bv;
                dm_dmu1[op] = (mu2 * 2.f * Dv) / AB - (mu2 * 2.f * Cv) / AB - (mu1 * 2.f * Cv * Dv) / (A * AB) +
                              (mu1 * 2.f * Cv * Dv) / (AB * Bv);
                dm_dsigma1_sq[op] = (-Cv * Dv) / (AB * Bv);
                dm_dsigma12[op] = (2.f * Cv) / AB;
            }
        }
    }
    if (LOSS) {
        const float a = block_sum(ssim_local, s_red);
        const float l = block_sum(l1_local, s_red);
        if (tid == 0) {
            atomicAdd(acc + 0, (double)a);
            atomicAdd(acc + 1, (double)l);
        }
    }
}

// LOSS: dL/dmap is the constant `chain` (= -lambda/n) and the epilogue adds l1_scale * sign(x - t).
template <bool LOSS>
__global__ void __launch_bounds__(SS_THREADS)
ssim_bwd_kernel(const int W, const int H, const int CH, const float *__restrict__ img, const Strides si,
                const float *__restrict__ ref, const Strides sr, const float *__restrict__ dL_dmap,
                const float chain_const, const float l1_scale, const float *__restrict__ dm_dmu1,
                const float *__restrict__ dm_dsigma1_sq, const float *__restrict__ dm_dsigma12,
                float *__restrict__ dL_dimg, const Strides so) {
    __shared__ __align__(16) float s_d[3][SS_TH][SS_TP];
    __shared__ __align__(16) float s_c[3][SS_TH][SS_X];
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int c = blockIdx.z % CH, b = blockIdx.z / CH;
    const int x0 = blockIdx.x * SS_X, y0 = blockIdx.y * SS_Y;
    const int64_t plane = ((int64_t)c + (int64_t)b * CH) * ((int64_t)W * H);

    // 1) load + fuse the chain multiplication (fused_ssim.jl:286-311): one warp per staged row, loads batched
    {
        const int gxa = x0 + tx - SS_HALO, gxb = gxa + 32;
        const bool oka = gxa >= 0 && gxa < W, okb = tx + 32 < SS_TW && gxb < W;
        constexpr int NR = (SS_TH + SS_THREADS / 32 - 1) / (SS_THREADS / 32);
#pragma unroll
        for (int i0 = 0; i0 < NR; i0 += 3) {  // 3 rows x 2 columns x (3 or 4) loads in flight
            float va[3][3], vb[3][3];
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const int ly = ty + (i0 + i) * (SS_THREADS / 32);
                const int gy = y0 + ly - SS_HALO;
                const bool row_ok = ly < SS_TH && gy >= 0 && gy < H;
                const int64_t p = plane + (int64_t)gy * W;
                float ca = 0.f, cb = 0.f;
                if (LOSS) {
                    ca = cb = chain_const;
                } else {
                    if (row_ok && oka) ca = __ldg(dL_dmap + p + gxa);
                    if (row_ok && okb) cb = __ldg(dL_dmap + p + gxb);
                }
                va[i][0] = (row_ok && oka) ? __ldg(dm_dmu1 + p + gxa) * ca : 0.f;
                va[i][1] = (row_ok && oka) ? __ldg(dm_dsigma1_sq + p + gxa) * ca : 0.f;
                va[i][2] = (row_ok && oka) ? __ldg(dm_dsigma12 + p + gxa) * ca : 0.f;
                vb[i][0] = (row_ok && okb) ? __ldg(dm_dmu1 + p + gxb) * cb : 0.f;
                vb[i][1] = (row_ok && okb) ? __ldg(dm_dsigma1_sq + p + gxb) * cb : 0.f;
                vb[i][2] = (row_ok && okb) ? __ldg(dm_dsigma12 + p + gxb) * cb : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const int ly = ty + (i0 + i) * (SS_THREADS / 32);
                if (ly < SS_TH) {
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        s_d[k][ly][tx] = va[i][k];
                        if (tx + 32 < SS_TP) s_d[k][ly][tx + 32] = vb[i][k];
                    }
                }
            }
        }
    }
    __syncthreads();

    // 2) horizontal pass (fused_ssim.jl:314-345): task = (row, pair of columns)
    for (int task = tid; task < SS_TH * (SS_X / 2); task += SS_THREADS) {
        const int r = task >> 4, cp = (task & 15) * 2;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            float v[12];
#pragma unroll
            for (int j = 0; j < 6; j++) {
                const float2 a = *reinterpret_cast<const float2 *>(&s_d[k][r][cp + 2 * j]);
                v[2 * j] = a.x; v[2 * j + 1] = a.y;
            }
            float res[2];
#pragma unroll
            for (int o = 0; o < 2; o++) {
                float a = 0.f;
#pragma unroll
                for (int d = 1; d <= SS_HALO; d++) a += (v[o + SS_HALO - d] + v[o + SS_HALO + d]) * c_gauss[SS_HALO - d];
                res[o] = a + v[o + SS_HALO] * c_gauss[SS_HALO];
            }
            *reinterpret_cast<float2 *>(&s_c[k][r][cp]) = make_float2(res[0], res[1]);
        }
    }
    __syncthreads();

    // 3) vertical pass, four adjacent rows per thread, and the final combination (fused_ssim.jl:348-381)
    float s[SS_VR][3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float v[SS_VR + 10];
#pragma unroll
        for (int j = 0; j < SS_VR + 10; j++) v[j] = s_c[k][SS_VR * ty + j][tx];
#pragma unroll
        for (int o = 0; o < SS_VR; o++) {
            float a = 0.f;
#pragma unroll
            for (int d = 1; d <= SS_HALO; d++) a += (v[o + SS_HALO - d] + v[o + SS_HALO + d]) * c_gauss[SS_HALO - d];
            s[o][k] = a + v[o + SS_HALO] * c_gauss[SS_HALO];
        }
    }
    const int px = x0 + tx;
    const float *ip = img + c * si.c + b * si.b, *rp = ref + c * sr.c + b * sr.b;
    float *op = dL_dimg + c * so.c + b * so.b;
#pragma unroll
    for (int o = 0; o < SS_VR; o++) {
        const int py = y0 + SS_VR * ty + o;
        if (px < W && py < H) {
            const float p1 = __ldg(ip + px * si.x + py * si.y), p2 = __ldg(rp + px * sr.x + py * sr.y);
            float g = s[o][0] + 2.f * p1 * s[o][1] + p2 * s[o][2];  // fused_ssim.jl:379
            if (LOSS) {
                const float d = p1 - p2;
                g += d > 0.f ? l1_scale : (d < 0.f ? -l1_scale : 0.f);  // pullback of mean|x - t|: sign(x - t)/n
            }
            op[px * so.x + py * so.y] = g;
        }
    }
}

__global__ void loss_finalize_kernel(const double *acc, const double inv_n, const float lambda, float *loss) {
    const double ssim_mean = acc[0] * inv_n, l1 = acc[1] * inv_n;
    loss[0] = (float)((1.0 - (double)lambda) * l1 + (double)lambda * (1.0 - ssim_mean));  // training.jl:690-699
    loss[1] = (float)l1;
    loss[2] = (float)ssim_mean;
}

dim3 ssim_grid(int W, int H, int CH, int B) {
    return dim3((W + SS_X - 1) / SS_X, (H + SS_Y - 1) / SS_Y, CH * B);
}

}  // namespace

int launch_ssim_forward(int W, int H, int CH, int B, const float *img, const float *ref, float C1, float C2, int train,
                        float *ssim_map, float *dm_dmu1, float *dm_dsigma1_sq, float *dm_dsigma12, cudaStream_t s) {
    const Strides p = planar(W, H, CH);
    if (train)
        ssim_fwd_kernel<true, false><<<ssim_grid(W, H, CH, B), SS_THREADS, 0, s>>>(W, H, CH, img, p, ref, p, C1, C2, ssim_map, dm_dmu1, dm_dsigma1_sq, dm_dsigma12, nullptr);
    else
        ssim_fwd_kernel<false, false><<<ssim_grid(W, H, CH, B), SS_THREADS, 0, s>>>(W, H, CH, img, p, ref, p, C1, C2, ssim_map, nullptr, nullptr, nullptr, nullptr);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_ssim_backward(int W, int H, int CH, int B, const float *img, const float *ref, const float *dL_dmap,
                         const float *dm_dmu1, const float *dm_dsigma1_sq, const float *dm_dsigma12, float *dL_dimg,
                         cudaStream_t s) {
    const Strides p = planar(W, H, CH);
    ssim_bwd_kernel<false><<<ssim_grid(W, H, CH, B), SS_THREADS, 0, s>>>(W, H, CH, img, p, ref, p, dL_dmap, 0.f, 0.f, dm_dmu1, dm_dsigma1_sq, dm_dsigma12, dL_dimg, p);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// image / vpixels: the rasterizer's (C,W,H) interleaved layout, rgb in channels 0..2; target: (W,H,3) planar.
// scratch: 3*W*H*3 floats (derivative maps) ; acc: 2 doubles ; loss: 3 floats {total, l1, mean ssim}.
int launch_photometric_loss(int W, int H, int C, const float *image, const float *target, float lambda, float C1,
                            float C2, float *scratch, double *acc, float *vpixels, float *loss, cudaStream_t s) {
    const Strides hwc{(int64_t)C, (int64_t)C * W, 1, 0};
    const Strides pl = planar(W, H, 3);
    const size_t plane3 = (size_t)W * H * 3;
    float *d0 = scratch, *d1 = scratch + plane3, *d2 = scratch + 2 * plane3;
    const double n = (double)plane3;
    if (cudaMemsetAsync(acc, 0, 2 * sizeof(double), s) != cudaSuccess) return -1;
    if (C > 3 && cudaMemsetAsync(vpixels, 0, (size_t)W * H * C * sizeof(float), s) != cudaSuccess) return -1;
    ssim_fwd_kernel<true, true><<<ssim_grid(W, H, 3, 1), SS_THREADS, 0, s>>>(W, H, 3, image, hwc, target, pl, C1, C2, nullptr, d0, d1, d2, acc);
    ssim_bwd_kernel<true><<<ssim_grid(W, H, 3, 1), SS_THREADS, 0, s>>>(W, H, 3, image, hwc, target, pl, nullptr, (float)(-(double)lambda / n), (float)((1.0 - (double)lambda) / n), d0, d1, d2, vpixels, hwc);
    loss_finalize_kernel<<<1, 1, 0, s>>>(acc, 1.0 / n, lambda, loss);
    count_launch(3);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
