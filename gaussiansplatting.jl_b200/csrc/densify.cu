// densify.cu — the device kernels of densify_and_prune! (src/densification.jl) for sm_100a.
//
// The reference composes clone / split / prune out of dozens of allocating broadcasts, `findall`-style boolean
// indexing and `cat`s, once per parameter array and once per Adam moment.  Here the same steps are four kinds of
// launches over plain device arrays:
//   densify_masks_kernel   ∇ = accum / denom (NaN -> 0), clone / split masks            densification.jl:8-11,35-38,78-81
//   prune_mask_kernel      opacity / screen-size / world-size validity mask             densification.jl:19-25
//   mask_offsets (3 tiny kernels)  exclusive prefix of a byte mask + count — the compaction index shared by all arrays
//   gather_rows_kernel     dst[:, j + c*count] = src[:, sel_j]  — `x[:, mask]`, `x[:, :, mask]`, `repeat(x[:, mask], 1, r)`
//                          for rows of any byte size (parameters, Adam moments, statistics, ids)
//   split_children_kernel  `_add_split_noise!` + the children's log-scales             densification.jl:83-95,123-136
// Byte movement is exact; exp / log / sigmoid are the same libdevice functions the fused-activation path uses.
// The normal deviates of the split are an INPUT (the reference draws them from the device RNG inside the kernel,
// which no other implementation can reproduce): the caller supplies N(0,1) samples, and parity tests feed the same
// samples to both sides.
#include "common.cuh"

namespace {

__device__ __forceinline__ float max_exp_scale(const float *__restrict__ scales, const int64_t i, const int isotropic) {
    if (isotropic) return expf(scales[i]);
    return fmaxf(fmaxf(expf(scales[3 * i]), expf(scales[3 * i + 1])), expf(scales[3 * i + 2]));
}

// n_grad <= n: Gaussians appended since the statistics were taken (clones) have gradient 0 (padded_grad, :74-75)
__global__ void densify_masks_kernel(const int64_t n, const int64_t n_grad, const float *__restrict__ accum,
                                     const float *__restrict__ denom, const float *__restrict__ scales, const int isotropic,
                                     const float grad_threshold, const float gamma, uint8_t *__restrict__ clone_mask,
                                     uint8_t *__restrict__ split_mask) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float g = 0.0f;
    if (i < n_grad) {
        g = accum[i] / denom[i];  // ./ (densification.jl:8)
        if (isnan(g)) g = 0.0f;   // :9-10
    }
    const float s = max_exp_scale(scales, i, isotropic);
    if (clone_mask) clone_mask[i] = (g > grad_threshold && s < gamma) ? 1 : 0;    // :36-38
    if (split_mask) split_mask[i] = (g >= grad_threshold && s > gamma) ? 1 : 0;   // :79-81
}

__global__ void prune_mask_kernel(const int64_t n, const float *__restrict__ opacities, const float *__restrict__ scales,
                                  const int isotropic, const int32_t *__restrict__ max_radii, const float min_opacity,
                                  const int32_t max_screen_size, const float gamma, uint8_t *__restrict__ valid) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool v = act_sigmoid(opacities[i]) > min_opacity;  // :19
    if (max_screen_size > 0)                           // :20-25
        v = v && (max_radii[i] < max_screen_size) && (max_exp_scale(scales, i, isotropic) < gamma);
    valid[i] = v ? 1 : 0;
}

// ---- exclusive prefix of a byte mask: block counts -> scan of the block counts -> offsets --------------------------
constexpr int MO_THREADS = 256, MO_IPT = 16, MO_TILE = MO_THREADS * MO_IPT;

__global__ void __launch_bounds__(MO_THREADS)
mask_block_count_kernel(const int64_t n, const uint8_t *__restrict__ mask, int32_t *__restrict__ block_sums) {
    __shared__ int s_w[MO_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * MO_TILE + (int64_t)threadIdx.x * MO_IPT;
    int c = 0;
#pragma unroll
    for (int k = 0; k < MO_IPT; k++) c += (base + k < n && mask[base + k]) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < MO_THREADS / 32; w++) t += s_w[w];
        block_sums[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024)
mask_block_scan_kernel(const int nblocks, int32_t *__restrict__ block_sums, int64_t *__restrict__ count_out) {
    // one CTA, sequential over chunks of 1024 blocks (4 M Gaussians per chunk): exclusive scan in place
    __shared__ int s_w[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nblocks; b0 += 1024) {
        const int b = b0 + threadIdx.x;
        const int v = b < nblocks ? block_sums[b] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += y;
        }
        if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) woff += s_w[w];
        const int carry = s_carry;
        if (b < nblocks) block_sums[b] = carry + woff + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + woff + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count_out = s_carry;
}

__global__ void __launch_bounds__(MO_THREADS)
mask_offsets_kernel(const int64_t n, const uint8_t *__restrict__ mask, const int32_t *__restrict__ block_excl,
                    int32_t *__restrict__ offsets) {
    __shared__ int s_w[MO_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * MO_TILE + (int64_t)threadIdx.x * MO_IPT;
    int m[MO_IPT], c = 0;
#pragma unroll
    for (int k = 0; k < MO_IPT; k++) {
        m[k] = (base + k < n && mask[base + k]) ? 1 : 0;
        c += m[k];
    }
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += y;
    }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) woff += s_w[w];
    int run = block_excl[blockIdx.x] + woff + incl - c;
#pragma unroll
    for (int k = 0; k < MO_IPT; k++) {
        if (base + k < n) offsets[base + k] = run;
        run += m[k];
    }
}

// dst row (offsets[i] + c*count) <- src row i for every selected i and copy c < repeat.  Rows are `row_words` 32-bit
// words (all of the reference's per-Gaussian arrays are 4-byte typed); one thread per (selected row, word).
__global__ void gather_rows_kernel(const int64_t n, const int row_words, const uint32_t *__restrict__ src,
                                   const uint8_t *__restrict__ mask, const int32_t *__restrict__ offsets,
                                   uint32_t *__restrict__ dst, const int repeat, const int64_t count) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = t / row_words;
    const int w = (int)(t - i * row_words);
    if (i >= n || !mask[i]) return;
    const uint32_t v = src[i * row_words + w];
    const int64_t j = offsets[i];
    for (int c = 0; c < repeat; c++) dst[(j + (int64_t)c * count) * row_words + w] = v;
}

// children of a split, in place on the gathered + repeated arrays (densification.jl:83-101,123-136):
//   stds = exp(scales) ; points += R(q) * (stds .* xi) ; scales = log(stds / (0.8 * n_split))
__global__ void split_children_kernel(const int64_t m, float *__restrict__ points, float *__restrict__ scales,
                                      const int isotropic, const float *__restrict__ rotations,
                                      const float *__restrict__ noise, const float shrink) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    float sd[3];
    if (isotropic) {
        sd[0] = sd[1] = sd[2] = expf(scales[i]);
    } else {
        sd[0] = expf(scales[3 * i]); sd[1] = expf(scales[3 * i + 1]); sd[2] = expf(scales[3 * i + 2]);
    }
    const float xi[3] = {sd[0] * noise[3 * i], sd[1] * noise[3 * i + 1], sd[2] * noise[3 * i + 2]};
    // unnorm_quat2rot (render.jl:322-333), column-major R
    const float4 q4 = *reinterpret_cast<const float4 *>(rotations + 4 * i);
    const float qi = 1.0f / sqrtf(((q4.x * q4.x + q4.y * q4.y) + q4.z * q4.z) + q4.w * q4.w);
    const float w = qi * q4.x, x = qi * q4.y, y = qi * q4.z, z = qi * q4.w;
    const float R00 = 1.0f - 2.0f * (y * y + z * z), R10 = 2.0f * (x * y + w * z), R20 = 2.0f * (x * z - w * y);
    const float R01 = 2.0f * (x * y - w * z), R11 = 1.0f - 2.0f * (x * x + z * z), R21 = 2.0f * (y * z + w * x);
    const float R02 = 2.0f * (x * z + w * y), R12 = 2.0f * (y * z - w * x), R22 = 1.0f - 2.0f * (x * x + y * y);
    points[3 * i] += (R00 * xi[0] + R01 * xi[1]) + R02 * xi[2];
    points[3 * i + 1] += (R10 * xi[0] + R11 * xi[1]) + R12 * xi[2];
    points[3 * i + 2] += (R20 * xi[0] + R21 * xi[1]) + R22 * xi[2];
    if (isotropic) {
        scales[i] = logf(sd[0] / shrink);  // log.(stds ./ (0.8f0 * n_split)), :91
    } else {
        scales[3 * i] = logf(sd[0] / shrink);
        scales[3 * i + 1] = logf(sd[1] / shrink);
        scales[3 * i + 2] = logf(sd[2] / shrink);
    }
}

unsigned blocks_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace

int launch_densify_masks(int64_t n, int64_t n_grad, const float *accum, const float *denom, const float *scales,
                         int isotropic, float grad_threshold, float gamma, uint8_t *clone_mask, uint8_t *split_mask,
                         cudaStream_t s) {
    if (n <= 0) return 0;
    densify_masks_kernel<<<blocks_for(n, 256), 256, 0, s>>>(n, n_grad, accum, denom, scales, isotropic, grad_threshold, gamma,
                                                          clone_mask, split_mask);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_prune_mask(int64_t n, const float *opacities, const float *scales, int isotropic, const int32_t *max_radii,
                      float min_opacity, int32_t max_screen_size, float gamma, uint8_t *valid, cudaStream_t s) {
    if (n <= 0) return 0;
    prune_mask_kernel<<<blocks_for(n, 256), 256, 0, s>>>(n, opacities, scales, isotropic, max_radii, min_opacity,
                                                       max_screen_size, gamma, valid);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

size_t mask_offsets_scratch_words(int64_t n) { return (size_t)((n + MO_TILE - 1) / MO_TILE) + 1; }

int launch_mask_offsets(int64_t n, const uint8_t *mask, int32_t *offsets, int64_t *count_dev, int32_t *scratch,
                        cudaStream_t s) {
    if (n <= 0) return cudaMemsetAsync(count_dev, 0, sizeof(int64_t), s) == cudaSuccess ? 0 : -1;
    const int nb = (int)((n + MO_TILE - 1) / MO_TILE);
    mask_block_count_kernel<<<nb, MO_THREADS, 0, s>>>(n, mask, scratch);
    mask_block_scan_kernel<<<1, 1024, 0, s>>>(nb, scratch, count_dev);
    mask_offsets_kernel<<<nb, MO_THREADS, 0, s>>>(n, mask, scratch, offsets);
    count_launch(3);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_gather_rows(int64_t n, int row_words, const void *src, const uint8_t *mask, const int32_t *offsets, void *dst,
                       int repeat, int64_t count, cudaStream_t s) {
    if (n <= 0 || row_words <= 0 || count <= 0) return 0;
    gather_rows_kernel<<<blocks_for(n * row_words, 256), 256, 0, s>>>(n, row_words, static_cast<const uint32_t *>(src), mask,
                                                                    offsets, static_cast<uint32_t *>(dst), repeat, count);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_split_children(int64_t m, float *points, float *scales, int isotropic, const float *rotations,
                          const float *noise, int n_split, cudaStream_t s) {
    if (m <= 0) return 0;
    split_children_kernel<<<blocks_for(m, 256), 256, 0, s>>>(m, points, scales, isotropic, rotations, noise,
                                                           0.8f * (float)n_split);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
