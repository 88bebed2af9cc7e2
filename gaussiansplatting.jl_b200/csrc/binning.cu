// binning.cu — scan, tile-instance duplication, onesweep radix sort and per-tile ranges for sm_100a.
//
// Replaces, on the reference's path (src/rasterization/rasterizer.jl:333-378):
//   cumsum!                      -> scan_kernel        single-pass decoupled look-back inclusive scan
//   duplicate_with_keys!         -> duplicate_kernel   (utils.jl:85-120) warp-cooperative for large rects
//   sortperm! + 2 x _permute!    -> hist_kernel + onesweep_kernel x passes   (rasterizer.jl:357-372)
//   identify_tile_range!         -> tile_ranges_kernel (utils.jl:56-78)
//
// Sort: least-significant-digit radix sort, 8-bit digits, one histogram pass over the keys for all digit
// positions, then ONE kernel per digit that ranks, looks back (chained scan over 4096-key tiles, per digit)
// and scatters — the "onesweep" scheme.  Only the key bits that can differ are sorted: ceil(log2 T) tile bits
// and the bits of (bits(depth) - bits(near)) (27 for near=0.2, far=1000), i.e. 40 bits -> 5 passes at
// 1920x1088 instead of 8 for the raw 64-bit key.  LSD radix is stable, so equal keys
// keep emission order (ascending Gaussian id), which the reference's sortperm! also guarantees.
// The published buffers hold the canonical (tile << 32 | bits(depth)) keys and 1-based ids.
//
//
// Depth pre-sort (default): every instance of a Gaussian carries the same depth, so sorting the M instances on the
// depth digits re-sorts M keys on bits that only the N Gaussians distinguish.  Instead the Gaussians are sorted by
// depth first (presort_keys_kernel + the same onesweep passes over N pairs), instances are emitted in that order
// (scan and duplicate read through the permutation), and the instance sort runs over the tile digits only:
// 2 passes over M + 4 over N instead of 5 over M at 1920x1088.  LSD radix sort is stable, so the result is the same
// permutation bit for bit: (tile, depth, emission order = Gaussian id).
//
// All integer work: bit-exact by construction; compiled with default flags.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace {

// ----------------------------------------------------------------------------------------------------------
// inclusive scan of tiles_touched (Int32) — decoupled look-back, 64-bit status words {flag:2 | value:62}
// ----------------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_IPT = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_IPT;
constexpr unsigned long long FLAG_AGG = 1ull << 62, FLAG_INC = 2ull << 62, VAL_MASK = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
    return *reinterpret_cast<const volatile unsigned long long *>(p);
}
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) {
    return *reinterpret_cast<const volatile uint32_t *>(p);
}

// perm != nullptr: element j of the scan is in[perm[j]] (tiles touched in depth order, see the pre-sort above)
__global__ void __launch_bounds__(SCAN_THREADS)
scan_kernel(const int64_t n, const int32_t *__restrict__ in, const uint32_t *__restrict__ perm, int32_t *__restrict__ out,
            unsigned long long *state /* [0] = tile counter, [1..] = status */, int64_t *total) {
    __shared__ long long s_warp[SCAN_THREADS / 32];
    __shared__ long long s_prefix;
    __shared__ unsigned s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = (unsigned)atomicAdd(&state[0], 1ull);
    __syncthreads();
    const int64_t tile = s_tile;
    unsigned long long *status = state + 1;
    const int64_t base = tile * SCAN_TILE + (int64_t)tid * SCAN_IPT;

    int32_t v[SCAN_IPT];
    const int32_t *src = perm ? reinterpret_cast<const int32_t *>(perm) : in;
    if (base + SCAN_IPT <= n) {
        const int4 a = *reinterpret_cast<const int4 *>(src + base);
        const int4 b = *reinterpret_cast<const int4 *>(src + base + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_IPT; k++) v[k] = (base + k < n) ? src[base + k] : 0;
    }
    if (perm) {  // eight independent gathers in flight
#pragma unroll
        for (int k = 0; k < SCAN_IPT; k++) v[k] = (base + k < n) ? __ldg(in + v[k]) : 0;
    }
    long long tsum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; k++) tsum += v[k];
    // block exclusive scan of thread sums
    long long incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    long long woff = 0, agg = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; w++) {
        const long long x = s_warp[w];
        if (w < warp) woff += x;
        agg += x;
    }
    const long long texcl = woff + incl - tsum;

    // look-back by warp 0
    if (warp == 0) {
        long long prefix = 0;
        if (tile == 0) {
            if (lane == 0) atomicExch(&status[0], FLAG_INC | (unsigned long long)agg);
        } else {
            if (lane == 0) atomicExch(&status[tile], FLAG_AGG | (unsigned long long)agg);
            int64_t look = tile - 1;
            while (true) {
                const int64_t t = look - lane;
                unsigned long long s = FLAG_INC;  // lanes before tile 0 act as an inclusive zero
                if (t >= 0) {
                    do { s = ld_volatile_u64(&status[t]); } while ((s >> 62) == 0);
                }
                const unsigned inc_mask = __ballot_sync(0xffffffffu, (s >> 62) == 2);
                const int first_inc = inc_mask ? (__ffs(inc_mask) - 1) : 32;
                long long contrib = (lane <= first_inc) ? (long long)(s & VAL_MASK) : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
                prefix += contrib;
                if (inc_mask) break;
                look -= 32;
            }
            if (lane == 0) atomicExch(&status[tile], FLAG_INC | (unsigned long long)(prefix + agg));
        }
        if (lane == 0) s_prefix = prefix;
    }
    __syncthreads();
    long long run = s_prefix + texcl;
    int32_t o[SCAN_IPT];
#pragma unroll
    for (int k = 0; k < SCAN_IPT; k++) { run += v[k]; o[k] = (int32_t)run; }
    if (base + SCAN_IPT <= n) {
        *reinterpret_cast<int4 *>(out + base) = make_int4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<int4 *>(out + base + 4) = make_int4(o[4], o[5], o[6], o[7]);
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_IPT; k++)
            if (base + k < n) out[base + k] = o[k];
    }
    if (tile == (n - 1) / SCAN_TILE && tid == SCAN_THREADS - 1) {
        *total = s_prefix + agg;  // `total` is host-mapped pinned memory (zero-copy)
        __threadfence_system();
    }
}

// ----------------------------------------------------------------------------------------------------------
// duplicate_with_keys! — utils.jl:85-120
// ----------------------------------------------------------------------------------------------------------
constexpr int DUP_THREADS = 256;
constexpr int DUP_SERIAL_MAX = 12;
constexpr int SORT_MAX_PASSES = 8;

// Digit histograms of the radix sort are accumulated while the keys are generated (no extra pass over the M
// keys): each thread emits ITS Gaussian's instances, so concurrent shared-memory atomics hit unrelated tiles
// (low conflict); digits that lie entirely inside the depth bits are added once per Gaussian with weight cnt.
__device__ __forceinline__ void hist_add_instance(uint32_t *sh, uint64_t ck, int first_pass, int passes, int bit_lo) {
    for (int p = first_pass; p < passes; p++) atomicAdd(&sh[p * 256 + (uint32_t)((ck >> (bit_lo + 8 * p)) & 255u)], 1u);
}

__global__ void __launch_bounds__(DUP_THREADS)
duplicate_kernel(const int64_t n, const int32_t grid_x, const int32_t grid_y, const int32_t *__restrict__ radii,
                 const float2 *__restrict__ means2d, const float *__restrict__ depths,
                 const int32_t *__restrict__ offsets, const uint32_t *__restrict__ perm, uint64_t *__restrict__ keys,
                 uint32_t *__restrict__ vals, const int depth_bits, const uint32_t depth_base, const int bit_lo,
                 const int passes, uint32_t *__restrict__ ghist) {
    __shared__ uint32_t sh[SORT_MAX_PASSES * 256];
    const bool do_hist = ghist != nullptr;
    if (do_hist) {
        for (int k = threadIdx.x; k < passes * 256; k += DUP_THREADS) sh[k] = 0;
        __syncthreads();
    }
    // passes whose 8-bit digit lies entirely in the depth bits (none when the instance sort starts at the tile bits)
    const int depth_only = depth_bits > bit_lo ? (depth_bits - bit_lo) / 8 : 0;
    const int64_t j = (int64_t)blockIdx.x * DUP_THREADS + threadIdx.x;  // emission slot
    const int lane = threadIdx.x & 31;
    int32_t x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    uint32_t dbits = 0;
    int64_t off = 0, i = j;
    int cnt = 0;
    if (j < n) {
        if (perm) i = (int64_t)perm[j];  // emission order = depth order
        const int32_t r = radii[i];
        if (r > 0) {
            const float2 m = means2d[i];
            get_rect(m.x, m.y, r, grid_x, grid_y, x0, y0, x1, y1);
            dbits = __float_as_uint(depths[i]);
            off = (j == 0) ? 0 : (int64_t)offsets[j - 1];
            cnt = (x1 - x0) * (y1 - y0);
        }
    }
    const uint32_t id = (uint32_t)(i + 1);  // 1-based, as the reference emits
    const uint32_t dlow = dbits - depth_base;
    const int first_pass = depth_only < passes ? depth_only : passes;
    if (do_hist && cnt > 0)
        for (int p = 0; p < first_pass; p++) atomicAdd(&sh[p * 256 + ((dlow >> (bit_lo + 8 * p)) & 255u)], (uint32_t)cnt);
    if (cnt > 0 && cnt <= DUP_SERIAL_MAX) {
        int64_t o = off;
        for (int32_t y = y0; y < y1; y++)
            for (int32_t x = x0; x < x1; x++) {
                const uint64_t tile = (uint64_t)y * (uint64_t)grid_x + (uint64_t)x;
                keys[o] = (tile << 32) | dbits;
                vals[o] = id;
                if (do_hist) hist_add_instance(sh, (tile << depth_bits) | dlow, first_pass, passes, bit_lo);
                o++;
            }
    }
    // large rects: the whole warp emits one Gaussian's tiles (the reference's serial loop is the load-imbalance
    // hot spot for big splats)
    unsigned big = __ballot_sync(0xffffffffu, cnt > DUP_SERIAL_MAX);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const int32_t bx0 = __shfl_sync(0xffffffffu, x0, src), by0 = __shfl_sync(0xffffffffu, y0, src);
        const int32_t bx1 = __shfl_sync(0xffffffffu, x1, src);
        const int bcnt = __shfl_sync(0xffffffffu, cnt, src);
        const uint32_t bd = __shfl_sync(0xffffffffu, dbits, src), bid = __shfl_sync(0xffffffffu, id, src);
        const int64_t boff = __shfl_sync(0xffffffffu, off, src);
        const int w = bx1 - bx0;
        const uint32_t bdlow = bd - depth_base;
        for (int t0 = 0; t0 < bcnt; t0 += 32) {  // warp-uniform trip count: match_any below needs convergence
            const int t = t0 + lane;
            const bool valid = t < bcnt;
            uint64_t tile = 0;
            if (valid) {
                const int ty = by0 + t / w, tx = bx0 + t % w;
                tile = (uint64_t)ty * (uint64_t)grid_x + (uint64_t)tx;
                keys[boff + t] = (tile << 32) | bd;
                vals[boff + t] = bid;
            }
            if (do_hist) {  // neighbouring tiles share their high digits: aggregate within the warp
                const uint64_t ck = (tile << depth_bits) | bdlow;
                for (int p = first_pass; p < passes; p++) {
                    const uint32_t d = valid ? (uint32_t)((ck >> (bit_lo + 8 * p)) & 255u) : 0xffffffffu;
                    const unsigned peers = __match_any_sync(0xffffffffu, d);
                    if (valid && lane == __ffs(peers) - 1) atomicAdd(&sh[p * 256 + d], (uint32_t)__popc(peers));
                }
            }
        }
    }
    if (do_hist) {
        __syncthreads();
        for (int k = threadIdx.x; k < passes * 256; k += DUP_THREADS)
            if (sh[k]) atomicAdd(&ghist[k], sh[k]);
    }
}

// ----------------------------------------------------------------------------------------------------------
// depth pre-sort keys: canonical (flag << 32 | bits(depth)) with flag = 1 for culled Gaussians (they sort behind every
// visible one and emit nothing); value = the 0-based Gaussian index
// ----------------------------------------------------------------------------------------------------------
// The digit histograms of the Gaussians' sort are accumulated here as well: depth digits are close to uniformly
// random, so plain shared-memory atomics see few conflicts (MATCH.ANY aggregation, which pays off on clustered tile
// digits, is at its slowest on such data).
__global__ void __launch_bounds__(256)
presort_keys_kernel(const int64_t n, const int32_t *__restrict__ radii, const float *__restrict__ depths,
                    const int depth_bits, const uint32_t depth_base, const int passes, uint64_t *__restrict__ keys,
                    uint32_t *__restrict__ vals, uint32_t *__restrict__ ghist /* [passes][256] */) {
    __shared__ uint32_t sh[SORT_MAX_PASSES * 256];
    for (int k = threadIdx.x; k < passes * 256; k += blockDim.x) sh[k] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const bool vis = radii[i] > 0;
        const uint32_t dbits = vis ? __float_as_uint(depths[i]) : depth_base;
        keys[i] = ((uint64_t)(vis ? 0u : 1u) << 32) | dbits;
        vals[i] = (uint32_t)i;
        const uint64_t ck = ((uint64_t)(vis ? 0u : 1u) << depth_bits) | (uint64_t)(dbits - depth_base);
        for (int p = 0; p < passes; p++) atomicAdd(&sh[p * 256 + (uint32_t)((ck >> (8 * p)) & 255u)], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < passes * 256; k += blockDim.x)
        if (sh[k]) atomicAdd(&ghist[k], sh[k]);
}

// ----------------------------------------------------------------------------------------------------------
// onesweep radix sort on the compact key  ck = (tile << depth_bits) | (bits(depth) - depth_base)
// ----------------------------------------------------------------------------------------------------------
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
// keys per thread (template parameter of the pass kernel): 8 -> 2048-key tiles, ~50 registers, 34 KB smem
// (5 CTAs/SM); 16 -> 4096-key tiles, 80 registers, 59 KB (3 CTAs/SM).  The pass is latency-bound, so occupancy wins.
constexpr uint32_t ST_AGG = 1u << 30, ST_INC = 2u << 30, ST_MASK = (1u << 30) - 1;
#ifndef GSR_SORT_LOOKBACK
#define GSR_SORT_LOOKBACK 4
#endif
constexpr int LOOKBACK = GSR_SORT_LOOKBACK;  // status words of that many preceding tiles fetched per round trip

__device__ __forceinline__ uint64_t compact_key(uint64_t key, int depth_bits, uint32_t depth_base) {
    const uint32_t lo = (uint32_t)key - depth_base;
    return ((key >> 32) << depth_bits) | (uint64_t)lo;
}

__global__ void __launch_bounds__(256)
hist_kernel(const uint64_t *__restrict__ keys, const int64_t m, const int depth_bits, const uint32_t depth_base,
            const int bit_lo, const int passes, uint32_t *__restrict__ ghist /* [passes][256] */) {
    __shared__ uint32_t sh[SORT_MAX_PASSES * 256];
    const int lane = threadIdx.x & 31;
    for (int k = threadIdx.x; k < passes * 256; k += blockDim.x) sh[k] = 0;
    __syncthreads();
    // each warp takes 128 consecutive keys per iteration (4 per lane, 128-bit loads): 4 independent chains of
    // match/atomic per lane hide the match latency; warps stay converged for match_any
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = warp_global * 128; base < m; base += n_warps * 128) {
        uint64_t k4[4];
        const int64_t i0 = base + lane * 4;
        if (i0 + 3 < m) {
            const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(keys + i0);
            const ulonglong2 b = *reinterpret_cast<const ulonglong2 *>(keys + i0 + 2);
            k4[0] = a.x; k4[1] = a.y; k4[2] = b.x; k4[3] = b.y;
        } else {
#pragma unroll
            for (int u = 0; u < 4; u++) k4[u] = (i0 + u < m) ? keys[i0 + u] : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const bool valid = i0 + u < m;
            const uint64_t ck = compact_key(k4[u], depth_bits, depth_base);
            for (int p = 0; p < passes; p++) {
                const uint32_t d = valid ? (uint32_t)((ck >> (bit_lo + 8 * p)) & 255u) : 0xffffffffu;
                const unsigned peers = __match_any_sync(0xffffffffu, d);  // warp-aggregated: tile digits cluster
                if (valid && lane == __ffs(peers) - 1) atomicAdd(&sh[p * 256 + d], (uint32_t)__popc(peers));
            }
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < passes * 256; k += blockDim.x)
        if (sh[k]) atomicAdd(&ghist[k], sh[k]);
}

// digit of a key at `shift`: 64-bit keys are the canonical (tile << 32 | depth bits), sorted through their compact form;
// 32-bit keys are bare tile ids (the instance sort after the depth pre-sort), sorted as they are
template <typename KeyT>
__device__ __forceinline__ uint32_t digit_of(KeyT key, int shift, int depth_bits, uint32_t depth_base);
template <>
__device__ __forceinline__ uint32_t digit_of<uint64_t>(uint64_t key, int shift, int depth_bits, uint32_t depth_base) {
    return (uint32_t)((compact_key(key, depth_bits, depth_base) >> shift) & 255u);
}
template <>
__device__ __forceinline__ uint32_t digit_of<uint32_t>(uint32_t key, int shift, int, uint32_t) {
    return (key >> shift) & 255u;
}

template <typename KeyT, int SORT_IPT>
struct SortSmem {
    KeyT keys[SORT_THREADS * SORT_IPT];
    uint32_t vals[SORT_THREADS * SORT_IPT];
    uint32_t whist[SORT_WARPS][256];
    uint32_t excl[256];
    uint32_t gbase[256];
    uint32_t wsum[SORT_WARPS];
    uint32_t wsum2[SORT_WARPS];
    uint32_t tile;
};

template <typename KeyT, int SORT_IPT, bool BALLOT>
__global__ void __launch_bounds__(SORT_THREADS)
onesweep_kernel(const KeyT *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                KeyT *__restrict__ keys_out, uint32_t *__restrict__ vals_out, const int64_t m, const int shift,
                const int depth_bits, const uint32_t depth_base, const uint32_t *__restrict__ ghist /* [256] */,
                uint32_t *status /* [tiles][256] */, uint32_t *tile_counter, const int ballot_bits) {
    constexpr int SORT_TILE = SORT_THREADS * SORT_IPT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SortSmem<KeyT, SORT_IPT> &S = *reinterpret_cast<SortSmem<KeyT, SORT_IPT> *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) S.tile = atomicAdd(tile_counter, 1u);
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++) S.whist[w][tid] = 0;
    __syncthreads();
    const int64_t tile = S.tile;
    const int64_t tile_base = tile * SORT_TILE;
    const int cnt = (int)((m - tile_base) < SORT_TILE ? (m - tile_base) : SORT_TILE);

    // ---- load (warp-striped inside each warp's contiguous 512-key chunk) and rank by digit ----------------
    KeyT key[SORT_IPT];
    uint32_t val[SORT_IPT];
    uint16_t rnk[SORT_IPT];
    uint32_t *whist = S.whist[warp];
    const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < SORT_IPT; j++) {
        const int local = warp * (32 * SORT_IPT) + j * 32 + lane;
        const bool valid = local < cnt;
        key[j] = valid ? keys_in[tile_base + local] : (KeyT)~(KeyT)0;
        val[j] = valid ? vals_in[tile_base + local] : 0u;
    }
#pragma unroll
    for (int j = 0; j < SORT_IPT; j++) {
        const int local = warp * (32 * SORT_IPT) + j * 32 + lane;
        const bool valid = local < cnt;
        const uint32_t d = valid ? digit_of<KeyT>(key[j], shift, depth_bits, depth_base) : 0xffffffffu;
        // lanes holding the same digit.  MATCH.ANY's latency is data dependent: on digits made of tile bits
        // (runs of neighbouring tiles from one Gaussian) it is ~3x that on depth bits (ncu: 32 % of the pass's
        // stall samples sit on its consumer), so those passes build the mask from one ballot per digit bit instead
        unsigned peers;
        if (!BALLOT) {
            peers = __match_any_sync(0xffffffffu, d);
        } else {
            const unsigned vmask = __ballot_sync(0xffffffffu, valid);
            peers = valid ? vmask : ~vmask;
            for (int b = 0; b < ballot_bits; b++) {
                const bool bit = (d >> b) & 1u;
                const unsigned bal = __ballot_sync(0xffffffffu, bit);
                peers &= bit ? bal : ~bal;
            }
        }
        const int leader = __ffs(peers) - 1;
        uint32_t pre = 0;
        if (valid && lane == leader) {
            pre = whist[d];
            whist[d] = pre + (uint32_t)__popc(peers);
        }
        pre = __shfl_sync(0xffffffffu, pre, leader);
        rnk[j] = (uint16_t)(pre + (uint32_t)__popc(peers & lt_mask));
        __syncwarp();
    }
    __syncthreads();

    // ---- per-digit: exclusive over warps, CTA count, chained look-back over tiles --------------------------
    uint32_t count = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++) {
        const uint32_t c = S.whist[w][tid];
        S.whist[w][tid] = count;
        count += c;
    }
    // publish this tile's per-digit count right away (aggregate), so successors can start summing
    uint32_t *my_status = status + tile * 256 + tid;
    atomicExch(my_status, (tile == 0 ? ST_INC : ST_AGG) | count);
    // two CTA-wide exclusive scans over the 256 digits: this tile's counts and the global histogram
    const uint32_t gh = ghist[tid];
    uint32_t inc1 = count, inc2 = gh;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y1 = __shfl_up_sync(0xffffffffu, inc1, o), y2 = __shfl_up_sync(0xffffffffu, inc2, o);
        if (lane >= o) { inc1 += y1; inc2 += y2; }
    }
    if (lane == 31) { S.wsum[warp] = inc1; S.wsum2[warp] = inc2; }
    __syncthreads();
    uint32_t off1 = 0, off2 = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++)
        if (w < warp) { off1 += S.wsum[w]; off2 += S.wsum2[w]; }
    const uint32_t excl = off1 + inc1 - count;
    const uint32_t gexcl = off2 + inc2 - gh;
    S.excl[tid] = excl;
    __syncthreads();

    // ---- scatter into shared memory in digit order (needs only CTA-local offsets) --------------------------
#pragma unroll
    for (int j = 0; j < SORT_IPT; j++) {
        const int local = warp * (32 * SORT_IPT) + j * 32 + lane;
        if (local < cnt) {
            const uint32_t d = digit_of<KeyT>(key[j], shift, depth_bits, depth_base);
            const uint32_t pos = S.excl[d] + whist[d] + rnk[j];
            S.keys[pos] = key[j];
            S.vals[pos] = val[j];
        }
    }

    // ---- chained look-back over the preceding tiles, per digit; done AFTER the local scatter so that the
    //      predecessors have had time to publish (LOOKBACK predecessors per round trip, independent loads in flight)
    uint32_t tiles_prefix = 0;
    if (tile > 0) {
        bool found = false;
        for (int64_t t = tile - 1; t >= 0 && !found; t -= LOOKBACK) {
            uint32_t sv[LOOKBACK];
#pragma unroll
            for (int j = 0; j < LOOKBACK; j++) sv[j] = (t - j >= 0) ? ld_volatile_u32(status + (t - j) * 256 + tid) : ST_INC;
#pragma unroll
            for (int j = 0; j < LOOKBACK; j++) {
                if (found) break;
                uint32_t sj = sv[j];
                while ((sj >> 30) == 0) sj = ld_volatile_u32(status + (t - j) * 256 + tid);
                tiles_prefix += sj & ST_MASK;
                found = (sj >> 30) == 2;
            }
        }
        atomicExch(my_status, ST_INC | (tiles_prefix + count));
    }
    S.gbase[tid] = gexcl + tiles_prefix - excl;  // global position = gbase[d] + local sorted position
    __syncthreads();

    // ---- stream out the digit runs ------------------------------------------------------------------------
    for (int p = tid; p < cnt; p += SORT_THREADS) {
        const KeyT k = S.keys[p];
        const uint32_t d = digit_of<KeyT>(k, shift, depth_bits, depth_base);
        const uint32_t gp = S.gbase[d] + (uint32_t)p;
        keys_out[gp] = k;
        vals_out[gp] = S.vals[p];
    }
}

// ----------------------------------------------------------------------------------------------------------
// identify_tile_range! — utils.jl:56-78   (ranges pre-zeroed by the caller, rasterizer.jl:375)
// ----------------------------------------------------------------------------------------------------------
constexpr int RANGES_IPT = 4;  // keys per thread: two 128-bit loads instead of 2 x 4 scalar ones
__global__ void __launch_bounds__(256)
tile_ranges_kernel(const int64_t m, const uint64_t *__restrict__ keys, uint32_t *__restrict__ ranges, const bool aligned16) {
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * RANGES_IPT;
    if (i0 >= m) return;
    uint64_t k[RANGES_IPT];
    if (aligned16 && i0 + RANGES_IPT <= m) {
        const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(keys + i0);
        const ulonglong2 b = *reinterpret_cast<const ulonglong2 *>(keys + i0 + 2);
        k[0] = a.x; k[1] = a.y; k[2] = b.x; k[3] = b.y;
    } else {
#pragma unroll
        for (int j = 0; j < RANGES_IPT; j++) k[j] = i0 + j < m ? keys[i0 + j] : 0ull;
    }
    uint32_t prev = i0 > 0 ? (uint32_t)(keys[i0 - 1] >> 32) : 0u;
#pragma unroll
    for (int j = 0; j < RANGES_IPT; j++) {
        const int64_t i = i0 + j;
        if (i >= m) break;
        const uint32_t tile = (uint32_t)(k[j] >> 32);
        if (i == 0) {
            ranges[2 * (int64_t)tile] = 0u;
        } else if (tile != prev) {
            ranges[2 * (int64_t)prev + 1] = (uint32_t)i;
            ranges[2 * (int64_t)tile] = (uint32_t)i;
        }
        if (i == m - 1) ranges[2 * (int64_t)tile + 1] = (uint32_t)m;
        prev = tile;
    }
}

// ----------------------------------------------------------------------------------------------------------
// duplicate_with_keys! after the depth pre-sort — utils.jl:85-120, warp-cooperative and load-balanced
//
// Emission order is depth order, so the instance sort only needs the TILE of an instance: the key is the bare 32-bit
// tile id (the canonical 64-bit key is rebuilt from it on demand, materialize_keys_kernel).  A warp takes 32 consecutive
// Gaussians of the depth order; their output ranges are adjacent, and lane l writes output slots begin + l, begin + 32
// + l, ...: the owner of a slot is found by a 5-step binary search over the warp's 32 inclusive offsets (shuffles), the
// tile follows from the slot's position in the owner's rectangle.  Every store is a coalesced 128-byte line whatever
// the sizes of the rectangles (the per-Gaussian loop writes 32 scattered lines per instruction and runs as long as
// the largest rectangle of the warp).
// ----------------------------------------------------------------------------------------------------------
constexpr int DUPC_THREADS = 256;
__global__ void __launch_bounds__(DUPC_THREADS)
duplicate_coop_kernel(const int64_t n, const int32_t grid_x, const int32_t grid_y, const int32_t *__restrict__ radii,
                      const float2 *__restrict__ means2d, const int32_t *__restrict__ offsets,
                      const uint32_t *__restrict__ perm, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals,
                      const int passes, uint32_t *__restrict__ ghist) {
    __shared__ uint32_t sh[4 * 256];
    for (int k = threadIdx.x; k < passes * 256; k += DUPC_THREADS) sh[k] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    // a CTA walks several 256-Gaussian chunks: its histogram is flushed once (the flush is 256 x passes global atomics per
    // CTA on the same few hundred addresses — with one chunk per CTA that contention, not the emission, set the time)
    for (int64_t chunk = blockIdx.x; chunk * DUPC_THREADS < n; chunk += gridDim.x) {
    const int64_t j = chunk * DUPC_THREADS + threadIdx.x;  // emission slot
    uint32_t pack = 0, id = 0, end = 0;
    int cnt = 0;
    {
        const int64_t jc = j < n ? j : n - 1;
        end = (uint32_t)offsets[jc];  // inclusive scan in emission order: monotone, so the search below needs no flags
        if (j < n) {
            const int64_t i = (int64_t)perm[j];
            const int32_t r = radii[i];
            if (r > 0) {
                const float2 m = means2d[i];
                int32_t x0, y0, x1, y1;
                get_rect(m.x, m.y, r, grid_x, grid_y, x0, y0, x1, y1);
                cnt = (x1 - x0) * (y1 - y0);
                pack = (uint32_t)(x1 - x0) | ((uint32_t)x0 << 10) | ((uint32_t)y0 << 20);  // tile grids up to 1023 x 4095
            }
            id = (uint32_t)(i + 1);  // 1-based, as the reference emits
        }
    }
    const uint32_t off = end - (uint32_t)cnt;
    const uint32_t wbegin = __shfl_sync(0xffffffffu, off, 0), wend = __shfl_sync(0xffffffffu, end, 31);
    for (uint32_t s0 = wbegin; s0 < wend; s0 += 32) {  // warp-uniform trip count
        const uint32_t s = s0 + lane;
        const bool valid = s < wend;
        int g = 0;  // first lane whose inclusive offset exceeds s
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const uint32_t e = __shfl_sync(0xffffffffu, end, g + step - 1);
            if (e <= s) g += step;
        }
        g = g > 31 ? 31 : g;  // lanes past the warp's range (not stored)
        const uint32_t gp = __shfl_sync(0xffffffffu, pack, g), go = __shfl_sync(0xffffffffu, off, g);
        const uint32_t gid = __shfl_sync(0xffffffffu, id, g);
        uint32_t tile = 0;
        if (valid) {
            const uint32_t w = gp & 1023u, t = s - go;
            const uint32_t q = t / w;
            tile = ((gp >> 20) + q) * (uint32_t)grid_x + ((gp >> 10) & 1023u) + (t - q * w);
            keys[s] = tile;
            vals[s] = gid;
        }
        // digit histograms of the tile passes: neighbouring slots are neighbouring tiles — all-distinct low digits
        // (plain atomics, no conflicts), near-constant high digits (aggregate with MATCH.ANY, fast on few values)
        if (valid) atomicAdd(&sh[tile & 255u], 1u);
        for (int p = 1; p < passes; p++) {
            const uint32_t d = valid ? ((tile >> (8 * p)) & 255u) : 0xffffffffu;
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            if (valid && lane == __ffs(peers) - 1) atomicAdd(&sh[p * 256 + d], (uint32_t)__popc(peers));
        }
    }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < passes * 256; k += DUPC_THREADS)
        if (sh[k]) atomicAdd(&ghist[k], sh[k]);
}

// identify_tile_range! on bare tile ids
__global__ void __launch_bounds__(256)
tile_ranges32_kernel(const int64_t m, const uint32_t *__restrict__ tiles, uint32_t *__restrict__ ranges) {
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= m) return;
    uint32_t k[4];
    if (i0 + 4 <= m) {
        const uint4 a = *reinterpret_cast<const uint4 *>(tiles + i0);
        k[0] = a.x; k[1] = a.y; k[2] = a.z; k[3] = a.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) k[j] = i0 + j < m ? tiles[i0 + j] : 0u;
    }
    uint32_t prev = i0 > 0 ? tiles[i0 - 1] : 0u;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int64_t i = i0 + j;
        if (i >= m) break;
        const uint32_t tile = k[j];
        if (i == 0) {
            ranges[2 * (int64_t)tile] = 0u;
        } else if (tile != prev) {
            ranges[2 * (int64_t)prev + 1] = (uint32_t)i;
            ranges[2 * (int64_t)tile] = (uint32_t)i;
        }
        if (i == m - 1) ranges[2 * (int64_t)tile + 1] = (uint32_t)m;
        prev = tile;
    }
}

// Tiles by falling instance count: a counting sort on 4 log2(count + 1) (64 buckets, 19 % resolution) by ONE CTA — a few
// thousand tiles.  The compositing kernels take their tiles in this order: CTAs are dispatched in launch order, so the
// grid ends on its shortest tiles instead of whatever the raster order leaves for last (the tail cost 3-5 % of both
// kernels on the uniform synthetic scenes; on real, clustered scenes the heaviest tile can be 10x the median).
__global__ void __launch_bounds__(1024)
tile_order_kernel(const int n_tiles, const uint2 *__restrict__ ranges, uint32_t *__restrict__ order) {
    __shared__ uint32_t s_cnt[64], s_base[64];
    if (threadIdx.x < 64) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    auto bucket = [](uint2 r) {
        const int b = (int)(4.0f * __log2f((float)(r.y - r.x) + 1.0f));
        return 63 - (b > 63 ? 63 : b);  // heaviest first
    };
    // neighbouring tiles carry similar loads: a warp's 32 tiles fall into a few buckets, so the shared-memory atomics are
    // aggregated per bucket with MATCH.ANY (32-way conflicts on 64 counters otherwise)
    const int lane = threadIdx.x & 31;
    const int n_round = (n_tiles + (int)blockDim.x - 1) / (int)blockDim.x * (int)blockDim.x;  // warp-uniform trip count
    for (int t = threadIdx.x; t < n_round; t += blockDim.x) {
        const int b = t < n_tiles ? bucket(ranges[t]) : 64 + lane;
        const unsigned peers = __match_any_sync(0xffffffffu, b);
        if (t < n_tiles && lane == __ffs(peers) - 1) atomicAdd(&s_cnt[b], (uint32_t)__popc(peers));
    }
    __syncthreads();
    if (threadIdx.x < 32) {  // exclusive scan of the 64 counters by one warp
        const uint32_t c0 = s_cnt[2 * lane], c1 = s_cnt[2 * lane + 1];
        uint32_t incl = c0 + c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        s_base[2 * lane] = incl - c0 - c1;
        s_base[2 * lane + 1] = incl - c1;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n_round; t += blockDim.x) {
        const int b = t < n_tiles ? bucket(ranges[t]) : 64 + lane;
        const unsigned peers = __match_any_sync(0xffffffffu, b);
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (t < n_tiles && lane == leader) base = atomicAdd(&s_base[b], (uint32_t)__popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (t < n_tiles) order[base + __popc(peers & ((1u << lane) - 1u))] = (uint32_t)t;
    }
}

// the canonical sorted keys (tile << 32 | bits(depth)) from the sorted tile ids and Gaussian ids
__global__ void __launch_bounds__(256)
materialize_keys_kernel(const int64_t m, const uint32_t *__restrict__ tiles, const uint32_t *__restrict__ vals,
                        const float *__restrict__ depths, uint64_t *__restrict__ keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) keys[i] = ((uint64_t)tiles[i] << 32) | (uint64_t)__float_as_uint(depths[vals[i] - 1u]);
}

int bit_length(uint64_t x) {
    int b = 0;
    while (x) { b++; x >>= 1; }
    return b;
}

}  // namespace

// ---- host side ---------------------------------------------------------------------------------------------
size_t scan_state_words(int64_t n) {  // in 64-bit words
    return 1 + (size_t)((n + SCAN_TILE - 1) / SCAN_TILE);
}

void launch_scan_tiles(int64_t n, const int32_t *tiles_touched, const uint32_t *perm, int32_t *offsets,
                       uint32_t *scan_state, int64_t *total_dev, cudaStream_t s) {
    if (n <= 0) return;
    const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    cudaMemsetAsync(scan_state, 0, scan_state_words(n) * sizeof(unsigned long long), s);
    scan_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(n, tiles_touched, perm, offsets,
                                                        reinterpret_cast<unsigned long long *>(scan_state), total_dev);
    count_launch();
}

void launch_duplicate(const DevCamera &cam, int64_t n, const GeomPtrs &g, const int32_t *offsets, const uint32_t *perm,
                      uint64_t *keys, uint32_t *vals, const SortPlan &plan, uint32_t *ghist, cudaStream_t s) {
    if (n <= 0) return;
    const int64_t blocks = (n + DUP_THREADS - 1) / DUP_THREADS;
    duplicate_kernel<<<(unsigned)blocks, DUP_THREADS, 0, s>>>(n, cam.grid_x, cam.grid_y, g.radii, g.means2d, g.depths,
                                                             offsets, perm, keys, vals, plan.depth_bits, plan.depth_base,
                                                             plan.bit_lo, plan.passes, ghist);
    count_launch();
}

// plan = presort_plan(...); ghist = sort_prepare(plan, ...): the histograms are ready for launch_sort_pairs afterwards
void launch_presort_keys(int64_t n, const GeomPtrs &g, const SortPlan &plan, uint64_t *keys, uint32_t *vals,
                         uint32_t *ghist, cudaStream_t s) {
    if (n <= 0) return;
    int64_t blocks = (n + 256 * 8 - 1) / (256 * 8);  // ~8 Gaussians per thread: few histogram flushes
    if (blocks > 148 * 8) blocks = 148 * 8;
    presort_keys_kernel<<<(unsigned)blocks, 256, 0, s>>>(n, g.radii, g.depths, plan.depth_bits, plan.depth_base,
                                                         plan.passes, keys, vals, ghist);
    count_launch();
}

// the Gaussians' own sort: flag bit + depth bits
SortPlan presort_plan(const SortPlan &plan) {
    SortPlan p = plan;
    p.tile_bits = 1;
    p.bit_lo = 0;
    p.small = 1;
    p.passes = (p.tile_bits + p.depth_bits + 7) / 8;
    return p;
}

// the instance sort after a depth pre-sort: tile digits only
SortPlan tile_only_plan(const SortPlan &plan) {
    SortPlan p = plan;
    p.bit_lo = plan.depth_bits;
    p.passes = (p.tile_bits + 7) / 8;
    return p;
}

SortPlan make_sort_plan(int64_t n_tiles, float near_plane, float far_plane) {
    SortPlan p;
    p.tile_bits = bit_length((uint64_t)(n_tiles > 1 ? n_tiles - 1 : 1));
    uint32_t nb, fb;
    memcpy(&nb, &near_plane, 4);
    memcpy(&fb, &far_plane, 4);
    if (near_plane > 0.f && far_plane > near_plane) {  // depths in (near, far): bits monotone, sign bit clear
        p.depth_base = nb;
        p.depth_bits = bit_length((uint64_t)(fb - nb));
    } else {  // degenerate planes: keep the raw 32 depth bits, exactly the reference's ordering
        p.depth_base = 0;
        p.depth_bits = 32;
    }
    if (p.depth_bits < 1) p.depth_bits = 1;
    p.bit_lo = 0;
    p.small = 0;
    p.passes = (p.tile_bits + p.depth_bits + 7) / 8;
    return p;
}

// keys per thread of the pass kernel: 16 for the instance sort; the Gaussians' sort (plan.small) is a single partial
// wave either way, so it takes the 2048-key tiles whose per-CTA latency is half
int sort_ipt(const SortPlan &plan) {
    static int v = -1, vs = -1;
    if (v < 0) {
        const char *e = getenv("GSR_SORT_IPT");
        v = (e && atoi(e) == 8) ? 8 : 16;
        const char *es = getenv("GSR_PRESORT_IPT");
        vs = (es && atoi(es) == 16) ? 16 : 8;
    }
    return plan.small ? vs : v;
}
static int64_t sort_tile_keys(const SortPlan &plan) { return (int64_t)SORT_THREADS * sort_ipt(plan); }

size_t sort_temp_words(int64_t m, const SortPlan &plan) {  // in 32-bit words
    const size_t tiles = (size_t)((m + sort_tile_keys(plan) - 1) / sort_tile_keys(plan));
    return (size_t)plan.passes * 256 /* ghist */ + (size_t)plan.passes * tiles * 256 /* status */ +
           SORT_MAX_PASSES /* tile counters */;
}

uint32_t *sort_prepare(const SortPlan &plan, int64_t m, uint32_t *temp_words, cudaStream_t s) {
    // zero {digit histograms, look-back status, tile counters}; returns the histogram buffer [passes][256]
    cudaMemsetAsync(temp_words, 0, sort_temp_words(m, plan) * sizeof(uint32_t), s);
    return temp_words;
}

void launch_sort_pairs(const SortPlan &plan, int64_t m, const uint64_t *keys_in, const uint32_t *vals_in,
                       uint64_t *keys_out, uint32_t *vals_out, uint64_t *keys_tmp, uint32_t *vals_tmp,
                       uint32_t *temp_words, bool hist_ready, cudaStream_t s) {
    if (m <= 0) return;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(onesweep_kernel<uint64_t, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortSmem<uint64_t, 8>));
        cudaFuncSetAttribute(onesweep_kernel<uint64_t, 16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortSmem<uint64_t, 16>));
        cudaFuncSetAttribute(onesweep_kernel<uint64_t, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortSmem<uint64_t, 8>));
        cudaFuncSetAttribute(onesweep_kernel<uint64_t, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortSmem<uint64_t, 16>));
        attr_set = true;
    }
    const int ipt = sort_ipt(plan);
    const size_t tiles = (size_t)((m + sort_tile_keys(plan) - 1) / sort_tile_keys(plan));
    uint32_t *ghist = temp_words;
    uint32_t *status = ghist + (size_t)plan.passes * 256;
    uint32_t *counters = status + (size_t)plan.passes * tiles * 256;
    if (!hist_ready) {
        int hb = (int)((m + 256 * 16 - 1) / (256 * 16));
        if (hb > 148 * 8) hb = 148 * 8;
        hist_kernel<<<hb, 256, 0, s>>>(keys_in, m, plan.depth_bits, plan.depth_base, plan.bit_lo, plan.passes, ghist);
        count_launch();
    }
    const uint64_t *ksrc = keys_in;
    const uint32_t *vsrc = vals_in;
    static int rank_mode = -1;  // GSR_SORT_RANK = match | ballot | auto (default)
    if (rank_mode < 0) {
        const char *e = getenv("GSR_SORT_RANK");
        rank_mode = (e && !strcmp(e, "match")) ? 0 : ((e && !strcmp(e, "ballot")) ? 1 : 2);
    }
    const int total_bits = plan.tile_bits + plan.depth_bits;
    for (int p = 0; p < plan.passes; p++) {
        const int shift = plan.bit_lo + 8 * p;
        const int digit_bits = total_bits - shift < 8 ? total_bits - shift : 8;
        const bool has_tile_bits = shift + 8 > plan.depth_bits;
        const int ballot_bits = (rank_mode == 1 || (rank_mode == 2 && has_tile_bits)) ? digit_bits : 0;
        // the chain must end in (keys_out, vals_out) and never write the input
        const bool to_out = ((plan.passes - 1 - p) % 2) == 0;
        uint64_t *kdst = to_out ? keys_out : keys_tmp;
        uint32_t *vdst = to_out ? vals_out : vals_tmp;
#define GSR_ONESWEEP(IPT, BAL)                                                                                       \
    onesweep_kernel<uint64_t, IPT, BAL><<<(unsigned)tiles, SORT_THREADS, sizeof(SortSmem<uint64_t, IPT>), s>>>(          \
        ksrc, vsrc, kdst, vdst, m, shift, plan.depth_bits, plan.depth_base, ghist + (size_t)p * 256,                    \
        status + (size_t)p * tiles * 256, counters + p, ballot_bits)
        if (ipt == 16) {
            if (ballot_bits) GSR_ONESWEEP(16, true); else GSR_ONESWEEP(16, false);
        } else {
            if (ballot_bits) GSR_ONESWEEP(8, true); else GSR_ONESWEEP(8, false);
        }
#undef GSR_ONESWEEP
        count_launch();
        ksrc = kdst;
        vsrc = vdst;
    }
}

// ---- instance binning on 32-bit tile keys (after the depth pre-sort) -----------------------------------------
void launch_duplicate_tiles(const DevCamera &cam, int64_t n, const GeomPtrs &g, const int32_t *offsets, const uint32_t *perm,
                            uint32_t *tiles, uint32_t *vals, const SortPlan &plan, uint32_t *ghist, cudaStream_t s) {
    if (n <= 0) return;
    int64_t blocks = (n + DUPC_THREADS - 1) / DUPC_THREADS;
    static int per_sm = -1;
    if (per_sm < 0) {
        const char *e = getenv("GSR_DUP_CTAS_PER_SM");
        per_sm = (e && atoi(e) > 0) ? atoi(e) : 4;
    }
    if (blocks > 148 * per_sm) blocks = 148 * per_sm;
    duplicate_coop_kernel<<<(unsigned)blocks, DUPC_THREADS, 0, s>>>(n, cam.grid_x, cam.grid_y, g.radii, g.means2d, offsets,
                                                                  perm, tiles, vals, plan.passes, ghist);
    count_launch();
}

// plan = tile_only_plan(...): radix passes over the tile id's bits, histograms ready (launch_duplicate_tiles).
// tiles_in / vals_in are left intact; (tiles_tmp, vals_tmp) is scratch of size m.
void launch_sort_tiles(const SortPlan &plan, int64_t m, const uint32_t *tiles_in, const uint32_t *vals_in, uint32_t *tiles_out,
                       uint32_t *vals_out, uint32_t *tiles_tmp, uint32_t *vals_tmp, uint32_t *temp_words, cudaStream_t s) {
    if (m <= 0) return;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(onesweep_kernel<uint32_t, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortSmem<uint32_t, 16>));
        cudaFuncSetAttribute(onesweep_kernel<uint32_t, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortSmem<uint32_t, 8>));
        attr_set = true;
    }
    const int ipt = sort_ipt(plan);
    const size_t tiles = (size_t)((m + sort_tile_keys(plan) - 1) / sort_tile_keys(plan));
    uint32_t *ghist = temp_words;
    uint32_t *status = ghist + (size_t)plan.passes * 256;
    uint32_t *counters = status + (size_t)plan.passes * tiles * 256;
    const uint32_t *ksrc = tiles_in, *vsrc = vals_in;
    for (int p = 0; p < plan.passes; p++) {
        const int shift = 8 * p;
        const int digit_bits = plan.tile_bits - shift < 8 ? plan.tile_bits - shift : 8;
        const bool to_out = ((plan.passes - 1 - p) % 2) == 0;  // the chain ends in (tiles_out, vals_out)
        uint32_t *kdst = to_out ? tiles_out : tiles_tmp, *vdst = to_out ? vals_out : vals_tmp;
        if (ipt == 16)
            onesweep_kernel<uint32_t, 16, true><<<(unsigned)tiles, SORT_THREADS, sizeof(SortSmem<uint32_t, 16>), s>>>(
                ksrc, vsrc, kdst, vdst, m, shift, 0, 0u, ghist + (size_t)p * 256, status + (size_t)p * tiles * 256, counters + p, digit_bits);
        else
            onesweep_kernel<uint32_t, 8, true><<<(unsigned)tiles, SORT_THREADS, sizeof(SortSmem<uint32_t, 8>), s>>>(
                ksrc, vsrc, kdst, vdst, m, shift, 0, 0u, ghist + (size_t)p * 256, status + (size_t)p * tiles * 256, counters + p, digit_bits);
        count_launch();
        ksrc = kdst;
        vsrc = vdst;
    }
}

void launch_tile_ranges32(int64_t m, const uint32_t *tiles_sorted, uint32_t *ranges, cudaStream_t s) {
    if (m <= 0) return;
    tile_ranges32_kernel<<<(unsigned)((m + 1023) / 1024), 256, 0, s>>>(m, tiles_sorted, ranges);
    count_launch();
}

void launch_tile_order(int64_t n_tiles, const uint32_t *ranges, uint32_t *order, cudaStream_t s) {
    if (n_tiles <= 0) return;
    tile_order_kernel<<<1, 1024, 0, s>>>((int)n_tiles, reinterpret_cast<const uint2 *>(ranges), order);
    count_launch();
}

void launch_materialize_keys(int64_t m, const uint32_t *tiles_sorted, const uint32_t *vals_sorted, const float *depths,
                             uint64_t *keys_sorted, cudaStream_t s) {
    if (m <= 0) return;
    materialize_keys_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(m, tiles_sorted, vals_sorted, depths, keys_sorted);
    count_launch();
}

void launch_tile_ranges(int64_t m, const uint64_t *keys_sorted, uint32_t *ranges, cudaStream_t s) {
    if (m <= 0) return;
    const int64_t per_block = 256 * RANGES_IPT;
    tile_ranges_kernel<<<(unsigned)((m + per_block - 1) / per_block), 256, 0, s>>>(m, keys_sorted, ranges,
                                                                                    (reinterpret_cast<uintptr_t>(keys_sorted) & 15) == 0);
    count_launch();
}
