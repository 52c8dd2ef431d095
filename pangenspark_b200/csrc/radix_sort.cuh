// radix_sort.cuh -- hand-written LSD radix sort of (key, uint32 value) pairs, 8-bit digits.
//
// Per pass: (1) per-tile digit histograms in shared memory -> digit-major matrix in HBM,
//           (2) exclusive scan of the matrix (scan.cuh),
//           (3) stable scatter: warp-level multisplit with __match_any_sync, warp-private digit
//               counters in shared memory, cross-warp prefix per digit, the tile staged in digit order in
//               shared memory, then stores of contiguous pieces (a tile of 4096 keys puts ~16 consecutive
//               keys into each of the 256 runs).
// Passes whose digit is constant over all keys are skipped by the caller through the bit range.
// Used by the suffix-array prefix doubling (sa_build.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "scan.cuh"

namespace drlz {

static const int kSortThreads = 256;
static const int kSortItems = 16;
static const int kSortTile = kSortThreads * kSortItems;   // 4096 keys per CTA
static const int kSortWarps = kSortThreads / 32;

template <typename K>
__global__ void __launch_bounds__(kSortThreads)
k_sort_hist(const K* __restrict__ keys, uint64_t n, int shift, uint32_t* __restrict__ hist, uint32_t ntiles) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    uint64_t base = (uint64_t)blockIdx.x * kSortTile;
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        uint64_t i = base + (uint64_t)k * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[(unsigned)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(uint64_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

// Stable scatter of one tile.  Ranks come from a warp-level multisplit (__match_any_sync) with warp-private digit
// counters; the tile is then put in digit order in shared memory, so that the stores to the 256 output runs are
// contiguous pieces (~16 keys each) written by consecutive threads instead of 32 unrelated 8-byte stores per warp.
template <typename K>
__global__ void __launch_bounds__(kSortThreads, 4)      // 64 registers: 4 CTAs per SM also fit their 43 KB of shared memory
k_sort_scatter(const K* __restrict__ keys, const uint32_t* __restrict__ vals, K* __restrict__ okeys,
               uint32_t* __restrict__ ovals, uint64_t n, int shift, const uint32_t* __restrict__ offs,
               uint32_t ntiles) {
    __shared__ uint32_t wcnt[kSortWarps][256];
    __shared__ uint32_t dstart[256];          // first tile-local slot of each digit
    __shared__ uint32_t goff[256];            // global slot of that first element
    __shared__ uint32_t sm[33];
    __shared__ K stage[kSortTile];            // the tile in digit order (keys first, then reused for the values)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&wcnt[0][0])[i] = 0;
    __syncthreads();
    // warp w owns the contiguous slice [w*512, w*512+512) of the tile; item k of lane l sits at k*32+l,
    // so (k, lane) order == memory order == stable order.
    const uint64_t tile0 = (uint64_t)blockIdx.x * kSortTile;
    const uint64_t base = tile0 + (uint64_t)warp * (32 * kSortItems);
    const uint32_t tile_n = (uint32_t)(n - tile0 < (uint64_t)kSortTile ? n - tile0 : (uint64_t)kSortTile);
    K key[kSortItems];
    uint32_t rank[kSortItems];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        uint64_t i = base + (uint64_t)k * 32 + lane;
        bool valid = i < n;
        key[k] = valid ? keys[i] : (K)0;
        unsigned d = valid ? ((unsigned)(key[k] >> shift) & 255u) : 256u;
        unsigned peers = __match_any_sync(0xffffffffu, d);
        uint32_t before = 0;
        if (valid) before = wcnt[warp][d];
        __syncwarp();
        rank[k] = before + __popc(peers & lt);
        if (valid && (peers & lt) == 0) wcnt[warp][d] = before + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    {   // digit t: exclusive prefix over warps (tile-local), digit totals -> exclusive scan over digits
        const int t = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) { uint32_t c = wcnt[w][t]; wcnt[w][t] = run; run += c; }
        uint32_t total;
        uint32_t ds = block_exclusive_scan<uint32_t, OpSum>(run, OpSum(), 0u, &total, sm);
        dstart[t] = ds;
        goff[t] = offs[(uint64_t)t * ntiles + blockIdx.x];
    }
    __syncthreads();
    uint32_t slot[kSortItems];
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        uint64_t i = base + (uint64_t)k * 32 + lane;
        slot[k] = 0;
        if (i < n) {
            unsigned d = (unsigned)(key[k] >> shift) & 255u;
            slot[k] = dstart[d] + wcnt[warp][d] + rank[k];
            stage[slot[k]] = key[k];
        }
    }
    __syncthreads();
    uint32_t dst[kSortItems];
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const uint32_t j = (uint32_t)k * kSortThreads + threadIdx.x;      // consecutive threads, consecutive slots
        if (j < tile_n) {
            const K kk = stage[j];
            const unsigned d = (unsigned)(kk >> shift) & 255u;
            dst[k] = goff[d] + (j - dstart[d]);
            okeys[dst[k]] = kk;
        }
    }
    __syncthreads();
    uint32_t* vstage = reinterpret_cast<uint32_t*>(stage);
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        uint64_t i = base + (uint64_t)k * 32 + lane;
        if (i < n) vstage[slot[k]] = vals[i];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const uint32_t j = (uint32_t)k * kSortThreads + threadIdx.x;
        if (j < tile_n) ovals[dst[k]] = vstage[j];
    }
}

// Second form of the scatter (sort_variant 1).  What the profile of the first one shows (ncu, 48 M pairs: 0.58 ms per pass,
// 2.4 TB/s of DRAM traffic, 23 % issue slots used): a third of the stall samples wait for a MATCH result (16 of them, each
// consumed at once), a sixth for the values, which are only requested after the keys have left, and the 64-register
// budget of 4 CTAs per SM spills.  Here all 16 matches are issued back to back before the first one is used, the values
// are requested as soon as the keys sit in shared memory (they arrive while the keys are being written out), and the
// output address of a slot is recomputed from a byte array of digits instead of being held in 16 registers.
template <typename K, int MINB>
__global__ void __launch_bounds__(kSortThreads, MINB)
k_sort_scatter2(const K* __restrict__ keys, const uint32_t* __restrict__ vals, K* __restrict__ okeys,
                uint32_t* __restrict__ ovals, uint64_t n, int shift, const uint32_t* __restrict__ offs,
                uint32_t ntiles) {
    __shared__ uint32_t wcnt[kSortWarps][256];
    __shared__ uint32_t dstart[256];
    __shared__ uint32_t goff[256];
    __shared__ uint32_t sm[33];
    __shared__ K stage[kSortTile];
    __shared__ uint8_t sdig[kSortTile];       // digit of the key in each slot of the staged tile
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&wcnt[0][0])[i] = 0;
    const uint64_t tile0 = (uint64_t)blockIdx.x * kSortTile;
    const uint64_t base = tile0 + (uint64_t)warp * (32 * kSortItems);
    const uint32_t tile_n = (uint32_t)(n - tile0 < (uint64_t)kSortTile ? n - tile0 : (uint64_t)kSortTile);
    K key[kSortItems];
    uint32_t rank[kSortItems];                // first the peer masks, then the ranks, then the slots
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const uint64_t i = base + (uint64_t)k * 32 + lane;
        key[k] = i < n ? keys[i] : (K)0;
    }
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const uint64_t i = base + (uint64_t)k * 32 + lane;
        const unsigned d = i < n ? ((unsigned)(key[k] >> shift) & 255u) : 256u;
        rank[k] = __match_any_sync(0xffffffffu, d);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const uint64_t i = base + (uint64_t)k * 32 + lane;
        const bool valid = i < n;
        const unsigned d = (unsigned)(key[k] >> shift) & 255u;
        const uint32_t peers = rank[k];
        uint32_t before = 0;
        if (valid) before = wcnt[warp][d];
        __syncwarp();
        rank[k] = before + __popc(peers & lt);
        if (valid && (peers & lt) == 0) wcnt[warp][d] = before + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    {
        const int t = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) { uint32_t c = wcnt[w][t]; wcnt[w][t] = run; run += c; }
        uint32_t total;
        uint32_t ds = block_exclusive_scan<uint32_t, OpSum>(run, OpSum(), 0u, &total, sm);
        dstart[t] = ds;
        goff[t] = offs[(uint64_t)t * ntiles + blockIdx.x];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const uint64_t i = base + (uint64_t)k * 32 + lane;
        if (i < n) {
            const unsigned d = (unsigned)(key[k] >> shift) & 255u;
            const uint32_t sl = dstart[d] + wcnt[warp][d] + rank[k];
            rank[k] = sl;
            stage[sl] = key[k];
            sdig[sl] = (uint8_t)d;
        }
    }
    uint32_t val[kSortItems];                 // requested now, used after the keys have been written out
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const uint64_t i = base + (uint64_t)k * 32 + lane;
        val[k] = i < n ? vals[i] : 0u;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const uint32_t j = (uint32_t)k * kSortThreads + threadIdx.x;
        if (j < tile_n) {
            const unsigned d = sdig[j];
            okeys[goff[d] + (j - dstart[d])] = stage[j];
        }
    }
    __syncthreads();
    uint32_t* vstage = reinterpret_cast<uint32_t*>(stage);
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const uint64_t i = base + (uint64_t)k * 32 + lane;
        if (i < n) vstage[rank[k]] = val[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const uint32_t j = (uint32_t)k * kSortThreads + threadIdx.x;
        if (j < tile_n) {
            const unsigned d = sdig[j];
            ovals[goff[d] + (j - dstart[d])] = vstage[j];
        }
    }
}

extern int g_sort_variant;                    // 0: k_sort_scatter, 1 / 2: k_sort_scatter2 at 4 / 3 CTAs per SM (api.cu, option sort_variant)

struct SortWorkspace {
    uint32_t* hist;      // 256 * ntiles (+ scan tmp behind it)
    uint32_t* scan_tmp;
};

static inline uint64_t sort_hist_elems(uint64_t n) { return 256ull * ((n + kSortTile - 1) / kSortTile); }

// Sorts by bits [begin_bit, end_bit) of the key.  Result lands in (*keys_a, *vals_a) (pointers are swapped
// as needed).  n < 2^32.  Returns the number of passes executed.
template <typename K>
int radix_sort_pairs(K** keys_a, K** keys_b, uint32_t** vals_a, uint32_t** vals_b, uint64_t n, int begin_bit,
                     int end_bit, uint32_t* hist, uint32_t* scan_tmp, cudaStream_t s) {
    if (n <= 1) return 0;
    uint32_t ntiles = (uint32_t)((n + kSortTile - 1) / kSortTile);
    uint64_t hn = 256ull * ntiles;
    int passes = 0;
    for (int shift = begin_bit; shift < end_bit; shift += 8) {
        g_launches += 2;
        k_sort_hist<K><<<ntiles, kSortThreads, 0, s>>>(*keys_a, n, shift, hist, ntiles);
        device_scan<uint32_t, LoadArr<uint32_t, uint32_t>, StoreArr<uint32_t>, OpSum, false>(
            LoadArr<uint32_t, uint32_t>{hist}, StoreArr<uint32_t>{hist}, hn, scan_tmp, OpSum(), 0u, s);
        if (g_sort_variant == 1) k_sort_scatter2<K, 4><<<ntiles, kSortThreads, 0, s>>>(*keys_a, *vals_a, *keys_b, *vals_b, n, shift, hist, ntiles);
        else if (g_sort_variant == 2) k_sort_scatter2<K, 3><<<ntiles, kSortThreads, 0, s>>>(*keys_a, *vals_a, *keys_b, *vals_b, n, shift, hist, ntiles);
        else k_sort_scatter<K><<<ntiles, kSortThreads, 0, s>>>(*keys_a, *vals_a, *keys_b, *vals_b, n, shift, hist, ntiles);
        K* tk = *keys_a; *keys_a = *keys_b; *keys_b = tk;
        uint32_t* tv = *vals_a; *vals_a = *vals_b; *vals_b = tv;
        ++passes;
    }
    return passes;
}

}  // namespace drlz
