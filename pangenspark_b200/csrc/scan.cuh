// scan.cuh -- device-wide scans (hand-written, three phases: tile reduce, sums scan, tile scan).
//
// Generic over a load functor (element i -> T), a store functor (element i, T) and an associative op,
// so the same kernels serve: exclusive sums of counts/flags (stream compaction), running maxima
// (group-head propagation in prefix doubling) and the right-to-left minimum that fills the k-mer table.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace drlz {

extern thread_local unsigned long long g_launches;      // kernels launched by the calling thread (defined in api.cu)

static const int kScanThreads = 256;
static const int kScanItems = 16;
static const int kScanTile = kScanThreads * kScanItems;

struct OpSum { template <typename T> __device__ __forceinline__ T operator()(T a, T b) const { return a + b; } };
struct OpMax { template <typename T> __device__ __forceinline__ T operator()(T a, T b) const { return a > b ? a : b; } };
struct OpMin { template <typename T> __device__ __forceinline__ T operator()(T a, T b) const { return a < b ? a : b; } };

template <typename T, typename Op>
__device__ __forceinline__ T warp_inclusive_scan(T v, Op op, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v = op(o, v);
    }
    return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, *total = block total
template <typename T, typename Op>
__device__ __forceinline__ T block_exclusive_scan(T v, Op op, T identity, T* total, T* smem /*[33]*/) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    T inc = warp_inclusive_scan(v, op, lane);
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        T w = lane < nw ? smem[lane] : identity;
        T winc = warp_inclusive_scan(w, op, lane);
        if (lane < nw) smem[lane] = winc;
        if (lane == nw - 1) smem[32] = winc;
    }
    __syncthreads();
    T exc = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) exc = identity;
    T wbase = warp ? smem[warp - 1] : identity;
    *total = smem[32];
    T r = op(wbase, exc);
    __syncthreads();
    return r;
}

template <typename T, typename Load, typename Op>
__global__ void __launch_bounds__(kScanThreads) k_scan_reduce(Load ld, uint64_t n, T* sums, Op op, T identity) {
    __shared__ T sm[33];
    uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    T acc = identity;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        uint64_t i = base + k;
        if (i < n) acc = op(acc, ld(i));
    }
    T total;
    block_exclusive_scan(acc, op, identity, &total, sm);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// single block: exclusive scan of sums[0..nb) in place; sums[nb] = grand total
template <typename T, typename Op>
__global__ void __launch_bounds__(1024) k_scan_sums(T* sums, uint64_t nb, Op op, T identity) {
    __shared__ T sm[33];
    __shared__ T carry_s;
    if (threadIdx.x == 0) carry_s = identity;
    __syncthreads();
    for (uint64_t b0 = 0; b0 < nb; b0 += blockDim.x) {
        uint64_t i = b0 + threadIdx.x;
        T v = i < nb ? sums[i] : identity;
        T total;
        T exc = block_exclusive_scan(v, op, identity, &total, sm);
        T carry = carry_s;
        if (i < nb) sums[i] = op(carry, exc);
        __syncthreads();
        if (threadIdx.x == 0) carry_s = op(carry, total);
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[nb] = carry_s;
}

template <typename T, typename Load, typename Store, typename Op, bool kInclusive>
__global__ void __launch_bounds__(kScanThreads)
k_scan_tile(Load ld, Store st, uint64_t n, const T* sums, Op op, T identity) {
    __shared__ T sm[33];
    uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    T v[kScanItems];
    T acc = identity;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        uint64_t i = base + k;
        v[k] = i < n ? ld(i) : identity;
        acc = op(acc, v[k]);
    }
    T total;
    T run = block_exclusive_scan(acc, op, identity, &total, sm);
    run = op(sums[blockIdx.x], run);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        uint64_t i = base + k;
        T nxt = op(run, v[k]);
        if (i < n) st(i, kInclusive ? nxt : run);
        run = nxt;
    }
}

static inline uint64_t scan_tmp_elems(uint64_t n) { return (n + kScanTile - 1) / kScanTile + 2; }

// out(i) = op over elements before i (exclusive) or up to i (inclusive).  tmp: scan_tmp_elems(n) of T.
// The grand total is left in tmp[number_of_tiles].
template <typename T, typename Load, typename Store, typename Op, bool kInclusive>
void device_scan(Load ld, Store st, uint64_t n, T* tmp, Op op, T identity, cudaStream_t s) {
    if (n == 0) { cudaMemsetAsync(tmp, 0, sizeof(T), s); return; }   // only meaningful for sums
    uint64_t nb = (n + kScanTile - 1) / kScanTile;
    g_launches += 3;
    k_scan_reduce<T, Load, Op><<<(unsigned)nb, kScanThreads, 0, s>>>(ld, n, tmp, op, identity);
    k_scan_sums<T, Op><<<1, 1024, 0, s>>>(tmp, nb, op, identity);
    k_scan_tile<T, Load, Store, Op, kInclusive><<<(unsigned)nb, kScanThreads, 0, s>>>(ld, st, n, tmp, op, identity);
}

// common functors
template <typename TIn, typename T> struct LoadArr {
    const TIn* p;
    __device__ __forceinline__ T operator()(uint64_t i) const { return (T)p[i]; }
};
template <typename T> struct StoreArr {
    T* p;
    __device__ __forceinline__ void operator()(uint64_t i, T v) const { p[i] = v; }
};

}  // namespace drlz
