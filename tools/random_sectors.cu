// random_sectors.cu -- what HBM gives a kernel that reads 32-byte sectors at random addresses.
//
// k_spec's lookups (k-mer table -> suffix array -> text) are reads of one or two sectors at addresses spread over
// ~0.6 GB; adding loads to it that nobody waits for makes it slower in proportion to the sectors, not to the dependent
// steps (profiles/README.md, round 2): its bound is the rate at which the memory system serves scattered sectors, far
// below the streaming bandwidth.  This tool measures that rate with nothing of the library involved:
//     every thread reads `per_thread` sectors of a `MiB` buffer at addresses from a hash of (thread, step),
//     `mlp` of them in flight per thread (independent), full occupancy; one 4-byte word of every sector is used.
// Prints sectors/s, the equivalent GB/s at 32 B per sector, and dram__bytes-style traffic is left to ncu.
// build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/bin/random_sectors tools/random_sectors.cu
// run:   tools/bin/random_sectors [MiB=1024] [per_thread=64] [chain=0|1] [l2_fetch_bytes=0|32|64|128]
//        chain=1: every address depends on the value read before it (one miss in flight per thread, as in a lookup)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <int MLP, bool CHAIN>
__global__ void __launch_bounds__(256) k_gather(const uint32_t* __restrict__ buf, uint32_t sectors, int per_thread, uint32_t* out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0, carry = 0;
    for (int s = 0; s < per_thread; s += MLP) {
        uint32_t v[MLP];
#pragma unroll
        for (int j = 0; j < MLP; ++j) {
            const uint32_t a = mix(t * 0x9E3779B9u + (uint32_t)(s + j) + carry) % sectors;
            v[j] = buf[(size_t)a * 8];
        }
#pragma unroll
        for (int j = 0; j < MLP; ++j) acc ^= v[j];
        if (CHAIN) carry = acc;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <int MLP, bool CHAIN>
static void run(const uint32_t* buf, uint32_t sectors, int per_thread, uint32_t* out, const char* name) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8 * 4;                       // four waves of full occupancy (8 x 256 threads per SM)
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; ++w) k_gather<MLP, CHAIN><<<blocks, 256>>>(buf, sectors, per_thread, out);
    cudaEventRecord(e0);
    const int reps = 5;
    for (int r = 0; r < reps; ++r) k_gather<MLP, CHAIN><<<blocks, 256>>>(buf, sectors, per_thread, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double n = (double)blocks * 256 * per_thread * reps;
    printf("{\"pattern\": \"%s\", \"mlp\": %d, \"sectors_per_s\": %.4g, \"gb_per_s_at_32B\": %.1f, \"ms_per_launch\": %.3f}\n", name, MLP,
           n / (ms * 1e-3), n * 32 / (ms * 1e-3) / 1e9, ms / reps);
}

int main(int argc, char** argv) {
    const size_t mib = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1024;
    const int per_thread = argc > 2 ? atoi(argv[2]) : 64;
    const int chain = argc > 3 ? atoi(argv[3]) : 0;
    const int fetch = argc > 4 ? atoi(argv[4]) : 0;            // cudaLimitMaxL2FetchGranularity: 32 / 64 / 128 bytes (0 = leave the default)
    if (fetch) {
        cudaError_t fe = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)fetch);
        size_t got = 0;
        cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        printf("{\"l2_fetch_granularity_requested\": %d, \"set\": \"%s\", \"now\": %zu}\n", fetch, cudaGetErrorString(fe), got);
    } else {
        size_t got = 0;
        cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        printf("{\"l2_fetch_granularity_default\": %zu}\n", got);
    }
    const size_t bytes = mib << 20;
    uint32_t *buf = nullptr, *out = nullptr;
    if (cudaMalloc(&buf, bytes) != cudaSuccess || cudaMalloc(&out, 64) != cudaSuccess) { fprintf(stderr, "cudaMalloc failed\n"); return 1; }
    cudaMemset(buf, 1, bytes);
    const uint32_t sectors = (uint32_t)(bytes / 32);
    printf("{\"buffer_mib\": %zu, \"per_thread\": %d}\n", mib, per_thread);
    if (chain) {
        run<1, true>(buf, sectors, per_thread, out, "dependent");
        run<2, true>(buf, sectors, per_thread, out, "dependent pairs");
        run<6, true>(buf, sectors, per_thread, out, "dependent groups of 6");
    } else {
        run<1, false>(buf, sectors, per_thread, out, "independent");
        run<4, false>(buf, sectors, per_thread, out, "independent");
        run<8, false>(buf, sectors, per_thread, out, "independent");
    }
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { fprintf(stderr, "%s\n", cudaGetErrorString(err)); return 1; }
    cudaFree(buf); cudaFree(out);
    return 0;
}
