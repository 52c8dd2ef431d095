// h2d_ceiling.cu -- what the box gives: pinned host -> device copy bandwidth of N GPUs, one at a time and all at once.
//
// the end-to-end curve of bench.py (53.7 -> 174 Gbp/s from 1 to 8 GPUs) follows the host->device
// copies; this tool measures the ceiling those copies can reach, with nothing of the library involved:
//     one thread per GPU, each with its own pinned buffer and stream, cudaMemcpyAsync of `bytes` x reps,
//     wall clock around a barrier (all GPUs start together), per-GPU and aggregate GB/s.
// Modes: "default"  cudaHostAlloc from the calling thread as it is
//        "affine"   the thread first binds itself to the CPUs next to its GPU (sysfs local_cpulist), so that the
//                   first-touch policy places the pinned pages on that GPU's NUMA node
//        "wc"       affine + cudaHostAllocWriteCombined
// build: nvcc -O2 -o tools/h2d_ceiling tools/h2d_ceiling.cu -lpthread      run: tools/h2d_ceiling [ngpu] [MiB] [reps]
#include <cuda_runtime.h>
#include <pthread.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

static std::string sysfs_read(const std::string& p) {
    FILE* f = fopen(p.c_str(), "r");
    if (!f) return "";
    char buf[4096]; size_t n = fread(buf, 1, sizeof buf - 1, f); fclose(f);
    buf[n] = 0;
    while (n && (buf[n - 1] == '\n' || buf[n - 1] == ' ')) buf[--n] = 0;
    return buf;
}

static std::string pci_dir(int dev) {
    char id[64] = {0};
    cudaDeviceGetPCIBusId(id, sizeof id, dev);
    for (char* c = id; *c; ++c) if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');
    return std::string("/sys/bus/pci/devices/") + id;
}

static bool bind_cpulist(const std::string& list) {
    cpu_set_t set; CPU_ZERO(&set);
    int n = 0;
    const char* s = list.c_str();
    while (*s) {
        char* e; long a = strtol(s, &e, 10), b = a;
        if (e == s) break;
        if (*e == '-') { s = e + 1; b = strtol(s, &e, 10); }
        for (long c = a; c <= b; ++c) { CPU_SET((int)c, &set); ++n; }
        s = *e == ',' ? e + 1 : e;
    }
    return n > 0 && sched_setaffinity(0, sizeof set, &set) == 0;
}

struct Barrier {
    std::atomic<int> count{0}; std::atomic<int> gen{0}; int n;
    explicit Barrier(int n_) : n(n_) {}
    void wait() {
        int g = gen.load();
        if (count.fetch_add(1) + 1 == n) { count = 0; gen.fetch_add(1); }
        else while (gen.load() == g) std::this_thread::yield();
    }
};

static void run(const std::vector<int>& devs, size_t bytes, int reps, int mode, std::vector<double>& gbs, double* agg) {
    int n = (int)devs.size();
    Barrier bar(n);
    std::vector<std::thread> th;
    std::vector<double> t0(n), t1(n);
    for (int i = 0; i < n; ++i) th.emplace_back([&, i] {
        int dev = devs[i];
        cudaSetDevice(dev);
        if (mode >= 1) bind_cpulist(sysfs_read(pci_dir(dev) + "/local_cpulist"));
        void *h = nullptr, *d = nullptr;
        cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
        if (cudaHostAlloc(&h, bytes, mode == 2 ? cudaHostAllocWriteCombined : cudaHostAllocDefault) != cudaSuccess) { fprintf(stderr, "hostalloc failed\n"); exit(1); }
        memset(h, 0x41, bytes);
        cudaMalloc(&d, bytes);
        cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s); cudaStreamSynchronize(s);     // warm-up
        bar.wait();
        auto a = std::chrono::steady_clock::now();
        for (int r = 0; r < reps; ++r) cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s);
        cudaStreamSynchronize(s);
        auto b = std::chrono::steady_clock::now();
        t0[i] = std::chrono::duration<double>(a.time_since_epoch()).count();
        t1[i] = std::chrono::duration<double>(b.time_since_epoch()).count();
        bar.wait();
        cudaFree(d); cudaFreeHost(h); cudaStreamDestroy(s);
    });
    for (auto& t : th) t.join();
    double lo = t0[0], hi = t1[0];
    gbs.assign(n, 0);
    for (int i = 0; i < n; ++i) {
        gbs[i] = (double)bytes * reps / (t1[i] - t0[i]) / 1e9;
        if (t0[i] < lo) lo = t0[i];
        if (t1[i] > hi) hi = t1[i];
    }
    *agg = (double)bytes * reps * n / (hi - lo) / 1e9;
}

int main(int argc, char** argv) {
    int ndev = 0; cudaGetDeviceCount(&ndev);
    int want = argc > 1 ? atoi(argv[1]) : ndev;
    if (want > ndev) want = ndev;
    size_t bytes = (size_t)(argc > 2 ? atoi(argv[2]) : 1024) << 20;
    int reps = argc > 3 ? atoi(argv[3]) : 4;
    printf("{\"gpus\": %d, \"bytes\": %zu, \"reps\": %d, \"devices\": [", want, bytes, reps);
    for (int d = 0; d < want; ++d)
        printf("%s{\"dev\": %d, \"numa_node\": \"%s\", \"local_cpulist\": \"%s\"}", d ? ", " : "", d,
               sysfs_read(pci_dir(d) + "/numa_node").c_str(), sysfs_read(pci_dir(d) + "/local_cpulist").c_str());
    printf("], \"online_nodes\": \"%s\", \"results\": [", sysfs_read("/sys/devices/system/node/online").c_str());
    const char* names[3] = {"default", "affine", "wc"};
    bool first = true;
    for (int mode = 0; mode < 3; ++mode) {
        std::vector<double> g; double agg;
        // one GPU at a time
        std::vector<double> solo;
        for (int d = 0; d < want; ++d) { run({d}, bytes, reps, mode, g, &agg); solo.push_back(g[0]); }
        printf("%s{\"mode\": \"%s\", \"solo_gbs\": [", first ? "" : ", ", names[mode]); first = false;
        for (int d = 0; d < want; ++d) printf("%s%.1f", d ? ", " : "", solo[d]);
        printf("]");
        for (int n = 2; n <= want; n *= 2) {
            std::vector<int> devs; for (int d = 0; d < n; ++d) devs.push_back(d);
            run(devs, bytes, reps, mode, g, &agg);
            printf(", \"together_%d\": {\"aggregate_gbs\": %.1f, \"per_gpu\": [", n, agg);
            for (int d = 0; d < n; ++d) printf("%s%.1f", d ? ", " : "", g[d]);
            printf("]}");
        }
        printf("}");
        fflush(stdout);
    }
    printf("]}\n");
    return 0;
}
