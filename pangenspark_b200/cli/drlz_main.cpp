// drlz -- native drop-in for `spark-submit --class org.ngseq.pangenspark.DistributedRLZ <jar> ...`.
//
// Same six positional arguments, same meaning (DistributedRLZ.scala:37-43, drlz.sh:15):
//     drlz <localradix> <localref> <dataGlob> <fsURI> <outDir> <gaplessOut>
//   localradix  path where the reference would write radixSA's text suffix array.  Nothing downstream reads
//               it, so it is only written when DRLZ_EMIT_RADIX=1 (one decimal index per line, :55).
//   localref    reference sequence: exactly the sequence bytes (single line, no header).  A trailing
//               newline is dropped (the Scala side reads the file with getLines.mkString(""), :61).
//   dataGlob    glob of gapped MSA rows, one single-line file per haplotype (<sample>.<NN>, vcf2msa.sh).  With an
//               hdfs:// fsURI a glob that matches nothing locally is listed and fetched with `hdfs dfs` (Spark reads
//               the rows from HDFS, drlz.sh:15).
//   fsURI       "" / file:/// = local file system; hdfs://... = outputs are staged locally and uploaded with
//               `hdfs dfs -put -f` (the reference writes through Hadoop's FileSystem API, :251-257).
//   outDir      receives <basename>.lz: int64 LE (pos, len) records, literal = (byte, 0)   (:263-284)
//   gaplessOut  directory of part-NNNNN files, each ">basename\n<gapless>\n", plus _SUCCESS   (:66-80)
// Extra arguments (the stray "2" that drlz.sh:15 passes) are ignored.  No input file => exit code 1 (Spark throws
// "Path does not exist").
//
// The encode phase is a streaming pipeline (BASELINE.json configs[3]: 1000 haplotypes streamed from the host):
//     I/O pool   pread()s file ranges straight into pinned slot buffers (drlz_host_alloc)           |  all three run
//     GPU        drlz_submit: one batch per slot, FIFO worker inside the library, >= 2 in flight     |  concurrently,
//     I/O pool   pwrite()s <name>.lz and the gapless FASTA part files from the pinned result buffers |  DRLZ_SLOTS deep
// Environment: DRLZ_DEVICE (default 0), DRLZ_GPUS (default 1; batches go to whichever device has a free slot, each
// device builds its own index), DRLZ_BATCH_BYTES (default 128 MiB of MSA bytes per batch), DRLZ_SLOTS (default 4),
// DRLZ_IO_THREADS (default: the CPUs available, at most 32), DRLZ_GAPLESS=0 (skip the FASTA part files),
// DRLZ_MERGED_LZ / DRLZ_MERGED_FA (also write the merged chrNN.lz / chrNN.fa that index_chr.sh:16,20 builds with
// getmerge; with DRLZ_MERGED_FA also <that>.offsets, the coordinate map CHIC needs: SURVEY 8(f) rank 3),
// DRLZ_GAPTABLE=1 (<outDir>/<basename>.gaps: int64 LE (column, length) of every run of '-'),
// DRLZ_STATS=<file> (timings of the run as one JSON object).
#include <fcntl.h>
#include <glob.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <zlib.h>

#include "drlz.h"

using Clock = std::chrono::steady_clock;
static double secs_since(Clock::time_point t0) { return std::chrono::duration<double>(Clock::now() - t0).count(); }

// ---- small file helpers ---------------------------------------------------------------------------------------------
static bool is_gzip(const std::string& path) {
    if (path.size() < 3 || path.compare(path.size() - 3, 3, ".gz") != 0) return false;      // by name AND by magic
    unsigned char m[2] = {0, 0};
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    size_t n = fread(m, 1, 2, f);
    fclose(f);
    return n == 2 && m[0] == 0x1f && m[1] == 0x8b;
}

// whole file into `out`; *.gz inputs (the legacy flow globs *.NN.gz, create.sh:27 -- Spark's text reader inflates
// them transparently) are inflated, everything else is read as it is
static bool read_file(const std::string& path, std::vector<uint8_t>& out) {
    if (!is_gzip(path)) {
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) return false;
        struct stat st;
        out.resize(fstat(fileno(f), &st) == 0 && st.st_size > 0 ? (size_t)st.st_size : 0);
        size_t got = out.empty() ? 0 : fread(out.data(), 1, out.size(), f);
        bool ok = got == out.size();
        fclose(f);
        return ok;
    }
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) return false;
    gzbuffer(f, 1 << 20);
    out.resize((size_t)1 << 20);
    size_t got = 0;
    for (;;) {
        if (got == out.size()) out.resize(out.size() * 2);
        size_t want = out.size() - got;
        if (want > (1u << 30)) want = 1u << 30;
        int n = gzread(f, out.data() + got, (unsigned)want);
        if (n < 0) { gzclose(f); return false; }
        if (n == 0) break;
        got += (size_t)n;
    }
    out.resize(got);
    return gzclose(f) == Z_OK;
}

// inflated size of a gzip file from its trailer (ISIZE, modulo 2^32; exact for rows below 4 GiB)
static uint64_t gzip_isize(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return 0;
    unsigned char b[4] = {0, 0, 0, 0};
    if (fseek(f, -4, SEEK_END) == 0 && fread(b, 1, 4, f) == 4) { fclose(f); return (uint64_t)b[0] | ((uint64_t)b[1] << 8) | ((uint64_t)b[2] << 16) | ((uint64_t)b[3] << 24); }
    fclose(f);
    return 0;
}

static bool write_all(int fd, const void* p, size_t n, off_t off) {
    const char* c = (const char*)p;
    while (n) {
        ssize_t w = pwrite(fd, c, n > ((size_t)1 << 30) ? ((size_t)1 << 30) : n, off);
        if (w <= 0) return false;
        c += w; n -= (size_t)w; off += w;
    }
    return true;
}

static bool write_file_atomic(const std::string& path, const void* a, size_t na, const void* b = nullptr, size_t nb = 0,
                              const void* c = nullptr, size_t nc = 0) {
    std::string tmp = path + ".tmp";
    int fd = open(tmp.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
    if (fd < 0) return false;
    bool ok = write_all(fd, a, na, 0) && write_all(fd, b, nb, (off_t)na) && write_all(fd, c, nc, (off_t)(na + nb));
    ok = close(fd) == 0 && ok;
    if (ok) ok = rename(tmp.c_str(), path.c_str()) == 0;
    return ok;
}

static void mkdirs(const std::string& p) {
    std::string cur;
    for (size_t i = 0; i <= p.size(); ++i) {
        if (i == p.size() || p[i] == '/') { if (!cur.empty()) mkdir(cur.c_str(), 0777); }
        if (i < p.size()) cur.push_back(p[i]);
    }
}

static void remove_tree(const std::string& dir) {
    glob_t gl; memset(&gl, 0, sizeof gl);
    if (glob((dir + "/*").c_str(), 0, nullptr, &gl) == 0)
        for (size_t i = 0; i < gl.gl_pathc; ++i) unlink(gl.gl_pathv[i]);
    globfree(&gl);
    rmdir(dir.c_str());
}

static std::string base_name(const std::string& p) {
    size_t s = p.find_last_of('/');
    return s == std::string::npos ? p : p.substr(s + 1);
}

static std::string strip_scheme(const std::string& p) {
    if (p.rfind("file://", 0) == 0) return p.substr(7);
    return p;
}

// run a command without a shell (no quoting problems with paths); optionally capture its stdout
static int run_cmd(const std::vector<std::string>& argv, std::string* capture = nullptr) {
    int pfd[2] = {-1, -1};
    if (capture && pipe(pfd) != 0) return -1;
    pid_t pid = fork();
    if (pid < 0) return -1;
    if (pid == 0) {
        if (capture) { dup2(pfd[1], 1); close(pfd[0]); close(pfd[1]); }
        std::vector<char*> a;
        for (const std::string& s : argv) a.push_back(const_cast<char*>(s.c_str()));
        a.push_back(nullptr);
        execvp(a[0], a.data());
        _exit(127);
    }
    if (capture) {
        close(pfd[1]);
        char buf[4096]; ssize_t n;
        while ((n = read(pfd[0], buf, sizeof buf)) > 0) capture->append(buf, (size_t)n);
        close(pfd[0]);
    }
    int st = 0;
    if (waitpid(pid, &st, 0) < 0) return -1;
    return WIFEXITED(st) ? WEXITSTATUS(st) : -1;
}

// ---- I/O thread pool ---------------------------------------------------------------------------------------------------
class Pool {
public:
    explicit Pool(int n) { for (int i = 0; i < n; ++i) th_.emplace_back([this] { loop(); }); }
    ~Pool() { { std::lock_guard<std::mutex> lk(mu_); stop_ = true; } cv_.notify_all(); for (auto& t : th_) t.join(); }
    void submit(std::function<void()> f) { { std::lock_guard<std::mutex> lk(mu_); q_.push_back(std::move(f)); } cv_.notify_one(); }
private:
    void loop() {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                f = std::move(q_.front()); q_.pop_front();
            }
            f();
        }
    }
    std::vector<std::thread> th_; std::deque<std::function<void()>> q_; std::mutex mu_; std::condition_variable cv_; bool stop_ = false;
};

// counts outstanding tasks of one pipeline stage of one slot
struct Latch {
    std::mutex mu; std::condition_variable cv; int pending = 0;
    void add(int n) { std::lock_guard<std::mutex> lk(mu); pending += n; }
    void done() { std::lock_guard<std::mutex> lk(mu); if (--pending == 0) cv.notify_all(); }
    void wait() { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return pending == 0; }); }
};

struct Job { std::string path; std::string name; size_t index; uint64_t size; bool gz; };

// one batch in flight: pinned input rows, pinned results
struct Slot {
    uint8_t* in = nullptr; size_t in_cap = 0;
    uint8_t* gl = nullptr; size_t gl_cap = 0;
    int64_t* pairs = nullptr; size_t pairs_cap = 0;            // records
    int64_t* gaps = nullptr; size_t gaps_cap = 0;              // records
    std::vector<Job> jobs;
    std::vector<size_t> off;                                    // byte offset of each row in `in` (and of its gapless bytes in `gl`)
    std::vector<const uint8_t*> rows; std::vector<uint8_t*> gptr; std::vector<uint64_t> cols, m, np, ng;
    std::vector<uint64_t> first_nl, last_body;                  // per row: first '\n', last byte that is not a line terminator (+1)
    std::vector<std::mutex> row_mu;
    uint64_t ticket = 0, total = 0;                             // total = padded bytes of the batch in `in`
    Latch reads, writes;
    bool failed = false;
};

// grow-only pinned buffer of `want` elements of `elem` bytes (phrase and gap records are 16 bytes: two int64)
template <typename T>
static bool ensure_pinned(T*& p, size_t& cap, size_t want, size_t elem = sizeof(T)) {
    if (want <= cap) return true;
    if (p) drlz_host_free(p);
    p = (T*)drlz_host_alloc(want * elem);
    cap = p ? want : 0;
    return p != nullptr;
}

struct Stats {
    std::mutex mu;
    uint64_t bases = 0, pairs = 0, in_bytes = 0, batches = 0;
    double read_wait_s = 0, gpu_wait_s = 0, write_wait_s = 0;
};

int main(int argc, char** argv) {
    if (argc < 7) {
        fprintf(stderr, "usage: %s <localradix> <localref> <dataGlob> <fsURI> <outDir> <gaplessOut>\n", argv[0]);
        return 2;
    }
    const auto t_start = Clock::now();
    std::string localradix = argv[1], localref = argv[2], data_glob = strip_scheme(argv[3]), fs_uri = argv[4];
    std::string out_dir = argv[5], gapless_dir = argv[6];
    const bool hdfs = fs_uri.rfind("hdfs://", 0) == 0;
    auto env_u64 = [](const char* k, uint64_t dflt) { const char* v = getenv(k); return v && *v ? strtoull(v, nullptr, 10) : dflt; };
    const int dev0 = (int)env_u64("DRLZ_DEVICE", 0);
    int ngpu = (int)env_u64("DRLZ_GPUS", 1); if (ngpu < 1) ngpu = 1;
    const uint64_t batch_bytes = env_u64("DRLZ_BATCH_BYTES", 128ull << 20);
    int nslots = (int)env_u64("DRLZ_SLOTS", 4); if (nslots < 2) nslots = 2;
    const bool want_gapless = env_u64("DRLZ_GAPLESS", 1) != 0;
    // DRLZ_IO_PROBE=1: the same pipeline with the device left out -- every batch is read into its pinned slot and
    // "results" of the same sizes (the raw row as the FASTA body, cols/1024 records) are written; measures what the file
    // system of the box gives this process (the outputs are NOT a parse)
    const bool io_probe = env_u64("DRLZ_IO_PROBE", 0) != 0;
    const bool want_gaptable = env_u64("DRLZ_GAPTABLE", 0) != 0;
    unsigned hw = std::thread::hardware_concurrency(); if (!hw) hw = 8;
    int io_threads = (int)env_u64("DRLZ_IO_THREADS", std::min<uint64_t>(hw, 32)); if (io_threads < 1) io_threads = 1;

    // staging directories of the hdfs:// mode, removed on every way out
    std::string stage_root;
    struct Cleanup { std::vector<std::string> dirs; ~Cleanup() { for (auto it = dirs.rbegin(); it != dirs.rend(); ++it) remove_tree(*it); } } cleanup;
    auto staging = [&](const char* what) {
        if (stage_root.empty()) {
            const char* t = getenv("TMPDIR");
            std::string tmpl = std::string(t && *t ? t : "/tmp") + "/drlz_stage_XXXXXX";
            std::vector<char> b(tmpl.begin(), tmpl.end()); b.push_back(0);
            if (!mkdtemp(b.data())) { fprintf(stderr, "drlz: cannot create a staging directory under %s\n", tmpl.c_str()); exit(1); }
            stage_root = b.data();
            cleanup.dirs.push_back(stage_root);
        }
        std::string d = stage_root + "/" + what;
        mkdir(d.c_str(), 0777);
        cleanup.dirs.push_back(d);
        return d;
    };
    std::string local_out = hdfs ? staging("out") : strip_scheme(out_dir);
    std::string local_gap = hdfs ? staging("gapless") : strip_scheme(gapless_dir);

    // reference: all lines concatenated (DistributedRLZ.scala:61)
    // ... into pinned memory (drlz_host_alloc): the index build starts with the host->device copy of the reference, and from
    // pageable memory the driver stages it (3 ms instead of 1 for 48 MB, a quarter of a second at 3.1 Gbp)
    std::vector<uint8_t> raw;
    if (!read_file(localref, raw)) { fprintf(stderr, "drlz: cannot read %s\n", localref.c_str()); return 1; }
    struct PinnedRef {
        uint8_t* p = nullptr; size_t n = 0; bool pinned = true;
        ~PinnedRef() { if (p) { if (pinned) drlz_host_free(p); else free(p); } }
        const uint8_t* data() const { return p; }
        size_t size() const { return n; }
    } ref;
    ref.p = (uint8_t*)drlz_host_alloc(raw.size() + 1);
    if (!ref.p) { ref.pinned = false; ref.p = (uint8_t*)malloc(raw.size() + 1); }      // pageable works too (no device: the build reports it)
    if (!ref.p) { fprintf(stderr, "drlz: cannot allocate %zu bytes for the reference\n", raw.size()); return 1; }
    size_t nl = 0;
    for (const uint8_t *t = raw.data(), *e = t + raw.size(); t < e;) {           // stretches between line terminators
        const uint8_t* a = (const uint8_t*)memchr(t, '\n', (size_t)(e - t));
        const uint8_t* stop = a ? a : e;
        const uint8_t* b = (const uint8_t*)memchr(t, '\r', (size_t)(stop - t));
        if (b) stop = b;
        memcpy(ref.p + ref.n, t, (size_t)(stop - t)); ref.n += (size_t)(stop - t);
        if (stop < e) ++nl;
        t = stop + 1;
    }
    if (nl > 1) fprintf(stderr, "drlz: warning: %s has %zu line terminators; the reference expects a single-line, header-less file\n",
                        localref.c_str(), nl);
    raw.clear(); raw.shrink_to_fit();

    // inputs, sorted by name (the order `hdfs dfs -getmerge <outDir>/*.NN.lz` concatenates them, index_chr.sh:16)
    std::vector<Job> jobs;
    auto expand_local = [&](const std::string& pattern) {
        glob_t gl; memset(&gl, 0, sizeof gl);
        if (glob(pattern.c_str(), 0, nullptr, &gl) == 0)
            for (size_t i = 0; i < gl.gl_pathc; ++i) jobs.push_back(Job{gl.gl_pathv[i], base_name(gl.gl_pathv[i]), 0, 0, false});
        globfree(&gl);
    };
    expand_local(data_glob);
    if (jobs.empty() && hdfs) {
        // Spark reads the rows from HDFS: list the glob there and fetch the files into the staging area
        std::string listing, pat = argv[3];
        if (run_cmd({"hdfs", "dfs", "-ls", "-C", pat}, &listing) != 0) {
            fprintf(stderr, "drlz: `hdfs dfs -ls -C %s` failed (is the `hdfs` CLI on PATH?)\n", pat.c_str());
            return 1;
        }
        std::string in_stage = staging("in");
        size_t a = 0;
        while (a < listing.size()) {
            size_t b = listing.find('\n', a); if (b == std::string::npos) b = listing.size();
            std::string remote = listing.substr(a, b - a); a = b + 1;
            while (!remote.empty() && (remote.back() == '\r' || remote.back() == ' ')) remote.pop_back();
            if (remote.empty()) continue;
            std::string local = in_stage + "/" + base_name(remote);
            if (run_cmd({"hdfs", "dfs", "-get", "-f", remote, local}) != 0) { fprintf(stderr, "drlz: cannot fetch %s\n", remote.c_str()); return 1; }
        }
        expand_local(in_stage + "/*");
    }
    if (jobs.empty()) {
        fprintf(stderr, "drlz: no input matches %s (Path does not exist)\n", argv[3]);
        return 1;
    }
    std::sort(jobs.begin(), jobs.end(), [](const Job& a, const Job& b) { return a.name < b.name; });
    for (size_t i = 0; i < jobs.size(); ++i) {
        jobs[i].index = i;
        jobs[i].gz = is_gzip(jobs[i].path);
        struct stat st;
        jobs[i].size = stat(jobs[i].path.c_str(), &st) == 0 ? (uint64_t)st.st_size : 0;
        if (jobs[i].gz) jobs[i].size = gzip_isize(jobs[i].path);       // inflated size (the slot layout needs it)
    }
    mkdirs(local_out); if (want_gapless) mkdirs(local_gap);
    // batches: consecutive files of the sorted list, at most batch_bytes each (a bigger row is a batch of its own);
    // whichever device has a free slot takes the next one
    std::vector<std::pair<size_t, size_t>> batches;
    uint64_t slot_bytes = 0, slot_rows = 0;
    for (size_t i = 0; i < jobs.size();) {
        size_t e = i; uint64_t acc = 0, padded = 0;
        while (e < jobs.size() && (e == i || acc + jobs[e].size <= batch_bytes)) { acc += jobs[e].size; padded += ((jobs[e].size + 15) & ~15ull) + 16; ++e; }
        batches.push_back({i, e});
        slot_bytes = std::max<uint64_t>(slot_bytes, padded + 16);
        slot_rows = std::max<uint64_t>(slot_rows, e - i);
        i = e;
    }
    if ((int)(batches.size() / ngpu) + 1 < nslots) nslots = std::max<int>(2, (int)(batches.size() / ngpu) + 1);

    printf("Create suffix\n");
    const auto t_index = Clock::now();
    std::vector<drlz_ctx*> ctxs(ngpu, nullptr);
    std::vector<std::vector<Slot>> all_slots(ngpu);
    std::atomic<int> failed{0};
    {
        std::vector<std::thread> th;
        for (int g = 0; g < ngpu; ++g) th.emplace_back([&, g] {
            if (drlz_create(&ctxs[g], dev0 + g) != DRLZ_OK) { fprintf(stderr, "drlz: no CUDA device %d (there is no CPU fallback)\n", dev0 + g); failed = 1; return; }
            if (const char* opts = getenv("DRLZ_OPTIONS")) {           // "key=value,key=value": library options (include/drlz.h), for measurements
                std::string o(opts);
                for (size_t a = 0; a < o.size();) {
                    size_t b = o.find(',', a); if (b == std::string::npos) b = o.size();
                    const size_t eq = o.find('=', a);
                    if (eq != std::string::npos && eq < b) drlz_set_option(ctxs[g], o.substr(a, eq - a).c_str(), atoll(o.substr(eq + 1, b - eq - 1).c_str()));
                    a = b + 1;
                }
            }
            // the pinned slot buffers are allocated while the index is being built (pinning gigabytes takes as long)
            std::thread alloc([&, g] {
                std::vector<Slot> slots(nslots);
                for (auto& sl : slots)
                    if (!ensure_pinned(sl.in, sl.in_cap, slot_bytes) || (want_gapless && !ensure_pinned(sl.gl, sl.gl_cap, slot_bytes)) ||
                        !ensure_pinned(sl.pairs, sl.pairs_cap, slot_bytes / 64 + 1024 * slot_rows, 16) ||
                        (want_gaptable && !ensure_pinned(sl.gaps, sl.gaps_cap, slot_bytes / 64 + 1024 * slot_rows, 16))) {
                        fprintf(stderr, "drlz: cannot allocate pinned host memory (%llu bytes per slot)\n", (unsigned long long)slot_bytes); failed = 1; break;
                    }
                all_slots[g].swap(slots);
            });
            int rc = drlz_build_index(ctxs[g], ref.data(), ref.size());
            if (rc != DRLZ_OK) { fprintf(stderr, "drlz: index build failed on device %d: %s\n", dev0 + g, drlz_last_error(ctxs[g])); failed = 1; }
            alloc.join();
        });
        for (auto& t : th) t.join();
    }
    auto free_slots_all = [&] { for (auto& v : all_slots) for (auto& sl : v) { drlz_host_free(sl.in); drlz_host_free(sl.gl); drlz_host_free(sl.pairs); drlz_host_free(sl.gaps); } };
    if (failed) { free_slots_all(); for (auto c : ctxs) if (c) drlz_destroy(c); return 1; }
    const double index_s = secs_since(t_index);
    double idx_ms[5] = {0, 0, 0, 0, 0};
    {
        uint64_t info[4];
        drlz_index_timings(ctxs[0], idx_ms); drlz_index_info(ctxs[0], info);
        printf("Load suffix\nbroadcasting\nindex: n=%llu k=%llu sigma=%llu bits=%d rounds=%llu build=%.1f ms on %d device(s)\n",
               (unsigned long long)info[0], (unsigned long long)(info[1] & 0xFF), (unsigned long long)(info[1] >> 16),
               1 << ((info[1] >> 8) & 0xFF), (unsigned long long)info[2], idx_ms[0], ngpu);
    }
    if (getenv("DRLZ_EMIT_RADIX") && atoi(getenv("DRLZ_EMIT_RADIX"))) {
        std::vector<uint32_t> sa(ref.size());
        if (drlz_index_get(ctxs[0], 0, sa.data(), sa.size() * 4) == DRLZ_OK) {
            FILE* f = fopen(localradix.c_str(), "w");
            if (f) { for (uint32_t v : sa) fprintf(f, "%u\n", v); fclose(f); }
        }
    }

    printf("started encoding\n");
    const auto t_encode = Clock::now();
    Stats stats;
    std::mutex job_mu;
    size_t next_job = 0;
    {
        Pool pool(io_threads);
        const size_t kRange = (size_t)env_u64("DRLZ_IO_RANGE", 4u << 20);   // bytes per read / write task
        auto device_pipeline = [&](int g) {
            drlz_ctx* c = ctxs[g];
            std::vector<Slot>& slots = all_slots[g];
            // a slot goes round: free -> reading (pool fills it) -> in_gpu (submitted) -> writing (pool writes its results) -> free;
            // one thread per arrow, so the reads of batch k+1, the parse of batch k and the writes of batch k-1 overlap
            std::deque<Slot*> free_slots, reading, in_gpu, writing;
            std::mutex smu; std::condition_variable scv;
            for (auto& s : slots) free_slots.push_back(&s);
            bool issuer_done = false, submitter_done = false, collector_done = false;

            // collector: waits for the batches in submission order, hands their results to the I/O pool
            std::thread collector([&] {
                for (;;) {
                    Slot* s;
                    {
                        std::unique_lock<std::mutex> lk(smu);
                        scv.wait(lk, [&] { return submitter_done || !in_gpu.empty(); });
                        if (in_gpu.empty()) { collector_done = true; scv.notify_all(); return; }
                        s = in_gpu.front(); in_gpu.pop_front();
                    }
                    const auto tw = Clock::now();
                    int rc = io_probe ? DRLZ_OK : drlz_wait(c, s->ticket);
                    const size_t H = s->jobs.size();
                    if (rc == DRLZ_E_NOSPACE) {               // more phrases than the slot's record buffer holds: grow and redo
                        uint64_t need = 16;
                        for (size_t h = 0; h < H; ++h) need += s->np[h];
                        if (!ensure_pinned(s->pairs, s->pairs_cap, need + need / 8, 16)) rc = DRLZ_E_NOMEM;
                        else rc = drlz_parse_batch(c, s->rows.data(), s->cols.data(), (uint32_t)H, 1, want_gapless ? s->gptr.data() : nullptr,
                                                   s->m.data(), s->pairs, s->pairs_cap, s->np.data());
                    }
                    { std::lock_guard<std::mutex> lk(stats.mu); stats.gpu_wait_s += secs_since(tw); }
                    if (rc != DRLZ_OK) { fprintf(stderr, "drlz: parse failed: %s\n", drlz_last_error(c)); failed = 1; s->failed = true; }
                    if (!s->failed) {
                        uint64_t poff = 0;
                        for (size_t h = 0; h < H; ++h) {
                            const int64_t* rec = s->pairs + 2 * poff;
                            const uint64_t nrec = s->np[h];
                            poff += nrec;
                            s->writes.add(1);
                            pool.submit([&, s, h, rec, nrec] {
                                const Job& j = s->jobs[h];
                                std::string lz = local_out + "/" + j.name + ".lz";
                                if (!write_file_atomic(lz, rec, (size_t)nrec * 16)) { fprintf(stderr, "drlz: cannot write %s\n", lz.c_str()); failed = 1; }
                                s->writes.done();
                            });
                        }
                    }
                    // (the FASTA part files are issued below, after the .lz tasks, so that small files finish first)
                    if (!s->failed && want_gapless) {
                        for (size_t h = 0; h < H; ++h) {
                            const Job& j = s->jobs[h];
                            char part[64]; snprintf(part, sizeof part, "/part-%05zu", j.index);
                            const std::string path = local_gap + part, tmp = path + ".tmp";
                            const std::string hdr = ">" + j.name + "\n";
                            const uint64_t body = s->m[h];
                            int fd = open(tmp.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
                            if (fd < 0 || !write_all(fd, hdr.data(), hdr.size(), 0) || !write_all(fd, "\n", 1, (off_t)(hdr.size() + body))) {
                                fprintf(stderr, "drlz: cannot write %s\n", path.c_str()); failed = 1;
                                if (fd >= 0) close(fd);
                                continue;
                            }
                            const size_t nr = body ? (size_t)((body + kRange - 1) / kRange) : 0;
                            auto left = std::make_shared<std::atomic<size_t>>(nr);
                            auto finish = [fd, tmp, path, &failed] {
                                if (close(fd) != 0 || rename(tmp.c_str(), path.c_str()) != 0) { fprintf(stderr, "drlz: cannot write %s\n", path.c_str()); failed = 1; }
                            };
                            if (nr == 0) { finish(); continue; }
                            const uint8_t* src = s->gptr[h];
                            const size_t hl = hdr.size();
                            for (size_t r = 0; r < nr; ++r) {
                                s->writes.add(1);
                                pool.submit([&, s, fd, src, r, body, hl, left, finish, path] {
                                    const uint64_t a = (uint64_t)r * kRange, b = std::min<uint64_t>(body, a + kRange);
                                    if (!write_all(fd, src + a, (size_t)(b - a), (off_t)(hl + a))) { fprintf(stderr, "drlz: cannot write %s\n", path.c_str()); failed = 1; }
                                    if (left->fetch_sub(1) == 1) finish();
                                    s->writes.done();
                                });
                            }
                        }
                    }
                    if (!s->failed && want_gaptable) {
                        uint64_t goff = 0;
                        for (size_t h = 0; h < H; ++h) {
                            std::string gp = local_out + "/" + s->jobs[h].name + ".gaps";
                            if (!write_file_atomic(gp, s->gaps + 2 * goff, (size_t)s->ng[h] * 16)) { fprintf(stderr, "drlz: cannot write %s\n", gp.c_str()); failed = 1; }
                            goff += s->ng[h];
                        }
                    }
                    {
                        std::lock_guard<std::mutex> lk(stats.mu);
                        if (!s->failed) for (size_t h = 0; h < H; ++h) { stats.bases += s->m[h]; stats.pairs += s->np[h]; stats.in_bytes += s->cols[h]; }
                        ++stats.batches;
                    }
                    { std::lock_guard<std::mutex> lk(smu); writing.push_back(s); }
                    scv.notify_all();
                }
            });

            // releaser: a slot is free again when the pool has written everything that points into it
            std::thread releaser([&] {
                for (;;) {
                    Slot* s;
                    {
                        std::unique_lock<std::mutex> lk(smu);
                        scv.wait(lk, [&] { return collector_done || !writing.empty(); });
                        if (writing.empty()) return;
                        s = writing.front(); writing.pop_front();
                    }
                    const auto tq = Clock::now();
                    s->writes.wait();
                    { std::lock_guard<std::mutex> lk(stats.mu); stats.write_wait_s += secs_since(tq); }
                    { std::lock_guard<std::mutex> lk(smu); free_slots.push_back(s); }
                    scv.notify_all();
                }
            });

            // submitter: waits for the reads of the batches in order, checks the rows, hands the slot to the library
            std::thread submitter([&] {
                for (;;) {
                    Slot* s;
                    {
                        std::unique_lock<std::mutex> lk(smu);
                        scv.wait(lk, [&] { return issuer_done || !reading.empty(); });
                        if (reading.empty()) { submitter_done = true; scv.notify_all(); return; }
                        s = reading.front(); reading.pop_front();
                    }
                    const size_t H = s->jobs.size();
                    const auto tr = Clock::now();
                    s->reads.wait();
                    { std::lock_guard<std::mutex> lk(stats.mu); stats.read_wait_s += secs_since(tr); }
                    for (size_t h = 0; h < H && !s->failed; ++h) {
                        // one record per line (Spark read.text); only single-line files are supported
                        const uint64_t first = std::min<uint64_t>(s->first_nl[h], s->jobs[h].size);
                        if (s->last_body[h] > first) {
                            fprintf(stderr, "drlz: %s has more than one line; multi-line rows are unsupported (each line would overwrite the same .lz in the reference)\n", s->jobs[h].path.c_str());
                            failed = 1; s->failed = true;
                        }
                        s->cols[h] = first;
                    }
                    int rc = DRLZ_OK;
                    if (s->failed || failed) rc = DRLZ_E_ARG;          // drain: nothing is submitted after a failure
                    else if (io_probe) {
                        for (size_t h = 0; h < H; ++h) {
                            s->m[h] = s->cols[h]; s->np[h] = s->cols[h] / 1024; s->ng[h] = 0;
                            if (want_gapless) s->gptr[h] = const_cast<uint8_t*>(s->rows[h]);
                        }
                        memset(s->pairs, 0, std::min<size_t>(s->pairs_cap, s->total / 64) * 16);
                    } else {
                        rc = want_gaptable
                            ? drlz_submit_gaps(c, s->rows.data(), s->cols.data(), (uint32_t)H, 1, want_gapless ? s->gptr.data() : nullptr, s->m.data(),
                                               s->pairs, s->pairs_cap, s->np.data(), s->gaps, s->gaps_cap, s->ng.data(), &s->ticket)
                            : drlz_submit(c, s->rows.data(), s->cols.data(), (uint32_t)H, 1, want_gapless ? s->gptr.data() : nullptr, s->m.data(),
                                          s->pairs, s->pairs_cap, s->np.data(), &s->ticket);
                        if (rc != DRLZ_OK) { fprintf(stderr, "drlz: submit failed: %s\n", drlz_last_error(c)); failed = 1; }
                    }
                    {
                        std::lock_guard<std::mutex> lk(smu);
                        if (rc == DRLZ_OK) in_gpu.push_back(s); else free_slots.push_back(s);
                    }
                    scv.notify_all();
                }
            });

            // issuer: takes the next batch of files and a free slot, hands the reads to the pool
            for (;;) {
                if (failed) break;
                Slot* s;
                {
                    std::unique_lock<std::mutex> lk(smu);
                    scv.wait(lk, [&] { return !free_slots.empty(); });
                    s = free_slots.front(); free_slots.pop_front();
                }
                s->jobs.clear(); s->failed = false;
                {
                    std::lock_guard<std::mutex> lk(job_mu);
                    if (next_job < batches.size()) {
                        for (size_t j = batches[next_job].first; j < batches[next_job].second; ++j) s->jobs.push_back(jobs[j]);
                        ++next_job;
                    }
                }
                if (s->jobs.empty()) { std::lock_guard<std::mutex> lk(smu); free_slots.push_back(s); break; }
                const size_t H = s->jobs.size();
                s->off.assign(H, 0); s->rows.assign(H, nullptr); s->gptr.assign(H, nullptr);
                s->cols.assign(H, 0); s->m.assign(H, 0); s->np.assign(H, 0); s->ng.assign(H, 0);
                s->first_nl.assign(H, UINT64_MAX); s->last_body.assign(H, 0);
                { std::vector<std::mutex> mus(H); s->row_mu.swap(mus); }
                uint64_t total = 0;
                for (size_t h = 0; h < H; ++h) { s->off[h] = total; total += ((s->jobs[h].size + 15) & ~15ull) + 16; }
                if (!ensure_pinned(s->in, s->in_cap, total + 16) || (want_gapless && !ensure_pinned(s->gl, s->gl_cap, total + 16)) ||
                    !ensure_pinned(s->pairs, s->pairs_cap, total / 64 + 1024 * H, 16) ||
                    (want_gaptable && !ensure_pinned(s->gaps, s->gaps_cap, total / 64 + 1024 * H, 16))) {
                    fprintf(stderr, "drlz: cannot allocate pinned host memory for a batch of %llu bytes\n", (unsigned long long)total); failed = 1; break;
                }
                s->total = total;
                for (size_t h = 0; h < H; ++h) {
                    const Job& j = s->jobs[h];
                    uint8_t* dst = s->in + s->off[h];
                    s->rows[h] = dst; s->gptr[h] = want_gapless ? s->gl + s->off[h] : nullptr;
                    // every task notes the first '\n' and the end of the last byte that is not a line terminator in its
                    // range: a row is single-line iff nothing but terminators follows its first '\n' 
                    auto scan = [s, h](const uint8_t* p, uint64_t a, uint64_t b) {
                        const void* q = memchr(p + a, '\n', (size_t)(b - a));
                        uint64_t e = b;
                        while (e > a && (p[e - 1] == '\n' || p[e - 1] == '\r')) --e;
                        std::lock_guard<std::mutex> lk(s->row_mu[h]);
                        if (q) s->first_nl[h] = std::min<uint64_t>(s->first_nl[h], (uint64_t)((const uint8_t*)q - p));
                        if (e > a) s->last_body[h] = std::max<uint64_t>(s->last_body[h], e);
                    };
                    if (j.gz) {
                        s->reads.add(1);
                        pool.submit([&, s, h, dst, scan] {
                            const Job& jj = s->jobs[h];
                            gzFile f = gzopen(jj.path.c_str(), "rb");
                            uint64_t got = 0; bool ok = f != nullptr;
                            if (ok) {
                                gzbuffer(f, 1 << 20);
                                while (got < jj.size) {
                                    int n = gzread(f, dst + got, (unsigned)std::min<uint64_t>(jj.size - got, 1u << 30));
                                    if (n <= 0) break;
                                    got += (uint64_t)n;
                                }
                                uint8_t extra;
                                ok = got == jj.size && gzread(f, &extra, 1) == 0;
                                gzclose(f);
                            }
                            if (!ok) { fprintf(stderr, "drlz: cannot inflate %s (rows of 4 GiB and more are not supported as .gz)\n", jj.path.c_str()); failed = 1; s->failed = true; }
                            else scan(dst, 0, got);
                            s->reads.done();
                        });
                        continue;
                    }
                    int fd = open(j.path.c_str(), O_RDONLY);
                    if (fd < 0) { fprintf(stderr, "drlz: cannot read %s\n", j.path.c_str()); failed = 1; s->failed = true; continue; }
                    const size_t nr = j.size ? (size_t)((j.size + kRange - 1) / kRange) : 0;
                    if (nr == 0) { close(fd); continue; }
                    auto left = std::make_shared<std::atomic<size_t>>(nr);
                    for (size_t r = 0; r < nr; ++r) {
                        s->reads.add(1);
                        pool.submit([&, s, h, fd, dst, r, left, scan] {
                            const Job& jj = s->jobs[h];
                            const uint64_t a = (uint64_t)r * kRange, b = std::min<uint64_t>(jj.size, a + kRange);
                            uint64_t got = a;
                            while (got < b) {
                                ssize_t n = pread(fd, dst + got, (size_t)(b - got), (off_t)got);
                                if (n <= 0) break;
                                got += (uint64_t)n;
                            }
                            if (got != b) { fprintf(stderr, "drlz: cannot read %s\n", jj.path.c_str()); failed = 1; s->failed = true; }
                            else scan(dst, a, b);
                            if (left->fetch_sub(1) == 1) close(fd);
                            s->reads.done();
                        });
                    }
                }
                { std::lock_guard<std::mutex> lk(smu); reading.push_back(s); }
                scv.notify_all();
            }
            { std::lock_guard<std::mutex> lk(smu); issuer_done = true; }
            scv.notify_all();
            submitter.join(); collector.join(); releaser.join();
        };
        std::vector<std::thread> th;
        for (int g = 0; g < ngpu; ++g) th.emplace_back(device_pipeline, g);
        for (auto& t : th) t.join();
    }
    const double encode_s = secs_since(t_encode);
    free_slots_all();
    for (auto c : ctxs) drlz_destroy(c);
    if (failed) return 1;
    if (want_gapless) write_file_atomic(local_gap + "/_SUCCESS", "", 0);

    // the two `hdfs dfs -getmerge` steps of components/index/index_chr.sh:16,20 done here --
    // concatenation in sorted-basename order -- when DRLZ_MERGED_LZ / DRLZ_MERGED_FA name the target files.
    // rank 3: with the merged FASTA comes <target>.offsets, one line per haplotype in file order:
    //     name <TAB> offset of '>' <TAB> offset of the first base <TAB> number of bases
    // Phrase positions are coordinates of the reference sequence; CHIC indexes the merged FASTA, in which the reference
    // is the first record (README.md:90: "the first MSA file alphabetically should be the reference genome"), so a
    // phrase (pos, len) is the text [offsets[0].first_base + pos, + len) of chrNN.fa.
    const auto t_merge = Clock::now();
    for (int which = 0; which < 2; ++which) {
        const char* target = getenv(which == 0 ? "DRLZ_MERGED_LZ" : "DRLZ_MERGED_FA");
        if (!target || !*target) continue;
        if (which == 1 && !want_gapless) { fprintf(stderr, "drlz: DRLZ_MERGED_FA needs the gapless output (DRLZ_GAPLESS=0 is set)\n"); return 1; }
        std::string tmp = std::string(target) + ".tmp";
        FILE* out = fopen(tmp.c_str(), "wb");
        if (!out) { fprintf(stderr, "drlz: cannot write %s\n", target); return 1; }
        FILE* map = nullptr;
        if (which == 1) { map = fopen((std::string(target) + ".offsets").c_str(), "w"); if (!map) { fprintf(stderr, "drlz: cannot write %s.offsets\n", target); fclose(out); return 1; } }
        std::vector<uint8_t> buf((size_t)8 << 20);
        uint64_t at = 0;
        for (size_t j = 0; j < jobs.size(); ++j) {
            char part[64]; snprintf(part, sizeof part, "/part-%05zu", jobs[j].index);
            std::string src = which == 0 ? local_out + "/" + jobs[j].name + ".lz" : local_gap + part;
            FILE* in = fopen(src.c_str(), "rb");          // plain bytes: a .lz may well begin with 0x1f 0x8b
            bool ok = in != nullptr;
            uint64_t len = 0;
            while (ok) {
                size_t n = fread(buf.data(), 1, buf.size(), in);
                if (n == 0) { ok = !ferror(in); break; }
                ok = fwrite(buf.data(), 1, n, out) == n;
                len += n;
            }
            if (in) fclose(in);
            if (!ok) { fprintf(stderr, "drlz: merge failed at %s\n", src.c_str()); fclose(out); if (map) fclose(map); return 1; }
            if (map) {
                const uint64_t hl = jobs[j].name.size() + 2;        // '>' name '\n'
                fprintf(map, "%s\t%llu\t%llu\t%llu\n", jobs[j].name.c_str(), (unsigned long long)at, (unsigned long long)(at + hl),
                        (unsigned long long)(len >= hl + 1 ? len - hl - 1 : 0));
            }
            at += len;
        }
        if (map) fclose(map);
        if (fclose(out) != 0 || rename(tmp.c_str(), target) != 0) { fprintf(stderr, "drlz: cannot write %s\n", target); return 1; }
    }
    const double merge_s = secs_since(t_merge);
    printf("encoded %zu haplotypes, %llu bases, %llu phrases in %.3f s (%.2f Gbp/s; read wait %.3f s, GPU wait %.3f s, write wait %.3f s)\n",
           jobs.size(), (unsigned long long)stats.bases, (unsigned long long)stats.pairs, encode_s, stats.bases / encode_s / 1e9,
           stats.read_wait_s, stats.gpu_wait_s, stats.write_wait_s);
    double upload_s = 0;
    if (hdfs) {
        const auto tu = Clock::now();
        bool ok = run_cmd({"hdfs", "dfs", "-mkdir", "-p", out_dir, gapless_dir}) == 0;
        for (int which = 0; which < 2 && ok; ++which) {
            if (which == 1 && !want_gapless) break;
            std::vector<std::string> cmd = {"hdfs", "dfs", "-put", "-f"};
            glob_t gl; memset(&gl, 0, sizeof gl);
            if (glob(((which == 0 ? local_out : local_gap) + "/*").c_str(), 0, nullptr, &gl) == 0)
                for (size_t i = 0; i < gl.gl_pathc; ++i) cmd.push_back(gl.gl_pathv[i]);
            globfree(&gl);
            if (cmd.size() == 4) continue;
            cmd.push_back((which == 0 ? out_dir : gapless_dir) + "/");
            ok = run_cmd(cmd) == 0;
        }
        upload_s = secs_since(tu);
        if (!ok) {
            fprintf(stderr, "drlz: upload to %s failed (is the `hdfs` CLI on PATH?); outputs kept in %s\n", fs_uri.c_str(), stage_root.c_str());
            cleanup.dirs.clear();
            return 1;
        }
    }
    if (const char* sp = getenv("DRLZ_STATS")) {
        FILE* f = fopen(sp, "w");
        if (f) {
            fprintf(f, "{\"haplotypes\": %zu, \"bases\": %llu, \"phrases\": %llu, \"input_bytes\": %llu, \"batches\": %llu, \"gpus\": %d, \"slots\": %d, "
                       "\"io_threads\": %d, \"batch_bytes\": %llu, \"gapless\": %d, \"index_s\": %.6f, \"index_device_ms\": %.3f, \"encode_s\": %.6f, "
                       "\"merge_s\": %.6f, \"upload_s\": %.6f, \"total_s\": %.6f, \"read_wait_s\": %.6f, \"gpu_wait_s\": %.6f, \"write_wait_s\": %.6f, "
                       "\"encode_gbp_per_s\": %.4f, \"encode_input_gb_per_s\": %.4f}\n",
                    jobs.size(), (unsigned long long)stats.bases, (unsigned long long)stats.pairs, (unsigned long long)stats.in_bytes,
                    (unsigned long long)stats.batches, ngpu, nslots, io_threads, (unsigned long long)batch_bytes, want_gapless ? 1 : 0, index_s,
                    idx_ms[0], encode_s, merge_s, upload_s, secs_since(t_start), stats.read_wait_s, stats.gpu_wait_s, stats.write_wait_s,
                    stats.bases / encode_s / 1e9, stats.in_bytes / encode_s / 1e9);
            fclose(f);
        }
    }
    return 0;
}
