// drlz -- native drop-in for `spark-submit --class org.ngseq.pangenspark.DistributedRLZ <jar> ...`.
//
// Same six positional arguments, same meaning (DistributedRLZ.scala:37-43, drlz.sh:15):
//     drlz <localradix> <localref> <dataGlob> <fsURI> <outDir> <gaplessOut>
//   localradix  path where the reference would write radixSA's text suffix array.  Nothing downstream reads
//               it, so it is only written when DRLZ_EMIT_RADIX=1 (one decimal index per line, :55).
//   localref    reference sequence: exactly the sequence bytes (single line, no header).  A trailing
//               newline is dropped (the Scala side reads the file with getLines.mkString(""), :61).
//   dataGlob    glob of gapped MSA rows, one single-line file per haplotype (<sample>.<NN>, vcf2msa.sh).
//   fsURI       "" / file:/// = local file system; hdfs://... = outputs are staged locally and uploaded with
//               `hdfs dfs -put -f` (the reference writes through Hadoop's FileSystem API, :251-257).
//   outDir      receives <basename>.lz: int64 LE (pos, len) records, literal = (byte, 0)   (:263-284)
//   gaplessOut  directory of part-NNNNN files, each ">basename\n<gapless>\n", plus _SUCCESS   (:66-80)
// Extra arguments (the stray "2" that drlz.sh:15 passes) are ignored.
// Environment: DRLZ_DEVICE (default 0), DRLZ_GPUS (default 1; haplotypes are dealt round-robin to the devices,
// each device builds its own index), DRLZ_BATCH_BYTES (default 1 GiB of MSA bytes per batch), DRLZ_MERGED_LZ /
// DRLZ_MERGED_FA (also write the merged chrNN.lz / chrNN.fa that index_chr.sh:16,20 builds with getmerge).
#include <glob.h>
#include <libgen.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <zlib.h>

#include "drlz.h"

// Reads a file whole; gzip files (the legacy flow globs *.NN.gz, create.sh:27 -- Spark's text reader inflates them
// transparently) are inflated, anything else is read as it is (zlib's gzread passes plain bytes through).
static bool read_file(const std::string& path, std::vector<uint8_t>& out) {
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) return false;
    gzbuffer(f, 1 << 20);
    struct stat st;
    size_t cap = stat(path.c_str(), &st) == 0 && st.st_size > 0 ? (size_t)st.st_size + 1 : (size_t)1 << 16;
    out.resize(cap);
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

static bool write_file_atomic(const std::string& path, const void* a, size_t na, const void* b = nullptr, size_t nb = 0,
                              const void* c = nullptr, size_t nc = 0) {
    std::string tmp = path + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return false;
    bool ok = (na == 0 || fwrite(a, 1, na, f) == na) && (nb == 0 || fwrite(b, 1, nb, f) == nb) &&
              (nc == 0 || fwrite(c, 1, nc, f) == nc);
    ok = fclose(f) == 0 && ok;
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

static std::string base_name(const std::string& p) {
    size_t s = p.find_last_of('/');
    return s == std::string::npos ? p : p.substr(s + 1);
}

static std::string strip_scheme(const std::string& p) {
    if (p.rfind("file://", 0) == 0) return p.substr(7);
    return p;
}

struct Job { std::string path; std::string name; size_t index; };

int main(int argc, char** argv) {
    if (argc < 7) {
        fprintf(stderr, "usage: %s <localradix> <localref> <dataGlob> <fsURI> <outDir> <gaplessOut>\n", argv[0]);
        return 2;
    }
    std::string localradix = argv[1], localref = argv[2], data_glob = strip_scheme(argv[3]), fs_uri = argv[4];
    std::string out_dir = argv[5], gapless_dir = argv[6];
    bool hdfs = fs_uri.rfind("hdfs://", 0) == 0;
    std::string local_out = hdfs ? std::string("/tmp/drlz_out_") + std::to_string(getpid()) : strip_scheme(out_dir);
    std::string local_gap = hdfs ? std::string("/tmp/drlz_gapless_") + std::to_string(getpid()) : strip_scheme(gapless_dir);
    int dev0 = getenv("DRLZ_DEVICE") ? atoi(getenv("DRLZ_DEVICE")) : 0;
    int ngpu = getenv("DRLZ_GPUS") ? atoi(getenv("DRLZ_GPUS")) : 1;
    if (ngpu < 1) ngpu = 1;
    uint64_t batch_bytes = getenv("DRLZ_BATCH_BYTES") ? strtoull(getenv("DRLZ_BATCH_BYTES"), nullptr, 10) : (1ull << 30);

    // reference: all lines concatenated (DistributedRLZ.scala:61)
    std::vector<uint8_t> raw, ref;
    if (!read_file(localref, raw)) { fprintf(stderr, "drlz: cannot read %s\n", localref.c_str()); return 1; }
    ref.reserve(raw.size());
    size_t nl = 0;
    for (uint8_t b : raw) { if (b == '\n' || b == '\r') { ++nl; continue; } ref.push_back(b); }
    if (nl > 1) fprintf(stderr, "drlz: warning: %s has %zu line terminators; the reference expects a single-line, header-less file\n",
                        localref.c_str(), nl);
    raw.clear(); raw.shrink_to_fit();

    // inputs, sorted by name (the order `hdfs dfs -getmerge <outDir>/*.NN.lz` concatenates them, index_chr.sh:16)
    glob_t gl; memset(&gl, 0, sizeof gl);
    std::vector<Job> jobs;
    if (glob(data_glob.c_str(), 0, nullptr, &gl) == 0)
        for (size_t i = 0; i < gl.gl_pathc; ++i) jobs.push_back(Job{gl.gl_pathv[i], base_name(gl.gl_pathv[i]), 0});
    globfree(&gl);
    std::sort(jobs.begin(), jobs.end(), [](const Job& a, const Job& b) { return a.name < b.name; });
    for (size_t i = 0; i < jobs.size(); ++i) jobs[i].index = i;
    if (jobs.empty()) fprintf(stderr, "drlz: warning: no input matches %s\n", data_glob.c_str());
    mkdirs(local_out); mkdirs(local_gap);

    printf("Create suffix\n");
    std::vector<drlz_ctx*> ctxs(ngpu, nullptr);
    std::atomic<int> failed{0};
    {
        std::vector<std::thread> th;
        for (int g = 0; g < ngpu; ++g) th.emplace_back([&, g] {
            if (drlz_create(&ctxs[g], dev0 + g) != DRLZ_OK) { fprintf(stderr, "drlz: no CUDA device %d (there is no CPU fallback)\n", dev0 + g); failed = 1; return; }
            int rc = drlz_build_index(ctxs[g], ref.data(), ref.size());
            if (rc != DRLZ_OK) { fprintf(stderr, "drlz: index build failed on device %d: %s\n", dev0 + g, drlz_last_error(ctxs[g])); failed = 1; }
        });
        for (auto& t : th) t.join();
    }
    if (failed) return 1;
    {
        double ms[5]; uint64_t info[4];
        drlz_index_timings(ctxs[0], ms); drlz_index_info(ctxs[0], info);
        printf("Load suffix\nbroadcasting\nindex: n=%llu k=%llu sigma=%llu bits=%d rounds=%llu build=%.1f ms on %d device(s)\n",
               (unsigned long long)info[0], (unsigned long long)(info[1] & 0xFF), (unsigned long long)(info[1] >> 16),
               1 << ((info[1] >> 8) & 0xFF), (unsigned long long)info[2], ms[0], ngpu);
    }
    if (getenv("DRLZ_EMIT_RADIX") && atoi(getenv("DRLZ_EMIT_RADIX"))) {
        std::vector<uint32_t> sa(ref.size());
        if (drlz_index_get(ctxs[0], 0, sa.data(), sa.size() * 4) == DRLZ_OK) {
            FILE* f = fopen(localradix.c_str(), "w");
            if (f) { for (uint32_t v : sa) fprintf(f, "%u\n", v); fclose(f); }
        }
    }

    printf("started encoding\n");
    std::mutex mu;
    size_t next = 0;
    uint64_t tot_bases = 0, tot_pairs = 0;
    auto worker = [&](int g) {
        drlz_ctx* c = ctxs[g];
        for (;;) {
            std::vector<Job> mine;
            {
                std::lock_guard<std::mutex> lk(mu);
                uint64_t acc = 0;
                while (next < jobs.size() && (mine.empty() || acc < batch_bytes)) {
                    struct stat st; uint64_t sz = stat(jobs[next].path.c_str(), &st) == 0 ? (uint64_t)st.st_size : 0;
                    acc += sz; mine.push_back(jobs[next++]);
                }
            }
            if (mine.empty()) return;
            size_t H = mine.size();
            std::vector<std::vector<uint8_t>> data(H), gap(H);
            std::vector<const uint8_t*> rows(H); std::vector<uint8_t*> gptr(H);
            std::vector<uint64_t> cols(H), m(H), np(H);
            uint64_t sum = 0;
            for (size_t h = 0; h < H; ++h) {
                if (!read_file(mine[h].path, data[h])) { fprintf(stderr, "drlz: cannot read %s\n", mine[h].path.c_str()); failed = 1; return; }
                // one record per line (Spark read.text); only single-line files are supported
                size_t len = data[h].size(), first = len;
                for (size_t i = 0; i < len; ++i) if (data[h][i] == '\n') { first = i; break; }
                for (size_t i = first; i < len; ++i) if (data[h][i] != '\n' && data[h][i] != '\r') {
                    fprintf(stderr, "drlz: %s has more than one line; multi-line rows are unsupported (each line would overwrite the same .lz in the reference)\n", mine[h].path.c_str());
                    failed = 1; return;
                }
                cols[h] = first; rows[h] = data[h].data();
                gap[h].resize(first + 16); gptr[h] = gap[h].data();
                sum += first;
            }
            uint64_t cap = sum / 32 + 1024 * H;
            std::vector<int64_t> pairs;
            int rc;
            for (;;) {
                pairs.resize(2 * cap);
                rc = drlz_parse_batch(c, rows.data(), cols.data(), (uint32_t)H, 1, gptr.data(), m.data(), pairs.data(), cap, np.data());
                if (rc != DRLZ_E_NOSPACE) break;
                cap = 0; for (size_t h = 0; h < H; ++h) cap += np[h];
                cap += 16;
            }
            if (rc != DRLZ_OK) { fprintf(stderr, "drlz: parse failed: %s\n", drlz_last_error(c)); failed = 1; return; }
            uint64_t off = 0;
            for (size_t h = 0; h < H; ++h) {
                std::string lz = local_out + "/" + mine[h].name + ".lz";
                if (!write_file_atomic(lz, pairs.data() + 2 * off, (size_t)np[h] * 16)) { fprintf(stderr, "drlz: cannot write %s\n", lz.c_str()); failed = 1; return; }
                char part[64]; snprintf(part, sizeof part, "/part-%05zu", mine[h].index);
                std::string hdr = ">" + mine[h].name + "\n";
                if (!write_file_atomic(local_gap + part, hdr.data(), hdr.size(), gap[h].data(), (size_t)m[h], "\n", 1)) { fprintf(stderr, "drlz: cannot write %s%s\n", local_gap.c_str(), part); failed = 1; return; }
                off += np[h];
            }
            std::lock_guard<std::mutex> lk(mu);
            for (size_t h = 0; h < H; ++h) { tot_bases += m[h]; tot_pairs += np[h]; }
        }
    };
    {
        std::vector<std::thread> th;
        for (int g = 0; g < ngpu; ++g) th.emplace_back(worker, g);
        for (auto& t : th) t.join();
    }
    for (auto c : ctxs) drlz_destroy(c);
    if (failed) return 1;
    write_file_atomic(local_gap + "/_SUCCESS", "", 0);
    // the two `hdfs dfs -getmerge` steps of components/index/index_chr.sh:16,20 done here --
    // concatenation in sorted-basename order -- when DRLZ_MERGED_LZ / DRLZ_MERGED_FA name the target files.
    for (int which = 0; which < 2; ++which) {
        const char* target = getenv(which == 0 ? "DRLZ_MERGED_LZ" : "DRLZ_MERGED_FA");
        if (!target || !*target) continue;
        std::string tmp = std::string(target) + ".tmp";
        FILE* out = fopen(tmp.c_str(), "wb");
        if (!out) { fprintf(stderr, "drlz: cannot write %s\n", target); return 1; }
        std::vector<uint8_t> buf;
        for (size_t j = 0; j < jobs.size(); ++j) {
            char part[64]; snprintf(part, sizeof part, "/part-%05zu", jobs[j].index);
            std::string src = which == 0 ? local_out + "/" + jobs[j].name + ".lz" : local_gap + part;
            if (!read_file(src, buf) || (buf.size() && fwrite(buf.data(), 1, buf.size(), out) != buf.size())) {
                fprintf(stderr, "drlz: merge failed at %s\n", src.c_str()); fclose(out); return 1;
            }
        }
        fclose(out);
        rename(tmp.c_str(), target);
    }
    printf("encoded %zu haplotypes, %llu bases, %llu phrases\n", jobs.size(), (unsigned long long)tot_bases, (unsigned long long)tot_pairs);
    if (hdfs) {
        std::string cmd = "hdfs dfs -mkdir -p '" + out_dir + "' '" + gapless_dir + "' && hdfs dfs -put -f " + local_out + "/* '" + out_dir +
                          "/' && hdfs dfs -put -f " + local_gap + "/* '" + gapless_dir + "/'";
        int rc = system(cmd.c_str());
        if (rc != 0) { fprintf(stderr, "drlz: upload to %s failed (is the `hdfs` CLI on PATH?); outputs kept in %s and %s\n", fs_uri.c_str(), local_out.c_str(), local_gap.c_str()); return 1; }
    }
    return 0;
}
