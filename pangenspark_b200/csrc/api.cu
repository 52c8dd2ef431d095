// api.cu -- C ABI of libdrlz.so (include/drlz.h): context, device buffers, index build, batched parse.
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <chrono>
#include <string>
#include <thread>
#include <vector>
#include <nvtx3/nvToolsExt.h>      // header-only NVTX 3: ranges show up in Nsight timelines, cost nothing without a tool attached
#include "../../include/drlz.h"
#include "drlz_internal.h"
#include "scan.cuh"

namespace drlz {
int g_sort_variant = 1;                              // which scatter kernel radix_sort_pairs launches (option sort_variant; process-wide): 1 = k_sort_scatter2
thread_local unsigned long long g_launches = 0;   // kernels launched by the calling thread (a parse call runs on one thread)

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};
struct HostBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};
}  // namespace drlz

using namespace drlz;

struct drlz_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr};
    // the gapless sequences of a group go back to the host on their own stream while the next group is parsed: two device
    // buffers in turn; ev_gl_done[b] = the copies out of buffer b have finished, ev_gl_ready = the group's kernels have
    cudaEvent_t ev_gl_done[2] = {nullptr, nullptr}, ev_gl_ready = nullptr;
    int gl_cur = 0; bool gl_pending = false;
    cudaEvent_t ev_stage[12] = {};
    std::string err;
    // options
    uint64_t x_zeroed = 0; bool keep_x = false; int opt_postzero = 1; int opt_keep_timings = 0; int opt_blind_rounds = 2; int opt_pack_variant = 2; uint32_t opt_prefetch3 = 2; int opt_occ_pack = 8; uint32_t opt_prefetch = 1200; int opt_k = 0; uint32_t opt_chunk = 8192; uint32_t opt_min_cap = 64; uint32_t opt_foreign_cap = 1u << 20; uint64_t opt_group_bytes = 256ull << 20; int opt_lcp = 0;
    // index
    bool have_index = false; uint64_t n = 0; int k = 0; int sa_rounds = 0;
    int opt_lcp_ties = 1;                  // LCP after <= 3 doubling rounds from the sort keys + k_lcp_ties (0 = always the Kasai-order kernel)
    int opt_l2_fetch = 0, opt_l2_fetch_parse = 0;   // L2 fetch granularity: for the whole device / around the parse kernels only (0 = leave it alone)
    bool isa_valid = false;                // the inverse suffix array is filled (a build without doubling rounds leaves it out)
    DevBuf blob, sa, isa, lcp;             // blob = [k-mer table | packed text] (one L2 access-policy window)
    uint64_t* d_text = nullptr; uint32_t* d_tab = nullptr; uint16_t* d_uniq = nullptr; uint8_t* d_lut = nullptr;   // lut = b2c[256], c2b[256]
    size_t blob_bytes = 0, text_bytes = 0, uniq_bytes = 0, tab_bytes = 0;
    uint32_t bl = 1, sigma = 4;        // symbol width (log2 bits) and alphabet size of the index
    bool has_uniq32 = false;           // some repeat of the reference reaches 65535 symbols: the unsaturated array (behind the SA) is in use
    int opt_predict = 1;
    cudaStream_t l2_stream = nullptr; int opt_l2persist = 0; int opt_occ = 8; int opt_occ_fix = 10; int opt_occ_diag = 12; int opt_spec_sync = 1; int opt_uniq_coarse = 1; int opt_tickets = 1; int opt_hint = 1; int opt_diag = 1; uint32_t opt_spec_store = 0;
    double idx_ms[5] = {0, 0, 0, 0, 0};
    // batch scratch (grow-only)
    SaWorkspace sa_ws;
    DevBuf stage[2], meta_dev, desc_cnt, desc_exc, tile_counter, X, mdev, exccnt, excpos, excbyte,
        flags, haps, gapless, gtile_cnt, gtile_exc, gtile_off, gtile_excoff, gscan_tmp, chunk_hap, first_pos, first_len, run_a, run_b, link_a, link_b, open_i, open_pos, open_len, spec_cnt, spec_exit, alt_entry, alt_cnt, alt_exit, alt_round, alt_own, alt_merge, spec_ph, alt_ph, hint, mm, true_slot, counters, jump_a,
        jump_b, mark, true_cnt, true_entry, true_off, scan_tmp, hap_pairs, out, ms_a, ms_b, tab_tmp, gap_list, gap_scratch, gap_counter, ext, hint_src, uniqc, gapless_b;
    HostBuf meta_host, small_host, gap_host;
    uint64_t exc_stride_hint = 0;
    // thread safety: every entry point takes `mu` (calls from several threads are serialised on the device)
    std::recursive_mutex mu;
    // async form (drlz_submit / drlz_wait): one worker thread per context, FIFO
    struct Job { uint64_t ticket; const uint8_t* const* rows; const uint64_t* cols; uint32_t H; int strip; uint8_t* const* gapless;
                 uint64_t* m_out; int64_t* pairs; uint64_t cap; uint64_t* n_pairs; int64_t* gaps; uint64_t gaps_cap; uint64_t* n_gaps; };
    std::mutex qmu; std::condition_variable qcv, done_cv;
    std::deque<Job> queue; std::vector<std::pair<uint64_t, int>> done;
    std::vector<uint64_t> outstanding;      // tickets submitted and not yet collected by drlz_wait
    std::thread worker; bool worker_started = false, stopping = false; uint64_t next_ticket = 1;
    double parse_ms[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
};

#define CK(ctx, x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { (ctx)->err = std::string(#x) + ": " + cudaGetErrorString(e_); return e_ == cudaErrorMemoryAllocation ? DRLZ_E_NOMEM : DRLZ_E_CUDA; } } while (0)

static const uint64_t kMaxLen = 0xFFFFFFFFull - 1024;

static uint64_t table_entries(uint32_t sigma, int k) {
    uint64_t e = 1;
    for (int t = 0; t < k; ++t) e *= sigma;
    return e + 1;
}

// blob = [k-mer table | packed text | uniqueness array | byte<->code tables]
static cudaError_t alloc_index_blob(drlz_ctx* c, uint64_t n, int k, uint32_t sigma, uint32_t bl) {
    size_t tab_bytes = (table_entries(sigma, k) * 4 + 255) & ~(size_t)255;
    size_t text_bytes = (((n >> (6 - bl)) + 8) * 8 + 64 + 255) & ~(size_t)255;
    size_t uniq_bytes = ((n + 1) * 2 + 255) & ~(size_t)255;
    cudaError_t e = c->blob.ensure(tab_bytes + text_bytes + uniq_bytes + 512);
    if (e != cudaSuccess) return e;
    c->d_tab = c->blob.as<uint32_t>();
    c->d_text = reinterpret_cast<uint64_t*>(c->blob.as<uint8_t>() + tab_bytes);
    c->d_uniq = reinterpret_cast<uint16_t*>(c->blob.as<uint8_t>() + tab_bytes + text_bytes);
    c->d_lut = c->blob.as<uint8_t>() + tab_bytes + text_bytes + uniq_bytes;
    c->tab_bytes = tab_bytes; c->text_bytes = text_bytes; c->uniq_bytes = uniq_bytes;
    c->blob_bytes = tab_bytes + text_bytes + uniq_bytes + 512;
    c->bl = bl; c->sigma = sigma;
    c->l2_stream = nullptr;
    return cudaSuccess;
}

// keep the packed reference (the target of every candidate probe) resident in L2 while haplotype data, table and
// suffix-array probes stream through
static void apply_l2_policy(drlz_ctx* c) {
    if (!c->blob.p || c->l2_stream == c->stream) return;
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof attr);
    if (c->opt_l2persist) {
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, c->device);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, c->device);
        size_t text_bytes = c->text_bytes;
        if (max_persist > 0 && max_window > 0) {
            size_t want = text_bytes < (size_t)max_persist ? text_bytes : (size_t)max_persist;
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
            size_t win = text_bytes < (size_t)max_window ? text_bytes : (size_t)max_window;
            attr.accessPolicyWindow.base_ptr = c->d_text;
            attr.accessPolicyWindow.num_bytes = win;
            attr.accessPolicyWindow.hitRatio = win <= want ? 1.0f : (float)((double)want / (double)win);
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
        }
    }
    cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
    c->l2_stream = c->stream;
}

extern "C" {

int drlz_create(drlz_ctx** out, int device) {
    if (!out) return DRLZ_E_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return DRLZ_E_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return DRLZ_E_CUDA;
    drlz_ctx* c = new drlz_ctx();
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return DRLZ_E_CUDA; }
    c->stream = c->own_stream;
    for (int i = 0; i < 2; ++i) cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming);
    for (int i = 0; i < 2; ++i) cudaEventCreateWithFlags(&c->ev_gl_done[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_gl_ready, cudaEventDisableTiming);
    for (int i = 0; i < 12; ++i) cudaEventCreate(&c->ev_stage[i]);
    *out = c;
    return DRLZ_OK;
}

void drlz_destroy(drlz_ctx* c) {
    if (!c) return;
    if (c->worker_started) {
        { std::lock_guard<std::mutex> lk(c->qmu); c->stopping = true; }
        c->qcv.notify_all();
        c->worker.join();
    }
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    c->sa_ws.release();
    DevBuf* bufs[] = {&c->blob, &c->sa, &c->isa, &c->lcp, &c->stage[0], &c->stage[1], &c->meta_dev, &c->desc_cnt,
                      &c->desc_exc, &c->tile_counter, &c->X, &c->mdev, &c->exccnt, &c->excpos,
                      &c->excbyte, &c->flags, &c->haps, &c->gapless, &c->gtile_cnt, &c->gtile_exc, &c->gtile_off, &c->gtile_excoff, &c->gscan_tmp, &c->chunk_hap, &c->first_pos, &c->first_len, &c->run_a, &c->run_b, &c->link_a, &c->link_b, &c->open_i, &c->open_pos, &c->open_len, &c->spec_cnt, &c->spec_exit,
                      &c->alt_entry, &c->alt_cnt, &c->alt_exit, &c->alt_round, &c->alt_own, &c->alt_merge, &c->spec_ph, &c->alt_ph, &c->hint, &c->mm, &c->true_slot, &c->counters, &c->jump_a, &c->jump_b,
                      &c->mark, &c->true_cnt, &c->true_entry, &c->true_off, &c->scan_tmp, &c->hap_pairs, &c->out, &c->ms_a, &c->ms_b, &c->tab_tmp, &c->gap_list, &c->gap_scratch, &c->gap_counter, &c->ext, &c->hint_src, &c->uniqc, &c->gapless_b};
    for (DevBuf* b : bufs) b->release();
    c->meta_host.release(); c->small_host.release(); c->gap_host.release();
    for (int i = 0; i < 2; ++i) if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]);
    for (int i = 0; i < 12; ++i) if (c->ev_stage[i]) cudaEventDestroy(c->ev_stage[i]);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    for (int i = 0; i < 2; ++i) if (c->ev_gl_done[i]) cudaEventDestroy(c->ev_gl_done[i]);
    if (c->ev_gl_ready) cudaEventDestroy(c->ev_gl_ready);
    delete c;
}

const char* drlz_last_error(const drlz_ctx* c) { return c ? c->err.c_str() : "null context"; }

int drlz_set_stream(drlz_ctx* c, void* s) {
    if (!c) return DRLZ_E_ARG;
    std::lock_guard<std::recursive_mutex> lk_(c->mu);
    cudaStream_t ns = s ? (cudaStream_t)s : c->own_stream;
    if (ns != c->stream) {
        // work may still be queued on the old stream (the zero-fill of the packed buffer behind the last batch is not
        // awaited): the first kernel on the new stream must not overtake it
        CK(c, cudaSetDevice(c->device));
        CK(c, cudaStreamSynchronize(c->stream));
        c->stream = ns;
    }
    return DRLZ_OK;
}

int drlz_set_option(drlz_ctx* c, const char* key, int64_t v) {
    if (!c || !key) return DRLZ_E_ARG;
    std::lock_guard<std::recursive_mutex> lk_(c->mu);
    if (!strcmp(key, "kmer")) { if (v < 0 || v > 32) return DRLZ_E_ARG; c->opt_k = (int)v; }
    else if (!strcmp(key, "chunk")) { if (v < 0 || v > 0x7FFFFFFF) return DRLZ_E_ARG; c->opt_chunk = v ? (uint32_t)v : 8192; }
    else if (!strcmp(key, "group_bytes")) { if (v < 0) return DRLZ_E_ARG; c->opt_group_bytes = v ? (uint64_t)v : (256ull << 20); }
    else if (!strcmp(key, "lcp")) c->opt_lcp = v != 0;
    else if (!strcmp(key, "l2persist")) { c->opt_l2persist = v != 0; c->l2_stream = nullptr; }
    else if (!strcmp(key, "lcp_ties")) c->opt_lcp_ties = v != 0;
    else if (!strcmp(key, "table_variant")) { if (v < 0 || v > 1) return DRLZ_E_ARG; g_table_variant = (int)v; }
    else if (!strcmp(key, "sort_variant")) { if (v < 0 || v > 2) return DRLZ_E_ARG; g_sort_variant = (int)v; }
    else if (!strcmp(key, "l2_fetch")) {
        // bytes the L2 fetches from HBM on a miss (cudaLimitMaxL2FetchGranularity: 32 / 64 / 128; a device-wide hint).  The
        // lookups of the parse read single 32-byte sectors at scattered addresses
        if (v != 32 && v != 64 && v != 128) return DRLZ_E_ARG;
        CK(c, cudaSetDevice(c->device));
        CK(c, cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)v));
        c->opt_l2_fetch = (int)v;
    }
    else if (!strcmp(key, "l2_fetch_parse")) { if (v != 0 && v != 32 && v != 64 && v != 128) return DRLZ_E_ARG; c->opt_l2_fetch_parse = (int)v; }
    else if (!strcmp(key, "occ")) c->opt_occ = (int)v;
    else if (!strcmp(key, "occ_fix")) c->opt_occ_fix = (int)v;
    else if (!strcmp(key, "occ_diag")) c->opt_occ_diag = (int)v;
    else if (!strcmp(key, "spec_sync")) c->opt_spec_sync = (int)v;
    else if (!strcmp(key, "uniq_coarse")) c->opt_uniq_coarse = (int)v;
    else if (!strcmp(key, "hint")) c->opt_hint = v != 0;
    else if (!strcmp(key, "diag")) c->opt_diag = v != 0;
    else if (!strcmp(key, "spec_store")) { if (v < 0 || v > 4096) return DRLZ_E_ARG; c->opt_spec_store = (uint32_t)v; }
    else if (!strcmp(key, "tickets")) c->opt_tickets = v != 0;
    else if (!strcmp(key, "predict")) c->opt_predict = v != 0;
    else if (!strcmp(key, "occ_pack")) c->opt_occ_pack = (int)v;
    else if (!strcmp(key, "keep_timings")) { c->opt_keep_timings = v != 0; for (int i = 0; i < 12; ++i) c->parse_ms[i] = 0; }
    else if (!strcmp(key, "prefetch3")) { if (v < 0 || v > 64) return DRLZ_E_ARG; c->opt_prefetch3 = (uint32_t)v; }
    else if (!strcmp(key, "pack_variant")) { if (v < 0 || v > 2) return DRLZ_E_ARG; c->opt_pack_variant = (int)v; }
    else if (!strcmp(key, "blind_rounds")) c->opt_blind_rounds = (int)v;
    else if (!strcmp(key, "postzero")) c->opt_postzero = v != 0;
    else if (!strcmp(key, "prefetch")) { if (v < 0 || v > (1 << 20)) return DRLZ_E_ARG; c->opt_prefetch = (uint32_t)v; }
    else if (!strcmp(key, "min_cap")) { if (v < 0 || v > 0x7FFFFFFF) return DRLZ_E_ARG; c->opt_min_cap = v ? (uint32_t)v : 64; }
    else if (!strcmp(key, "foreign_cap")) { if (v < 0 || v > 0x7FFFFFFF) return DRLZ_E_ARG; c->opt_foreign_cap = v ? (uint32_t)v : (1u << 20); }
    else return DRLZ_E_ARG;
    return DRLZ_OK;
}

void* drlz_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void drlz_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// device pipeline for one group of rows already resident in device memory
// ---------------------------------------------------------------------------------------------------
struct GroupResult { uint64_t total_pairs = 0; };
struct NvtxRange { explicit NvtxRange(const char* name) { nvtxRangePushA(name); } ~NvtxRange() { nvtxRangePop(); } };

static int run_device_group(drlz_ctx* c, const uint8_t* rows_dev, const uint64_t* row_off, const uint64_t* cols, uint32_t H,
                            int strip, bool want_gapless, uint64_t* m_out, uint64_t* n_pairs, uint64_t* total_pairs,
                            uint64_t* gapless_off_out /*[H], host*/) {
    cudaStream_t s = c->stream;
    NvtxRange nvtx_group("drlz.group: strip+pack, parse, records");
    const auto t_entry = std::chrono::steady_clock::now();
    *total_pairs = 0;
    if (H == 0) return DRLZ_OK;
    apply_l2_policy(c);
    if (H > 65535) { c->err = "more than 65535 rows in one device group"; return DRLZ_E_LIMIT; }
    uint64_t maxcols = 0, sumcols = 0;
    for (uint32_t h = 0; h < H; ++h) {
        if (cols[h] > kMaxLen) { c->err = "row longer than 2^32-1024"; return DRLZ_E_LIMIT; }
        if (row_off[h] & 15) { c->err = "row offset not a multiple of 16"; return DRLZ_E_ARG; }
        if (cols[h] > maxcols) maxcols = cols[h];
        sumcols += cols[h];
    }
    uint32_t T = (uint32_t)((maxcols + kPackTileBytes - 1) / kPackTileBytes);
    if (T == 0) T = 1;
    uint64_t nt = (uint64_t)H * T;
    uint32_t chunk_shift = 0;
    while ((2u << chunk_shift) <= c->opt_chunk && chunk_shift < 30) ++chunk_shift;    // chunk size: power of two
    uint32_t C = 1u << chunk_shift;
    bool single_chunk = false;
    unsigned long long launches0 = g_launches;

    // host metadata: row_off[H] cols[H] xoff[H] gapless_off[H] (u64) then base[H+1] (u32)
    size_t meta_bytes = sizeof(uint64_t) * 4 * H + sizeof(uint32_t) * (H + 1) + 64;
    CK(c, c->meta_host.ensure(meta_bytes));
    CK(c, c->meta_dev.ensure(meta_bytes));
    uint64_t* mh = c->meta_host.as<uint64_t>();
    uint64_t* h_row_off = mh; uint64_t* h_cols = mh + H; uint64_t* h_xoff = mh + 2 * H; uint64_t* h_goff = mh + 3 * H;
    uint32_t* h_base = reinterpret_cast<uint32_t*>(mh + 4 * H);
    uint64_t xwords = 0, gbytes = 0;
    for (uint32_t h = 0; h < H; ++h) {
        h_row_off[h] = row_off[h]; h_cols[h] = cols[h];
        h_xoff[h] = xwords; xwords += (cols[h] >> (6 - c->bl)) + 8;
        h_goff[h] = gbytes; gbytes += (cols[h] + 15) & ~15ull;
        if (gapless_off_out) gapless_off_out[h] = h_goff[h];
    }
    const size_t res_bytes = sizeof(uint64_t) * (3 * (size_t)H + 2) + sizeof(uint32_t) * (4 + kMaxRounds);   // see k_results
    CK(c, c->small_host.ensure(res_bytes + 64));
    uint64_t* sh_m = c->small_host.as<uint64_t>();            // [H]
    uint64_t* sh_pairs = sh_m + H;                            // [H+1]
    uint64_t* sh_total = sh_pairs + H + 1;                    // [1]
    uint64_t* sh_exccnt = sh_total + 1;                       // [H] exceptions per row
    uint32_t* sh_flags = reinterpret_cast<uint32_t*>(sh_exccnt + H);   // [1] (+1 pad)
    uint32_t* sh_counters = sh_flags + 2;                     // [2 + kMaxRounds]

    CK(c, c->desc_cnt.ensure(nt * 8)); CK(c, c->desc_exc.ensure(nt * 8)); CK(c, c->tile_counter.ensure(16));
    {
        void* before = c->X.p;
        CK(c, c->X.ensure(xwords * 8 + 256));
        if (c->X.p != before) c->x_zeroed = 0;
    }
    uint32_t GT = (uint32_t)((maxcols + kPackGTileBytes - 1) / kPackGTileBytes); if (GT == 0) GT = 1;
    if (c->bl != 1) {
        uint64_t gnt = (uint64_t)H * GT;
        CK(c, c->gtile_cnt.ensure(gnt * 4)); CK(c, c->gtile_exc.ensure(gnt * 4));
        CK(c, c->gtile_off.ensure(gnt * 8)); CK(c, c->gtile_excoff.ensure(gnt * 8));
        CK(c, c->gscan_tmp.ensure(scan_tmp_elems(gnt) * 8 * 2));
    }
    CK(c, c->mdev.ensure((size_t)H * 8)); CK(c, c->exccnt.ensure(((size_t)H + 1) * 8));
    CK(c, c->flags.ensure(16)); CK(c, c->haps.ensure((size_t)H * sizeof(HapView)));
    DevBuf& glbuf = c->gl_cur ? c->gapless_b : c->gapless;
    if (want_gapless) {
        CK(c, glbuf.ensure(gbytes + 16));
        CK(c, cudaStreamWaitEvent(c->stream, c->ev_gl_done[c->gl_cur], 0));      // the copies of the group before last out of this buffer
    }
    if (c->exc_stride_hint < maxcols / 256 + 256) c->exc_stride_hint = maxcols / 256 + 256;

    // a group is redone when (a) the fix-up did not converge with chunks (once: one chunk per haplotype, always exact),
    // (b) the look-back of k_pack timed out under blockIdx tile ids (once: atomic tickets cannot time out), (c) a row
    // has more exceptions than the list holds (the list is then sized from the exact counts; twice at most, the
    // second time only if a ticket retry came in between).  Each cause has its own budget and its own message.
    int retry_ticket = 0, retry_exc = 0;
    for (;;) {
        uint64_t exc_stride = c->exc_stride_hint;
        CK(c, c->excpos.ensure(exc_stride * H * 4)); CK(c, c->excbyte.ensure(exc_stride * H));
        // chunk layout (upper bound from cols; chunks past the gapless length stay empty)
        uint64_t NC64 = 0; uint32_t maxc = 0;
        for (uint32_t h = 0; h < H; ++h) {
            h_base[h] = (uint32_t)NC64;
            uint64_t nc = single_chunk ? (cols[h] ? 1 : 0) : (cols[h] + C - 1) / C;
            if (nc > maxc) maxc = (uint32_t)nc;
            NC64 += nc;
            if (NC64 * (kAltSlots + 1) >= 0xFFFFFFF0ull) { c->err = "too many chunks in one group"; return DRLZ_E_LIMIT; }
        }
        h_base[H] = (uint32_t)NC64;
        uint32_t NC = (uint32_t)NC64;
        uint32_t Ceff = single_chunk ? 0xFFFFFFFFu : C;
        CK(c, cudaMemcpyAsync(c->meta_dev.p, c->meta_host.p, meta_bytes, cudaMemcpyHostToDevice, s));
        const uint64_t* d_row_off = c->meta_dev.as<uint64_t>();
        const uint64_t* d_cols = d_row_off + H; const uint64_t* d_xoff = d_row_off + 2 * H; const uint64_t* d_goff = d_row_off + 3 * H;
        const uint32_t* d_base = reinterpret_cast<const uint32_t*>(d_row_off + 4 * H);

        cudaEventRecord(c->ev_stage[0], s);
        const auto t_first = std::chrono::steady_clock::now();
        // the packed buffer must be zero where k_pack ORs in the words at tile borders; in the steady state the previous
        // batch left it zeroed behind its last kernel (below), while the host was busy between the two calls
        if (c->x_zeroed < xwords) CK(c, cudaMemsetAsync(c->X.p, 0, xwords * 8, s));
        c->x_zeroed = 0;
        CK(c, cudaMemsetAsync(c->flags.p, 0, 16, s));
        cudaEventRecord(c->ev_stage[1], s);
        PackJob job;
        job.rows = rows_dev; job.row_off = d_row_off; job.cols = d_cols; job.H = H; job.tiles_per_row = T; job.strip = strip;
        job.desc_cnt = c->desc_cnt.as<uint64_t>(); job.desc_exc = c->desc_exc.as<uint64_t>();
        job.tile_counter = c->tile_counter.as<uint32_t>(); job.X = c->X.as<uint64_t>(); job.xoff = d_xoff;
        job.occ = c->opt_occ_pack; job.prefetch = c->opt_pack_variant == 2 ? c->opt_prefetch3 : c->opt_prefetch; job.use_tickets = c->opt_tickets; job.variant = c->opt_pack_variant;
        job.uniform_cols = 0; job.uniform_stride = 0; job.uniform_off0 = 0;
        if (H > 0 && cols[0] > 0) {           // an MSA: rows of one length at a constant spacing (the usual case)
            bool uni = true;
            const uint64_t st = H > 1 ? row_off[1] - row_off[0] : 0;
            for (uint32_t h = 0; h < H && uni; ++h) uni = cols[h] == cols[0] && row_off[h] == row_off[0] + (uint64_t)h * st;
            if (uni) { job.uniform_cols = (uint32_t)cols[0]; job.uniform_stride = st; job.uniform_off0 = row_off[0]; }
        }
        job.m_out = c->mdev.as<uint64_t>(); job.exc_cnt = c->exccnt.as<uint64_t>();
        job.exc_pos = c->excpos.as<uint32_t>(); job.exc_byte = c->excbyte.as<uint8_t>(); job.exc_stride = exc_stride;
        job.flags = c->flags.as<uint32_t>();
        job.gapless = want_gapless ? glbuf.as<uint8_t>() : nullptr; job.gapless_off = d_goff;
        job.bl = c->bl; job.b2c = c->d_lut; job.gtiles_per_row = GT;
        job.gtile_cnt = c->gtile_cnt.as<uint32_t>(); job.gtile_exc = c->gtile_exc.as<uint32_t>();
        job.gtile_off = c->gtile_off.as<uint64_t>(); job.gtile_excoff = c->gtile_excoff.as<uint64_t>();
        job.gscan_tmp = c->gscan_tmp.as<uint64_t>();
        { NvtxRange r("drlz.pack"); launch_pack(job, s, &c->ev_stage[2]); }                            // ev 2,3,4
        launch_make_hapviews(c->haps.as<HapView>(), c->X.as<uint64_t>(), d_xoff, c->mdev.as<uint64_t>(), c->excpos.as<uint32_t>(),
                             c->excbyte.as<uint8_t>(), c->exccnt.as<uint64_t>(), exc_stride, H, s);

        size_t nn = (size_t)NC * (kAltSlots + 1), na = (size_t)NC * kAltSlots;
        CK(c, c->first_pos.ensure((size_t)NC * 4 + 4)); CK(c, c->first_len.ensure((size_t)NC * 4 + 4));
        CK(c, c->run_a.ensure((size_t)NC * 4 + 4)); CK(c, c->run_b.ensure((size_t)NC * 4 + 4));
        CK(c, c->link_a.ensure((size_t)NC * 4 + 4)); CK(c, c->link_b.ensure((size_t)NC * 4 + 4));
        CK(c, c->open_i.ensure((size_t)NC * 4 + 4)); CK(c, c->open_pos.ensure((size_t)NC * 4 + 4)); CK(c, c->open_len.ensure((size_t)NC * 4 + 4));
        CK(c, c->chunk_hap.ensure((size_t)NC * 4 + 4)); CK(c, c->spec_cnt.ensure((size_t)NC * 4 + 4)); CK(c, c->spec_exit.ensure((size_t)NC * 4 + 4));
        CK(c, c->alt_entry.ensure(na * 4 + 4)); CK(c, c->alt_cnt.ensure(na * 4 + 4)); CK(c, c->alt_exit.ensure(na * 4 + 4));
        CK(c, c->alt_own.ensure(na * 4 + 4)); CK(c, c->alt_merge.ensure(na * 4 + 4));
        uint32_t spec_store = c->opt_spec_store ? c->opt_spec_store : (C / 32 > 16 ? C / 32 : 16);   // 2 phrases per SNP: covers ~1.5 % divergence
        if (single_chunk) spec_store = 1;
        CK(c, c->spec_ph.ensure((size_t)NC * spec_store * 8 + 8)); CK(c, c->alt_ph.ensure(na * kAltStore * 8 + 8));
        CK(c, c->hint.ensure(((size_t)NC / kHintGroup + 2) * 8)); CK(c, c->mm.ensure(((size_t)NC + 1) * kMmStride * 2));
        CK(c, c->true_slot.ensure((size_t)NC * 4 + 4));
        if (c->has_uniq32) { CK(c, c->ext.ensure((size_t)NC * 4 + 4)); CK(c, c->hint_src.ensure((size_t)NC * 4 + 4)); }
        CK(c, c->alt_round.ensure(na * 4 + 4)); CK(c, c->counters.ensure(sizeof(uint32_t) * (2 + kMaxRounds)));
        CK(c, c->jump_a.ensure(nn * 4 + 4)); CK(c, c->jump_b.ensure(nn * 4 + 4)); CK(c, c->mark.ensure(nn + 4));
        CK(c, c->true_cnt.ensure((size_t)NC * 4 + 4)); CK(c, c->true_entry.ensure((size_t)NC * 4 + 4)); CK(c, c->true_off.ensure((size_t)NC * 8 + 8));
        CK(c, c->scan_tmp.ensure(scan_tmp_elems(NC) * 8)); CK(c, c->hap_pairs.ensure(res_bytes + 64));
        launch_fill_chunk_hap(c->chunk_hap.as<uint32_t>(), d_base, H, NC, s);

        ParseBuffers pb;
        pb.pv.ix = IndexView{c->d_text, c->sa.as<uint32_t>(), c->d_tab, c->n, c->k, c->opt_predict ? c->d_uniq : nullptr,
                             c->bl, c->sigma, c->d_lut + 256, (c->opt_predict && c->has_uniq32) ? c->sa.as<uint32_t>() + c->n : nullptr,
                             (c->opt_predict && c->opt_uniq_coarse && c->uniqc.p) ? c->uniqc.as<uint16_t>() : nullptr};
        pb.pv.haps = c->haps.as<HapView>(); pb.pv.chunk_hap = c->chunk_hap.as<uint32_t>(); pb.pv.hap_chunk_base = d_base;
        pb.pv.chunk_bases = Ceff; pb.pv.chunk_shift = single_chunk ? 32u : chunk_shift; pb.pv.n_chunks = NC;
        pb.pv.min_cap = single_chunk ? 0xFFFFFFFFu : c->opt_min_cap; pb.pv.foreign_cap = c->opt_foreign_cap;
        pb.pv.first_pos = c->first_pos.as<uint32_t>(); pb.pv.first_len = c->first_len.as<uint32_t>();
        pb.pv.run_len = c->run_a.as<uint32_t>(); pb.pv.link = c->link_a.as<uint32_t>();
        pb.link_b = c->link_b.as<uint32_t>(); pb.run_b = c->run_b.as<uint32_t>();
        pb.pv.open_i = c->open_i.as<uint32_t>(); pb.pv.open_pos = c->open_pos.as<uint32_t>(); pb.pv.open_len = c->open_len.as<uint32_t>();
        pb.pv.spec_cnt = c->spec_cnt.as<uint32_t>(); pb.pv.spec_exit = c->spec_exit.as<uint32_t>();
        pb.pv.alt_entry = c->alt_entry.as<uint32_t>(); pb.pv.alt_cnt = c->alt_cnt.as<uint32_t>();
        pb.pv.alt_exit = c->alt_exit.as<uint32_t>(); pb.pv.alt_round = c->alt_round.as<uint32_t>();
        pb.pv.alt_own = c->alt_own.as<uint32_t>(); pb.pv.alt_merge = c->alt_merge.as<uint32_t>();
        pb.pv.spec_ph = c->spec_ph.as<uint2>(); pb.pv.alt_ph = c->alt_ph.as<uint2>();
        pb.pv.spec_store = spec_store; pb.pv.hint = nullptr; pb.pv.mm = nullptr;
        pb.mm = (c->opt_hint && c->opt_diag && !single_chunk) ? c->mm.as<uint16_t>() : nullptr;
        pb.hint = (c->opt_hint && !single_chunk) ? c->hint.as<int64_t>() : nullptr;
        pb.ext = c->has_uniq32 ? c->ext.as<uint32_t>() : nullptr; pb.hint_src = c->has_uniq32 ? c->hint_src.as<uint32_t>() : nullptr;
        pb.true_slot = c->true_slot.as<uint32_t>();
        pb.pv.counters = c->counters.as<uint32_t>(); pb.pv.spec_sync = (uint32_t)c->opt_spec_sync;
        pb.H = H; pb.max_chunks_per_hap = maxc ? maxc : 1; pb.occ = c->opt_occ; pb.occ_fix = c->opt_occ_fix; pb.occ_diag = c->opt_occ_diag; pb.blind_rounds = c->opt_blind_rounds;
        pb.jump_a = c->jump_a.as<uint32_t>(); pb.jump_b = c->jump_b.as<uint32_t>(); pb.mark = c->mark.as<uint8_t>();
        pb.true_cnt = c->true_cnt.as<uint32_t>(); pb.true_entry = c->true_entry.as<uint32_t>(); pb.true_off = c->true_off.as<uint64_t>();
        pb.scan_tmp = c->scan_tmp.as<uint64_t>();
        pb.res = c->hap_pairs.as<uint64_t>(); pb.host_res = sh_m; pb.res_bytes = res_bytes;
        pb.m_dev = c->mdev.as<uint64_t>(); pb.exccnt_dev = c->exccnt.as<uint64_t>(); pb.flags_dev = c->flags.as<uint32_t>();
        pb.host_counters = sh_counters; pb.host_total = sh_total;
        int rounds = 0;
        pb.spec_out = c->out.as<int64_t>(); pb.spec_out_cap = c->out.cap >= 32 ? (c->out.cap - 16) / 16 : 0;
        pb.ev_emitted = c->ev_stage[11]; pb.emitted = false;
        int prc;
        if (c->opt_l2_fetch_parse) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)c->opt_l2_fetch_parse);
        { NvtxRange r("drlz.parse: pre-scan, speculative parse, fix-up, marking, emit"); prc = parse_stage_count(pb, s, &rounds, &c->ev_stage[5]); }
        if (c->opt_l2_fetch_parse) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)(c->opt_l2_fetch ? c->opt_l2_fetch : 64));   // ev 5,6,7,8 and 9 (after stage A)
        CK(c, cudaGetLastError());
        if (prc == 2) {
            if (single_chunk) { c->err = "parse fix-up did not converge"; return DRLZ_E_INTERNAL; }
            single_chunk = true;        // exact sequential fallback on the device: one chunk per haplotype
            continue;
        }
        auto t_done = std::chrono::steady_clock::now();            // parse_stage_count returned after its sync
        uint64_t total = *sh_total;
        if (!pb.emitted || total > pb.spec_out_cap) {                    // first batch, or more records than ever before
            CK(c, c->out.ensure(total * 16 + total * 2 + 16));
            parse_stage_emit(pb, c->out.as<int64_t>(), s);
            cudaEventRecord(c->ev_stage[11], s);
            CK(c, cudaStreamSynchronize(s));
            t_done = std::chrono::steady_clock::now();
        }
        if (*sh_flags & 2u) {           // look-back timed out (CTAs not dispatched in index order): redo with tickets
            if (c->opt_tickets || ++retry_ticket > 1) { c->err = "strip/pack look-back timed out although tile ids come from atomic tickets (is the GPU shared with another process?)"; return DRLZ_E_INTERNAL; }
            c->opt_tickets = 1;
            continue;
        }
        if (*sh_flags & 1u) {           // exception list overflowed: size it from the exact counts and redo the group
            if (++retry_exc > 3) { c->err = "exception list kept overflowing"; return DRLZ_E_INTERNAL; }
            uint64_t mx = 0;
            for (uint32_t h = 0; h < H; ++h) if (sh_exccnt[h] > mx) mx = sh_exccnt[h];
            c->exc_stride_hint = mx + mx / 8 + 1024;
            continue;
        }
        for (uint32_t h = 0; h < H; ++h) {
            if (m_out) m_out[h] = sh_m[h];
            if (n_pairs) n_pairs[h] = sh_pairs[h + 1] - sh_pairs[h];
        }
        *total_pairs = total;
        if (c->opt_postzero && !c->keep_x) {      // nothing reads the packed haplotypes any more: zero them for the next batch, unawaited
            CK(c, cudaMemsetAsync(c->X.p, 0, xwords * 8, s));
            c->x_zeroed = xwords;
        }
        // timings
        float t;
        double* pm = c->parse_ms;
        static const int iv[10][2] = {{0, 1}, {5, 10}, {2, 3}, {3, 4}, {10, 6}, {9, 7}, {7, 8}, {8, 11}, {0, 11}, {6, 9}};
        for (int i = 0; i < 10; ++i) { cudaEventElapsedTime(&t, c->ev_stage[iv[i][0]], c->ev_stage[iv[i][1]]); pm[i] += t; }
        pm[10] += rounds; pm[11] += (double)(g_launches - launches0);
        // [2]: host time of this call during which the GPU had nothing of it to run (before the first launch, after the last sync)
        pm[2] = (c->opt_keep_timings ? pm[2] : 0.0) + std::chrono::duration<double, std::milli>(t_first - t_entry).count() +
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_done).count();
        return DRLZ_OK;
    }
}

extern "C" {

// ---------------------------------------------------------------------------------------------------
// index
// ---------------------------------------------------------------------------------------------------
// smallest k with sigma^k >= n/2 (about one suffix per table slot; measured best on B200), bounded by the symbols
// of one packed word and by a 1 GiB table
static int auto_k(uint64_t n, uint32_t sigma, uint32_t bl) {
    int kmax = (int)(64u >> bl);
    if (kmax > 15) kmax = 15;
    int k = 1;
    double e = sigma;
    while (k < kmax && e < (double)n / 2 && e * sigma <= 268435456.0) { ++k; e *= sigma; }
    return k;
}

int drlz_build_index(drlz_ctx* c, const uint8_t* ref, uint64_t n) {
    if (!c || (!ref && n)) return DRLZ_E_ARG;
    std::lock_guard<std::recursive_mutex> lk_(c->mu);
    if (n > kMaxLen) { c->err = "reference longer than 2^32-1024"; return DRLZ_E_LIMIT; }
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    NvtxRange nvtx_index("drlz.build_index: copy, pack, suffix array, k-mer table, LCP, uniqueness array");
    c->have_index = false;
    cudaEvent_t e[5];
    for (int i = 0; i < 5; ++i) cudaEventCreate(&e[i]);
    // everything from here on is inside idx_ms[0]: the host->device copy of the reference, the alphabet scan, the
    // packing, the suffix sort, the table and the uniqueness array
    cudaEventRecord(e[0], s);
    // stage the reference and pack it with the row kernels (no gap stripping)
    uint64_t padded = ((n + 15) & ~15ull) + 16;
    CK(c, c->stage[0].ensure(padded));
    if (n) CK(c, cudaMemcpyAsync(c->stage[0].p, ref, n, cudaMemcpyHostToDevice, s));
    // alphabet: pure {A,C,G,T} -> 2-bit fast path; otherwise the distinct bytes of the reference get dense
    // order-preserving codes (suffix order under unsigned bytes is preserved), 4 bits if sigma <= 16 else 8 bits.
    // Which byte values occur is found on the device (one HBM pass; a host loop over 3.1 Gbp takes seconds).
    uint8_t lut[512];
    memset(lut, 0xFF, 256); memset(lut + 256, 0, 256);
    uint32_t sigma = 4, bl = 1;
    {
        CK(c, c->counters.ensure(sizeof(uint32_t) * (2 + kMaxRounds) > 32 ? sizeof(uint32_t) * (2 + kMaxRounds) : 32));
        CK(c, c->small_host.ensure(64));
        launch_byte_presence(c->stage[0].as<uint8_t>(), n, c->counters.as<uint32_t>(), s);
        uint32_t* hp = c->small_host.as<uint32_t>();
        CK(c, cudaMemcpyAsync(hp, c->counters.p, 32, cudaMemcpyDeviceToHost, s));
        CK(c, cudaStreamSynchronize(s));
        bool present[256];
        for (int b = 0; b < 256; ++b) present[b] = (hp[b >> 5] >> (b & 31)) & 1u;
        bool acgt = true;
        for (int b = 0; b < 256; ++b) if (present[b] && b != 'A' && b != 'C' && b != 'G' && b != 'T') acgt = false;
        if (acgt) { const char* a4 = "ACGT"; for (int q = 0; q < 4; ++q) { lut[(uint8_t)a4[q]] = (uint8_t)q; lut[256 + q] = (uint8_t)a4[q]; } }
        else {
            sigma = 0;
            for (int b = 0; b < 256; ++b) if (present[b]) { lut[b] = (uint8_t)sigma; lut[256 + sigma] = (uint8_t)b; ++sigma; }
            bl = sigma <= 16 ? 2 : 3;
            if (sigma == 256) { c->err = "reference uses all 256 byte values (0xFF is reserved)"; for (int i = 0; i < 5; ++i) cudaEventDestroy(e[i]); return DRLZ_E_ALPHABET; }
        }
    }
    uint64_t words = (n >> (6 - bl)) + 8;
    int k = c->opt_k ? c->opt_k : auto_k(n, sigma, bl);
    if (k > (int)(64u >> bl)) k = (int)(64u >> bl);
    while (k > 1 && table_entries(sigma, k) > (1ull << 30)) --k;
    CK(c, alloc_index_blob(c, n, k, sigma, bl));
    CK(c, c->tab_tmp.ensure(kmer_table_tmp_bytes(table_entries(sigma, k))));
    CK(c, cudaMemcpyAsync(c->d_lut, lut, 512, cudaMemcpyHostToDevice, s));
    CK(c, c->sa.ensure(n * 8 + 8));          // suffix array, and behind it the unsaturated uniqueness array (used only if needed)
    CK(c, c->isa.ensure(n * 4 + 4));
    c->n = 0; c->k = 1;
    uint64_t off0 = 0, m = 0, np = 0, tot = 0;
    // pack through the generic group path needs an index view only for parsing; call the pack pieces directly
    {
        uint32_t T = (uint32_t)((n + kPackTileBytes - 1) / kPackTileBytes); if (!T) T = 1;
        CK(c, c->desc_cnt.ensure((size_t)T * 8)); CK(c, c->desc_exc.ensure((size_t)T * 8)); CK(c, c->tile_counter.ensure(16));
        CK(c, c->mdev.ensure(8)); CK(c, c->exccnt.ensure(16)); CK(c, c->flags.ensure(16));
        CK(c, c->excpos.ensure(64 * 4)); CK(c, c->excbyte.ensure(64));
        CK(c, c->meta_host.ensure(64)); CK(c, c->meta_dev.ensure(64));
        CK(c, c->small_host.ensure(64));
        uint64_t* mh = c->meta_host.as<uint64_t>();
        mh[0] = 0; mh[1] = n; mh[2] = 0; mh[3] = 0;
        CK(c, cudaMemcpyAsync(c->meta_dev.p, mh, 32, cudaMemcpyHostToDevice, s));
        CK(c, cudaMemsetAsync(c->d_text, 0, words * 8, s));
        CK(c, cudaMemsetAsync(c->flags.p, 0, 16, s));
        PackJob job;
        const uint64_t* dm = c->meta_dev.as<uint64_t>();
        job.rows = c->stage[0].as<uint8_t>(); job.row_off = dm; job.cols = dm + 1; job.H = 1; job.tiles_per_row = T; job.strip = 0;
        job.desc_cnt = c->desc_cnt.as<uint64_t>(); job.desc_exc = c->desc_exc.as<uint64_t>();
        job.tile_counter = c->tile_counter.as<uint32_t>(); job.X = c->d_text; job.xoff = dm + 2;
        job.occ = c->opt_occ_pack; job.prefetch = 0; job.use_tickets = 1; job.variant = c->opt_pack_variant; job.uniform_cols = 0; job.uniform_stride = 0; job.uniform_off0 = 0;
        job.m_out = c->mdev.as<uint64_t>(); job.exc_cnt = c->exccnt.as<uint64_t>();
        job.exc_pos = c->excpos.as<uint32_t>(); job.exc_byte = c->excbyte.as<uint8_t>(); job.exc_stride = 64;
        job.flags = c->flags.as<uint32_t>(); job.gapless = nullptr; job.gapless_off = dm + 3;
        uint32_t GT = (uint32_t)((n + kPackGTileBytes - 1) / kPackGTileBytes); if (!GT) GT = 1;
        if (bl != 1) {
            CK(c, c->gtile_cnt.ensure((size_t)GT * 4)); CK(c, c->gtile_exc.ensure((size_t)GT * 4));
            CK(c, c->gtile_off.ensure((size_t)GT * 8)); CK(c, c->gtile_excoff.ensure((size_t)GT * 8));
            CK(c, c->gscan_tmp.ensure(scan_tmp_elems(GT) * 8 * 2));
        }
        job.bl = bl; job.b2c = c->d_lut; job.gtiles_per_row = GT;
        job.gtile_cnt = c->gtile_cnt.as<uint32_t>(); job.gtile_exc = c->gtile_exc.as<uint32_t>();
        job.gtile_off = c->gtile_off.as<uint64_t>(); job.gtile_excoff = c->gtile_excoff.as<uint64_t>();
        job.gscan_tmp = c->gscan_tmp.as<uint64_t>();
        launch_pack(job, s, nullptr);
        uint64_t* sh = c->small_host.as<uint64_t>();
        CK(c, cudaMemcpyAsync(sh, c->exccnt.p, 8, cudaMemcpyDeviceToHost, s));
        CK(c, cudaStreamSynchronize(s));
        if (sh[0] != 0) { c->err = "internal: reference byte without a code"; for (int i = 0; i < 5; ++i) cudaEventDestroy(e[i]); return DRLZ_E_INTERNAL; }
    }
    (void)off0; (void)m; (void)np; (void)tot;
    cudaEventRecord(e[1], s);
    char ebuf[256]; ebuf[0] = 0;
    CK(c, c->lcp.ensure(n * 4 + 8));
    int fast = 0;
    IndexView tix{c->d_text, c->sa.as<uint32_t>(), nullptr, n, k, nullptr, bl, sigma, c->d_lut + 256};
    int rc = build_suffix_array(c->d_text, n, bl, c->sa.as<uint32_t>(), c->isa.as<uint32_t>(), s, &c->sa_rounds, ebuf, sizeof ebuf, &c->sa_ws,
                                c->lcp.as<uint32_t>(), c->d_uniq, &fast, &tix, c->d_tab, c->tab_tmp.as<uint32_t>());
    c->isa_valid = !fast;
    if (c->sa_ws.bytes() > (16ull << 30)) {          // whole-genome scale: the parse needs that memory back
        cudaStreamSynchronize(s);
        c->sa_ws.release();
    }
    if (rc != 0) { c->err = std::string("suffix array: ") + ebuf; for (int i = 0; i < 5; ++i) cudaEventDestroy(e[i]); return rc == (int)cudaErrorMemoryAllocation ? DRLZ_E_NOMEM : DRLZ_E_CUDA; }
    cudaEventRecord(e[2], s);

    if (fast != 2) rc = build_kmer_table(tix, c->d_tab, c->tab_tmp.as<uint32_t>(), s);       // 2: written with the suffix array (k_sa_finish0)
    if (rc != 0) { c->err = "k-mer table build failed"; for (int i = 0; i < 5; ++i) cudaEventDestroy(e[i]); return DRLZ_E_CUDA; }
    cudaEventRecord(e[3], s);
    {   // LCP (kept only on request) and the uniqueness array derived from it
        // no doubling round was needed: every neighbouring pair of suffixes differs inside its first sort key, and the
        // LCP array came out of the sort (StoreRank); otherwise the Kasai-order kernel computes it
        // ... and when no two first keys were equal the uniqueness array came out of it as well (k_sa_finish0; values < 64)
        // a few rounds: only the pairs inside groups of equal first keys are compared on (k_lcp_ties); the bound on their
        // LCP (S0 * 2^rounds symbols) keeps that cheap.  Repetitive references (many rounds) take the Kasai-order kernel
        if (c->sa_rounds >= 1 && c->sa_rounds <= 3 && c->opt_lcp_ties) rc = build_lcp_ties(c->d_text, n, bl, c->sa.as<uint32_t>(), c->lcp.as<uint32_t>(), s);
        else rc = c->sa_rounds == 0 ? 0 : build_lcp(c->d_text, n, bl, c->sa.as<uint32_t>(), c->isa.as<uint32_t>(), c->lcp.as<uint32_t>(), s);
        if (fast) cudaMemsetAsync(c->flags.as<uint32_t>() + 2, 0, 4, s);
        else if (rc == 0) rc = build_uniq(n, c->isa.as<uint32_t>(), c->lcp.as<uint32_t>(), c->d_uniq, c->sa.as<uint32_t>() + n, c->flags.as<uint32_t>() + 2, s);
        if (rc == 0 && c->uniqc.ensure(((n >> kUniqBlockShift) + 2) * 2) != cudaSuccess) rc = 1;
        if (rc == 0) rc = build_uniq_coarse(n, c->d_uniq, c->uniqc.as<uint16_t>(), s);
        if (rc != 0) { c->err = "lcp / uniqueness build failed"; for (int i = 0; i < 5; ++i) cudaEventDestroy(e[i]); return DRLZ_E_CUDA; }
    }
    cudaEventRecord(e[4], s);
    CK(c, cudaMemcpyAsync(c->small_host.p, c->flags.as<uint32_t>() + 2, 4, cudaMemcpyDeviceToHost, s));
    CK(c, cudaStreamSynchronize(s));
    CK(c, cudaGetLastError());
    c->has_uniq32 = *c->small_host.as<uint32_t>() != 0;
    float t;
    cudaEventElapsedTime(&t, e[0], e[4]); c->idx_ms[0] = t;
    cudaEventElapsedTime(&t, e[0], e[1]); c->idx_ms[1] = t;
    cudaEventElapsedTime(&t, e[1], e[2]); c->idx_ms[2] = t;
    cudaEventElapsedTime(&t, e[2], e[3]); c->idx_ms[3] = t;
    cudaEventElapsedTime(&t, e[3], e[4]); c->idx_ms[4] = t;
    for (int i = 0; i < 5; ++i) cudaEventDestroy(e[i]);
    c->n = n; c->k = k; c->have_index = true;
    return DRLZ_OK;
}

int drlz_index_get(drlz_ctx* c, int what, void* host_out, uint64_t bytes) {
    if (!c || !host_out) return DRLZ_E_ARG;
    std::lock_guard<std::recursive_mutex> lk_(c->mu);
    if (!c->have_index) { c->err = "index not built"; return DRLZ_E_STATE; }
    CK(c, cudaSetDevice(c->device));
    const void* src = nullptr; uint64_t need = 0;
    switch (what) {
        case 0: src = c->sa.p; need = c->n * 4; break;
        case 1:
            need = c->n * 4;
            if (!c->isa_valid) CK(c, c->isa.ensure(c->n * 4 + 4));      // also on the receiving side of a broadcast
            src = c->isa.p;
            if (!c->isa_valid) { launch_isa_from_sa(c->sa.as<uint32_t>(), c->n, c->isa.as<uint32_t>(), c->stream); c->isa_valid = true; }   // the build did not need it
            break;
        case 2: if (!c->opt_lcp || !c->lcp.p) { c->err = "lcp not built (option lcp=1)"; return DRLZ_E_STATE; } src = c->lcp.p; need = c->n * 4; break;
        case 3: src = c->d_tab; need = table_entries(c->sigma, c->k) * 4; break;
        case 4: src = c->d_text; need = ((c->n + (64u >> c->bl) - 1) >> (6 - c->bl)) * 8; break;
        case 5: src = c->d_uniq; need = c->n * 2; break;
        default: return DRLZ_E_ARG;
    }
    if (bytes < need) { c->err = "host buffer too small"; return DRLZ_E_NOSPACE; }
    if (need) CK(c, cudaMemcpyAsync(host_out, src, need, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return DRLZ_OK;
}

int drlz_index_info(const drlz_ctx* c, uint64_t info[4]) {
    if (!c || !info) return DRLZ_E_ARG;
    info[0] = c->n; info[1] = (uint64_t)c->k | ((uint64_t)c->bl << 8) | ((uint64_t)c->sigma << 16) | ((uint64_t)(c->has_uniq32 ? 1 : 0) << 28);
    info[2] = (uint64_t)c->sa_rounds;
    info[3] = c->have_index ? c->blob_bytes + c->n * 4 : 0;
    return DRLZ_OK;
}

int drlz_index_timings(const drlz_ctx* c, double ms[5]) {
    if (!c || !ms) return DRLZ_E_ARG;
    for (int i = 0; i < 5; ++i) ms[i] = c->idx_ms[i];
    return DRLZ_OK;
}

int drlz_index_reserve(drlz_ctx* c, uint64_t n, int kinfo) {
    int k = kinfo & 0xFF;
    uint32_t bl = (uint32_t)(kinfo >> 8) & 0xFF, sigma = ((uint32_t)kinfo >> 16) & 0x1FFu;
    const bool has32 = ((uint32_t)kinfo >> 28) & 1u;
    if (bl == 0 && sigma == 0) { bl = 1; sigma = 4; }          // plain k: pure-ACGT index
    if (!c || k < 1 || k > 32 || bl < 1 || bl > 3 || sigma < 1 || sigma > 256) return DRLZ_E_ARG;
    std::lock_guard<std::recursive_mutex> lk_(c->mu);
    if (n > kMaxLen) return DRLZ_E_LIMIT;
    CK(c, cudaSetDevice(c->device));
    c->have_index = false;
    CK(c, alloc_index_blob(c, n, k, sigma, bl));
    CK(c, c->sa.ensure(n * 8 + 8));
    c->has_uniq32 = has32;
    c->isa_valid = false;
    CK(c, cudaMemsetAsync(c->d_text, 0, c->text_bytes, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    c->n = n; c->k = k;
    return DRLZ_OK;
}

int drlz_index_buffers(drlz_ctx* c, void* dev_ptrs[3], uint64_t bytes[3]) {
    if (!c || !dev_ptrs || !bytes) return DRLZ_E_ARG;
    std::lock_guard<std::recursive_mutex> lk_(c->mu);
    if (!c->d_text || !c->d_tab) { c->err = "no index buffers (build or reserve first)"; return DRLZ_E_STATE; }
    dev_ptrs[0] = c->d_text; bytes[0] = c->text_bytes + c->uniq_bytes + 512;   // packed text, uniqueness array, byte<->code tables
    dev_ptrs[1] = c->sa.p; bytes[1] = c->n * (c->has_uniq32 ? 8 : 4);      // suffix array (+ the unsaturated uniqueness array if in use)
    dev_ptrs[2] = c->d_tab; bytes[2] = table_entries(c->sigma, c->k) * 4;
    return DRLZ_OK;
}

int drlz_index_commit(drlz_ctx* c) {
    if (!c) return DRLZ_E_ARG;
    std::lock_guard<std::recursive_mutex> lk_(c->mu);
    if (!c->d_text || !c->d_tab) return DRLZ_E_STATE;
    // what is derived from the received buffers is made here: the block maxima of the uniqueness array.  The buffers were
    // filled by the caller on streams of its own, hence the device-wide synchronisation (once per index).
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaDeviceSynchronize());
    CK(c, c->uniqc.ensure(((c->n >> kUniqBlockShift) + 2) * 2));
    if (build_uniq_coarse(c->n, c->d_uniq, c->uniqc.as<uint16_t>(), c->stream) != 0) { c->err = "uniqueness summary failed"; return DRLZ_E_CUDA; }
    CK(c, cudaStreamSynchronize(c->stream));
    c->have_index = true;
    return DRLZ_OK;
}

// ---------------------------------------------------------------------------------------------------
// parse
// ---------------------------------------------------------------------------------------------------
int drlz_parse_batch_device(drlz_ctx* c, const uint8_t* rows_dev, const uint64_t* row_off, const uint64_t* cols, uint32_t H,
                            int strip_gaps, uint64_t* m_out, const int64_t** pairs_dev, uint64_t* n_pairs) {
    if (!c || (H && (!rows_dev || !row_off || !cols))) return DRLZ_E_ARG;
    std::lock_guard<std::recursive_mutex> lk_(c->mu);
    if (!c->have_index) { c->err = "index not built"; return DRLZ_E_STATE; }
    CK(c, cudaSetDevice(c->device));
    if (!c->opt_keep_timings) for (int i = 0; i < 12; ++i) c->parse_ms[i] = 0;
    uint64_t total = 0;
    int rc = run_device_group(c, rows_dev, row_off, cols, H, strip_gaps, false, m_out, n_pairs, &total, nullptr);
    if (rc != DRLZ_OK) return rc;
    if (pairs_dev) *pairs_dev = c->out.as<int64_t>();
    return DRLZ_OK;
}

// Gap runs of the rows of the group that run_device_group has just parsed (its row offsets and column counts are still in
// meta_dev): events on the device, sorted there, paired on the host.  Appends (column, length) records to gaps_out.
static int collect_gap_runs(drlz_ctx* c, const uint8_t* rows_dev, const uint64_t* tc, uint32_t H, int64_t* gaps_out, uint64_t gaps_cap,
                            uint64_t* done_gaps, uint64_t* n_gaps, bool* nospace) {
    cudaStream_t s = c->stream;
    uint64_t maxcols = 0, sumcols = 0;
    for (uint32_t h = 0; h < H; ++h) { n_gaps[h] = 0; sumcols += tc[h]; if (tc[h] > maxcols) maxcols = tc[h]; }
    const uint64_t* d_row_off = c->meta_dev.as<uint64_t>();
    const uint64_t* d_cols = d_row_off + H;
    CK(c, c->gap_counter.ensure(16));
    uint64_t cap = sumcols / 256 + 4096;
    if (c->gap_list.cap / 8 > cap) cap = c->gap_list.cap / 8;
    uint64_t found = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        CK(c, c->gap_list.ensure(cap * 8));
        CK(c, c->gap_host.ensure(16));
        launch_gap_events(rows_dev, d_row_off, d_cols, H, maxcols, c->gap_list.as<uint64_t>(), cap, c->gap_counter.as<unsigned long long>(), s);
        CK(c, cudaMemcpyAsync(c->gap_host.p, c->gap_counter.p, 8, cudaMemcpyDeviceToHost, s));
        CK(c, cudaStreamSynchronize(s));
        found = *c->gap_host.as<uint64_t>();
        if (found <= cap) break;
        cap = found + found / 8;             // the exact count is known now
    }
    if (found == 0) return DRLZ_OK;
    CK(c, c->gap_scratch.ensure(sort_keys_scratch_bytes(found)));
    int bits = 33;
    for (uint32_t hh = H - 1; hh; hh >>= 1) ++bits;
    if (sort_keys_u64(c->gap_list.as<uint64_t>(), found, bits, c->gap_scratch.p, s) != 0) { c->err = "gap table: sort failed"; return DRLZ_E_CUDA; }
    CK(c, c->gap_host.ensure(found * 8));
    CK(c, cudaMemcpyAsync(c->gap_host.p, c->gap_list.p, found * 8, cudaMemcpyDeviceToHost, s));
    CK(c, cudaStreamSynchronize(s));
    const uint64_t* ev = c->gap_host.as<uint64_t>();
    if (found & 1) { c->err = "gap table: odd number of run events"; return DRLZ_E_INTERNAL; }
    for (uint64_t e = 0; e < found; e += 2) {
        const uint64_t a = ev[e], b = ev[e + 1];
        if ((a & 1) != 0 || (b & 1) != 1 || (a >> 33) != (b >> 33) || b < a) { c->err = "gap table: run events do not pair up"; return DRLZ_E_INTERNAL; }
        const uint32_t h = (uint32_t)(a >> 33);
        const uint64_t start = (a >> 1) & 0xFFFFFFFFull, end = (b >> 1) & 0xFFFFFFFFull;
        if (*done_gaps < gaps_cap && gaps_out) { gaps_out[2 * *done_gaps] = (int64_t)start; gaps_out[2 * *done_gaps + 1] = (int64_t)(end - start + 1); }
        else *nospace = true;
        ++*done_gaps; ++n_gaps[h];
    }
    return DRLZ_OK;
}

static int parse_batch_impl(drlz_ctx* c, const uint8_t* const* rows, const uint64_t* cols, uint32_t H, int strip_gaps,
                            uint8_t* const* gapless_out, uint64_t* m_out, int64_t* pairs_out, uint64_t pairs_cap, uint64_t* n_pairs,
                            int64_t* gaps_out, uint64_t gaps_cap, uint64_t* n_gaps) {
    if (!c || (H && (!rows || !cols || !n_pairs))) return DRLZ_E_ARG;
    std::lock_guard<std::recursive_mutex> lk_(c->mu);
    if (!c->have_index) { c->err = "index not built"; return DRLZ_E_STATE; }
    CK(c, cudaSetDevice(c->device));
    NvtxRange nvtx_batch("drlz.parse_batch: host rows in, records out");
    // whatever way out: no copy into the caller's gapless buffers is still in flight when the call returns
    struct GlGuard { drlz_ctx* c; ~GlGuard() { if (c->gl_pending) { c->gl_pending = false; cudaStreamSynchronize(c->d2h_stream); } } } gl_guard{c};
    if (!c->opt_keep_timings) for (int i = 0; i < 12; ++i) c->parse_ms[i] = 0;
    // trimmed column counts (Spark's text reader drops the line terminator)
    std::vector<uint64_t> tc(H);
    for (uint32_t h = 0; h < H; ++h) {
        uint64_t n = cols[h];
        if (n && !rows[h]) return DRLZ_E_ARG;
        while (n > 0 && (rows[h][n - 1] == '\n' || rows[h][n - 1] == '\r')) --n;
        if (n > kMaxLen) { c->err = "row longer than 2^32-1024"; return DRLZ_E_LIMIT; }
        tc[h] = n;
    }
    // groups of consecutive rows
    std::vector<uint32_t> gstart;
    {
        uint64_t acc = 0; uint32_t cnt = 0;
        for (uint32_t h = 0; h < H; ++h) {
            if (cnt == 0) gstart.push_back(h);
            acc += tc[h]; ++cnt;
            if (acc >= c->opt_group_bytes || cnt >= 4096) { acc = 0; cnt = 0; }
        }
        gstart.push_back(H);
    }
    size_t G = gstart.size() - 1;
    std::vector<std::vector<uint64_t>> goffs(G);
    auto issue_copy = [&](size_t gi) -> int {
        uint32_t a = gstart[gi], b = gstart[gi + 1];
        uint64_t bytes = 0;
        goffs[gi].resize(b - a);
        for (uint32_t h = a; h < b; ++h) { goffs[gi][h - a] = bytes; bytes += ((tc[h] + 15) & ~15ull) + 16; }
        DevBuf& st = c->stage[gi & 1];
        CK(c, st.ensure(bytes + 16));
        // rows that follow each other in host memory with the staging layout's own spacing (a caller that fills one
        // pinned buffer, as the streaming CLI does) go over in one copy: thousands of small rows cost one call
        for (uint32_t h = a; h < b;) {
            uint32_t e = h + 1;
            while (e < b && rows[e] == rows[h] + (goffs[gi][e - a] - goffs[gi][h - a])) ++e;
            uint64_t span = goffs[gi][e - 1 - a] - goffs[gi][h - a] + tc[e - 1];
            if (span) CK(c, cudaMemcpyAsync(st.as<uint8_t>() + goffs[gi][h - a], rows[h], span, cudaMemcpyHostToDevice, c->copy_stream));
            h = e;
        }
        CK(c, cudaEventRecord(c->ev_copied[gi & 1], c->copy_stream));
        return DRLZ_OK;
    };
    uint64_t done_pairs = 0, done_gaps = 0;
    bool nospace = false;
    if (G) { int rc = issue_copy(0); if (rc) return rc; }
    std::vector<uint64_t> gl_off;
    for (size_t gi = 0; gi < G; ++gi) {
        uint32_t a = gstart[gi], b = gstart[gi + 1];
        // the other staging buffer was last read by group gi-1, whose pipeline has been synchronised already
        if (gi + 1 < G) { int rc = issue_copy(gi + 1); if (rc) return rc; }
        CK(c, cudaStreamWaitEvent(c->stream, c->ev_copied[gi & 1], 0));
        bool want_gl = false;
        if (gapless_out) for (uint32_t h = a; h < b; ++h) if (gapless_out[h]) want_gl = true;
        uint64_t total = 0;
        gl_off.resize(b - a);
        int rc = run_device_group(c, c->stage[gi & 1].as<uint8_t>(), goffs[gi].data(), tc.data() + a, b - a, strip_gaps, want_gl,
                                  m_out ? m_out + a : nullptr, n_pairs + a, &total, gl_off.data());
        if (rc != DRLZ_OK) return rc;
        if (done_pairs + total <= pairs_cap && pairs_out) {
            if (total) CK(c, cudaMemcpyAsync(pairs_out + 2 * done_pairs, c->out.p, total * 16, cudaMemcpyDeviceToHost, c->stream));
        } else nospace = true;
        if (want_gl) {
            // on their own stream: the next group's copy in, its kernels and these copies out run side by side
            uint64_t* sh_m = c->small_host.as<uint64_t>();
            const uint8_t* glsrc = (c->gl_cur ? c->gapless_b : c->gapless).as<uint8_t>();
            CK(c, cudaEventRecord(c->ev_gl_ready, c->stream));
            CK(c, cudaStreamWaitEvent(c->d2h_stream, c->ev_gl_ready, 0));
            c->gl_pending = true;
            for (uint32_t h = a; h < b; ++h)
                if (gapless_out[h] && sh_m[h - a])
                    CK(c, cudaMemcpyAsync(gapless_out[h], glsrc + gl_off[h - a], sh_m[h - a], cudaMemcpyDeviceToHost, c->d2h_stream));
            CK(c, cudaEventRecord(c->ev_gl_done[c->gl_cur], c->d2h_stream));
            c->gl_cur ^= 1;
        }
        if (n_gaps) {
            if (strip_gaps) { rc = collect_gap_runs(c, c->stage[gi & 1].as<uint8_t>(), tc.data() + a, b - a, gaps_out, gaps_cap, &done_gaps, n_gaps + a, &nospace); if (rc != DRLZ_OK) return rc; }
            else for (uint32_t h = a; h < b; ++h) n_gaps[h] = 0;
        }
        CK(c, cudaStreamSynchronize(c->stream));
        done_pairs += total;
    }
    if (c->gl_pending) { c->gl_pending = false; CK(c, cudaStreamSynchronize(c->d2h_stream)); }
    if (nospace) { c->err = "pairs_cap or gaps_cap too small"; return DRLZ_E_NOSPACE; }
    return DRLZ_OK;
}

int drlz_parse_batch(drlz_ctx* c, const uint8_t* const* rows, const uint64_t* cols, uint32_t H, int strip_gaps,
                     uint8_t* const* gapless_out, uint64_t* m_out, int64_t* pairs_out, uint64_t pairs_cap, uint64_t* n_pairs) {
    return parse_batch_impl(c, rows, cols, H, strip_gaps, gapless_out, m_out, pairs_out, pairs_cap, n_pairs, nullptr, 0, nullptr);
}

int drlz_parse_batch_gaps(drlz_ctx* c, const uint8_t* const* rows, const uint64_t* cols, uint32_t H, int strip_gaps,
                          uint8_t* const* gapless_out, uint64_t* m_out, int64_t* pairs_out, uint64_t pairs_cap, uint64_t* n_pairs,
                          int64_t* gaps_out, uint64_t gaps_cap, uint64_t* n_gaps) {
    if (!n_gaps) return DRLZ_E_ARG;
    return parse_batch_impl(c, rows, cols, H, strip_gaps, gapless_out, m_out, pairs_out, pairs_cap, n_pairs, gaps_out, gaps_cap, n_gaps);
}

int drlz_parse(drlz_ctx* c, const uint8_t* msa_row, uint64_t cols, int strip_gaps, uint8_t* gapless_out, uint64_t* m_out,
               int64_t* pairs_out, uint64_t pairs_cap, uint64_t* n_pairs) {
    if (!c || !n_pairs) return DRLZ_E_ARG;
    const uint8_t* rows[1] = {msa_row};
    uint8_t* gl[1] = {gapless_out};
    return drlz_parse_batch(c, rows, &cols, 1, strip_gaps, gapless_out ? gl : nullptr, m_out, pairs_out, pairs_cap, n_pairs);
}

// ---- async form: FIFO worker thread per context -----------------------------------------------------------
static void drlz_worker(drlz_ctx* c) {
    for (;;) {
        drlz_ctx::Job j;
        {
            std::unique_lock<std::mutex> lk(c->qmu);
            c->qcv.wait(lk, [&] { return c->stopping || !c->queue.empty(); });
            if (c->queue.empty()) return;
            j = c->queue.front(); c->queue.pop_front();
        }
        int rc = parse_batch_impl(c, j.rows, j.cols, j.H, j.strip, j.gapless, j.m_out, j.pairs, j.cap, j.n_pairs, j.gaps, j.gaps_cap, j.n_gaps);
        { std::lock_guard<std::mutex> lk(c->qmu); c->done.push_back(std::make_pair(j.ticket, rc)); }
        c->done_cv.notify_all();
    }
}

int drlz_submit_gaps(drlz_ctx* c, const uint8_t* const* rows, const uint64_t* cols, uint32_t H, int strip_gaps,
                     uint8_t* const* gapless_out, uint64_t* m_out, int64_t* pairs_out, uint64_t pairs_cap, uint64_t* n_pairs,
                     int64_t* gaps_out, uint64_t gaps_cap, uint64_t* n_gaps, uint64_t* ticket) {
    if (!c || !ticket || (H && (!rows || !cols || !n_pairs))) return DRLZ_E_ARG;
    std::lock_guard<std::mutex> lk(c->qmu);
    if (!c->worker_started) { c->worker = std::thread(drlz_worker, c); c->worker_started = true; }
    drlz_ctx::Job j{c->next_ticket++, rows, cols, H, strip_gaps, gapless_out, m_out, pairs_out, pairs_cap, n_pairs, gaps_out, gaps_cap, n_gaps};
    *ticket = j.ticket;
    c->outstanding.push_back(j.ticket);
    c->queue.push_back(j);
    c->qcv.notify_one();
    return DRLZ_OK;
}

int drlz_submit(drlz_ctx* c, const uint8_t* const* rows, const uint64_t* cols, uint32_t H, int strip_gaps,
                uint8_t* const* gapless_out, uint64_t* m_out, int64_t* pairs_out, uint64_t pairs_cap, uint64_t* n_pairs,
                uint64_t* ticket) {
    return drlz_submit_gaps(c, rows, cols, H, strip_gaps, gapless_out, m_out, pairs_out, pairs_cap, n_pairs, nullptr, 0, nullptr, ticket);
}

int drlz_wait(drlz_ctx* c, uint64_t ticket) {
    if (!c || ticket == 0) return DRLZ_E_ARG;
    std::unique_lock<std::mutex> lk(c->qmu);
    // unknown ticket, or one that was collected already (waiting for it again would never return)
    size_t oi = 0;
    while (oi < c->outstanding.size() && c->outstanding[oi] != ticket) ++oi;
    if (oi == c->outstanding.size()) return DRLZ_E_ARG;
    c->outstanding.erase(c->outstanding.begin() + oi);
    for (;;) {
        for (size_t i = 0; i < c->done.size(); ++i)
            if (c->done[i].first == ticket) { int rc = c->done[i].second; c->done.erase(c->done.begin() + i); return rc; }
        c->done_cv.wait(lk);
    }
}

int drlz_copy_to_host(drlz_ctx* c, void* dst_host, const void* src_dev, uint64_t bytes) {
    if (!c || (bytes && (!dst_host || !src_dev))) return DRLZ_E_ARG;
    std::lock_guard<std::recursive_mutex> lk_(c->mu);
    CK(c, cudaSetDevice(c->device));
    if (bytes) CK(c, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return DRLZ_OK;
}

int drlz_parse_timings(const drlz_ctx* c, double ms[12]) {
    if (!c || !ms) return DRLZ_E_ARG;
    for (int i = 0; i < 12; ++i) ms[i] = c->parse_ms[i];
    return DRLZ_OK;
}

int drlz_matching_stats(drlz_ctx* c, const uint8_t* x, uint64_t m, uint32_t* ms_out, uint32_t* pos_out) {
    if (!c || (m && (!x || !ms_out || !pos_out))) return DRLZ_E_ARG;
    std::lock_guard<std::recursive_mutex> lk_(c->mu);
    if (!c->have_index) { c->err = "index not built"; return DRLZ_E_STATE; }
    if (m > kMaxLen) return DRLZ_E_LIMIT;
    if (m == 0) return DRLZ_OK;
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    // pack x on the host side of the device: reuse the row packer with H = 1, no gap stripping
    uint64_t padded = ((m + 15) & ~15ull) + 16;
    CK(c, c->stage[0].ensure(padded));
    CK(c, cudaMemcpyAsync(c->stage[0].p, x, m, cudaMemcpyHostToDevice, s));
    uint64_t off = 0, mm = 0, np = 0, total = 0;
    // run the normal group path to get the packed haplotype (its phrase output is ignored)
    c->keep_x = true;                       // k_ms_dense below reads the packed sequence
    int rc = run_device_group(c, c->stage[0].as<uint8_t>(), &off, &m, 1, 0, false, &mm, &np, &total, nullptr);
    c->keep_x = false;
    if (rc != DRLZ_OK) return rc;
    uint64_t* sh = c->small_host.as<uint64_t>();
    CK(c, cudaMemcpyAsync(sh, c->exccnt.p, 8, cudaMemcpyDeviceToHost, s));
    CK(c, cudaStreamSynchronize(s));
    if (sh[0] != 0) { c->err = "sequence contains bytes outside the reference alphabet"; return DRLZ_E_ALPHABET; }
    CK(c, c->ms_a.ensure(m * 4)); CK(c, c->ms_b.ensure(m * 4));
    IndexView ix{c->d_text, c->sa.as<uint32_t>(), c->d_tab, c->n, c->k, nullptr, c->bl, c->sigma, c->d_lut + 256};
    launch_ms_dense(ix, c->X.as<uint64_t>(), m, c->ms_a.as<uint32_t>(), c->ms_b.as<uint32_t>(), s);
    CK(c, cudaMemcpyAsync(ms_out, c->ms_a.p, m * 4, cudaMemcpyDeviceToHost, s));
    CK(c, cudaMemcpyAsync(pos_out, c->ms_b.p, m * 4, cudaMemcpyDeviceToHost, s));
    CK(c, cudaStreamSynchronize(s));
    CK(c, cudaGetLastError());
    return DRLZ_OK;
}

}  // extern "C"
