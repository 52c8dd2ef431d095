// sa_build.cu -- suffix array of the 2-bit packed reference by radix-sort prefix doubling, the k-mer jump
// table derived from it, and the LCP array.  Replaces the external `radixSA` binary
// (DistributedRLZ.scala:45,52) and the text SA load (:55).
//
//   round 0 : key = first S0 symbols (S0 = 58 / bits per symbol: 29 bases at 2 bits) | min(remaining, S0) (6 bits)
//             -> 8-pass LSD radix sort.  The length field makes every suffix shorter than S0 unique and sorts a
//             proper prefix first, which is exactly the "shorter suffix first" order of the reference's SA.
//   round r : only suffixes whose group is not yet a singleton stay active (early retirement);
//             key = (group head index << 32) | (rank of suffix i+h, 0 past the end), radix sorted over the
//             significant bits only; positions come back through the (sorted) list of active SA slots.
//             h = S0, 2 S0, 4 S0, ...   One 8-byte device->host read per round decides termination.
//   rank[]  is the inverse suffix array once every group is a singleton.
#include <stdio.h>
#include "drlz_internal.h"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace drlz {


#define SA_CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { snprintf(err, errlen, "%s: %s", #x, cudaGetErrorString(e_)); rc = (int)e_; goto done; } } while (0)

__global__ void k_sa_init_keys(const uint64_t* __restrict__ text, uint64_t n, uint64_t* keys, uint32_t* vals, uint32_t bl) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s0 = 58u >> bl;                              // symbols in the first key
    const uint64_t keep = ~((1ull << (64 - (s0 << bl))) - 1);   // their bits (58 or 56 of them)
    uint64_t rem = n - i;
    uint64_t len = rem < (uint64_t)s0 ? rem : (uint64_t)s0;
    keys[i] = (get64(text, i, bl) & keep) | len;
    vals[i] = (uint32_t)i;
}

// head flag of sorted position j
struct LoadHeadIdx {
    const uint64_t* keys;
    __device__ __forceinline__ uint32_t operator()(uint64_t j) const {
        return (j == 0 || keys[j] != keys[j - 1]) ? (uint32_t)j : 0u;
    }
};
// ISA[sa[j]] = group head; also copies the sorted values into SA.  With lcp != null the LCP of neighbouring suffixes is
// read off their sort keys on the way (first S0 symbols + length): exact wherever the two keys differ, which is everywhere
// once no group is left for the doubling rounds (the caller knows; otherwise build_lcp recomputes the array).
struct StoreRank {
    const uint32_t* vals; uint32_t* sa; uint32_t* isa;
    const uint64_t* keys; uint32_t* lcp; uint32_t bl;
    __device__ __forceinline__ void operator()(uint64_t j, uint32_t head) const {
        uint32_t s = vals[j];
        sa[j] = s; isa[s] = head;
        if (lcp) {
            uint32_t l = 0;
            if (j) {
                const uint32_t s0 = 58u >> bl;
                const uint64_t keep = ~((1ull << (64 - (s0 << bl))) - 1);
                const uint64_t a = keys[j - 1], b = keys[j], x = (a ^ b) & keep;
                const uint32_t la = (uint32_t)(a & 63u), lb = (uint32_t)(b & 63u);
                l = x ? (uint32_t)(__clzll((long long)x) >> bl) : s0;
                l = l < la ? l : la; l = l < lb ? l : lb;
            }
            lcp[j] = l;
        }
    }
};
// active flag: position j is in a group of size > 1  (keys = sorted keys of the element list, nel elements)
struct LoadActive {
    const uint64_t* keys; uint64_t nel;
    __device__ __forceinline__ uint32_t operator()(uint64_t j) const {
        bool head = j == 0 || keys[j] != keys[j - 1];
        bool next_head = j + 1 >= nel || keys[j + 1] != keys[j];
        return (head && next_head) ? 0u : 1u;
    }
};
// compaction: active element j of the current list -> slot list of the next round
struct StoreCompact {
    const uint64_t* keys; uint64_t nel; const uint32_t* cur_slots; /* null in round 0: slot == j */ uint32_t* next_slots;
    __device__ __forceinline__ void operator()(uint64_t j, uint32_t dst) const {
        bool head = j == 0 || keys[j] != keys[j - 1];
        bool next_head = j + 1 >= nel || keys[j + 1] != keys[j];
        if (!(head && next_head)) next_slots[dst] = cur_slots ? cur_slots[j] : (uint32_t)j;
    }
};

// ---- the same table in one pass (table_variant 1) ---------------------------------------------------------------
// Keys are non-decreasing along the suffix array (a proper prefix sorts first and its zero padding is the smallest code),
// so the entries (key[j-1], key[j]] all equal j: the thread of suffix j writes them, the thread of the last suffix also
// writes n into everything above its key.  Every table entry is written exactly once -- no fill, no scattered marks, no
// scan (fill + mark + scan: 1.05 ms and 1.9 GB of traffic for a 268 MB table).  A range is a couple of entries as a rule;
// ranges above 16 entries are written by the whole warp, one after the other, and those above kTabLongRange (a symbol
// that hardly occurs, like N, leaves millions of consecutive keys empty) go to a list for k_tab_ranges.
static const uint32_t kTabLongRange = 1u << 16;
static const uint32_t kTabListCap = 1u << 15;          // a table has at most 2^30 + 1 entries: at most 2^14 ranges above 2^16
struct TabRange { uint32_t start, len, value, pad; };

__device__ __forceinline__ void tab_fill_range(uint32_t* __restrict__ tab, uint32_t start, uint32_t len, uint32_t value,
                                               TabRange* list, uint32_t* list_cnt, int lane) {
    if (len && len <= 16u) for (uint32_t t = 0; t < len; ++t) tab[start + t] = value;
    unsigned big = __ballot_sync(0xffffffffu, len > 16u);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const uint32_t st = __shfl_sync(0xffffffffu, start, src), ln = __shfl_sync(0xffffffffu, len, src),
                       v = __shfl_sync(0xffffffffu, value, src);
        if (ln > kTabLongRange) {
            uint32_t slot = 0;
            if (lane == 0) slot = atomicAdd(list_cnt, 1u);
            slot = __shfl_sync(0xffffffffu, slot, 0);
            if (slot < kTabListCap) { if (lane == 0) list[slot] = TabRange{st, ln, v, 0u}; continue; }
        }
        for (uint32_t t = (uint32_t)lane; t < ln; t += 32u) tab[st + t] = v;
    }
}

__global__ void __launch_bounds__(256) k_tab_ranges(uint32_t* __restrict__ tab, const TabRange* __restrict__ list, const uint32_t* __restrict__ list_cnt) {
    uint32_t c = *list_cnt;
    if (c > kTabListCap) c = kTabListCap;
    const uint32_t stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t e = 0; e < c; ++e) {
        const TabRange r = list[e];
        for (uint32_t t = t0; t < r.len; t += stride) tab[r.start + t] = r.value;
    }
}

// Round 0 without ties (random-like references: every suffix is told apart by its first key).  One pass over the sorted
// pairs finishes the index arrays that the two scans, the inverse permutation and k_uniq used to produce: SA, the LCP of
// neighbours read off their keys, and -- scattered through SA instead of gathered through ISA -- the uniqueness array
// uniq[SA[j]] = max(LCP[j], LCP[j+1]) (< 64, never saturated).  *ties is raised if two neighbouring keys are equal; the
// caller then takes the general path, which overwrites everything written here.
// With tab != null the k-mer table is written on the way: the table key of a suffix is the head of its sort key (k <= S0),
// so the ranges of tab_fill_range() come from the two keys the thread holds anyway -- no pass over SA, no reads of the text.
__global__ void __launch_bounds__(256) k_sa_finish0(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint64_t n,
                                                    uint32_t bl, uint32_t* __restrict__ sa, uint32_t* __restrict__ lcp,
                                                    uint16_t* __restrict__ uniq, uint32_t* ties,
                                                    IndexView tix, uint32_t* __restrict__ tab, uint32_t tab_cnt, TabRange* list, uint32_t* list_cnt) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool tie = false;
    if (tab) {
        const uint32_t s0 = 58u >> bl;
        const uint64_t keep = ~((1ull << (64 - (s0 << bl))) - 1);
        uint64_t key = 0, prev = 0, span;
        const bool valid = j < n;
        if (valid) {
            prefix_keys(tix, keys[j] & keep, (uint32_t)tix.k, &key, &span);
            if (j) prefix_keys(tix, keys[j - 1] & keep, (uint32_t)tix.k, &prev, &span);
        }
        uint32_t start = 0, len = 0;
        if (valid && prev <= key) { start = j ? (uint32_t)prev + 1u : 0u; len = (uint32_t)key + 1u - start; }
        tab_fill_range(tab, start, len, (uint32_t)j, list, list_cnt, threadIdx.x & 31);
        const bool last = valid && j + 1 == n;
        start = last ? (uint32_t)key + 1u : 0u; len = last ? tab_cnt - start : 0u;
        tab_fill_range(tab, start, len, (uint32_t)n, list, list_cnt, threadIdx.x & 31);
    }
    if (j < n) {
        const uint32_t s0 = 58u >> bl;
        const uint64_t keep = ~((1ull << (64 - (s0 << bl))) - 1);
        const uint64_t b = keys[j];
        const uint32_t lb = (uint32_t)(b & 63u);
        uint32_t la_ = 0, ln_ = 0;                      // LCP with the suffix in front, with the one behind
        if (j) {
            const uint64_t a = keys[j - 1], x = (a ^ b) & keep;
            const uint32_t la = (uint32_t)(a & 63u);
            tie = a == b;
            uint32_t l = x ? (uint32_t)(__clzll((long long)x) >> bl) : s0;
            l = l < la ? l : la; la_ = l < lb ? l : lb;
        }
        if (j + 1 < n) {
            const uint64_t c = keys[j + 1], x = (c ^ b) & keep;
            const uint32_t lc = (uint32_t)(c & 63u);
            uint32_t l = x ? (uint32_t)(__clzll((long long)x) >> bl) : s0;
            l = l < lc ? l : lc; ln_ = l < lb ? l : lb;
        }
        const uint32_t s = vals[j];
        sa[j] = s;
        if (lcp) lcp[j] = la_;
        uniq[s] = (uint16_t)(la_ > ln_ ? la_ : ln_);
    }
    if (__syncthreads_or(tie) && threadIdx.x == 0) *ties = 1u;
}
__global__ void k_isa_from_sa(const uint32_t* __restrict__ sa, uint64_t n, uint32_t* __restrict__ isa) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) isa[sa[j]] = (uint32_t)j;
}
void launch_isa_from_sa(const uint32_t* sa, uint64_t n, uint32_t* isa, cudaStream_t s) {
    if (!n) return;
    g_launches += 1;
    k_isa_from_sa<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(sa, n, isa);
}

__global__ void k_sa_round_keys(const uint32_t* __restrict__ slots, uint64_t na, const uint32_t* __restrict__ sa,
                                const uint32_t* __restrict__ isa, uint64_t n, uint64_t h, uint64_t* keys, uint32_t* vals) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= na) return;
    uint32_t j = slots[t];
    uint32_t s = sa[j];
    uint64_t g = isa[s];
    uint64_t p = (uint64_t)s + h;
    uint64_t r2 = p < n ? (uint64_t)isa[p] + 1 : 0;
    keys[t] = (g << 32) | r2;
    vals[t] = s;
}

// new group head (as an index t into the sorted active list) by running maximum
struct LoadHeadT {
    const uint64_t* keys;
    __device__ __forceinline__ uint32_t operator()(uint64_t t) const {
        return (t == 0 || keys[t] != keys[t - 1]) ? (uint32_t)t : 0u;
    }
};
struct StoreRound {
    const uint32_t* slots; const uint32_t* vals; uint32_t* sa; uint32_t* isa;
    __device__ __forceinline__ void operator()(uint64_t t, uint32_t head_t) const {
        uint32_t s = vals[t];
        sa[slots[t]] = s;
        isa[s] = slots[head_t];
    }
};

static int bits_for(uint64_t v) { int b = 0; while (v) { ++b; v >>= 1; } return b ? b : 1; }

int build_suffix_array(const uint64_t* text, uint64_t n, uint32_t bl, uint32_t* sa, uint32_t* isa_out, cudaStream_t s,
                       int* rounds_out, char* err, int errlen, SaWorkspace* ws_in, uint32_t* lcp_from_keys, uint16_t* uniq_fast,
                       int* fast_out, const IndexView* tab_ix, uint32_t* tab, uint32_t* tab_tmp) {
    int rc = 0;
    if (rounds_out) *rounds_out = 0;
    if (fast_out) *fast_out = 0;
    if (n == 0) return 0;
    SaWorkspace local;
    SaWorkspace& ws = ws_in ? *ws_in : local;
    uint64_t *keys_a = nullptr, *keys_b = nullptr, *tot_host = nullptr;
    uint32_t *vals_a = nullptr, *vals_b = nullptr, *hist = nullptr, *scan_tmp = nullptr, *slots_a = nullptr,
             *slots_b = nullptr, *isa = nullptr;
    uint64_t hist_n = sort_hist_elems(n);
    uint64_t tmp_n = scan_tmp_elems(hist_n > n ? hist_n : n) + 8;
    SA_CK(ws.ensure(0, n * 8));
    SA_CK(ws.ensure(1, n * 8));
    SA_CK(ws.ensure(2, n * 4));
    SA_CK(ws.ensure(3, n * 4));
    SA_CK(ws.ensure(4, hist_n * 4 + 64));
    SA_CK(ws.ensure(5, tmp_n * 4));
    SA_CK(ws.ensure(6, n * 4));
    SA_CK(ws.ensure(7, n * 4));
    if (!isa_out) SA_CK(ws.ensure(8, n * 4));
    if (!ws.host) SA_CK(cudaMallocHost(&ws.host, 8));
    keys_a = (uint64_t*)ws.p[0]; keys_b = (uint64_t*)ws.p[1]; vals_a = (uint32_t*)ws.p[2]; vals_b = (uint32_t*)ws.p[3];
    hist = (uint32_t*)ws.p[4]; scan_tmp = (uint32_t*)ws.p[5]; slots_a = (uint32_t*)ws.p[6]; slots_b = (uint32_t*)ws.p[7];
    isa = isa_out ? isa_out : (uint32_t*)ws.p[8];
    tot_host = (uint64_t*)ws.host;
    {
        const unsigned T = 256;
        unsigned nb = (unsigned)((n + T - 1) / T);
        g_launches += 1;
        k_sa_init_keys<<<nb, T, 0, s>>>(text, n, keys_a, vals_a, bl);
        radix_sort_pairs<uint64_t>(&keys_a, &keys_b, &vals_a, &vals_b, n, 0, 64, hist, scan_tmp, s);
        if (uniq_fast && fast_out) {
            // no two equal keys (the usual outcome on a random-like reference): SA, LCP and the uniqueness array in one pass,
            // no inverse permutation (api.cu derives it from SA if somebody asks for it)
            uint32_t* ties = scan_tmp;
            SA_CK(cudaMemsetAsync(ties, 0, 4, s));
            // the k-mer table in the same pass when its key is a head of the sort key (k symbols <= S0) and the caller wants it
            uint64_t tab_cnt = 0;
            IndexView tix{};
            bool with_tab = tab_ix && tab && tab_tmp && g_table_variant == 1 && (uint32_t)tab_ix->k <= (58u >> bl);
            if (with_tab) {
                tix = *tab_ix;
                tab_cnt = 1;
                for (int t = 0; t < tix.k; ++t) tab_cnt *= tix.sigma;
                tab_cnt += 1;
                if (tab_cnt > 0xFFFFFFFFull) with_tab = false;
            }
            uint32_t* list_cnt = with_tab ? tab_tmp : nullptr;
            TabRange* list = with_tab ? reinterpret_cast<TabRange*>(reinterpret_cast<uint8_t*>(tab_tmp) + 64) : nullptr;
            if (with_tab) SA_CK(cudaMemsetAsync(list_cnt, 0, 4, s));
            g_launches += 1;
            k_sa_finish0<<<nb, T, 0, s>>>(keys_a, vals_a, n, bl, sa, lcp_from_keys, uniq_fast, ties, tix, with_tab ? tab : nullptr, (uint32_t)tab_cnt,
                                          list, list_cnt);
            if (with_tab) { g_launches += 1; k_tab_ranges<<<592, 256, 0, s>>>(tab, list, list_cnt); }
            SA_CK(cudaMemcpyAsync(tot_host, ties, 4, cudaMemcpyDeviceToHost, s));
            SA_CK(cudaStreamSynchronize(s));
            if (*(uint32_t*)tot_host == 0) { *fast_out = with_tab ? 2 : 1; SA_CK(cudaGetLastError()); goto done; }
        }
        // ranks = group head index (inclusive running max of head indices); SA = sorted values
        device_scan<uint32_t, LoadHeadIdx, StoreRank, OpMax, true>(LoadHeadIdx{keys_a}, StoreRank{vals_a, sa, isa, keys_a, lcp_from_keys, bl}, n,
                                                                   scan_tmp, OpMax(), 0u, s);
        // active slots
        device_scan<uint32_t, LoadActive, StoreCompact, OpSum, false>(
            LoadActive{keys_a, n}, StoreCompact{keys_a, n, nullptr, slots_a}, n, scan_tmp, OpSum(), 0u, s);
        uint64_t ntile = (n + kScanTile - 1) / kScanTile;
        uint32_t na32 = 0;
        SA_CK(cudaMemcpyAsync(tot_host, scan_tmp + ntile, 4, cudaMemcpyDeviceToHost, s));
        SA_CK(cudaStreamSynchronize(s));
        na32 = *(uint32_t*)tot_host;
        uint64_t na = na32;
        uint64_t h = 58u >> bl;
        int rounds = 0;
        int rank_bits = bits_for(n);        // ranks+1 <= n
        while (na > 0) {
            ++rounds;
            unsigned nba = (unsigned)((na + T - 1) / T);
            g_launches += 1;
            k_sa_round_keys<<<nba, T, 0, s>>>(slots_a, na, sa, isa, n, h, keys_a, vals_a);
            // low half: rank of i+h; high half: group head.  Only the significant bytes are sorted.
            radix_sort_pairs<uint64_t>(&keys_a, &keys_b, &vals_a, &vals_b, na, 0, rank_bits, hist, scan_tmp, s);
            radix_sort_pairs<uint64_t>(&keys_a, &keys_b, &vals_a, &vals_b, na, 32, 32 + rank_bits, hist, scan_tmp, s);
            device_scan<uint32_t, LoadHeadT, StoreRound, OpMax, true>(LoadHeadT{keys_a}, StoreRound{slots_a, vals_a, sa, isa},
                                                                     na, scan_tmp, OpMax(), 0u, s);
            device_scan<uint32_t, LoadActive, StoreCompact, OpSum, false>(
                LoadActive{keys_a, na}, StoreCompact{keys_a, na, slots_a, slots_b}, na, scan_tmp, OpSum(), 0u, s);
            uint64_t nt = (na + kScanTile - 1) / kScanTile;
            SA_CK(cudaMemcpyAsync(tot_host, scan_tmp + nt, 4, cudaMemcpyDeviceToHost, s));
            SA_CK(cudaStreamSynchronize(s));
            na = *(uint32_t*)tot_host;
            uint32_t* tsl = slots_a; slots_a = slots_b; slots_b = tsl;
            h *= 2;
            if (rounds > 40) { snprintf(err, errlen, "prefix doubling did not converge"); rc = -1; goto done; }
        }
        if (rounds_out) *rounds_out = rounds;
        SA_CK(cudaGetLastError());
    }
done:
    if (!ws_in) { cudaStreamSynchronize(s); local.release(); }
    return rc;
}

// ---- plain key sort for other parts of the library (the gap-run events of pack.cu) ------------------------------------
// scratch layout: keys_b[n] | vals_a[n] | vals_b[n] | hist | scan_tmp
size_t sort_keys_scratch_bytes(uint64_t n) {
    uint64_t hist_n = sort_hist_elems(n);
    return (size_t)(n * 8 + n * 4 * 2 + hist_n * 4 + 64 + (scan_tmp_elems(hist_n) + 8) * 4 + 256);
}
int sort_keys_u64(uint64_t* keys, uint64_t n, int end_bit, void* scratch, cudaStream_t s) {
    if (n <= 1) return 0;
    uint64_t hist_n = sort_hist_elems(n);
    uint8_t* p = (uint8_t*)scratch;
    uint64_t* keys_b = (uint64_t*)p; p += n * 8;
    uint32_t* vals_a = (uint32_t*)p; p += n * 4;
    uint32_t* vals_b = (uint32_t*)p; p += n * 4;
    uint32_t* hist = (uint32_t*)p; p += hist_n * 4 + 64;
    uint32_t* tmp = (uint32_t*)p;
    uint64_t* ka = keys; uint64_t* kb = keys_b;
    int passes = radix_sort_pairs<uint64_t>(&ka, &kb, &vals_a, &vals_b, n, 0, end_bit, hist, tmp, s);
    if (passes & 1) cudaMemcpyAsync(keys, ka, n * 8, cudaMemcpyDeviceToDevice, s);       // result sits in the scratch half
    return (int)cudaGetLastError();
}

// ---- k-mer jump table ---------------------------------------------------------------------------
__global__ void k_tab_fill(uint32_t* tab, uint64_t cnt, uint32_t v) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cnt) tab[i] = v;
}
__global__ void k_tab_mark(IndexView ix, uint32_t* tab) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ix.n) return;
    uint64_t key, span, prev = 0;
    prefix_keys(ix, get64(ix.text, ix.sa[j], ix.bl), (uint32_t)ix.k, &key, &span);
    if (j) prefix_keys(ix, get64(ix.text, ix.sa[j - 1], ix.bl), (uint32_t)ix.k, &prev, &span);
    if (j == 0 || prev != key) tab[key] = (uint32_t)j;
}
// right-to-left running minimum: element e of the scan is table entry (cnt-1-e)
struct LoadRev { const uint32_t* tab; uint64_t cnt; __device__ __forceinline__ uint32_t operator()(uint64_t e) const { return tab[cnt - 1 - e]; } };
struct StoreRev { uint32_t* tab; uint64_t cnt; __device__ __forceinline__ void operator()(uint64_t e, uint32_t v) const { tab[cnt - 1 - e] = v; } };

__global__ void __launch_bounds__(256) k_tab_build(IndexView ix, uint32_t* __restrict__ tab, uint32_t cnt, TabRange* list, uint32_t* list_cnt) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = j < ix.n;
    uint64_t key = 0, prev = 0, span;
    if (valid) {
        prefix_keys(ix, get64(ix.text, ix.sa[j], ix.bl), (uint32_t)ix.k, &key, &span);
        if (j) prefix_keys(ix, get64(ix.text, ix.sa[j - 1], ix.bl), (uint32_t)ix.k, &prev, &span);
    }
    // (prev, key] <- j ; for the first suffix [0, key] <- 0
    uint32_t start = 0, len = 0;
    if (valid && prev <= key) { start = j ? (uint32_t)prev + 1u : 0u; len = (uint32_t)key + 1u - start; }
    tab_fill_range(tab, start, len, (uint32_t)j, list, list_cnt, lane);
    // (key of the last suffix, cnt) <- n
    const bool last = valid && j + 1 == ix.n;
    start = last ? (uint32_t)key + 1u : 0u; len = last ? cnt - start : 0u;
    tab_fill_range(tab, start, len, (uint32_t)ix.n, list, list_cnt, lane);
}

size_t kmer_table_tmp_bytes(uint64_t entries) {
    const size_t scan = scan_tmp_elems(entries) * 4 + 64, list = (size_t)kTabListCap * sizeof(TabRange) + 64;
    return scan > list ? scan : list;
}

int g_table_variant = 1;                   // 1: k_tab_build (one pass), 0: fill + mark + reverse minimum scan (option table_variant; process-wide)

int build_kmer_table(const IndexView& ix, uint32_t* tab, uint32_t* tmp, cudaStream_t s) {
    uint64_t cnt = 1;
    for (int t = 0; t < ix.k; ++t) cnt *= ix.sigma;
    cnt += 1;
    if (g_table_variant == 1 && cnt <= 0xFFFFFFFFull) {
        if (ix.n == 0) { cudaMemsetAsync(tab, 0, cnt * 4, s); return (int)cudaGetLastError(); }
        uint32_t* list_cnt = tmp;                                   // [0]: number of listed ranges, the list 64 bytes behind it
        TabRange* list = reinterpret_cast<TabRange*>(reinterpret_cast<uint8_t*>(tmp) + 64);
        cudaMemsetAsync(list_cnt, 0, 4, s);
        g_launches += 2;
        k_tab_build<<<(unsigned)((ix.n + 255) / 256), 256, 0, s>>>(ix, tab, (uint32_t)cnt, list, list_cnt);
        k_tab_ranges<<<592, 256, 0, s>>>(tab, list, list_cnt);
        return (int)cudaGetLastError();
    }
    g_launches += 2;
    k_tab_fill<<<(unsigned)((cnt + 255) / 256), 256, 0, s>>>(tab, cnt, (uint32_t)ix.n);
    if (ix.n) k_tab_mark<<<(unsigned)((ix.n + 255) / 256), 256, 0, s>>>(ix, tab);
    device_scan<uint32_t, LoadRev, StoreRev, OpMin, true>(LoadRev{tab, cnt}, StoreRev{tab, cnt}, cnt, tmp, OpMin(),
                                                          0xFFFFFFFFu, s);
    return (int)cudaGetLastError();
}

// ---- LCP (Kasai order inside fixed blocks of text positions) ----------------------------------------
static const int kLcpBlock = 64;
__global__ void k_lcp(const uint64_t* __restrict__ text, uint64_t n, const uint32_t* __restrict__ sa,
                      const uint32_t* __restrict__ isa, uint32_t* __restrict__ lcp, uint32_t bl) {
    uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t i0 = b * kLcpBlock;
    if (i0 >= n) return;
    uint64_t i1 = i0 + kLcpBlock < n ? i0 + kLcpBlock : n;
    uint64_t hh = 0;
    for (uint64_t i = i0; i < i1; ++i) {
        uint32_t r = isa[i];
        if (r == 0) { lcp[0] = 0; hh = 0; continue; }
        uint64_t j = sa[r - 1];
        // extend from hh (Kasai: lcp of the successor suffix is at least lcp-1)
        uint64_t lim = n - (i > j ? i : j);
        while (hh < lim) {
            uint64_t xr = get64(text, i + hh, bl) ^ get64(text, j + hh, bl);
            if (xr) { hh += (uint64_t)(clz64(xr) >> bl); break; }
            hh += 64u >> bl;
        }
        if (hh > lim) hh = lim;
        lcp[r] = (uint32_t)hh;
        if (hh) --hh;
    }
}

// After a few doubling rounds (a random-like reference at whole-genome scale: a handful of equal first keys) the LCP array
// read off the first-round sort keys (StoreRank) is already exact wherever the two keys differ.  That includes every pair
// across a group border, whatever the final order inside the groups, because all members of a group carry the same key.
// Pairs inside a group of equal keys share their first S0 symbols and were written as exactly S0: those, and only those,
// are compared on from symbol S0.  The caller uses this only when the rounds bound the answer (rounds r => LCP < S0 * 2^r).
__global__ void k_lcp_ties(const uint64_t* __restrict__ text, uint64_t n, const uint32_t* __restrict__ sa, uint32_t* __restrict__ lcp, uint32_t bl) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0 || j >= n) return;
    const uint32_t s0 = 58u >> bl;
    if (lcp[j] != s0) return;
    const uint64_t a = sa[j - 1], b = sa[j];
    const uint64_t lim = n - (a > b ? a : b);
    uint64_t hh = s0;
    while (hh < lim) {
        const uint64_t xr = get64(text, a + hh, bl) ^ get64(text, b + hh, bl);
        if (xr) { hh += (uint64_t)(clz64(xr) >> bl); break; }
        hh += 64u >> bl;
    }
    lcp[j] = (uint32_t)(hh > lim ? lim : hh);
}
int build_lcp_ties(const uint64_t* text, uint64_t n, uint32_t bl, const uint32_t* sa, uint32_t* lcp, cudaStream_t s) {
    if (n == 0) return 0;
    g_launches += 1;
    k_lcp_ties<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(text, n, sa, lcp, bl);
    return (int)cudaGetLastError();
}

int build_lcp(const uint64_t* text, uint64_t n, uint32_t bl, const uint32_t* sa, const uint32_t* isa, uint32_t* lcp, cudaStream_t s) {
    if (n == 0) return 0;
    uint64_t nb = (n + kLcpBlock - 1) / kLcpBlock;
    g_launches += 1;
    k_lcp<<<(unsigned)((nb + 127) / 128), 128, 0, s>>>(text, n, sa, isa, lcp, bl);
    return (int)cudaGetLastError();
}

}  // namespace drlz

namespace drlz {
// ---- uniqueness array: longest repeat starting at each text position --------------------------------------------
// uniq[p] = min(65535, max(LCP[ISA[p]], LCP[ISA[p]+1])): a match of more than uniq[p] bases that starts at reference
// position p cannot be matched as long anywhere else (used by the predicted-diagonal path of the parse).
__global__ void k_uniq(uint64_t n, const uint32_t* __restrict__ isa, const uint32_t* __restrict__ lcp, uint16_t* __restrict__ uniq,
                       uint32_t* __restrict__ uniq32, uint32_t* saturated) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint32_t r = isa[p];
    uint32_t a = lcp[r];
    uint32_t b = (uint64_t)r + 1 < n ? lcp[r + 1] : 0u;
    uint32_t u = a > b ? a : b;
    uniq[p] = (uint16_t)(u > 65535u ? 65535u : u);
    uniq32[p] = u;
    if (u >= 65535u) *saturated = 1u;                       // (every writer stores the same value)
}

int build_uniq(uint64_t n, const uint32_t* isa, const uint32_t* lcp, uint16_t* uniq, uint32_t* uniq32, uint32_t* saturated, cudaStream_t s) {
    cudaMemsetAsync(saturated, 0, 4, s);
    if (n == 0) return 0;
    g_launches += 1;
    k_uniq<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, isa, lcp, uniq, uniq32, saturated);
    return (int)cudaGetLastError();
}

// uniqc[j] = max of uniq over positions [64 j, 64 j + 64): the parse tests c > uniq[p] against it first (rlz_core.cuh)
__global__ void k_uniq_coarse(uint64_t n, const uint16_t* __restrict__ uniq, uint16_t* __restrict__ uniqc) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t a = j << kUniqBlockShift;
    if (a >= n) return;
    uint32_t mx = 0;
    if (a + kUniqBlock <= n) {
        const uint4* v = reinterpret_cast<const uint4*>(uniq + a);        // 128 bytes, 16-byte aligned (the array is 256-byte aligned)
#pragma unroll
        for (int t = 0; t < (int)(kUniqBlock / 8); ++t) {
            const uint4 q = v[t];
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int z = 0; z < 4; ++z) { const uint32_t lo = w[z] & 0xFFFFu, hi = w[z] >> 16; mx = mx > lo ? mx : lo; mx = mx > hi ? mx : hi; }
        }
    } else for (uint64_t q = a; q < n; ++q) mx = mx > uniq[q] ? mx : uniq[q];
    uniqc[j] = (uint16_t)mx;
}

int build_uniq_coarse(uint64_t n, const uint16_t* uniq, uint16_t* uniqc, cudaStream_t s) {
    if (n == 0) return 0;
    const uint64_t nb = (n + kUniqBlock - 1) >> kUniqBlockShift;
    g_launches += 1;
    k_uniq_coarse<<<(unsigned)((nb + 255) / 256), 256, 0, s>>>(n, uniq, uniqc);
    return (int)cudaGetLastError();
}

}  // namespace drlz

// ---- debug/test entry points (not part of include/drlz.h): exercise the primitives in isolation ----
extern "C" int drlz_debug_sort_pairs(uint64_t* keys_host, uint32_t* vals_host, uint64_t n, int begin_bit, int end_bit) {
    using namespace drlz;
    if (n == 0) return 0;
    uint64_t *ka = nullptr, *kb = nullptr; uint32_t *va = nullptr, *vb = nullptr, *hist = nullptr, *tmp = nullptr;
    uint64_t hist_n = sort_hist_elems(n);
    if (cudaMalloc(&ka, n * 8) || cudaMalloc(&kb, n * 8) || cudaMalloc(&va, n * 4) || cudaMalloc(&vb, n * 4) ||
        cudaMalloc(&hist, hist_n * 4 + 64) || cudaMalloc(&tmp, (scan_tmp_elems(hist_n) + 8) * 4)) return -1;
    cudaMemcpy(ka, keys_host, n * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(va, vals_host, n * 4, cudaMemcpyHostToDevice);
    radix_sort_pairs<uint64_t>(&ka, &kb, &va, &vb, n, begin_bit, end_bit, hist, tmp, 0);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(keys_host, ka, n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(vals_host, va, n * 4, cudaMemcpyDeviceToHost);
    cudaFree(ka); cudaFree(kb); cudaFree(va); cudaFree(vb); cudaFree(hist); cudaFree(tmp);
    return (int)e;
}

extern "C" int drlz_debug_scan_u32(const uint32_t* in_host, uint64_t* out_host, uint64_t n, uint64_t* total) {
    using namespace drlz;
    uint32_t* din = nullptr; uint64_t *dout = nullptr, *tmp = nullptr;
    if (cudaMalloc(&din, n * 4 + 4) || cudaMalloc(&dout, n * 8 + 8) || cudaMalloc(&tmp, scan_tmp_elems(n) * 8)) return -1;
    cudaMemcpy(din, in_host, n * 4, cudaMemcpyHostToDevice);
    device_scan<uint64_t, LoadArr<uint32_t, uint64_t>, StoreArr<uint64_t>, OpSum, false>(
        LoadArr<uint32_t, uint64_t>{din}, StoreArr<uint64_t>{dout}, n, tmp, OpSum(), 0ull, 0);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(out_host, dout, n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(total, tmp + (n + kScanTile - 1) / kScanTile, 8, cudaMemcpyDeviceToHost);
    cudaFree(din); cudaFree(dout); cudaFree(tmp);
    return (int)e;
}
