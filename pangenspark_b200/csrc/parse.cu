// parse.cu -- kernels of the chunk-parallel exact greedy RLZ parse (per-thread logic in rlz_core.cuh).
//
//   k_hint      H : a 64-base lookup at every chunk start gives the chunk's predicted diagonal
//   k_diag_w    D : mismatch offsets of every chunk along that diagonal (the streaming compare, done once, up front)
//   k_spec      P1: one thread per chunk, speculative greedy parse from the chunk's first base with chunk-capped
//                   lookups; phrases stored; the first phrase doubles as stage A's phrase at the chunk start
//   k_link      A : open first phrases linked to the chunk their match continues in
//   scan/k_rank A : list ranking of the diagonal runs (a right-to-left segmented sum; pointer jumping when chunks
//                   are smaller than min_cap) -> exact lengths of arbitrarily long matches
//   k_resolve   A : total length of each chunk's open last phrase -> exit of the chunk
//   k_fix       F : one thread per (chunk, alternative entry) created in the previous round
//   k_mark_*    J : records form a functional graph; the true chain is marked in two levels (shared-memory pointer
//                   jumping per block of 1024 chunks + one hop per block), then each chunk's true record is picked
//   device_scan E : exclusive prefix of the counts = output slot of every chunk (stream compaction)
//   k_emit      E : gather the stored phrases of the true records into (int64 pos, int64 len) records
//                   (DistributedRLZ.scala:263-284 record layout); re-parses only what was not stored
//
// Two warp-synchronous variants of k_spec (a lock-step per-lane state machine, and cooperative 1024-base
// extension) were built and measured slower (2.7 ms / 1.6 ms vs 1.2 ms per step): the parse is bound by chains
// of dependent cache misses, and divergent scalar lanes overlap those chains better (profiles/README.md).  So was
// a kernel that did the lookups at the listed mismatch positions up front, one thread each (0.35 ms for 2.3 M
// lookups, k_spec only 0.08 ms faster).
#include "drlz_internal.h"
#include "scan.cuh"

namespace drlz {

// The symbol width is a run-time field of the index (2, 4 or 8 bits), but nearly every batch runs on a pure ACGT
// index: the hot kernels carry a copy of their body in which the width is a compile-time constant (shifts and
// masks fold, the generic key arithmetic disappears) and fall back to the run-time body otherwise.
#define DRLZ_WITH_BL(pv, PV, ...)                                                               \
    do {                                                                                         \
        if ((pv).ix.bl == 1) { ParseView PV = (pv); PV.ix.bl = 1; PV.ix.sigma = 4; __VA_ARGS__; } \
        else { const ParseView& PV = (pv); __VA_ARGS__; }                                        \
    } while (0)

__global__ void __launch_bounds__(128) k_link(ParseView pv) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < pv.n_chunks) DRLZ_WITH_BL(pv, q, link_chunk(q, g));
}

__global__ void k_rank(const uint32_t* __restrict__ link_in, const uint32_t* __restrict__ run_in, uint32_t* link_out,
                       uint32_t* run_out, uint32_t nc) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < nc) rank_round(link_in, run_in, link_out, run_out, g);
}

__global__ void __launch_bounds__(128) k_hint(ParseView pv, int64_t* hint, uint32_t nq) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) DRLZ_WITH_BL(pv, v, hint_group(v, hint, q));
}

// H, second part: a chunk whose probes found nothing (it starts inside a run of N, or in a long repeat, where no
// 64-symbol pattern is unique) takes the diagonal of the nearest chunk in front of it in the same haplotype -- inside a
// run of millions of N the alignment does not change.  "Nearest valid chunk in front" = an inclusive running maximum over
// (g + 1 if the chunk has a hint of its own, else 0).
struct LoadHintIdx {
    const int64_t* hint;
    __device__ __forceinline__ uint32_t operator()(uint64_t g) const { return hint[g] != kNoDiag ? (uint32_t)g + 1u : 0u; }
};
struct StoreHintSrc { uint32_t* src; __device__ __forceinline__ void operator()(uint64_t g, uint32_t v) const { src[g] = v; } };
__global__ void k_hint_fill(ParseView pv, int64_t* hint, const uint32_t* __restrict__ src) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= pv.n_chunks || hint[g] != kNoDiag) return;
    const uint32_t f = src[g];                              // 1 + the nearest chunk with a hint (in front of g, or behind it), 0 = none
    // (reads hint[f - 1], which this kernel never writes: chunks that have a hint return above)
    if (f && f <= pv.n_chunks && pv.chunk_hap[f - 1] == pv.chunk_hap[g]) hint[g] = hint[f - 1];
}
// the same from the other side, for the chunks at the head of a haplotype (a chromosome that begins with a run of N has no
// chunk in front of it to ask): a right-to-left running minimum over (g + 1 if the chunk has a hint, else "none")
struct LoadHintIdxRev {
    const int64_t* hint; uint64_t nc;
    __device__ __forceinline__ uint32_t operator()(uint64_t e) const { const uint64_t g = nc - 1 - e; return hint[g] != kNoDiag ? (uint32_t)g + 1u : 0xFFFFFFFFu; }
};
struct StoreHintSrcRev { uint32_t* src; uint64_t nc; __device__ __forceinline__ void operator()(uint64_t e, uint32_t v) const { src[nc - 1 - e] = v; } };

// ext[]: how far the match along a chunk's diagonal goes, counted through the chunks behind it (rlz_core.cuh: ext_term):
// a right-to-left segmented sum like the one that ranks the diagonal runs of stage A, with one more sticky bit: a count
// that ends for no reason (the next chunk was hinted another diagonal) is only a lower bound.
// element: bits 0..39 symbols, bit 62 inexact, bit 63 "the count ends with this chunk"
struct OpExtSum {
    __device__ __forceinline__ uint64_t operator()(uint64_t a, uint64_t b) const {
        const uint64_t F = 1ull << 63, X = 1ull << 62, M = (1ull << 40) - 1;
        if (b & F) return b;
        return (a & F) | ((a | b) & X) | (((a & M) + (b & M)) & M);
    }
};
struct LoadExtRev {
    ParseView pv; uint64_t nc;
    __device__ __forceinline__ uint64_t operator()(uint64_t e) const {
        bool on, exact;
        const uint32_t v = ext_term(pv, (uint32_t)(nc - 1 - e), &on, &exact);
        return (uint64_t)v | (on ? 0ull : (1ull << 63)) | ((on || exact) ? 0ull : (1ull << 62));
    }
};
struct StoreExtRev {
    uint32_t* out; uint64_t nc;
    __device__ __forceinline__ void operator()(uint64_t e, uint64_t v) const {
        const uint64_t val = v & ((1ull << 40) - 1);
        const bool inexact = (v >> 62) & 1ull;
        out[nc - 1 - e] = val >= 0x7FFFFFFFull ? (0x7FFFFFFFu | kExtInexact) : ((uint32_t)val | (inexact ? kExtInexact : 0u));
    }
};

// D: one warp per chunk; lane l compares packed word l of each 1024-base slab (the haplotype side is word aligned
// because chunks start at multiples of 32 bases), so both streams are read with coalesced 256-byte requests at
// 32 active lanes -- this is the whole streaming compare of the parse, done once.  Mismatching words are decoded
// in lane order (ballot), which keeps the offsets sorted.  Same result as diag_scan_chunk().
//
// The compare is bound by instruction issue, so the loop over complete slabs carries nothing but the loads, the
// funnel shift and the XOR: two running pointers, the half of the funnel fixed at compile time (ODD), no bounds or
// tail-mask logic (that lives in the loop over the last, partial group of slabs).
struct DiagList {            // the chunk's mismatch list under construction (warp-uniform)
    uint16_t* mm; uint32_t cnt, V; bool full;
};
template <int BL>
__device__ __forceinline__ void diag_decode(DiagList& L, const uint64_t (&xr)[4], uint32_t w0, uint32_t bl, int lane) {
    const uint32_t sl = 6 - bl, B = 1u << bl;
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const bool hit = xr[r] != 0;
        unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (!bal || L.full) continue;
        // the usual case -- one mismatching symbol per mismatching word, and room in the list for all of them: every
        // lane stores its own offset at its rank among the mismatching lanes (lane order == offset order)
        const uint32_t b1 = hit ? (uint32_t)(clz64(xr[r]) >> bl) : 0u;
        const uint64_t rest = hit ? xr[r] & ~(((1ull << B) - 1) << (64 - B - B * b1)) : 0ull;
        const uint32_t nb = (uint32_t)__popc(bal);
        if (L.cnt + nb <= (uint32_t)kMmMax && !__any_sync(0xffffffffu, rest != 0)) {
            if (hit) L.mm[2 + L.cnt + __popc(bal & lt)] = (uint16_t)(((w0 + 32u * r + (uint32_t)lane) << sl) + b1);
            L.cnt += nb;
            continue;
        }
        while (bal && !L.full) {
            const int src = __ffs(bal) - 1;
            bal &= bal - 1;
            uint64_t w = __shfl_sync(0xffffffffu, xr[r], src);
            const uint32_t ob = (w0 + 32u * r + (uint32_t)src) << sl;
            while (w) {
                uint32_t b = (uint32_t)(clz64(w) >> bl);
                if (L.cnt == (uint32_t)kMmMax) { L.V = ob + b; L.full = true; break; }
                if (lane == 0) L.mm[2 + L.cnt] = (uint16_t)(ob + b);
                ++L.cnt;
                w &= ~(((1ull << B) - 1) << (64 - B - B * b));
            }
        }
    }
}
template <int BL, bool ODD>
__device__ __forceinline__ void diag_compare(DiagList& L, const uint64_t* __restrict__ xw, const uint64_t* __restrict__ Dw,
                                             unsigned s5, uint32_t nfull, uint32_t rem, uint32_t bl, int lane) {
    const uint32_t B = 1u << bl;
    const uint32_t nwords = nfull + (rem ? 1u : 0u);
    const uint64_t* __restrict__ xl = xw + lane;
    const uint64_t* __restrict__ dl = Dw + lane;
    uint32_t w0 = 0;
    for (; w0 + 128u <= nfull && !L.full; w0 += 128u, xl += 128, dl += 128) {      // four complete slabs of 32 words
        uint64_t xr[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) xr[r] = xl[32 * r] ^ funnel_fixed<ODD>(dl[32 * r], dl[32 * r + 1], s5);
        if (!__any_sync(0xffffffffu, (xr[0] | xr[1] | xr[2] | xr[3]) != 0)) continue;
        diag_decode<BL>(L, xr, w0, bl, lane);
    }
    if (w0 < nwords && !L.full) {                                  // the last, partial group (the text ends in zero pad words)
        const uint64_t tailmask = rem ? ~0ull << (64 - B * rem) : ~0ull;
        uint64_t xr[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint32_t wi = w0 + 32u * r + (uint32_t)lane;
            xr[r] = 0;
            if (wi < nwords) {
                xr[r] = xl[32 * r] ^ funnel_fixed<ODD>(dl[32 * r], dl[32 * r + 1], s5);
                if (wi == nfull) xr[r] &= tailmask;
            }
        }
        if (__any_sync(0xffffffffu, (xr[0] | xr[1] | xr[2] | xr[3]) != 0)) diag_decode<BL>(L, xr, w0, bl, lane);
    }
}

template <int BL, int MINB>        // BL = 1: 2-bit symbols known at compile time; 0: width read from the index
__global__ void __launch_bounds__(128, MINB) k_diag_w(ParseView pv) {
    const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= pv.n_chunks) return;
    uint16_t* mm = pv.mm + (size_t)g * kMmStride;
    const int64_t dg = pv.hint[g / kHintGroup];
    const uint32_t h = pv.chunk_hap[g];
    const HapView& hv = pv.haps[h];
    const uint64_t s = (uint64_t)(g - pv.hap_chunk_base[h]) << pv.chunk_shift;
    const int64_t p0 = (int64_t)s + dg;
    if (dg == kNoDiag || s >= hv.m || p0 < 0 || (uint64_t)p0 >= pv.ix.n) {
        if (lane == 0) { mm[0] = 0; mm[1] = 0; }
        return;
    }
    uint64_t e = s + (1ull << pv.chunk_shift); if (e > hv.m) e = hv.m;
    const uint32_t p = (uint32_t)p0, len = (uint32_t)(e - s);
    const uint32_t avail = (uint32_t)pv.ix.n - p;
    const uint32_t cmplen = len < avail ? len : avail;
    const uint32_t bl = BL ? (uint32_t)BL : pv.ix.bl, sl = 6 - bl, spw = 1u << sl;
    const uint64_t* __restrict__ xw = hv.x + (s >> sl);
    // the reference side: same word grid shifted by a bit offset that is constant along the diagonal
    const uint64_t* __restrict__ Dw = pv.ix.text + (p >> sl);
    const uint32_t dsh = (p & (spw - 1)) << bl;                        // 0..62
    const uint32_t nfull = cmplen >> sl, rem = cmplen & (spw - 1);     // whole words, symbols of the last partial word
    DiagList L; L.mm = mm; L.cnt = 0; L.V = len; L.full = false;
    if (dsh >= 32) diag_compare<BL, true>(L, xw, Dw, dsh & 31, nfull, rem, bl, lane);
    else diag_compare<BL, false>(L, xw, Dw, dsh & 31, nfull, rem, bl, lane);
    if (!L.full && cmplen < len) {
        if (L.cnt == (uint32_t)kMmMax) L.V = cmplen;
        else { if (lane == 0) mm[2 + L.cnt] = (uint16_t)cmplen; ++L.cnt; }
    }
    if (lane == 0) { mm[0] = (uint16_t)L.cnt; mm[1] = (uint16_t)L.V; }
}

__global__ void __launch_bounds__(128) k_resolve(ParseView pv) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < pv.n_chunks) DRLZ_WITH_BL(pv, q, resolve_chunk(q, g));
}
// P1 of one chunk per lane, the 32 chunks of a warp in step.  spec_parse_chunk() as it stands in rlz_core.cuh lets every
// lane run ahead on its own; the lanes then sit in different places of the same code, the warp has one scoreboard, and
// every dependent miss of every lane is a stall of all 32 (1600 load requests per warp, 8 lanes per instruction).  Here
// an iteration is: head of the phrase at i (literal / predicted phrase: one read of the uniqueness array) for all
// lanes; for the lanes that got a phrase, the head of the next one; then the index lookup for all lanes that stand
// on a variant site.  Same functions, same order per chunk, same phrases: only the interleaving across lanes differs.
template <typename PV>
__device__ __forceinline__ void spec_parse_warp(const PV& pv, uint32_t g, bool valid, const uint16_t* mm_copy) {
    uint32_t h = 0; HapView hv; hv.m = 0; hv.n_exc = 0;
    uint64_t s = 0, e = 0;
    bool live = false;
    if (valid) {
        h = pv.chunk_hap[g];
        hv = pv.haps[h];
        s = (uint64_t)(g - pv.hap_chunk_base[h]) << pv.chunk_shift;
        e = s + (1ull << pv.chunk_shift); if (e > hv.m) e = hv.m;
        pv.link[g] = kEmpty;
        pv.open_i[g] = kEmpty;
        if (s >= hv.m) {
            pv.first_pos[g] = 0; pv.first_len[g] = 0; pv.run_len[g] = 0;
            pv.spec_cnt[g] = 0; pv.spec_exit[g] = (uint32_t)(s > 0xFFFFFFFEull ? 0xFFFFFFFEull : s);
        } else live = true;
    }
    const bool parsed = live;
    uint64_t i = s; uint32_t cnt = 0;
    uint32_t cur = (live && hv.n_exc) ? exc_lower_bound(hv, i) : 0;
    bool open = false;
    const int64_t hint_g = (valid && pv.hint) ? pv.hint[g / kHintGroup] : kNoDiag;
    int64_t diag = hint_g;
    const uint16_t* mml = !pv.mm ? nullptr : mm_copy;
    auto take = [&](const Match& mt) {
        if (i == s) { pv.first_pos[g] = mt.pos; pv.first_len[g] = mt.len; pv.run_len[g] = open ? kEmpty : mt.len; }
        if (cnt < pv.spec_store) pv.spec_ph[(size_t)g * pv.spec_store + cnt] = make_uint2(mt.pos, mt.len);
        ++cnt;
        if (open) { pv.open_i[g] = (uint32_t)i; pv.open_pos[g] = mt.pos; pv.open_len[g] = mt.len; live = false; return; }
        i += mt.len ? mt.len : 1;
        if (i >= e) live = false;
    };
    while (__any_sync(0xffffffffu, live)) {
        Match mt; uint32_t maxlen = 0, cap = 0;
        int st = -1;
        if (live) {
            st = phrase_head(pv, hv, i, &cur, &open, &diag, mml, hint_g, s, g, &mt, &maxlen, &cap);
            if (st != kHeadLookup) take(mt);
        }
        __syncwarp();
        if (live && st != kHeadLookup) {
            st = phrase_head(pv, hv, i, &cur, &open, &diag, mml, hint_g, s, g, &mt, &maxlen, &cap);
            if (st != kHeadLookup) take(mt);
        }
        __syncwarp();
        if (live && st == kHeadLookup) {
            mt = phrase_tail(pv, hv, i, maxlen, cap, &open, &diag);
            take(mt);
        }
        __syncwarp();
    }
    if (parsed) {
        pv.spec_cnt[g] = cnt;
        if (open) pv.spec_exit[g] = kEmpty;                       // resolved later
        else { pv.spec_exit[g] = (uint32_t)i; register_exit(pv, h, i, 1); }
    }
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_spec(ParseView pv) {
    // every thread keeps its chunk's mismatch record (32 bytes) in shared memory: it is consulted at every predicted
    // phrase, and between two of them a lookup's misses push it out of L1 (14 % of the kernel's stall samples were reads
    // of the record through L2)
    __shared__ __align__(16) uint16_t smm[128 * kMmStride];
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = g < pv.n_chunks;
    uint16_t* mine = nullptr;
    if (pv.mm) {
        static_assert(kMmStride == 16, "the record is copied as two 16-byte vectors");
        mine = smm + threadIdx.x * kMmStride;
        if (valid) {
            const uint4* src = reinterpret_cast<const uint4*>(pv.mm + (size_t)g * kMmStride);
            reinterpret_cast<uint4*>(mine)[0] = src[0];
            reinterpret_cast<uint4*>(mine)[1] = src[1];
        }
    }
    if (pv.spec_sync) { DRLZ_WITH_BL(pv, q, spec_parse_warp(q, g, valid, mine)); }
    else if (valid) DRLZ_WITH_BL(pv, q, spec_parse_chunk(q, g, mine));
}
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_fix(ParseView pv, uint32_t round) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < pv.n_chunks * kAltSlots) DRLZ_WITH_BL(pv, q, fix_parse_record(q, r / kAltSlots, (int)(r % kAltSlots), round));
}

// ---- J: marking the true chain, two levels ------------------------------------------------------------------
// A record's successor always lies in a later chunk, so inside a block of kMarkChunks consecutive chunks every
// chain is resolved by pointer jumping in shared memory (k_mark_local: record -> last record of the block on its
// chain).  One thread per haplotype then hops from block to block (k_mark_walk), marking the record through which
// the true chain enters each block, and k_mark_spread propagates those marks inside the blocks (again in shared
// memory) and picks each chunk's true record.  3 launches instead of 2 + log2(chunks) global passes.
static const int kMarkChunks = 1024;
static const int kMarkThreads = 1024;     // one thread per chunk of the block: the 10 jumping rounds are latency, not work (256 threads: 62 us for both kernels)
static const int kMarkNodes = kMarkChunks * (kAltSlots + 1);

__device__ __forceinline__ void mark_load_links(const ParseView& pv, const uint32_t* succ_in, uint32_t* succ_out, uint32_t base,
                                                uint32_t nn, uint16_t* ja) {
    for (uint32_t l = threadIdx.x; l < (uint32_t)kMarkNodes; l += blockDim.x) {
        uint32_t v = base + l;
        uint16_t loc = (uint16_t)l;
        if (v < nn) {
            uint32_t sv = succ_in ? succ_in[v] : node_succ(pv, v);
            if (succ_out) succ_out[v] = sv;
            if (sv != v && sv >= base && sv - base < (uint32_t)kMarkNodes) loc = (uint16_t)(sv - base);
        }
        ja[l] = loc;
    }
}

__global__ void __launch_bounds__(kMarkThreads) k_mark_local(ParseView pv, uint32_t* succ_glob, uint32_t* block_last) {
    __shared__ uint16_t ja[kMarkNodes], jb[kMarkNodes];
    const uint32_t nn = pv.n_chunks * (kAltSlots + 1);
    const uint32_t base = blockIdx.x * kMarkNodes;
    mark_load_links(pv, nullptr, succ_glob, base, nn, ja);
    __syncthreads();
    uint16_t *a = ja, *b = jb;
    for (int span = 1; span < kMarkChunks; span *= 2) {
        for (uint32_t l = threadIdx.x; l < (uint32_t)kMarkNodes; l += blockDim.x) b[l] = a[a[l]];
        __syncthreads();
        uint16_t* t = a; a = b; b = t;
    }
    for (uint32_t l = threadIdx.x; l < (uint32_t)kMarkNodes; l += blockDim.x)
        if (base + l < nn) block_last[base + l] = base + a[l];
}

__global__ void k_mark_walk(ParseView pv, uint32_t H, const uint32_t* __restrict__ succ_glob,
                            const uint32_t* __restrict__ block_last, uint8_t* mark) {
    uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= H) return;
    if (pv.haps[h].m == 0 || pv.hap_chunk_base[h] >= pv.hap_chunk_base[h + 1]) return;
    uint32_t cur = pv.hap_chunk_base[h] * (kAltSlots + 1);       // root: speculative record of the first chunk
    for (;;) {
        mark[cur] = 1;
        uint32_t ex = block_last[cur];
        uint32_t nx = succ_glob[ex];
        if (nx == ex) break;                                     // parse finished
        cur = nx;
    }
}

__global__ void __launch_bounds__(kMarkThreads) k_mark_spread(ParseView pv, const uint32_t* __restrict__ succ_glob,
                                                      const uint8_t* __restrict__ mark_in, uint32_t* true_cnt,
                                                      uint32_t* true_entry, uint32_t* true_slot) {
    __shared__ uint16_t ja[kMarkNodes], jb[kMarkNodes];
    __shared__ uint8_t mk[kMarkNodes];
    const uint32_t nn = pv.n_chunks * (kAltSlots + 1);
    const uint32_t base = blockIdx.x * kMarkNodes;
    mark_load_links(pv, succ_glob, nullptr, base, nn, ja);
    for (uint32_t l = threadIdx.x; l < (uint32_t)kMarkNodes; l += blockDim.x) mk[l] = base + l < nn ? mark_in[base + l] : 0;
    __syncthreads();
    uint16_t *a = ja, *b = jb;
    for (int span = 1; span < kMarkChunks; span *= 2) {
        for (uint32_t l = threadIdx.x; l < (uint32_t)kMarkNodes; l += blockDim.x) {
            uint16_t j = a[l];
            if (mk[l]) mk[j] = 1;           // marks only ever travel along the chain: racing writes are benign
            b[l] = a[j];
        }
        __syncthreads();
        uint16_t* t = a; a = b; b = t;
    }
    for (uint32_t c = threadIdx.x; c < (uint32_t)kMarkChunks; c += blockDim.x) {
        uint32_t g = blockIdx.x * kMarkChunks + c;
        if (g >= pv.n_chunks) continue;
        uint32_t e, sl;
        true_cnt[g] = chunk_true_record(pv, mk + c * (kAltSlots + 1), g, &e, &sl);
        true_entry[g] = e;
        true_slot[g] = sl;
    }
}

// ---- A: run lengths of open first phrases as a right-to-left segmented sum ---------------------------------------
// With min_cap <= chunk size an open first phrase is verified exactly to its chunk end, so link[g] == g + 1 and the
// list ranking is a segmented suffix sum over consecutive chunks: one 3-launch scan instead of log2(chunks) rounds.
struct OpSegSum {
    __device__ __forceinline__ uint64_t operator()(uint64_t a, uint64_t b) const {
        const uint64_t F = 1ull << 63, M = F - 1;
        uint64_t sum = ((b & F) ? 0ull : (a & M)) + (b & M);
        return ((a | b) & F) | (sum & M);
    }
};
struct LoadRunRev {
    const uint32_t* run; const uint32_t* link; uint64_t nc;
    __device__ __forceinline__ uint64_t operator()(uint64_t e) const {
        uint64_t g = nc - 1 - e;
        return (uint64_t)run[g] | (link[g] == kEmpty ? (1ull << 63) : 0ull);
    }
};
struct StoreRunRev {
    uint32_t* out; uint64_t nc;
    __device__ __forceinline__ void operator()(uint64_t e, uint64_t v) const { out[nc - 1 - e] = (uint32_t)v; }
};

// Everything the host reads back after a batch, gathered into one block so that it is one device-to-host copy:
//   u64 [0, H) gapless lengths | [H, 2H+1) first record of each haplotype, [2H] = total | [2H+1] total |
//   [2H+2, 3H+2) exceptions per row | u32 flags[2] | u32 counters[2 + kMaxRounds]        (same layout on the host)
__global__ void k_results(ParseView pv, uint32_t H, const uint64_t* true_off, const uint64_t* total, const uint64_t* m,
                          const uint64_t* exccnt, const uint32_t* flags, uint64_t* res) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t tot = total ? *total : 0;
    if (t < H) { res[t] = m[t]; res[2 * (size_t)H + 2 + t] = exccnt[t]; }
    if (t <= H) {
        uint32_t g = t < H ? pv.hap_chunk_base[t] : pv.n_chunks;
        res[(size_t)H + t] = g < pv.n_chunks ? true_off[g] : tot;
    }
    if (t == 0) res[2 * (size_t)H + 1] = tot;
    uint32_t* r32 = reinterpret_cast<uint32_t*>(res + 3 * (size_t)H + 2);
    if (t < 2) r32[t] = flags[t];
    if (t < 2u + (uint32_t)kMaxRounds) r32[2 + t] = pv.counters[t];
}
static void launch_results(ParseBuffers& pb, const uint64_t* total, cudaStream_t s) {
    uint32_t nthr = pb.H + 1 > 2u + (uint32_t)kMaxRounds ? pb.H + 1 : 2u + (uint32_t)kMaxRounds;
    g_launches += 1;
    k_results<<<(nthr + 127) / 128, 128, 0, s>>>(pb.pv, pb.H, pb.true_off, total, pb.m_dev, pb.exccnt_dev, pb.flags_dev, pb.res);
    cudaMemcpyAsync(pb.host_res, pb.res, pb.res_bytes, cudaMemcpyDeviceToHost, s);
}

// E: one warp per 32 consecutive chunks.  The lanes first fetch the 32 chunks' (count, offset, slot) in one coalesced
// round trip; then the warp goes through the chunks, lane t copying stored phrase t of the chunk's true record into
// record t of its output run, so that the 8-byte phrase reads and the 16-byte record writes of a warp instruction are
// contiguous (one thread per chunk wrote 32 separate runs per instruction).  Same selection as emit_stored(); a
// record whose phrases were not all stored is re-parsed by its lane first (emit_chunk).
__device__ __noinline__ void emit_chunk_slow(const ParseView& pv, uint32_t g, uint32_t entry, uint32_t cnt, int64_t* out) {
    emit_chunk(pv, g, entry, cnt, out);
}
__global__ void __launch_bounds__(128, 16) k_emit(ParseView pv, const uint32_t* __restrict__ true_cnt, const uint32_t* __restrict__ true_entry,
                                              const uint32_t* __restrict__ true_slot, const uint64_t* __restrict__ true_off,
                                              int64_t* out, uint64_t cap) {
    const uint32_t g0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32u;
    const uint32_t lane = threadIdx.x & 31;
    if (g0 >= pv.n_chunks) return;
    const uint32_t gl = g0 + lane;
    uint32_t c_l = 0, slot_l = 0, own_l = 0, mg_l = kEmpty;
    uint64_t o_l = 0;
    if (gl < pv.n_chunks) c_l = true_cnt[gl];
    if (c_l) {
        o_l = true_off[gl]; slot_l = true_slot[gl];
        if (o_l + c_l > cap) c_l = 0;                        // speculative launch into a buffer that may be too small
    }
    if (c_l) {
        bool stored;
        if (slot_l == 0) stored = c_l <= pv.spec_store;
        else {
            const uint32_t rid = gl * kAltSlots + slot_l - 1;
            own_l = pv.alt_own[rid]; mg_l = pv.alt_merge[rid];
            const uint32_t sc = mg_l != kEmpty ? pv.spec_cnt[gl] : 0u;
            stored = own_l <= (uint32_t)kAltStore && sc <= pv.spec_store && (mg_l == kEmpty || mg_l <= sc) &&
                     own_l + (mg_l != kEmpty ? sc - mg_l : 0u) == c_l;         // the copy below trusts these three
        }
        if (!stored) { emit_chunk_slow(pv, gl, true_entry[gl], c_l, out + 2 * o_l); c_l = 0; }
    }
    unsigned todo = __ballot_sync(0xffffffffu, c_l != 0);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t c = __shfl_sync(0xffffffffu, c_l, src), slot = __shfl_sync(0xffffffffu, slot_l, src);
        const uint64_t o = __shfl_sync(0xffffffffu, o_l, src);
        const uint32_t g = g0 + (uint32_t)src;
        const uint2* __restrict__ sp = pv.spec_ph + (size_t)g * pv.spec_store;
        int64_t* dst = out + 2 * o;
        if (slot == 0) {
            for (uint32_t t = lane; t < c; t += 32) put_record(dst, t, sp[t]);
        } else {
            const uint32_t own = __shfl_sync(0xffffffffu, own_l, src), mg = __shfl_sync(0xffffffffu, mg_l, src);
            const uint2* __restrict__ ap = pv.alt_ph + (size_t)(g * kAltSlots + slot - 1) * kAltStore;
            for (uint32_t t = lane; t < c; t += 32) put_record(dst, t, t < own ? ap[t] : sp[mg + (t - own)]);
        }
    }
}

int parse_stage_count(ParseBuffers& pb, cudaStream_t s, int* rounds_out, cudaEvent_t* ev) {
    ParseView& pv = pb.pv;
    uint32_t NC = pv.n_chunks;
    *pb.host_total = 0;
    if (rounds_out) *rounds_out = 0;
    if (NC == 0) {
        if (ev) for (int i = 0; i < 6; ++i) cudaEventRecord(ev[i], s);
        cudaMemsetAsync(pv.counters, 0, sizeof(uint32_t) * (2 + kMaxRounds), s);
        launch_results(pb, nullptr, s);
        cudaStreamSynchronize(s);
        return 0;
    }
    cudaMemsetAsync(pv.alt_entry, 0xFF, sizeof(uint32_t) * (size_t)NC * kAltSlots, s);
    cudaMemsetAsync(pv.counters, 0, sizeof(uint32_t) * (2 + kMaxRounds), s);
    if (ev) cudaEventRecord(ev[0], s);
    pv.ext = nullptr;
    if (pv.ix.uniq && pb.hint) {            // H: one 64-base lookup per group of chunks seeds the predicted diagonals
        uint32_t nq = (NC + kHintGroup - 1) / kHintGroup;
        g_launches += 1;
        k_hint<<<(nq + 127) / 128, 128, 0, s>>>(pv, pb.hint, nq);
        pv.hint = pb.hint;
        // references with repeats beyond the 16-bit range of the uniqueness array (runs of N): chunks without a hint of
        // their own take the one in front of them
        const bool long_repeats = pv.ix.uniq32 != nullptr && pb.ext && pb.hint_src && kHintGroup == 1;
        if (long_repeats) {
            // first from behind (a chunk inside a run is placed by what follows the run), then from the front for what is left
            g_launches += 1;
            device_scan<uint32_t, LoadHintIdxRev, StoreHintSrcRev, OpMin, true>(LoadHintIdxRev{pb.hint, NC}, StoreHintSrcRev{pb.hint_src, NC}, NC,
                                                                                reinterpret_cast<uint32_t*>(pb.scan_tmp), OpMin(), 0xFFFFFFFFu, s);
            k_hint_fill<<<(NC + 255) / 256, 256, 0, s>>>(pv, pb.hint, pb.hint_src);
            device_scan<uint32_t, LoadHintIdx, StoreHintSrc, OpMax, true>(LoadHintIdx{pb.hint}, StoreHintSrc{pb.hint_src}, NC,
                                                                          reinterpret_cast<uint32_t*>(pb.scan_tmp), OpMax(), 0u, s);
            k_hint_fill<<<(NC + 255) / 256, 256, 0, s>>>(pv, pb.hint, pb.hint_src);
            g_launches += 1;
        }
        if (pb.mm && pv.chunk_shift >= 5 && pv.chunk_shift <= 15) {   // D: mismatch lists along the hinted diagonals
            pv.mm = pb.mm;
            g_launches += 1;
            const unsigned nb = (unsigned)(((uint64_t)NC * 32 + 127) / 128);
            if (pv.ix.bl == 2) k_diag_w<2, 12><<<nb, 128, 0, s>>>(pv);          // 4-bit symbols: references with N, the real-data path
            else if (pv.ix.bl != 1) k_diag_w<0, 16><<<nb, 128, 0, s>>>(pv);
            else if (pb.occ_diag >= 16) k_diag_w<1, 16><<<nb, 128, 0, s>>>(pv);
            else if (pb.occ_diag >= 12) k_diag_w<1, 12><<<nb, 128, 0, s>>>(pv);
            else k_diag_w<1, 10><<<nb, 128, 0, s>>>(pv);
            if (long_repeats) {                                 // ... and the matches along the diagonals are measured across chunks
                device_scan<uint64_t, LoadExtRev, StoreExtRev, OpExtSum, true>(LoadExtRev{pv, NC}, StoreExtRev{pb.ext, NC}, NC, pb.scan_tmp,
                                                                               OpExtSum(), 0ull, s);
                pv.ext = pb.ext;
            }
        } else pv.mm = nullptr;
    } else { pv.hint = nullptr; pv.mm = nullptr; }
    if (ev) cudaEventRecord(ev[5], s);
    g_launches += 1;
    if (pb.occ >= 16) k_spec<16><<<(NC + 127) / 128, 128, 0, s>>>(pv);
    else if (pb.occ >= 12) k_spec<12><<<(NC + 127) / 128, 128, 0, s>>>(pv);
    else if (pb.occ >= 10) k_spec<10><<<(NC + 127) / 128, 128, 0, s>>>(pv);
    else if (pb.occ >= 8) k_spec<8><<<(NC + 127) / 128, 128, 0, s>>>(pv);
    else k_spec<1><<<(NC + 127) / 128, 128, 0, s>>>(pv);
    if (ev) cudaEventRecord(ev[1], s);
    // stage A: diagonal-run links, list ranking, exits of the chunks whose last phrase is open
    g_launches += 2;
    k_link<<<(NC + 127) / 128, 128, 0, s>>>(pv);
    if (pv.min_cap <= pv.chunk_bases) {
        device_scan<uint64_t, LoadRunRev, StoreRunRev, OpSegSum, true>(LoadRunRev{pv.run_len, pv.link, NC}, StoreRunRev{pb.run_b, NC},
                                                                        NC, pb.scan_tmp, OpSegSum(), 0ull, s);
        pv.run_len = pb.run_b;
    } else {
        uint32_t *li = pv.link, *ri = pv.run_len, *lo = pb.link_b, *ro = pb.run_b;
        for (uint64_t span = 1; span < pb.max_chunks_per_hap; span *= 2) {
            g_launches += 1;
            k_rank<<<(NC + 255) / 256, 256, 0, s>>>(li, ri, lo, ro, NC);
            uint32_t* t = li; li = lo; lo = t; t = ri; ri = ro; ro = t;
        }
        pv.link = li; pv.run_len = ri;
    }
    k_resolve<<<(NC + 127) / 128, 128, 0, s>>>(pv);
    if (ev) cudaEventRecord(ev[4], s);
    // F: fix-up rounds.  Whether another round is needed is known on the device only (counters[1 + r] = records
    // created for round r); reading it back after every round idles the GPU for a round trip each time.  So the
    // usual number of rounds is launched blind (a round without records is an empty launch), the marking and the
    // compaction follow in the same breath, and one read-back at the end tells whether that was enough.  If not
    // (never seen on real rates), the remaining rounds run with a read-back each and J is redone.
    auto launch_fix = [&](int r) {
        uint32_t nr = NC * kAltSlots;
        g_launches += 1;
        if (pb.occ_fix >= 16) k_fix<16><<<(nr + 127) / 128, 128, 0, s>>>(pv, (uint32_t)r);
        else if (pb.occ_fix >= 10) k_fix<10><<<(nr + 127) / 128, 128, 0, s>>>(pv, (uint32_t)r);
        else k_fix<1><<<(nr + 127) / 128, 128, 0, s>>>(pv, (uint32_t)r);
    };
    int round = 1;
    const int blind = pb.blind_rounds < 0 ? 0 : (pb.blind_rounds >= kMaxRounds ? kMaxRounds - 1 : pb.blind_rounds);
    for (; round <= blind; ++round) launch_fix(round);
    const uint32_t nn = NC * (kAltSlots + 1);
    const uint32_t nmb = (nn + kMarkNodes - 1) / kMarkNodes;
    const uint64_t nt = ((uint64_t)NC + kScanTile - 1) / kScanTile;
    for (int pass = 0;; ++pass) {
        if (ev && pass == 0) cudaEventRecord(ev[2], s);
        pv.parsed_rounds = (uint32_t)(round - 1);
        g_launches += 3;
        cudaMemsetAsync(pb.mark, 0, nn, s);
        k_mark_local<<<nmb, kMarkThreads, 0, s>>>(pv, pb.jump_a, pb.jump_b);
        k_mark_walk<<<(pb.H + 127) / 128, 128, 0, s>>>(pv, pb.H, pb.jump_a, pb.jump_b, pb.mark);
        k_mark_spread<<<nmb, kMarkThreads, 0, s>>>(pv, pb.jump_a, pb.mark, pb.true_cnt, pb.true_entry, pb.true_slot);
        device_scan<uint64_t, LoadArr<uint32_t, uint64_t>, StoreArr<uint64_t>, OpSum, false>(
            LoadArr<uint32_t, uint64_t>{pb.true_cnt}, StoreArr<uint64_t>{pb.true_off}, NC, pb.scan_tmp, OpSum(), 0ull, s);
        if (ev && pass == 0) cudaEventRecord(ev[3], s);
        pb.emitted = false;
        if (pb.spec_out && pb.spec_out_cap) {
            // E without waiting for the host: the records go to the caller's buffer if they fit (checked per chunk on
            // the device, and by the host once the total is known; too small => the host grows it and emits again)
            g_launches += 1;
            k_emit<<<(NC + 127) / 128, 128, 0, s>>>(pv, pb.true_cnt, pb.true_entry, pb.true_slot, pb.true_off, pb.spec_out, pb.spec_out_cap);
            if (pb.ev_emitted) cudaEventRecord(pb.ev_emitted, s);
            pb.emitted = true;
        }
        launch_results(pb, pb.scan_tmp + nt, s);             // totals, per-haplotype offsets, lengths, flags, counters: one copy
        cudaStreamSynchronize(s);
        if (pb.host_counters[0]) return 2;
        if (pb.host_counters[1 + round] == 0) break;        // nothing was registered for the next round: converged
        while (pb.host_counters[1 + round] != 0) {           // rare: more rounds than launched blind
            if (round >= kMaxRounds) return 2;
            launch_fix(round);
            ++round;
            cudaMemcpyAsync(pb.host_counters, pv.counters, sizeof(uint32_t) * (2 + kMaxRounds), cudaMemcpyDeviceToHost, s);
            cudaStreamSynchronize(s);
            if (pb.host_counters[0]) return 2;
        }
    }
    if (rounds_out) {                                        // rounds that had work + 1, as before
        int r = 1;
        while (r < kMaxRounds && pb.host_counters[1 + r] != 0) ++r;
        *rounds_out = r;
    }
    if (pb.host_counters[0]) return 2;
    return 0;
}

void parse_stage_emit(ParseBuffers& pb, int64_t* out, cudaStream_t s) {
    uint32_t NC = pb.pv.n_chunks;
    if (NC == 0) return;
    g_launches += 1;
    k_emit<<<(NC + 127) / 128, 128, 0, s>>>(pb.pv, pb.true_cnt, pb.true_entry, pb.true_slot, pb.true_off, out, ~0ull);
}

__global__ void k_fill_chunk_hap(uint32_t* chunk_hap, const uint32_t* __restrict__ base, uint32_t H, uint32_t NC) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= NC) return;
    uint32_t a = 0, b = H;               // last h with base[h] <= g
    while (b - a > 1) { uint32_t mid = a + ((b - a) >> 1); if (base[mid] <= g) a = mid; else b = mid; }
    chunk_hap[g] = a;
}
void launch_fill_chunk_hap(uint32_t* chunk_hap, const uint32_t* hap_chunk_base, uint32_t H, uint32_t NC, cudaStream_t s) {
    if (NC == 0) return;
    g_launches += 1;
    k_fill_chunk_hap<<<(NC + 255) / 256, 256, 0, s>>>(chunk_hap, hap_chunk_base, H, NC);
}

// dense matching statistics: one thread per position (diagnostic entry point drlz_matching_stats)
__global__ void __launch_bounds__(128) k_ms_dense(IndexView ix, const uint64_t* __restrict__ X, uint64_t m, uint32_t* ms, uint32_t* pos) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    bool open;
    Match mt = lookup(ix, X, i, (uint32_t)(m - i), (uint32_t)(m - i), &open);
    if (mt.len == 0) mt.pos = (uint32_t)ix.c2b[base_at(X, i, ix.bl)];
    ms[i] = mt.len; pos[i] = mt.pos;
}
void launch_ms_dense(const IndexView& ix, const uint64_t* X, uint64_t m, uint32_t* ms, uint32_t* pos, cudaStream_t s) {
    if (m == 0) return;
    g_launches += 1;
    k_ms_dense<<<(unsigned)((m + 127) / 128), 128, 0, s>>>(ix, X, m, ms, pos);
}

}  // namespace drlz
