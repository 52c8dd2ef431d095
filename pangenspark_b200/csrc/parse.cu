// parse.cu -- kernels of the chunk-parallel exact greedy RLZ parse (scalar logic in rlz_core.cuh).
//
//   k_first_w   A : phrase at the first base of every chunk, compared no further than the chunk end
//   k_link      A : open first phrases linked to the chunk their match continues in
//   k_rank      A : list ranking (pointer jumping) -> exact lengths of arbitrarily long matches
//   k_spec_w    P1: one lane per chunk, speculative greedy parse from the chunk's first base; phrases stored
//   k_fix       F : one thread per (chunk, alternative entry) created in the previous round
//   k_succ      J : successor of every record; roots marked
//   k_jump      J : pointer jumping (mark[jump[v]] |= mark[v]; jump = jump o jump), log2(chunks) rounds
//   k_true      E : per chunk the marked record's phrase count, entry and slot
//   device_scan E : exclusive prefix of the counts = output slot of every chunk (stream compaction)
//   k_emit      E : gather the stored phrases of the true records into (int64 pos, int64 len) records
//                   (DistributedRLZ.scala:263-284 record layout); re-parses only what was not stored
//
// k_first_w / k_spec_w use warp_lookup(): the scalar lookup() of rlz_core.cuh re-expressed as a warp-synchronous
// state machine.  Every lane owns one chunk and performs at most ONE suffix-array probe per iteration, all at the
// same program location (no divergence between lanes in different stages of their binary searches); a probe
// compares 64 bases per lane from two packed words, and matches that run past 64 bases are extended by the whole
// warp together: 32 lanes x 32 bases = 1024 bases per step with coalesced 256-byte reads of both streams.
// (ncu on the scalar version: 4.4 of 32 lanes active per instruction, see profiles/README.md.)
#include "drlz_internal.h"
#include "scan.cuh"

namespace drlz {

#define FULLMASK 0xffffffffu

// All 32 lanes must call.  Lanes with active == true look up x[i..i+maxlen) capped at cap (semantics of lookup()).
__device__ __forceinline__ Match warp_lookup(const IndexView& ix, bool active, const uint64_t* __restrict__ X, uint64_t i,
                                             uint32_t maxlen, uint32_t cap, bool* open_out) {
    const int lane = threadIdx.x & 31;
    const int k = ix.k;
    const uint64_t n = ix.n;
    enum { BS = 0, PL1 = 1, PL2 = 2, SINGLE = 3, LEFT = 4, DONE = 5 };
    int phase = active ? BS : DONE;
    uint64_t xw0 = 0, xw1 = 0, kmer = 0;
    uint32_t plen = 0, kk = 0, lo = 0, l = 0, h = 0, r = 0, l1 = 0, l2 = 0, lcp_left = 0, lcp_right = 0, sar = 0, a = 0, b = 0, L = 0;
    bool have_left = false, have_right = false, open = false;
    Match mt; mt.len = 0; mt.pos = 0;
    if (active) {
        xw0 = get64(X, i);
        xw1 = get64(X, i + 32);
        kmer = xw0 >> (64 - 2 * k);
        plen = cap < maxlen ? cap : maxlen;
        kk = plen < (uint32_t)k ? plen : (uint32_t)k;
        uint64_t lowmask = (1ull << (2 * (k - (int)kk))) - 1;
        lo = ix.tab[kmer & ~lowmask];
        l = lo; h = ix.tab[(kmer | lowmask) + 1];
    }
    while (__any_sync(FULLMASK, phase != DONE)) {
        // ---- per-lane transitions until a probe is needed (no memory traffic except rare table reads)
        bool need = false;
        uint32_t idx = 0, lim_pat = 0;
        if (phase != DONE) {
            for (;;) {
                if (phase == BS) {
                    if (l < h) { idx = l + ((h - l) >> 1); lim_pat = plen; need = true; break; }
                    r = l; phase = PL1;
                }
                if (phase == PL1) {
                    if (r > 0 && !have_left) { idx = r - 1; lim_pat = plen; need = true; break; }
                    l1 = r > 0 ? lcp_left : 0; phase = PL2;
                }
                if (phase == PL2) {
                    if ((uint64_t)r < n && !have_right) { idx = r; lim_pat = plen; need = true; break; }
                    l2 = (uint64_t)r < n ? lcp_right : 0;
                    if (l2 > l1) {
                        if (l2 == plen && plen < maxlen) phase = SINGLE;
                        else { mt.len = l2; mt.pos = sar; phase = DONE; break; }
                    } else if (l1 == 0) { mt.len = 0; mt.pos = 0; phase = DONE; break; }
                    else {
                        L = l1;
                        uint32_t Lk = L < (uint32_t)k ? L : (uint32_t)k;
                        a = lo;
                        if (Lk != kk) { uint64_t m2 = (1ull << (2 * (k - (int)Lk))) - 1; a = ix.tab[kmer & ~m2]; }
                        b = r - 1; phase = LEFT;
                    }
                }
                if (phase == SINGLE) {
                    if ((uint64_t)r + 1 < n) { idx = r + 1; lim_pat = plen; need = true; break; }
                    open = true; mt.len = l2; mt.pos = sar; phase = DONE; break;
                }
                if (phase == LEFT) {
                    if (a < b) { idx = a + ((b - a) >> 1); lim_pat = L; need = true; break; }
                    mt.len = L; mt.pos = ix.sa[a]; phase = DONE; break;
                }
                break;
            }
        }
        // ---- one probe per lane: 64 bases from two packed words
        uint32_t p = 0, c = 0, lim = 0;
        bool less = false, sat = false;
        if (need) {
            p = ix.sa[idx];
            uint64_t rem = n - p;
            lim = rem < (uint64_t)lim_pat ? (uint32_t)rem : lim_pat;
            uint64_t b0 = get64(ix.text, p);
            uint64_t x0 = xw0 ^ b0;
            if (x0) { c = (uint32_t)(clz64(x0) >> 1); less = b0 < xw0; }
            else {
                c = 32;
                if (lim > 32) {
                    uint64_t b1 = get64(ix.text, (uint64_t)p + 32);
                    uint64_t x1 = xw1 ^ b1;
                    if (x1) { c = 32 + (uint32_t)(clz64(x1) >> 1); less = b1 < xw1; }
                    else c = 64;
                }
            }
            if (c >= lim) { c = lim; less = (lim != lim_pat); }
            else if (c == 64) sat = true;
        }
        // ---- matches longer than 64 bases: the warp extends them together, 1024 bases per step
        unsigned em = __ballot_sync(FULLMASK, sat);
        while (em) {
            int src = __ffs(em) - 1;
            em &= em - 1;
            uint64_t xi = __shfl_sync(FULLMASK, (unsigned long long)i, src) + 64;
            uint64_t dp = (uint64_t)__shfl_sync(FULLMASK, p, src) + 64;
            uint32_t limE = __shfl_sync(FULLMASK, lim, src) - 64;
            const uint64_t* Xs = reinterpret_cast<const uint64_t*>(__shfl_sync(FULLMASK, (unsigned long long)X, src));
            uint32_t off = 0, cext = limE;
            bool lext = false, found = false;
            while (off < limE) {
                uint32_t pos = off + 32u * (uint32_t)lane;
                uint64_t xr = 0; bool bl = false;
                if (pos < limE) {
                    uint64_t aa = get64(Xs, xi + pos), bb = get64(ix.text, dp + pos);
                    xr = aa ^ bb; bl = bb < aa;
                }
                unsigned bal = __ballot_sync(FULLMASK, xr != 0);
                if (bal) {
                    int t0 = __ffs(bal) - 1;
                    uint32_t lc = __shfl_sync(FULLMASK, (uint32_t)(clz64(xr) >> 1), t0);
                    lext = __shfl_sync(FULLMASK, (int)bl, t0) != 0;
                    cext = off + 32u * (uint32_t)t0 + lc;
                    found = true;
                    break;
                }
                off += 1024;
            }
            if (!found || cext >= limE) { cext = limE; found = false; }
            if (lane == src) { c = 64 + cext; less = found ? lext : (lim != lim_pat); }
        }
        // ---- consume the probe
        if (need) {
            if (phase == BS) {
                if (less) { l = idx + 1; lcp_left = c; have_left = true; }
                else { h = idx; lcp_right = c; have_right = true; sar = p; }
            } else if (phase == PL1) { lcp_left = c; have_left = true; }
            else if (phase == PL2) { lcp_right = c; have_right = true; sar = p; }
            else if (phase == SINGLE) {
                if (c >= plen) {            // several suffixes share the capped pattern: redo uncapped (exact)
                    plen = maxlen;
                    kk = plen < (uint32_t)k ? plen : (uint32_t)k;
                    uint64_t lowmask = (1ull << (2 * (k - (int)kk))) - 1;
                    lo = ix.tab[kmer & ~lowmask];
                    l = lo; h = ix.tab[(kmer | lowmask) + 1];
                    have_left = have_right = false;
                    phase = BS;
                } else { open = true; mt.len = l2; mt.pos = sar; phase = DONE; }
            } else if (phase == LEFT) { if (c >= L) b = idx; else a = idx + 1; }
        }
    }
    *open_out = open;
    return mt;
}

// warp version of phrase_at(): all lanes call; `act` lanes get their exact phrase at position i
__device__ __forceinline__ Match warp_phrase_at(const ParseView& pv, bool act, uint32_t h, const HapView& hv, uint64_t i,
                                                uint32_t* cursor) {
    Match mt; mt.len = 0; mt.pos = 0;
    bool lit = false;
    uint32_t maxlen = 0, cap = 0;
    if (act) {
        uint64_t stop = stop_after(hv, i, cursor);
        if (stop == i) { lit = true; mt.pos = hv.exc_byte[*cursor]; }
        else {
            maxlen = (uint32_t)(stop - i);
            uint64_t C = pv.chunk_bases;
            uint64_t to_end = (i / C + 1) * C - i;
            cap = to_end > (uint64_t)pv.min_cap ? (to_end > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)to_end) : pv.min_cap;
        }
    }
    bool open;
    Match lm = warp_lookup(pv.ix, act && !lit, hv.x, i, maxlen, cap, &open);
    if (act && !lit) {
        mt = lm;
        if (mt.len == 0) mt.pos = (uint32_t)("ACGT"[base_at(hv.x, i)]);
        else if (open) {
            uint64_t j; bool same;
            uint32_t g2 = open_target(pv, h, i, mt, &j, &same);
            uint32_t rest = same ? pv.run_len[g2] : foreign_run(pv, hv, i, mt, j);
            mt.len = (uint32_t)(j - i) + rest;
        }
    }
    return mt;
}

__global__ void __launch_bounds__(128) k_first_w(ParseView pv) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = g < pv.n_chunks;
    uint32_t h = 0; uint64_t s = 0; HapView hv; hv.x = nullptr; hv.m = 0; hv.exc_pos = nullptr; hv.exc_byte = nullptr; hv.n_exc = 0;
    bool act = false, lit = false;
    uint32_t maxlen = 0, cap = 0, litbyte = 0;
    if (valid) {
        h = pv.chunk_hap[g];
        hv = pv.haps[h];
        s = (uint64_t)(g - pv.hap_chunk_base[h]) * pv.chunk_bases;
        pv.link[g] = kEmpty;
        if (s >= hv.m) { pv.first_pos[g] = 0; pv.first_len[g] = 0; pv.run_len[g] = 0; }
        else {
            uint32_t cur = hv.n_exc ? exc_lower_bound(hv, s) : 0;
            uint64_t stop = stop_after(hv, s, &cur);
            if (stop == s) { lit = true; litbyte = hv.exc_byte[cur]; }
            else {
                act = true;
                maxlen = (uint32_t)(stop - s);
                uint64_t C = pv.chunk_bases;
                cap = C > (uint64_t)pv.min_cap ? (uint32_t)C : pv.min_cap;
            }
        }
    }
    bool open;
    Match mt = warp_lookup(pv.ix, act, hv.x, s, maxlen, cap, &open);
    if (lit) { pv.first_pos[g] = litbyte; pv.first_len[g] = 0; pv.run_len[g] = 0; }
    if (act) {
        if (mt.len == 0) mt.pos = (uint32_t)("ACGT"[base_at(hv.x, s)]);
        pv.first_pos[g] = mt.pos;
        pv.first_len[g] = mt.len;
        pv.run_len[g] = open ? kEmpty : mt.len;
    }
}

__global__ void __launch_bounds__(128) k_link(ParseView pv) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < pv.n_chunks) link_chunk(pv, g);
}

__global__ void k_rank(const uint32_t* __restrict__ link_in, const uint32_t* __restrict__ run_in, uint32_t* link_out,
                       uint32_t* run_out, uint32_t nc) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < nc) rank_round(link_in, run_in, link_out, run_out, g);
}

__global__ void __launch_bounds__(128) k_spec_w(ParseView pv) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = g < pv.n_chunks;
    uint32_t h = 0, cnt = 0, cur = 0;
    uint64_t s = 0, e = 0, i = 0;
    HapView hv; hv.x = nullptr; hv.m = 0; hv.exc_pos = nullptr; hv.exc_byte = nullptr; hv.n_exc = 0;
    if (valid) {
        h = pv.chunk_hap[g];
        hv = pv.haps[h];
        s = (uint64_t)(g - pv.hap_chunk_base[h]) * pv.chunk_bases;
        e = s + pv.chunk_bases; if (e > hv.m) e = hv.m;
        i = s;
        if (i < e) {                                   // first phrase: stage A's result
            Match mt; mt.pos = pv.first_pos[g]; mt.len = pv.run_len[g];
            pv.spec_ph[(size_t)g * kSpecStore] = make_uint2(mt.pos, mt.len);
            cnt = 1;
            i += mt.len ? mt.len : 1;
            cur = hv.n_exc ? exc_lower_bound(hv, i) : 0;
        }
    }
    for (;;) {
        bool act = valid && i < e;
        if (!__any_sync(FULLMASK, act)) break;
        Match mt = warp_phrase_at(pv, act, h, hv, i, &cur);
        if (act) {
            if (cnt < (uint32_t)kSpecStore) pv.spec_ph[(size_t)g * kSpecStore + cnt] = make_uint2(mt.pos, mt.len);
            ++cnt;
            i += mt.len ? mt.len : 1;
        }
    }
    if (valid) {
        pv.spec_cnt[g] = cnt;
        pv.spec_exit[g] = (uint32_t)i;
        register_exit(pv, h, i, 1);
    }
}

// scalar versions (used by the exact sequential fallback, where one chunk = one haplotype)
__global__ void __launch_bounds__(128) k_first(ParseView pv) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < pv.n_chunks) first_phrase_chunk(pv, g);
}
__global__ void __launch_bounds__(128) k_spec(ParseView pv) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < pv.n_chunks) spec_parse_chunk(pv, g);
}

__global__ void __launch_bounds__(128) k_fix(ParseView pv, uint32_t round) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < pv.n_chunks * kAltSlots) fix_parse_record(pv, r / kAltSlots, (int)(r % kAltSlots), round);
}

__global__ void k_succ(ParseView pv, uint32_t* jump, uint8_t* mark) {
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t nn = pv.n_chunks * (kAltSlots + 1);
    if (v >= nn) return;
    jump[v] = node_succ(pv, v);
    uint32_t g = v / (kAltSlots + 1);
    uint32_t h = pv.chunk_hap[g];
    // root: speculative record of the first chunk of a non-empty haplotype
    mark[v] = (v % (kAltSlots + 1) == 0 && g == pv.hap_chunk_base[h] && pv.haps[h].m > 0) ? 1 : 0;
}

__global__ void k_jump(const uint32_t* __restrict__ jump, uint32_t* __restrict__ jump2, uint8_t* mark, uint32_t nn) {
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nn) return;
    uint32_t j = jump[v];
    if (mark[v]) mark[j] = 1;
    jump2[v] = jump[j];
}

__global__ void k_true(ParseView pv, const uint8_t* mark, uint32_t* true_cnt, uint32_t* true_entry, uint32_t* true_slot) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= pv.n_chunks) return;
    uint32_t e, sl;
    true_cnt[g] = chunk_true_record(pv, mark, g, &e, &sl);
    true_entry[g] = e;
    true_slot[g] = sl;
}

__global__ void k_hap_pairs(ParseView pv, uint32_t H, const uint64_t* true_off, const uint64_t* total, uint64_t* hap_pairs) {
    uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h > H) return;
    uint32_t g = h < H ? pv.hap_chunk_base[h] : pv.n_chunks;
    hap_pairs[h] = g < pv.n_chunks ? true_off[g] : *total;
}

__global__ void __launch_bounds__(128) k_emit(ParseView pv, const uint32_t* true_cnt, const uint32_t* true_entry,
                                              const uint32_t* true_slot, const uint64_t* true_off, int64_t* out) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= pv.n_chunks) return;
    uint32_t c = true_cnt[g];
    if (!c) return;
    int64_t* o = out + 2 * true_off[g];
    if (!emit_stored(pv, g, true_slot[g], c, o)) emit_chunk(pv, g, true_entry[g], c, o);
}

int parse_stage_count(ParseBuffers& pb, cudaStream_t s, int* rounds_out, cudaEvent_t* ev) {
    ParseView& pv = pb.pv;
    uint32_t NC = pv.n_chunks;
    *pb.host_total = 0;
    if (rounds_out) *rounds_out = 0;
    if (NC == 0) {
        if (ev) for (int i = 0; i < 5; ++i) cudaEventRecord(ev[i], s);
        cudaMemsetAsync(pb.hap_pairs, 0, sizeof(uint64_t) * (pb.H + 1), s);
        cudaStreamSynchronize(s);
        return 0;
    }
    bool scalar = pv.min_cap == 0xFFFFFFFFu;          // exact sequential fallback
    cudaMemsetAsync(pv.alt_entry, 0xFF, sizeof(uint32_t) * (size_t)NC * kAltSlots, s);
    cudaMemsetAsync(pv.counters, 0, sizeof(uint32_t) * (2 + kMaxRounds), s);
    if (ev) cudaEventRecord(ev[0], s);
    // stage A: first phrase of every chunk, diagonal-run links, list ranking
    g_launches += 2;
    if (scalar) k_first<<<(NC + 127) / 128, 128, 0, s>>>(pv);
    else k_first_w<<<(NC + 127) / 128, 128, 0, s>>>(pv);
    k_link<<<(NC + 127) / 128, 128, 0, s>>>(pv);
    {
        uint32_t *li = pv.link, *ri = pv.run_len, *lo = pb.link_b, *ro = pb.run_b;
        for (uint64_t span = 1; span < pb.max_chunks_per_hap; span *= 2) {
            g_launches += 1;
            k_rank<<<(NC + 255) / 256, 256, 0, s>>>(li, ri, lo, ro, NC);
            uint32_t* t = li; li = lo; lo = t; t = ri; ri = ro; ro = t;
        }
        pv.link = li; pv.run_len = ri;
    }
    if (ev) cudaEventRecord(ev[4], s);
    g_launches += 1;
    if (scalar) k_spec<<<(NC + 127) / 128, 128, 0, s>>>(pv);
    else k_spec_w<<<(NC + 127) / 128, 128, 0, s>>>(pv);
    if (ev) cudaEventRecord(ev[1], s);
    int round = 1;
    for (;;) {
        cudaMemcpyAsync(pb.host_counters, pv.counters, sizeof(uint32_t) * (2 + kMaxRounds), cudaMemcpyDeviceToHost, s);
        cudaStreamSynchronize(s);
        if (pb.host_counters[0]) return 2;
        if (round >= kMaxRounds) return 2;
        if (pb.host_counters[1 + round] == 0) break;
        uint32_t nr = NC * kAltSlots;
        g_launches += 1;
        k_fix<<<(nr + 127) / 128, 128, 0, s>>>(pv, (uint32_t)round);
        ++round;
    }
    if (rounds_out) *rounds_out = round;
    if (ev) cudaEventRecord(ev[2], s);
    uint32_t nn = NC * (kAltSlots + 1);
    g_launches += 3;
    k_succ<<<(nn + 255) / 256, 256, 0, s>>>(pv, pb.jump_a, pb.mark);
    uint32_t* ja = pb.jump_a; uint32_t* jb = pb.jump_b;
    for (uint64_t span = 1; span < 2ull * pb.max_chunks_per_hap; span *= 2) {
        g_launches += 1;
        k_jump<<<(nn + 255) / 256, 256, 0, s>>>(ja, jb, pb.mark, nn);
        uint32_t* t = ja; ja = jb; jb = t;
    }
    k_true<<<(NC + 255) / 256, 256, 0, s>>>(pv, pb.mark, pb.true_cnt, pb.true_entry, pb.true_slot);
    device_scan<uint64_t, LoadArr<uint32_t, uint64_t>, StoreArr<uint64_t>, OpSum, false>(
        LoadArr<uint32_t, uint64_t>{pb.true_cnt}, StoreArr<uint64_t>{pb.true_off}, NC, pb.scan_tmp, OpSum(), 0ull, s);
    uint64_t nt = ((uint64_t)NC + kScanTile - 1) / kScanTile;
    k_hap_pairs<<<(pb.H + 1 + 127) / 128, 128, 0, s>>>(pv, pb.H, pb.true_off, pb.scan_tmp + nt, pb.hap_pairs);
    if (ev) cudaEventRecord(ev[3], s);
    cudaMemcpyAsync(pb.host_total, pb.scan_tmp + nt, sizeof(uint64_t), cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(pb.host_counters, pv.counters, sizeof(uint32_t) * 2, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    if (pb.host_counters[0]) return 2;
    return 0;
}

void parse_stage_emit(ParseBuffers& pb, int64_t* out, cudaStream_t s) {
    uint32_t NC = pb.pv.n_chunks;
    if (NC == 0) return;
    g_launches += 1;
    k_emit<<<(NC + 127) / 128, 128, 0, s>>>(pb.pv, pb.true_cnt, pb.true_entry, pb.true_slot, pb.true_off, out);
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
    if (mt.len == 0) mt.pos = (uint32_t)("ACGT"[base_at(X, i)]);
    ms[i] = mt.len; pos[i] = mt.pos;
}
void launch_ms_dense(const IndexView& ix, const uint64_t* X, uint64_t m, uint32_t* ms, uint32_t* pos, cudaStream_t s) {
    if (m == 0) return;
    g_launches += 1;
    k_ms_dense<<<(unsigned)((m + 127) / 128), 128, 0, s>>>(ix, X, m, ms, pos);
}

}  // namespace drlz
