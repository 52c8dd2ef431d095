// parse.cu -- kernels of the chunk-parallel exact greedy RLZ parse (scalar logic in rlz_core.cuh).
//
//   k_chunk_machine<true>  A : phrase at the first base of every chunk, compared no further than the chunk end
//   k_link      A : open first phrases linked to the chunk their match continues in
//   k_rank      A : list ranking (pointer jumping) -> exact lengths of arbitrarily long matches
//   k_chunk_machine<false> P1: one lane per chunk, speculative greedy parse from the chunk's first base; phrases stored
//   k_fix       F : one thread per (chunk, alternative entry) created in the previous round
//   k_succ      J : successor of every record; roots marked
//   k_jump      J : pointer jumping (mark[jump[v]] |= mark[v]; jump = jump o jump), log2(chunks) rounds
//   k_true      E : per chunk the marked record's phrase count, entry and slot
//   device_scan E : exclusive prefix of the counts = output slot of every chunk (stream compaction)
//   k_emit      E : gather the stored phrases of the true records into (int64 pos, int64 len) records
//                   (DistributedRLZ.scala:263-284 record layout); re-parses only what was not stored
//
// k_chunk_machine is the scalar lookup()/spec_parse_chunk() of rlz_core.cuh re-expressed as a per-lane state
// machine inside one warp-uniform loop (ncu on the scalar version: 4.4 of 32 lanes active per instruction,
// see profiles/README.md).
#include "drlz_internal.h"
#include "scan.cuh"

namespace drlz {

#define FULLMASK 0xffffffffu

// Persistent per-lane state machine: one lane = one chunk.  Every iteration of the single loop below each lane
// does at most ONE suffix probe (64 bases from two packed words) or ONE extension slice (128 bases), all lanes at
// the same program locations, so lanes in different stages of different phrases do not serialise each other and
// their memory requests overlap.  kFirstOnly = stage A (one capped lookup at the chunk start); otherwise P1
// (speculative parse of the whole chunk, phrases stored for emission).
template <bool kFirstOnly>
__global__ void __launch_bounds__(128) k_chunk_machine(ParseView pv) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = pv.ix.k;
    const uint64_t n = pv.ix.n;
    const uint64_t* __restrict__ D = pv.ix.text;
    const uint32_t* __restrict__ SA = pv.ix.sa;
    const uint32_t* __restrict__ TAB = pv.ix.tab;
    enum { START = 0, BS, PL1, PL2, SINGLE, LEFT, FINISH, DONE };
    int phase = DONE;
    // chunk state
    uint32_t h = 0, cnt = 0, cur = 0;
    uint64_t s = 0, e = 0, i = 0;
    HapView hv; hv.x = nullptr; hv.m = 0; hv.exc_pos = nullptr; hv.exc_byte = nullptr; hv.n_exc = 0;
    // lookup state
    uint64_t xw0 = 0, xw1 = 0, kmer = 0;
    uint32_t maxlen = 0, plen = 0, kk = 0, lo = 0, l = 0, hh = 0, r = 0, l1 = 0, l2 = 0, lcp_left = 0, lcp_right = 0, sar = 0,
             a = 0, b = 0, L = 0, pidx = 0, plim_pat = 0;
    bool have_left = false, have_right = false, open = false, lit_exc = false;
    Match mt; mt.len = 0; mt.pos = 0;
    // extension state
    bool in_ext = false;
    uint32_t ext_c = 0, ext_lim = 0, ext_p = 0;

    if (g < pv.n_chunks) {
        h = pv.chunk_hap[g];
        hv = pv.haps[h];
        s = (uint64_t)(g - pv.hap_chunk_base[h]) * pv.chunk_bases;
        e = s + pv.chunk_bases; if (e > hv.m) e = hv.m;
        i = s;
        if (kFirstOnly) {
            pv.link[g] = kEmpty;
            if (s >= hv.m) { pv.first_pos[g] = 0; pv.first_len[g] = 0; pv.run_len[g] = 0; }
            else { cur = hv.n_exc ? exc_lower_bound(hv, s) : 0; phase = START; }
        } else {
            if (i < e) {                               // first phrase: stage A's result
                uint32_t fp = pv.first_pos[g], fl = pv.run_len[g];
                pv.spec_ph[(size_t)g * kSpecStore] = make_uint2(fp, fl);
                cnt = 1;
                i += fl ? fl : 1;
                cur = hv.n_exc ? exc_lower_bound(hv, i) : 0;
            }
            if (i < e) phase = START;
            else { pv.spec_cnt[g] = cnt; pv.spec_exit[g] = (uint32_t)i; register_exit(pv, h, i, 1); }
        }
    }

    while (__any_sync(FULLMASK, phase != DONE)) {
        // ---- per-lane transitions until a probe is needed ---------------------------------------------------
        bool need = false;
        if (phase != DONE && !in_ext) {
            for (;;) {
                if (phase == START) {
                    open = false; lit_exc = false;
                    uint64_t stop = stop_after(hv, i, &cur);
                    if (stop == i) { mt.len = 0; mt.pos = hv.exc_byte[cur]; lit_exc = true; phase = FINISH; }
                    else {
                        maxlen = (uint32_t)(stop - i);
                        uint64_t C = pv.chunk_bases;
                        uint64_t to_end = (i / C + 1) * C - i;
                        uint32_t cap = to_end > (uint64_t)pv.min_cap ? (to_end > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)to_end) : pv.min_cap;
                        plen = cap < maxlen ? cap : maxlen;
                        xw0 = get64(hv.x, i);
                        xw1 = get64(hv.x, i + 32);
                        kmer = xw0 >> (64 - 2 * k);
                        kk = plen < (uint32_t)k ? plen : (uint32_t)k;
                        uint64_t lowmask = (1ull << (2 * (k - (int)kk))) - 1;
                        lo = TAB[kmer & ~lowmask];
                        l = lo; hh = TAB[(kmer | lowmask) + 1];
                        have_left = have_right = false;
                        phase = BS;
                    }
                }
                if (phase == BS) {
                    if (l < hh) { pidx = l + ((hh - l) >> 1); plim_pat = plen; need = true; break; }
                    r = l; phase = PL1;
                }
                if (phase == PL1) {
                    if (r > 0 && !have_left) { pidx = r - 1; plim_pat = plen; need = true; break; }
                    l1 = r > 0 ? lcp_left : 0; phase = PL2;
                }
                if (phase == PL2) {
                    if ((uint64_t)r < n && !have_right) { pidx = r; plim_pat = plen; need = true; break; }
                    l2 = (uint64_t)r < n ? lcp_right : 0;
                    if (l2 > l1) {
                        if (l2 == plen && plen < maxlen) phase = SINGLE;
                        else { mt.len = l2; mt.pos = sar; phase = FINISH; }
                    } else if (l1 == 0) { mt.len = 0; mt.pos = 0; phase = FINISH; }
                    else {
                        L = l1;
                        uint32_t Lk = L < (uint32_t)k ? L : (uint32_t)k;
                        a = lo;
                        if (Lk != kk) { uint64_t m2 = (1ull << (2 * (k - (int)Lk))) - 1; a = TAB[kmer & ~m2]; }
                        b = r - 1; phase = LEFT;
                    }
                }
                if (phase == SINGLE) {
                    if ((uint64_t)r + 1 < n) { pidx = r + 1; plim_pat = plen; need = true; break; }
                    open = true; mt.len = l2; mt.pos = sar; phase = FINISH;
                }
                if (phase == LEFT) {
                    if (a < b) { pidx = a + ((b - a) >> 1); plim_pat = L; need = true; break; }
                    mt.len = L; mt.pos = SA[a]; phase = FINISH;
                }
                if (phase == FINISH) {
                    if (kFirstOnly) {
                        if (mt.len == 0 && !lit_exc) mt.pos = (uint32_t)("ACGT"[base_at(hv.x, i)]);
                        pv.first_pos[g] = mt.pos;
                        pv.first_len[g] = mt.len;
                        pv.run_len[g] = open ? kEmpty : mt.len;
                        phase = DONE;
                        break;
                    }
                    if (mt.len == 0) { if (!lit_exc) mt.pos = (uint32_t)("ACGT"[base_at(hv.x, i)]); }
                    else if (open) {
                        uint64_t j; bool same;
                        uint32_t g2 = open_target(pv, h, i, mt, &j, &same);
                        uint32_t rest = same ? pv.run_len[g2] : foreign_run(pv, hv, i, mt, j);
                        mt.len = (uint32_t)(j - i) + rest;
                    }
                    if (cnt < (uint32_t)kSpecStore) pv.spec_ph[(size_t)g * kSpecStore + cnt] = make_uint2(mt.pos, mt.len);
                    ++cnt;
                    i += mt.len ? mt.len : 1;
                    if (i < e) { phase = START; continue; }
                    pv.spec_cnt[g] = cnt;
                    pv.spec_exit[g] = (uint32_t)i;
                    register_exit(pv, h, i, 1);
                    phase = DONE;
                }
                break;
            }
        }
        // ---- one probe: 64 bases from two packed words -------------------------------------------------------
        uint32_t c = 0, lim = 0;
        bool less = false, ready = false;
        if (need) {
            uint32_t p = SA[pidx];
            uint64_t rem = n - p;
            lim = rem < (uint64_t)plim_pat ? (uint32_t)rem : plim_pat;
            uint64_t b0 = get64(D, p);
            uint64_t x0 = xw0 ^ b0;
            ext_p = p;
            if (x0) { c = (uint32_t)(clz64(x0) >> 1); less = b0 < xw0; }
            else {
                c = 32;
                if (lim > 32) {
                    uint64_t b1 = get64(D, (uint64_t)p + 32);
                    uint64_t x1 = xw1 ^ b1;
                    if (x1) { c = 32 + (uint32_t)(clz64(x1) >> 1); less = b1 < xw1; }
                    else c = 64;
                }
            }
            if (c >= lim) { c = lim; less = (lim != plim_pat); ready = true; }
            else if (c == 64) { in_ext = true; ext_c = 64; ext_lim = lim; }
            else ready = true;
        }
        // ---- one extension slice: up to 128 more bases of a match that ran past the probe ----------------------
        if (in_ext) {
            uint64_t xa[4], xb[4];
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                xa[w] = get64(hv.x, i + ext_c + 32u * w);
                xb[w] = get64(D, (uint64_t)ext_p + ext_c + 32u * w);
            }
            bool fin = false, lessb = false;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                if (!fin) {
                    uint64_t xr = xa[w] ^ xb[w];
                    if (xr) { ext_c += (uint32_t)(clz64(xr) >> 1); lessb = xb[w] < xa[w]; fin = true; }
                    else { ext_c += 32; if (ext_c >= ext_lim) fin = true; }
                }
            }
            if (fin) {
                if (ext_c >= ext_lim) { c = ext_lim; less = (ext_lim != plim_pat); }
                else { c = ext_c; less = lessb; }
                in_ext = false; ready = true;
            }
        }
        // ---- consume the finished compare ----------------------------------------------------------------------
        if (ready) {
            if (phase == BS) {
                if (less) { l = pidx + 1; lcp_left = c; have_left = true; }
                else { hh = pidx; lcp_right = c; have_right = true; sar = ext_p; }
            } else if (phase == PL1) { lcp_left = c; have_left = true; }
            else if (phase == PL2) { lcp_right = c; have_right = true; sar = ext_p; }
            else if (phase == SINGLE) {
                if (c >= plen) {            // several suffixes share the capped pattern: redo uncapped (exact)
                    plen = maxlen;
                    kk = plen < (uint32_t)k ? plen : (uint32_t)k;
                    uint64_t lowmask = (1ull << (2 * (k - (int)kk))) - 1;
                    lo = TAB[kmer & ~lowmask];
                    l = lo; hh = TAB[(kmer | lowmask) + 1];
                    have_left = have_right = false;
                    phase = BS;
                } else { open = true; mt.len = l2; mt.pos = sar; phase = FINISH; }
            } else if (phase == LEFT) { if (c >= L) b = pidx; else a = pidx + 1; }
        }
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

// scalar versions (used by the exact sequential fallback, where one chunk = one haplotype)
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_first(ParseView pv) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < pv.n_chunks) first_phrase_chunk(pv, g);
}
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_spec(ParseView pv) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < pv.n_chunks) spec_parse_chunk(pv, g);
}
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_fix(ParseView pv, uint32_t round) {
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
    bool scalar = pv.min_cap == 0xFFFFFFFFu || !pb.use_machine;   // exact sequential fallback, or scalar kernels by choice
    cudaMemsetAsync(pv.alt_entry, 0xFF, sizeof(uint32_t) * (size_t)NC * kAltSlots, s);
    cudaMemsetAsync(pv.counters, 0, sizeof(uint32_t) * (2 + kMaxRounds), s);
    if (ev) cudaEventRecord(ev[0], s);
    // stage A: first phrase of every chunk, diagonal-run links, list ranking
    g_launches += 2;
    if (scalar && pb.occ >= 16) k_first<16><<<(NC + 127) / 128, 128, 0, s>>>(pv);
    else if (scalar) k_first<1><<<(NC + 127) / 128, 128, 0, s>>>(pv);
    else k_chunk_machine<true><<<(NC + 127) / 128, 128, 0, s>>>(pv);
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
    if (scalar && pb.occ >= 16) k_spec<16><<<(NC + 127) / 128, 128, 0, s>>>(pv);
    else if (scalar) k_spec<1><<<(NC + 127) / 128, 128, 0, s>>>(pv);
    else k_chunk_machine<false><<<(NC + 127) / 128, 128, 0, s>>>(pv);
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
        if (pb.occ >= 16) k_fix<16><<<(nr + 127) / 128, 128, 0, s>>>(pv, (uint32_t)round);
        else k_fix<1><<<(nr + 127) / 128, 128, 0, s>>>(pv, (uint32_t)round);
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
