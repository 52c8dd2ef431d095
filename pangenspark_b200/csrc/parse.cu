// parse.cu -- kernels of the chunk-parallel exact greedy RLZ parse (per-thread logic in rlz_core.cuh).
//
//   k_spec      P1: one thread per chunk, speculative greedy parse from the chunk's first base with chunk-capped
//                   lookups; phrases stored; the first phrase doubles as stage A's phrase at the chunk start
//   k_link      A : open first phrases linked to the chunk their match continues in
//   k_rank      A : list ranking (pointer jumping) -> exact lengths of arbitrarily long matches
//   k_resolve   A : total length of each chunk's open last phrase -> exit of the chunk
//   k_fix       F : one thread per (chunk, alternative entry) created in the previous round
//   k_succ      J : successor of every record; roots marked
//   k_jump      J : pointer jumping (mark[jump[v]] |= mark[v]; jump = jump o jump), log2(chunks) rounds
//   k_true      E : per chunk the marked record's phrase count, entry and slot
//   device_scan E : exclusive prefix of the counts = output slot of every chunk (stream compaction)
//   k_emit      E : gather the stored phrases of the true records into (int64 pos, int64 len) records
//                   (DistributedRLZ.scala:263-284 record layout); re-parses only what was not stored
//
// Two warp-synchronous variants of k_spec (a lock-step per-lane state machine, and cooperative 1024-base
// extension) were built and measured slower (2.7 ms / 1.6 ms vs 1.2 ms per step): the parse is bound by chains
// of dependent cache misses, and divergent scalar lanes overlap those chains better (profiles/README.md).
#include "drlz_internal.h"
#include "scan.cuh"

namespace drlz {

__global__ void __launch_bounds__(128) k_link(ParseView pv) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < pv.n_chunks) link_chunk(pv, g);
}

__global__ void k_rank(const uint32_t* __restrict__ link_in, const uint32_t* __restrict__ run_in, uint32_t* link_out,
                       uint32_t* run_out, uint32_t nc) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < nc) rank_round(link_in, run_in, link_out, run_out, g);
}

__global__ void __launch_bounds__(128) k_resolve(ParseView pv) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < pv.n_chunks) resolve_chunk(pv, g);
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
    cudaMemsetAsync(pv.alt_entry, 0xFF, sizeof(uint32_t) * (size_t)NC * kAltSlots, s);
    cudaMemsetAsync(pv.counters, 0, sizeof(uint32_t) * (2 + kMaxRounds), s);
    if (ev) cudaEventRecord(ev[0], s);
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
    {
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
        else if (pb.occ >= 10) k_fix<10><<<(nr + 127) / 128, 128, 0, s>>>(pv, (uint32_t)round);
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
