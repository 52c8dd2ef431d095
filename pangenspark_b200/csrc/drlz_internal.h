// drlz_internal.h -- host-side declarations shared by the .cu translation units of libdrlz.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "rlz_core.cuh"

namespace drlz {

extern thread_local unsigned long long g_launches;

static const int kPackGTileBytes = 4096;     // tile of the generic-alphabet pack kernels (256 threads x 16 B)
static const int kPackTileBytes = 16384;     // MSA bytes per CTA in the strip/pack kernel (256 threads x 64 B; 8 KB and 32 KB measured slower)

// ---- strip + pack (pack.cu) --------------------------------------------------------------------
struct PackJob {
    const uint8_t* rows;          // device; row h starts at rows + row_off[h] (16-byte aligned, readable up to
                                  // the next multiple of 16 past cols[h])
    const uint64_t* row_off;      // [H] device
    const uint64_t* cols;         // [H] device
    uint32_t H;
    uint32_t tiles_per_row;       // ceil(max cols / kPackTileBytes)
    int strip;                    // 1: '-' is a gap column (DistributedRLZ.scala:72,86); 0: keep it (a literal)
    uint64_t* desc_cnt;           // [H*tiles_per_row] look-back descriptors: non-gap bytes   (zeroed by launch_pack)
    uint64_t* desc_exc;           // [H*tiles_per_row] look-back descriptors: exceptions
    uint32_t* tile_counter;       // [1] dynamic tile ids
    int occ;                      // min CTAs per SM the kernel is compiled for (6 or 8)
    uint32_t prefetch;            // L2 prefetch distance in tiles (0 = off): CTA i prefetches the bytes of tile i + prefetch
    int use_tickets;              // 1: tile ids from the atomic counter; 0: from blockIdx (dispatch order), guarded by a timeout
    int variant;                  // 0: k_pack (round 1), 1: k_pack2 (lean plain-warp path; exception counts by atomics), 2: k_pack3 (a CTA streams 8 tiles)
    uint32_t uniform_cols;        // != 0: every row has this many columns and row h starts at rows + uniform_off0 + h * uniform_stride (no per-row loads)
    uint64_t uniform_stride, uniform_off0;
    uint64_t* X;                  // packed output, zero-filled by the caller
    const uint64_t* xoff;         // [H] word offset of haplotype h inside X (device)
    uint64_t* m_out;              // [H] gapless lengths (device)
    uint64_t* exc_cnt;            // [H] exceptions per row (device)
    uint32_t* exc_pos;            // [H*exc_stride] row h owns [h*exc_stride, (h+1)*exc_stride)
    uint8_t* exc_byte;            // [H*exc_stride]
    uint64_t exc_stride;
    uint32_t* flags;              // bit 0: a row has more exceptions than exc_stride; bit 1: look-back timed out
    // generic alphabets (bl > 1): byte -> code table and the buffers of the two-pass generic kernels
    uint32_t bl;                  // log2(bits per symbol) of the packed output: 1 (ACGT fast path), 2 or 3
    const uint8_t* b2c;           // [256] byte -> code, 0xFF = not in the reference alphabet (device)
    uint32_t gtiles_per_row;      // ceil(max cols / kPackGTileBytes)
    uint32_t* gtile_cnt;          // [H*gtiles_per_row]
    uint32_t* gtile_exc;          // [H*gtiles_per_row]
    uint64_t* gtile_off;          // [H*gtiles_per_row]
    uint64_t* gtile_excoff;       // [H*gtiles_per_row]
    uint64_t* gscan_tmp;          // 2 * scan_tmp_elems(H*gtiles_per_row)
    uint8_t* gapless;             // optional gapless bytes out (device) or null
    const uint64_t* gapless_off;  // [H] byte offsets into gapless
};
// ev (nullable): 3 events: [0],[1] before the kernel, [2] after it
void launch_pack(const PackJob& job, cudaStream_t s, cudaEvent_t* ev);

// gap runs of the rows of a group as unsorted events (see k_gap_events); *counter = events found (may exceed cap)
void launch_gap_events(const uint8_t* rows, const uint64_t* row_off, const uint64_t* cols, uint32_t H, uint64_t maxcols,
                       uint64_t* list, uint64_t cap, unsigned long long* counter, cudaStream_t s);
// sorts n 64-bit keys in place over bits [0, end_bit) (hand-written LSD radix sort of sa_build.cu); scratch from the caller
size_t sort_keys_scratch_bytes(uint64_t n);
int sort_keys_u64(uint64_t* keys, uint64_t n, int end_bit, void* scratch, cudaStream_t s);

// presence[8]: bit b of the 256-bit set = byte value b occurs in bytes[0, n) (bytes 16-byte aligned)
void launch_byte_presence(const uint8_t* bytes, uint64_t n, uint32_t* presence, cudaStream_t s);

// ---- suffix array + k-mer table (sa_build.cu) ---------------------------------------------------
// Scratch of build_suffix_array (sort keys / values, active-slot lists, histograms: ~32 bytes per reference base).
// Kept by the caller between builds, grow-only, so that only the first build of a context allocates; cudaMalloc and
// cudaFree of gigabytes otherwise sit inside the build (measured: 14 ms of kernels became 34 - 101 ms on some boxes).
struct SaWorkspace {
    void* p[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t cap[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    void* host = nullptr;                  // 8 pinned bytes for the per-round read-back
    cudaError_t ensure(int i, size_t bytes) {
        if (bytes <= cap[i]) return cudaSuccess;
        if (p[i]) cudaFree(p[i]);
        p[i] = nullptr; cap[i] = 0;
        cudaError_t e = cudaMalloc(&p[i], bytes + 256);
        if (e == cudaSuccess) cap[i] = bytes;
        return e;
    }
    size_t bytes() const { size_t t = 0; for (int i = 0; i < 9; ++i) t += cap[i]; return t; }
    void release() {
        for (int i = 0; i < 9; ++i) { if (p[i]) cudaFree(p[i]); p[i] = nullptr; cap[i] = 0; }
        if (host) cudaFreeHost(host);
        host = nullptr;
    }
};
// Builds SA (uint32[n]) of the packed text and, when isa != null, the inverse.  Returns 0 or a cudaError.
// bl = log2(bits per symbol) of the packed text
// ws: scratch kept by the caller (null = allocated and freed inside the call)
// lcp_from_keys (nullable, uint32[n]): the LCP array as read off the first-round sort keys; final iff *rounds_out == 0
// (every suffix was distinguished by its first key), otherwise build_lcp must recompute it
int build_suffix_array(const uint64_t* text, uint64_t n, uint32_t bl, uint32_t* sa, uint32_t* isa, cudaStream_t s,
                       int* rounds_out, char* err, int errlen, SaWorkspace* ws = nullptr, uint32_t* lcp_from_keys = nullptr,
                       uint16_t* uniq_fast = nullptr, int* fast_out = nullptr, const IndexView* tab_ix = nullptr, uint32_t* tab = nullptr,
                       uint32_t* tab_tmp = nullptr);
// tab_ix (k, bl, sigma are read) + tab + tab_tmp (kmer_table_tmp_bytes): in that same pass the k-mer table is written too when its
// key is a head of the sort key; *fast_out = 2 says so (1: SA / LCP / uniqueness array only, build_kmer_table is still to run)
// uniq_fast + fast_out (both or neither): when the first sort leaves no two equal keys, SA, lcp_from_keys and the uniqueness
// array uniq_fast[n] are finished in one pass, *fast_out = 1 and ISA IS NOT WRITTEN (launch_isa_from_sa derives it)
void launch_isa_from_sa(const uint32_t* sa, uint64_t n, uint32_t* isa, cudaStream_t s);
// tab: sigma^k + 1 entries (ix.tab is ignored, ix.text / ix.sa / ix.k / ix.bl / ix.sigma are used)
// tmp: kmer_table_tmp_bytes(entries) bytes of scratch owned by the caller; the call does not synchronise
extern int g_table_variant;               // sa_build.cu: 1 = the table in one pass (k_tab_build), 0 = fill + mark + scan
size_t kmer_table_tmp_bytes(uint64_t entries);
int build_kmer_table(const IndexView& ix, uint32_t* tab, uint32_t* tmp, cudaStream_t s);
// LCP[j] = lcp(suffix SA[j-1], suffix SA[j]), LCP[0] = 0
int build_lcp(const uint64_t* text, uint64_t n, uint32_t bl, const uint32_t* sa, const uint32_t* isa, uint32_t* lcp,
              cudaStream_t s);
// the same array from lcp_from_keys of build_suffix_array when only a few doubling rounds ran: the entries equal to S0
// (pairs with equal first keys) are compared on from symbol S0, all others are already exact
int build_lcp_ties(const uint64_t* text, uint64_t n, uint32_t bl, const uint32_t* sa, uint32_t* lcp, cudaStream_t s);
// uniq[p] = min(65535, max(LCP[ISA[p]], LCP[ISA[p]+1])); uniq32[p] = the same unsaturated; *saturated (device) is set to 1 if
// any value reached 65535
int build_uniq(uint64_t n, const uint32_t* isa, const uint32_t* lcp, uint16_t* uniq, uint32_t* uniq32, uint32_t* saturated, cudaStream_t s);
int build_uniq_coarse(uint64_t n, const uint16_t* uniq, uint16_t* uniqc, cudaStream_t s);   // max of uniq over blocks of kUniqBlock positions

// ---- parse pipeline (parse.cu) -----------------------------------------------------------------
struct ParseBuffers {
    ParseView pv;                 // device pointers (see rlz_core.cuh)
    uint32_t H;
    uint32_t max_chunks_per_hap;
    int occ;                      // min CTAs per SM of k_spec (0 = compiler default; 8 / 10 / 12 / 16)
    int occ_fix;                  // same for k_fix
    int occ_diag;                 // same for k_diag_w (16 / 12 / 10)
    uint64_t* res; void* host_res; size_t res_bytes;   // result block (layout at k_results) and its pinned host mirror
    const uint64_t* m_dev; const uint64_t* exccnt_dev; const uint32_t* flags_dev;   // what k_pack left for it
    int64_t* spec_out;            // where parse_stage_count may emit the records right away (null = do not)
    uint64_t spec_out_cap;        // its capacity in records
    cudaEvent_t ev_emitted;       // recorded after that emission (nullable)
    bool emitted;                 // out: the records were emitted (valid if the total fits spec_out_cap)
    int blind_rounds;             // fix-up rounds launched without reading the counters back (see parse_stage_count)
    int64_t* hint;                // [ceil(NC / kHintGroup)] diagonal hints (null = no hint pass)
    uint16_t* mm;                 // [NC*kMmStride] mismatch lists (null = no diagonal pre-scan)
    uint32_t* ext;                // [NC] extents across chunks, hint_src [NC]: scratch of the hint fill (both only used when the index has uniq32)
    uint32_t* hint_src;
    uint32_t* link_b;             // [NC] second buffers of the list ranking
    uint32_t* run_b;              // [NC]
    uint32_t* jump_a;             // [NC*(kAltSlots+1)]
    uint32_t* jump_b;
    uint8_t* mark;                // [NC*(kAltSlots+1)]
    uint32_t* true_cnt;           // [NC]
    uint32_t* true_entry;         // [NC]
    uint32_t* true_slot;          // [NC] 0 = speculative record, 1+s = alternative record s
    uint64_t* true_off;           // [NC]
    uint64_t* scan_tmp;           // scan_tmp_elems(NC)
    uint32_t* host_counters;      // pinned host mirror of pv.counters (2 + kMaxRounds)
    uint64_t* host_total;         // pinned host: total phrase count
};
// Stage A: speculative parse, fix-up rounds, pointer jumping, per-chunk counts and offsets.
// Synchronises the stream (reads round counters and the phrase total).  Returns 0, or 2 on overflow.
// ev (nullable): 6 events: [0] start, [5] after the pre-scan (k_hint, k_diag_w), [1] after P1, [4] after stage A,
// [2] after the fix-up rounds, [3] after compaction.
int parse_stage_count(ParseBuffers& pb, cudaStream_t s, int* rounds_out, cudaEvent_t* ev);
// Stage B: emission into `out` (device, >= 2*total int64).
void parse_stage_emit(ParseBuffers& pb, int64_t* out, cudaStream_t s);

void launch_make_hapviews(HapView* haps, const uint64_t* X, const uint64_t* xoff, const uint64_t* m,
                          const uint32_t* exc_pos, const uint8_t* exc_byte, const uint64_t* exc_cnt, uint64_t exc_stride,
                          uint32_t H, cudaStream_t s);

}  // namespace drlz

namespace drlz {
void launch_fill_chunk_hap(uint32_t* chunk_hap, const uint32_t* hap_chunk_base, uint32_t H, uint32_t NC, cudaStream_t s);
void launch_ms_dense(const IndexView& ix, const uint64_t* X, uint64_t m, uint32_t* ms, uint32_t* pos, cudaStream_t s);
}
