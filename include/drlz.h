/*
 * drlz.h -- C ABI of libdrlz.so: the B200-native drop-in for PanGenSpark's distributed RLZ parse.
 *
 * The reference has no FFI for this path: its only boundary is the process boundary
 *     spark-submit --class org.ngseq.pangenspark.DistributedRLZ <jar> localradix localref dataGlob fsURI outDir gaplessOut
 * (drlz.sh:15, DistributedRLZ.scala:37-43) and the arithmetic lives in three local closures
 * (binarySearch :108-164, factor :176-217, encode :222-242).  This header is the interface a maintainer binds
 * from Java (JNI / Panama), from the native `drlz` CLI, or from Python (ctypes).
 * Each entry point names the reference code it replaces.
 *
 * Conventions: plain pointers and sizes; the caller owns every buffer; 0 = success, negative = DRLZ_E_*;
 * no exceptions or exit() cross this boundary; one context per GPU (one process per GPU).  A context may be shared
 * by several threads (Spark runs several tasks per executor JVM): calls are serialised on the device by a mutex.
 * There is no CPU fallback: without a CUDA device drlz_create fails.
 */
#ifndef DRLZ_H
#define DRLZ_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRLZ_OK 0
#define DRLZ_E_ARG (-1)        /* bad argument */
#define DRLZ_E_CUDA (-2)       /* CUDA runtime error (see drlz_last_error) */
#define DRLZ_E_NOMEM (-3)
#define DRLZ_E_NOSPACE (-4)    /* pairs_cap too small; required count returned */
#define DRLZ_E_ALPHABET (-5)   /* reference uses all 256 byte values (one code is reserved) */
#define DRLZ_E_STATE (-6)      /* index not built / wrong call order */
#define DRLZ_E_LIMIT (-7)      /* n or m >= 2^32 - 1024, or more than 65535 rows in one batch */
#define DRLZ_E_INTERNAL (-8)

typedef struct drlz_ctx drlz_ctx;

/* ---- lifecycle -------------------------------------------------------------------------------------- */
int drlz_create(drlz_ctx** out, int device);
void drlz_destroy(drlz_ctx* ctx);
const char* drlz_last_error(const drlz_ctx* ctx);
/* Run all kernels of this context on the caller's CUDA stream (cudaStream_t as void*), e.g. torch's current
 * stream, so that caller-side CUDA events bracket them.  NULL restores the context's own stream. */
int drlz_set_stream(drlz_ctx* ctx, void* cuda_stream);
/* options: "kmer" (jump-table k, 0 = auto), "chunk" (bases per parse chunk, rounded down to a power of two, 0 = default 8192),
 *          "group_bytes" (host-path pipeline granularity), "lcp" (1 = also build the LCP array),
 *          "min_cap" (lower bound of the per-lookup compare cap, default 64), "foreign_cap" (longest off-chain
 *          diagonal compare a single thread may do before the exact sequential fallback is taken),
 *          tuning / A-B switches: "occ", "occ_fix", "occ_diag" (min CTAs per SM of k_spec / k_fix / k_diag_w), "predict" (uniqueness-array
 *          prediction), "hint" + "diag" (per-chunk diagonal hint and mismatch pre-scan), "tickets" (1 = atomic tile tickets in k_pack), "occ_pack" (6 | 8 CTAs per SM k_pack is
 *          compiled for), "prefetch" (L2 prefetch distance of k_pack in tiles, 0 = off),
 *          "spec_store", "l2persist", "pack_variant" (2 = k_pack3, 1 = k_pack2, 0 = the round-1 kernels; generic alphabets: 0 = two passes, else k_packg1), "prefetch3",
 *          "spec_sync" (1 = the chunks of a warp parsed in step), "uniq_coarse" (1 = block maxima of the uniqueness array consulted first); "keep_timings" (1: drlz_parse_timings sums over all
 *          calls since the option was set instead of the last call only), "sort_variant" / "table_variant" (scatter kernel of the radix sort, one-pass k-mer table; process-wide),
 *          "lcp_ties" (0 = the LCP array always by the Kasai-order kernel), "l2_fetch" / "l2_fetch_parse" (cudaLimitMaxL2FetchGranularity for
 *          the device / around the parse kernels: 32 | 64 | 128; measured without effect on B200).  None of them changes results. */
int drlz_set_option(drlz_ctx* ctx, const char* key, int64_t value);

/* pinned host memory for the streaming path (cudaHostAlloc); plain malloc memory also works, slower */
void* drlz_host_alloc(size_t bytes);
void drlz_host_free(void* p);

/* ---- index: replaces `./radixSA localref localradix` + the SA text load + both broadcasts -------------
 * (DistributedRLZ.scala:45,52,55,58,61,63).  ref = exactly n sequence bytes (no header, no newline).
 * Any byte alphabet works, like the reference's byte-wise binary search: a pure {A,C,G,T} reference is packed at
 * 2 bits per base (the measured fast path); otherwise the distinct bytes of the reference get dense, order-
 * preserving codes at 4 bits (<= 16 symbols, e.g. ACGTN + lower case) or 8 bits per symbol.  Haplotype bytes
 * that do not occur in the reference become literal phrases, as in the reference.  The build starts with the host->device
 * copy of ref: from drlz_host_alloc memory it runs at PCIe speed, from pageable memory the driver stages it (3 x slower). */
int drlz_build_index(drlz_ctx* ctx, const uint8_t* ref, uint64_t n);
/* copy index arrays to the host.  what: 0 = SA (uint32[n]), 1 = inverse SA (uint32[n]), 2 = LCP (uint32[n],
 * needs option lcp=1), 3 = k-mer table (uint32[sigma^k+1]), 4 = packed text (uint64[ceil(n * bits / 64)]),
 * 5 = uniqueness array (uint16[n]: the longest repeat that starts at each reference position, saturated at 65535).
 * The inverse SA is derived from SA on the first request if the build did not need it.  */
int drlz_index_get(drlz_ctx* ctx, int what, void* host_out, uint64_t bytes);
/* info[0]=n, [1]=k | log2(bits per symbol) << 8 | alphabet size sigma << 16 | (long repeats: 1) << 28, [2]=prefix-doubling rounds,
 * [3]=index bytes on device */
int drlz_index_info(const drlz_ctx* ctx, uint64_t info[4]);
/* time of the last drlz_build_index in ms (CUDA events on the context stream): [0] total, from before the host->device
 * copy of the reference to the last kernel, [1] copy + alphabet scan + pack, [2] suffix array, [3] k-mer table, [4] lcp */
int drlz_index_timings(const drlz_ctx* ctx, double ms[5]);

/* Multi-GPU: rank 0 builds, the others call drlz_index_reserve(n, info[1] of rank 0) and receive the three
 * buffers by an NCCL broadcast (done by the caller, e.g. torch.distributed) into the device pointers returned by
 * drlz_index_buffers, then call drlz_index_commit.  bufs: [0] packed text + uniqueness array + byte<->code
 * tables, [1] SA (followed by the unsaturated uniqueness array when the reference has repeats of 65535 symbols and more: bit
 * 28 of kinfo), [2] k-mer table.  kinfo = k | log2(bits per symbol) << 8 | alphabet size << 16 | flags, exactly
 * drlz_index_info's info[1]; a plain k (upper bits 0) means the 2-bit {A,C,G,T} layout.  drlz_index_commit waits for the
 * device (the buffers may have been filled on the caller's streams) and derives what is not broadcast (the block maxima
 * of the uniqueness array). */
int drlz_index_reserve(drlz_ctx* ctx, uint64_t n, int kinfo);
int drlz_index_buffers(drlz_ctx* ctx, void* dev_ptrs[3], uint64_t bytes[3]);
int drlz_index_commit(drlz_ctx* ctx);

/* ---- parse: replaces the gap strip (:72,86) + encode (:222-242) + record encoding (:263-284) ----------
 * One MSA row (one haplotype).  strip_gaps != 0 removes every '-' first.  A trailing "\n" / "\r\n" is ignored
 * (Spark's text reader drops it).  pairs_out receives n_pairs records of (int64 pos, int64 len), i.e. the exact
 * bytes of <outDir>/<basename>.lz on a little-endian host.  gapless_out (nullable, >= cols bytes) receives the
 * gapless sequence (the body of the gapless FASTA record, :75-80).  If pairs_cap is too small the call
 * returns DRLZ_E_NOSPACE with the required count in *n_pairs. */
int drlz_parse(drlz_ctx* ctx, const uint8_t* msa_row, uint64_t cols, int strip_gaps, uint8_t* gapless_out,
               uint64_t* m_out, int64_t* pairs_out, uint64_t pairs_cap, uint64_t* n_pairs);

/* H rows at once (the data-parallel axis of the Spark job, :247-248).  Phrase records of row h start at
 * pairs_out + 2 * (n_pairs[0] + ... + n_pairs[h-1]).  gapless_out (nullable) is an array of H row buffers
 * (entries may be NULL).  Host->device copies of the rows are pipelined with the kernels. */
int drlz_parse_batch(drlz_ctx* ctx, const uint8_t* const* rows, const uint64_t* cols, uint32_t H, int strip_gaps,
                     uint8_t* const* gapless_out, uint64_t* m_out, int64_t* pairs_out, uint64_t pairs_cap,
                     uint64_t* n_pairs);

/* Asynchronous form of drlz_parse_batch for streaming producers: the call returns at once with a
 * ticket; one worker thread per context runs the submitted batches in FIFO order.  Every buffer (including the
 * rows/cols arrays) must stay valid and untouched until drlz_wait(ticket) returns that batch's status code. */
int drlz_submit(drlz_ctx* ctx, const uint8_t* const* rows, const uint64_t* cols, uint32_t H, int strip_gaps,
                uint8_t* const* gapless_out, uint64_t* m_out, int64_t* pairs_out, uint64_t pairs_cap, uint64_t* n_pairs,
                uint64_t* ticket);
int drlz_wait(drlz_ctx* ctx, uint64_t ticket);     /* a ticket can be collected once; an unknown or collected one => DRLZ_E_ARG */

/* drlz_parse_batch / drlz_submit that also return the gap-position side table of every row:
 * one (int64 first column, int64 length) record per maximal run of '-' in the row, in column order -- what
 * ScoreMatrix.scala:89-127 recomputes from the rows (`gap`, `consecutive`) to map gapless positions back to MSA
 * columns.  Records of row h start at gaps_out + 2 * (n_gaps[0] + ... + n_gaps[h-1]); too small a gaps_cap =>
 * DRLZ_E_NOSPACE with the required counts in n_gaps.  strip_gaps == 0 => no gaps by definition. */
int drlz_parse_batch_gaps(drlz_ctx* ctx, const uint8_t* const* rows, const uint64_t* cols, uint32_t H, int strip_gaps,
                          uint8_t* const* gapless_out, uint64_t* m_out, int64_t* pairs_out, uint64_t pairs_cap,
                          uint64_t* n_pairs, int64_t* gaps_out, uint64_t gaps_cap, uint64_t* n_gaps);
int drlz_submit_gaps(drlz_ctx* ctx, const uint8_t* const* rows, const uint64_t* cols, uint32_t H, int strip_gaps,
                     uint8_t* const* gapless_out, uint64_t* m_out, int64_t* pairs_out, uint64_t pairs_cap, uint64_t* n_pairs,
                     int64_t* gaps_out, uint64_t gaps_cap, uint64_t* n_gaps, uint64_t* ticket);

/* Same with the rows already resident in device memory: row h = rows_dev + row_off[h]; every row_off is a
 * multiple of 16 and the buffer is readable up to the next multiple of 16 past each row's end; cols excludes
 * any line terminator.  The phrase records stay on the device: *pairs_dev points into context-owned memory
 * valid until the next parse call (layout as in drlz_parse_batch). */
int drlz_parse_batch_device(drlz_ctx* ctx, const uint8_t* rows_dev, const uint64_t* row_off, const uint64_t* cols,
                            uint32_t H, int strip_gaps, uint64_t* m_out, const int64_t** pairs_dev,
                            uint64_t* n_pairs);

/* Copy bytes from device memory owned by the context (e.g. *pairs_dev) to the host, on the context's stream. */
int drlz_copy_to_host(drlz_ctx* ctx, void* dst_host, const void* src_dev, uint64_t bytes);

/* Per-kernel timings (ms, CUDA events on the context's stream) summed over the groups of the last parse call:
 * [0] zero-fill of the packed buffer, [1] pre-scan (k_hint + k_diag_w), [2] host time of the last group outside the GPU work (before its first launch + after its last synchronisation), [3] k_pack (single pass), [4] k_spec (speculative
 * parse), [5] k_fix rounds, [6] pointer jumping + compaction scan, [7] k_emit (incl. the host read of the phrase
 * total), [8] whole device pipeline, [9] stage A (k_link + list ranking + k_resolve), [10] fix-up rounds executed,
 * [11] kernels launched. */
int drlz_parse_timings(const drlz_ctx* ctx, double ms[12]);

/* Dense matching statistics of one gapless, pure-ACGT sequence in host memory: for every position i
 * ms[i] = longest match length of x[i..] in the reference and pos[i] = SA[leftmost index of its interval]
 * (literal => ms 0, pos = byte value).  Diagnostic / validation entry point. */
int drlz_matching_stats(drlz_ctx* ctx, const uint8_t* x, uint64_t m, uint32_t* ms_out, uint32_t* pos_out);

#ifdef __cplusplus
}
#endif
#endif
