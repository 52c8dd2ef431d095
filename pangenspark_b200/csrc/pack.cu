// pack.cu -- gap strip + 2-bit pack + exception list for MSA rows (DistributedRLZ.scala:72,86), one pass.
//
// A row is a gapped MSA line; every '-' is removed (stream compaction) and the remaining bases are packed
// 2 bits each, 32 per 64-bit word, first base in the top bits.  Bytes outside {A,C,G,T} are stored as code 0
// and listed (position, byte) in a sorted per-row exception list -- they become literal phrases.
//
// k_pack: one CTA per 16 KB tile, 64 bytes per thread (four 128-bit loads).  The tile's output offset (number of
// non-gap bytes before it in the row) comes from a decoupled look-back over per-tile descriptors
// {status:2, value:62} in the same kernel (tiles take their ids from an atomic counter, so every predecessor is
// already running): the MSA bytes are read from HBM exactly once.  ACGT validation and code extraction are
// SIMD-in-register (~11 instructions per 4 bytes).  Every warp whose 2 KB hold neither a gap nor a foreign byte
// assembles its 64-bit output words in registers with one shuffle; a warp that holds irregular bytes goes through
// its own shared-memory bit buffer.  Partially covered words at tile / warp borders are OR-ed atomically into the
// zero-filled output (in a tile without any irregular byte only the tile's two edge words are).
// HBM traffic = M read + m/4 written (+ m/4 zero fill): the algorithmic bytes of SURVEY.md 8(d).
#include "drlz_internal.h"
#include "scan.cuh"

namespace drlz {

struct Seg16 {
    uint32_t cnt;      // non-gap bytes
    uint32_t exc;      // exceptions among them
    uint32_t codes;    // cnt codes, left aligned in 32 bits
};

__device__ __forceinline__ uint4 load16(const uint8_t* p) { return *reinterpret_cast<const uint4*>(p); }

static const uint64_t kDescAgg = 1ull << 62, kDescPrefix = 2ull << 62, kDescMask = (1ull << 62) - 1;

__device__ __forceinline__ uint64_t ld_desc(const uint64_t* p) { return *reinterpret_cast<const volatile uint64_t*>(p); }
__device__ __forceinline__ void st_desc(uint64_t* p, uint64_t v) { *reinterpret_cast<volatile uint64_t*>(p) = v; }

// decoupled look-back by one warp: exclusive prefix of `agg` over tiles [0, t) of the row; publishes the tile
// `flags` bit 1 is raised (and the wait abandoned) if a predecessor does not publish within ~0.2 s: with ids taken
// from blockIdx the kernel relies on the hardware dispatching CTAs in index order; the host then redoes the batch
// with atomic tickets.
__device__ __forceinline__ uint64_t lookback(uint64_t* desc_row, uint32_t t, uint64_t agg, uint32_t* flags) {
    const int lane = threadIdx.x & 31;
    long long t0 = 0; bool timing = false;
    if (t == 0) { if (lane == 0) st_desc(&desc_row[0], kDescPrefix | agg); return 0; }
    if (lane == 0) st_desc(&desc_row[t], kDescAgg | agg);
    uint64_t excl = 0;
    int64_t look = (int64_t)t - 1;
    for (;;) {
        int64_t idx = look - lane;
        uint64_t v = idx >= 0 ? ld_desc(&desc_row[idx]) : kDescPrefix;
        unsigned st = (unsigned)(v >> 62);
        if (__any_sync(0xffffffffu, st == 0)) {                  // a predecessor has not published yet
            if (!timing) { t0 = clock64(); timing = true; }
            else if (clock64() - t0 > 400000000ll) { if (lane == 0) atomicOr(flags, 2u); break; }
            continue;
        }
        unsigned pm = __ballot_sync(0xffffffffu, st == 2);
        int first = pm ? __ffs(pm) - 1 : 32;
        uint64_t val = lane <= first ? (v & kDescMask) : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
        excl += val;
        if (pm) break;
        look -= 32;
    }
    if (lane == 0) st_desc(&desc_row[t], kDescPrefix | (excl + agg));
    return excl;
}

static const int kPackThreads = kPackTileBytes / 64;

// ---- classification -------------------------------------------------------------------------------------------------
// The kernel is bound by integer issue (profiles/README.md items 13 and 18), hence:
//   * ACGT validation from the masked byte itself: (b & 0xF9) == 0x41 for A, C, G and 0x50 for T, where "T" is
//     bits 2:1 == 10 -- two shifts, four logic operations and two multiply-adds per 4 bytes;
//   * a 16-byte segment with a gap is compacted word by word (clean words are appended 4 codes at a time);
//   * block scan with one barrier (every warp sums the 8 warp totals itself).
__device__ __forceinline__ uint32_t codes4(uint32_t w, uint32_t* bad) {
    const uint32_t s1 = w >> 1, s2 = w >> 2;
    const uint32_t t = s2 & ~s1 & 0x01010101u;                        // bits 2:1 == 10: 'T' among the valid bytes
    *bad = (w & 0xF9F9F9F9u) ^ (t * 15u + 0x41414141u);               // per byte: zero iff the byte is A, C, G or T
    return (s1 ^ (s2 & 0x01010101u)) & 0x03030303u;                   // code = [bit 2, bit 1 ^ bit 2]: A0 C1 G2 T3
}
// top bytes of four products -> one word, p0's first
__device__ __forceinline__ uint32_t top_bytes(uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3) {
    return __byte_perm(__byte_perm(p3, p2, 0x0073), __byte_perm(p1, p0, 0x0073), 0x5410);
}

__device__ __forceinline__ Seg16 classify16(uint4 v, int nvalid, int strip) {
    Seg16 s;
    uint32_t b0, b1, b2, b3;
    const uint32_t c0 = codes4(v.x, &b0), c1 = codes4(v.y, &b1), c2 = codes4(v.z, &b2), c3 = codes4(v.w, &b3);
    const uint32_t p0 = c0 * 0x40100401u, p1 = c1 * 0x40100401u, p2 = c2 * 0x40100401u, p3 = c3 * 0x40100401u;
    if (nvalid == 16 && (b0 | b1 | b2 | b3) == 0) {
        s.cnt = 16; s.exc = 0; s.codes = top_bytes(p0, p1, p2, p3);
        return s;
    }
    const uint32_t w[4] = {v.x, v.y, v.z, v.w}, bad[4] = {b0, b1, b2, b3}, p[4] = {p0, p1, p2, p3};
    uint32_t cnt = 0, exc = 0, codes = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (bad[q] == 0 && 4 * q + 4 <= nvalid) { codes = (codes << 8) | (p[q] >> 24); cnt += 4; continue; }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint32_t byte = (w[q] >> (8 * b)) & 255u;
            if (4 * q + b >= nvalid || (strip && byte == '-')) continue;
            uint32_t c = (byte >> 1) & 3u; c ^= c >> 1;
            if ((bad[q] >> (8 * b)) & 255u) { c = 0; ++exc; }
            codes = (codes << 2) | c;
            ++cnt;
        }
    }
    s.cnt = cnt; s.exc = exc;
    s.codes = cnt ? codes << (2 * (16 - cnt)) : 0u;
    return s;
}

static const int kPackWarps = kPackThreads / 32;
static const int kWarpBits = 132;       // uint32 words of a warp's bit buffer: 2048 bases / 16 + the pad words the flush reads

// The 64-bit output words of T bases whose codes sit left-aligned in the bit buffer sb (16 per uint32, zero behind
// the last one), first base at row position o.  Thread tid of nthr takes words tid, tid + nthr, ...; the partially
// covered first / last word is OR-ed into the zero-filled output, the others are stored.
__device__ __forceinline__ void flush_bits(const uint32_t* sb, uint64_t* X, uint64_t o, uint32_t T, int tid, int nthr) {
    const uint32_t lead = (uint32_t)(o & 31);
    const uint64_t W0 = o >> 5;
    const uint32_t nwords = (lead + T + 31) >> 5;
    for (uint32_t j = tid; j < nwords; j += nthr) {
        const int64_t lb = (int64_t)j * 32 - (int64_t)lead;      // local base index of the word's first base
        const uint32_t b0 = lb < 0 ? 0u : (uint32_t)lb;
        const uint32_t wi = b0 >> 4, sh = (b0 & 15u) * 2;
        const uint64_t hi = ((uint64_t)sb[wi] << 32) | sb[wi + 1];
        uint64_t val = sh ? (hi << sh) | ((uint64_t)sb[wi + 2] >> (32 - sh)) : hi;
        if (lb < 0) val >>= 2 * lead;
        const bool first_partial = j == 0 && lead != 0;
        const bool last_partial = j == nwords - 1 && ((lead + T) & 31u) != 0;
        if (first_partial || last_partial) atomicOr(reinterpret_cast<unsigned long long*>(&X[W0 + j]), (unsigned long long)val);
        else X[W0 + j] = val;
    }
}

__device__ __forceinline__ void pack_tile(const PackJob& job, uint32_t h, uint32_t t, uint32_t* sbits, uint32_t* sm,
                                           uint64_t* s_bcast, uint64_t* s_wlast) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t cols = job.cols[h];
    const uint64_t tile_start = (uint64_t)t * kPackTileBytes;
    uint64_t* dcnt = job.desc_cnt + (uint64_t)h * job.tiles_per_row;
    uint64_t* dexc = job.desc_exc + (uint64_t)h * job.tiles_per_row;
    const bool last_tile = t + 1 == job.tiles_per_row;

    // ---- load and classify 64 bytes per thread -------------------------------------------------------------
    const uint64_t off = tile_start + (uint64_t)tid * 64;
    Seg16 sg[4];
    uint32_t cnt = 0, exc = 0;
    const uint8_t* rowp = job.rows + job.row_off[h];
    if (tile_start + kPackTileBytes <= cols) {            // interior tile (block-uniform): no bounds logic at all
        const uint4* src = reinterpret_cast<const uint4*>(rowp + off);
        uint4 v0 = src[0], v1 = src[1], v2 = src[2], v3 = src[3];
        sg[0] = classify16(v0, 16, job.strip); sg[1] = classify16(v1, 16, job.strip);
        sg[2] = classify16(v2, 16, job.strip); sg[3] = classify16(v3, 16, job.strip);
        cnt = sg[0].cnt + sg[1].cnt + sg[2].cnt + sg[3].cnt;
        exc = sg[0].exc + sg[1].exc + sg[2].exc + sg[3].exc;
    } else {
        uint4 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint64_t o = off + 16u * q;
            v[q] = make_uint4(0, 0, 0, 0);
            if (o < cols) v[q] = load16(rowp + o);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint64_t o = off + 16u * q;
            sg[q].cnt = 0; sg[q].exc = 0; sg[q].codes = 0;
            if (o < cols) {
                int nvalid = cols - o >= 16 ? 16 : (int)(cols - o);
                sg[q] = classify16(v[q], nvalid, job.strip);
            }
            cnt += sg[q].cnt; exc += sg[q].exc;
        }
    }
    // ---- block scan with one barrier: warp scan, then every warp adds up the warp totals it needs ----------------
    const uint32_t v = cnt | (exc << 16);                 // <= 16384 bases and <= 16384 exceptions per tile: no carry
    const uint32_t inc = warp_inclusive_scan<uint32_t, OpSum>(v, OpSum(), lane);
    if (lane == 31) sm[warp] = inc;
    __syncthreads();
    uint32_t wbase = 0, total = 0, wtot = 0;
#pragma unroll
    for (int w = 0; w < kPackWarps; ++w) {
        const uint32_t tw = sm[w];
        total += tw;
        if (w < warp) wbase += tw;
        if (w == warp) wtot = tw;
    }
    const uint32_t ex = wbase + inc - v;
    const uint32_t lo = ex & 0xFFFFu, elo = ex >> 16;     // local base offset, local exception offset
    const uint32_t T = total & 0xFFFFu, TE = total >> 16;
    const bool fast = T == kPackTileBytes && TE == 0;     // no gap, no foreign byte, full tile
    const uint32_t lw = wbase & 0xFFFFu, Tw = wtot & 0xFFFFu;     // the warp's first local base and its base count
    const bool wclean = wtot == 2048u;                    // 32 lanes x 64 bases and no exception
    uint32_t* sb = sbits + warp * kWarpBits;

    // ---- tile offset inside the row: decoupled look-back by warp 0, while the warps that hold irregular bytes put
    //      their bases into their bit buffers (that needs local offsets only) --------------------------------------
    if (warp == 0) {
        uint64_t o = lookback(dcnt, t, T, job.flags);
        uint64_t eo = 0;
        if (TE || last_tile) eo = lookback(dexc, t, TE, job.flags);
        else if (lane == 0) st_desc(&dexc[t], (t == 0 ? kDescPrefix : kDescAgg) | 0ull);
        if (lane == 0) {
            s_bcast[1] = o; s_bcast[2] = eo;
            if (last_tile) {
                job.m_out[h] = o + T;
                uint64_t ne = eo + TE;
                job.exc_cnt[h] = ne;
                if (ne > job.exc_stride) atomicOr(job.flags, 1u);
            }
        }
    }
    if (!wclean && Tw) {
        for (int i = lane; i < kWarpBits; i += 32) sb[i] = 0;
        __syncwarp();
        uint32_t l = lo - lw;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (sg[q].cnt) {
                uint32_t wi = l >> 4, sh = (l & 15u) * 2;
                atomicOr(&sb[wi], sg[q].codes >> sh);
                if (sh && sg[q].cnt > 16 - (l & 15u)) atomicOr(&sb[wi + 1], sg[q].codes << (32 - sh));
                l += sg[q].cnt;
            }
        }
    }
    if (fast && lane == 31) s_wlast[warp] = ((uint64_t)sg[2].codes << 32) | sg[3].codes;
    __syncthreads();
    const uint64_t o = s_bcast[1];
    if (T == 0) return;
    uint64_t* X = job.X + job.xoff[h];

    if (job.gapless || exc) {          // optional gapless bytes / rare exceptions: plain byte loops
        uint64_t e = s_bcast[2] + elo;
        uint64_t p = o + lo;
        uint8_t* gl = job.gapless ? job.gapless + job.gapless_off[h] : nullptr;
        uint32_t* ep = job.exc_pos + (uint64_t)h * job.exc_stride;
        uint8_t* eb = job.exc_byte + (uint64_t)h * job.exc_stride;
        for (int q = 0; q < 4; ++q) {
            uint64_t ob = off + 16u * q;
            int nvalid = ob >= cols ? 0 : (cols - ob >= 16 ? 16 : (int)(cols - ob));
            for (int b = 0; b < nvalid; ++b) {
                uint32_t byte = rowp[ob + b];              // rare path: re-read (L1/L2 hit)
                if (job.strip && byte == '-') continue;
                bool valid = byte == 'A' || byte == 'C' || byte == 'G' || byte == 'T';
                if (!valid) {
                    if (e < job.exc_stride) { ep[e] = (uint32_t)p; eb[e] = (uint8_t)byte; }
                    ++e;
                }
                if (gl) gl[p] = (uint8_t)byte;
                ++p;
            }
        }
    }

    if (wclean) {
        // thread owns local 64-bit words 2*lane, 2*lane+1 of its warp; output word j = (local[j-1] << 2(32-lead)) |
        // (local[j] >> 2 lead).  In a fully clean tile the word before lane 0's comes from the previous warp through
        // shared memory and only the tile's two edge words are OR-ed; otherwise each clean warp ORs its own two.
        const uint64_t ow = o + lw;
        const uint32_t lead = (uint32_t)(ow & 31);
        const uint64_t w0 = ((uint64_t)sg[0].codes << 32) | sg[1].codes;
        const uint64_t w1 = ((uint64_t)sg[2].codes << 32) | sg[3].codes;
        uint64_t prev = __shfl_up_sync(0xffffffffu, w1, 1);
        if (lane == 0) prev = (fast && warp) ? s_wlast[warp - 1] : 0ull;
        uint64_t g0 = w0, g1 = w1;
        if (lead) {
            const unsigned sh = 2 * lead;
            g0 = (prev << (64 - sh)) | (w0 >> sh);
            g1 = (w0 << (64 - sh)) | (w1 >> sh);
        }
        uint64_t* dst = X + (ow >> 5) + 2 * (uint64_t)lane;
        const bool first_edge = lane == 0 && lead != 0 && !(fast && warp);
        const bool last_edge = lane == 31 && lead != 0 && !(fast && warp != kPackWarps - 1);
        if (first_edge) atomicOr(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)g0);
        else dst[0] = g0;
        dst[1] = g1;
        if (last_edge) atomicOr(reinterpret_cast<unsigned long long*>(dst + 2), (unsigned long long)(w1 << (64 - 2 * lead)));
        return;
    }
    if (Tw) flush_bits(sb, X, o + lw, Tw, lane, 32);
}

template <int MINB>
__global__ void __launch_bounds__(kPackThreads, MINB) k_pack(PackJob job, uint32_t n_tiles) {
    __shared__ uint32_t sbits[kPackWarps * kWarpBits];
    __shared__ uint32_t sm[kPackWarps];
    __shared__ uint64_t s_bcast[4];
    __shared__ uint64_t s_wlast[kPackWarps];
    __shared__ uint32_t s_ticket[5];
    if (threadIdx.x == 0) {
        uint32_t id = job.use_tickets ? atomicAdd(job.tile_counter, 1u) : blockIdx.x;
        s_ticket[0] = id; s_ticket[1] = id % job.H; s_ticket[2] = id / job.H;
        const uint32_t idp = id + job.prefetch;
        s_ticket[3] = idp % job.H; s_ticket[4] = (job.prefetch && idp < n_tiles) ? idp / job.H : 0xFFFFFFFFu;
    }
    __syncthreads();
    if (s_ticket[0] >= n_tiles) return;
    if (s_ticket[4] != 0xFFFFFFFFu && threadIdx.x < kPackTileBytes / 128) {
        const uint32_t hp = s_ticket[3];
        const uint64_t o = (uint64_t)s_ticket[4] * kPackTileBytes + (uint64_t)threadIdx.x * 128;
        if (o < job.cols[hp]) asm volatile("prefetch.global.L2 [%0];" ::"l"(job.rows + job.row_off[hp] + o));
    }
    pack_tile(job, s_ticket[1], s_ticket[2], sbits, sm, s_bcast, s_wlast);
}

// =====================================================================================================================
// k_pack2 -- the same pass with the common case written for the instruction issue slots it costs.
//
// ncu on k_pack (profiles/README.md, round 1): 8.1 thread-instructions per input byte, ALU pipe 2/3 busy, 344 bytes of
// spills at 32 registers -- the general classification (gaps, foreign bytes, ragged tile ends) sat inline four times in
// every thread's path.  Here a warp whose 2 KB are plain A/C/G/T -- 97 % of the warps at BASELINE's indel rates -- runs
// straight-line code: four 128-bit loads, 7 operations per 4 bytes for the exact ACGT test and the codes, one vote, its
// words shifted onto the output grid with four funnel shifts and stored; no per-lane counts, no scan.  Everything else
// (a gap or foreign byte anywhere in the warp's 2 KB, the last tile of a row) lives in two __noinline__ functions that
// re-read the thread's 64 bytes (L1 hits) and keep their registers to themselves.  32-bit positions throughout (rows
// are shorter than 2^32 - 1024).  The exception count of a row is a plain atomic sum (only tiles that hold a foreign byte
// look back over the exception descriptors, and a tile without one copies a finished predecessor's prefix when it is
// there): the last tile of a row no longer walks the whole descriptor row at the tail of the kernel.
// =====================================================================================================================

// per-lane result of the general classification, kept between the two slow-path calls
struct SlowLane { uint32_t cnt, exc; };

// 16 bytes of plain bases: codes left aligned in 32 bits.  5 ALU-pipe + 2 FMA-pipe operations per 4 bytes:
//   t    = bits 2:1 == 10 ('T' among the valid bytes), E = the byte a valid input must equal once its bits 2:1 are
//          masked away (0x41 for A, C, G; 0x50 for T), acc collects w ^ E unmasked -- the mask 0xF9 is applied once, to
//          the OR over all 16 words of the thread;
//   code = [bit 2, bit 1 ^ bit 2] = the low two bits of (w >> 1) ^ (w >> 2) for every byte whose bit 3 is clear, which
//          holds for A, C, G and T (a byte that fails the test never reaches the output through this path).
__device__ __forceinline__ uint32_t code_word(uint32_t w, uint32_t& acc) {
    const uint32_t s1 = w >> 1, s2 = w >> 2;
    const uint32_t t = s2 & ~s1 & 0x01010101u;
    acc |= w ^ (t * 15u + 0x41414141u);
    return ((s1 ^ s2) & 0x03030303u) * 0x40100401u;
}
__device__ __forceinline__ uint32_t fast16(const uint4 v, uint32_t& acc) {
    const uint32_t p0 = code_word(v.x, acc), p1 = code_word(v.y, acc), p2 = code_word(v.z, acc), p3 = code_word(v.w, acc);
    return top_bytes(p0, p1, p2, p3);
}

// the thread's 64 bytes again, bounds-checked (bytes at and past `cols` read as zero and do not count)
__device__ __forceinline__ void load64_checked(const uint8_t* rowp, uint32_t off, uint32_t cols, uint4 (&v)[4], int (&nvalid)[4]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t o = off + 16u * q;
        v[q] = make_uint4(0, 0, 0, 0);
        nvalid[q] = 0;
        if (o < cols) { v[q] = load16(rowp + o); nvalid[q] = cols - o >= 16u ? 16 : (int)(cols - o); }
    }
}

// slow path, part 1: kept bases and exceptions of this lane (general classification)
__device__ __noinline__ SlowLane pack2_slow_count(const uint8_t* rowp, uint32_t off, uint32_t cols, int strip) {
    uint4 v[4]; int nv[4];
    load64_checked(rowp, off, cols, v, nv);
    SlowLane r; r.cnt = 0; r.exc = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (!nv[q]) continue;
        const Seg16 sg = classify16(v[q], nv[q], strip);
        r.cnt += sg.cnt; r.exc += sg.exc;
    }
    return r;
}

// slow path, part 2: the warp's bases through its bit buffer into the packed output, its exceptions into the row's list.
//   l       local offset of this lane's first kept base inside the warp
//   ow      row position (gapless) of the warp's first kept base; Tw kept bases in the warp
//   e       index of this lane's first exception in the row's list
__device__ __noinline__ void pack2_slow_emit(int strip, uint64_t* X, uint32_t* ep, uint8_t* eb, uint64_t exc_stride, const uint8_t* rowp,
                                             uint32_t off, uint32_t cols, uint32_t l, uint64_t ow, uint32_t Tw, uint64_t e, uint32_t nexc,
                                             uint32_t* sb) {
    const int lane = threadIdx.x & 31;
    uint4 v[4]; int nv[4];
    load64_checked(rowp, off, cols, v, nv);
    for (int i = lane; i < kWarpBits; i += 32) sb[i] = 0;
    __syncwarp();
    uint32_t ll = l;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (!nv[q]) continue;
        const Seg16 sg = classify16(v[q], nv[q], strip);
        if (sg.cnt) {
            const uint32_t wi = ll >> 4, sh = (ll & 15u) * 2;
            atomicOr(&sb[wi], sg.codes >> sh);
            if (sh && sg.cnt > 16 - (ll & 15u)) atomicOr(&sb[wi + 1], sg.codes << (32 - sh));
            ll += sg.cnt;
        }
    }
    __syncwarp();
    if (Tw) flush_bits(sb, X, ow, Tw, lane, 32);
    if (nexc) {                       // foreign bytes of this lane: (gapless position, byte) into the sorted list of the row
        uint64_t p = ow + l;
        for (int q = 0; q < 4; ++q) {
            const uint32_t w[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
            for (int b = 0; b < nv[q]; ++b) {
                const uint32_t byte = (w[b >> 2] >> (8 * (b & 3))) & 255u;
                if (strip && byte == '-') continue;
                if (!(byte == 'A' || byte == 'C' || byte == 'G' || byte == 'T')) {
                    if (e < exc_stride) { ep[e] = (uint32_t)p; eb[e] = (uint8_t)byte; }
                    ++e;
                }
                ++p;
            }
        }
    }
}

// gapless bytes (the body of the gapless FASTA record, DistributedRLZ.scala:75-80): the warp's kept bytes are staged in
// shared memory and leave as aligned 16-byte stores (the round-1 kernel wrote them byte by byte from every lane).
//   gb   the warp's staging buffer (2048 + 16 bytes, 16-byte aligned); n bytes staged; dst = their place in the output
__device__ __forceinline__ void flush_bytes(const uint8_t* gb, uint8_t* dst, uint32_t n, int lane) {
    const uint32_t head = min(n, (uint32_t)((16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u));
    if ((uint32_t)lane < head) dst[lane] = gb[lane];
    const uint32_t nfull = (n - head) >> 4;
    const uint32_t* gw = reinterpret_cast<const uint32_t*>(gb);
    const uint32_t w0 = head >> 2, sel = 0x3210u + 0x1111u * (head & 3u);
    uint4* d16 = reinterpret_cast<uint4*>(dst + head);
    for (uint32_t c = lane; c < nfull; c += 32) {
        const uint32_t* s = gw + w0 + 4 * c;
        const uint32_t a0 = s[0], a1 = s[1], a2 = s[2], a3 = s[3], a4 = s[4];
        d16[c] = make_uint4(__byte_perm(a0, a1, sel), __byte_perm(a1, a2, sel), __byte_perm(a2, a3, sel), __byte_perm(a3, a4, sel));
    }
    const uint32_t done = head + (nfull << 4);
    if (done + (uint32_t)lane < n) dst[done + lane] = gb[done + lane];
}
__device__ __noinline__ void pack2_slow_gapless(int strip, const uint8_t* rowp, uint32_t off, uint32_t cols, uint32_t l, uint8_t* gb) {
    uint4 v[4]; int nv[4];
    load64_checked(rowp, off, cols, v, nv);
    uint32_t p = l;
    for (int q = 0; q < 4; ++q) {
        const uint32_t w[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
        for (int b = 0; b < nv[q]; ++b) {
            const uint32_t byte = (w[b >> 2] >> (8 * (b & 3))) & 255u;
            if (strip && byte == '-') continue;
            gb[p++] = (uint8_t)byte;
        }
    }
}

static const int kGaplessStage = 2048 + 32;       // bytes of a warp's staging buffer

template <int MINB, bool GAPLESS>
__global__ void __launch_bounds__(kPackThreads, MINB) k_pack2(PackJob job, uint32_t n_tiles) {
    __shared__ uint32_t sbits[kPackWarps * kWarpBits];
    __shared__ __align__(16) uint32_t s_w[kPackWarps];
    __shared__ uint64_t s_o[2];
    __shared__ uint64_t s_wlast[kPackWarps];
    __shared__ uint32_t s_ticket[5];
    extern __shared__ __align__(16) uint8_t s_gl[];          // GAPLESS: kPackWarps staging buffers
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        const uint32_t id = job.use_tickets ? atomicAdd(job.tile_counter, 1u) : blockIdx.x;
        s_ticket[0] = id; s_ticket[1] = id % job.H; s_ticket[2] = id / job.H;
        const uint32_t idp = id + job.prefetch;
        s_ticket[3] = idp % job.H; s_ticket[4] = (job.prefetch && idp < n_tiles) ? idp / job.H : 0xFFFFFFFFu;
    }
    __syncthreads();
    if (s_ticket[0] >= n_tiles) return;
    if (s_ticket[4] != 0xFFFFFFFFu && tid < kPackTileBytes / 128) {
        const uint32_t hp = s_ticket[3];
        const uint64_t o = (uint64_t)s_ticket[4] * kPackTileBytes + (uint64_t)tid * 128;
        if (o < job.cols[hp]) asm volatile("prefetch.global.L2 [%0];" ::"l"(job.rows + job.row_off[hp] + o));
    }
    const uint32_t h = s_ticket[1], t = s_ticket[2];
    const uint32_t cols = (uint32_t)job.cols[h];
    const uint8_t* __restrict__ rowp = job.rows + job.row_off[h];
    const uint32_t tile_start = t * (uint32_t)kPackTileBytes;          // < 2^32: the host bounds cols by 2^32 - 1024
    const uint32_t off = tile_start + (uint32_t)tid * 64u;
    const bool last_tile = t + 1 == job.tiles_per_row;

    // ---- load + classify: a warp of plain bases needs nothing but its codes ------------------------------------------
    uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0, acc = 0xFFFFFFFFu;
    if (tile_start + (uint32_t)kPackTileBytes <= cols && tile_start + (uint32_t)kPackTileBytes > tile_start) {
        const uint4* src = reinterpret_cast<const uint4*>(rowp + off);
        const uint4 raw0 = src[0], raw1 = src[1], raw2 = src[2], raw3 = src[3];
        if (GAPLESS) {                 // the bytes themselves are the gapless output of a plain warp: staged right away
            uint4* g4 = reinterpret_cast<uint4*>(s_gl + warp * kGaplessStage + lane * 64);
            g4[0] = raw0; g4[1] = raw1; g4[2] = raw2; g4[3] = raw3;
        }
        acc = 0;
        c0 = fast16(raw0, acc); c1 = fast16(raw1, acc); c2 = fast16(raw2, acc); c3 = fast16(raw3, acc);
    }
    const bool wclean = __all_sync(0xffffffffu, (acc & 0xF9F9F9F9u) == 0);
    uint32_t lane_cnt = 64, lane_exc = 0, lane_lo = (uint32_t)lane * 64u, lane_elo = 0, wtot = 2048u;
    if (!wclean) {
        const SlowLane sl = pack2_slow_count(rowp, off, cols, job.strip);
        lane_cnt = sl.cnt; lane_exc = sl.exc;
        const uint32_t v = sl.cnt | (sl.exc << 16);
        const uint32_t inc = warp_inclusive_scan<uint32_t, OpSum>(v, OpSum(), lane);
        wtot = __shfl_sync(0xffffffffu, inc, 31);
        lane_lo = (inc - v) & 0xFFFFu; lane_elo = (inc - v) >> 16;
    }
    if (lane == 0) s_w[warp] = wtot;
    __syncthreads();
    const uint4 wa = *reinterpret_cast<const uint4*>(&s_w[0]), wb = *reinterpret_cast<const uint4*>(&s_w[4]);
    const uint32_t total = wa.x + wa.y + wa.z + wa.w + wb.x + wb.y + wb.z + wb.w;
    const bool fast = total == (uint32_t)kPackTileBytes;              // every warp plain, full tile
    uint32_t wbase = (uint32_t)warp * 2048u;
    if (!fast) {
        const uint32_t tw[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
        wbase = 0;
#pragma unroll
        for (int w = 0; w < kPackWarps; ++w) if (w < warp) wbase += tw[w];
    }
    const uint32_t T = total & 0xFFFFu, TE = total >> 16;
    const uint32_t lw = wbase & 0xFFFFu, lew = wbase >> 16, Tw = wtot & 0xFFFFu;

    // ---- tile offset inside the row: decoupled look-back by warp 0 ---------------------------------------------------
    if (warp == 0) {
        uint64_t* dcnt = job.desc_cnt + (uint64_t)h * job.tiles_per_row;
        uint64_t* dexc = job.desc_exc + (uint64_t)h * job.tiles_per_row;
        const uint64_t o = lookback(dcnt, t, T, job.flags);
        uint64_t eo = 0;
        if (TE) {
            eo = lookback(dexc, t, TE, job.flags);
            if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&job.exc_cnt[h]), (unsigned long long)TE);
        } else if (lane == 0) {
            // no foreign byte here: pass a finished predecessor's prefix on if it is there, else publish "0 of mine"
            const uint64_t pv = t ? ld_desc(&dexc[t - 1]) : kDescPrefix;
            st_desc(&dexc[t], (pv >> 62) == 2 ? pv : kDescAgg);
        }
        if (lane == 0) {
            s_o[0] = o; s_o[1] = eo;
            if (last_tile) job.m_out[h] = o + T;
        }
    }
    if (fast && lane == 31) s_wlast[warp] = ((uint64_t)c2 << 32) | c3;
    __syncthreads();
    if (T == 0) return;
    const uint64_t ow = s_o[0] + lw;                                   // row position of the warp's first kept base

    if (GAPLESS) {
        uint8_t* gb = s_gl + warp * kGaplessStage;
        if (!wclean && lane_cnt) pack2_slow_gapless(job.strip, rowp, off, cols, lane_lo, gb);
        __syncwarp();
        if (Tw) flush_bytes(gb, job.gapless + job.gapless_off[h] + ow, Tw, lane);
    }

    if (wclean) {
        // The lane owns the warp's local 64-bit words 2*lane and 2*lane+1 = the 32-bit pieces c0..c3; the output grid is
        // shifted by lead = ow % 32 bases, so output word j = the 64 bits that start 64 - 2*lead bits into local words
        // (j-1, j).  lead is the same for the whole warp: the half of the funnel is picked by a uniform branch, and every
        // output half is one funnel shift of two neighbouring pieces.  In a fully plain tile the pieces in front of
        // lane 0 come from the previous warp through shared memory and only the tile's two edge words are OR-ed into
        // the zero-filled output; otherwise each plain warp ORs its own two.
        const uint32_t lead = (uint32_t)ow & 31u;
        uint32_t pm1 = __shfl_up_sync(0xffffffffu, c3, 1), pm2 = __shfl_up_sync(0xffffffffu, c2, 1);   // pieces in front of c0
        if (lane == 0) {
            const uint64_t pw = (fast && warp) ? s_wlast[warp - 1] : 0ull;
            pm1 = (uint32_t)pw; pm2 = (uint32_t)(pw >> 32);
        }
        uint32_t g0h = c0, g0l = c1, g1h = c2, g1l = c3, tail_h = 0, tail_l = 0;
        if (lead) {
            const unsigned sr = (2u * lead) & 31u;          // right shift inside a 32-bit piece
            if (lead < 16) {                                 // less than one piece: output piece k = (piece k-1 : piece k) >> sr
                g0h = __funnelshift_r(c0, pm1, sr); g0l = __funnelshift_r(c1, c0, sr);
                g1h = __funnelshift_r(c2, c1, sr);  g1l = __funnelshift_r(c3, c2, sr);
                tail_h = c3 << (32 - sr);
            } else {                                         // one piece and more: output piece k = (piece k-2 : piece k-1) >> sr
                g0h = __funnelshift_r(pm1, pm2, sr); g0l = __funnelshift_r(c0, pm1, sr);
                g1h = __funnelshift_r(c1, c0, sr);   g1l = __funnelshift_r(c2, c1, sr);
                tail_h = __funnelshift_r(c3, c2, sr); tail_l = sr ? c3 << (32 - sr) : 0u;
            }
        }
        const uint64_t g0 = ((uint64_t)g0h << 32) | g0l, g1 = ((uint64_t)g1h << 32) | g1l;
        uint64_t* dst = job.X + job.xoff[h] + (ow >> 5) + 2 * (uint32_t)lane;
        const bool first_edge = lane == 0 && lead != 0 && !(fast && warp);
        const bool last_edge = lane == 31 && lead != 0 && !(fast && warp != kPackWarps - 1);
        if (first_edge) atomicOr(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)g0);
        else dst[0] = g0;
        dst[1] = g1;
        if (last_edge) atomicOr(reinterpret_cast<unsigned long long*>(dst + 2), ((unsigned long long)tail_h << 32) | tail_l);
        return;
    }
    pack2_slow_emit(job.strip, job.X + job.xoff[h], job.exc_pos + (uint64_t)h * job.exc_stride, job.exc_byte + (uint64_t)h * job.exc_stride,
                    job.exc_stride, rowp, off, cols, lane_lo, ow, Tw, s_o[1] + lew + lane_elo, lane_exc, sbits + warp * kWarpBits);
}

// =====================================================================================================================
// k_pack3 -- streaming form: a CTA takes 8 consecutive 16 KB tiles of one row.
//
// What bounds the tile-per-CTA kernels once their instructions are cut (k_pack2: 40 % fewer instructions, 3 % less
// time) is the life cycle of a CTA: ticket -> loads -> classification -> look-back -> stores, ~12 000 cycles per 16 KB
// with every step waiting for the one before and 8 CTAs per SM to hide it.  Here
//   phase 1  every warp streams its 8 pieces (2 KB each: warp w of tile k) on its own -- no barrier, the loads of piece
//            k+1 in flight while piece k is classified -- and leaves the packed codes of each piece in shared memory
//            (plain piece: one 128-bit store per lane; a piece with a gap or foreign byte: compacted by the slow path);
//   phase 2  one barrier, one scan over the 64 piece counts, ONE descriptor and one look-back per CTA (per 128 KB);
//   phase 3  every warp shifts its pieces onto the output grid and stores them.
// The fixed costs of a CTA are paid once per 128 KB instead of once per 16 KB, and the loads never wait for a look-back.
// =====================================================================================================================
static const int kP3Pieces = 8;                                   // tiles per CTA = pieces per warp
static const uint32_t kP3Bytes = kP3Pieces * kPackTileBytes;      // row bytes per CTA
static const int kP3Stash = kWarpBits;                            // words per stashed piece (128 + what flush_bits reads past the end)

// slow path of phase 1: the piece compacted into its stash through the general classification; returns the warp's
// kept bases | exceptions << 16 (warp-uniform).  Called by the whole warp.
__device__ __noinline__ uint32_t pack3_slow_piece(const uint8_t* rowp, uint32_t off, uint32_t cols, int strip, uint32_t* sb) {
    const int lane = threadIdx.x & 31;
    uint4 v[4]; int nv[4];
    load64_checked(rowp, off, cols, v, nv);
    Seg16 sg[4];
    uint32_t cnt = 0, exc = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        sg[q].cnt = 0; sg[q].exc = 0; sg[q].codes = 0;
        if (nv[q]) sg[q] = classify16(v[q], nv[q], strip);
        cnt += sg[q].cnt; exc += sg[q].exc;
    }
    const uint32_t val = cnt | (exc << 16);
    const uint32_t inc = warp_inclusive_scan<uint32_t, OpSum>(val, OpSum(), lane);
    const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
    for (int i = lane; i < kP3Stash; i += 32) sb[i] = 0;
    __syncwarp();
    uint32_t l = (inc - val) & 0xFFFFu;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (sg[q].cnt) {
            const uint32_t wi = l >> 4, sh = (l & 15u) * 2;
            atomicOr(&sb[wi], sg[q].codes >> sh);
            if (sh && sg[q].cnt > 16 - (l & 15u)) atomicOr(&sb[wi + 1], sg[q].codes << (32 - sh));
            l += sg[q].cnt;
        }
    }
    __syncwarp();
    return tot;
}

// slow path of phase 3: the foreign bytes of a piece into the row's exception list (ow = row position of the piece's
// first kept base, e0 = index of its first exception).  Called by the whole warp.
__device__ __noinline__ void pack3_slow_exc(const uint8_t* rowp, uint32_t off, uint32_t cols, int strip, uint32_t* ep, uint8_t* eb,
                                            uint64_t exc_stride, uint64_t ow, uint64_t e0) {
    const int lane = threadIdx.x & 31;
    uint4 v[4]; int nv[4];
    load64_checked(rowp, off, cols, v, nv);
    uint32_t cnt = 0, exc = 0;
    for (int q = 0; q < 4; ++q) {
        const uint32_t w[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
        for (int b = 0; b < nv[q]; ++b) {
            const uint32_t byte = (w[b >> 2] >> (8 * (b & 3))) & 255u;
            if (strip && byte == '-') continue;
            ++cnt;
            if (!(byte == 'A' || byte == 'C' || byte == 'G' || byte == 'T')) ++exc;
        }
    }
    const uint32_t val = cnt | (exc << 16);
    const uint32_t inc = warp_inclusive_scan<uint32_t, OpSum>(val, OpSum(), lane);
    uint64_t p = ow + ((inc - val) & 0xFFFFu), e = e0 + ((inc - val) >> 16);
    if (!exc) return;
    for (int q = 0; q < 4; ++q) {
        const uint32_t w[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
        for (int b = 0; b < nv[q]; ++b) {
            const uint32_t byte = (w[b >> 2] >> (8 * (b & 3))) & 255u;
            if (strip && byte == '-') continue;
            if (!(byte == 'A' || byte == 'C' || byte == 'G' || byte == 'T')) {
                if (e < exc_stride) { ep[e] = (uint32_t)p; eb[e] = (uint8_t)byte; }
                ++e;
            }
            ++p;
        }
    }
}

template <int MINB>
__global__ void __launch_bounds__(kPackThreads, MINB) k_pack3(PackJob job, uint32_t n_ctas, uint32_t ctas_per_row) {
    extern __shared__ __align__(16) uint32_t p3_stash[];               // [kPackWarps][kP3Pieces][kP3Stash]
    __shared__ uint32_t s_cnt[kP3Pieces * kPackWarps];                 // piece (k, w) at k * 8 + w: kept bases | exceptions << 16
    __shared__ uint32_t s_pre[kP3Pieces * kPackWarps], s_pre_e[kP3Pieces * kPackWarps];   // exclusive prefixes inside the CTA
    __shared__ uint64_t s_o[2];
    __shared__ uint32_t s_ticket[3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        const uint32_t id = job.use_tickets ? atomicAdd(job.tile_counter, 1u) : blockIdx.x;
        s_ticket[0] = id; s_ticket[1] = id % job.H; s_ticket[2] = id / job.H;
    }
    __syncthreads();
    if (s_ticket[0] >= n_ctas) return;
    const uint32_t h = s_ticket[1], c = s_ticket[2];
    const uint32_t cols = job.uniform_cols ? job.uniform_cols : (uint32_t)job.cols[h];
    const uint8_t* __restrict__ rowp = job.rows + (job.uniform_cols ? job.uniform_off0 + (uint64_t)h * job.uniform_stride : job.row_off[h]);
    const uint32_t base = c * kP3Bytes + (uint32_t)warp * 2048u;       // first byte of this warp's piece 0 (< 2^32, as cols is)
    uint32_t* my = p3_stash + warp * (kP3Pieces * kP3Stash);

    // ---- phase 1: stream the pieces --------------------------------------------------------------------------------
    // piece k is "full" when all of its 2 KB lie inside the row; (only) full pieces take the plain path
    // the registers hold piece k+1 while piece k is classified; lanes 0..15 ask L2 for piece k+1+pf (one 128-byte line
    // each) so that the register loads find their bytes on the chip
    const uint32_t pf = job.prefetch;
    auto l2_prefetch = [&](uint32_t pb) {
        if (lane < 16 && pb < cols && cols - pb > (uint32_t)lane * 128u) asm volatile("prefetch.global.L2 [%0];" ::"l"(rowp + pb + (uint32_t)lane * 128u));
    };
    uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0, n2 = n0, n3 = n0;
    {
        const uint32_t pb = base;
        if (pb < cols && cols - pb >= 2048u) {
            const uint4* src = reinterpret_cast<const uint4*>(rowp + pb + (uint32_t)lane * 64u);
            n0 = src[0]; n1 = src[1]; n2 = src[2]; n3 = src[3];
        }
        for (uint32_t j = 1; j <= pf && j < (uint32_t)kP3Pieces; ++j) l2_prefetch(base + j * (uint32_t)kPackTileBytes);
    }
#pragma unroll 1
    for (int k = 0; k < kP3Pieces; ++k) {
        const uint32_t pb = base + (uint32_t)k * (uint32_t)kPackTileBytes;
        const bool full = pb < cols && cols - pb >= 2048u;
        const uint4 r0 = n0, r1 = n1, r2 = n2, r3 = n3;
        {
            const uint32_t nb = pb + (uint32_t)kPackTileBytes;
            if (k + 1 < kP3Pieces && nb < cols && cols - nb >= 2048u) {
                const uint4* src = reinterpret_cast<const uint4*>(rowp + nb + (uint32_t)lane * 64u);
                n0 = src[0]; n1 = src[1]; n2 = src[2]; n3 = src[3];
            }
            if (pf && k + 1 + (int)pf < kP3Pieces) l2_prefetch(nb + pf * (uint32_t)kPackTileBytes);
        }
        uint32_t cnt;
        bool plain = false;
        if (full) {
            uint32_t acc = 0;
            const uint32_t c0 = fast16(r0, acc), c1 = fast16(r1, acc), c2 = fast16(r2, acc), c3 = fast16(r3, acc);
            plain = __all_sync(0xffffffffu, (acc & 0xF9F9F9F9u) == 0);
            if (plain) *reinterpret_cast<uint4*>(my + k * kP3Stash + 4 * lane) = make_uint4(c0, c1, c2, c3);
        }
        if (plain) cnt = 2048u;
        else if (pb < cols) cnt = pack3_slow_piece(rowp, pb + (uint32_t)lane * 64u, cols, job.strip, my + k * kP3Stash);
        else cnt = 0;
        if (lane == 0) s_cnt[k * kPackWarps + warp] = cnt;
    }
    __syncthreads();

    // ---- phase 2: offsets of the 64 pieces inside the CTA, one look-back for the CTA -------------------------------------
    if (warp == 0) {
        static_assert(kP3Pieces * kPackWarps == 64, "two pieces per lane in the scan below");
        const uint32_t v0 = s_cnt[2 * lane], v1 = s_cnt[2 * lane + 1];
        const uint32_t b0 = v0 & 0xFFFFu, b1 = v1 & 0xFFFFu, e0 = v0 >> 16, e1 = v1 >> 16;
        const uint32_t sb = b0 + b1, se = e0 + e1;
        const uint32_t ib = warp_inclusive_scan<uint32_t, OpSum>(sb, OpSum(), lane), ie = warp_inclusive_scan<uint32_t, OpSum>(se, OpSum(), lane);
        s_pre[2 * lane] = ib - sb; s_pre[2 * lane + 1] = ib - sb + b0;
        s_pre_e[2 * lane] = ie - se; s_pre_e[2 * lane + 1] = ie - se + e0;
        const uint32_t T = __shfl_sync(0xffffffffu, ib, 31), TE = __shfl_sync(0xffffffffu, ie, 31);
        uint64_t* dcnt = job.desc_cnt + (uint64_t)h * ctas_per_row;
        uint64_t* dexc = job.desc_exc + (uint64_t)h * ctas_per_row;
        const uint64_t o = lookback(dcnt, c, T, job.flags);
        uint64_t eo = 0;
        if (TE) {
            eo = lookback(dexc, c, TE, job.flags);
            if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&job.exc_cnt[h]), (unsigned long long)TE);
        } else if (lane == 0) {
            const uint64_t pv = c ? ld_desc(&dexc[c - 1]) : kDescPrefix;
            st_desc(&dexc[c], (pv >> 62) == 2 ? pv : kDescAgg);
        }
        if (lane == 0) {
            s_o[0] = o; s_o[1] = eo;
            if (c + 1 == ctas_per_row) job.m_out[h] = o + T;
        }
    }
    __syncthreads();

    // ---- phase 3: the pieces onto the output grid --------------------------------------------------------------------
    uint64_t* X = job.X + job.xoff[h];
    const uint64_t o_cta = s_o[0];
#pragma unroll 1
    for (int k = 0; k < kP3Pieces; ++k) {
        const uint32_t v = s_cnt[k * kPackWarps + warp];
        const uint32_t Tw = v & 0xFFFFu;
        if (!Tw) continue;
        const uint64_t ow = o_cta + s_pre[k * kPackWarps + warp];
        const uint32_t* sb = my + k * kP3Stash;
        if (v == 2048u) {
            // a plain piece: the lane owns the 32-bit pieces c0..c3 (its 64 bases); see k_pack2 for the funnel
            const uint4 cc = *reinterpret_cast<const uint4*>(sb + 4 * lane);
            const uint32_t lead = (uint32_t)ow & 31u;
            uint32_t pm1 = __shfl_up_sync(0xffffffffu, cc.w, 1), pm2 = __shfl_up_sync(0xffffffffu, cc.z, 1);
            if (lane == 0) { pm1 = 0; pm2 = 0; }
            uint32_t g0h = cc.x, g0l = cc.y, g1h = cc.z, g1l = cc.w, tail_h = 0, tail_l = 0;
            if (lead) {
                const unsigned sr = (2u * lead) & 31u;
                if (lead < 16) {
                    g0h = __funnelshift_r(cc.x, pm1, sr); g0l = __funnelshift_r(cc.y, cc.x, sr);
                    g1h = __funnelshift_r(cc.z, cc.y, sr); g1l = __funnelshift_r(cc.w, cc.z, sr);
                    tail_h = cc.w << (32 - sr);
                } else {
                    g0h = __funnelshift_r(pm1, pm2, sr); g0l = __funnelshift_r(cc.x, pm1, sr);
                    g1h = __funnelshift_r(cc.y, cc.x, sr); g1l = __funnelshift_r(cc.z, cc.y, sr);
                    tail_h = __funnelshift_r(cc.w, cc.z, sr); tail_l = sr ? cc.w << (32 - sr) : 0u;
                }
            }
            const uint64_t g0 = ((uint64_t)g0h << 32) | g0l, g1 = ((uint64_t)g1h << 32) | g1l;
            uint64_t* dst = X + (ow >> 5) + 2 * (uint32_t)lane;
            if (lane == 0 && lead) atomicOr(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)g0);
            else dst[0] = g0;
            dst[1] = g1;
            if (lane == 31 && lead) atomicOr(reinterpret_cast<unsigned long long*>(dst + 2), ((unsigned long long)tail_h << 32) | tail_l);
        } else {
            flush_bits(sb, X, ow, Tw, lane, 32);
            if (v >> 16)
                pack3_slow_exc(rowp, base + (uint32_t)k * (uint32_t)kPackTileBytes + (uint32_t)lane * 64u, cols, job.strip,
                               job.exc_pos + (uint64_t)h * job.exc_stride, job.exc_byte + (uint64_t)h * job.exc_stride, job.exc_stride,
                               ow, s_o[1] + s_pre_e[k * kPackWarps + warp]);
        }
    }
}

// exceptions per row are summed by atomics in k_pack2: the overflow test moves here (one thread per row, after the kernel)
__global__ void k_pack2_finish(PackJob job) {
    const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h < job.H && job.exc_cnt[h] > job.exc_stride) atomicOr(job.flags, 1u);
}

// ---- generic alphabets (references with N, soft-masking, IUPAC codes ...): 4 or 8 bits per symbol ----------------
// Two passes (count, scan, write) with a byte -> code table; slower than k_pack (the MSA bytes are read twice) but
// alphabet-agnostic.  Bytes outside the reference alphabet become exceptions exactly as in the 2-bit path.
__global__ void __launch_bounds__(256) k_packg_count(PackJob job) {
    __shared__ uint32_t sm[33];
    __shared__ uint8_t s_b2c[256];                           // the byte -> code table, once per CTA
    s_b2c[threadIdx.x] = job.b2c[threadIdx.x];
    __syncthreads();
    uint32_t h = blockIdx.y, t = blockIdx.x;
    uint64_t cols = job.cols[h];
    uint64_t off = (uint64_t)t * kPackGTileBytes + (uint64_t)threadIdx.x * 16;
    const uint8_t* rowp = job.rows + job.row_off[h];
    uint32_t cnt = 0, exc = 0;
    if (off < cols) {
        const uint4 q = load16(rowp + off);                  // one 128-bit load per thread (rows are readable up to the next multiple of 16)
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        const int nv = cols - off >= 16 ? 16 : (int)(cols - off);
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            const uint32_t byte = (w[b >> 2] >> (8 * (b & 3))) & 255u;
            if (b >= nv || (job.strip && byte == '-')) continue;
            ++cnt;
            if (s_b2c[byte] == 0xFF) ++exc;
        }
    }
    uint32_t total;
    block_exclusive_scan<uint32_t, OpSum>(cnt | (exc << 16), OpSum(), 0u, &total, sm);
    if (threadIdx.x == 0) {
        job.gtile_cnt[(uint64_t)h * job.gtiles_per_row + t] = total & 0xFFFFu;
        job.gtile_exc[(uint64_t)h * job.gtiles_per_row + t] = total >> 16;
    }
}

__global__ void k_packg_rows(PackJob job, const uint64_t* total_cnt, const uint64_t* total_exc) {
    uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= job.H) return;
    uint64_t T = job.gtiles_per_row, nt = (uint64_t)job.H * T;
    uint64_t a = job.gtile_off[h * T], b = (uint64_t)(h + 1) * T < nt ? job.gtile_off[(uint64_t)(h + 1) * T] : *total_cnt;
    uint64_t ea = job.gtile_excoff[h * T], eb = (uint64_t)(h + 1) * T < nt ? job.gtile_excoff[(uint64_t)(h + 1) * T] : *total_exc;
    job.m_out[h] = b - a;
    job.exc_cnt[h] = eb - ea;
    if (eb - ea > job.exc_stride) atomicOr(job.flags, 1u);
}

__global__ void __launch_bounds__(256) k_packg_write(PackJob job) {
    __shared__ uint8_t scode[kPackGTileBytes];
    __shared__ uint32_t sm[33];
    __shared__ uint8_t s_b2c[256];
    uint32_t h = blockIdx.y, t = blockIdx.x;
    uint64_t cols = job.cols[h];
    uint64_t tile_start = (uint64_t)t * kPackGTileBytes;
    if (tile_start >= cols) return;
    s_b2c[threadIdx.x] = job.b2c[threadIdx.x];
    __syncthreads();
    const uint8_t* rowp = job.rows + job.row_off[h];
    uint64_t off = tile_start + (uint64_t)threadIdx.x * 16;
    uint8_t code[16], raw[16];
    uint32_t cnt = 0, exc = 0;
    if (off < cols) {
        const uint4 q = load16(rowp + off);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        const int nv = cols - off >= 16 ? 16 : (int)(cols - off);
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            const uint32_t byte = (w[b >> 2] >> (8 * (b & 3))) & 255u;
            if (b >= nv || (job.strip && byte == '-')) continue;
            const uint32_t c = s_b2c[byte];
            if (c == 0xFF) { ++exc; }
            code[cnt] = (uint8_t)c; raw[cnt] = (uint8_t)byte;
            ++cnt;
        }
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan<uint32_t, OpSum>(cnt | (exc << 16), OpSum(), 0u, &total, sm);
    const uint32_t lo = ex & 0xFFFFu, elo = ex >> 16, T = total & 0xFFFFu;
    uint64_t gidx = (uint64_t)h * job.gtiles_per_row + t, g0 = (uint64_t)h * job.gtiles_per_row;
    uint64_t o = job.gtile_off[gidx] - job.gtile_off[g0];            // first output symbol of the tile in the row
    uint64_t e = job.gtile_excoff[gidx] - job.gtile_excoff[g0] + elo;
    uint8_t* gl = job.gapless ? job.gapless + job.gapless_off[h] : nullptr;
    uint32_t* ep = job.exc_pos + (uint64_t)h * job.exc_stride;
    uint8_t* eb = job.exc_byte + (uint64_t)h * job.exc_stride;
    for (uint32_t q = 0; q < cnt; ++q) {
        uint32_t c = code[q];
        if (c == 0xFF) {
            if (e < job.exc_stride) { ep[e] = (uint32_t)(o + lo + q); eb[e] = raw[q]; }
            ++e; c = 0;
        }
        scode[lo + q] = (uint8_t)c;
        if (gl) gl[o + lo + q] = raw[q];
    }
    __syncthreads();
    if (T == 0) return;
    const uint32_t bl = job.bl, sl = 6 - bl, spw = 1u << sl, B = 1u << bl;
    uint64_t* X = job.X + job.xoff[h];
    const uint32_t lead = (uint32_t)(o & (spw - 1));
    const uint64_t W0 = o >> sl;
    uint32_t nwords = (lead + T + spw - 1) >> sl;
    for (uint32_t j = threadIdx.x; j < nwords; j += 256) {
        uint64_t val = 0;
        for (uint32_t q = 0; q < spw; ++q) {
            int64_t li = (int64_t)j * spw + q - (int64_t)lead;
            if (li >= 0 && li < (int64_t)T) val |= (uint64_t)scode[li] << (64 - B - B * q);
        }
        bool first_partial = j == 0 && lead != 0;
        bool last_partial = j == nwords - 1 && ((lead + T) & (spw - 1)) != 0;
        if (first_partial || last_partial) atomicOr(reinterpret_cast<unsigned long long*>(&X[W0 + j]), (unsigned long long)val);
        else X[W0 + j] = val;
    }
}

// =====================================================================================================================
// k_packg1 -- generic alphabets in one pass.
//
// The two-pass kernels above read the rows twice in 4 KB tiles (586 k CTAs of 256 threads per 2.4 GB, twice, with two
// device scans in between) and keep their codes in dynamically indexed local arrays: 0.37 TB/s.  A reference assembly
// with N in it -- every real chromosome -- takes this path, so it gets the structure of k_pack2: 16 KB tiles handed out by
// a ticket (row = id % H), 64 bytes per thread, the tile's place in the row by decoupled look-back over the same
// descriptors, exception counts by atomics.  Codes come from the 256-entry table in shared memory; every 64-bit word a
// thread completes on its own (the rule when the tile has no gap column in front of it) is a plain shared-memory store,
// the others are OR-ed in at their symbol offset; the tile's words then leave shifted onto the output's word grid.
// =====================================================================================================================
__device__ __forceinline__ uint32_t comp4(const uint4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

// the block's staged bytes (16-byte aligned, readable 16 bytes past n) to an arbitrarily aligned destination as 16-byte stores
__device__ __forceinline__ void flush_bytes_block(const uint8_t* gb, uint8_t* dst, uint32_t n, int tid, int nthr) {
    const uint32_t head = min(n, (uint32_t)((16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u));
    if ((uint32_t)tid < head) dst[tid] = gb[tid];
    const uint32_t nfull = (n - head) >> 4;
    const uint32_t* gw = reinterpret_cast<const uint32_t*>(gb);
    const uint32_t w0 = head >> 2, sel = 0x3210u + 0x1111u * (head & 3u);
    uint4* d16 = reinterpret_cast<uint4*>(dst + head);
    for (uint32_t c = tid; c < nfull; c += nthr) {
        const uint32_t* q = gw + w0 + 4 * c;
        const uint32_t a0 = q[0], a1 = q[1], a2 = q[2], a3 = q[3], a4 = q[4];
        d16[c] = make_uint4(__byte_perm(a0, a1, sel), __byte_perm(a1, a2, sel), __byte_perm(a2, a3, sel), __byte_perm(a3, a4, sel));
    }
    const uint32_t done = head + (nfull << 4);
    if (done + (uint32_t)tid < n) dst[done + tid] = gb[done + tid];
}

template <int BL, int MINB>        // log2 of the bits per symbol: 2 or 3; min CTAs per SM
__global__ void __launch_bounds__(kPackThreads, MINB) k_packg1(PackJob job, uint32_t n_tiles) {
    constexpr int B = 1 << BL, SPW = 64 / B, NG = 64 / SPW;          // bits per symbol, symbols per word, words per 64 input bytes
    constexpr int NW = kPackTileBytes / SPW;                          // words of a full tile
    __shared__ uint64_t sw[NW + 2];                                   // [0]: the empty word in front; the tile's words from [1]
    __shared__ uint8_t s_b2c[256];
    __shared__ uint32_t sm[33];
    __shared__ uint64_t s_o[2];
    __shared__ uint32_t s_ticket[3];
    extern __shared__ __align__(16) uint8_t s_gl[];                   // gapless output asked for: kPackTileBytes + 32 bytes
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        const uint32_t id = atomicAdd(job.tile_counter, 1u);
        s_ticket[0] = id; s_ticket[1] = id % job.H; s_ticket[2] = id / job.H;
    }
    s_b2c[tid] = job.b2c[tid];
    if (tid == 0) { sw[0] = 0; sw[NW + 1] = 0; }
    __syncthreads();
    if (s_ticket[0] >= n_tiles) return;
    const uint32_t h = s_ticket[1], t = s_ticket[2];
    const uint32_t cols = (uint32_t)job.cols[h];
    const uint8_t* __restrict__ rowp = job.rows + job.row_off[h];
    const uint32_t off = t * (uint32_t)kPackTileBytes + (uint32_t)tid * 64u;     // < 2^32: the host bounds cols by 2^32 - 1024
    const bool last_tile = t + 1 == job.tiles_per_row;
    const bool gl_on = job.gapless != nullptr;

    // ---- own 64 bytes: one word per SPW input bytes, kept symbols left aligned ----------------------------------------
    uint4 v[4]; int nv[4];
    load64_checked(rowp, off, cols, v, nv);
    uint64_t gw[NG]; uint32_t gc[NG];
    uint32_t cnt = 0, exc = 0;
    // the rule: 64 bytes of the alphabet, no gap column -- every symbol's place in its word is known at compile time
    bool plain = nv[0] == 16 && nv[1] == 16 && nv[2] == 16 && nv[3] == 16;
    if (plain && job.strip) {
        uint32_t any = 0;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const uint32_t x = comp4(v[q >> 2], q & 3) ^ 0x2D2D2D2Du;             // a zero byte = a '-'
            any |= (x - 0x01010101u) & ~x;
        }
        plain = (any & 0x80808080u) == 0;
    }
    if (plain) {
        uint32_t foreign = 0;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            uint32_t hi = 0, lw = 0;
#pragma unroll
            for (int q = 0; q < SPW; ++q) {
                const int b = g * SPW + q;
                const uint32_t code = s_b2c[(comp4(v[b >> 4], (b >> 2) & 3) >> (8 * (b & 3))) & 255u];
                foreign |= (code + 1u) >> 8;
                if (q < SPW / 2) hi |= code << (32 - B - B * q); else lw |= code << (32 - B - B * (q - SPW / 2));
            }
            gw[g] = ((uint64_t)hi << 32) | lw; gc[g] = SPW;
        }
        plain = foreign == 0;
        if (plain) cnt = 64;
    }
    if (!plain)
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        uint64_t w = 0; uint32_t c = 0;
#pragma unroll
        for (int q = 0; q < SPW; ++q) {
            const int b = g * SPW + q;
            const uint32_t byte = (comp4(v[b >> 4], (b >> 2) & 3) >> (8 * (b & 3))) & 255u;
            if ((b & 15) < nv[b >> 4] && !(job.strip && byte == '-')) {
                uint32_t code = s_b2c[byte];
                if (code == 0xFFu) { ++exc; code = 0; }
                w |= (uint64_t)code << (64 - B - B * (int)c);
                ++c;
            }
        }
        gw[g] = w; gc[g] = c; cnt += c;
    }
    // a tile of nothing but plain threads -- the rule -- needs no scan and no cleared word buffer: every thread's place is
    // known and every word is written whole (3 barriers per tile instead of 6)
    const bool tile_plain = __syncthreads_and(plain) != 0;
    uint32_t lo = (uint32_t)tid * 64u, elo = 0, T = (uint32_t)kPackTileBytes, TE = 0;
    if (!tile_plain) {
        for (int i = tid; i < NW + 2; i += kPackThreads) sw[i] = 0;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan<uint32_t, OpSum>(cnt | (exc << 16), OpSum(), 0u, &total, sm);     // (its barriers order the clearing too)
        lo = ex & 0xFFFFu; elo = ex >> 16; T = total & 0xFFFFu; TE = total >> 16;
    }

    // ---- the tile's place in the row: look-back by warp 0, while the others fill the word buffer ----------------------
    if (warp == 0) {
        uint64_t* dcnt = job.desc_cnt + (uint64_t)h * job.tiles_per_row;
        uint64_t* dexc = job.desc_exc + (uint64_t)h * job.tiles_per_row;
        const uint64_t o = lookback(dcnt, t, T, job.flags);
        uint64_t eo = 0;
        if (TE) {
            eo = lookback(dexc, t, TE, job.flags);
            if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&job.exc_cnt[h]), (unsigned long long)TE);
        } else if (lane == 0) {
            const uint64_t pv = t ? ld_desc(&dexc[t - 1]) : kDescPrefix;
            st_desc(&dexc[t], (pv >> 62) == 2 ? pv : kDescAgg);
        }
        if (lane == 0) {
            s_o[0] = o; s_o[1] = eo;
            if (last_tile) job.m_out[h] = o + T;
        }
    }
    {
        uint32_t pos = lo;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const uint32_t c = gc[g];
            if (!c) continue;
            const uint32_t wi = 1u + pos / SPW, r = pos % SPW;
            if (r == 0 && c == (uint32_t)SPW) sw[wi] = gw[g];                     // the whole word is this thread's
            else {
                atomicOr(reinterpret_cast<unsigned long long*>(&sw[wi]), (unsigned long long)(gw[g] >> (B * r)));
                if (r + c > (uint32_t)SPW) atomicOr(reinterpret_cast<unsigned long long*>(&sw[wi + 1]), (unsigned long long)(gw[g] << (64 - B * r)));
            }
            pos += c;
        }
    }
    if (gl_on) {                                                      // kept bytes of the tile, in order
        if (cnt == 64u && (lo & 15u) == 0) {
            uint4* d = reinterpret_cast<uint4*>(s_gl + lo);
            d[0] = v[0]; d[1] = v[1]; d[2] = v[2]; d[3] = v[3];
        } else if (cnt) {
            uint32_t p = lo;
            for (int q = 0; q < 4; ++q) {
                const uint32_t w4[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
                for (int b = 0; b < nv[q]; ++b) {
                    const uint32_t byte = (w4[b >> 2] >> (8 * (b & 3))) & 255u;
                    if (job.strip && byte == '-') continue;
                    s_gl[p++] = (uint8_t)byte;
                }
            }
        }
    }
    __syncthreads();
    if (T == 0) return;
    const uint64_t o = s_o[0];

    if (exc) {                                                        // foreign bytes: (gapless position, byte) into the row's sorted list
        uint32_t* ep = job.exc_pos + (uint64_t)h * job.exc_stride;
        uint8_t* eb = job.exc_byte + (uint64_t)h * job.exc_stride;
        uint64_t e = s_o[1] + elo, p = o + lo;
        for (int q = 0; q < 4; ++q) {
            const uint32_t w4[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
            for (int b = 0; b < nv[q]; ++b) {
                const uint32_t byte = (w4[b >> 2] >> (8 * (b & 3))) & 255u;
                if (job.strip && byte == '-') continue;
                if (s_b2c[byte] == 0xFFu) {
                    if (e < job.exc_stride) { ep[e] = (uint32_t)p; eb[e] = (uint8_t)byte; }
                    ++e;
                }
                ++p;
            }
        }
    }
    // ---- the tile's words onto the output's word grid -------------------------------------------------------------------
    const uint32_t lead = (uint32_t)(o % SPW), sh = B * lead;
    const uint32_t nwords = (lead + T + SPW - 1) / SPW;
    uint64_t* X = job.X + job.xoff[h] + o / SPW;
    for (uint32_t j = tid; j < nwords; j += kPackThreads) {
        const uint64_t a = sw[j], b = sw[j + 1];                      // the tile's words j - 1 and j
        const uint64_t val = lead ? (a << (64 - sh)) | (b >> sh) : b;
        const bool edge = (j == 0 && lead != 0) || (j == nwords - 1 && ((lead + T) % SPW) != 0);
        if (edge) atomicOr(reinterpret_cast<unsigned long long*>(&X[j]), (unsigned long long)val);
        else X[j] = val;
    }
    if (gl_on) flush_bytes_block(s_gl, job.gapless + job.gapless_off[h] + o, T, tid, kPackThreads);
}

static void launch_pack_generic(const PackJob& job, cudaStream_t s, cudaEvent_t* ev) {
    dim3 grid(job.gtiles_per_row, job.H);
    uint64_t nt = (uint64_t)job.H * job.gtiles_per_row;
    if (ev) { cudaEventRecord(ev[0], s); cudaEventRecord(ev[1], s); }
    g_launches += 3;
    k_packg_count<<<grid, 256, 0, s>>>(job);
    uint64_t* tmp1 = job.gscan_tmp;
    uint64_t* tmp2 = job.gscan_tmp + scan_tmp_elems(nt);
    device_scan<uint64_t, LoadArr<uint32_t, uint64_t>, StoreArr<uint64_t>, OpSum, false>(
        LoadArr<uint32_t, uint64_t>{job.gtile_cnt}, StoreArr<uint64_t>{job.gtile_off}, nt, tmp1, OpSum(), 0ull, s);
    device_scan<uint64_t, LoadArr<uint32_t, uint64_t>, StoreArr<uint64_t>, OpSum, false>(
        LoadArr<uint32_t, uint64_t>{job.gtile_exc}, StoreArr<uint64_t>{job.gtile_excoff}, nt, tmp2, OpSum(), 0ull, s);
    uint64_t nts = (nt + kScanTile - 1) / kScanTile;
    k_packg_rows<<<(job.H + 127) / 128, 128, 0, s>>>(job, tmp1 + nts, tmp2 + nts);
    k_packg_write<<<grid, 256, 0, s>>>(job);
    if (ev) cudaEventRecord(ev[2], s);
}

void launch_pack(const PackJob& job, cudaStream_t s, cudaEvent_t* ev) {
    if (job.H == 0 || job.tiles_per_row == 0) { if (ev) for (int i = 0; i < 3; ++i) cudaEventRecord(ev[i], s); return; }
    if (job.bl != 1 && job.variant == 0) { launch_pack_generic(job, s, ev); return; }      // the two-pass kernels, kept for comparison
    uint64_t nt = (uint64_t)job.H * job.tiles_per_row;
    cudaMemsetAsync(job.desc_cnt, 0, nt * 8, s);
    cudaMemsetAsync(job.desc_exc, 0, nt * 8, s);
    cudaMemsetAsync(job.tile_counter, 0, 4, s);
    if (ev) { cudaEventRecord(ev[0], s); cudaEventRecord(ev[1], s); }
    if (job.bl != 1) {
        cudaMemsetAsync(job.exc_cnt, 0, (size_t)job.H * 8, s);
        const size_t smem = job.gapless ? (size_t)kPackTileBytes + 32 : 0;
        g_launches += 2;
        if (job.bl == 2) {
            if (job.occ == 6) k_packg1<2, 6><<<(unsigned)nt, kPackThreads, smem, s>>>(job, (uint32_t)nt);
            else if (job.occ == 4) k_packg1<2, 4><<<(unsigned)nt, kPackThreads, smem, s>>>(job, (uint32_t)nt);
            else if (job.occ == 3) k_packg1<2, 3><<<(unsigned)nt, kPackThreads, smem, s>>>(job, (uint32_t)nt);
            else k_packg1<2, 5><<<(unsigned)nt, kPackThreads, smem, s>>>(job, (uint32_t)nt);      // measured: 3 / 4 / 5 / 6 CTAs per SM = 2.13 / 1.82 / 1.74 / 1.74 ms
        } else k_packg1<3, 4><<<(unsigned)nt, kPackThreads, smem, s>>>(job, (uint32_t)nt);
        k_pack2_finish<<<(job.H + 127) / 128, 128, 0, s>>>(job);
        if (ev) cudaEventRecord(ev[2], s);
        return;
    }
    g_launches += 1;
    const unsigned grid = (unsigned)nt;          // one tile per CTA (4 per ticket measured 18x slower: a tile is published only
                                                 // after its CTA finished the earlier ones, which serialises every row's look-back)
    if (job.variant == 0) {
        if (job.occ >= 8) k_pack<8><<<grid, kPackThreads, 0, s>>>(job, (uint32_t)nt);
        else k_pack<6><<<grid, kPackThreads, 0, s>>>(job, (uint32_t)nt);
    } else if (job.variant == 2 && !job.gapless) {
        // k_pack3: one CTA per 8 tiles of a row; its descriptors (one per CTA) live at the head of the tile descriptors
        const uint32_t cpr = (job.tiles_per_row + kP3Pieces - 1) / kP3Pieces;
        const uint64_t nc = (uint64_t)job.H * cpr;
        cudaMemsetAsync(job.exc_cnt, 0, (size_t)job.H * 8, s);
        const size_t smem = (size_t)kPackWarps * kP3Pieces * kP3Stash * 4;
        if (job.occ == 5) k_pack3<5><<<(unsigned)nc, kPackThreads, smem, s>>>(job, (uint32_t)nc, cpr);
        else if (job.occ == 3) k_pack3<3><<<(unsigned)nc, kPackThreads, smem, s>>>(job, (uint32_t)nc, cpr);
        else k_pack3<4><<<(unsigned)nc, kPackThreads, smem, s>>>(job, (uint32_t)nc, cpr);      // 64 registers: measured best
        g_launches += 1;
        k_pack2_finish<<<(job.H + 127) / 128, 128, 0, s>>>(job);
    } else {
        cudaMemsetAsync(job.exc_cnt, 0, (size_t)job.H * 8, s);
        if (job.gapless) {
            static bool attr_set = false;
            const size_t smem = (size_t)kPackWarps * kGaplessStage;
            if (!attr_set) { cudaFuncSetAttribute(k_pack2<6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_set = true; }
            k_pack2<6, true><<<grid, kPackThreads, smem, s>>>(job, (uint32_t)nt);
        } else if (job.occ >= 8) k_pack2<8, false><<<grid, kPackThreads, 0, s>>>(job, (uint32_t)nt);
        else k_pack2<6, false><<<grid, kPackThreads, 0, s>>>(job, (uint32_t)nt);
        g_launches += 1;
        k_pack2_finish<<<(job.H + 127) / 128, 128, 0, s>>>(job);
    }
    if (ev) cudaEventRecord(ev[2], s);
}

// ---- gap-position side table (the consumer ScoreMatrix.scala:89-127 recomputes it from the rows) ----
// Every maximal run of '-' in a row gives two events, (column of its first byte, 0) and (column of its last byte, 1),
// appended to one list as keys  row << 33 | column << 1 | kind ; the caller sorts the list, after which every row's
// events alternate start, end, start, end ... in column order.  An optional pass over the rows (they are read once
// more; the product's hot path does not pay for it).
__global__ void __launch_bounds__(256) k_gap_events(const uint8_t* __restrict__ rows, const uint64_t* __restrict__ row_off,
                                                    const uint64_t* __restrict__ cols, uint64_t* __restrict__ list, uint64_t cap,
                                                    unsigned long long* __restrict__ counter) {
    const uint32_t h = blockIdx.y;
    const uint64_t nc = cols[h];
    const uint8_t* __restrict__ rowp = rows + row_off[h];
    for (uint64_t o = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16; o < nc; o += (uint64_t)gridDim.x * blockDim.x * 16) {
        const uint4 q = load16(rowp + o);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        const uint32_t nvalid = nc - o >= 16 ? 16u : (uint32_t)(nc - o);
        uint32_t mask = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) if ((uint32_t)j < nvalid && ((w[j >> 2] >> (8 * (j & 3))) & 255u) == (uint32_t)'-') mask |= 1u << j;
        if (!mask) continue;
        const uint32_t prev = (mask & 1u) && o > 0 && rowp[o - 1] == '-';
        const uint32_t next = (mask >> 15) && o + 16 < nc && rowp[o + 16] == '-';
        uint32_t starts = mask & ~((mask << 1) | prev), ends = mask & ~((mask >> 1) | (next << 15));
        const uint32_t cnt = (uint32_t)(__popc(starts) + __popc(ends));
        if (!cnt) continue;
        unsigned long long at = atomicAdd(counter, (unsigned long long)cnt);
        const uint64_t hi = (uint64_t)h << 33;
        while (starts) { const uint32_t b = (uint32_t)__ffs(starts) - 1; starts &= starts - 1; if (at < cap) list[at] = hi | ((o + b) << 1); ++at; }
        while (ends) { const uint32_t b = (uint32_t)__ffs(ends) - 1; ends &= ends - 1; if (at < cap) list[at] = hi | ((o + b) << 1) | 1u; ++at; }
    }
}

void launch_gap_events(const uint8_t* rows, const uint64_t* row_off, const uint64_t* cols, uint32_t H, uint64_t maxcols,
                       uint64_t* list, uint64_t cap, unsigned long long* counter, cudaStream_t s) {
    cudaMemsetAsync(counter, 0, 8, s);
    if (H == 0 || maxcols == 0) return;
    uint64_t want = (maxcols / 16 + 255) / 256;
    dim3 grid((unsigned)(want < 592 ? (want ? want : 1) : 592), H);
    g_launches += 1;
    k_gap_events<<<grid, 256, 0, s>>>(rows, row_off, cols, list, cap, counter);
}

// ---- alphabet of the reference: which of the 256 byte values occur (drlz_build_index) --------------------------------
// One bit per byte value, 8 words.  A thread looks at 16 bytes per step; a bit that is already set in the CTA's
// shared copy costs a shared-memory read, so after the first few hundred bytes the kernel is a plain HBM read.
__global__ void __launch_bounds__(256) k_byte_presence(const uint8_t* __restrict__ bytes, uint64_t n, uint32_t* __restrict__ presence) {
    __shared__ uint32_t sp[8];
    if (threadIdx.x < 8) sp[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t nvec = n >> 4;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 q = reinterpret_cast<const uint4*>(bytes)[v];
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const uint32_t c = (w[j] >> (8 * b)) & 255u;
                if (!((sp[c >> 5] >> (c & 31u)) & 1u)) atomicOr(&sp[c >> 5], 1u << (c & 31u));
            }
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 15)) {
        const uint32_t c = bytes[(nvec << 4) + threadIdx.x];
        atomicOr(&sp[c >> 5], 1u << (c & 31u));
    }
    __syncthreads();
    if (threadIdx.x < 8 && sp[threadIdx.x]) atomicOr(&presence[threadIdx.x], sp[threadIdx.x]);
}

void launch_byte_presence(const uint8_t* bytes, uint64_t n, uint32_t* presence, cudaStream_t s) {
    cudaMemsetAsync(presence, 0, 32, s);
    if (n == 0) return;
    uint64_t want = (n / 16 + 255) / 256;
    unsigned grid = (unsigned)(want < 148 * 8 ? (want ? want : 1) : 148 * 8);
    g_launches += 1;
    k_byte_presence<<<grid, 256, 0, s>>>(bytes, n, presence);
}

__global__ void k_make_hapviews(HapView* haps, const uint64_t* X, const uint64_t* xoff, const uint64_t* m,
                                const uint32_t* exc_pos, const uint8_t* exc_byte, const uint64_t* exc_cnt,
                                uint64_t exc_stride, uint32_t H) {
    uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= H) return;
    HapView hv;
    hv.x = X + xoff[h];
    hv.m = m[h];
    hv.exc_pos = exc_pos + (uint64_t)h * exc_stride;
    hv.exc_byte = exc_byte + (uint64_t)h * exc_stride;
    uint64_t ne = exc_cnt[h];
    hv.n_exc = (uint32_t)(ne < exc_stride ? ne : exc_stride);
    haps[h] = hv;
}

void launch_make_hapviews(HapView* haps, const uint64_t* X, const uint64_t* xoff, const uint64_t* m,
                          const uint32_t* exc_pos, const uint8_t* exc_byte, const uint64_t* exc_cnt, uint64_t exc_stride,
                          uint32_t H, cudaStream_t s) {
    if (H == 0) return;
    g_launches += 1;
    k_make_hapviews<<<(H + 127) / 128, 128, 0, s>>>(haps, X, xoff, m, exc_pos, exc_byte, exc_cnt, exc_stride, H);
}

}  // namespace drlz
