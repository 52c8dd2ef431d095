// rlz_core.cuh -- per-thread RLZ logic shared by every parse kernel.
//
// Everything here is __host__ __device__ so that tests/emu (a test-only host harness) can run the
// exact same statements on the CPU; the product (libdrlz.so) only ever calls them from kernels.
//
// Semantics reproduced (DistributedRLZ.scala:108-242, SURVEY.md 8(a) "clean spec"):
//   at haplotype position i:  L = longest l such that x[i..i+l) occurs in d
//                             pos = SA[leftmost index of the SA interval of x[i..i+L)]      (quirk Q1)
//                             L == 0  =>  literal (pos = byte value, len = 0)               (:210-211)
//   greedy: i += max(1, L)                                                                  (:226-240)
//
// Data layout (all in HBM):
//   text / haplotypes : B = 2, 4 or 8 bits per symbol (B = 1 << bl), 64/B symbols per 64-bit word, first symbol in
//                       the most significant bits, so integer order of words == lexicographic order of symbols;
//                       arrays carry >= 2 zero pad words.  B = 2: A,C,G,T -> 0..3 (pure-ACGT references, the fast
//                       path).  Otherwise the sigma distinct bytes of the reference get order-preserving codes
//                       0..sigma-1 (4 bits if sigma <= 16, else 8 bits).
//   SA                : uint32 suffix array of the reference (n < 2^32 - 1024)
//   tab               : sigma^k + 1 entries; tab[K] = #suffixes whose zero-padded radix-sigma k-mer key is < K
//                       (a jump table replacing the top ~2k levels of the SA binary search)
//   exceptions        : sorted positions (+ bytes) of haplotype bytes outside the reference alphabet; such a byte
//                       can never match, so it ends the running phrase and is a literal.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DRLZ_HD __host__ __device__ __forceinline__
#else
#define DRLZ_HD inline
#endif

#if !defined(__CUDACC__)
struct uint2 { uint32_t x, y; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 r; r.x = x; r.y = y; return r; }
#endif

namespace drlz {

#if !defined(__CUDACC__)
static uint64_t g_emu_long_hits;          // host harness (tests/emu) only: phrases taken from the extents across chunks
#endif

static const uint32_t kEmpty = 0xFFFFFFFFu;
static const int kAltSlots = 3;           // alternative entry records per chunk (see parse pipeline)
static const int kMaxRounds = 30;
static const int kHintGroup = 1;          // chunks per diagonal hint (a 64-base lookup at the chunk start)
static const int kMmMax = 14;             // mismatch offsets kept per chunk by the diagonal pre-scan
static const int kMmStride = 16;          // uint16 per chunk: [count, valid_until, offsets...]
static const int kAltStore = 4;           // own phrases of an alternative record kept for emission
static const uint32_t kExtInexact = 0x80000000u;   // ParseView::ext: the count is only a lower bound (see ext_term)

DRLZ_HD int clz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)v);
#else
    return __builtin_clzll(v);
#endif
}

// one word of symbols (64 / B of them) starting at symbol `pos`;  bl = log2(B)
DRLZ_HD uint64_t get64(const uint64_t* __restrict__ w, uint64_t pos, uint32_t bl) {
    const uint32_t sl = 6 - bl;
    uint64_t wi = pos >> sl;
    unsigned sh = (unsigned)(pos & ((1u << sl) - 1)) << bl;
    uint64_t a = w[wi];
    if (sh == 0) return a;
    return (a << sh) | (w[wi + 1] >> (64 - sh));
}

DRLZ_HD unsigned base_at(const uint64_t* __restrict__ w, uint64_t pos, uint32_t bl) {
    const uint32_t sl = 6 - bl, B = 1u << bl;
    return (unsigned)(w[pos >> sl] >> (64 - B - ((unsigned)(pos & ((1u << sl) - 1)) << bl))) & ((1u << B) - 1u);
}

struct IndexView {
    const uint64_t* text;   // packed reference
    const uint32_t* sa;
    const uint32_t* tab;    // sigma^k + 1
    uint64_t n;
    int k;
    const uint16_t* uniq;   // [n] min(65535, longest repeat starting at p) = max(LCP[ISA[p]], LCP[ISA[p]+1]); may be null
    uint32_t bl;            // log2(bits per symbol): 1, 2 or 3
    uint32_t sigma;         // alphabet size (4 when bl == 1)
    const uint8_t* c2b;     // code -> byte (literal phrases)
    const uint32_t* uniq32; // [n] the same without the 16-bit saturation; only present (non-null) if some repeat reaches 65535 symbols
    const uint16_t* uniqc;  // [ceil(n / kUniqBlock)] max of uniq over blocks of kUniqBlock positions (stays in L2); may be null
};
static const uint32_t kUniqBlockShift = 6, kUniqBlock = 1u << kUniqBlockShift;

DRLZ_HD uint32_t fsl32(uint32_t hi, uint32_t lo, unsigned s);

// key of the first kk symbols of the word xw padded with code 0 to k symbols, and the number of keys sharing that
// prefix: the table brackets are tab[lokey] and tab[lokey + span]
DRLZ_HD void prefix_keys(const IndexView& ix, uint64_t xw, uint32_t kk, uint64_t* lokey, uint64_t* span) {
    const int k = ix.k;
    if (ix.bl == 1) {
        uint64_t kmer = xw >> (64 - 2 * k);
        uint64_t lowmask = (1ull << (2 * (k - (int)kk))) - 1;
        *lokey = kmer & ~lowmask; *span = lowmask + 1;
        return;
    }
    // radix-sigma keys in 32-bit arithmetic (the table has at most 2^30 + 1 entries): the next symbol is always the top B
    // bits of the word, which moves up by one symbol per step (the 64-bit form of this loop was 12 % of k_spec's
    // instructions on a 4-bit index)
    const uint32_t B = 1u << ix.bl, sigma = ix.sigma;
    uint32_t hi = (uint32_t)(xw >> 32), lo = (uint32_t)xw, key = 0, sp = 1;
    for (uint32_t t = 0; t < kk; ++t) {
        key = key * sigma + (hi >> (32 - B));
        hi = fsl32(hi, lo, B); lo <<= B;
    }
    for (uint32_t t = kk; t < (uint32_t)k; ++t) { key *= sigma; sp *= sigma; }
    *lokey = key; *span = sp;
}

struct HapView {
    const uint64_t* x;        // packed haplotype
    uint64_t m;               // gapless length
    const uint32_t* exc_pos;  // sorted exception positions (may be null when n_exc == 0)
    const uint8_t* exc_byte;
    uint32_t n_exc;
};

struct Match {
    uint32_t len;   // 0 => literal
    uint32_t pos;   // SA[leftmost] or the literal byte
};

// high 32 bits of (hi:lo) << s, s in 0..31: one funnel-shift instruction on the device
DRLZ_HD uint32_t fsl32(uint32_t hi, uint32_t lo, unsigned s) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, s);
#else
    return s ? (hi << s) | (lo >> (32 - s)) : hi;
#endif
}

// shifted word: the 64 bits that start sh bits (0..63) into the 128-bit string a:b.  Done on 32-bit halves: the
// string is four halves, the result two funnel shifts of three of them (no 64-bit shifts, no special case for 0).
DRLZ_HD uint64_t funnel64(uint64_t a, uint64_t b, unsigned sh) {
    const uint32_t ah = (uint32_t)(a >> 32), al = (uint32_t)a, bh = (uint32_t)(b >> 32), bl_ = (uint32_t)b;
    const bool odd = sh >= 32;
    const unsigned s = sh & 31;
    const uint32_t w0 = odd ? al : ah, w1 = odd ? bh : al, w2 = odd ? bl_ : bh;
    return ((uint64_t)fsl32(w0, w1, s) << 32) | fsl32(w1, w2, s);
}

DRLZ_HD void prefetch_l2(const void* p) {
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// get64 with a 32-bit position (n, m < 2^32: half the integer work of the 64-bit form).  Always reads the next word
// too (every packed array ends in zero pad words): no branch on the alignment.
DRLZ_HD uint64_t get64u(const uint64_t* __restrict__ w, uint32_t pos, uint32_t bl) {
    const uint32_t sl = 6 - bl;
    uint32_t wi = pos >> sl;
    unsigned sh = (pos & ((1u << sl) - 1)) << bl;
    return funnel64(w[wi], w[wi + 1], sh);
}

// funnel64 for a shift that is fixed along a stream: ODD = (sh >= 32) picks the halves at compile time, s = sh & 31
template <bool ODD>
DRLZ_HD uint64_t funnel_fixed(uint64_t a, uint64_t b, unsigned s) {
    const uint32_t ah = (uint32_t)(a >> 32), al = (uint32_t)a, bh = (uint32_t)(b >> 32), bl_ = (uint32_t)b;
    return ODD ? ((uint64_t)fsl32(al, bh, s) << 32) | fsl32(bh, bl_, s)
               : ((uint64_t)fsl32(ah, al, s) << 32) | fsl32(al, bh, s);
}

// Streaming part of cmp_suffix_w: whole haplotype words xp[0..] against the reference words dp[0..] shifted by a
// fixed amount, from pattern offset `base` up to lim.  Returns the offset of the first difference (>= lim if none).
template <bool ODD>
DRLZ_HD uint32_t cmp_stream(const uint64_t* __restrict__ xp, const uint64_t* __restrict__ dp, unsigned s, uint32_t base,
                            uint32_t lim, uint32_t bl, bool* found, bool* diff_less) {
    const uint32_t spw = 64u >> bl;
    uint64_t cd = dp[0];
    while (base + 4 * spw <= lim) {                           // 4 words per iteration, 8 independent loads in flight
        uint64_t x0 = xp[0], x1 = xp[1], x2 = xp[2], x3 = xp[3];
        uint64_t d1 = dp[1], d2 = dp[2], d3 = dp[3], d4 = dp[4];
        uint64_t r0 = x0 ^ funnel_fixed<ODD>(cd, d1, s), r1 = x1 ^ funnel_fixed<ODD>(d1, d2, s);
        uint64_t r2 = x2 ^ funnel_fixed<ODD>(d2, d3, s), r3 = x3 ^ funnel_fixed<ODD>(d3, d4, s);
        if (r0 | r1 | r2 | r3) break;                         // the word loop below locates the mismatch
        base += 4 * spw; xp += 4; dp += 4; cd = d4;
    }
    while (base < lim) {
        const uint64_t a = xp[0], d1 = dp[1];
        const uint64_t b = funnel_fixed<ODD>(cd, d1, s);
        const uint64_t xr = a ^ b;
        if (xr) { base += (uint32_t)(clz64(xr) >> bl); *diff_less = b < a; *found = true; break; }
        base += spw; ++xp; ++dp; cd = d1;
    }
    return base;
}

// lcp(x[i..i+maxlen), d[p..n)) and whether the suffix d[p..] sorts before the pattern; xw0 = get64u(X, i).
// The first word is compared alone (most probes differ there); longer matches stream 4 words per iteration.
// dw0 = get64u(D, p): the caller may have requested it together with other candidates' first words
DRLZ_HD uint32_t cmp_suffix_pre(uint64_t xw0, uint64_t dw0, const uint64_t* __restrict__ X, uint32_t i,
                                const uint64_t* __restrict__ D, uint32_t p, uint32_t n, uint32_t maxlen, bool* less, uint32_t bl) {
    const uint32_t sl = 6 - bl, spw = 1u << sl;          // symbols per word
    uint32_t rem = n - p;
    uint32_t lim = rem < maxlen ? rem : maxlen;
    uint32_t l = 0;
    bool diff_less = false, found = false;
    if (lim > 0) {
        uint64_t b = dw0;
        uint64_t xr = xw0 ^ b;
        if (xr) { l = (uint32_t)(clz64(xr) >> bl); diff_less = b < xw0; found = true; }
        else l = spw;
    }
    if (!found && l < lim) {
        // Continue on the grid of the haplotype's packed words: position i + l lies q symbols into a word, and those q
        // symbols were just verified equal, so whole haplotype words are compared and only the reference is shifted
        // onto that grid (one funnel shift per word).
        const uint32_t q = (i + l) & (spw - 1);
        uint32_t base = l - q;                                // pattern offset of that word's first symbol (l == spw > q)
        const uint64_t* xp = X + ((i + base) >> sl);
        const uint32_t pd = p + base;
        const unsigned shd = (pd & (spw - 1)) << bl;
        const uint64_t* dp = D + (pd >> sl);
        base = shd >= 32 ? cmp_stream<true>(xp, dp, shd & 31, base, lim, bl, &found, &diff_less)
                         : cmp_stream<false>(xp, dp, shd & 31, base, lim, bl, &found, &diff_less);
        l = base;
    }
    if (l >= lim) {
        l = lim;
        // pattern exhausted => suffix >= pattern; else the suffix ran out => suffix < pattern
        *less = (lim != maxlen);
    } else {
        *less = diff_less;
    }
    return l;
}

DRLZ_HD uint32_t cmp_suffix_w(uint64_t xw0, const uint64_t* __restrict__ X, uint32_t i, const uint64_t* __restrict__ D,
                              uint32_t p, uint32_t n, uint32_t maxlen, bool* less, uint32_t bl) {
    return cmp_suffix_pre(xw0, get64u(D, p, bl), X, i, D, p, n, maxlen, less, bl);
}

DRLZ_HD uint32_t cmp_suffix(const uint64_t* __restrict__ X, uint64_t i, const uint64_t* __restrict__ D,
                            uint64_t p, uint64_t n, uint32_t maxlen, bool* less, uint32_t bl) {
    return cmp_suffix_w(get64u(X, (uint32_t)i, bl), X, (uint32_t)i, D, (uint32_t)p, (uint32_t)n, maxlen, less, bl);
}

// Longest match of x[i..i+maxlen) in d with the reference's tie-break.  maxlen >= 1.
// cap (>= 1) bounds the compared length: if the pattern's first min(cap, maxlen) bases occur in d, cap < maxlen
// and exactly ONE suffix carries them, the call stops there and reports *open = true with len = cap -- the
// match continues along that unique alignment and its full length is the diagonal run resolved by the caller.
// If several suffixes share the capped prefix the search is redone uncapped (exact, slower; long exact repeats).
// probe_only: give up (len = 0) instead of redoing the search uncapped when several suffixes carry the capped pattern --
// for callers that only want a prediction (k_hint): inside a run of millions of N every uncapped compare streams the run.
DRLZ_HD Match lookup(const IndexView& ix, const uint64_t* __restrict__ X, uint64_t i64, uint32_t maxlen,
                     uint32_t cap, bool* open, bool probe_only = false) {
    const int k = ix.k;
    const uint32_t n = (uint32_t)ix.n;
    const uint32_t i = (uint32_t)i64;
    const uint64_t* __restrict__ D = ix.text;
    const uint32_t* __restrict__ SA = ix.sa;
    *open = false;
    const uint32_t bl = ix.bl;
    const uint64_t xw = get64u(X, i, bl);
    Match mt;
    uint32_t plen = cap < maxlen ? cap : maxlen;
    for (int pass = 0; pass < 2; ++pass) {
        uint32_t kk = plen < (uint32_t)k ? plen : (uint32_t)k;
        uint64_t lokey, span;
        prefix_keys(ix, xw, kk, &lokey, &span);
        uint32_t lo = ix.tab[lokey], hi = ix.tab[lokey + span];
        // r = #suffixes < pattern lies in [lo, hi]
        uint32_t l = lo, h = hi;
        uint32_t lcp_left = 0, lcp_right = 0, sar = 0;
        bool have_left = false, have_right = false;
        while (l < h) {
            uint32_t mid = l + ((h - l) >> 1);
            bool less;
            uint32_t pm = SA[mid];
            // every suffix between two that share `skip` symbols with the pattern shares them too: inside a long repeat
            // (the uncapped pass over a run of N) the compares go on where the bounds left off instead of streaming the
            // repeat again in every step
            const uint32_t skip = (have_left && have_right) ? (lcp_left < lcp_right ? lcp_left : lcp_right) : 0u;
            uint32_t c;
            if (skip >= 64u && skip < plen && skip < n - pm)
                c = skip + cmp_suffix(X, (uint64_t)i + skip, D, (uint64_t)pm + skip, n, plen - skip, &less, bl);
            else c = cmp_suffix_w(xw, X, i, D, pm, n, plen, &less, bl);
            if (less) { l = mid + 1; lcp_left = c; have_left = true; }
            else { h = mid; lcp_right = c; have_right = true; sar = pm; }
        }
        uint32_t r = l;
        bool dummy;
        uint32_t l1 = 0, l2 = 0;
        // the candidates are the neighbours r-1 and r (and r+1 for the uniqueness test of a capped match): their SA
        // entries share a line and their first reference words are requested together, so the compares below start
        // from data that arrives in two round trips instead of six
        const bool needL = r > 0 && !have_left, needR = r < n && !have_right, haveN = r + 1 < n;
        const uint32_t pL = needL ? SA[r - 1] : 0u;
        if (needR) sar = SA[r];
        const uint32_t pN = haveN ? SA[r + 1] : 0u;
        const uint64_t wL = needL ? get64u(D, pL, bl) : 0ull, wR = needR ? get64u(D, sar, bl) : 0ull;
        const uint64_t wN = haveN ? get64u(D, pN, bl) : 0ull;
        if (r > 0) l1 = have_left ? lcp_left : cmp_suffix_pre(xw, wL, X, i, D, pL, n, plen, &dummy, bl);
        if (r < n) {
            if (!have_right) lcp_right = cmp_suffix_pre(xw, wR, X, i, D, sar, n, plen, &dummy, bl);
            l2 = lcp_right;
        }
        if (l2 > l1) {
            if (l2 == plen && plen < maxlen) {
                // capped: is SA[r] the only suffix carrying the capped pattern?
                bool multi = false;
                if (haveN) multi = cmp_suffix_pre(xw, wN, X, i, D, pN, n, plen, &dummy, bl) >= plen;
                if (multi) { if (probe_only) { mt.len = 0; mt.pos = 0; return mt; } plen = maxlen; continue; }
                *open = true;
            }
            mt.len = l2; mt.pos = sar; return mt;
        }
        if (l1 == 0) { mt.len = 0; mt.pos = 0; return mt; }
        // leftmost index whose suffix shares >= L bases with the pattern; it lies in [tab[key(L)], r-1]
        uint32_t L = l1;
        uint32_t Lk = L < (uint32_t)k ? L : (uint32_t)k;
        uint32_t a = lo;
        if (Lk != kk) {
            uint64_t lk2, sp2;
            prefix_keys(ix, xw, Lk, &lk2, &sp2);
            a = ix.tab[lk2];
        }
        uint32_t b = r - 1;
        if (a < b) {
            // The suffixes that start with the pattern's first L symbols are the whole table bracket of that prefix,
            // except for at most L-1 shorter suffixes at its start whose zero padding imitates it: the bracket's
            // first entry is the answer nearly always -- one probe instead of log2(bracket) dependent SA / text misses.
            const uint32_t pa = SA[a];
            if (cmp_suffix_w(xw, X, i, D, pa, n, L, &dummy, bl) >= L) { mt.len = L; mt.pos = pa; return mt; }
            ++a;
        }
        while (a < b) {
            uint32_t mid = a + ((b - a) >> 1);
            uint32_t c = cmp_suffix_w(xw, X, i, D, SA[mid], n, L, &dummy, bl);
            if (c >= L) b = mid; else a = mid + 1;
        }
        mt.len = L; mt.pos = SA[a];
        return mt;
    }
    mt.len = 0; mt.pos = 0;      // not reached
    return mt;
}

// lookup() for a pattern of at least kLookShort symbols whose longest match is shorter than kLookShort, on a 2-bit
// index, in a fixed number of dependent steps: the table bracket of the pattern's k-mer holds at most four suffixes, so
// the six suffix-array entries [lo-1, lo+4] contain both neighbours of the pattern's rank; the longest match is the best
// of their first words.  Its leftmost suffix: for a match of k symbols or more it is among those entries; for a shorter
// one it is the first entry of the table bracket of the matched prefix that is long enough to carry it (shorter suffixes
// imitate the prefix through their zero padding and sort in front; a suffix of that bracket that is long enough shares
// the prefix).  Returns false if it cannot answer (then lookup() does).
static const int kLookCand = 6;
static const uint32_t kLookShort = 32;         // symbols of one 2-bit word (wider symbols: 64 >> bl, see lookup_short_len)
DRLZ_HD uint32_t lookup_short_len(uint32_t bl) { return 64u >> bl; }
// xw = the pattern's first word (64 >> bl symbols).  Symbols of 2 bits: keys are bit fields; wider symbols (references with
// N: 4 bits): radix-sigma keys through prefix_keys(), one word holds 16 symbols, so the answer covers matches below 16.
DRLZ_HD bool lookup_short(const IndexView& ix, uint64_t xw, Match* out) {
    const int k = ix.k;
    const uint32_t n = (uint32_t)ix.n, bl = ix.bl, wsym = 64u >> bl;
    if ((uint32_t)k > wsym) return false;
    uint64_t key, span;
    if (bl == 1) key = xw >> (64 - 2 * k);
    else prefix_keys(ix, xw, (uint32_t)k, &key, &span);
    const uint32_t lo = ix.tab[key], hi = ix.tab[key + 1];
    if (hi - lo > (uint32_t)kLookCand - 2u) return false;
    uint32_t best = 0, lcp[kLookCand], sp[kLookCand];
    bool val[kLookCand];
#pragma unroll
    for (int c = 0; c < kLookCand; ++c) {                    // entries [lo-1, hi]: both neighbours of the pattern's rank are among them
        const uint32_t q = lo + (uint32_t)c - 1u;
        val[c] = (c > 0 || lo > 0) && q <= hi && q < n;
        sp[c] = val[c] ? ix.sa[q] : 0u;
    }
#pragma unroll
    for (int c = 0; c < kLookCand; ++c) {
        lcp[c] = 0;
        if (val[c]) {
            const uint64_t xr = xw ^ get64u(ix.text, sp[c], bl);
            const uint32_t l = xr ? (uint32_t)(clz64(xr) >> bl) : wsym, avail = n - sp[c];
            lcp[c] = l < avail ? l : avail;
        }
        if (lcp[c] > best) best = lcp[c];
    }
    if (best >= wsym) return false;
    out->len = best; out->pos = 0;
    if (best == 0) return true;                              // literal; the caller fills in the byte
    if (best >= (uint32_t)k) {                               // the interval lies inside the bracket: entries 1 .. hi - lo of the candidates
        bool found = false;
#pragma unroll
        for (int c = kLookCand - 2; c >= 1; --c) if (lcp[c] >= best) { out->pos = sp[c]; found = true; }
        return found;
    }
    uint64_t pkey;
    if (bl == 1) pkey = key & ~((1ull << (2 * (k - (int)best))) - 1);
    else prefix_keys(ix, xw, best, &pkey, &span);
    const uint32_t a = ix.tab[pkey];                         // first suffix whose padded k-mer starts with the matched prefix
    if (a + 1 >= n) return false;
    const uint32_t p0 = ix.sa[a], p1 = ix.sa[a + 1];
    if (p0 + best <= n) { out->pos = p0; return true; }
    if (p1 + best <= n) { out->pos = p1; return true; }
    return false;
}

// index of the first exception at or after position i (binary search)
DRLZ_HD uint32_t exc_lower_bound(const HapView& hv, uint64_t i) {
    uint32_t a = 0, b = hv.n_exc;
    while (a < b) {
        uint32_t mid = a + ((b - a) >> 1);
        if ((uint64_t)hv.exc_pos[mid] < i) a = mid + 1; else b = mid;
    }
    return a;
}

// ---------------------------------------------------------------------------------------------
// Chunk-parallel exact greedy parse.
//
// The greedy parse is a deterministic function of its start position, so two parses that ever share
// a phrase start coincide from there on.  Each haplotype is cut into chunks of C bases:
//   A   every chunk looks up the phrase at its first base, comparing no further than the chunk end.
//       A match that reaches the chunk end is "open": it continues along its alignment (diagonal
//       pos - i).  Chunks whose successor starts on the same diagonal are linked; list ranking by
//       pointer jumping turns the links into exact run lengths (run_len) -- so no thread ever compares
//       more than ~one chunk of bases, however long the match (a haplotype identical to the reference is
//       one 48 Mbp phrase).
//   P1  every chunk parses speculatively from its first base            -> spec_cnt, spec_exit
//   R   every exit position e is registered as an "alternative entry" of the chunk containing e
//   F   every new alternative entry is parsed in tandem with the chunk's speculative parse until the
//       two share a phrase start (merge) or the chunk ends              -> alt_cnt, alt_exit
//       (R and F repeat until no new entry appears; two rounds in practice)
//   J   records form a functional graph (record -> record entered at its exit); pointer jumping from
//       (chunk 0, speculative record) marks the true chain
//   E   per chunk the marked record's phrase count is prefix-summed (stream compaction) and the chunk is
//       parsed once more from its true entry straight into the output slots.
// ---------------------------------------------------------------------------------------------

struct ParseView {
    IndexView ix;
    const HapView* haps;            // [H]
    const uint32_t* chunk_hap;      // [NC] haplotype of global chunk g
    const uint32_t* hap_chunk_base; // [H+1]
    uint32_t chunk_bases;           // C = 2^chunk_shift (0xFFFFFFFF when chunk_shift == 32: one chunk per haplotype)
    uint32_t chunk_shift;           // chunk index of position x = x >> chunk_shift (no divisions in the kernels)
    uint32_t n_chunks;              // NC
    uint32_t min_cap;               // lower bound of the compare cap (64; 0xFFFFFFFF disables capping)
    uint32_t foreign_cap;           // bound of an off-chain diagonal compare before the exact fallback is taken
    // stage A
    uint32_t* first_pos;            // [NC] phrase at the chunk's first base: position (or literal byte)
    uint32_t* first_len;            // [NC] its verified length (0 = literal); >= chunk end when open
    uint32_t* run_len;              // [NC] its exact total length (after list ranking)
    uint32_t* link;                 // [NC] next chunk of the diagonal run, kEmpty = end
    uint32_t* open_i;               // [NC] start of the chunk's last phrase if it is still open (kEmpty otherwise)
    uint32_t* open_pos;             // [NC] its reference position
    uint32_t* open_len;             // [NC] its verified length
    // stages P1 / R / F
    uint32_t* spec_cnt;             // [NC]
    uint32_t* spec_exit;            // [NC]
    uint32_t* alt_entry;            // [NC*kAltSlots], kEmpty when unused
    uint32_t* alt_cnt;              // [NC*kAltSlots]
    uint32_t* alt_exit;             // [NC*kAltSlots]
    uint32_t* alt_round;            // [NC*kAltSlots] round in which the record must be fix-parsed
    uint32_t* alt_own;              // [NC*kAltSlots] phrases before the merge point
    uint32_t* alt_merge;            // [NC*kAltSlots] index into the speculative phrase list where the tail starts, kEmpty = no merge
    uint2* spec_ph;                 // [NC*spec_store] (pos, len) of the first phrases of the speculative parse
    uint32_t spec_store;            // phrases of a chunk's speculative parse kept for emission
    const int64_t* hint;            // [ceil(NC / kHintGroup)] predicted diagonal of each chunk group (kNoDiag = none); may be null
    uint16_t* mm;                   // [NC*kMmStride] mismatch offsets of each chunk along its hinted diagonal; may be null
    uint2* alt_ph;                  // [NC*kAltSlots*kAltStore] own phrases of the alternative records
    uint32_t* counters;             // [0] overflow flag, [1 + r] records created for round r
    uint32_t parsed_rounds;         // fix-up rounds run so far: records registered for a later round are not parsed yet
    const uint32_t* ext;            // [NC] symbols from the START of chunk g to the first mismatch along its hinted diagonal, counted through
                                    // the following chunks as long as they lie on the same diagonal (0 = unknown); may be null
    uint32_t spec_sync;             // k_spec: the chunks of a warp parsed in step (phrase heads together, lookups together)
};

DRLZ_HD uint32_t atomic_cas_u32(uint32_t* p, uint32_t cmp, uint32_t val) {
#if defined(__CUDA_ARCH__)
    return atomicCAS(p, cmp, val);
#else
    uint32_t old = *p; if (old == cmp) *p = val; return old;
#endif
}
DRLZ_HD void atomic_inc_u32(uint32_t* p) {
#if defined(__CUDA_ARCH__)
    atomicAdd(p, 1u);
#else
    ++*p;
#endif
}

// first position past i where the running phrase must stop (next exception or end of the haplotype)
DRLZ_HD uint64_t stop_after(const HapView& hv, uint64_t i, uint32_t* cursor) {
    uint64_t stop = hv.m;
    if (hv.n_exc) {
        uint32_t c = *cursor;
        while (c < hv.n_exc && (uint64_t)hv.exc_pos[c] < i) ++c;
        *cursor = c;
        if (c < hv.n_exc) stop = hv.exc_pos[c];
    }
    return stop;
}

// capped phrase lookup at position i (chunk-end cap); literals carry their byte in .pos
static const int64_t kNoDiag = INT64_MIN;
static const uint32_t kDiagMinLen = 32;   // phrases at least this long update the predicted diagonal

// capped phrase lookup at position i (chunk-end cap); literals carry their byte in .pos.
// *diag (in/out, kNoDiag = none): the alignment pos - i of the last long phrase.  The next long phrase nearly always
// continues on it (a SNP costs one short phrase in between), so the candidate p = i + *diag is tried first: if the
// match along that alignment is longer than uniq[p] -- the longest repeat starting at p anywhere else in the
// reference -- then every other suffix shares fewer bases with the pattern, i.e. p is the unique longest match and
// the SA interval is a singleton: exactly what lookup() would return, for one 2-byte read instead of the
// table -> SA -> text probe chain.  Anything else falls through to lookup().
// verified length of the match of x[i..i+plen) along the chunk's hinted diagonal, from the pre-scanned mismatch list;
// returns false if the list cannot answer (position past the classified range, or plen reaches beyond the chunk)
DRLZ_HD bool list_extent(const uint16_t* mm, uint64_t s, uint64_t i, uint32_t plen, uint32_t* c_out) {
    uint32_t cnt = mm[0], V = mm[1];
    uint32_t off = (uint32_t)(i - s);
    if (off >= V) return false;
    uint32_t c = V - off;
    bool closed = false;
    for (uint32_t t = 0; t < cnt; ++t) {
        uint32_t mo = mm[2 + t];
        if (mo >= off) { c = mo - off; closed = true; break; }
    }
    if (c >= plen) { *c_out = plen; return true; }
    if (closed) { *c_out = c; return true; }
    return false;                                            // classified range ends before the pattern does
}

// mml / hint_g: the chunk's mismatch record and hinted diagonal (mml == nullptr: none).  k_spec keeps its copy of the
// record in shared memory -- it is read once per predicted phrase, and in L1 it does not survive between two of them.
// The phrase at i comes in two halves, so that k_spec can run each with its warp converged: phrase_head() answers the
// literals at exceptions and the predicted phrases (one read of the uniqueness array), phrase_tail() is the index lookup.
enum { kHeadLiteral = 0, kHeadPredicted = 1, kHeadLookup = 2 };
DRLZ_HD int phrase_head(const ParseView& pv, const HapView& hv, uint64_t i, uint32_t* cursor, bool* open, int64_t* diag,
                        const uint16_t* mml, int64_t hint_g, uint64_t s, uint32_t g, Match* out, uint32_t* maxlen_out, uint32_t* cap_out) {
    *open = false;
    uint64_t stop = stop_after(hv, i, cursor);
    if (stop == i) { out->len = 0; out->pos = hv.exc_byte[*cursor]; return kHeadLiteral; }
    uint32_t maxlen = (uint32_t)(stop - i);
    uint64_t to_end = (((i >> pv.chunk_shift) + 1) << pv.chunk_shift) - i;
    uint32_t cap = to_end > (uint64_t)pv.min_cap ? (to_end > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)to_end) : pv.min_cap;
    *maxlen_out = maxlen; *cap_out = cap;
    Match mt;
    if (pv.ix.uniq && *diag != kNoDiag) {
        int64_t p64 = (int64_t)i + *diag;
        if (p64 >= 0 && (uint64_t)p64 < pv.ix.n) {
            uint32_t p = (uint32_t)p64;
            uint32_t plen = cap < maxlen ? cap : maxlen;
            uint32_t c;
            bool less;
            // the list first (shared memory): at a listed variant site nothing is predicted, and a read of the uniqueness
            // array whose answer nobody wants still holds its register -- and the warp -- until it returns
            const bool listed = mml && hint_g == *diag && list_extent(mml, s, i, plen, &c);
            if (listed && c == 0) return kHeadLookup;
            // The test is c > uniq[p], and uniq[p] is a DRAM miss as a rule (2 bytes of a sector at a random place of an
            // array the size of the reference; k_spec runs at the rate HBM serves such reads).  A phrase of hundreds of
            // bases against a longest repeat of a dozen: the maximum of uniq over the 64 positions around p, from an array
            // small enough to stay in L2, nearly always settles it.
            uint32_t u = pv.ix.uniqc ? pv.ix.uniqc[p >> kUniqBlockShift] : 0xFFFFu;
            if (!(listed && u != 0xFFFFu && c > u)) u = pv.ix.uniq[p];
            if (!listed)
                c = cmp_suffix_w(get64u(hv.x, (uint32_t)i, pv.ix.bl), hv.x, (uint32_t)i, pv.ix.text, p, (uint32_t)pv.ix.n, plen, &less, pv.ix.bl);
            // The predicted path never touches the haplotype (the mismatch list answers for it), so the lookup at the variant
            // site behind a predicted phrase would start with a miss on its own pattern: ask for it while the uniqueness
            // array is still answering (the phrase after this one starts at i + c, whatever it is)
            if (c < plen) prefetch_l2(hv.x + (((uint32_t)i + c) >> (6 - pv.ix.bl)));
            if (c > 0) {
                if (u != 0xFFFFu && c > u) {
                    mt.len = c; mt.pos = p;
                    *open = (c == plen && plen < maxlen);
                    *out = mt; return kHeadPredicted;
                }
                // Long repeats (N runs of millions of bases, segmental duplications): the match reaches the cap and the
                // capped pattern is NOT unique (the longest repeat at p is at least as long), so lookup() would search
                // again uncapped -- streaming the whole repeat in every step of its binary search, from every chunk start
                // inside it.  The pre-scan knows how far the match along this diagonal really goes (this chunk is clean up
                // to its end, ext[] counts on through the chunks behind it); if that is longer than the longest repeat at p
                // it is the unique longest match of the uncapped pattern: what the second pass of lookup() returns.
                if (c == plen && plen < maxlen && pv.ext && g != kEmpty && g + 1 < pv.n_chunks) {
                    const uint32_t ur = u != 0xFFFFu ? u : (pv.ix.uniq32 ? pv.ix.uniq32[p] : 0xFFFFFFFFu);
                    if (ur >= plen && ur != 0xFFFFFFFFu && pv.hint && pv.hint[(g + 1) / kHintGroup] == *diag) {
                        const uint32_t ex = pv.ext[g + 1];
                        uint64_t cf = to_end + (uint64_t)(ex & ~kExtInexact);
                        if (cf > maxlen) cf = maxlen;
                        if (!(ex & kExtInexact) && cf > ur && cf >= plen) {
#if !defined(__CUDACC__)
                            ++g_emu_long_hits;
#endif
                            mt.len = (uint32_t)cf; mt.pos = p; *out = mt; return kHeadPredicted;
                        }
                    }
                }
            }
        }
    }
    return kHeadLookup;
}

DRLZ_HD Match phrase_tail(const ParseView& pv, const HapView& hv, uint64_t i, uint32_t maxlen, uint32_t cap, bool* open, int64_t* diag) {
    Match mt;
    // most lookups of the parse stand on a variant site: a match of a dozen symbols somewhere else, which the short form
    // finds in five dependent loads instead of eight
    if (!(pv.ix.bl <= 2 && (cap < maxlen ? cap : maxlen) >= lookup_short_len(pv.ix.bl) &&
          lookup_short(pv.ix, get64u(hv.x, (uint32_t)i, pv.ix.bl), &mt)))
        mt = lookup(pv.ix, hv.x, i, maxlen, cap, open);
    if (mt.len == 0) mt.pos = (uint32_t)pv.ix.c2b[base_at(hv.x, i, pv.ix.bl)];
    else if (mt.len >= kDiagMinLen) *diag = (int64_t)mt.pos - (int64_t)i;
    return mt;
}

DRLZ_HD Match phrase_capped_l(const ParseView& pv, const HapView& hv, uint64_t i, uint32_t* cursor, bool* open, int64_t* diag,
                              const uint16_t* mml, int64_t hint_g, uint64_t s, uint32_t g = kEmpty) {
    Match mt; uint32_t maxlen, cap;
    if (phrase_head(pv, hv, i, cursor, open, diag, mml, hint_g, s, g, &mt, &maxlen, &cap) != kHeadLookup) return mt;
    return phrase_tail(pv, hv, i, maxlen, cap, open, diag);
}

DRLZ_HD Match phrase_capped(const ParseView& pv, const HapView& hv, uint64_t i, uint32_t* cursor, bool* open, int64_t* diag,
                            uint32_t g, uint64_t s) {
    const bool listed = pv.mm && g != kEmpty;
    return phrase_capped_l(pv, hv, i, cursor, open, diag, listed ? pv.mm + (size_t)g * kMmStride : nullptr,
                           listed ? pv.hint[g / kHintGroup] : kNoDiag, s, g);
}

// Where an open match (start i, verified len, position pos) continues: the last chunk start j <= i+len.
// Returns the global chunk id; *j_out = its first base; *same = that chunk's own first phrase lies on the
// same diagonal (then the remaining length is that chunk's run).
DRLZ_HD uint32_t open_target(const ParseView& pv, uint32_t h, uint64_t i, const Match& mt, uint64_t* j_out, bool* same) {
    uint64_t cc = (i + mt.len) >> pv.chunk_shift;
    uint64_t j = cc << pv.chunk_shift;
    uint32_t g2 = pv.hap_chunk_base[h] + (uint32_t)cc;
    *j_out = j;
    *same = pv.first_len[g2] > 0 && (int64_t)pv.first_pos[g2] - (int64_t)j == (int64_t)mt.pos - (int64_t)i;
    return g2;
}

// off-chain continuation: compare along the diagonal from j; flags the exact fallback if it runs too long
DRLZ_HD uint32_t foreign_run(const ParseView& pv, const HapView& hv, uint64_t i, const Match& mt, uint64_t j) {
    uint32_t cur = hv.n_exc ? exc_lower_bound(hv, j) : 0;
    uint64_t stop = stop_after(hv, j, &cur);
    uint32_t maxlen = (uint32_t)(stop - j);
    uint32_t cap = maxlen < pv.foreign_cap ? maxlen : pv.foreign_cap;
    bool less;
    uint64_t p = (uint64_t)mt.pos + (j - i);
    uint32_t l = cap ? cmp_suffix(hv.x, j, pv.ix.text, p, pv.ix.n, cap, &less, pv.ix.bl) : 0;
    if (l == cap && cap < maxlen) pv.counters[0] = 3;        // too long for one thread: take the exact fallback
    return l;
}

// exact phrase at position i (uses run_len of the chunk starts; valid after stage A)
DRLZ_HD Match phrase_at(const ParseView& pv, uint32_t h, const HapView& hv, uint64_t i, uint32_t* cursor, int64_t* diag,
                        uint32_t g, uint64_t s) {
    bool open;
    Match mt = phrase_capped(pv, hv, i, cursor, &open, diag, g, s);
    if (open) {
        uint64_t j; bool same;
        uint32_t g2 = open_target(pv, h, i, mt, &j, &same);
        uint32_t rest = same ? pv.run_len[g2] : foreign_run(pv, hv, i, mt, j);
        mt.len = (uint32_t)(j - i) + rest;
    }
    return mt;
}

// A2: link open chunks to the chunk their match continues in, or finish them by a direct compare
DRLZ_HD void link_chunk(const ParseView& pv, uint32_t g) {
    if (pv.run_len[g] != kEmpty) return;
    uint32_t h = pv.chunk_hap[g];
    const HapView& hv = pv.haps[h];
    uint64_t s = (uint64_t)(g - pv.hap_chunk_base[h]) << pv.chunk_shift;
    Match mt; mt.pos = pv.first_pos[g]; mt.len = pv.first_len[g];
    uint64_t j; bool same;
    uint32_t g2 = open_target(pv, h, s, mt, &j, &same);
    if (same) { pv.link[g] = g2; pv.run_len[g] = (uint32_t)(j - s); }          // partial sum; rest follows the link
    else pv.run_len[g] = (uint32_t)(j - s) + foreign_run(pv, hv, s, mt, j);
}

// A3: one pointer-jumping round of the list ranking (double buffered by the caller)
DRLZ_HD void rank_round(const uint32_t* link_in, const uint32_t* run_in, uint32_t* link_out, uint32_t* run_out, uint32_t g) {
    uint32_t l = link_in[g];
    if (l == kEmpty) { link_out[g] = kEmpty; run_out[g] = run_in[g]; return; }
    run_out[g] = run_in[g] + run_in[l];
    link_out[g] = link_in[l];
}

// R: register exit position x of haplotype h as an entry; new records are scheduled for `next_round`.
DRLZ_HD void register_exit(const ParseView& pv, uint32_t h, uint64_t x, uint32_t next_round) {
    const HapView& hv = pv.haps[h];
    if (x >= hv.m) return;                                   // parse finished
    if ((x & ((1ull << pv.chunk_shift) - 1)) == 0) return;   // that is the chunk's speculative record
    uint32_t g = pv.hap_chunk_base[h] + (uint32_t)(x >> pv.chunk_shift);
    for (int s = 0; s < kAltSlots; ++s) {
        uint32_t old = atomic_cas_u32(&pv.alt_entry[g * kAltSlots + s], kEmpty, (uint32_t)x);
        if (old == kEmpty) {
            pv.alt_round[g * kAltSlots + s] = next_round;
            atomic_inc_u32(&pv.counters[1 + (next_round < (uint32_t)kMaxRounds ? next_round : kMaxRounds)]);
            return;
        }
        if (old == (uint32_t)x) return;
    }
    pv.counters[0] = 1;                                      // more distinct entries than slots
}

// one greedy step of the parse of chunk g at position i (i == chunk start reuses stage A's result)
DRLZ_HD Match chunk_step(const ParseView& pv, uint32_t g, uint32_t h, const HapView& hv, uint64_t s, uint64_t i, uint32_t* cursor,
                         int64_t* diag) {
    if (i == s) {
        Match mt; mt.pos = pv.first_pos[g]; mt.len = pv.run_len[g];
        if (mt.len >= kDiagMinLen) *diag = (int64_t)mt.pos - (int64_t)i;
        return mt;
    }
    return phrase_at(pv, h, hv, i, cursor, diag, g, s);
}

// H: diagonal hint of chunk group q = alignment of the phrase at the first base of its first chunk (64-base probe)
DRLZ_HD void hint_group(const ParseView& pv, int64_t* hint, uint32_t q) {
    uint32_t g = q * kHintGroup;
    hint[q] = kNoDiag;
    if (g >= pv.n_chunks) return;
    uint32_t h = pv.chunk_hap[g];
    const HapView& hv = pv.haps[h];
    uint64_t s = (uint64_t)(g - pv.hap_chunk_base[h]) << pv.chunk_shift;
    if (s >= hv.m) return;
    uint32_t cur = hv.n_exc ? exc_lower_bound(hv, s) : 0;
    uint64_t stop = stop_after(hv, s, &cur);
    if (stop == s) return;
    uint64_t e = s + (1ull << pv.chunk_shift); if (e > hv.m) e = hv.m;
    // A chunk that starts inside a run of one symbol (the N runs of a reference assembly): a match from inside a run is
    // placed by what follows the run, so the probe goes behind the run's end if the chunk sees it; the chunks further
    // inside, which see nothing but the run, get no hint of their own and inherit this diagonal from behind (k_hint_fill,
    // right to left first).  The end is found by bisection over the chunk's words (first word all run, last word not): a
    // second run behind the first can mislead it to a later boundary, which is as good a place for a prediction.
    const uint32_t bl = pv.ix.bl, sl = 6 - bl, spw = 1u << sl;
    const uint64_t w0 = get64u(hv.x, (uint32_t)s, bl);
    const uint64_t unit = bl == 1 ? 0x5555555555555555ull : (bl == 2 ? 0x1111111111111111ull : 0x0101010101010101ull);
    const uint64_t rep = (w0 >> (64 - (1u << bl))) * unit;
    uint64_t first = s;
    if (w0 == rep && e - s >= 2 * spw) {
        uint64_t a = 0, b = (e - s) / spw - 1;               // words (at s + t * spw) known all-run / known not
        if (get64u(hv.x, (uint32_t)(s + b * spw), bl) == rep) return;
        while (b - a > 1) {
            const uint64_t mid = a + ((b - a) >> 1);
            if (get64u(hv.x, (uint32_t)(s + mid * spw), bl) == rep) a = mid; else b = mid;
        }
        const uint64_t wb = get64u(hv.x, (uint32_t)(s + b * spw), bl);
        first = s + b * spw + (uint64_t)(clz64(wb ^ rep) >> bl);
        if (first >= e) return;
    }
    // a variant site in the first bases leaves the probe without a long match (3 % of the chunks at BASELINE's rates,
    // which then parse without a prediction at all): two more tries further in -- the hint is a prediction for the whole
    // chunk, wherever it was taken
    for (uint32_t probe = 0; probe < 3; ++probe) {
        const uint64_t at = first + 96u * probe;
        if (at >= e || stop < at + 64) return;
        bool open;
        Match mt = lookup(pv.ix, hv.x, at, (uint32_t)(stop - at), 64u, &open, true);
        if (mt.len >= kDiagMinLen) { hint[q] = (int64_t)mt.pos - (int64_t)at; return; }
    }
}

// D: mismatch offsets of chunk g along its hinted diagonal (scalar form; the kernel k_diag_w does the same with one
// warp per chunk and coalesced loads).  mm = [count, valid_until, offsets...]: every position below valid_until is
// classified (mismatch iff listed); a reference that ends inside the chunk counts as a mismatch there.
DRLZ_HD void diag_scan_chunk(const ParseView& pv, uint32_t g) {
    uint16_t* mm = pv.mm + (size_t)g * kMmStride;
    mm[0] = 0; mm[1] = 0;
    int64_t dg = pv.hint[g / kHintGroup];
    if (dg == kNoDiag) return;
    uint32_t h = pv.chunk_hap[g];
    const HapView& hv = pv.haps[h];
    uint64_t s = (uint64_t)(g - pv.hap_chunk_base[h]) << pv.chunk_shift;
    if (s >= hv.m) return;
    uint64_t e = s + (1ull << pv.chunk_shift); if (e > hv.m) e = hv.m;
    int64_t p0 = (int64_t)s + dg;
    if (p0 < 0 || (uint64_t)p0 >= pv.ix.n) return;
    uint32_t p = (uint32_t)p0, len = (uint32_t)(e - s);
    uint32_t avail = (uint32_t)pv.ix.n - p;
    uint32_t cmplen = len < avail ? len : avail;
    uint32_t cnt = 0, V = len;
    bool full = false;
    const uint32_t bl = pv.ix.bl, B = 1u << bl, spw = 64u >> bl;
    for (uint32_t o = 0; o < cmplen && !full; o += spw) {
        uint64_t xr = get64u(hv.x, (uint32_t)s + o, bl) ^ get64u(pv.ix.text, p + o, bl);
        uint32_t nb = cmplen - o < spw ? cmplen - o : spw;
        if (nb < spw) xr &= ~0ull << (64 - B * nb);
        while (xr) {
            uint32_t b = (uint32_t)(clz64(xr) >> bl);
            if (cnt == (uint32_t)kMmMax) { V = o + b; full = true; break; }
            mm[2 + cnt++] = (uint16_t)(o + b);
            xr &= ~(((1ull << B) - 1) << (64 - B - B * b));
        }
    }
    if (!full && cmplen < len) {
        if (cnt == (uint32_t)kMmMax) V = cmplen; else mm[2 + cnt++] = (uint16_t)cmplen;
    }
    mm[0] = (uint16_t)cnt; mm[1] = (uint16_t)V;
}

// one term of ext[]: symbols of chunk g up to its first mismatch along its hinted diagonal (its whole length if there is
// none), whether the count goes on into chunk g + 1 (same haplotype, same diagonal, nothing in between), and whether a
// count that ends here ends for a REASON the match cannot outlive -- a mismatch, the end of the haplotype.  If it ends
// because the next chunk was hinted another diagonal (or none), the match may well go on: the sum is then only a lower
// bound and carries kExtInexact, and nobody may take it for the length of a phrase.
DRLZ_HD uint32_t ext_term(const ParseView& pv, uint32_t g, bool* goes_on, bool* exact_end) {
    *goes_on = false; *exact_end = false;
    const int64_t dg = pv.hint[g / kHintGroup];
    const uint32_t h = pv.chunk_hap[g];
    const HapView& hv = pv.haps[h];
    const uint64_t s = (uint64_t)(g - pv.hap_chunk_base[h]) << pv.chunk_shift;
    if (dg == kNoDiag || s >= hv.m) return 0u;
    uint64_t e = s + (1ull << pv.chunk_shift); if (e > hv.m) e = hv.m;
    const uint16_t* mm = pv.mm + (size_t)g * kMmStride;
    const uint32_t len = (uint32_t)(e - s), cnt = mm[0], V = mm[1];
    if (cnt > 0) { *exact_end = true; return mm[2]; }
    if (V != len) return V;                                  // (not classified at all: V == 0)
    if (e == hv.m) { *exact_end = true; return len; }        // the haplotype ends here
    *goes_on = g + 1 < pv.n_chunks && pv.chunk_hap[g + 1] == h && pv.hint[(g + 1) / kHintGroup] == dg;
    return len;
}

// P1 for global chunk g: speculative greedy parse from the chunk's first base with chunk-capped lookups.  Only
// the last phrase of a chunk can be open (it reaches the chunk end); its total length -- hence the exit of the
// chunk -- is filled in by resolve_chunk() once the diagonal runs are ranked.  The first phrase doubles as
// stage A's "phrase at the chunk start" (first_pos / first_len / run_len).
DRLZ_HD void spec_parse_chunk(const ParseView& pv, uint32_t g, const uint16_t* mm_copy = nullptr) {
    uint32_t h = pv.chunk_hap[g];
    const HapView hv = pv.haps[h];          // by value: the fields stay in registers instead of being re-read through L1
    uint64_t s = (uint64_t)(g - pv.hap_chunk_base[h]) << pv.chunk_shift;
    uint64_t e = s + (1ull << pv.chunk_shift); if (e > hv.m) e = hv.m;
    pv.link[g] = kEmpty;
    pv.open_i[g] = kEmpty;
    if (s >= hv.m) {
        pv.first_pos[g] = 0; pv.first_len[g] = 0; pv.run_len[g] = 0;
        pv.spec_cnt[g] = 0; pv.spec_exit[g] = (uint32_t)(s > 0xFFFFFFFEull ? 0xFFFFFFFEull : s);
        return;
    }
    uint64_t i = s; uint32_t cnt = 0;
    uint32_t cur = hv.n_exc ? exc_lower_bound(hv, i) : 0;
    bool open = false;
    const int64_t hint_g = pv.hint ? pv.hint[g / kHintGroup] : kNoDiag;
    int64_t diag = hint_g;
    const uint16_t* mml = !pv.mm ? nullptr : (mm_copy ? mm_copy : pv.mm + (size_t)g * kMmStride);
    while (i < e) {
        Match mt = phrase_capped_l(pv, hv, i, &cur, &open, &diag, mml, hint_g, s, g);
        if (i == s) { pv.first_pos[g] = mt.pos; pv.first_len[g] = mt.len; pv.run_len[g] = open ? kEmpty : mt.len; }
        if (cnt < pv.spec_store) pv.spec_ph[(size_t)g * pv.spec_store + cnt] = make_uint2(mt.pos, mt.len);
        ++cnt;
        if (open) { pv.open_i[g] = (uint32_t)i; pv.open_pos[g] = mt.pos; pv.open_len[g] = mt.len; break; }
        i += mt.len ? mt.len : 1;
    }
    pv.spec_cnt[g] = cnt;
    if (open) pv.spec_exit[g] = kEmpty;                       // resolved later
    else { pv.spec_exit[g] = (uint32_t)i; register_exit(pv, h, i, 1); }
}

// after link_chunk + list ranking: total length of the chunk's open last phrase, exit of the chunk
DRLZ_HD void resolve_chunk(const ParseView& pv, uint32_t g) {
    uint32_t oi = pv.open_i[g];
    if (oi == kEmpty) return;
    uint32_t h = pv.chunk_hap[g];
    const HapView& hv = pv.haps[h];
    uint64_t s = (uint64_t)(g - pv.hap_chunk_base[h]) << pv.chunk_shift;
    uint64_t i = oi;
    uint32_t total;
    if (i == s) total = pv.run_len[g];
    else {
        Match mt; mt.pos = pv.open_pos[g]; mt.len = pv.open_len[g];
        uint64_t j; bool same;
        uint32_t g2 = open_target(pv, h, i, mt, &j, &same);
        total = (uint32_t)(j - i) + (same ? pv.run_len[g2] : foreign_run(pv, hv, i, mt, j));
    }
    uint32_t cnt = pv.spec_cnt[g];
    if (cnt - 1 < pv.spec_store) pv.spec_ph[(size_t)g * pv.spec_store + cnt - 1].y = total;
    uint64_t x = i + total;
    pv.spec_exit[g] = (uint32_t)x;
    register_exit(pv, h, x, 1);
}

// F for alternative record (g, slot) scheduled for `round`
DRLZ_HD void fix_parse_record(const ParseView& pv, uint32_t g, int slot, uint32_t round) {
    uint32_t rid = g * kAltSlots + slot;
    if (pv.alt_entry[rid] == kEmpty || pv.alt_round[rid] != round) return;
    uint32_t h = pv.chunk_hap[g];
    const HapView& hv = pv.haps[h];
    uint64_t s = (uint64_t)(g - pv.hap_chunk_base[h]) << pv.chunk_shift;
    uint64_t e = s + (1ull << pv.chunk_shift); if (e > hv.m) e = hv.m;
    uint64_t iA = s, iB = pv.alt_entry[rid];
    uint32_t cntA = 0, cntB = 0;
    uint32_t curA = hv.n_exc ? exc_lower_bound(hv, iA) : 0;
    uint32_t curB = hv.n_exc ? exc_lower_bound(hv, iB) : 0;
    bool merged = false;
    int64_t diagA = kNoDiag, diagB = kNoDiag;
    if (pv.first_len[g] >= kDiagMinLen) diagA = diagB = (int64_t)pv.first_pos[g] - (int64_t)s;
    while (iB < e) {
        while (iA < iB) {
            uint32_t lenA;
            if (cntA < pv.spec_store && cntA < pv.spec_cnt[g]) lenA = pv.spec_ph[(size_t)g * pv.spec_store + cntA].y;
            else lenA = chunk_step(pv, g, h, hv, s, iA, &curA, &diagA).len;
            ++cntA;
            iA += lenA ? lenA : 1;
        }
        if (iA == iB) { merged = true; break; }
        Match mt = phrase_at(pv, h, hv, iB, &curB, &diagB, g, s);
        if (cntB < (uint32_t)kAltStore) pv.alt_ph[(size_t)rid * kAltStore + cntB] = make_uint2(mt.pos, mt.len);
        ++cntB;
        iB += mt.len ? mt.len : 1;
    }
    pv.alt_own[rid] = cntB;
    pv.alt_merge[rid] = merged ? cntA : kEmpty;
    if (merged) {
        pv.alt_cnt[rid] = cntB + (pv.spec_cnt[g] - cntA);
        pv.alt_exit[rid] = pv.spec_exit[g];                  // already registered by P1
    } else {
        pv.alt_cnt[rid] = cntB;
        pv.alt_exit[rid] = (uint32_t)iB;
        register_exit(pv, h, iB, round + 1);
    }
}

// J: node ids.  node = g*(kAltSlots+1) + 0 (speculative) | 1+slot (alternative)
DRLZ_HD uint32_t node_succ(const ParseView& pv, uint32_t node) {
    uint32_t g = node / (kAltSlots + 1);
    int r = (int)(node % (kAltSlots + 1));
    uint32_t h = pv.chunk_hap[g];
    const HapView& hv = pv.haps[h];
    uint64_t x;
    if (r == 0) x = pv.spec_exit[g];
    else {
        // unused slot, or a record whose fix-up round has not run (possible only while the host launches rounds blind
        // and checks convergence afterwards): no exit yet, a terminal
        if (pv.alt_entry[g * kAltSlots + r - 1] == kEmpty || pv.alt_round[g * kAltSlots + r - 1] > pv.parsed_rounds) return node;
        x = pv.alt_exit[g * kAltSlots + r - 1];
    }
    if (x >= hv.m) return node;                              // terminal: self loop
    uint32_t g2 = pv.hap_chunk_base[h] + (uint32_t)(x >> pv.chunk_shift);
    if ((x & ((1ull << pv.chunk_shift) - 1)) == 0) return g2 * (kAltSlots + 1);
    for (int s = 0; s < kAltSlots; ++s)
        if (pv.alt_entry[g2 * kAltSlots + s] == (uint32_t)x) return g2 * (kAltSlots + 1) + 1 + s;
    pv.counters[0] = 2;                                      // unresolved link (cannot happen after convergence)
    return node;
}

// E: after marking, pick the chunk's true record: returns phrase count, *entry = true entry position
// *slot = 0 for the speculative record, 1 + s for alternative record s.  m4 = the chunk's kAltSlots+1 mark bytes.
DRLZ_HD uint32_t chunk_true_record(const ParseView& pv, const uint8_t* m4, uint32_t g, uint32_t* entry, uint32_t* slot) {
    *slot = 0;
    if (m4[0]) {
        uint32_t h = pv.chunk_hap[g];
        *entry = (uint32_t)((uint64_t)(g - pv.hap_chunk_base[h]) << pv.chunk_shift);
        return pv.spec_cnt[g];
    }
    for (int s = 0; s < kAltSlots; ++s)
        if (m4[1 + s]) {
            // a record registered for a round that has not run yet (blind rounds, see parse_stage_count) has no phrases:
            // its counters are whatever an earlier batch left there, and this marking pass is provisional anyway
            const uint32_t rid = g * kAltSlots + s;
            *entry = pv.alt_entry[rid]; *slot = 1 + s;
            return pv.alt_round[rid] > pv.parsed_rounds ? 0u : pv.alt_cnt[rid];
        }
    *entry = 0;
    return 0;
}

DRLZ_HD void put_record(int64_t* out, uint32_t t, uint2 ph) {
    out[2 * (uint64_t)t] = (int64_t)ph.x;
    out[2 * (uint64_t)t + 1] = (int64_t)ph.y;
}

// E (fast path): copy the stored phrases of the true record; returns false if some were not stored
DRLZ_HD bool emit_stored(const ParseView& pv, uint32_t g, uint32_t slot, uint32_t cnt, int64_t* out) {
    const uint2* sp = pv.spec_ph + (size_t)g * pv.spec_store;
    if (slot == 0) {
        if (cnt > pv.spec_store) return false;
        for (uint32_t t = 0; t < cnt; ++t) put_record(out, t, sp[t]);
        return true;
    }
    uint32_t rid = g * kAltSlots + slot - 1;
    uint32_t own = pv.alt_own[rid], mg = pv.alt_merge[rid];
    if (own > (uint32_t)kAltStore) return false;
    if (mg != kEmpty && pv.spec_cnt[g] > pv.spec_store) return false;
    const uint2* ap = pv.alt_ph + (size_t)rid * kAltStore;
    for (uint32_t t = 0; t < own; ++t) put_record(out, t, ap[t]);
    if (mg != kEmpty) for (uint32_t t = mg; t < pv.spec_cnt[g]; ++t) put_record(out, own + t - mg, sp[t]);
    return true;
}

// E (slow path): write `cnt` phrases of chunk g starting at `entry` as (int64 pos, int64 len) records
DRLZ_HD void emit_chunk(const ParseView& pv, uint32_t g, uint32_t entry, uint32_t cnt, int64_t* out) {
    uint32_t h = pv.chunk_hap[g];
    const HapView& hv = pv.haps[h];
    uint64_t s = (uint64_t)(g - pv.hap_chunk_base[h]) << pv.chunk_shift;
    uint64_t i = entry;
    uint32_t cur = hv.n_exc ? exc_lower_bound(hv, i) : 0;
    int64_t diag = kNoDiag;
    for (uint32_t t = 0; t < cnt; ++t) {
        Match mt = chunk_step(pv, g, h, hv, s, i, &cur, &diag);
        out[2 * (uint64_t)t] = (int64_t)mt.pos;
        out[2 * (uint64_t)t + 1] = (int64_t)mt.len;
        i += mt.len ? mt.len : 1;
    }
}

}  // namespace drlz
