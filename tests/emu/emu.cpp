// emu.cpp -- TEST-ONLY host harness that runs the statements of pangenspark_b200/csrc/rlz_core.cuh
// on the CPU (sequentially, one "thread" after another) so the chunk-parallel parse logic can be checked
// against the oracle without a GPU.  It is never linked into libdrlz.so and nothing in the product calls it.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include "../../pangenspark_b200/csrc/rlz_core.cuh"

using namespace drlz;

struct Alphabet {
    uint32_t bl, sigma;
    uint8_t b2c[256], c2b[256];
};

// same rule as the library: pure ACGT -> fixed 2-bit codes; else order-preserving dense codes, 4 bits if sigma <= 16
static Alphabet make_alphabet(const uint8_t* ref, uint64_t n) {
    Alphabet a; memset(a.b2c, 0xFF, 256); memset(a.c2b, 0, 256);
    bool present[256] = {false};
    for (uint64_t i = 0; i < n; ++i) present[ref[i]] = true;
    bool acgt = true;
    for (int b = 0; b < 256; ++b) if (present[b] && b != 'A' && b != 'C' && b != 'G' && b != 'T') acgt = false;
    if (acgt) {
        a.bl = 1; a.sigma = 4;
        const char* s = "ACGT";
        for (int c = 0; c < 4; ++c) { a.b2c[(uint8_t)s[c]] = (uint8_t)c; a.c2b[c] = (uint8_t)s[c]; }
        return a;
    }
    uint32_t sg = 0;
    for (int b = 0; b < 256; ++b) if (present[b]) { a.b2c[b] = (uint8_t)sg; a.c2b[sg] = (uint8_t)b; ++sg; }
    a.sigma = sg; a.bl = sg <= 16 ? 2 : 3;
    return a;
}

static void pack_seq(const Alphabet& al, const uint8_t* s, uint64_t m, std::vector<uint64_t>& w, std::vector<uint32_t>* ep,
                     std::vector<uint8_t>* eb) {
    const uint32_t B = 1u << al.bl, spw = 64 / B;
    w.assign(m / spw + 8, 0);
    for (uint64_t i = 0; i < m; ++i) {
        uint32_t c = al.b2c[s[i]];
        if (c == 0xFF) { if (ep) { ep->push_back((uint32_t)i); eb->push_back(s[i]); } c = 0; }
        w[i / spw] |= (uint64_t)c << (64 - B - B * (i % spw));
    }
}

static void build_tab(const IndexView& ix, const std::vector<uint32_t>& sa, std::vector<uint32_t>& tab) {
    uint64_t nk = 1;
    for (int t = 0; t < ix.k; ++t) nk *= ix.sigma;
    tab.assign(nk + 1, (uint32_t)ix.n);
    uint64_t prev = 0;  // next K to fill
    for (uint64_t j = 0; j < ix.n; ++j) {
        uint64_t key, span;
        prefix_keys(ix, get64(ix.text, sa[j], ix.bl), (uint32_t)ix.k, &key, &span);   // symbols past n are code 0
        for (uint64_t K = prev; K <= key; ++K) tab[K] = (uint32_t)j;
        if (key + 1 > prev) prev = key + 1;
    }
}

extern "C" {

void emu_free(void* p) { free(p); }
uint64_t emu_long_hits() { return drlz::g_emu_long_hits; }

// use_uniq: 0 = no uniqueness array, 1 = with it, 2 = also the long-repeat machinery (hint fill, extents across chunks)
// returns 0 ok, 1 reference not ACGT, 2 overflow/unresolved
int emu_parse(const uint8_t* ref, uint64_t n, const int64_t* sa64, int k, uint32_t chunk, uint32_t min_cap, uint32_t foreign_cap, int use_uniq, int H,
              const uint8_t* const* rows, const uint64_t* cols, int64_t** pairs_out, uint64_t* npairs,
              uint64_t* m_out, int* rounds_out) {
    Alphabet al = make_alphabet(ref, n);
    const uint32_t spw = 64u >> al.bl;
    while (k > 1 && ((uint32_t)k > spw || pow((double)al.sigma, k) > 3e7)) --k;
    uint32_t chunk_shift = 0;
    while ((2u << chunk_shift) <= chunk && chunk_shift < 31) ++chunk_shift;     // round down to a power of two
    chunk = 1u << chunk_shift;
    std::vector<uint64_t> text; pack_seq(al, ref, n, text, nullptr, nullptr);
    std::vector<uint32_t> sa(n); for (uint64_t i = 0; i < n; ++i) sa[i] = (uint32_t)sa64[i];
    // uniqueness array: longest repeat starting at p (Kasai LCP on the host; test harness only)
    std::vector<uint16_t> uniq(n + 1, 0);
    std::vector<uint32_t> uniq32(n + 1, 0);
    {
        std::vector<uint32_t> isa(n), lcp(n + 1, 0);
        for (uint64_t j = 0; j < n; ++j) isa[sa[j]] = (uint32_t)j;
        uint64_t hh = 0;
        for (uint64_t p = 0; p < n; ++p) {
            uint32_t r = isa[p];
            if (r == 0) { hh = 0; continue; }
            uint64_t q = sa[r - 1];
            while (p + hh < n && q + hh < n && ref[p + hh] == ref[q + hh]) ++hh;
            lcp[r] = (uint32_t)hh;
            if (hh) --hh;
        }
        for (uint64_t p = 0; p < n; ++p) {
            uint32_t r = isa[p];
            uint32_t u = lcp[r] > lcp[r + 1] ? lcp[r] : lcp[r + 1];
            uniq[p] = (uint16_t)(u > 65535u ? 65535u : u);
            uniq32[p] = u;
        }
    }
    std::vector<uint16_t> uniqc((n >> kUniqBlockShift) + 2, 0);
    if (use_uniq) for (uint64_t p = 0; p < n; ++p) if (uniq[p] > uniqc[p >> kUniqBlockShift]) uniqc[p >> kUniqBlockShift] = uniq[p];
    IndexView ix{text.data(), sa.data(), nullptr, n, k, use_uniq ? uniq.data() : nullptr, al.bl, al.sigma, al.c2b,
                 use_uniq ? uniq32.data() : nullptr, use_uniq ? uniqc.data() : nullptr};
    std::vector<uint32_t> tab; build_tab(ix, sa, tab);
    ix.tab = tab.data();

    std::vector<std::vector<uint64_t>> xs(H);
    std::vector<std::vector<uint32_t>> eps(H);
    std::vector<std::vector<uint8_t>> ebs(H);
    std::vector<HapView> haps(H);
    std::vector<uint32_t> base(H + 1, 0);
    for (int h = 0; h < H; ++h) {
        std::vector<uint8_t> g; g.reserve(cols[h]);
        uint64_t c = cols[h];
        while (c > 0 && (rows[h][c - 1] == '\n' || rows[h][c - 1] == '\r')) --c;
        for (uint64_t t = 0; t < c; ++t) if (rows[h][t] != '-') g.push_back(rows[h][t]);
        pack_seq(al, g.data(), g.size(), xs[h], &eps[h], &ebs[h]);
        haps[h] = HapView{xs[h].data(), (uint64_t)g.size(), eps[h].data(), ebs[h].data(), (uint32_t)eps[h].size()};
        m_out[h] = g.size();
        base[h + 1] = base[h] + (uint32_t)((g.size() + chunk - 1) / chunk);
    }
    uint32_t NC = base[H];
    std::vector<uint32_t> chunk_hap(NC);
    for (int h = 0; h < H; ++h) for (uint32_t g = base[h]; g < base[h + 1]; ++g) chunk_hap[g] = h;
    std::vector<uint32_t> first_pos(NC), first_len(NC), run_a(NC), run_b(NC), link_a(NC), link_b(NC), open_i(NC), open_pos(NC), open_len(NC);
    std::vector<uint32_t> spec_cnt(NC), spec_exit(NC), alt_entry((size_t)NC * kAltSlots, kEmpty),
        alt_cnt((size_t)NC * kAltSlots, 0), alt_exit((size_t)NC * kAltSlots, 0), alt_round((size_t)NC * kAltSlots, 0),
        counters(2 + kMaxRounds, 0), alt_own((size_t)NC * kAltSlots, 0), alt_merge((size_t)NC * kAltSlots, kEmpty);
    uint32_t spec_store = 4;   // small on purpose: exercises the re-parse path of the emission
    std::vector<uint2> spec_ph((size_t)NC * spec_store), alt_ph((size_t)NC * kAltSlots * kAltStore);
    std::vector<int64_t> hint(NC / kHintGroup + 1, kNoDiag);
    std::vector<uint16_t> mm((size_t)NC * kMmStride + 16, 0);
    std::vector<uint32_t> ext;
    ParseView pv{ix, haps.data(), chunk_hap.data(), base.data(), chunk, chunk_shift, NC, min_cap, foreign_cap,
                 first_pos.data(), first_len.data(), run_a.data(), link_a.data(), open_i.data(), open_pos.data(), open_len.data(), spec_cnt.data(), spec_exit.data(),
                 alt_entry.data(), alt_cnt.data(), alt_exit.data(), alt_round.data(), alt_own.data(), alt_merge.data(),
                 spec_ph.data(), spec_store, nullptr, nullptr, alt_ph.data(), counters.data(), (uint32_t)kMaxRounds};
    uint32_t maxc0 = 1;
    for (int h = 0; h < H; ++h) if (base[h + 1] - base[h] > maxc0) maxc0 = base[h + 1] - base[h];
    if (use_uniq && chunk_shift >= 5 && chunk_shift <= 15) {
        for (uint32_t q = 0; q * kHintGroup < NC; ++q) hint_group(pv, hint.data(), q);
        if (use_uniq > 1) {                                  // k_hint_fill: chunks without a hint take the one behind them (same haplotype) ...
            std::vector<int64_t> own(hint);
            for (uint32_t g = 0; g < NC; ++g) {
                if (own[g] != kNoDiag) continue;
                uint32_t f = g;
                while (f + 1 < NC && own[f] == kNoDiag) ++f;
                if (own[f] != kNoDiag && chunk_hap[f] == chunk_hap[g]) hint[g] = own[f];
            }
            own = hint;                                      // ... and what is left (the tail of a haplotype) the one in front
            for (uint32_t g = 0; g < NC; ++g) {
                if (own[g] != kNoDiag) continue;
                uint32_t f = g;
                while (f > 0 && own[f] == kNoDiag) --f;
                if (own[f] != kNoDiag && chunk_hap[f] == chunk_hap[g]) hint[g] = own[f];
            }
        }
        pv.hint = hint.data(); pv.mm = mm.data();
        for (uint32_t g = 0; g < NC; ++g) diag_scan_chunk(pv, g);
        if (use_uniq > 1) {                                  // ext[]: right-to-left segmented sum of ext_term
            ext.assign(NC + 1, 0);
            for (uint32_t g = NC; g-- > 0;) {
                bool on, exact;
                const uint32_t v = ext_term(pv, g, &on, &exact);
                const uint64_t sum = (uint64_t)v + (on ? (uint64_t)(ext[g + 1] & ~kExtInexact) : 0ull);
                const bool inexact = on ? (ext[g + 1] & kExtInexact) != 0 : !exact;
                ext[g] = sum >= 0x7FFFFFFFull ? (0x7FFFFFFFu | kExtInexact) : ((uint32_t)sum | (inexact ? kExtInexact : 0u));
            }
            pv.ext = ext.data();
        }
    }
    for (uint32_t g = 0; g < NC; ++g) spec_parse_chunk(pv, g);
    for (uint32_t g = 0; g < NC; ++g) link_chunk(pv, g);
    if (counters[0]) return 2;
    {
        uint32_t *li = link_a.data(), *ri = run_a.data(), *lo = link_b.data(), *ro = run_b.data();
        for (uint32_t span = 1; span < maxc0; span *= 2) {
            for (uint32_t g = 0; g < NC; ++g) rank_round(li, ri, lo, ro, g);
            std::swap(li, lo); std::swap(ri, ro);
        }
        pv.run_len = ri; pv.link = li;
    }
    for (uint32_t g = 0; g < NC; ++g) resolve_chunk(pv, g);
    if (getenv("DRLZ_EMU_DEBUG"))
        for (uint32_t g = 0; g < NC; ++g)
            fprintf(stderr, "chunk %u s=%u hint=%lld mm=[%u,%u,%u] ext=%u first=(%u,%u) run=%u spec_cnt=%u exit=%u open_i=%u\n", g, g * chunk, (long long)hint[g / kHintGroup],
                    mm[(size_t)g * kMmStride], mm[(size_t)g * kMmStride + 1], mm[(size_t)g * kMmStride + 2], ext.empty() ? 0u : (ext[g] & 0x7FFFFFFFu), first_pos[g], first_len[g], pv.run_len[g], spec_cnt[g], spec_exit[g], open_i[g]);
    if (counters[0]) return 2;
    {   // A provisional marking pass before any fix-up round has run (what the library does when it launches fewer
        // rounds blind than the input needs): records registered for round 1 are not parsed yet, their counters are
        // whatever an earlier batch left behind -- poisoned here.  Such a record must count zero phrases; a nonzero
        // count can only be a speculative record's.
        ParseView pq = pv; pq.parsed_rounds = 0;
        std::fill(alt_cnt.begin(), alt_cnt.end(), 0xDEADBEEFu);
        std::fill(alt_own.begin(), alt_own.end(), 0xDEADBEEFu);
        std::fill(alt_merge.begin(), alt_merge.end(), 0x00ADBEEFu);
        uint32_t NNp = NC * (kAltSlots + 1);
        std::vector<uint32_t> jp(NNp), jp2(NNp);
        std::vector<uint8_t> mk(NNp, 0);
        const uint32_t c0 = counters[0];
        for (uint32_t v = 0; v < NNp; ++v) jp[v] = node_succ(pq, v);
        counters[0] = c0;                  // links into records that do not exist yet are expected in this pass
        uint32_t mc = 1;
        for (int h = 0; h < H; ++h) { if (base[h + 1] > base[h]) mk[base[h] * (kAltSlots + 1)] = 1; if (base[h + 1] - base[h] > mc) mc = base[h + 1] - base[h]; }
        for (uint32_t span = 1; span < 2 * mc; span *= 2) {
            for (uint32_t v = 0; v < NNp; ++v) { uint32_t j = jp[v]; if (mk[v]) mk[j] = 1; jp2[v] = jp[j]; }
            jp.swap(jp2);
        }
        for (uint32_t g = 0; g < NC; ++g) {
            uint32_t e, sl;
            uint32_t c = chunk_true_record(pq, mk.data() + (size_t)g * (kAltSlots + 1), g, &e, &sl);
            if (c != 0 && !(sl == 0 && c == spec_cnt[g])) return 3;
        }
        std::fill(alt_cnt.begin(), alt_cnt.end(), 0u);
        std::fill(alt_own.begin(), alt_own.end(), 0u);
        std::fill(alt_merge.begin(), alt_merge.end(), kEmpty);
    }
    int round = 1;
    while (counters[1 + round] > 0 && round < kMaxRounds) {
        for (uint32_t g = 0; g < NC; ++g) for (int s = 0; s < kAltSlots; ++s) fix_parse_record(pv, g, s, round);
        ++round;
    }
    *rounds_out = round;
    if (counters[0] || round >= kMaxRounds) return 2;
    uint32_t NN = NC * (kAltSlots + 1);
    std::vector<uint32_t> jump(NN), jump2(NN);
    std::vector<uint8_t> mark(NN, 0);
    for (uint32_t v = 0; v < NN; ++v) jump[v] = node_succ(pv, v);
    if (counters[0]) return 2;
    uint32_t maxc = 1;
    for (int h = 0; h < H; ++h) { if (base[h + 1] > base[h]) mark[base[h] * (kAltSlots + 1)] = 1; if (base[h+1]-base[h] > maxc) maxc = base[h+1]-base[h]; }
    for (uint32_t span = 1; span < 2 * maxc; span *= 2) {
        for (uint32_t v = 0; v < NN; ++v) { uint32_t j = jump[v]; if (mark[v]) mark[j] = 1; jump2[v] = jump[j]; }
        jump.swap(jump2);
    }
    std::vector<uint64_t> off(NC + 1, 0);
    std::vector<uint32_t> ent(NC), cnt(NC), slot(NC);
    for (uint32_t g = 0; g < NC; ++g) { cnt[g] = chunk_true_record(pv, mark.data() + (size_t)g * (kAltSlots + 1), g, &ent[g], &slot[g]); off[g + 1] = off[g] + cnt[g]; }
    std::vector<int64_t> out(2 * off[NC] + 2);
    for (uint32_t g = 0; g < NC; ++g)
        if (cnt[g] && !emit_stored(pv, g, slot[g], cnt[g], out.data() + 2 * off[g])) emit_chunk(pv, g, ent[g], cnt[g], out.data() + 2 * off[g]);
    for (int h = 0; h < H; ++h) {
        uint64_t a = off[base[h]], b = off[base[h + 1]];
        npairs[h] = b - a;
        pairs_out[h] = (int64_t*)malloc(sizeof(int64_t) * 2 * (b - a) + 16);
        memcpy(pairs_out[h], out.data() + 2 * a, sizeof(int64_t) * 2 * (b - a));
    }
    return 0;
}

}  // extern "C"
