/*
 * drlz_oracle.c -- CPU oracle for the distributed RLZ parse.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this file's shared object.  The product path (libdrlz.so) never does.
 *
 * PARITY UNPINNED: the reference (NGSeq/PanGenSpark) ships no tests, no golden vectors and
 * cannot be executed here (no JVM/Scala/Spark; `radixSA` is an un-vendored, un-pinned
 * third-party binary, compile.sh:11).  This file is a line-faithful restatement of the
 * reference *source*:
 *
 *   orc_binary_search   <- DistributedRLZ.scala:108-164  (binarySearch)
 *   orc_factor          <- DistributedRLZ.scala:176-217  (factor)
 *   orc_encode_literal  <- DistributedRLZ.scala:222-242  (encode)
 *   orc_strip_gaps      <- DistributedRLZ.scala:72,86    (replaceAll("-", ""))
 *   record layout       <- DistributedRLZ.scala:263-284  (int64 LE pos, int64 LE len;
 *                                                         literal => pos = char code, len = 0)
 *   orc_sa_build        <- stands in for `radixSA` (DistributedRLZ.scala:52,55): the suffix
 *                          array of a byte string is unique, so any correct builder is an
 *                          exact stand-in provided the file holds exactly the sequence bytes.
 *
 * Widenings w.r.t. the Scala source: Int -> int64 everywhere; the StringIndexOutOfBounds
 * crash at DistributedRLZ.scala:186 is reported through *would_throw and then treated as a
 * mismatch (the clean-spec behaviour).
 *
 * A second, independent formulation (orc_encode_a3: one lexicographic lower-bound search,
 * both SA neighbours, LCP walk to the leftmost interval index) is kept here to cross-check
 * the restatement.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t i64;
typedef uint64_t u64;
typedef uint8_t u8;

/* ------------------------------------------------------------------------------------ */
/* Suffix array construction: packed-key radix sort + prefix doubling (Larsson-Sadakane  */
/* flavour).  Order: unsigned bytes, a proper prefix sorts before its extensions.        */
/* ------------------------------------------------------------------------------------ */

typedef struct { u64 key; i64 idx; } kv_t;

static int kv_cmp(const void* a, const void* b) {
    const kv_t* x = (const kv_t*)a; const kv_t* y = (const kv_t*)b;
    if (x->key < y->key) return -1;
    if (x->key > y->key) return 1;
    return 0;
}

static void radix_sort_kv(kv_t* a, kv_t* tmp, u64 n) {
    /* LSD, 16-bit digits, 4 passes; stable. */
    u64* cnt = (u64*)malloc(sizeof(u64) * 65536);
    for (int pass = 0; pass < 4; ++pass) {
        int sh = pass * 16;
        memset(cnt, 0, sizeof(u64) * 65536);
        for (u64 i = 0; i < n; ++i) cnt[(a[i].key >> sh) & 0xFFFF]++;
        int skip = 0;
        for (u64 dgt = 0; dgt < 65536; ++dgt) if (cnt[dgt] == n) { skip = 1; break; }
        if (skip) continue;
        u64 s = 0;
        for (u64 dgt = 0; dgt < 65536; ++dgt) { u64 c = cnt[dgt]; cnt[dgt] = s; s += c; }
        for (u64 i = 0; i < n; ++i) tmp[cnt[(a[i].key >> sh) & 0xFFFF]++] = a[i];
        memcpy(a, tmp, sizeof(kv_t) * n);
    }
    free(cnt);
}

int orc_sa_build(const u8* d, u64 n, i64* sa) {
    if (n == 0) return 0;
    /* dense order-preserving codes, 0 reserved for "past the end" (acts as a unique sentinel) */
    int present[256] = {0}; int code[256]; int sigma = 0;
    for (u64 i = 0; i < n; ++i) present[d[i]] = 1;
    for (int b = 0; b < 256; ++b) code[b] = present[b] ? ++sigma : 0;
    int bits = 1; while ((1 << bits) < sigma + 1) ++bits;
    int q = 64 / bits;                       /* symbols per key */
    kv_t* a = (kv_t*)malloc(sizeof(kv_t) * n);
    kv_t* tmp = (kv_t*)malloc(sizeof(kv_t) * n);
    i64* rank = (i64*)malloc(sizeof(i64) * n);
    if (!a || !tmp || !rank) { free(a); free(tmp); free(rank); return -1; }
    {
        u64 key = 0; int top = bits * (q - 1);
        u64 keep = (bits * q == 64) ? ~0ull : ((1ull << (bits * q)) - 1);
        for (u64 ii = n; ii-- > 0;) {
            key = ((key >> bits) | ((u64)code[d[ii]] << top)) & keep;
            a[ii].key = key; a[ii].idx = (i64)ii;
        }
    }
    radix_sort_kv(a, tmp, n);
    /* rank = index of group head */
    u64 nactive_groups = 0;
    {
        u64 head = 0;
        for (u64 j = 0; j < n; ++j) {
            if (j > 0 && a[j].key != a[j - 1].key) head = j;
            rank[a[j].idx] = (i64)head;
        }
    }
    for (u64 j = 0; j < n; ++j) sa[j] = a[j].idx;
    /* active groups (start,end) of size > 1 */
    u64 cap = 1024, ng = 0;
    u64* gs = (u64*)malloc(sizeof(u64) * cap * 2);
    for (u64 j = 0; j < n;) {
        u64 e = j + 1;
        while (e < n && a[e].key == a[j].key) ++e;
        if (e - j > 1) {
            if (ng == cap) { cap *= 2; gs = (u64*)realloc(gs, sizeof(u64) * cap * 2); }
            gs[2 * ng] = j; gs[2 * ng + 1] = e; ++ng;
        }
        j = e;
    }
    (void)nactive_groups;
    u64 h = (u64)q;
    while (ng > 0) {
        /* pass 1: secondary keys from the ranks of the previous round (no in-round updates) */
        for (u64 g = 0; g < ng; ++g)
            for (u64 j = gs[2 * g]; j < gs[2 * g + 1]; ++j) {
                u64 p = (u64)sa[j] + h;
                a[j].key = p < n ? (u64)rank[p] + 1 : 0;
                a[j].idx = sa[j];
            }
        /* pass 2: sort every group by secondary key, split */
        u64 cap2 = 1024, ng2 = 0;
        u64* gs2 = (u64*)malloc(sizeof(u64) * cap2 * 2);
        for (u64 g = 0; g < ng; ++g) {
            u64 s = gs[2 * g], e = gs[2 * g + 1];
            if (e - s > 4096) radix_sort_kv(a + s, tmp, e - s);
            else qsort(a + s, e - s, sizeof(kv_t), kv_cmp);
            for (u64 j = s; j < e;) {
                u64 e2 = j + 1;
                while (e2 < e && a[e2].key == a[j].key) ++e2;
                for (u64 t = j; t < e2; ++t) { sa[t] = a[t].idx; }
                if (e2 - j > 1) {
                    if (ng2 == cap2) { cap2 *= 2; gs2 = (u64*)realloc(gs2, sizeof(u64) * cap2 * 2); }
                    gs2[2 * ng2] = j; gs2[2 * ng2 + 1] = e2; ++ng2;
                }
                j = e2;
            }
        }
        /* pass 3: rank updates (after all keys of this round were consumed) */
        for (u64 g = 0; g < ng; ++g) {
            u64 s = gs[2 * g], e = gs[2 * g + 1];
            u64 head = s;
            for (u64 j = s; j < e; ++j) {
                if (j > s && a[j].key != a[j - 1].key) head = j;
                rank[sa[j]] = (i64)head;
            }
        }
        free(gs); gs = gs2; ng = ng2; h *= 2;
    }
    free(gs); free(a); free(tmp); free(rank);
    return 0;
}

/* Kasai LCP: lcp[k] = lcp(d[sa[k-1]..], d[sa[k]..]) for k >= 1, lcp[0] = 0. */
int orc_lcp_build(const u8* d, u64 n, const i64* sa, i64* lcp) {
    if (n == 0) return 0;
    i64* isa = (i64*)malloc(sizeof(i64) * n);
    if (!isa) return -1;
    for (u64 k = 0; k < n; ++k) isa[sa[k]] = (i64)k;
    u64 hh = 0;
    for (u64 i = 0; i < n; ++i) {
        i64 k = isa[i];
        if (k == 0) { lcp[0] = 0; hh = 0; continue; }
        u64 j = (u64)sa[k - 1];
        while (i + hh < n && j + hh < n && d[i + hh] == d[j + hh]) ++hh;
        lcp[k] = (i64)hh;
        if (hh > 0) --hh;
    }
    free(isa);
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* A1: literal restatement of DistributedRLZ.scala                                        */
/* ------------------------------------------------------------------------------------ */

/* DistributedRLZ.scala:108-164.  Returns 0 for None, 1 for Some((*olb,*orb)). */
static int orc_binary_search(i64 lb, i64 rb, const u8* d, i64 n, u8 cur, i64 i, const i64* SA,
                             i64* olb, i64* orb) {
    i64 low = lb, high = rb;
    while (low < high) {                                   /* :112-130 */
        i64 mid = low + ((high - low) / 2);
        i64 midKey = SA[mid] + i;
        u8 midValue = midKey < n ? d[midKey] : (u8)'1';    /* :118-122 */
        if (cur <= midValue) high = mid; else low = mid + 1;
    }
    i64 low_res = low;
    if (SA[low_res] + i >= n || d[SA[low_res] + i] != cur) return 0;   /* :136-138 */
    high = rb;
    while (low < high) {                                   /* :140-155 */
        i64 mid = low + ((high - low) / 2);
        i64 midKey = SA[mid] + i;
        u8 midValue = midKey < n ? d[midKey] : (u8)'1';
        if (cur >= midValue) low = mid + 1; else high = mid;
    }
    if (SA[low] != n - 1 && SA[low] + i < n && d[SA[low] + i] != cur) {  /* :157 */
        *olb = low_res; *orb = low - 1; return 1;
    }
    *olb = low_res; *orb = low; return 1;
}

/* DistributedRLZ.scala:176-217.  Literal => *pos = byte value, *len = 0. */
static void orc_factor(i64 i, const u8* x, i64 m, const u8* d, i64 n, const i64* SA,
                       i64* pos, i64* len, u64* would_throw) {
    i64 lb = 0, rb = n - 1, j = i;
    while (j < m) {
        if (lb == rb) {                                    /* :186 */
            i64 p = SA[lb] + j - i;
            if (p >= n) { if (would_throw) ++*would_throw; break; }   /* Scala throws here */
            if (d[p] != x[j]) break;
        }
        i64 nlb, nrb;
        if (!orc_binary_search(lb, rb, d, n, x[j], j - i, SA, &nlb, &nrb)) break;   /* :190-194 */
        lb = nlb; rb = nrb;
        ++j;
        if (j == m) break;                                 /* :204-206 */
    }
    if (j == i) { *pos = (i64)x[j]; *len = 0; }            /* :210-211 */
    else { *pos = SA[lb]; *len = j - i; }                  /* :215 */
}

/* DistributedRLZ.scala:222-242.  pairs = (pos,len) int64 records, returns count (or needed
 * count if cap is too small; nothing is written past cap). */
u64 orc_encode_literal(const u8* d, u64 n, const i64* SA, const u8* x, u64 m,
                       i64* pairs, u64 cap, u64* would_throw) {
    u64 np = 0; i64 i = 0;
    if (would_throw) *would_throw = 0;
    if (n == 0) {   /* the Scala code would index SA(…) of an empty array; define: all literals */
        for (u64 k = 0; k < m; ++k) { if (np < cap) { pairs[2 * np] = x[k]; pairs[2 * np + 1] = 0; } ++np; }
        return np;
    }
    while (i < (i64)m) {
        i64 pos, len;
        orc_factor(i, x, (i64)m, d, (i64)n, SA, &pos, &len, would_throw);
        if (np < cap) { pairs[2 * np] = pos; pairs[2 * np + 1] = len; }
        ++np;
        i += len == 0 ? 1 : len;
    }
    return np;
}

/* ------------------------------------------------------------------------------------ */
/* A3: position-independent formulation                           */
/* ------------------------------------------------------------------------------------ */

/* compare pattern P=x[i..i+plen) with suffix d[p..]; returns lcp, *less = suffix < P */
static u64 suf_cmp(const u8* d, u64 n, u64 p, const u8* P, u64 plen, int* less) {
    u64 l = 0, lim = n - p < plen ? n - p : plen;
    while (l < lim && d[p + l] == P[l]) ++l;
    if (l == plen) *less = 0;
    else if (l == n - p) *less = 1;
    else *less = d[p + l] < P[l];
    return l;
}

void orc_ms_at(const u8* d, u64 n, const i64* SA, const i64* LCP, const u8* x, u64 m, u64 i,
               i64* ms, i64* pos) {
    const u8* P = x + i; u64 plen = m - i;
    u64 lo = 0, hi = n;                       /* r = #suffixes < P */
    while (lo < hi) {
        u64 mid = lo + (hi - lo) / 2; int less;
        suf_cmp(d, n, (u64)SA[mid], P, plen, &less);
        if (less) lo = mid + 1; else hi = mid;
    }
    u64 r = lo; int dummy;
    u64 l1 = r > 0 ? suf_cmp(d, n, (u64)SA[r - 1], P, plen, &dummy) : 0;
    u64 l2 = r < n ? suf_cmp(d, n, (u64)SA[r], P, plen, &dummy) : 0;
    if (l2 > l1) { *ms = (i64)l2; *pos = SA[r]; }
    else if (l1 > 0) {
        u64 k = r - 1;
        while (k > 0 && (u64)LCP[k] >= l1) --k;
        *ms = (i64)l1; *pos = SA[k];
    } else { *ms = 0; *pos = (i64)x[i]; }
}

u64 orc_encode_a3(const u8* d, u64 n, const i64* SA, const i64* LCP, const u8* x, u64 m,
                  i64* pairs, u64 cap) {
    u64 np = 0, i = 0;
    while (i < m) {
        i64 ms, pos;
        if (n == 0) { ms = 0; pos = x[i]; } else orc_ms_at(d, n, SA, LCP, x, m, i, &ms, &pos);
        if (np < cap) { pairs[2 * np] = pos; pairs[2 * np + 1] = ms; }
        ++np;
        i += ms == 0 ? 1 : (u64)ms;
    }
    return np;
}

/* dense matching statistics for every position (used to check the dense GPU kernel) */
void orc_ms_all(const u8* d, u64 n, const i64* SA, const i64* LCP, const u8* x, u64 m,
                i64* ms, i64* pos) {
    #pragma omp parallel for schedule(dynamic, 1024)
    for (i64 i = 0; i < (i64)m; ++i) orc_ms_at(d, n, SA, LCP, x, m, (u64)i, &ms[i], &pos[i]);
}

/* ------------------------------------------------------------------------------------ */
/* helpers: gap strip, decode, batch driver                                               */
/* ------------------------------------------------------------------------------------ */

/* DistributedRLZ.scala:72,86: remove every '-'.  Also drops one trailing '\n' / "\r\n" the way
 * Spark's text reader strips the line terminator. */
u64 orc_strip_gaps(const u8* row, u64 cols, u8* out) {
    while (cols > 0 && (row[cols - 1] == '\n' || row[cols - 1] == '\r')) --cols;
    u64 m = 0;
    for (u64 c = 0; c < cols; ++c) if (row[c] != '-') out[m++] = row[c];
    return m;
}

/* Rebuild the haplotype from (ref, phrases): independent validity check. returns length or -1 */
i64 orc_decode(const u8* d, u64 n, const i64* pairs, u64 np, u8* out, u64 cap) {
    u64 m = 0;
    for (u64 k = 0; k < np; ++k) {
        i64 pos = pairs[2 * k], len = pairs[2 * k + 1];
        if (len == 0) { if (m >= cap) return -1; out[m++] = (u8)pos; }
        else {
            if (pos < 0 || (u64)pos + (u64)len > n || m + (u64)len > cap) return -1;
            memcpy(out + m, d + pos, (size_t)len); m += (u64)len;
        }
    }
    return (i64)m;
}

static double now_s(void) {
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* One haplotype per task over nthreads threads: the shape of spark-submit --master local[N]
 * (DistributedRLZ.scala:247-248).  Timed region = gap strip + encode + 16-byte records.
 * pairs_out[h] is malloc'd here (free with orc_free). */
int orc_encode_rows_mt(const u8* d, u64 n, const i64* SA, const u8* const* rows, const u64* cols,
                       u64 H, int nthreads, i64** pairs_out, u64* npairs, u64* m_out,
                       u64* would_throw, double* seconds) {
    int err = 0;
    double t0 = now_s();
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    #pragma omp parallel for schedule(dynamic, 1)
    for (i64 h = 0; h < (i64)H; ++h) {
        u8* x = (u8*)malloc(cols[h] + 1);
        if (!x) { err = -1; continue; }
        u64 m = orc_strip_gaps(rows[h], cols[h], x);
        u64 cap = m / 64 + 1024; u64 wt = 0;
        i64* pr = (i64*)malloc(sizeof(i64) * 2 * cap);
        u64 np = pr ? orc_encode_literal(d, n, SA, x, m, pr, cap, &wt) : 0;
        if (pr && np > cap) {
            free(pr); cap = np; pr = (i64*)malloc(sizeof(i64) * 2 * cap);
            if (pr) np = orc_encode_literal(d, n, SA, x, m, pr, cap, &wt);
        }
        if (!pr) err = -1;
        pairs_out[h] = pr; npairs[h] = np; m_out[h] = m;
        if (would_throw) would_throw[h] = wt;
        free(x);
    }
    if (seconds) *seconds = now_s() - t0;
    return err;
}

void orc_free(void* p) { free(p); }

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
