/*
 * synth.c -- deterministic synthetic pan-genome generator.
 * TEST / BENCH INFRASTRUCTURE ONLY (input data, not the product).
 *
 * Counter based, so any slice can be generated independently:
 *     u(seed, k) = mix64(seed + (k + 1) * GAMMA)          (k-th output of a splitmix64 stream)
 * Reference   d[k]  = "ACGT"[u(S_ref, k) >> 62]
 * Insertion columns are global (shared by all rows, as in a gapped MSA, README.md:89-90):
 *     site after ref position k (k < n-1) iff u(S_ins, k) < p_ins * 2^64, length 1 + (u(S_len,k) & 7),
 *     inserted base t = "ACGT"[u(S_insb, 8k + t) >> 62]; haplotype h carries it iff u(S_h^2, k) & 1.
 *     No site after k = n-1 (keeps inputs away from the reference's crash, quirk Q2).
 * Haplotype h >= 1 (seed S_h = S_ref ^ (h+1) * 0xA24BAED4963EE407):
 *     SNP at k iff u(S_h, k) < p_snp * 2^64, alt = (ref + 1 + (u(S_h^5, k) >> 33) % 3) & 3
 *     deletion starts at k iff u(S_h^1, k) < p_del * 2^64, length 1 + (u(S_h^3, k) & 7), written '-'
 * Row 0 is the reference itself with '-' in every insertion column (README.md:90).
 * All rows have the same number of columns M.
 * Optional repeat overlay on the reference (genome-like LCP tails): `repeats` segments,
 * each a copy of an earlier segment (len 300..6000) with 1 % substitutions.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;
typedef uint8_t u8;

#define GAMMA 0x9E3779B97F4A7C15ull

static inline u64 mix64(u64 z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline u64 uu(u64 seed, u64 k) { return mix64(seed + (k + 1) * GAMMA); }

u64 syn_u(u64 seed, u64 k) { return uu(seed, k); }

static const char ACGT[4] = {'A', 'C', 'G', 'T'};

static inline u64 thresh(double p) {
    if (p <= 0) return 0;
    if (p >= 1) return ~0ull;
    return (u64)(p * 18446744073709551616.0);
}

void syn_reference(u64 seed_ref, u64 n, u64 repeats, u8* d) {
    #pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < (int64_t)n; ++k) d[k] = (u8)ACGT[uu(seed_ref, (u64)k) >> 62];
    /* sequential overlay: segment r copies [src, src+len) to dst with 1 % substitutions */
    u64 sr = seed_ref ^ 0x5EEDBEEF00C0FFEEull;
    for (u64 r = 0; r < repeats && n > 16000; ++r) {
        u64 len = 300 + uu(sr, 4 * r) % 5701;
        u64 src = uu(sr, 4 * r + 1) % (n - len);
        u64 dst = uu(sr, 4 * r + 2) % (n - len);
        for (u64 t = 0; t < len; ++t) {
            u8 c = d[src + t];
            if (uu(sr ^ (r + 1) * 0xD6E8FEB86659FD93ull, t) < thresh(0.01)) {
                int code = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3;
                c = (u8)ACGT[(code + 1 + (int)(uu(sr ^ r, t) % 3)) & 3];
            }
            d[dst + t] = c;
        }
    }
}

static inline u64 S_INS(u64 s) { return s ^ 0x1111111111111111ull; }
static inline u64 S_LEN(u64 s) { return s ^ 0x2222222222222222ull; }
static inline u64 S_INSB(u64 s) { return s ^ 0x3333333333333333ull; }

/* number of MSA columns */
u64 syn_columns(u64 seed_ref, u64 n, double p_ins) {
    u64 M = n, th = thresh(p_ins);
    if (n < 2 || th == 0) return M;
    #pragma omp parallel for schedule(static) reduction(+:M)
    for (int64_t k = 0; k < (int64_t)n - 1; ++k)
        if (uu(S_INS(seed_ref), (u64)k) < th) M += 1 + (uu(S_LEN(seed_ref), (u64)k) & 7);
    return M;
}

/* Write row h (M columns, no terminator).  d = reference bytes.  Returns columns written. */
u64 syn_row(u64 seed_ref, const u8* d, u64 n, u64 h, double p_snp, double p_ins, double p_del, u8* row) {
    u64 sh = seed_ref ^ ((h + 1) * 0xA24BAED4963EE407ull);
    u64 t_snp = thresh(p_snp), t_ins = thresh(p_ins), t_del = thresh(p_del);
    u64 c = 0, del_until = 0;
    for (u64 k = 0; k < n; ++k) {
        u8 base = d[k];
        if (h > 0) {
            if (t_del && uu(sh ^ 1, k) < t_del) {
                u64 e = k + 1 + (uu(sh ^ 3, k) & 7);
                if (e > del_until) del_until = e;
            }
            if (k < del_until) base = '-';
            else if (t_snp && uu(sh, k) < t_snp) {
                int code = base == 'A' ? 0 : base == 'C' ? 1 : base == 'G' ? 2 : 3;
                base = (u8)ACGT[(code + 1 + (int)((uu(sh ^ 5, k) >> 33) % 3)) & 3];
            }
        }
        row[c++] = base;
        if (k + 1 < n && t_ins && uu(S_INS(seed_ref), k) < t_ins) {
            u64 len = 1 + (uu(S_LEN(seed_ref), k) & 7);
            int carry = h > 0 && (uu(sh ^ 2, k) & 1);
            for (u64 t = 0; t < len; ++t)
                row[c++] = carry ? (u8)ACGT[uu(S_INSB(seed_ref), 8 * k + t) >> 62] : (u8)'-';
        }
    }
    return c;
}

/* rows h0..h0+H-1 into one buffer with stride `stride` bytes (>= M), in parallel */
void syn_rows(u64 seed_ref, const u8* d, u64 n, u64 h0, u64 H, double p_snp, double p_ins,
              double p_del, u8* rows, u64 stride) {
    #pragma omp parallel for schedule(dynamic, 1)
    for (int64_t h = 0; h < (int64_t)H; ++h)
        syn_row(seed_ref, d, n, h0 + (u64)h, p_snp, p_ins, p_del, rows + (u64)h * stride);
}
