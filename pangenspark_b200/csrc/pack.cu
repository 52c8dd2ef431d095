// pack.cu -- gap strip + 2-bit pack + exception list for MSA rows (DistributedRLZ.scala:72,86).
//
// A row is a gapped MSA line; every '-' is removed (stream compaction) and the remaining bases are packed
// 2 bits each, 32 per 64-bit word, first base in the top bits.  Bytes outside {A,C,G,T} are stored as code 0
// and listed (position, byte) in a sorted exception list -- they become literal phrases.
//
//   k_pack_count : one CTA per 4096-byte tile; 16 bytes per thread (one 128-bit load);
//                  SIMD-in-register validity test handles the all-ACGT case in ~10 integer ops per 4 bytes
//   device_scan  : tile counts -> output base offsets (whole batch at once)
//   k_pack_write : same tiling; block scan of per-thread counts, bases assembled in a shared-memory bit
//                  buffer, then written as aligned 64-bit words (interior words plain stores, the two
//                  boundary words OR-ed atomically into the zero-filled output)
// HBM traffic: 2 * M read (count + write pass) + m/4 written; a single-pass chained-scan variant is the
// obvious next step.
#include "drlz_internal.h"
#include "scan.cuh"

namespace drlz {

// 4 ASCII bytes -> per-byte 2-bit codes (A0 C1 G2 T3) in the low 2 bits of each byte; *ok = all four valid
__device__ __forceinline__ uint32_t codes4(uint32_t w, bool* ok) {
    uint32_t c = (w >> 1) & 0x03030303u;                  // A0 C1 T2 G3
    uint32_t m = (c >> 1) & ~c & 0x01010101u;             // 1 where T
    uint32_t expect = 0x41414141u + (c << 1) + (m << 4) - m;
    *ok = expect == w;
    return c ^ ((c >> 1) & 0x01010101u);                  // swap 2 <-> 3
}
// four per-byte codes (first byte = lowest address) -> 8 bits, first base in the top bits
__device__ __forceinline__ uint32_t squeeze4(uint32_t c) { return (c * 0x40100401u) >> 24; }

struct Seg16 {
    uint32_t cnt;      // non-gap bytes
    uint32_t exc;      // exceptions among them
    uint32_t codes;    // cnt codes, left aligned in 32 bits
};

__device__ __forceinline__ uint4 load16(const uint8_t* p) { return *reinterpret_cast<const uint4*>(p); }

__device__ __forceinline__ Seg16 classify16(uint4 v, int nvalid, int strip) {
    Seg16 s;
    bool k0, k1, k2, k3;
    uint32_t c0 = codes4(v.x, &k0), c1 = codes4(v.y, &k1), c2 = codes4(v.z, &k2), c3 = codes4(v.w, &k3);
    if (nvalid == 16 && k0 && k1 && k2 && k3) {
        s.cnt = 16; s.exc = 0;
        s.codes = (squeeze4(c0) << 24) | (squeeze4(c1) << 16) | (squeeze4(c2) << 8) | squeeze4(c3);
        return s;
    }
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t cnt = 0, exc = 0, codes = 0;
#pragma unroll
    for (int b = 0; b < 16; ++b) {
        uint32_t byte = (w[b >> 2] >> (8 * (b & 3))) & 255u;
        bool gap = b >= nvalid || (strip && byte == '-');
        if (!gap) {
            uint32_t c = (byte >> 1) & 3u; c ^= c >> 1;
            bool valid = byte == 'A' || byte == 'C' || byte == 'G' || byte == 'T';
            if (!valid) { c = 0; ++exc; }
            codes = (codes << 2) | c;
            ++cnt;
        }
    }
    s.cnt = cnt; s.exc = exc;
    s.codes = cnt ? codes << (2 * (16 - cnt)) : 0u;
    return s;
}

__global__ void __launch_bounds__(256) k_pack_count(PackJob job) {
    uint32_t h = blockIdx.y, t = blockIdx.x;
    uint64_t cols = job.cols[h];
    uint64_t off = (uint64_t)t * kPackTileBytes + (uint64_t)threadIdx.x * 16;
    uint32_t packed = 0;
    if (off < cols) {
        int nvalid = cols - off >= 16 ? 16 : (int)(cols - off);
        Seg16 s = classify16(load16(job.rows + job.row_off[h] + off), nvalid, job.strip);
        packed = s.cnt | (s.exc << 16);
    }
    __shared__ uint32_t sm[33];
    uint32_t total;
    block_exclusive_scan<uint32_t, OpSum>(packed, OpSum(), 0u, &total, sm);
    if (threadIdx.x == 0) {
        job.tile_cnt[(uint64_t)h * job.tiles_per_row + t] = total & 0xFFFFu;
        job.tile_exc[(uint64_t)h * job.tiles_per_row + t] = total >> 16;
    }
}

// per-row totals from the scanned tile offsets
__global__ void k_pack_rows(PackJob job, const uint64_t* total_cnt, const uint64_t* total_exc) {
    uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h > job.H) return;
    uint64_t T = job.tiles_per_row;
    uint64_t nt = (uint64_t)job.H * T;
    uint64_t eo = h < job.H ? job.tile_excoff[h * T] : *total_exc;
    job.exc_off[h] = eo;
    if (h < job.H) {
        uint64_t a = job.tile_off[h * T];
        uint64_t b = (uint64_t)(h + 1) * T < nt ? job.tile_off[(uint64_t)(h + 1) * T] : *total_cnt;
        job.m_out[h] = b - a;
    } else if (eo > job.exc_cap) atomicOr(job.flags, 1u);
}

__global__ void __launch_bounds__(256) k_pack_write(PackJob job) {
    __shared__ uint32_t sbits[kPackTileBytes / 16 + 8];
    __shared__ uint32_t sm[33];
    uint32_t h = blockIdx.y, t = blockIdx.x;
    uint64_t cols = job.cols[h];
    uint64_t tile_start = (uint64_t)t * kPackTileBytes;
    if (tile_start >= cols) return;
    for (int i = threadIdx.x; i < kPackTileBytes / 16 + 8; i += 256) sbits[i] = 0;
    uint64_t off = tile_start + (uint64_t)threadIdx.x * 16;
    Seg16 s; s.cnt = 0; s.exc = 0; s.codes = 0;
    uint4 v = make_uint4(0, 0, 0, 0);
    int nvalid = 0;
    if (off < cols) {
        nvalid = cols - off >= 16 ? 16 : (int)(cols - off);
        v = load16(job.rows + job.row_off[h] + off);
        s = classify16(v, nvalid, job.strip);
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan<uint32_t, OpSum>(s.cnt | (s.exc << 16), OpSum(), 0u, &total, sm);
    uint32_t lo = ex & 0xFFFFu, elo = ex >> 16;     // local base offset, local exception offset
    uint32_t T = total & 0xFFFFu;
    if (s.cnt) {
        uint32_t wi = lo >> 4, sh = (lo & 15u) * 2;
        atomicOr(&sbits[wi], s.codes >> sh);
        if (sh && s.cnt > 16 - (lo & 15u)) atomicOr(&sbits[wi + 1], s.codes << (32 - sh));
    }
    uint64_t gidx = (uint64_t)h * job.tiles_per_row + t;
    uint64_t rowbase = job.tile_off[(uint64_t)h * job.tiles_per_row];
    uint64_t o = job.tile_off[gidx] - rowbase;      // first output base of this tile within the haplotype
    // exceptions and optional gapless bytes (rare / optional paths: plain byte loops)
    if (s.exc || job.gapless) {
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint64_t e = job.tile_excoff[gidx] + elo;
        uint64_t p = o + lo;
        uint8_t* gl = job.gapless ? job.gapless + job.gapless_off[h] : nullptr;
        for (int b = 0; b < nvalid; ++b) {
            uint32_t byte = (w[b >> 2] >> (8 * (b & 3))) & 255u;
            if (job.strip && byte == '-') continue;
            bool valid = byte == 'A' || byte == 'C' || byte == 'G' || byte == 'T';
            if (!valid) {
                if (e < job.exc_cap) { job.exc_pos[e] = (uint32_t)p; job.exc_byte[e] = (uint8_t)byte; }
                ++e;
            }
            if (gl) gl[p] = (uint8_t)byte;
            ++p;
        }
    }
    __syncthreads();
    if (T == 0) return;
    // assemble aligned 64-bit output words
    uint32_t lead = (uint32_t)(o & 31);
    uint64_t W0 = o >> 5;
    uint32_t nwords = (lead + T + 31) >> 5;
    uint64_t* X = job.X + job.xoff[h];
    for (uint32_t j = threadIdx.x; j < nwords; j += 256) {
        int64_t lb = (int64_t)j * 32 - (int64_t)lead;      // local base index of the word's first base
        uint64_t val;
        uint32_t b0 = lb < 0 ? 0u : (uint32_t)lb;
        uint32_t wi = b0 >> 4, sh = (b0 & 15u) * 2;
        uint64_t hi = ((uint64_t)sbits[wi] << 32) | sbits[wi + 1];
        val = sh ? (hi << sh) | ((uint64_t)sbits[wi + 2] >> (32 - sh)) : hi;
        if (lb < 0) val >>= 2 * lead;
        bool first_partial = j == 0 && lead != 0;
        bool last_partial = j == nwords - 1 && ((lead + T) & 31u) != 0;
        if (first_partial || last_partial) atomicOr(reinterpret_cast<unsigned long long*>(&X[W0 + j]), (unsigned long long)val);
        else X[W0 + j] = val;
    }
}

void launch_pack(const PackJob& job, cudaStream_t s, cudaEvent_t* ev) {
    if (job.H == 0 || job.tiles_per_row == 0) { if (ev) for (int i = 0; i < 3; ++i) cudaEventRecord(ev[i], s); return; }
    dim3 grid(job.tiles_per_row, job.H);
    uint64_t nt = (uint64_t)job.H * job.tiles_per_row;
    g_launches += 3;
    k_pack_count<<<grid, 256, 0, s>>>(job);
    if (ev) cudaEventRecord(ev[0], s);
    uint64_t* tmp1 = job.scan_tmp;
    uint64_t* tmp2 = job.scan_tmp + scan_tmp_elems(nt);
    device_scan<uint64_t, LoadArr<uint32_t, uint64_t>, StoreArr<uint64_t>, OpSum, false>(
        LoadArr<uint32_t, uint64_t>{job.tile_cnt}, StoreArr<uint64_t>{job.tile_off}, nt, tmp1, OpSum(), 0ull, s);
    device_scan<uint64_t, LoadArr<uint32_t, uint64_t>, StoreArr<uint64_t>, OpSum, false>(
        LoadArr<uint32_t, uint64_t>{job.tile_exc}, StoreArr<uint64_t>{job.tile_excoff}, nt, tmp2, OpSum(), 0ull, s);
    uint64_t ntiles_scan = (nt + kScanTile - 1) / kScanTile;
    k_pack_rows<<<(job.H + 1 + 127) / 128, 128, 0, s>>>(job, tmp1 + ntiles_scan, tmp2 + ntiles_scan);
    if (ev) cudaEventRecord(ev[1], s);
    k_pack_write<<<grid, 256, 0, s>>>(job);
    if (ev) cudaEventRecord(ev[2], s);
}

__global__ void k_make_hapviews(HapView* haps, const uint64_t* X, const uint64_t* xoff, const uint64_t* m,
                                const uint32_t* exc_pos, const uint8_t* exc_byte, const uint64_t* exc_off, uint32_t H) {
    uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= H) return;
    HapView hv;
    hv.x = X + xoff[h];
    hv.m = m[h];
    hv.exc_pos = exc_pos + exc_off[h];
    hv.exc_byte = exc_byte + exc_off[h];
    hv.n_exc = (uint32_t)(exc_off[h + 1] - exc_off[h]);
    haps[h] = hv;
}

void launch_make_hapviews(HapView* haps, const uint64_t* X, const uint64_t* xoff, const uint64_t* m,
                          const uint32_t* exc_pos, const uint8_t* exc_byte, const uint64_t* exc_off, uint32_t H,
                          cudaStream_t s) {
    if (H == 0) return;
    g_launches += 1;
    k_make_hapviews<<<(H + 127) / 128, 128, 0, s>>>(haps, X, xoff, m, exc_pos, exc_byte, exc_off, H);
}

}  // namespace drlz
