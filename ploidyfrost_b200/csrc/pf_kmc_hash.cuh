// pf_kmc_hash.cuh -- the one-sector k-mer index of libpfgpu.so and the lookup kernel that walks it.
//
// Why a second index layout.  The verbatim KMC image (prefix table + sorted suffix records, pf_kmc.cu) costs the
// reference's own dependent chain per lookup -- signature map -> LUT pair -> ceil(log2(bucket)) record probes -- which
// on the 196 M k-mer database of BASELINE config 2 is 196 bytes of DRAM traffic (six 32-byte sectors) and ~900
// thread-instructions per lookup (profiles/r01d_lookup_sorted_ncu.txt).  Random 32-byte sectors are the scarce
// resource of this path (the box sustains ~38 G of them per second), so the index is re-hashed once at open time
// into buckets of exactly one sector:
//
//   bucket = 4 slots x 8 bytes, 32-byte aligned; table = 2^b buckets, b chosen so that a bucket holds <= 2 keys on average
//   h      = mix(key) -- a bijection on the 2k-bit key space (odd multiply, xor-shift, odd multiply, xor-shift)
//   home   = top b bits of h;   rem = the other 2k-b bits (quotienting: home + rem identify the key)
//   slot   = dist << (rem_bits + 8C) | rem << 8C | counter        (all ones = empty)
//   a key that finds its home bucket full goes to home+1, home+2, ... (dist <= 7); slots only ever go from empty to
//   full, so "this bucket still has an empty slot" ends an unsuccessful search.
//
// One lookup = one 256-bit load (LDG.E.256) in ~98 % of the cases.  The answer is the reference's
// (CheckKmer, kmc_file.cpp:330-366: counter of the key if present and min_count <= counter <= max_count) whenever the
// database is one the reference itself can search -- records strictly ascending inside every prefix bucket and, for
// KMC2, every record in the bin its signature maps to.  The build kernel verifies exactly that for every record; a
// database that fails the check keeps the verbatim index and the chain of pf_kmc.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pf_types.h"

namespace pfkmc {

constexpr unsigned long long H_EMPTY = ~0ull;
constexpr uint32_t H_DIST_BITS = 3, H_MAX_DIST = 7, H_SLOTS = 4;

struct HashView {
    unsigned long long *tab;   // 2^bucket_bits buckets x 4 slots
    uint32_t bucket_bits;      // b
    uint32_t rem_bits;         // 2k - b
    uint32_t cbits;            // 8 * counter_size
    uint32_t kbits;            // 2k
};

// bijection on [0, 2^kbits): every step is invertible (odd multiply mod 2^kbits; x ^= x >> s)
__host__ __device__ __forceinline__ uint64_t hash_mix(uint64_t key, uint32_t kbits) {
    const uint64_t mask = kbits >= 64 ? ~0ull : ((1ull << kbits) - 1);
    const uint32_t s = (kbits + 1) / 2;
    uint64_t h = key;
    h = (h * 0x9E3779B97F4A7C15ull) & mask;
    h ^= h >> s;
    h = (h * 0xD6E8FEB86659FD93ull) & mask;
    h ^= h >> s;
    return h;
}

// which partition of a partitioned hash index owns a key: the low end of the mixed value (the bucket uses the high end)
__host__ __device__ __forceinline__ uint32_t hash_owner(uint64_t key, uint32_t kbits, uint32_t n_parts) {
    return (uint32_t)(hash_mix(key, kbits) % n_parts);
}

__device__ __forceinline__ void ld_bucket(const unsigned long long *p, unsigned long long v[4]) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3])
                 : "l"(p));
}

__device__ __forceinline__ uint64_t revcomp2(uint64_t v, uint32_t k) {
    uint64_t x = __brevll(~v);
    x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
    return x >> (64 - 2 * k);
}

// Search for the key whose mixed value is h, the home bucket's four slots already in v[].
__device__ __forceinline__ bool hash_resolve(const HashView &hv, const unsigned long long *tab, uint64_t h, unsigned long long v[4],
                                             uint32_t &cnt) {
    const uint64_t nb_mask = (1ull << hv.bucket_bits) - 1;
    const uint64_t home = h >> hv.rem_bits;
    const uint64_t rem = hv.rem_bits ? (h & ((1ull << hv.rem_bits) - 1)) : 0ull;
    const uint64_t cmask = (1ull << hv.cbits) - 1;
    for (uint32_t d = 0;; d++) {
        const uint64_t tag = ((uint64_t)d << hv.rem_bits) | rem;
        bool open = false;
#pragma unroll
        for (int s = 0; s < 4; s++) {
            if ((v[s] >> hv.cbits) == tag) { cnt = (uint32_t)(v[s] & cmask); return true; }
            open |= v[s] == H_EMPTY;
        }
        if (open || d == H_MAX_DIST) return false;
        ld_bucket(tab + 4 * ((home + d + 1) & nb_mask), v);
    }
}
__device__ __forceinline__ bool hash_resolve(const HashView &hv, uint64_t h, unsigned long long v[4], uint32_t &cnt) {
    return hash_resolve(hv, hv.tab, h, v, cnt);
}

__device__ __forceinline__ bool hash_find(const HashView &hv, const unsigned long long *tab, uint64_t key, uint32_t &cnt) {
    const uint64_t h = hash_mix(key, hv.kbits);
    unsigned long long v[4];
    ld_bucket(tab + 4 * (h >> hv.rem_bits), v);
    return hash_resolve(hv, tab, h, v, cnt);
}
__device__ __forceinline__ bool hash_find(const HashView &hv, uint64_t key, uint32_t &cnt) { return hash_find(hv, hv.tab, key, cnt); }

// Insert (build time).  Returns false when the key would sit more than H_MAX_DIST buckets from home.
__device__ __forceinline__ bool hash_insert(const HashView &hv, uint64_t key, uint64_t counter) {
    const uint64_t h = hash_mix(key, hv.kbits);
    const uint64_t nb_mask = (1ull << hv.bucket_bits) - 1;
    const uint64_t home = h >> hv.rem_bits;
    const uint64_t rem = hv.rem_bits ? (h & ((1ull << hv.rem_bits) - 1)) : 0ull;
    for (uint32_t d = 0; d <= H_MAX_DIST; d++) {
        const unsigned long long val = ((((uint64_t)d << hv.rem_bits) | rem) << hv.cbits) | counter;
        unsigned long long *b = hv.tab + 4 * ((home + d) & nb_mask);
        for (int s = 0; s < 4; s++) {
            if (b[s] != H_EMPTY) continue;
            if (atomicCAS(b + s, H_EMPTY, val) == H_EMPTY) return true;
        }
    }
    return false;
}

// ---- the lookup kernel -----------------------------------------------------------------------------------------------
constexpr int HL_THREADS = 256;
constexpr int HL_PPT = 8;                          // window starts per thread
constexpr int HL_TILE = HL_THREADS * HL_PPT;       // 2048 base positions per tile
constexpr int HL_WORDS = HL_TILE / 16 + 3;         // 16 bases per packed word + 48 bases of halo (k-1+7 <= 38)
constexpr int HL_NSEQ = 512;
constexpr int PF_MAX_PEERS = 16;                       // sequence ends staged per tile (more than that: per-thread search)

struct HashLookupArgs {
    HashView hv;
    uint32_t k;
    uint32_t min_count;
    uint64_t max_count;
    int mode;          // PF_LOOKUP_*; FWD_THEN_RC arrives as CANONICAL when the database is verified canonical
    const uint8_t *bases;
    uint64_t n_bases;
    const uint64_t *seq_off;
    const uint64_t *win_off;
    uint32_t n_seq;
    uint32_t low, up;
    uint32_t *counts;
    uint8_t *found;
    pf_cov_t *cov;
    uint64_t n_tiles;
    const unsigned long long *peer_tab[PF_MAX_PEERS];   // peer-memory form: the slice of every partition (own slice included), else unused
    uint32_t peer_lookup;                  // 1: every lookup loads the owner's bucket through peer_tab (NVLink), n_parts partitions
    uint32_t n_parts;                      // ROUTE instantiation / peer lookups: partitions of the index
    unsigned long long *route_keys;        // ROUTE: key of every live window ...
    uint8_t *route_owner;                  // ... and the partition that owns it (windows that are not looked up keep 0xFF)
    const uint32_t *tile_seq;   // [n_tiles + 1] sequence containing the first base of each tile (tile_seq_kernel)
};

// tile_seq[t] = last s with seq_off[s] <= t * HL_TILE (clamped to the last sequence); one thread per tile
__global__ void tile_seq_kernel(const uint64_t *__restrict__ seq_off, uint32_t n_seq, uint64_t n_tiles, uint32_t *__restrict__ tile_seq) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    const uint64_t g = t * HL_TILE;
    uint32_t lo = 0, hi = n_seq;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (seq_off[mid] <= g) lo = mid; else hi = mid;
    }
    tile_seq[t] = lo;
}

struct CovRun {   // readCov partials of one (thread, sequence) run  (CDBG.cpp:29-120)
    uint32_t s;
    uint64_t sum;
    uint32_t mn, fm, fo;
    __device__ __forceinline__ void reset(uint32_t seq) { s = seq; sum = 0; mn = fm = fo = 0xFFFFFFFFu; }
    __device__ __forceinline__ void flush(pf_cov_t *cov) const {
        if (s == 0xFFFFFFFFu) return;
        if (sum) atomicAdd((unsigned long long *)&cov[s].sum, (unsigned long long)sum);
        if (mn != 0xFFFFFFFFu) atomicMin(&cov[s].min, mn);
        if (fm != 0xFFFFFFFFu) atomicMin((unsigned int *)&cov[s].first_missing, fm);
        if (fo != 0xFFFFFFFFu) atomicMin((unsigned int *)&cov[s].first_outside, fo);
    }
};

// One CTA walks 2048-base tiles.  The tile is staged once into shared memory as 2-bit codes (16 per word, first base
// in the top bits) plus a not-a-symbol bit per base; after that every thread is on its own: it owns 8 consecutive
// window starts, takes its first k-mer from three packed words, rolls the forward and reverse-complement values from
// base to base in registers, and keeps four bucket loads in flight at a time.  The readCov reductions are accumulated
// per (thread, sequence) run and leave the warp as one set of atomics per (warp, sequence).
// MODE 1 = ROUTE (partitioned index): nothing is looked up here -- the key of every live window and the partition that owns it
// are written out for the exchange (pf_kmc_route_dev).
// MODE 2 (peer-memory form): like 0, but the bucket of a key is loaded from the slice of the partition that owns it.
template <int MODE>
__global__ void __launch_bounds__(HL_THREADS, 3) kmc_hash_lookup_kernel(const HashLookupArgs a) {
    constexpr bool ROUTE = MODE == 1, PEER = MODE == 2;
    __shared__ uint32_t s_pk[HL_WORDS];
    __shared__ uint32_t s_bad[HL_WORDS];
    __shared__ uint32_t s_end[HL_NSEQ];   // end of the tile's sequences relative to the tile start (saturated)
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t k = a.k;
    const uint64_t kmask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1);
    const bool aligned16 = ((uintptr_t)a.bases & 15) == 0;
    const uint64_t n_bases = min(a.n_bases, (uint64_t)__ldg((const unsigned long long *)a.seq_off + a.n_seq));
    const uint32_t n_seq = a.n_seq;

    for (uint64_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const uint64_t p0 = tile * HL_TILE;
        // ---- stage: 16 characters -> one packed word + 16 validity bits (kmer_api.h:264-275: ACGT/acgt are symbols) ----
        if (tid < HL_WORDS) {
            const uint64_t g0 = p0 + (uint64_t)tid * 16;
            uint32_t ch[4];
            if (aligned16 && g0 + 16 <= n_bases) {
                const uint4 v = __ldg((const uint4 *)(a.bases + g0));
                ch[0] = v.x; ch[1] = v.y; ch[2] = v.z; ch[3] = v.w;
            } else {
#pragma unroll
                for (int w = 0; w < 4; w++) {
                    uint32_t x = 0;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint64_t g = g0 + w * 4 + j;
                        const uint32_t c = g < n_bases ? (uint32_t)__ldg(a.bases + g) : (uint32_t)'N';
                        x |= c << (8 * j);
                    }
                    ch[w] = x;
                }
            }
            uint32_t pk = 0, bad = 0;
#pragma unroll
            for (int w = 0; w < 4; w++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t c = (ch[w] >> (8 * j)) & 0xFFu;
                    const uint32_t x = (c >> 1) & 3u;                        // A 0, C 1, T 2, G 3 (either case)
                    const uint32_t u = (c | 0x20u) - (uint32_t)'a';
                    const uint32_t ok = u < 20u ? ((0x80045u >> u) & 1u) : 0u;   // a, c, g, t
                    pk = (pk << 2) | (x ^ (x >> 1));                         // A 0, C 1, G 2, T 3
                    bad = (bad << 1) | (ok ^ 1u);
                }
            s_pk[tid] = pk;
            s_bad[tid] = bad;
        }
        // the sequences this tile touches: [s_lo, s_hi]; their ends, relative to p0, go to shared memory
        const uint32_t s_lo = __ldg(a.tile_seq + tile), s_hi = __ldg(a.tile_seq + tile + 1);
        const uint32_t ns = s_hi - s_lo + 1;
        const bool staged = ns <= (uint32_t)HL_NSEQ;
        if (staged)
            for (uint32_t i = tid; i < ns; i += HL_THREADS) {
                const uint64_t e = __ldg((const unsigned long long *)a.seq_off + s_lo + i + 1);
                s_end[i] = e - p0 > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)(e - p0);   // e > p0 except for empty leading sequences
            }
        __syncthreads();

        CovRun run;
        run.reset(0xFFFFFFFFu);
        const uint32_t q0 = tid * HL_PPT;
        uint64_t g = p0 + q0;
        if (g < n_bases) {
            // sequence containing base g: last s with seq_off[s] <= g   (seq_off[0] == 0)
            uint32_t s;
            if (staged) {                                                    // first staged sequence that ends after q0
                uint32_t lo = 0, hi = ns - 1;                                // the last one ends after every base of the tile
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (s_end[mid] > q0) hi = mid; else lo = mid + 1;
                }
                s = s_lo + lo;
            } else {
                uint32_t lo = s_lo, hi = n_seq;
                while (hi - lo > 1) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (__ldg((const unsigned long long *)a.seq_off + mid) <= g) lo = mid; else hi = mid;
                }
                s = lo;
            }
            uint64_t sb = __ldg((const unsigned long long *)a.seq_off + s);
            uint64_t se = __ldg((const unsigned long long *)a.seq_off + s + 1);
            uint64_t wo = __ldg((const unsigned long long *)a.win_off + s);

            const uint32_t wi0 = q0 >> 4, intra = q0 & 15u;                  // intra is 0 or 8
            const uint32_t w0 = s_pk[wi0], w1 = s_pk[wi0 + 1], w2 = s_pk[wi0 + 2];
            uint64_t v = ((uint64_t)w0 << 32) | w1;                          // 32 bases from the word boundary
            uint64_t nx = (uint64_t)w2 << 32;                                // the 16 after them, left-aligned
            uint64_t bb = ((uint64_t)s_bad[wi0] << 48) | ((uint64_t)s_bad[wi0 + 1] << 32) | ((uint64_t)s_bad[wi0 + 2] << 16);
            if (intra) { v = (v << 16) | (w2 >> 16); nx <<= 16; bb <<= 8; }
            uint64_t fwd = v >> (64 - 2 * k);
            uint64_t rc = revcomp2(fwd, k);
            uint64_t rest = k >= 32 ? nx : ((v << (2 * k)) | (nx >> (64 - 2 * k)));   // bases after the first window
            const uint32_t rc_shift = 2 * k - 2;

#pragma unroll
            for (int half = 0; half < 2; half++) {
                uint64_t hk[4];
                uint64_t alt[4];
                uint32_t wi[4], wl[4], sq[4];
                int st[4];                                                   // 0 no window, 1 window with a non-symbol, 2 lookup
                unsigned long long bk[4][4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int t = half * 4 + u;
                    const uint64_t gq = g + t;
                    st[u] = 0;
                    wi[u] = 0; wl[u] = 0; sq[u] = 0;
                    if (gq < n_bases) {
                        while (gq >= se) {                                   // next sequence (skips empty ones)
                            s++;
                            sb = se;
                            se = __ldg((const unsigned long long *)a.seq_off + s + 1);
                            wo = __ldg((const unsigned long long *)a.win_off + s);
                        }
                        if (gq + k <= se) {
                            wl[u] = (uint32_t)(gq - sb);
                            wi[u] = (uint32_t)(wo + (gq - sb));
                            sq[u] = s;
                            st[u] = ((bb << t) >> (64 - k)) ? 1 : 2;
                        }
                    }
                    const uint64_t key = a.mode == PF_LOOKUP_CANONICAL ? (fwd < rc ? fwd : rc) : fwd;
                    alt[u] = rc;
                    hk[u] = hash_mix(key, a.hv.kbits);
                    if (ROUTE) {
                        if (st[u] == 2) { a.route_keys[wi[u]] = key; a.route_owner[wi[u]] = (uint8_t)(hk[u] % a.n_parts); }
                    } else if (st[u] == 2) {
                        // peer-memory form: the bucket lives in the slice of the partition that owns the key (a load over NVLink)
                        const unsigned long long *tab = PEER ? a.peer_tab[hk[u] % a.n_parts] : a.hv.tab;
                        ld_bucket(tab + 4 * (hk[u] >> a.hv.rem_bits), bk[u]);
                    }
                    const uint64_t c = rest >> 62;                           // roll to the next window
                    rest <<= 2;
                    fwd = ((fwd << 2) | c) & kmask;
                    rc = (rc >> 2) | ((3ull - c) << rc_shift);
                }
                if (ROUTE) continue;
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (st[u] == 0) continue;
                    uint32_t cnt = 0;
                    bool ok = false;
                    if (st[u] == 2) {
                        const unsigned long long *tab = PEER ? a.peer_tab[hk[u] % a.n_parts] : a.hv.tab;
                        ok = hash_resolve(a.hv, tab, hk[u], bk[u], cnt);
                        ok = ok && cnt >= a.min_count && (uint64_t)cnt <= a.max_count;            // kmc_file.cpp:1459
                        if (!ok && a.mode == PF_LOOKUP_FWD_THEN_RC) {                             // CDBG.cpp:38-43
                            const unsigned long long *tab2 = PEER ? a.peer_tab[hash_mix(alt[u], a.hv.kbits) % a.n_parts] : a.hv.tab;
                            ok = hash_find(a.hv, tab2, alt[u], cnt);
                            ok = ok && cnt >= a.min_count && (uint64_t)cnt <= a.max_count;
                        }
                        if (!ok) cnt = 0;
                    }
                    if (a.counts) a.counts[wi[u]] = cnt;
                    if (a.found) a.found[wi[u]] = ok ? 1 : 0;
                    if (a.cov) {
                        if (sq[u] != run.s) { run.flush(a.cov); run.reset(sq[u]); }
                        if (ok) {
                            run.sum += cnt;
                            run.mn = min(run.mn, cnt);
                            if (!(cnt > a.low && cnt < a.up)) run.fo = min(run.fo, wl[u]);
                        } else run.fm = min(run.fm, wl[u]);
                    }
                }
            }
        }
        if (!ROUTE && a.cov) {   // the thread's last run leaves through the warp: one set of atomics per (warp, sequence)
            const uint32_t grp = __match_any_sync(0xffffffffu, run.s);
            const uint32_t leader = __ffs(grp) - 1;
            const uint32_t s_lo = __reduce_add_sync(grp, (uint32_t)(run.sum & 0xFFFFFu));
            const uint32_t s_hi = __reduce_add_sync(grp, (uint32_t)(run.sum >> 20));
            CovRun tot;
            tot.s = run.s;
            tot.sum = (uint64_t)s_lo + ((uint64_t)s_hi << 20);
            tot.mn = __reduce_min_sync(grp, run.mn);
            tot.fm = __reduce_min_sync(grp, run.fm);
            tot.fo = __reduce_min_sync(grp, run.fo);
            if (lane == leader) tot.flush(a.cov);
        }
        __syncthreads();   // the staging arrays are reused by the next tile
    }
}

// u64 keys (right-aligned 2-bit k-mers) -> counters; the form the partition owner's side uses
__global__ void kmc_hash_lookup_keys_kernel(const HashView hv, uint32_t min_count, uint64_t max_count,
                                            const unsigned long long *__restrict__ keys, uint64_t n,
                                            uint32_t *__restrict__ counts, uint8_t *__restrict__ found) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t c = 0;
    bool ok = hash_find(hv, keys[i], c);
    ok = ok && c >= min_count && (uint64_t)c <= max_count;
    counts[i] = ok ? c : 0;
    found[i] = ok ? 1 : 0;
}

}  // namespace pfkmc
