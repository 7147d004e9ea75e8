// pf_kmc.cu -- HBM-resident KMC index + batched k-mer lookup kernel for sm_100a.
//
// Stands in for CKMCFile's random-access path (KMC/kmc_api/kmc_file.cpp): OpenForRA :27,
// ReadParamsFrom_prefix_file_buf :185, CheckKmer :330, BinarySearch :1383, GetCountersForRead :904,
// and for the coverage reducers built on it (src/CDBG.cpp:29-120).  Semantics are the reference's;
// the data layout and the execution are not:
//
//  * .kmc_pre  -> `lut`   : the prefix table verbatim (u32 entries when N+1 < 2^32, else u64), with the
//                           N+1 end sentinel the reader plants (:233, :292); `sigmap` (KMC2) verbatim;
//                           `norm` = the m-mer normalisation table of mmer.h:77-87, built on the host.
//  * .kmc_suf  -> `rec`   : every R = S+C byte record re-packed on the GPU into one aligned u64
//                           (suffix << 8C | counter) when R <= 8 -- one 8-byte load returns key and
//                           counter, a 32-byte sector holds four records; otherwise split into
//                           `suf` (u64 suffix) + `cnt` (u32 counter).  Order is unchanged, so the
//                           reference's bucket ranges and search outcome are unchanged.
//
//  * lookup kernel: one CTA walks 2048-base tiles of the flat `bases` array.  The tile is staged into
//    shared memory with 128-bit coalesced loads and turned into 2-bit codes; the m-mer normalisation
//    is evaluated once per base position (not once per m-mer per window as get_signature does,
//    kmer_api.h:653-672); window starts are compacted so that every thread of the search phase owns a
//    live lookup; the search phase issues the dependent chain sigmap -> LUT pair -> records, and the
//    readCov reductions are folded in with warp-aggregated atomics.  HBM-bound integer work: no
//    tensor cores.
#include "pf_common.cuh"
#include "pf_kmc_hash.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstring>
#include <memory>

__global__ void pf_iota_u32(uint32_t *p, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}

namespace {

struct KmcView {  // by-value kernel argument
    uint32_t k, p, S, C, sig_len, is_kmc2, lut64, packed;
    uint32_t min_count;
    uint64_t max_count;
    uint64_t N;           // records held by THIS index (all of the database, or one partition of it)
    uint64_t single_lut;  // 4^p
    uint64_t lut_n;       // entries including the sentinel
    uint32_t n_parts, part;   // partitioned index: KMC2 bins with bin % n_parts == part, KMC1 prefixes of one range
    uint64_t prefix_lo, prefix_cnt;   // KMC1 partition: owned prefixes [prefix_lo, prefix_lo + prefix_cnt)
    uint64_t prefix_per_part;         // KMC1: ceil(4^p / n_parts)
    const void *lut;
    const uint32_t *sigmap;
    const uint32_t *norm;
    const uint64_t *rec;  // packed layout
    const uint64_t *suf;  // split layout
    const uint32_t *cnt;
};

constexpr int LK_THREADS = 256;
constexpr int LK_PPT = 8;                       // base positions per thread in the classification phase
constexpr int LK_TILE = LK_THREADS * LK_PPT;    // 2048 base positions per tile
constexpr int LK_HALO = 32;                     // >= k-1 for k <= 32

__device__ __forceinline__ uint32_t base_code(uint32_t c) {  // kmer_api.h:264-275; 4 = not a symbol
    c |= 0x20u;
    return c == 'a' ? 0u : c == 'c' ? 1u : c == 'g' ? 2u : c == 't' ? 3u : 4u;
}

__device__ __forceinline__ uint64_t revcomp64(uint64_t v, uint32_t k) {
    uint64_t x = __brevll(~v);  // reverses bit order; fix the order inside each 2-bit symbol
    x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
    return x >> (64 - 2 * k);
}

__device__ __forceinline__ uint64_t lut_at(const KmcView &db, uint64_t i) {
    return db.lut64 ? __ldg((const unsigned long long *)db.lut + i) : (uint64_t)__ldg((const uint32_t *)db.lut + i);
}

// Which partition owns a key (KMC2: by bin, KMC1: by prefix range); 0 for an unpartitioned index.
__device__ __forceinline__ uint32_t kmc_owner(const KmcView &db, uint64_t key, uint32_t bin) {
    if (db.n_parts <= 1) return 0;
    if (db.is_kmc2) return bin % db.n_parts;
    return (uint32_t)((key >> (2 * (db.k - db.p))) / db.prefix_per_part);
}

// CheckKmer (:330-366) + BinarySearch (:1383-1462) for one key; `bin` = signature_map[signature] (0 for KMC1).
// A key whose bin / prefix is not held by this (partitioned) index is reported absent.
__device__ __forceinline__ bool kmc_search(const KmcView &db, uint64_t key, uint32_t bin, uint32_t &count) {
    const uint32_t sbits = 2 * (db.k - db.p);
    const uint64_t prefix = key >> sbits;  // sbits < 64 because p >= 1
    const uint64_t suffix = key & ((1ull << sbits) - 1);
    uint64_t slot;
    if (db.is_kmc2) {
        if (db.n_parts > 1) {
            if (bin % db.n_parts != db.part) return false;
            bin /= db.n_parts;
        }
        slot = (uint64_t)bin * db.single_lut + prefix;
    } else {
        if (prefix < db.prefix_lo || prefix - db.prefix_lo >= db.prefix_cnt) return false;
        slot = prefix - db.prefix_lo;
    }
    if (slot + 1 >= db.lut_n) return false;
    long long lo = (long long)lut_at(db, slot);
    long long hi = (long long)lut_at(db, slot + 1) - 1;
    if (lo >= (long long)db.N) return false;                 // :1385
    if (hi > (long long)db.N - 1) hi = (long long)db.N - 1;  // the last bucket's stop is one past the end (:233,:292)
    const uint32_t cbits = 8 * db.C;
    while (lo <= hi) {
        const long long mid = (lo + hi) >> 1;
        uint64_t rs, c;
        if (db.packed) {
            const uint64_t r = __ldg((const unsigned long long *)db.rec + mid);
            rs = r >> cbits;
            c = r & ((1ull << cbits) - 1);
        } else {
            rs = __ldg((const unsigned long long *)db.suf + mid);
            c = 0;
        }
        if (rs == suffix) {
            if (!db.packed) c = __ldg(db.cnt + mid);
            count = (uint32_t)c;
            return c >= db.min_count && c <= db.max_count;  // :1459
        }
        if (rs < suffix) lo = mid + 1; else hi = mid - 1;
    }
    return false;
}

// ROUTE = false: look every window up (this index holds the whole database).
// ROUTE = true : partitioned database -- emit the key of every live window and the partition that owns it
//                (route_keys[wi], route_owner[wi]; 0xFF = window is not looked up) instead of searching.
template <bool ROUTE>
__global__ void __launch_bounds__(LK_THREADS)
kmc_lookup_kernel(const KmcView db, const uint8_t *__restrict__ bases, const uint64_t n_bases_arg,
                  const uint64_t *__restrict__ seq_off, const uint64_t *__restrict__ win_off, const uint32_t n_seq,
                  const int mode, const uint32_t low, const uint32_t up, uint32_t *__restrict__ counts,
                  uint8_t *__restrict__ found, pf_cov_t *__restrict__ cov, const uint64_t n_tiles,
                  unsigned long long *__restrict__ route_keys, uint8_t *__restrict__ route_owner) {
    __shared__ __align__(16) uint8_t s_code[LK_TILE + LK_HALO];
    __shared__ uint32_t s_nv[LK_TILE + LK_HALO];
    __shared__ uint16_t s_q[LK_TILE];
    __shared__ uint32_t s_wi[LK_TILE];
    __shared__ uint32_t s_sq[LK_TILE];
    __shared__ uint32_t s_warp_tot[LK_THREADS / 32];
    __shared__ uint32_t s_total;

    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t k = db.k, m = db.sig_len;
    const bool aligned16 = ((uintptr_t)bases & 15) == 0;
    // a caller may hand over a padded buffer: bases past the last sequence belong to no window
    const uint64_t n_bases = min(n_bases_arg, (uint64_t)__ldg((const unsigned long long *)seq_off + n_seq));

    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t p0 = tile * LK_TILE;
        // ---- phase A: stage bases -> 2-bit codes (128-bit coalesced loads for the tile body) ----
        if (aligned16 && p0 + LK_TILE <= n_bases) {
            if (tid < LK_TILE / 16) {
                const uint4 v = __ldg((const uint4 *)(bases + p0) + tid);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 4; j++)
                    o[j] = base_code(w[j] & 0xff) | (base_code((w[j] >> 8) & 0xff) << 8) |
                           (base_code((w[j] >> 16) & 0xff) << 16) | (base_code(w[j] >> 24) << 24);
                *((uint4 *)s_code + tid) = make_uint4(o[0], o[1], o[2], o[3]);
            }
            if (tid < LK_HALO) {
                const uint64_t g = p0 + LK_TILE + tid;
                s_code[LK_TILE + tid] = g < n_bases ? (uint8_t)base_code(__ldg(bases + g)) : (uint8_t)4;
            }
        } else {
            for (uint32_t q = tid; q < LK_TILE + LK_HALO; q += LK_THREADS) {
                const uint64_t g = p0 + q;
                s_code[q] = g < n_bases ? (uint8_t)base_code(__ldg(bases + g)) : (uint8_t)4;
            }
        }
        __syncthreads();
        // ---- phase B (KMC2): normalised m-mer value once per base position (mmer.h:117-124) ----
        if (db.is_kmc2) {
            for (uint32_t q = tid; q + m <= LK_TILE + LK_HALO; q += LK_THREADS) {
                uint32_t v = 0, bad = 0;
                for (uint32_t j = 0; j < m; j++) {
                    const uint32_t c = s_code[q + j];
                    bad |= c >> 2;
                    v = (v << 2) | (c & 3);
                }
                s_nv[q] = bad ? 0xFFFFFFFFu : __ldg(db.norm + v);
            }
        }
        // ---- phase C: which positions start a window?  compact them ----
        uint32_t my_wi[LK_PPT], my_sq[LK_PPT], my_mask = 0;  // statically indexed -> registers
        const uint32_t q0 = tid * LK_PPT;
        {
            uint64_t g = p0 + q0;
            if (g < n_bases) {
                // sequence containing base g: last s with seq_off[s] <= g   (seq_off[0] == 0)
                uint32_t lo = 0, hi = n_seq;  // invariant: seq_off[lo] <= g < seq_off[hi]
                while (hi - lo > 1) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (__ldg((const unsigned long long *)seq_off + mid) <= g) lo = mid; else hi = mid;
                }
                uint32_t s = lo;
                uint64_t sb = __ldg((const unsigned long long *)seq_off + s);
                uint64_t se = __ldg((const unsigned long long *)seq_off + s + 1);
                uint64_t wo = __ldg((const unsigned long long *)win_off + s);
                int last_bad = -1;  // largest tile-local index of a non-symbol among the codes scanned so far
                for (uint32_t j = 0; j + 1 < k; j++)
                    if (s_code[q0 + j] > 3) last_bad = (int)(q0 + j);
#pragma unroll
                for (uint32_t t = 0; t < LK_PPT; t++) {
                    const uint32_t q = q0 + t;
                    g = p0 + q;
                    if (s_code[q + k - 1] > 3) last_bad = (int)(q + k - 1);
                    if (g < n_bases) {
                        while (g >= se) {  // next sequence (skips empty ones)
                            s++;
                            sb = se;
                            se = __ldg((const unsigned long long *)seq_off + s + 1);
                            wo = __ldg((const unsigned long long *)win_off + s);
                        }
                        if (g + k <= se) {
                            const uint32_t wi = (uint32_t)(wo + (g - sb));
                            if (last_bad >= (int)q) {  // window touches a non-ACGT character: not found
                                if (ROUTE) route_owner[wi] = 0xFF;
                                else {
                                    if (counts) counts[wi] = 0;
                                    if (found) found[wi] = 0;
                                    if (cov) atomicMin((unsigned int *)&cov[s].first_missing, (unsigned int)(g - sb));
                                }
                            } else {
                                my_wi[t] = wi;
                                my_sq[t] = s;
                                my_mask |= 1u << t;
                            }
                        }
                    }
                }
            }
        }
        const uint32_t my_n = __popc(my_mask);
        uint32_t incl = my_n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += o;
        }
        if (lane == 31) s_warp_tot[wid] = incl;
        __syncthreads();  // also publishes s_nv
        if (tid == 0) {
            uint32_t acc = 0;
            for (int w = 0; w < LK_THREADS / 32; w++) {
                const uint32_t t = s_warp_tot[w];
                s_warp_tot[w] = acc;
                acc += t;
            }
            s_total = acc;
        }
        __syncthreads();
        {
            uint32_t o = s_warp_tot[wid] + incl - my_n;
#pragma unroll
            for (uint32_t t = 0; t < LK_PPT; t++)
                if (my_mask & (1u << t)) {
                    s_q[o] = (uint16_t)(q0 + t);
                    s_wi[o] = my_wi[t];
                    s_sq[o] = my_sq[t];
                    o++;
                }
        }
        __syncthreads();
        // ---- phase D: dense search phase ----
        const uint32_t n_valid = s_total;
        for (uint32_t base = 0; base < n_valid; base += LK_THREADS) {
            const uint32_t idx = base + tid;
            const bool live = idx < n_valid;
            uint32_t cnt = 0, sq = 0xFFFFFFFFu;
            bool ok = false;
            if (live) {
                const uint32_t q = s_q[idx];
                sq = s_sq[idx];
                uint64_t fwd = 0;
                for (uint32_t j = 0; j < k; j++) fwd = (fwd << 2) | s_code[q + j];
                uint32_t bin = 0;
                if (db.is_kmc2) {  // signature = min over the window's m-mers (kmer_api.h:653-672); strand-symmetric
                    uint32_t sig = 0xFFFFFFFFu;
                    for (uint32_t j = 0; j + m <= k; j++) sig = min(sig, s_nv[q + j]);
                    bin = __ldg(db.sigmap + sig);  // kmc_file.cpp:349-351
                }
                if (ROUTE) {  // the owner searches; FWD_THEN_RC is routed as the canonical key (equal on a both-strands DB)
                    const uint64_t rc = revcomp64(fwd, k);
                    const uint64_t key = mode == PF_LOOKUP_FWD ? fwd : (fwd < rc ? fwd : rc);
                    const uint32_t wi = s_wi[idx];
                    route_keys[wi] = key;
                    route_owner[wi] = (uint8_t)kmc_owner(db, key, bin);
                    continue;
                }
                if (mode == PF_LOOKUP_FWD) {
                    ok = kmc_search(db, fwd, bin, cnt);
                } else if (mode == PF_LOOKUP_FWD_THEN_RC) {  // CDBG.cpp:38-43
                    ok = kmc_search(db, fwd, bin, cnt);
                    if (!ok) ok = kmc_search(db, revcomp64(fwd, k), bin, cnt);
                } else {  // canonical key, kmc_file.cpp:1060 / :1290
                    const uint64_t rc = revcomp64(fwd, k);
                    ok = kmc_search(db, fwd < rc ? fwd : rc, bin, cnt);
                }
                if (!ok) cnt = 0;
                const uint32_t wi = s_wi[idx];
                if (counts) counts[wi] = cnt;
                if (found) found[wi] = ok ? 1 : 0;
            }
            if (!ROUTE && cov) {  // readCov reductions (CDBG.cpp:29-120), aggregated per (warp, sequence) segment
                const uint32_t grp = __match_any_sync(0xffffffffu, sq);
                const uint32_t leader = __ffs(grp) - 1;
                const uint32_t s_lo = __reduce_add_sync(grp, ok ? (cnt & 0xffffu) : 0u);
                const uint32_t s_hi = __reduce_add_sync(grp, ok ? (cnt >> 16) : 0u);
                const uint32_t mn = __reduce_min_sync(grp, ok ? cnt : 0xFFFFFFFFu);
                uint32_t fm = 0xFFFFFFFFu, fo = 0xFFFFFFFFu;
                if (live) {
                    // window index inside its sequence = wi - win_off[sq]
                    const uint32_t wloc = s_wi[idx] - (uint32_t)__ldg((const unsigned long long *)win_off + sq);
                    if (!ok) fm = wloc;
                    else if (!(cnt > low && cnt < up)) fo = wloc;
                }
                fm = __reduce_min_sync(grp, fm);
                fo = __reduce_min_sync(grp, fo);
                if (live && lane == leader) {
                    const uint64_t sum = (uint64_t)s_lo + ((uint64_t)s_hi << 16);
                    if (sum) atomicAdd((unsigned long long *)&cov[sq].sum, (unsigned long long)sum);
                    if (mn != 0xFFFFFFFFu) atomicMin(&cov[sq].min, mn);
                    if (fm != 0xFFFFFFFFu) atomicMin((unsigned int *)&cov[sq].first_missing, fm);
                    if (fo != 0xFFFFFFFFu) atomicMin((unsigned int *)&cov[sq].first_outside, fo);
                }
            }
        }
        __syncthreads();  // smem is reused by the next tile
    }
}

// window counts max(len - k + 1, 0) of every sequence (+ a 0 for the end of the exclusive scan that makes win_off)
__global__ void win_len_kernel(const uint64_t *__restrict__ seq_off, uint32_t n_seq, uint32_t k, uint64_t *__restrict__ wlen) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > n_seq) return;
    if (s == n_seq) { wlen[s] = 0; return; }
    const uint64_t len = seq_off[s + 1] - seq_off[s];
    wlen[s] = len >= k ? len - k + 1 : 0;
}

__global__ void cov_init_kernel(pf_cov_t *cov, const uint64_t *__restrict__ win_off, uint32_t n_seq) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seq) return;
    pf_cov_t c;
    c.sum = 0;
    c.min = 10000;  // CDBG.cpp:71
    c.n_kmers = (uint32_t)(win_off[s + 1] - win_off[s]);
    c.first_missing = -1;
    c.first_outside = -1;
    cov[s] = c;
}

// .kmc_suf records (S suffix bytes MSB-first + C counter bytes little-endian, kmc_file.cpp:1405-1452)
// -> one u64 per record, or suffix/counter arrays.
__global__ void repack_records_kernel(const uint8_t *__restrict__ raw, uint64_t n, uint32_t S, uint32_t C, int packed,
                                      uint64_t *__restrict__ rec, uint64_t *__restrict__ suf, uint32_t *__restrict__ cnt) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *r = raw + i * (S + C);
    uint64_t s = 0, c = 0;
    for (uint32_t a = 0; a < S; a++) s = (s << 8) | r[a];
    for (uint32_t b = 0; b < C; b++) c |= (uint64_t)r[S + b] << (8 * b);
    if (packed) rec[i] = (s << (8 * C)) | c;
    else { suf[i] = s; cnt[i] = (uint32_t)c; }
}

// ---- one-sector hash index (pf_kmc_hash.cuh): build + verification, one thread per record ------------------------------
// Every record is re-keyed (prefix from its position in the prefix table, suffix from the record) and inserted; on the
// way the kernel checks that the reference's own search (BinarySearch, kmc_file.cpp:1383) would find this record:
//   bit 0  suffixes not strictly ascending inside a prefix bucket      -> binary search outcome undefined
//   bit 1  KMC2: the record is not in the bin its signature maps to     -> CheckKmer (:349-354) looks elsewhere
//   bit 3  a key would sit more than H_MAX_DIST buckets from home       -> table too dense
// (any of these keeps the verbatim index) and whether every key is the canonical form of its k-mer:
//   bit 2  key > reverse complement                                      -> FWD_THEN_RC stays a two-probe lookup
enum { HB_UNSORTED = 1, HB_WRONG_BIN = 2, HB_NOT_CANONICAL = 4, HB_OVERFLOW = 8 };
// part / n_parts: a partition of the index inserts only the keys it owns (owner = mix(key) % n_parts, pfkmc::hash_owner); every
// record is still verified.  status[1] (as u64 at status + 2) counts the inserted keys.
// Chunked form: the thread block range covers records [i0, i0 + n); db.rec / db.suf / db.cnt hold the records from index `base`
// on (base = i0 - 1 when a carried predecessor record precedes the chunk, else i0); the whole image is (i0, n, base) = (0, N, 0).
__global__ void kmc_hash_build_kernel(const KmcView db, const pfkmc::HashView hv, uint32_t *__restrict__ status, uint32_t part,
                                      uint32_t n_parts, uint64_t i0, uint64_t n, uint64_t base) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint64_t i = i0 + t;
    if (i >= db.N) return;
    if (lut_at(db, 0) > i) return;                     // before the first bucket: no prefix reaches it
    uint64_t lo = 0, hi = db.lut_n - 1;               // lut[lut_n-1] = N+1 > i
    while (hi - lo > 1) {                              // last slot with lut[slot] <= i  (empty slots share a value)
        const uint64_t mid = (lo + hi) >> 1;
        if (lut_at(db, mid) <= i) lo = mid; else hi = mid;
    }
    const uint64_t slot = lo;
    if (!db.is_kmc2 && slot >= db.single_lut) return;  // KMC1: entries past 4^p are never indexed (:353)
    const uint64_t prefix = db.is_kmc2 ? slot % db.single_lut : slot;
    const uint32_t bin = db.is_kmc2 ? (uint32_t)(slot / db.single_lut) : 0u;
    const uint32_t cbits = 8 * db.C, sbits = 2 * (db.k - db.p);
    uint64_t rs, c, prev = 0;
    const bool has_prev = i > lut_at(db, slot);
    if (db.packed) {
        const uint64_t r = db.rec[i - base];
        rs = r >> cbits; c = r & ((1ull << cbits) - 1);
        if (has_prev) prev = db.rec[i - 1 - base] >> cbits;
    } else {
        rs = db.suf[i - base]; c = db.cnt[i - base];
        if (has_prev) prev = db.suf[i - 1 - base];
    }
    uint32_t st = 0;
    if (has_prev && !(prev < rs)) st |= HB_UNSORTED;
    const uint64_t key = (prefix << sbits) | rs;
    if (db.is_kmc2) {                                  // kmer_api.h:653-672 on the packed key
        const uint32_t m = db.sig_len;
        const uint64_t mask = (1ull << (2 * m)) - 1;
        uint32_t sig = 0xFFFFFFFFu;
        for (uint32_t j = 0; j + m <= db.k; j++) sig = min(sig, __ldg(db.norm + ((key >> (2 * (db.k - m - j))) & mask)));
        if (__ldg(db.sigmap + sig) != bin) st |= HB_WRONG_BIN;
    }
    if (key > revcomp64(key, db.k)) st |= HB_NOT_CANONICAL;
    if (n_parts <= 1 || pfkmc::hash_owner(key, hv.kbits, n_parts) == part) {
        if (!pfkmc::hash_insert(hv, key, c)) st |= HB_OVERFLOW;
        else atomicAdd((unsigned long long *)(status + 2), 1ull);
    }
    if (st) atomicOr(status, st);
}

// ---- partitioned database: the owner's side and the way back ---------------------------------------------------
// One thread per routed key: signature -> bin (KMC2) -> local LUT -> search.
__global__ void kmc_lookup_keys_kernel(const KmcView db, const unsigned long long *__restrict__ keys, const uint64_t n,
                                       uint32_t *__restrict__ counts, uint8_t *__restrict__ found) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t key = keys[i];
    uint32_t bin = 0;
    if (db.is_kmc2) {  // kmer_api.h:653-672 on the packed key
        const uint32_t m = db.sig_len;
        const uint64_t mask = (1ull << (2 * m)) - 1;
        uint32_t sig = 0xFFFFFFFFu;
        for (uint32_t j = 0; j + m <= db.k; j++) sig = min(sig, __ldg(db.norm + ((key >> (2 * (db.k - m - j))) & mask)));
        bin = __ldg(db.sigmap + sig);
    }
    uint32_t c = 0;
    const bool ok = kmc_search(db, key, bin, c);
    counts[i] = ok ? c : 0;
    found[i] = ok ? 1 : 0;
}

// send_keys[t] = keys[idx[t]] for the t-th window in owner order
__global__ void kmc_gather_keys_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ idx, uint64_t n,
                                       unsigned long long *__restrict__ out) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) out[t] = keys[idx[t]];
}

// bounds[o] = first position of owner o in the sorted owner list, o = 0 .. n_parts (0xFF entries sort last)
__global__ void kmc_owner_bounds_kernel(const uint8_t *__restrict__ owners, uint64_t n, uint32_t n_parts, uint64_t *bounds) {
    const uint32_t o = threadIdx.x;
    if (o > n_parts) return;
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (owners[mid] < o) lo = mid + 1; else hi = mid;
    }
    bounds[o] = lo;
}

// replies (in send order) -> per-window outputs
__global__ void kmc_scatter_kernel(const uint32_t *__restrict__ idx, uint64_t n, const uint32_t *__restrict__ r_counts,
                                   const uint8_t *__restrict__ r_found, uint32_t *__restrict__ counts, uint8_t *__restrict__ found) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint32_t wi = idx[t];
    counts[wi] = r_counts[t];
    found[wi] = r_found[t];
}

// readCov reductions (CDBG.cpp:29-120) from per-window results, one thread per sequence
__global__ void kmc_cov_from_counts_kernel(const uint64_t *__restrict__ win_off, uint32_t n_seq, const uint32_t *__restrict__ counts,
                                           const uint8_t *__restrict__ found, uint32_t low, uint32_t up, pf_cov_t *__restrict__ cov) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seq) return;
    const uint64_t w0 = win_off[s], w1 = win_off[s + 1];
    pf_cov_t c;
    c.sum = 0; c.min = 10000; c.n_kmers = (uint32_t)(w1 - w0); c.first_missing = -1; c.first_outside = -1;
    for (uint64_t w = w0; w < w1; w++) {
        if (!found[w]) { if (c.first_missing < 0) c.first_missing = (int32_t)(w - w0); continue; }
        const uint32_t v = counts[w];
        c.sum += v;
        if (v < c.min) c.min = v;
        if (!(v > low && v < up) && c.first_outside < 0) c.first_outside = (int32_t)(w - w0);
    }
    cov[s] = c;
}

// ---- lookup phase B: site k-mers of branching bubbles --------------------------------------------------------------------
// CDBG.cpp:2295-2509.  For every variable column of an aligned bubble the reference builds, per row, the
// k-mer that ENDS at the site (SNP) or ends where the rows start to differ after the gap (indel), de-duplicates the strings per
// allele class in a std::set, looks each distinct string up with readCov(string, lower, upper) -- 'as written, else reverse
// complement', strict gate -- and sums per class; a count outside the gate drops the site, a missing k-mer ends the program.
// One thread per SITE (site_map_kernel first writes the bubble of every variable column): the number of indel sites before a
// site -- the only thing that couples the sites of a bubble -- is a count over var_kind.  Everything else is a few dozen byte
// loads per row.  Row r of bubble b is rows[rows_off[b] + r * aln_len[b] ...].
constexpr uint32_t SITE_MAX_ROWS = 16;

__global__ void site_map_kernel(const uint64_t *__restrict__ var_off, uint32_t n, uint32_t *__restrict__ site_bubble) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    for (uint64_t v = var_off[b]; v < var_off[b + 1]; v++) site_bubble[v] = b;
}

struct SiteArgs {
    KmcView db;
    pfkmc::HashView hv;
    int hash_on, both_strands;
    uint32_t n;
    const int32_t *status;
    const uint32_t *n_rows, *aln_len;
    const uint64_t *rows_off;
    const char *rows;
    const uint64_t *var_off;
    const uint32_t *var_col;
    const uint8_t *var_kind;
    const uint64_t *cls_off;
    const uint16_t *cls;
    const uint8_t *skip;
    uint32_t low, up;
    uint8_t *site_status, *site_ncls;
    unsigned long long *site_cov;
    const uint32_t *site_bubble;
    uint64_t n_sites;
    const unsigned long long *peer_tab[pfkmc::PF_MAX_PEERS];   // peer-memory form of a partitioned index (else unused)
    uint32_t peer_lookup, n_parts;
};

__device__ __forceinline__ bool site_lookup_one(const SiteArgs &a, uint64_t key, uint32_t &cnt) {
    if (a.hash_on) {
        const unsigned long long *tab = a.peer_lookup ? a.peer_tab[pfkmc::hash_owner(key, a.hv.kbits, a.n_parts)] : a.hv.tab;
        const bool ok = pfkmc::hash_find(a.hv, tab, key, cnt);
        return ok && cnt >= a.db.min_count && (uint64_t)cnt <= a.db.max_count;
    }
    uint32_t bin = 0;
    if (a.db.is_kmc2) {
        const uint32_t m = a.db.sig_len;
        const uint64_t mask = (1ull << (2 * m)) - 1;
        uint32_t sig = 0xFFFFFFFFu;
        for (uint32_t j = 0; j + m <= a.db.k; j++) sig = min(sig, __ldg(a.db.norm + ((key >> (2 * (a.db.k - m - j))) & mask)));
        bin = __ldg(a.db.sigmap + sig);
    }
    return kmc_search(a.db, key, bin, cnt);
}

// One row's site k-mer.  `end` = one past the last column that belongs to the left part, `need` = characters wanted from the
// left, `tail` / `tail_len` = the characters already fixed after them (the indel extension), `fwd` = where to go on reading if
// the row's left part is too short.  contiguous: the left part is the `need` columns before `end` exactly as they are
// (no indel site so far, :2358-2365, :2469-2472); otherwise gaps are skipped (:2366-2388, :2433-2465).
// Returns 0 ok, PF_SITE_UNDEFINED when the reference itself would read outside the row.
__device__ __forceinline__ int site_row_kmer(const char *row, uint32_t L, uint32_t k, uint32_t end, uint32_t need, uint64_t tail,
                                             uint32_t tail_len, uint32_t fwd, bool contiguous, uint64_t &key) {
    uint64_t back = 0;
    uint32_t got = 0;
    if (contiguous) {
        if (end < need) return PF_SITE_UNDEFINED;                // substr with a wrapped start position throws
        for (uint32_t q = 0; q < need; q++) {
            const uint32_t code = base_code((uint8_t)row[end - 1 - q]);
            if (code > 3) return PF_SITE_UNDEFINED;              // a '-' inside the window: not a k-mer, outcome order dependent
            back |= (uint64_t)code << (2 * q);
        }
        got = need;
    } else {
        for (uint32_t j = end; j > 0 && got < need; j--) {
            const uint8_t ch = (uint8_t)row[j - 1];
            if (ch == '-') continue;
            const uint32_t code = base_code(ch);
            if (code > 3) return PF_SITE_UNDEFINED;
            back |= (uint64_t)code << (2 * got);
            got++;
        }
    }
    key = tail_len ? (((tail_len >= 32 ? 0ull : back << (2 * tail_len))) | tail) : back;
    uint32_t have = got + tail_len;
    for (uint32_t x = fwd; have < k; x++) {                       // the left part was shorter than wanted: extend to the right
        if (x >= L) return PF_SITE_UNDEFINED;                    // the reference never terminates / throws here
        const uint8_t ch = (uint8_t)row[x];
        if (ch == '-') continue;
        const uint32_t code = base_code(ch);
        if (code > 3) return PF_SITE_UNDEFINED;
        key = (key << 2) | code;
        have++;
    }
    return PF_SITE_OK;
}

// The site k-mer of every row at variable column v of bubble b (CDBG.cpp:2338-2388 indel sites, :2433-2472 SNP sites).
// Returns PF_SITE_OK or PF_SITE_UNDEFINED (the reference itself would read outside the row / a non-ACGT character).
__device__ __forceinline__ int site_build_keys(const SiteArgs &a, uint32_t b, uint64_t v, uint32_t n_ind, uint64_t key[SITE_MAX_ROWS]) {
    const uint32_t nr = a.n_rows[b], L = a.aln_len[b], k = a.db.k;
    const char *R = a.rows + a.rows_off[b];
    const uint32_t c = a.var_col[v];
    const bool is_ind = a.var_kind[v] == 1;
    int st = PF_SITE_OK;
    if (is_ind) {
        uint32_t cur[SITE_MAX_ROWS];
        uint64_t ext[SITE_MAX_ROWS];
        for (uint32_t r = 0; r < nr; r++) { cur[r] = c; ext[r] = 0; }
        uint32_t e = 0;
        for (;;) {                                            // :2338-2357: one more base per row until the rows differ
            bool differ = false;
            uint32_t first = 0;
            for (uint32_t r = 0; r < nr && st == PF_SITE_OK; r++) {
                const char *row = R + (uint64_t)r * L;
                while (cur[r] < L && row[cur[r]] == '-') cur[r]++;
                if (cur[r] >= L) { st = PF_SITE_UNDEFINED; break; }
                const uint32_t code = base_code((uint8_t)row[cur[r]]);
                if (code > 3) { st = PF_SITE_UNDEFINED; break; }
                cur[r]++;
                ext[r] = (ext[r] << 2) | code;
                if (r == 0) first = code; else differ |= code != first;
            }
            if (st != PF_SITE_OK) break;
            e++;
            if (differ) break;
            if (e >= k) { st = PF_SITE_UNDEFINED; break; }
        }
        for (uint32_t r = 0; r < nr && st == PF_SITE_OK; r++)
            st = site_row_kmer(R + (uint64_t)r * L, L, k, c, k - e, ext[r], e, cur[r], n_ind == 0, key[r]);
    } else {
        for (uint32_t r = 0; r < nr && st == PF_SITE_OK; r++)
            st = site_row_kmer(R + (uint64_t)r * L, L, k, c + 1, k, 0, 0, c + 1, n_ind == 0, key[r]);
    }
    return st;
}

__global__ void site_cov_kernel(const SiteArgs a) {
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= a.n_sites) return;
    const uint32_t b = a.site_bubble[v];
    const uint64_t v0 = a.var_off[b];
    const uint32_t nr = a.n_rows[b], k = a.db.k;
    const uint16_t *C = a.cls + a.cls_off[b];
    unsigned long long *cov_out = a.site_cov + a.cls_off[b];
    const bool skipped = a.skip && a.skip[b];
    uint32_t n_ind = 0;
    if (!skipped)
        for (uint64_t u = v0; u < v; u++) n_ind += a.var_kind[u] == 1;     // indel sites before this one (:2390)
    {
        const uint16_t *cl = C + (v - v0) * nr;
        unsigned long long *cov = cov_out + (v - v0) * nr;
        uint32_t ncls = 0;
        for (uint32_t r = 0; r < nr; r++) { ncls = max(ncls, (uint32_t)cl[r]); cov[r] = 0; }
        a.site_ncls[v] = (uint8_t)min(ncls, 255u);
        int st = PF_SITE_OK;
        uint64_t key[SITE_MAX_ROWS];
        if (skipped) st = PF_SITE_SKIPPED;
        else if (nr > SITE_MAX_ROWS || k > 32) st = PF_SITE_UNDEFINED;
        else if (!a.both_strands) st = PF_SITE_OK;               // readCov(string) does nothing on a strand-specific database (CDBG.cpp:34)
        else st = site_build_keys(a, b, v, n_ind, key);
        if (st == PF_SITE_OK && a.both_strands) {
            for (uint32_t q = 1; q <= ncls && st == PF_SITE_OK; q++) {          // classes in order, strings in std::set order
                unsigned long long acc = 0;
                bool have_last = false;
                uint64_t last = 0;
                for (;;) {
                    bool found = false;
                    uint64_t best = 0;
                    for (uint32_t r = 0; r < nr; r++)
                        if (cl[r] == q && (!have_last || key[r] > last) && (!found || key[r] < best)) { best = key[r]; found = true; }
                    if (!found) break;
                    last = best; have_last = true;
                    uint32_t cnt = 0;
                    bool ok = site_lookup_one(a, best, cnt);                       // CDBG.cpp:38-43
                    if (!ok) ok = site_lookup_one(a, revcomp64(best, k), cnt);
                    if (!ok) { st = PF_SITE_MISSING; break; }                      // the reference exits (CDBG.cpp:52-56)
                    if (!(cnt > a.low && cnt < a.up)) { st = PF_SITE_DROPPED; break; }
                    acc += cnt;
                }
                cov[q - 1] = acc;
            }
        }
        a.site_status[v] = (uint8_t)st;
    }
}

// The site k-mers themselves (coloured graphs, CCDBG.cpp:1057-1376: the caller needs the strings -- it asks the graph which colours
// hold each one before any database is read): key of every (variable column, row) as a right-aligned 2-bit k-mer, one status per column.
__global__ void site_keys_kernel(const SiteArgs a, unsigned long long *__restrict__ keys) {
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= a.n_sites) return;
    const uint32_t b = a.site_bubble[v];
    const uint64_t v0 = a.var_off[b];
    const uint32_t nr = a.n_rows[b];
    unsigned long long *out = keys + a.cls_off[b] + (v - v0) * nr;
    const bool skipped = a.skip && a.skip[b];
    int st;
    uint64_t key[SITE_MAX_ROWS];
    if (skipped) st = PF_SITE_SKIPPED;
    else if (nr > SITE_MAX_ROWS || a.db.k > 32) st = PF_SITE_UNDEFINED;
    else {
        uint32_t n_ind = 0;
        for (uint64_t u = v0; u < v; u++) n_ind += a.var_kind[u] == 1;
        st = site_build_keys(a, b, v, n_ind, key);
    }
    for (uint32_t r = 0; r < nr; r++) out[r] = st == PF_SITE_OK ? key[r] : 0ull;
    a.site_status[v] = (uint8_t)st;
}

bool slurp(const std::string &path, std::vector<unsigned char> &buf) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    if (fseek(f, 0, SEEK_END) != 0) { fclose(f); return false; }
    const long long sz = ftell(f);
    rewind(f);
    if (sz < 0) { fclose(f); return false; }
    buf.resize((size_t)sz);
    const size_t got = sz ? fread(buf.data(), 1, (size_t)sz, f) : 0;
    fclose(f);
    return got == (size_t)sz;
}

inline uint32_t le32(const unsigned char *p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint64_t le64(const unsigned char *p) { uint64_t v; memcpy(&v, p, 8); return v; }

// mmer.h:34-57
bool sig_allowed(uint32_t mm, uint32_t len) {
    if ((mm & 0x3f) == 0x3f || (mm & 0x3f) == 0x3b || (mm & 0x3c) == 0x3c) return false;
    for (uint32_t j = 0; j + 3 < len; ++j) {
        if ((mm & 0xf) == 0) return false;
        mm >>= 2;
    }
    return !(mm == 0 || mm == 0x04 || (mm & 0xf) == 0);
}

}  // namespace

struct pf_kmc {
    pf_ctx *ctx = nullptr;
    pf_kmc_info_t info{};
    uint32_t orig_min = 0;
    uint64_t orig_max = 0;
    KmcView view{};
    void *d_lut = nullptr, *d_sigmap = nullptr, *d_norm = nullptr, *d_rec = nullptr, *d_suf = nullptr, *d_cnt = nullptr;
    uint64_t device_bytes = 0;
    uint64_t local_kmers = 0;
    struct pf_kmc_route_state *route = nullptr;
    // one-sector hash index (pf_kmc_hash.cuh); when active it replaces lut/rec/suf/cnt
    bool hash_on = false, canonical_ok = false;
    pfkmc::HashView hview{};
    void *d_hash = nullptr;
    uint32_t build_status = 0;
    uint64_t hash_inserted = 0;
    // peer-memory form of a partitioned index: the slices of the other partitions mapped through CUDA IPC
    const unsigned long long *peer_tab[pfkmc::PF_MAX_PEERS] = {nullptr};
    void *peer_mapped[pfkmc::PF_MAX_PEERS] = {nullptr};
    bool peers_attached = false;   // keys in the hash index (== total_kmers unless this handle is one partition)
    bool borrowed = false;         // pf_kmc_share: the index memory belongs to another handle, only the per-call buffers are ours
    pf::DevBuf tile_seq;   // per-call scratch of the hash lookup (grow-only)
    pf::DevBuf site_status, site_ncls, site_cov, site_skip, site_map;   // pf_site_cov outputs (grow-only)
    pf::PinnedBuf h_site[5];
    uint64_t site_totals[2] = {0, 0};
    // host-pointer calls: the handle's own stream, staging and device buffers
    cudaStream_t k_stream = nullptr;
    bool k_stream_borrowed = false;
    cudaEvent_t k_staged_ev = nullptr;   // recorded when the sequences of the last coverage-only host call are on the device
    uint32_t k_staged_seq = 0;
    uint64_t k_staged_bases = 0;   // the context's partition stream (pf_lookup_partition): not ours to destroy
    pf::PinnedBuf k_stage;
    pf::DevBuf k_in[3], k_out[3];   // variable columns, class entries of the last pf_site_cov*
};

struct pf_kmc_route_state {   // scratch of pf_kmc_route_dev (grow-only)
    pf::DevBuf keys, owner, owner_sorted, idx, idx_sorted, bounds, cub_tmp;
    pf::PinnedBuf h_bounds;
};

namespace {

// The hash index built WITHOUT staging the database in HBM: the records are streamed from the file through two pinned buffers
// (64 MB of records each), re-packed, verified and inserted chunk by chunk -- the predecessor of a chunk's first record is
// carried along for the ascending-suffix check.  Peak HBM = the table + the prefix table + two chunk buffers, so a partition
// (part / n_parts: only the keys with mix(key) % n_parts == part are inserted) opens a database that is larger than one GPU's
// memory, and no rank holds the file in host memory.  Leaves db->hash_on false (and everything released) when the database
// fails the verification or the slot does not fit 63 bits: the caller falls back to the verbatim image.
int kmc_stream_hash(pf_kmc *db, const std::vector<uint64_t> &lut, const std::vector<uint32_t> &sigmap, const std::vector<uint32_t> &norm,
                    FILE *suf_f, const char *prefix, uint32_t part, uint32_t n_parts) {
    pf_ctx *ctx = db->ctx;
    const pf_kmc_info_t &I = db->info;
    const uint32_t k = I.kmer_length, p = I.lut_prefix_length, C = I.counter_size, S = (k - p) / 4;
    const uint64_t N = I.total_kmers, R = S + C;
    if (!N) return PF_OK;
    KmcView V{};
    V.k = k; V.p = p; V.S = S; V.C = C; V.sig_len = I.signature_len; V.is_kmc2 = I.kmc_version == 0x200;
    V.min_count = I.min_count; V.max_count = I.max_count; V.N = N; V.single_lut = 1ull << (2 * p); V.lut_n = lut.size();
    V.lut64 = (N + 1 >= (1ull << 32)) ? 1 : 0;
    V.packed = R <= 8 ? 1 : 0;
    V.n_parts = 1; V.part = 0; V.prefix_lo = 0; V.prefix_cnt = V.single_lut; V.prefix_per_part = V.single_lut;
    const uint64_t n_local = n_parts > 1 ? N / n_parts + N / (4 * n_parts) + 1024 : N;
    uint32_t b = 4;
    while (b < 40 && (3ull << b) < 2 * n_local) b++;                        // <= 1.5 keys per 4-slot bucket on average
    const int need = (int)(2 * k + 8 * C + pfkmc::H_DIST_BITS) - 63;        // dist + rem + counter must leave the all-ones slot free
    if (need > (int)b) {
        if (need > 23 || need > 2 * (int)k) return PF_OK;
        b = (uint32_t)need;
    }
    if (b > 2 * k) b = 2 * k;
    cudaStream_t st = ctx->stream;
    void *d_lut = nullptr, *d_sigmap = nullptr, *d_norm = nullptr, *d_status = nullptr, *tab = nullptr;
    void *d_raw[2] = {nullptr, nullptr}, *d_rec = nullptr, *d_suf = nullptr, *d_cnt = nullptr;
    pf::PinnedBuf stage[2];
    cudaEvent_t done[2] = {nullptr, nullptr};
    // records per chunk: 64 MB of the file; PF_OPEN_CHUNK_RECORDS overrides it (tests drive the chunk seams with tiny chunks)
    const char *ch_env = getenv("PF_OPEN_CHUNK_RECORDS");
    const uint64_t CH_REC = ch_env && atoll(ch_env) > 0 ? (uint64_t)atoll(ch_env) : std::max<uint64_t>(1, (64ull << 20) / R);
    auto cleanup = [&](bool keep_index) {
        for (int i = 0; i < 2; i++) { if (done[i]) cudaEventDestroy(done[i]); stage[i].release(); cudaFree(d_raw[i]); }
        cudaFree(d_rec); cudaFree(d_suf); cudaFree(d_cnt); cudaFree(d_status); cudaFree(d_lut);
        if (!keep_index) { cudaFree(tab); cudaFree(d_sigmap); cudaFree(d_norm); }
    };
#define PF_TRY_CLEAN(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { pf::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); cleanup(false); return PF_E_CUDA; } } while (0)
    if (V.lut64) {
        PF_TRY_CLEAN(cudaMalloc(&d_lut, lut.size() * 8));
        PF_TRY_CLEAN(cudaMemcpyAsync(d_lut, lut.data(), lut.size() * 8, cudaMemcpyHostToDevice, st));
        PF_TRY_CLEAN(pf::stream_sync(st));
    } else {
        std::vector<uint32_t> l32(lut.size());
        for (size_t i = 0; i < lut.size(); i++) l32[i] = (uint32_t)lut[i];
        PF_TRY_CLEAN(cudaMalloc(&d_lut, l32.size() * 4));
        PF_TRY_CLEAN(cudaMemcpyAsync(d_lut, l32.data(), l32.size() * 4, cudaMemcpyHostToDevice, st));
        PF_TRY_CLEAN(pf::stream_sync(st));
    }
    V.lut = d_lut;
    if (V.is_kmc2) {
        PF_TRY_CLEAN(cudaMalloc(&d_sigmap, sigmap.size() * 4));
        PF_TRY_CLEAN(cudaMemcpyAsync(d_sigmap, sigmap.data(), sigmap.size() * 4, cudaMemcpyHostToDevice, st));
        PF_TRY_CLEAN(cudaMalloc(&d_norm, norm.size() * 4));
        PF_TRY_CLEAN(cudaMemcpyAsync(d_norm, norm.data(), norm.size() * 4, cudaMemcpyHostToDevice, st));
        PF_TRY_CLEAN(pf::stream_sync(st));
        V.sigmap = (const uint32_t *)d_sigmap; V.norm = (const uint32_t *)d_norm;
    }
    for (int i = 0; i < 2; i++) {
        if (stage[i].reserve(CH_REC * R)) { cleanup(false); return PF_E_NOMEM; }
        PF_TRY_CLEAN(cudaMalloc(&d_raw[i], CH_REC * R));
        PF_TRY_CLEAN(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
    }
    if (V.packed) PF_TRY_CLEAN(cudaMalloc(&d_rec, (CH_REC + 1) * 8));
    else { PF_TRY_CLEAN(cudaMalloc(&d_suf, (CH_REC + 1) * 8)); PF_TRY_CLEAN(cudaMalloc(&d_cnt, (CH_REC + 1) * 4)); }
    PF_TRY_CLEAN(cudaMalloc(&d_status, 16));
    V.rec = (const uint64_t *)d_rec; V.suf = (const uint64_t *)d_suf; V.cnt = (const uint32_t *)d_cnt;
    pfkmc::HashView hv{};
    uint64_t bytes = 0;
    uint32_t h_status[4] = {0, 0, 0, 0};
    for (int attempt = 0;; attempt++) {   // a table that overflows (a key more than H_MAX_DIST buckets from home) is rebuilt once, twice as large
        if (b > 2 * k || (2 * k - b) + 8 * C + pfkmc::H_DIST_BITS > 63) { cleanup(false); return PF_OK; }
        hv.bucket_bits = b; hv.rem_bits = 2 * k - b; hv.cbits = 8 * C; hv.kbits = 2 * k;
        bytes = 32ull << b;
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || free_b < bytes + (256ull << 20)) {
            cudaGetLastError();
            pf::set_error("%s: the hash index needs %.1f GB of device memory, %.1f GB are free (partition the index: pf_kmc_open_part)", prefix,
                          bytes / 1e9, free_b / 1e9);
            cleanup(false);
            return PF_E_NOMEM;
        }
        PF_TRY_CLEAN(cudaMalloc(&tab, bytes));
        hv.tab = (unsigned long long *)tab;
        PF_TRY_CLEAN(cudaMemsetAsync(d_status, 0, 16, st));
        PF_TRY_CLEAN(cudaMemsetAsync(tab, 0xFF, bytes, st));
        if (fseeko(suf_f, 4, SEEK_SET) != 0) { pf::set_error("%s.kmc_suf: seek failed", prefix); cleanup(false); return PF_E_IO; }
        int slot = 0;
        for (uint64_t i0 = 0; i0 < N; i0 += CH_REC) {
            const uint64_t n = std::min<uint64_t>(CH_REC, N - i0);
            cudaEventSynchronize(done[slot]);                              // the copy that last used this staging buffer is over
            if (fread(stage[slot].p, 1, n * R, suf_f) != n * R) { pf::set_error("%s.kmc_suf: short read", prefix); cleanup(false); return PF_E_IO; }
            PF_TRY_CLEAN(cudaMemcpyAsync(d_raw[slot], stage[slot].p, n * R, cudaMemcpyHostToDevice, st));
            PF_TRY_CLEAN(cudaEventRecord(done[slot], st));
            // slot 0 of the chunk arrays holds the predecessor of the chunk's first record (carried over on the device)
            if (i0) {
                if (V.packed) PF_TRY_CLEAN(cudaMemcpyAsync(d_rec, (uint64_t *)d_rec + CH_REC, 8, cudaMemcpyDeviceToDevice, st));
                else {
                    PF_TRY_CLEAN(cudaMemcpyAsync(d_suf, (uint64_t *)d_suf + CH_REC, 8, cudaMemcpyDeviceToDevice, st));
                    PF_TRY_CLEAN(cudaMemcpyAsync(d_cnt, (uint32_t *)d_cnt + CH_REC, 4, cudaMemcpyDeviceToDevice, st));
                }
            }
            repack_records_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const uint8_t *)d_raw[slot], n, S, C, (int)V.packed,
                                                                              V.packed ? (uint64_t *)d_rec + 1 : nullptr,
                                                                              V.packed ? nullptr : (uint64_t *)d_suf + 1, V.packed ? nullptr : (uint32_t *)d_cnt + 1);
            kmc_hash_build_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(V, hv, (uint32_t *)d_status, part, n_parts, i0, n, i0 - 1);
            ctx->launches += 2;
            slot ^= 1;
        }
        PF_TRY_CLEAN(cudaGetLastError());
        PF_TRY_CLEAN(cudaMemcpyAsync(h_status, d_status, 16, cudaMemcpyDeviceToHost, st));
        PF_TRY_CLEAN(pf::stream_sync(st));
        db->build_status = h_status[0];
        if (!(h_status[0] & (HB_UNSORTED | HB_WRONG_BIN | HB_OVERFLOW))) break;
        cudaFree(tab);
        tab = nullptr;
        if ((h_status[0] & (HB_UNSORTED | HB_WRONG_BIN)) || attempt == 1) { cleanup(false); return PF_OK; }
        b++;
    }
#undef PF_TRY_CLEAN
    db->hash_inserted = (uint64_t)h_status[2] | ((uint64_t)h_status[3] << 32);
    db->hash_on = true;
    db->canonical_ok = !(h_status[0] & HB_NOT_CANONICAL);
    db->hview = hv;
    db->d_hash = tab;
    db->d_sigmap = d_sigmap; db->d_norm = d_norm;
    V.lut = nullptr; V.rec = nullptr; V.suf = nullptr; V.cnt = nullptr;
    V.n_parts = n_parts; V.part = part;
    db->view = V;
    db->orig_min = I.min_count; db->orig_max = I.max_count;
    db->device_bytes = bytes + (V.is_kmc2 ? ((1ull << (2 * V.sig_len)) * 8 + 4) : 0);
    db->local_kmers = db->hash_inserted;
    cleanup(true);
    return PF_OK;
}

int kmc_open_impl(pf_ctx *ctx, const char *prefix, uint32_t part, uint32_t n_parts, uint32_t flags, pf_kmc **out);

int kmc_open_impl(pf_ctx *ctx, const char *prefix, uint32_t part, uint32_t n_parts, uint32_t flags, pf_kmc **out) {
    if (!ctx || !prefix || !out) { pf::set_error("pf_kmc_open: null argument"); return PF_E_INVALID; }
    *out = nullptr;
    if (n_parts == 0 || part >= n_parts || n_parts > 254) { pf::set_error("pf_kmc_open: bad partition %u of %u", part, n_parts); return PF_E_INVALID; }
    PF_CUDA_TRY(cudaSetDevice(ctx->device));
    std::vector<unsigned char> pre;
    const std::string base(prefix);
    if (!slurp(base + ".kmc_pre", pre)) { pf::set_error("cannot read %s.kmc_pre", prefix); return PF_E_IO; }
    const size_t fs = pre.size();
    if (fs < 28 || memcmp(&pre[0], "KMCP", 4) || memcmp(&pre[fs - 4], "KMCP", 4)) {
        pf::set_error("%s.kmc_pre: bad KMCP markers", prefix);
        return PF_E_IO;
    }
    std::unique_ptr<pf_kmc> db(new pf_kmc());
    db->ctx = ctx;
    pf_kmc_info_t &I = db->info;
    I.kmc_version = le32(&pre[fs - 12]);          // kmc_file.cpp:188-192
    const uint64_t hoff = pre[fs - 8];            // one byte (:200, :257)
    std::vector<uint64_t> lut;
    std::vector<uint32_t> sigmap, norm;
    if (I.kmc_version == 0x200) {                 // :196-245
        if (fs < hoff + 12 || hoff < 37) { pf::set_error("%s.kmc_pre: truncated KMC2 header", prefix); return PF_E_IO; }
        const unsigned char *h = &pre[fs - 8 - hoff];
        I.kmer_length = le32(h); I.mode = le32(h + 4); I.counter_size = le32(h + 8);
        I.lut_prefix_length = le32(h + 12); I.signature_len = le32(h + 16); I.min_count = le32(h + 20);
        I.max_count = le32(h + 24); I.total_kmers = le64(h + 28); I.both_strands = h[36] ? 0 : 1;
        if (I.signature_len < 5 || I.signature_len > 11) {
            pf::set_error("%s.kmc_pre: signature length %u outside 5..11 (mmer.cpp:27-58)", prefix, I.signature_len);
            return PF_E_UNSUPPORTED;
        }
        const uint64_t sig_n = (1ull << (2 * I.signature_len)) + 1;
        const uint64_t body = fs - 12;
        if (body < sig_n * 4 + hoff + 8 + 8) { pf::set_error("%s.kmc_pre: file too small", prefix); return PF_E_IO; }
        const uint64_t lut_n = (body - (sig_n * 4 + hoff + 8)) / 8;  // index of the guard word (:224-233)
        lut.resize(lut_n + 1);
        memcpy(lut.data(), &pre[4], (lut_n + 1) * 8);
        lut[lut_n] = I.total_kmers + 1;
        sigmap.resize(sig_n);
        memcpy(sigmap.data(), &pre[4 + (lut_n + 1) * 8], sig_n * 4);
        const uint32_t special = 1u << (2 * I.signature_len);
        norm.resize(special);
        for (uint32_t x = 0; x < special; x++) {  // mmer.h:61-87
            uint32_t rc = 0, t = x;
            for (uint32_t i = 0; i < I.signature_len; i++) { rc = (rc << 2) | (3 - (t & 3)); t >>= 2; }
            const uint32_t a = sig_allowed(x, I.signature_len) ? x : special;
            const uint32_t b = sig_allowed(rc, I.signature_len) ? rc : special;
            norm[x] = a < b ? a : b;
        }
    } else if (I.kmc_version == 0) {              // :246-300
        const uint64_t body = fs - 12;
        if (body < hoff || hoff < 40) { pf::set_error("%s.kmc_pre: truncated KMC1 header", prefix); return PF_E_IO; }
        const uint64_t hi = (body - hoff) / 8;
        const unsigned char *h = &pre[4 + hi * 8];
        const uint64_t w0 = le64(h), w1 = le64(h + 8), w2 = le64(h + 16), w3 = le64(h + 24), w4 = le64(h + 32);
        I.kmer_length = (uint32_t)w0; I.mode = (uint32_t)(w0 >> 32);
        I.counter_size = (uint32_t)w1; I.lut_prefix_length = (uint32_t)(w1 >> 32);
        I.min_count = (uint32_t)w2; I.max_count = (w2 >> 32) + (w4 & 0xFFFFFFFF00000000ull);
        I.total_kmers = w3; I.both_strands = ((w4 & 0xF) == 1) ? 0 : 1;
        I.signature_len = 0;
        lut.resize(hi + 1);
        memcpy(lut.data(), &pre[4], hi * 8);
        lut[hi] = I.total_kmers + 1;              // sentinel over the first header word (:292)
    } else {
        pf::set_error("%s.kmc_pre: unsupported kmc_version 0x%x", prefix, I.kmc_version);
        return PF_E_IO;
    }
    if (I.mode != 0) { pf::set_error("%s: quake-mode (float) counters are not supported", prefix); return PF_E_UNSUPPORTED; }
    const uint32_t k = I.kmer_length, p = I.lut_prefix_length, C = I.counter_size;
    if (k == 0 || k > 32) { pf::set_error("%s: k=%u outside 1..32 (reference build: MAX_KMER_SIZE=32)", prefix, k); return PF_E_UNSUPPORTED; }
    if (p < 1 || p >= k || (k - p) % 4 != 0 || p > 15) { pf::set_error("%s: bad lut_prefix_length %u for k=%u", prefix, p, k); return PF_E_IO; }
    if (C < 1 || C > 4) { pf::set_error("%s: counter_size %u outside 1..4", prefix, C); return PF_E_UNSUPPORTED; }
    const uint32_t S = (k - p) / 4;
    const uint64_t single = 1ull << (2 * p);
    I.n_bins = I.kmc_version == 0x200 ? (uint32_t)((lut.size() - 1) / single) : 1;
    if (I.kmc_version == 0 && lut.size() < single + 1) { pf::set_error("%s.kmc_pre: prefix table shorter than 4^p", prefix); return PF_E_IO; }
    if (I.kmc_version == 0x200)
        for (uint32_t v : sigmap)
            if ((uint64_t)v >= I.n_bins) { pf::set_error("%s.kmc_pre: signature map points past the last bin", prefix); return PF_E_IO; }
    pre.clear();
    pre.shrink_to_fit();

    // .kmc_suf is not read into host memory as a whole: its markers and size are checked here, the records are streamed to
    // the device further down through two pinned staging buffers (read of chunk i + 1 overlaps the copy of chunk i)
    const std::string suf_path = base + ".kmc_suf";
    FILE *suf_f = fopen(suf_path.c_str(), "rb");
    if (!suf_f) { pf::set_error("cannot read %s.kmc_suf", prefix); return PF_E_IO; }
    struct FileCloser { FILE *f; ~FileCloser() { if (f) fclose(f); } } suf_closer{suf_f};
    long long suf_size = -1;
    {
        char m0[4] = {0}, m1[4] = {0};
        if (fseeko(suf_f, 0, SEEK_END) == 0) suf_size = (long long)ftello(suf_f);
        bool ok = suf_size >= 8;
        ok = ok && fseeko(suf_f, 0, SEEK_SET) == 0 && fread(m0, 1, 4, suf_f) == 4;
        ok = ok && fseeko(suf_f, (off_t)(suf_size - 4), SEEK_SET) == 0 && fread(m1, 1, 4, suf_f) == 4;
        if (!ok || memcmp(m0, "KMCS", 4) || memcmp(m1, "KMCS", 4)) { pf::set_error("%s.kmc_suf: bad KMCS markers", prefix); return PF_E_IO; }
    }
    const uint64_t Nall = I.total_kmers, R = S + C;
    if ((uint64_t)suf_size - 8 < Nall * R) { pf::set_error("%s.kmc_suf: %lld bytes, expected %llu records of %llu bytes", prefix, suf_size, (unsigned long long)Nall, (unsigned long long)R); return PF_E_IO; }
    for (size_t i = 0; i + 1 < lut.size(); i++)
        if (lut[i] > lut[i + 1] || lut[i] > Nall) { pf::set_error("%s.kmc_pre: prefix table is not monotone", prefix); return PF_E_IO; }

    // ---- default layout: the one-sector hash index, built while the records stream in (no verbatim image in HBM) ----
    if (!(flags & PF_KMC_INDEX_VERBATIM)) {
        const int hrc = kmc_stream_hash(db.get(), lut, sigmap, norm, suf_f, prefix, part, n_parts);
        if (hrc) return hrc;
        if (db->hash_on) { *out = db.release(); return PF_OK; }
        // not hashable (fails the verification, or the slot does not fit): the verbatim image below, results are identical
    }
    // ---- partition: local prefix table (rebased record indices) + the byte ranges of the owned records ----
    KmcView &V = db->view;
    V.n_parts = n_parts; V.part = part; V.prefix_lo = 0; V.prefix_cnt = single;
    V.prefix_per_part = (single + n_parts - 1) / n_parts;
    std::vector<std::pair<uint64_t, uint64_t>> ranges;   // [first, last) global record indices, in local order
    uint64_t N = Nall;
    if (n_parts > 1) {
        std::vector<uint64_t> loc;
        uint64_t acc = 0;
        auto rec_end = [&](uint64_t slot) { return std::min<uint64_t>(lut[slot], Nall); };
        if (I.kmc_version == 0x200) {
            for (uint32_t b = part; b < I.n_bins; b += n_parts) {
                const uint64_t s0 = (uint64_t)b * single, g0 = rec_end(s0), g1 = rec_end(s0 + single);
                for (uint64_t x = 0; x < single; x++) loc.push_back(rec_end(s0 + x) - g0 + acc);
                ranges.emplace_back(g0, g1);
                acc += g1 - g0;
            }
        } else {
            V.prefix_lo = std::min<uint64_t>((uint64_t)part * V.prefix_per_part, single);
            V.prefix_cnt = std::min<uint64_t>(V.prefix_per_part, single - V.prefix_lo);
            const uint64_t g0 = rec_end(V.prefix_lo), g1 = rec_end(V.prefix_lo + V.prefix_cnt);
            for (uint64_t x = 0; x < V.prefix_cnt; x++) loc.push_back(rec_end(V.prefix_lo + x) - g0);
            ranges.emplace_back(g0, g1);
            acc = g1 - g0;
        }
        loc.push_back(acc + 1);   // same N+1 end sentinel convention as the reader (:233, :292)
        lut.swap(loc);
        N = acc;
    } else {
        ranges.emplace_back(0, Nall);
    }

    // ---- device image ----
    V.k = k; V.p = p; V.S = S; V.C = C; V.sig_len = I.signature_len; V.is_kmc2 = I.kmc_version == 0x200;
    V.min_count = I.min_count; V.max_count = I.max_count; V.N = N; V.single_lut = single; V.lut_n = lut.size();
    V.lut64 = (N + 1 >= (1ull << 32)) ? 1 : 0;
    V.packed = R <= 8 ? 1 : 0;
    db->orig_min = I.min_count; db->orig_max = I.max_count;
    cudaStream_t st = ctx->stream;
    uint64_t bytes = 0;
    if (V.lut64) {
        PF_CUDA_TRY(cudaMalloc(&db->d_lut, lut.size() * 8));
        PF_CUDA_TRY(cudaMemcpyAsync(db->d_lut, lut.data(), lut.size() * 8, cudaMemcpyHostToDevice, st));
        PF_CUDA_TRY(pf::stream_sync(st));
        bytes += lut.size() * 8;
    } else {
        std::vector<uint32_t> l32(lut.size());
        for (size_t i = 0; i < lut.size(); i++) l32[i] = (uint32_t)lut[i];
        PF_CUDA_TRY(cudaMalloc(&db->d_lut, l32.size() * 4));
        PF_CUDA_TRY(cudaMemcpyAsync(db->d_lut, l32.data(), l32.size() * 4, cudaMemcpyHostToDevice, st));
        PF_CUDA_TRY(pf::stream_sync(st));
        bytes += l32.size() * 4;
    }
    V.lut = db->d_lut;
    if (V.is_kmc2) {
        PF_CUDA_TRY(cudaMalloc(&db->d_sigmap, sigmap.size() * 4));
        PF_CUDA_TRY(cudaMemcpyAsync(db->d_sigmap, sigmap.data(), sigmap.size() * 4, cudaMemcpyHostToDevice, st));
        PF_CUDA_TRY(cudaMalloc(&db->d_norm, norm.size() * 4));
        PF_CUDA_TRY(cudaMemcpyAsync(db->d_norm, norm.data(), norm.size() * 4, cudaMemcpyHostToDevice, st));
        PF_CUDA_TRY(pf::stream_sync(st));
        bytes += (sigmap.size() + norm.size()) * 4;
        V.sigmap = (const uint32_t *)db->d_sigmap;
        V.norm = (const uint32_t *)db->d_norm;
    }
    if (N) {
        void *d_raw = nullptr;
        PF_CUDA_TRY(cudaMalloc(&d_raw, N * R));
        {
            const size_t CH = 64u << 20;
            pf::PinnedBuf stage[2];
            cudaEvent_t done[2];
            int rcs = 0;
            for (int i = 0; i < 2 && !rcs; i++) { rcs = stage[i].reserve(CH); if (!rcs && cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming) != cudaSuccess) rcs = PF_E_CUDA; }
            uint64_t at = 0;
            int slot = 0;
            bool io_ok = true;
            for (auto &r : ranges) {
                uint64_t left = (r.second - r.first) * R;
                if (left && fseeko(suf_f, (off_t)(4 + r.first * R), SEEK_SET) != 0) io_ok = false;
                while (left && io_ok && !rcs) {
                    const size_t nb = (size_t)std::min<uint64_t>(left, CH);
                    cudaEventSynchronize(done[slot]);                      // the copy that last used this staging buffer is over
                    if (fread(stage[slot].p, 1, nb, suf_f) != nb) { io_ok = false; break; }
                    if (cudaMemcpyAsync((uint8_t *)d_raw + at, stage[slot].p, nb, cudaMemcpyHostToDevice, st) != cudaSuccess) { rcs = PF_E_CUDA; break; }
                    cudaEventRecord(done[slot], st);
                    at += nb; left -= nb; slot ^= 1;
                }
            }
            pf::stream_sync(st);
            for (int i = 0; i < 2; i++) { cudaEventDestroy(done[i]); stage[i].release(); }
            if (!io_ok || rcs) { cudaFree(d_raw); if (!io_ok) { pf::set_error("%s.kmc_suf: short read", prefix); return PF_E_IO; } return rcs; }
        }
        if (V.packed) {
            PF_CUDA_TRY(cudaMalloc(&db->d_rec, N * 8));
            bytes += N * 8;
        } else {
            PF_CUDA_TRY(cudaMalloc(&db->d_suf, N * 8));
            PF_CUDA_TRY(cudaMalloc(&db->d_cnt, N * 4));
            bytes += N * 12;
        }
        const uint64_t nb = (N + 255) / 256;
        repack_records_kernel<<<(unsigned)nb, 256, 0, st>>>((const uint8_t *)d_raw, N, S, C, (int)V.packed,
                                                            (uint64_t *)db->d_rec, (uint64_t *)db->d_suf, (uint32_t *)db->d_cnt);
        ctx->launches++;
        PF_CUDA_TRY(cudaGetLastError());
        PF_CUDA_TRY(pf::stream_sync(st));
        PF_CUDA_TRY(cudaFree(d_raw));
    }
    PF_CUDA_TRY(pf::stream_sync(st));
    V.rec = (const uint64_t *)db->d_rec; V.suf = (const uint64_t *)db->d_suf; V.cnt = (const uint32_t *)db->d_cnt;
    db->device_bytes = bytes;
    db->local_kmers = N;
    *out = db.release();
    return PF_OK;
}

}  // namespace

extern "C" {

static uint32_t env_open_flags() {   // PF_KMC_INDEX=verbatim keeps the prefix-table + sorted-record image (diagnostics, A/B runs)
    const char *e = getenv("PF_KMC_INDEX");
    return (e && (!strcmp(e, "verbatim") || !strcmp(e, "sorted"))) ? (uint32_t)PF_KMC_INDEX_VERBATIM : 0u;
}

int pf_kmc_open(pf_ctx *ctx, const char *prefix, pf_kmc **out) { return kmc_open_impl(ctx, prefix, 0, 1, env_open_flags(), out); }

int pf_kmc_open_ex(pf_ctx *ctx, const char *prefix, uint32_t flags, pf_kmc **out) { return kmc_open_impl(ctx, prefix, 0, 1, flags, out); }

/* diagnostics: what the open-time verification found (bit 0 unsorted bucket, 1 record in the wrong bin, 2 non-canonical key,
 * 3 table overflow); 0 for a handle that never tried the hash index */
uint32_t pf_kmc_build_status(const pf_kmc *db) { return db ? db->build_status : 0; }

int pf_kmc_index_kind(const pf_kmc *db) { return db ? (db->hash_on ? PF_KMC_INDEX_HASH : PF_KMC_INDEX_VERBATIM) : PF_E_INVALID; }

int pf_kmc_open_part_ex(pf_ctx *ctx, const char *prefix, uint32_t part, uint32_t n_parts, uint32_t flags, pf_kmc **out) {
    if (!ctx || !prefix || !out) { pf::set_error("pf_kmc_open_part: null argument"); return PF_E_INVALID; }
    if (n_parts == 0 || part >= n_parts || n_parts > 254) { pf::set_error("pf_kmc_open_part: bad partition %u of %u", part, n_parts); return PF_E_INVALID; }
    return kmc_open_impl(ctx, prefix, part, n_parts, flags, out);
}

int pf_kmc_open_part(pf_ctx *ctx, const char *prefix, uint32_t part, uint32_t n_parts, pf_kmc **out) {
    return pf_kmc_open_part_ex(ctx, prefix, part, n_parts, env_open_flags(), out);
}

uint64_t pf_kmc_local_kmers(const pf_kmc *db) { return db ? db->local_kmers : 0; }

int pf_kmc_close(pf_kmc *db) {
    if (!db) return PF_OK;
    cudaSetDevice(db->ctx->device);
    if (db->route) {
        pf::DevBuf *d[] = {&db->route->keys, &db->route->owner, &db->route->owner_sorted, &db->route->idx, &db->route->idx_sorted,
                           &db->route->bounds, &db->route->cub_tmp};
        for (auto *b : d) b->release();
        db->route->h_bounds.release();
        delete db->route;
    }
    if (!db->borrowed) {
        cudaFree(db->d_lut); cudaFree(db->d_sigmap); cudaFree(db->d_norm);
        cudaFree(db->d_rec); cudaFree(db->d_suf); cudaFree(db->d_cnt); cudaFree(db->d_hash);
        for (auto &m : db->peer_mapped) if (m) cudaIpcCloseMemHandle(m);
    }
    db->tile_seq.release();
    db->site_status.release(); db->site_ncls.release(); db->site_cov.release(); db->site_skip.release(); db->site_map.release();
    for (auto &b : db->h_site) b.release();
    if (db->k_stream) { pf::stream_sync(db->k_stream); if (!db->k_stream_borrowed) cudaStreamDestroy(db->k_stream); }
    db->k_stage.release();
    if (db->k_staged_ev) cudaEventDestroy(db->k_staged_ev);
    for (auto &b : db->k_in) b.release();
    for (auto &b : db->k_out) b.release();
    delete db;
    return PF_OK;
}

// A second handle on the SAME index for another context of the same device (one pf_ctx per host thread: every thread drives
// its own batches through its own stream and buffers while the HBM-resident index exists once).  The borrowed handle must be
// closed before the handle it was taken from.
int pf_kmc_share(pf_kmc *db, pf_ctx *ctx, pf_kmc **out) {
    if (!db || !ctx || !out) { pf::set_error("pf_kmc_share: null argument"); return PF_E_INVALID; }
    *out = nullptr;
    if (ctx->device != db->ctx->device) { pf::set_error("pf_kmc_share: the context is on another device than the index"); return PF_E_INVALID; }
    pf_kmc *h = new pf_kmc();
    h->ctx = ctx;
    h->info = db->info; h->orig_min = db->orig_min; h->orig_max = db->orig_max;
    h->view = db->view;
    h->d_lut = db->d_lut; h->d_sigmap = db->d_sigmap; h->d_norm = db->d_norm; h->d_rec = db->d_rec; h->d_suf = db->d_suf; h->d_cnt = db->d_cnt;
    h->device_bytes = db->device_bytes; h->local_kmers = db->local_kmers;
    h->hash_on = db->hash_on; h->canonical_ok = db->canonical_ok; h->hview = db->hview; h->d_hash = db->d_hash;
    h->build_status = db->build_status; h->hash_inserted = db->hash_inserted;
    for (int i = 0; i < pfkmc::PF_MAX_PEERS; i++) h->peer_tab[i] = db->peer_tab[i];
    h->peers_attached = db->peers_attached;
    h->borrowed = true;
    *out = h;
    return PF_OK;
}

int pf_kmc_info(const pf_kmc *db, pf_kmc_info_t *info) {
    if (!db || !info) { pf::set_error("pf_kmc_info: null argument"); return PF_E_INVALID; }
    *info = db->info;
    return PF_OK;
}

int pf_kmc_set_min_count(pf_kmc *db, uint32_t x) {
    if (!db) return PF_E_INVALID;
    db->info.min_count = x; db->view.min_count = x;
    return PF_OK;
}
int pf_kmc_set_max_count(pf_kmc *db, uint32_t x) {
    if (!db) return PF_E_INVALID;
    db->info.max_count = x; db->view.max_count = x;
    return PF_OK;
}
int pf_kmc_reset_min_max(pf_kmc *db) {
    if (!db) return PF_E_INVALID;
    db->info.min_count = db->view.min_count = db->orig_min;
    db->info.max_count = db->view.max_count = db->orig_max;
    return PF_OK;
}
uint64_t pf_kmc_device_bytes(const pf_kmc *db) { return db ? db->device_bytes : 0; }

uint64_t pf_window_offsets(const uint64_t *seq_off, uint32_t n_seq, uint32_t k, uint64_t *win_off) {
    uint64_t acc = 0;
    for (uint32_t s = 0; s < n_seq; s++) {
        win_off[s] = acc;
        const uint64_t len = seq_off[s + 1] - seq_off[s];
        if (len >= k) acc += len - k + 1;
    }
    win_off[n_seq] = acc;
    return acc;
}

int pf_kmc_lookup_dev(pf_kmc *db, const void *d_bases, uint64_t n_bases, const void *d_seq_off, const void *d_win_off,
                      uint32_t n_seq, uint64_t n_windows, int mode, uint32_t low, uint32_t up, void *d_counts,
                      void *d_found, void *d_cov, void *cuda_stream) {
    if (!db) { pf::set_error("pf_kmc_lookup_dev: null database"); return PF_E_INVALID; }
    if (mode < PF_LOOKUP_CANONICAL || mode > PF_LOOKUP_FWD) { pf::set_error("pf_kmc_lookup_dev: bad mode %d", mode); return PF_E_INVALID; }
    if (n_windows >= (1ull << 32)) { pf::set_error("pf_kmc_lookup_dev: more than 2^32-1 windows in one call; split the batch"); return PF_E_INVALID; }
    if (db->view.n_parts > 1 && !db->peers_attached) { pf::set_error("pf_kmc_lookup_dev: this index holds one partition of the database; use pf_kmc_route_dev / pf_kmc_lookup_keys_dev, or attach the other partitions (pf_kmc_attach_peers)"); return PF_E_INVALID; }
    pf_ctx *ctx = db->ctx;
    PF_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    if (n_seq == 0) return PF_OK;
    if (d_cov) {
        cov_init_kernel<<<(n_seq + 255) / 256, 256, 0, st>>>((pf_cov_t *)d_cov, (const uint64_t *)d_win_off, n_seq);
        ctx->launches++;
    }
    if (n_bases == 0 || n_windows == 0) { PF_CUDA_TRY(cudaGetLastError()); return PF_OK; }
    if (db->hash_on) {
        pfkmc::HashLookupArgs a;
        a.hv = db->hview; a.k = db->view.k; a.min_count = db->view.min_count; a.max_count = db->view.max_count;
        // on a verified canonical both-strands database "as written, else reverse complement" IS the canonical lookup
        a.mode = (mode == PF_LOOKUP_FWD_THEN_RC && db->info.both_strands && db->canonical_ok) ? (int)PF_LOOKUP_CANONICAL : mode;
        a.bases = (const uint8_t *)d_bases; a.n_bases = n_bases; a.seq_off = (const uint64_t *)d_seq_off;
        a.win_off = (const uint64_t *)d_win_off; a.n_seq = n_seq; a.low = low; a.up = up;
        a.counts = (uint32_t *)d_counts; a.found = (uint8_t *)d_found; a.cov = (pf_cov_t *)d_cov;
        a.n_tiles = (n_bases + pfkmc::HL_TILE - 1) / pfkmc::HL_TILE;
        if (int rc = db->tile_seq.reserve((a.n_tiles + 1) * 4)) return rc;
        a.tile_seq = db->tile_seq.as<uint32_t>();
        pfkmc::tile_seq_kernel<<<(unsigned)((a.n_tiles + 1 + 255) / 256), 256, 0, st>>>(a.seq_off, n_seq, a.n_tiles, db->tile_seq.as<uint32_t>());
        ctx->launches++;
        static int per_sm = 0;
        if (!per_sm) {
            PF_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pfkmc::kmc_hash_lookup_kernel<0>, pfkmc::HL_THREADS, 0));
            if (per_sm < 1) per_sm = 1;
        }
        const unsigned hgrid = (unsigned)std::min<uint64_t>(a.n_tiles, (uint64_t)ctx->sm_count * per_sm);
        a.n_parts = db->view.n_parts; a.route_keys = nullptr; a.route_owner = nullptr;
        a.peer_lookup = db->peers_attached ? 1u : 0u;
        for (int i = 0; i < pfkmc::PF_MAX_PEERS; i++) a.peer_tab[i] = db->peer_tab[i];
        if (db->peers_attached) pfkmc::kmc_hash_lookup_kernel<2><<<hgrid, pfkmc::HL_THREADS, 0, st>>>(a);
        else pfkmc::kmc_hash_lookup_kernel<0><<<hgrid, pfkmc::HL_THREADS, 0, st>>>(a);
        ctx->launches++;
        PF_CUDA_TRY(cudaGetLastError());
        return PF_OK;
    }
    const uint64_t n_tiles = (n_bases + LK_TILE - 1) / LK_TILE;
    const unsigned grid = (unsigned)std::min<uint64_t>(n_tiles, (uint64_t)ctx->sm_count * 6);
    kmc_lookup_kernel<false><<<grid, LK_THREADS, 0, st>>>(db->view, (const uint8_t *)d_bases, n_bases, (const uint64_t *)d_seq_off,
                                                         (const uint64_t *)d_win_off, n_seq, mode, low, up, (uint32_t *)d_counts,
                                                         (uint8_t *)d_found, (pf_cov_t *)d_cov, n_tiles, nullptr, nullptr);
    ctx->launches++;
    PF_CUDA_TRY(cudaGetLastError());
    return PF_OK;
}

// ---- partitioned database: route -> [all-to-all] -> lookup at the owner -> [all-to-all] -> scatter ----
int pf_kmc_route_dev(pf_kmc *db, const void *d_bases, uint64_t n_bases, const void *d_seq_off, const void *d_win_off,
                     uint32_t n_seq, uint64_t n_windows, int mode, void *d_send_keys, void *d_send_idx, uint64_t *h_send_off,
                     void *cuda_stream) {
    if (!db || !h_send_off) { pf::set_error("pf_kmc_route_dev: null argument"); return PF_E_INVALID; }
    if (mode < PF_LOOKUP_CANONICAL || mode > PF_LOOKUP_FWD) { pf::set_error("pf_kmc_route_dev: bad mode %d", mode); return PF_E_INVALID; }
    if (mode == PF_LOOKUP_FWD_THEN_RC && !db->info.both_strands) {
        pf::set_error("pf_kmc_route_dev: FWD_THEN_RC is routed as the canonical key, which needs a both-strands database");
        return PF_E_UNSUPPORTED;
    }
    if (n_windows >= (1ull << 32)) { pf::set_error("pf_kmc_route_dev: more than 2^32-1 windows in one call; split the batch"); return PF_E_INVALID; }
    pf_ctx *ctx = db->ctx;
    PF_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    const uint32_t P = db->view.n_parts;
    for (uint32_t o = 0; o <= P; o++) h_send_off[o] = 0;
    if (n_seq == 0 || n_windows == 0 || n_bases == 0) return PF_OK;
    if (!db->route) db->route = new pf_kmc_route_state();
    pf_kmc_route_state *R = db->route;
    int rc;
    if ((rc = R->keys.reserve(n_windows * 8))) return rc;
    if ((rc = R->owner.reserve(n_windows + 16))) return rc;
    if ((rc = R->owner_sorted.reserve(n_windows + 16))) return rc;
    if ((rc = R->idx.reserve(n_windows * 4))) return rc;
    if ((rc = R->bounds.reserve((P + 2) * 8))) return rc;
    if ((rc = R->h_bounds.reserve((P + 2) * 8))) return rc;
    // windows shorter sequences never produce stay 0xFF (not sent)
    PF_CUDA_TRY(cudaMemsetAsync(R->owner.p, 0xFF, n_windows, st));
    if (db->hash_on) {   // partition of the hash index: keys are owned by mix(key) % n_parts
        pfkmc::HashLookupArgs a;
        a.hv = db->hview; a.k = db->view.k; a.min_count = db->view.min_count; a.max_count = db->view.max_count;
        a.mode = mode == PF_LOOKUP_FWD ? (int)PF_LOOKUP_FWD : (int)PF_LOOKUP_CANONICAL;   // FWD_THEN_RC travels as the canonical key
        a.bases = (const uint8_t *)d_bases; a.n_bases = n_bases; a.seq_off = (const uint64_t *)d_seq_off;
        a.win_off = (const uint64_t *)d_win_off; a.n_seq = n_seq; a.low = 0; a.up = 0;
        a.counts = nullptr; a.found = nullptr; a.cov = nullptr;
        a.n_tiles = (n_bases + pfkmc::HL_TILE - 1) / pfkmc::HL_TILE;
        if ((rc = db->tile_seq.reserve((a.n_tiles + 1) * 4))) return rc;
        a.tile_seq = db->tile_seq.as<uint32_t>();
        a.n_parts = P; a.route_keys = R->keys.as<unsigned long long>(); a.route_owner = R->owner.as<uint8_t>();
        a.peer_lookup = 0;
        for (int i = 0; i < pfkmc::PF_MAX_PEERS; i++) a.peer_tab[i] = nullptr;
        pfkmc::tile_seq_kernel<<<(unsigned)((a.n_tiles + 1 + 255) / 256), 256, 0, st>>>(a.seq_off, n_seq, a.n_tiles, db->tile_seq.as<uint32_t>());
        const unsigned hgrid = (unsigned)std::min<uint64_t>(a.n_tiles, (uint64_t)ctx->sm_count * 3);
        pfkmc::kmc_hash_lookup_kernel<1><<<hgrid, pfkmc::HL_THREADS, 0, st>>>(a);
        ctx->launches++;
    } else {
    const uint64_t n_tiles = (n_bases + LK_TILE - 1) / LK_TILE;
    const unsigned grid = (unsigned)std::min<uint64_t>(n_tiles, (uint64_t)ctx->sm_count * 6);
    kmc_lookup_kernel<true><<<grid, LK_THREADS, 0, st>>>(db->view, (const uint8_t *)d_bases, n_bases, (const uint64_t *)d_seq_off,
                                                        (const uint64_t *)d_win_off, n_seq, mode, 0, 0, nullptr, nullptr, nullptr, n_tiles,
                                                        R->keys.as<unsigned long long>(), R->owner.as<uint8_t>());
    }
    // bucket by owner: stable radix sort of (owner, window index) on the 8 owner bits
    {
        uint32_t *idx = R->idx.as<uint32_t>();   // window indices 0..n-1 as the sort's value array
        pf_iota_u32<<<(unsigned)((n_windows + 255) / 256), 256, 0, st>>>(idx, n_windows);
        size_t tmp = 0;
        PF_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp, R->owner.as<uint8_t>(), R->owner_sorted.as<uint8_t>(), idx,
                                                    (uint32_t *)d_send_idx, (int)n_windows, 0, 8, st));
        if ((rc = R->cub_tmp.reserve(tmp + 16))) return rc;
        tmp = R->cub_tmp.cap;
        PF_CUDA_TRY(cub::DeviceRadixSort::SortPairs(R->cub_tmp.p, tmp, R->owner.as<uint8_t>(), R->owner_sorted.as<uint8_t>(), idx,
                                                    (uint32_t *)d_send_idx, (int)n_windows, 0, 8, st));
    }
    kmc_owner_bounds_kernel<<<1, 256, 0, st>>>(R->owner_sorted.as<uint8_t>(), n_windows, P, R->bounds.as<uint64_t>());
    kmc_gather_keys_kernel<<<(unsigned)((n_windows + 255) / 256), 256, 0, st>>>(R->keys.as<unsigned long long>(), (const uint32_t *)d_send_idx,
                                                                               n_windows, (unsigned long long *)d_send_keys);
    ctx->launches += 6;
    PF_CUDA_TRY(cudaGetLastError());
    PF_CUDA_TRY(cudaMemcpyAsync(R->h_bounds.p, R->bounds.p, (P + 1) * 8, cudaMemcpyDeviceToHost, st));
    PF_CUDA_TRY(pf::stream_sync(st));
    for (uint32_t o = 0; o <= P; o++) h_send_off[o] = R->h_bounds.as<uint64_t>()[o];
    return PF_OK;
}

int pf_kmc_lookup_keys_dev(pf_kmc *db, const void *d_keys, uint64_t n, void *d_counts, void *d_found, void *cuda_stream) {
    if (!db || (n && (!d_keys || !d_counts || !d_found))) { pf::set_error("pf_kmc_lookup_keys_dev: null argument"); return PF_E_INVALID; }
    pf_ctx *ctx = db->ctx;
    PF_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    if (!n) return PF_OK;
    if (db->hash_on) {
        pfkmc::kmc_hash_lookup_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(db->hview, db->view.min_count, db->view.max_count,
                                                                                       (const unsigned long long *)d_keys, n,
                                                                                       (uint32_t *)d_counts, (uint8_t *)d_found);
        ctx->launches++;
        PF_CUDA_TRY(cudaGetLastError());
        return PF_OK;
    }
    kmc_lookup_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(db->view, (const unsigned long long *)d_keys, n, (uint32_t *)d_counts,
                                                                       (uint8_t *)d_found);
    ctx->launches++;
    PF_CUDA_TRY(cudaGetLastError());
    return PF_OK;
}

int pf_kmc_scatter_dev(pf_kmc *db, const void *d_send_idx, uint64_t n_sent, const void *d_reply_counts, const void *d_reply_found,
                       const void *d_win_off, uint32_t n_seq, uint64_t n_windows, uint32_t low, uint32_t up, void *d_counts,
                       void *d_found, void *d_cov, void *cuda_stream) {
    if (!db || !d_counts || !d_found) { pf::set_error("pf_kmc_scatter_dev: null argument (d_counts and d_found are required)"); return PF_E_INVALID; }
    pf_ctx *ctx = db->ctx;
    PF_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    if (n_windows) {
        PF_CUDA_TRY(cudaMemsetAsync(d_counts, 0, n_windows * 4, st));   // windows that were not routed are "not found"
        PF_CUDA_TRY(cudaMemsetAsync(d_found, 0, n_windows, st));
    }
    if (n_sent) {
        kmc_scatter_kernel<<<(unsigned)((n_sent + 255) / 256), 256, 0, st>>>((const uint32_t *)d_send_idx, n_sent, (const uint32_t *)d_reply_counts,
                                                                            (const uint8_t *)d_reply_found, (uint32_t *)d_counts, (uint8_t *)d_found);
        ctx->launches++;
    }
    if (d_cov && n_seq) {
        kmc_cov_from_counts_kernel<<<(n_seq + 255) / 256, 256, 0, st>>>((const uint64_t *)d_win_off, n_seq, (const uint32_t *)d_counts,
                                                                       (const uint8_t *)d_found, low, up, (pf_cov_t *)d_cov);
        ctx->launches++;
    }
    PF_CUDA_TRY(cudaGetLastError());
    return PF_OK;
}

// ---- peer-memory form of the partitioned index ---------------------------------------------------------------------------
struct PeerBlob {   // 128 bytes exchanged between the ranks (any transport)
    cudaIpcMemHandle_t handle;       // 64 bytes
    uint32_t bucket_bits, rem_bits, cbits, kbits, part, n_parts, hash_on, pad;
    uint8_t reserved[32];
};
static_assert(sizeof(PeerBlob) == 128, "PeerBlob layout");

int pf_kmc_export_ipc(pf_kmc *db, void *blob128) {
    if (!db || !blob128) { pf::set_error("pf_kmc_export_ipc: null argument"); return PF_E_INVALID; }
    if (!db->hash_on || !db->d_hash) { pf::set_error("pf_kmc_export_ipc: only a hash-index partition can be shared"); return PF_E_UNSUPPORTED; }
    PF_CUDA_TRY(cudaSetDevice(db->ctx->device));
    PeerBlob b;
    memset(&b, 0, sizeof(b));
    PF_CUDA_TRY(cudaIpcGetMemHandle(&b.handle, db->d_hash));
    b.bucket_bits = db->hview.bucket_bits; b.rem_bits = db->hview.rem_bits; b.cbits = db->hview.cbits; b.kbits = db->hview.kbits;
    b.part = db->view.part; b.n_parts = db->view.n_parts; b.hash_on = 1;
    memcpy(blob128, &b, sizeof(b));
    return PF_OK;
}

int pf_kmc_attach_peers(pf_kmc *db, const void *blobs, uint32_t n_parts) {
    if (!db || !blobs) { pf::set_error("pf_kmc_attach_peers: null argument"); return PF_E_INVALID; }
    if (!db->hash_on || n_parts != db->view.n_parts || n_parts > (uint32_t)pfkmc::PF_MAX_PEERS) {
        pf::set_error("pf_kmc_attach_peers: needs a hash-index partition and one blob per partition (<= %d)", pfkmc::PF_MAX_PEERS);
        return PF_E_INVALID;
    }
    PF_CUDA_TRY(cudaSetDevice(db->ctx->device));
    const PeerBlob *B = (const PeerBlob *)blobs;
    for (uint32_t r = 0; r < n_parts; r++) {
        if (!B[r].hash_on || B[r].part != r || B[r].n_parts != n_parts || B[r].bucket_bits != db->hview.bucket_bits ||
            B[r].rem_bits != db->hview.rem_bits || B[r].cbits != db->hview.cbits) {
            pf::set_error("pf_kmc_attach_peers: partition %u has a different table geometry", r);
            return PF_E_INVALID;
        }
        if (r == db->view.part) { db->peer_tab[r] = (const unsigned long long *)db->d_hash; continue; }
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, B[r].handle, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            pf::set_error("pf_kmc_attach_peers: cudaIpcOpenMemHandle for partition %u failed: %s", r, cudaGetErrorString(e));
            cudaGetLastError();
            return PF_E_CUDA;
        }
        db->peer_mapped[r] = p;
        db->peer_tab[r] = (const unsigned long long *)p;
    }
    db->peers_attached = true;
    return PF_OK;
}

// device-resident form: d_skip is a device pointer (or NULL); results stay in the handle's device buffers, `out_dev` gets DEVICE pointers
int pf_site_cov_dev(pf_kmc *db, uint32_t low, uint32_t up, const void *d_skip, pf_site_batch_t *out_dev, void *cuda_stream) {
    if (!db) { pf::set_error("pf_site_cov_dev: null database"); return PF_E_INVALID; }
    if (db->view.n_parts > 1 && !db->peers_attached) { pf::set_error("pf_site_cov: this index holds one partition of the database (attach the others with pf_kmc_attach_peers)"); return PF_E_INVALID; }
    pf_ctx *ctx = db->ctx;
    PF_CUDA_TRY(cudaSetDevice(ctx->device));
    pf_msa_batch_t m;
    uint64_t tot[4];
    if (pf_align_last_dev(ctx, &m, tot) != PF_OK) { pf::set_error("pf_site_cov: no alignment result on this context (call pf_align / pf_align_dev first)"); return PF_E_INVALID; }
    const uint32_t n = m.n_bubbles;
    const uint64_t n_var = tot[1], n_cls = tot[2];
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    int rc;
    if ((rc = db->site_status.reserve(n_var + 16))) return rc;
    if ((rc = db->site_ncls.reserve(n_var + 16))) return rc;
    if ((rc = db->site_cov.reserve(n_cls * 8 + 16))) return rc;
    SiteArgs a;
    a.db = db->view; a.hv = db->hview; a.hash_on = db->hash_on ? 1 : 0; a.both_strands = (int)db->info.both_strands;
    a.n = n; a.status = m.status; a.n_rows = m.n_rows; a.aln_len = m.aln_len; a.rows_off = m.rows_off; a.rows = m.rows;
    a.var_off = m.var_off; a.var_col = m.var_col; a.var_kind = m.var_kind; a.cls_off = m.cls_off; a.cls = m.cls;
    a.skip = (const uint8_t *)d_skip; a.low = low; a.up = up;
    a.peer_lookup = db->peers_attached ? 1u : 0u; a.n_parts = db->view.n_parts;
    for (int i = 0; i < pfkmc::PF_MAX_PEERS; i++) a.peer_tab[i] = db->peer_tab[i];
    a.site_status = db->site_status.as<uint8_t>(); a.site_ncls = db->site_ncls.as<uint8_t>();
    a.site_cov = db->site_cov.as<unsigned long long>();
    if ((rc = db->site_map.reserve(n_var * 4 + 16))) return rc;
    a.site_bubble = db->site_map.as<uint32_t>(); a.n_sites = n_var;
    if (n_var) {
        site_map_kernel<<<(n + 255) / 256, 256, 0, st>>>(m.var_off, n, db->site_map.as<uint32_t>());
        site_cov_kernel<<<(unsigned)((n_var + 127) / 128), 128, 0, st>>>(a);
        ctx->launches++;
    }
    ctx->launches++;
    PF_CUDA_TRY(cudaGetLastError());
    if (out_dev) {
        memset(out_dev, 0, sizeof(*out_dev));
        out_dev->n_bubbles = n;
        out_dev->site_off = m.var_off; out_dev->status = db->site_status.as<uint8_t>(); out_dev->n_class = db->site_ncls.as<uint8_t>();
        out_dev->cov_off = m.cls_off; out_dev->cov = db->site_cov.as<uint64_t>();
    }
    db->site_totals[0] = n_var; db->site_totals[1] = n_cls;
    return PF_OK;
}

int pf_site_cov(pf_kmc *db, uint32_t low, uint32_t up, const uint8_t *skip, pf_site_batch_t *out) {
    if (!db || !out) { pf::set_error("pf_site_cov: null argument"); return PF_E_INVALID; }
    pf_ctx *ctx = db->ctx;
    PF_CUDA_TRY(cudaSetDevice(ctx->device));
    memset(out, 0, sizeof(*out));
    cudaStream_t st = ctx->stream;
    pf_msa_batch_t m;
    uint64_t tot[4];
    if (pf_align_last_dev(ctx, &m, tot) != PF_OK) { pf::set_error("pf_site_cov: no alignment result on this context (call pf_align / pf_align_dev first)"); return PF_E_INVALID; }
    const uint32_t n = m.n_bubbles;
    int rc;
    if (skip) {
        if ((rc = db->site_skip.reserve(n + 16))) return rc;
        PF_CUDA_TRY(cudaMemcpyAsync(db->site_skip.p, skip, n, cudaMemcpyHostToDevice, st));
    }
    pf_site_batch_t d;
    if ((rc = pf_site_cov_dev(db, low, up, skip ? db->site_skip.p : nullptr, &d, st))) return rc;
    const uint64_t n1 = (uint64_t)n + 1, n_var = db->site_totals[0], n_cls = db->site_totals[1];
    // site_off / cov_off ARE var_off / cls_off of the alignment: when that came through a host-pointer call they already sit in the
    // context's pinned arena (valid until its next pf_align*), and 16 bytes per bubble need not cross the bus a second time
    const uint64_t *h_var_off = nullptr, *h_cls_off = nullptr;
    const bool have_off = pf_align_last_host_offsets(ctx, n, &h_var_off, &h_cls_off) == PF_OK;
    const void *src[5] = {d.site_off, d.status, d.n_class, d.cov_off, d.cov};
    const uint64_t bytes[5] = {have_off ? 0 : n1 * 8, n_var, n_var, have_off ? 0 : n1 * 8, n_cls * 8};
    for (int i = 0; i < 5; i++) {
        if ((rc = db->h_site[i].reserve(bytes[i] + 16))) return rc;
        if (bytes[i]) PF_CUDA_TRY(cudaMemcpyAsync(db->h_site[i].p, src[i], bytes[i], cudaMemcpyDeviceToHost, st));
    }
    PF_CUDA_TRY(pf::stream_sync(st));
    out->n_bubbles = n;
    out->site_off = have_off ? h_var_off : db->h_site[0].as<uint64_t>(); out->status = db->h_site[1].as<uint8_t>(); out->n_class = db->h_site[2].as<uint8_t>();
    out->cov_off = have_off ? h_cls_off : db->h_site[3].as<uint64_t>(); out->cov = db->h_site[4].as<uint64_t>();
    return PF_OK;
}

// internal (pf_align.cu: pf_align_staged): the sequences the last coverage-only host call left on the device
int pf_kmc_staged_dev(pf_kmc *db, const uint8_t **bases, const uint64_t **seq_off, uint32_t *n_seq, cudaEvent_t *ready) {
    if (!db || !db->k_staged_seq || !db->k_staged_ev) return PF_E_INVALID;
    *bases = db->k_in[0].as<uint8_t>(); *seq_off = db->k_in[1].as<uint64_t>(); *n_seq = db->k_staged_seq; *ready = db->k_staged_ev;
    return PF_OK;
}

// pf_site_kmers: the site k-mers of the context's last alignment, without any lookup (see site_keys_kernel).
int pf_site_kmers(pf_ctx *ctx, uint32_t k, const uint8_t *skip, pf_site_kmers_t *out) {
    if (!ctx || !out) { pf::set_error("pf_site_kmers: null argument"); return PF_E_INVALID; }
    if (k < 1 || k > 32) { pf::set_error("pf_site_kmers: k=%u outside 1..32", k); return PF_E_INVALID; }
    PF_CUDA_TRY(cudaSetDevice(ctx->device));
    memset(out, 0, sizeof(*out));
    cudaStream_t st = ctx->stream;
    pf_msa_batch_t m;
    uint64_t tot[4];
    if (pf_align_last_dev(ctx, &m, tot) != PF_OK) { pf::set_error("pf_site_kmers: no alignment result on this context (call pf_align first)"); return PF_E_INVALID; }
    const uint32_t n = m.n_bubbles;
    const uint64_t n_var = tot[1], n_cls = tot[2], n1 = (uint64_t)n + 1;
    int rc;
    pf::DevBuf &d_status = ctx->d_out[0], &d_keys = ctx->d_out[1], &d_map = ctx->d_out[2], &d_skip = ctx->d_in[0];
    if ((rc = d_status.reserve(n_var + 16))) return rc;
    if ((rc = d_keys.reserve(n_cls * 8 + 16))) return rc;
    if ((rc = d_map.reserve(n_var * 4 + 16))) return rc;
    if (skip) {
        if ((rc = d_skip.reserve(n + 16))) return rc;
        PF_CUDA_TRY(cudaMemcpyAsync(d_skip.p, skip, n, cudaMemcpyHostToDevice, st));
    }
    SiteArgs a{};
    a.db.k = k; a.both_strands = 1;
    a.n = n; a.status = m.status; a.n_rows = m.n_rows; a.aln_len = m.aln_len; a.rows_off = m.rows_off; a.rows = m.rows;
    a.var_off = m.var_off; a.var_col = m.var_col; a.var_kind = m.var_kind; a.cls_off = m.cls_off; a.cls = m.cls;
    a.skip = skip ? d_skip.as<uint8_t>() : nullptr;
    a.site_status = d_status.as<uint8_t>(); a.site_bubble = d_map.as<uint32_t>(); a.n_sites = n_var;
    if (n_var) {
        site_map_kernel<<<(n + 255) / 256, 256, 0, st>>>(m.var_off, n, d_map.as<uint32_t>());
        site_keys_kernel<<<(unsigned)((n_var + 127) / 128), 128, 0, st>>>(a, d_keys.as<unsigned long long>());
        ctx->launches += 2;
    }
    PF_CUDA_TRY(cudaGetLastError());
    pf::PinnedBuf &h_off = ctx->h_stage[0], &h_rest = ctx->h_stage[1];
    if ((rc = h_off.reserve(n1 * 16))) return rc;
    if ((rc = h_rest.reserve(n_cls * 8 + n_var + 32))) return rc;
    uint64_t *h_site_off = h_off.as<uint64_t>(), *h_key_off = h_site_off + n1;
    uint64_t *h_keys = h_rest.as<uint64_t>();
    uint8_t *h_status = (uint8_t *)(h_keys + n_cls);
    PF_CUDA_TRY(cudaMemcpyAsync(h_site_off, m.var_off, n1 * 8, cudaMemcpyDeviceToHost, st));
    PF_CUDA_TRY(cudaMemcpyAsync(h_key_off, m.cls_off, n1 * 8, cudaMemcpyDeviceToHost, st));
    if (n_cls) PF_CUDA_TRY(cudaMemcpyAsync(h_keys, d_keys.p, n_cls * 8, cudaMemcpyDeviceToHost, st));
    if (n_var) PF_CUDA_TRY(cudaMemcpyAsync(h_status, d_status.p, n_var, cudaMemcpyDeviceToHost, st));
    PF_CUDA_TRY(pf::stream_sync(st));
    out->n_bubbles = n;
    out->site_off = h_site_off; out->key_off = h_key_off; out->keys = h_keys; out->status = h_status;
    return PF_OK;
}

// Host-pointer lookups.  Staging and device buffers belong to the database handle (not to the context) so that an asynchronous
// call can run beside pf_align / pf_site_cov of the same context; `async`: enqueue on the handle's own stream and return.
static int kmc_host_call(pf_kmc *db, const char *bases, const uint64_t *seq_off, uint32_t n_seq, int mode, uint32_t low,
                         uint32_t up, uint32_t *counts, uint8_t *found, pf_cov_t *cov, bool async) {
    if (!db || !seq_off || (!bases && n_seq && seq_off[n_seq] > 0)) { pf::set_error("pf_kmc: null argument"); return PF_E_INVALID; }
    pf_ctx *ctx = db->ctx;
    PF_CUDA_TRY(cudaSetDevice(ctx->device));
    if (n_seq == 0) return PF_OK;
    if (db->view.n_parts > 1 && !db->peers_attached) { pf::set_error("pf_kmc_counts/cov: this index holds one partition of the database; use the route / lookup_keys / scatter calls or pf_kmc_attach_peers"); return PF_E_INVALID; }
    if (!db->k_stream) {
        // PF_LOOKUP_SMS=<n>: the handle's stream is confined to an n-SM partition (pf_lookup_partition), so the lookups of a batch
        // run beside the alignment kernels of the same context instead of competing with them for every SM
        static const int part_env = getenv("PF_LOOKUP_SMS") ? atoi(getenv("PF_LOOKUP_SMS")) : 0;
        void *ps = nullptr;
        if (part_env >= 8 && pf_lookup_partition(ctx, (uint32_t)part_env, &ps) == PF_OK && ps) { db->k_stream = (cudaStream_t)ps; db->k_stream_borrowed = true; }
        else PF_CUDA_TRY(cudaStreamCreateWithFlags(&db->k_stream, cudaStreamNonBlocking));
    }
    PF_CUDA_TRY(pf::stream_sync(db->k_stream));   // an earlier asynchronous call still owns the staging buffers
    const uint64_t n_bases = seq_off[n_seq] - seq_off[0];
    int rc;
    if (cov && !counts && !found && seq_off[0] == 0) {
        // coverage only, offsets already zero-based: nothing is touched on the host -- the caller's offsets are copied as they
        // are and the window offsets are an exclusive scan on the device
        cudaStream_t st = db->k_stream;
        if ((rc = db->k_in[0].reserve(n_bases + 16))) return rc;
        if ((rc = db->k_in[1].reserve((uint64_t)(n_seq + 1) * 8))) return rc;
        if ((rc = db->k_in[2].reserve((uint64_t)(n_seq + 1) * 8))) return rc;
        if ((rc = db->k_out[1].reserve((uint64_t)(n_seq + 1) * 8))) return rc;                 // window counts
        if ((rc = db->k_out[2].reserve((uint64_t)n_seq * sizeof(pf_cov_t)))) return rc;
        if (n_bases) PF_CUDA_TRY(cudaMemcpyAsync(db->k_in[0].p, bases, n_bases, cudaMemcpyHostToDevice, st));
        PF_CUDA_TRY(cudaMemcpyAsync(db->k_in[1].p, seq_off, (uint64_t)(n_seq + 1) * 8, cudaMemcpyHostToDevice, st));
        if (!db->k_staged_ev) PF_CUDA_TRY(cudaEventCreateWithFlags(&db->k_staged_ev, cudaEventDisableTiming));
        PF_CUDA_TRY(cudaEventRecord(db->k_staged_ev, st));      // the batch is on the device: pf_align_staged may read it
        db->k_staged_seq = n_seq; db->k_staged_bases = n_bases;
        win_len_kernel<<<(n_seq + 1 + 255) / 256, 256, 0, st>>>(db->k_in[1].as<uint64_t>(), n_seq, db->info.kmer_length, db->k_out[1].as<uint64_t>());
        size_t tmp = 0;
        PF_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp, db->k_out[1].as<uint64_t>(), db->k_in[2].as<uint64_t>(), (int)(n_seq + 1), st));
        if ((rc = db->k_out[0].reserve(tmp + 16))) return rc;
        tmp = db->k_out[0].cap;
        PF_CUDA_TRY(cub::DeviceScan::ExclusiveSum(db->k_out[0].p, tmp, db->k_out[1].as<uint64_t>(), db->k_in[2].as<uint64_t>(), (int)(n_seq + 1), st));
        ctx->launches += 3;
        // the number of windows is only an upper bound here (no per-window output is written)
        rc = pf_kmc_lookup_dev(db, db->k_in[0].p, n_bases, db->k_in[1].p, db->k_in[2].p, n_seq, std::max<uint64_t>(n_bases, 1), mode, low, up,
                               nullptr, nullptr, db->k_out[2].p, st);
        if (rc) return rc;
        PF_CUDA_TRY(cudaMemcpyAsync(cov, db->k_out[2].p, (uint64_t)n_seq * sizeof(pf_cov_t), cudaMemcpyDeviceToHost, st));
        if (!async) PF_CUDA_TRY(pf::stream_sync(st));
        return PF_OK;
    }
    db->k_staged_seq = 0;
    // rebased offsets + window offsets are built straight into pinned staging (one async copy each, no bounce buffer)
    if ((rc = db->k_stage.reserve((uint64_t)(n_seq + 1) * 16))) return rc;
    uint64_t *off = db->k_stage.as<uint64_t>(), *woff = off + (n_seq + 1);
    for (uint32_t s = 0; s <= n_seq; s++) off[s] = seq_off[s] - seq_off[0];
    const uint64_t W = pf_window_offsets(off, n_seq, db->info.kmer_length, woff);
    cudaStream_t st = db->k_stream;
    if ((rc = db->k_in[0].reserve(n_bases + 16))) return rc;
    if ((rc = db->k_in[1].reserve((n_seq + 1) * 8))) return rc;
    if ((rc = db->k_in[2].reserve((n_seq + 1) * 8))) return rc;
    if (counts && (rc = db->k_out[0].reserve(W * 4 + 4))) return rc;
    if (found && (rc = db->k_out[1].reserve(W + 4))) return rc;
    if (cov && (rc = db->k_out[2].reserve((uint64_t)n_seq * sizeof(pf_cov_t)))) return rc;
    if (n_bases) PF_CUDA_TRY(cudaMemcpyAsync(db->k_in[0].p, bases + seq_off[0], n_bases, cudaMemcpyHostToDevice, st));
    PF_CUDA_TRY(cudaMemcpyAsync(db->k_in[1].p, off, (n_seq + 1) * 8, cudaMemcpyHostToDevice, st));
    PF_CUDA_TRY(cudaMemcpyAsync(db->k_in[2].p, woff, (n_seq + 1) * 8, cudaMemcpyHostToDevice, st));
    rc = pf_kmc_lookup_dev(db, db->k_in[0].p, n_bases, db->k_in[1].p, db->k_in[2].p, n_seq, W, mode, low, up,
                           counts ? db->k_out[0].p : nullptr, found ? db->k_out[1].p : nullptr,
                           cov ? db->k_out[2].p : nullptr, st);
    if (rc) return rc;
    if (counts && W) PF_CUDA_TRY(cudaMemcpyAsync(counts, db->k_out[0].p, W * 4, cudaMemcpyDeviceToHost, st));
    if (found && W) PF_CUDA_TRY(cudaMemcpyAsync(found, db->k_out[1].p, W, cudaMemcpyDeviceToHost, st));
    if (cov) PF_CUDA_TRY(cudaMemcpyAsync(cov, db->k_out[2].p, (uint64_t)n_seq * sizeof(pf_cov_t), cudaMemcpyDeviceToHost, st));
    if (!async) PF_CUDA_TRY(pf::stream_sync(st));
    return PF_OK;
}

int pf_kmc_cov_async(pf_kmc *db, const char *bases, const uint64_t *seq_off, uint32_t n_seq, int mode, uint32_t low, uint32_t up,
                     pf_cov_t *out) {
    if (!out) { pf::set_error("pf_kmc_cov_async: null output"); return PF_E_INVALID; }
    return kmc_host_call(db, bases, seq_off, n_seq, mode, low, up, nullptr, nullptr, out, true);
}

int pf_kmc_wait(pf_kmc *db) {
    if (!db) return PF_E_INVALID;
    PF_CUDA_TRY(cudaSetDevice(db->ctx->device));
    if (db->k_stream) PF_CUDA_TRY(pf::stream_sync(db->k_stream));
    return PF_OK;
}

int pf_kmc_counts(pf_kmc *db, const char *bases, const uint64_t *seq_off, uint32_t n_seq, int mode, uint32_t *counts,
                  uint8_t *found) {
    if (!counts && !found) { pf::set_error("pf_kmc_counts: no output requested"); return PF_E_INVALID; }
    return kmc_host_call(db, bases, seq_off, n_seq, mode, 0, 0xFFFFFFFFu, counts, found, nullptr, false);
}

int pf_kmc_cov(pf_kmc *db, const char *bases, const uint64_t *seq_off, uint32_t n_seq, int mode, uint32_t low, uint32_t up,
               pf_cov_t *out) {
    if (!out) { pf::set_error("pf_kmc_cov: null output"); return PF_E_INVALID; }
    return kmc_host_call(db, bases, seq_off, n_seq, mode, low, up, nullptr, nullptr, out, false);
}

}  // extern "C"
