/*
 * pf_gpu.h -- C ABI of libpfgpu.so, the B200 (sm_100a) implementation of PloidyFrost's per-superbubble
 * hot path.  Plain pointers and sizes only; no C++/torch/CUDA types cross this boundary.
 *
 * The reference has no FFI for this path: the seam is two C++ classes used by value,
 *   CKMCFile  (KMC/kmc_api/kmc_file.h:105-167; member CDBG.hpp:13, vector<CKMCFile*> CCDBG.hpp:12)
 *   SeqAlign  (src/SeqAlign.hpp:7-22; one stack object per bubble, CDBG.cpp:2036, :2265)
 * so every entry point below names the reference member it stands in for.  Calls are *batched*
 * (one k-mer or one bubble per call cannot feed a GPU); INTEGRATION.md shows the binding a maintainer
 * of the reference would write (a CKMCFile / SeqAlign look-alike over these calls).
 *
 * Conventions: int return, 0 = PF_OK, negative = error class, text via pf_last_error() (thread-local).
 * No exceptions or aborts cross the ABI.  Inputs are caller-owned.  Unless a function says otherwise,
 * pointers are HOST pointers and the call includes the host<->device copies.  Functions ending in
 * `_dev` take DEVICE pointers (inputs already resident in HBM) plus a CUDA stream handle (a
 * cudaStream_t passed as void*, NULL = the context's own stream) and are asynchronous on that stream.
 * A pf_ctx is bound to one GPU; one process per GPU is the intended deployment (bubbles shard
 * across processes with no data-path collective).  A pf_ctx may be used from one host thread at a time.
 */
#ifndef PF_GPU_H
#define PF_GPU_H

#include <stddef.h>
#include <stdint.h>
#include "pf_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pf_ctx pf_ctx;
typedef struct pf_kmc pf_kmc;

enum {
    PF_OK = 0,
    PF_E_INVALID = -1, /* bad argument                                  */
    PF_E_IO = -2,      /* file missing / malformed KMC database          */
    PF_E_CUDA = -3,    /* CUDA runtime or kernel failure                 */
    PF_E_NOMEM = -4,   /* host or device allocation failed               */
    PF_E_UNSUPPORTED = -5 /* valid for the reference but outside this build (e.g. float counters, k>32) */
};

/* ---- context -------------------------------------------------------------------------------- */
int pf_init(int device, pf_ctx **ctx);
void pf_shutdown(pf_ctx *ctx);
const char *pf_last_error(void);
/* library + device description as a short static string ("libpfgpu <ver> sm_100a ...") */
const char *pf_version(void);
/* number of kernel launches issued by this context since creation (bench.py reports the delta) */
uint64_t pf_launch_count(const pf_ctx *ctx);
int pf_sync(pf_ctx *ctx);

/* ---- KMC database: CKMCFile ------------------------------------------------------------------- */
/* CKMCFile::OpenForRA (kmc_file.cpp:27): parse <prefix>.kmc_pre/.kmc_suf, build the HBM-resident index. */
int pf_kmc_open(pf_ctx *ctx, const char *prefix, pf_kmc **db);
/*
 * Index layout held in HBM.  By default (PF_KMC_INDEX_AUTO) the records are re-hashed at open time into one-sector
 * buckets (one 32-byte load per lookup instead of the reference's signature -> prefix table -> binary search chain;
 * ploidyfrost_b200/csrc/pf_kmc_hash.cuh) after every record has been verified to be where the reference's own search
 * (CheckKmer kmc_file.cpp:330, BinarySearch :1383) would find it; a database that fails the verification, and every
 * partitioned index, keeps the verbatim image (prefix table + sorted records) and the reference's chain.  Results are
 * the same either way.  PF_KMC_INDEX_VERBATIM as `flags` (or PF_KMC_INDEX=verbatim in the environment of pf_kmc_open)
 * forces the verbatim image.
 */
enum { PF_KMC_INDEX_AUTO = 0, PF_KMC_INDEX_VERBATIM = 1, PF_KMC_INDEX_HASH = 2 };
int pf_kmc_open_ex(pf_ctx *ctx, const char *prefix, uint32_t flags, pf_kmc **db);
/* PF_KMC_INDEX_VERBATIM or PF_KMC_INDEX_HASH: the layout this handle ended up with */
int pf_kmc_index_kind(const pf_kmc *db);
/* diagnostics: result of the open-time verification (bit 0: a prefix bucket is not strictly ascending, bit 1: a record is not
 * in the bin its signature maps to, bit 2: a key is not the canonical form, bit 3: hash table overflow) */
uint32_t pf_kmc_build_status(const pf_kmc *db);
/* CKMCFile::Close (kmc_file.cpp:631) */
int pf_kmc_close(pf_kmc *db);
/* CKMCFile::Info (kmc_file.cpp Info(CKMCFileInfo&)) */
int pf_kmc_info(const pf_kmc *db, pf_kmc_info_t *info);
/* CKMCFile::SetMinCount / SetMaxCount / ResetMinMaxCounts (kmc_file.h:118-128,157) */
int pf_kmc_set_min_count(pf_kmc *db, uint32_t x);
int pf_kmc_set_max_count(pf_kmc *db, uint32_t x);
int pf_kmc_reset_min_max(pf_kmc *db);
/* bytes of HBM held by the index */
uint64_t pf_kmc_device_bytes(const pf_kmc *db);

/*
 * Batched CKMCFile::GetCountersForRead (kmc_file.cpp:904) / CheckKmer (:330) / IsKmer (:775).
 * For every k-mer window of every sequence (len-k+1 windows, none when len < k), in sequence order:
 *   counts[w] = counter if the key is present and min_count <= counter <= max_count, else 0
 *   found[w]  = 1 / 0 likewise (may be NULL)
 * `mode` is PF_LOOKUP_* (pf_types.h).  Windows containing a non-ACGT character are not found
 * (kmc_file.cpp:1036-1047, kmer_api.h:502-509).  counts/found must hold sum(max(len-k+1,0)) entries.
 */
int pf_kmc_counts(pf_kmc *db, const char *bases, const uint64_t *seq_off, uint32_t n_seq, int mode,
                  uint32_t *counts, uint8_t *found);

/*
 * Batched coverage reducers CDBG::readCov (CDBG.cpp:29, :66) / CCDBG::readCov, readCovUni
 * (CCDBG.cpp:89, :123): one pf_cov_t per sequence.  The caller derives the reference's return values:
 *   unitig form  : fatal if first_missing >= 0, else (sum / n_kmers, min)
 *   string form  : scan order decides -- fatal if first_missing >= 0 and (first_outside < 0 or
 *                  first_missing < first_outside); (0,false) if first_outside >= 0; else (sum/n_kmers, true)
 */
int pf_kmc_cov(pf_kmc *db, const char *bases, const uint64_t *seq_off, uint32_t n_seq, int mode,
               uint32_t low, uint32_t up, pf_cov_t *out);

/*
 * Asynchronous pf_kmc_cov: enqueues copy-in, lookup and copy-out on the handle's own stream and returns; `bases`, `seq_off` and
 * `out` must stay valid (and should be pinned for the copies to be truly asynchronous) until pf_kmc_wait(db) returns.  Lets the
 * lookups of a batch run beside pf_align / pf_site_cov of the same context.  One outstanding call per handle.
 */
int pf_kmc_cov_async(pf_kmc *db, const char *bases, const uint64_t *seq_off, uint32_t n_seq, int mode, uint32_t low, uint32_t up,
                     pf_cov_t *out);
int pf_kmc_wait(pf_kmc *db);

/*
 * Device-resident forms.  d_bases (n_bases chars), d_seq_off (n_seq+1 x u64) and d_win_off
 * (n_seq+1 x u64, exclusive prefix sum of max(len-k+1,0); pf_window_offsets computes it on the host)
 * are device pointers; d_seq_off[0] must be 0 and d_bases should be 16-byte aligned (it is staged with
 * 128-bit loads; an unaligned pointer falls back to byte loads).  Outputs are device pointers; d_counts/d_found may be NULL when only d_cov is
 * wanted and vice versa.  d_cov must be n_seq x pf_cov_t.
 */
int pf_kmc_lookup_dev(pf_kmc *db, const void *d_bases, uint64_t n_bases, const void *d_seq_off,
                      const void *d_win_off, uint32_t n_seq, uint64_t n_windows, int mode, uint32_t low,
                      uint32_t up, void *d_counts, void *d_found, void *d_cov, void *cuda_stream);
/* host helper: win_off[n_seq+1] from seq_off and k; returns the total number of windows */
uint64_t pf_window_offsets(const uint64_t *seq_off, uint32_t n_seq, uint32_t k, uint64_t *win_off);

/* ---- partitioned database (BASELINE config 3) --------------------------------------------------------
 * When the index does not fit one GPU it is partitioned -- KMC2: bins with bin % n_parts == part (a k-mer's bin is
 * signature_map[signature], kmc_file.cpp:349-351); KMC1: the part-th range of ceil(4^p / n_parts) prefixes -- and every
 * query is answered by the rank that owns its bin.  Per batch, on every rank:
 *   pf_kmc_route_dev        keys of all live windows, bucketed by owner      (kernel + sort)
 *   [all-to-all of keys]    8 bytes per query; the caller's transport (NCCL via torch.distributed in sharded.py)
 *   pf_kmc_lookup_keys_dev  CheckKmer of the received keys against the local partition
 *   [all-to-all of replies] 4-byte counter + 1-byte found per query
 *   pf_kmc_scatter_dev      replies back into window order + the readCov reductions
 * Results equal the unpartitioned calls (PF_LOOKUP_FWD_THEN_RC is routed as the canonical key, which is the same
 * lookup on a both-strands database and refused otherwise). */
int pf_kmc_open_part(pf_ctx *ctx, const char *prefix, uint32_t part, uint32_t n_parts, pf_kmc **db);
/* same with PF_KMC_INDEX_* flags.  By default a partition is a slice of the one-sector hash index -- keys are owned by
 * mix(key) % n_parts, every rank stages the whole database once at open time and keeps its own keys -- and falls back to the
 * bin / prefix partition of the verbatim image (PF_KMC_INDEX_VERBATIM forces it) when the database fails the verification.
 * All ranks of a job must use the same layout (they route by the same owner function). */
int pf_kmc_open_part_ex(pf_ctx *ctx, const char *prefix, uint32_t part, uint32_t n_parts, uint32_t flags, pf_kmc **db);
/*
 * Peer-memory form (one process per GPU of one NVLink / NVSwitch box): every rank exports its slice of the hash index as a
 * 128-byte blob (CUDA IPC handle + table geometry), the ranks exchange the blobs by any transport, and pf_kmc_attach_peers
 * maps the other slices.  After that the ordinary calls on the partition handle -- pf_kmc_lookup_dev, pf_kmc_counts, pf_kmc_cov,
 * pf_kmc_cov_async -- answer EVERY query in one kernel: the owner of a key is mix(key) % n_parts and its bucket is loaded from
 * the owner's HBM over NVLink (one 32-byte read per lookup); there is no route / all-to-all / scatter step.
 */
int pf_kmc_export_ipc(pf_kmc *db, void *blob128);
int pf_kmc_attach_peers(pf_kmc *db, const void *blobs, uint32_t n_parts);
/* records held by this index (== total_kmers unless partitioned) */
uint64_t pf_kmc_local_kmers(const pf_kmc *db);
/* d_send_keys: u64[n_windows], d_send_idx: u32[n_windows] (device, caller-allocated).  On return h_send_off[0..n_parts]
 * (host) holds the bucket boundaries: keys for partition o are d_send_keys[h_send_off[o] .. h_send_off[o+1]), and
 * d_send_idx[t] is the window each sent key came from.  Windows that are not looked up (non-ACGT) are not sent.
 * Synchronises the stream once (the boundaries are needed to size the exchange). */
int pf_kmc_route_dev(pf_kmc *db, const void *d_bases, uint64_t n_bases, const void *d_seq_off, const void *d_win_off,
                     uint32_t n_seq, uint64_t n_windows, int mode, void *d_send_keys, void *d_send_idx, uint64_t *h_send_off,
                     void *cuda_stream);
/* d_keys: u64[n] right-aligned 2-bit k-mers -> d_counts u32[n], d_found u8[n] (0 / 0 when absent, out of
 * [min_count,max_count] or not owned by this partition) */
int pf_kmc_lookup_keys_dev(pf_kmc *db, const void *d_keys, uint64_t n, void *d_counts, void *d_found, void *cuda_stream);
/* replies in send order -> d_counts u32[n_windows], d_found u8[n_windows] (both required) and, when d_cov != NULL,
 * one pf_cov_t per sequence */
int pf_kmc_scatter_dev(pf_kmc *db, const void *d_send_idx, uint64_t n_sent, const void *d_reply_counts, const void *d_reply_found,
                       const void *d_win_off, uint32_t n_seq, uint64_t n_windows, uint32_t low, uint32_t up, void *d_counts,
                       void *d_found, void *d_cov, void *cuda_stream);

/* ---- SeqAlign ------------------------------------------------------------------------------- */
/*
 * Batched SeqAlign::SequenceAlignment (SeqAlign.cpp:550) with scoring SeqAlign(M, D, G)
 * (SeqAlign.hpp:10): for every bubble, aligns its sequences (already ordered by the caller) and calls
 * sites exactly as needlemanWunch (:480) -> traceback (:306) -> variantAnalyze (:237) ->
 * compareStrPair (:8) do.  `out` views context-owned pinned host memory, valid until the next
 * pf_align* call on the same context.
 */
int pf_align(pf_ctx *ctx, double M, double D, double G, const char *bases, const uint64_t *seq_off,
             const uint32_t *bubble_off, uint32_t n_bubbles, pf_msa_batch_t *out);

/*
 * Device-resident form: inputs are device pointers; results stay in context-owned device memory and
 * `out_dev` receives DEVICE pointers (same layout).  Asynchronous on the stream except for one
 * scalar read-back that sizes the compacted result.  `max_len` / `max_rows` are the longest sequence
 * and the largest bubble in the batch (the host knows them from the offsets it built).
 */
int pf_align_dev(pf_ctx *ctx, double M, double D, double G, const void *d_bases, uint64_t n_bases,
                 const void *d_seq_off, uint32_t n_seq, const void *d_bubble_off, uint32_t n_bubbles,
                 uint32_t max_len, uint32_t max_rows, pf_msa_batch_t *out_dev, void *cuda_stream);
/* diagnostics: bubbles per execution tier of the last pf_align* call: out[0..4] = thread-per-bubble kernel, size classes
 * (longest branch <= 64/96/128/192/256); out[5] = warp-per-bubble kernel; out[6] = re-runs with the flag matrix in shared
 * memory (long co-optimal searches); out[7] = re-runs with the large limits */
int pf_align_last_tier_counts(const pf_ctx *ctx, uint32_t *out, int n);
/* diagnostics: bubbles of the last pf_align* call whose traceback exceeded the first-pass step budget and were handed to the
 * heavy queue (re-run beside the first pass by a few dedicated one-warp CTAs with the flag matrix in shared memory) */
uint32_t pf_align_last_heavy_queued(const pf_ctx *ctx);
/* diagnostics: bubbles of the last pf_align / pf_align_dev call that needed the large (tier-2) work area */
uint32_t pf_align_last_retry_count(const pf_ctx *ctx);
/* diagnostics: DP cells (m*n summed over every needlemanWunch fill) executed by the last pf_align* call */
uint64_t pf_align_last_cells(const pf_ctx *ctx);

/* ---- lookup phase B ------------------------------------------------------------------------------------------------ */
/*
 * The per-site part of the branching-bubble caller (CDBG.cpp:2295-2509; coloured: CCDBG.cpp:1057-1241): for every variable
 * column of every bubble of the LAST pf_align / pf_align_dev call on the database's context, the site k-mers are built on
 * the device from the aligned rows (still resident), the distinct k-mers of each allele class are looked up ('as written,
 * else reverse complement', CDBG.cpp:38-43) and gated with the strict (low, up) of readCov(string, lower, upper), and the
 * class coverages are returned.  `skip` (host, one byte per bubble, may be NULL) marks bubbles not to compute -- the strict
 * bubbles, whose class coverages are sums of branch means (CDBG.cpp:2105-2108).  `out` views pinned host memory owned by the
 * database handle, valid until the next pf_site_cov on it.
 */
int pf_site_cov(pf_kmc *db, uint32_t low, uint32_t up, const uint8_t *skip, pf_site_batch_t *out);
/* device-resident form: `d_skip` is a device pointer (or NULL), `out_dev` (may be NULL) receives DEVICE pointers; asynchronous on the stream */
int pf_site_cov_dev(pf_kmc *db, uint32_t low, uint32_t up, const void *d_skip, pf_site_batch_t *out_dev, void *cuda_stream);

/*
 * pf_align_staged -- pf_align over branches that are already on the device: the sequences first_seq, first_seq + 1, ... of the
 * batch that the PRECEDING pf_kmc_cov / pf_kmc_cov_async call on `db` copied there (zero-based offsets).  The reference looks
 * the branch unitigs up and then aligns the same strings (CDBG.cpp:2016-2036): a host that puts the entrance unitigs first and
 * the branches after them into ONE lookup batch sends every base once, and only bubble_off (relative to first_seq) travels here.
 * max_len / max_rows: longest branch and most branches per bubble of the batch (upper bounds are fine).  The staged batch
 * stays valid until the next host-pointer lookup call on `db`.  Result and ownership as pf_align.
 */
int pf_align_staged(pf_ctx *ctx, pf_kmc *db, double M, double D, double G, uint32_t first_seq, const uint32_t *bubble_off,
                    uint32_t n_bubbles, uint32_t max_len, uint32_t max_rows, pf_msa_batch_t *out);

/*
 * pf_site_kmers -- the site k-mers themselves, no database involved: the coloured caller (CCDBG.cpp:1057-1376) asks the GRAPH
 * which colours hold each site k-mer (findUnitig, :1127, :1258) before it reads any database, so it needs the strings.  Works on
 * the last pf_align of the context; `skip` as in pf_site_cov.  `out` views pinned memory of the context, valid until its next call.
 */
int pf_site_kmers(pf_ctx *ctx, uint32_t k, const uint8_t *skip, pf_site_kmers_t *out);

/*
 * pf_kmc_share -- a second handle on the same HBM-resident index for ANOTHER pf_ctx of the same device.  The reference keeps
 * one CKMCFile per CDBG and walks the graph with N threads (CDBG.cpp:1929-1945); here every host thread owns its own context -- stream,
 * staging, result arena -- and a shared handle, so the copy-in of one thread's batch overlaps the kernels and the copy-out of
 * another's.  Close the shared handle before the handle it was taken from.
 */
int pf_kmc_share(pf_kmc *db, pf_ctx *ctx, pf_kmc **out);

/*
 * pf_lookup_partition -- a CUDA stream confined to `n_sm` SMs of the context's device (a green context, CUDA 12.4+; n_sm is
 * rounded by the driver to its partition granularity, pf_lookup_partition_sms reports what was granted).  Pass it as the
 * stream of pf_kmc_lookup_dev while pf_align_dev of the same step runs on another stream: the random-access-bound lookups
 * keep to their SMs and the issue-bound alignment kernels fill the others at the same time, instead of each phase leaving the
 * other's resource idle.  The handle's own stream (pf_kmc_cov_async) uses such a partition when PF_LOOKUP_SMS=<n> is set.
 * One partition per context; it lives until pf_shutdown.  PF_E_UNSUPPORTED when the driver has no green contexts.
 */
int pf_lookup_partition(pf_ctx *ctx, uint32_t n_sm, void **stream_out);
uint32_t pf_lookup_partition_sms(const pf_ctx *ctx);

/* ---- roofline denominators measured on this device (bench.py reports them next to the kernels) ------- */
/* random 32-byte-sector gather rate over a `bytes`-sized table (GB/s of sectors touched) */
int pf_bench_random_gather(pf_ctx *ctx, uint64_t bytes, double *gb_per_s);
/* sustained INT32 ALU rate of an IADD3/VIMNMX/LOP3 mix (Gop/s, counting one op per lane-instruction) */
int pf_bench_int32(pf_ctx *ctx, double *gop_per_s);
/* diagnostics: random-access ceiling -- accesses of `width` (16 / 32 / 64) bytes, `ilp` (4 / 8 / 16) in flight per thread,
 * `ctas_per_sm` 256-thread CTAs per SM, over a power-of-two table of <= `bytes`; G accesses per second */
int pf_bench_gather_sweep(pf_ctx *ctx, uint64_t bytes, uint32_t width, uint32_t ilp, uint32_t ctas_per_sm, double *g_access_per_s);

#ifdef __cplusplus
}
#endif
#endif /* PF_GPU_H */
