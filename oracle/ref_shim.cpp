// ref_shim.cpp -- TEST INFRASTRUCTURE.  Thin extern "C" wrappers around the UNMODIFIED reference
// classes CKMCFile (KMC/kmc_api/kmc_file.h:105-167) and SeqAlign (src/SeqAlign.hpp:7-22).
// Compiled together with the reference's own sources, taken where they lie under /root/reference,
// into oracle/_ref/libpfref.so by oracle/Makefile.  No reference source is copied into this repo.
//
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / `--impl reference` legs may load
// this library; the product (libpfgpu.so) never does.
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <atomic>
#include <algorithm>

#include "kmc_file.h"    // -I /root/reference/KMC/kmc_api
#include "SeqAlign.hpp"  // -I /root/reference/src

#include "../include/pf_types.h"
#include "msa_pack.hpp"

namespace {

struct RefDb {
    CKMCFile f;
    uint32_t k = 0;
};

template <class F>
void parallel_for(size_t n, int n_threads, F fn) {
    if (n_threads <= 1 || n < 2) {
        for (size_t i = 0; i < n; i++) fn(i);
        return;
    }
    std::atomic<size_t> next(0);
    const size_t chunk = std::max<size_t>(1, n / (size_t(n_threads) * 16));
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++)
        th.emplace_back([&] {
            for (;;) {
                size_t b = next.fetch_add(chunk);
                if (b >= n) return;
                size_t e = std::min(n, b + chunk);
                for (size_t i = b; i < e; i++) fn(i);
            }
        });
    for (auto &t : th) t.join();
}

// One window, the three lookup dialects (pf_types.h PF_LOOKUP_*), through the reference API only.
inline bool ref_lookup_window(RefDb *db, CKmerAPI &km, const std::string &w, int mode, uint32_t &count) {
    count = 0;
    if (!km.from_string(w)) return false;  // non-ACGT: from_string refuses (kmer_api.h:502-509)
    if (mode == PF_LOOKUP_FWD) {
        uint32_t c = 0;
        bool ok = db->f.CheckKmer(km, c);
        if (ok) count = c;
        return ok;
    }
    if (mode == PF_LOOKUP_FWD_THEN_RC) {  // CDBG.cpp:38-43
        if (!db->f.IsKmer(km)) km.reverse();
        uint32_t c = 0;
        bool ok = db->f.CheckKmer(km, c);
        if (ok) count = c;
        return ok;
    }
    // canonical: min(kmer, rc) as GetCountersForRead does (kmc_file.cpp:1060)
    CKmerAPI rc(km);
    rc.reverse();
    uint32_t c = 0;
    bool ok = (km < rc) ? db->f.CheckKmer(km, c) : db->f.CheckKmer(rc, c);
    if (ok) count = c;
    return ok;
}

}  // namespace

extern "C" {

void *pfref_kmc_open(const char *prefix) {
    RefDb *db = new RefDb();
    if (!db->f.OpenForRA(prefix)) {
        delete db;
        return nullptr;
    }
    db->k = db->f.KmerLength();
    return db;
}

void pfref_kmc_close(void *h) {
    RefDb *db = (RefDb *)h;
    if (!db) return;
    db->f.Close();
    delete db;
}

int pfref_kmc_info(void *h, pf_kmc_info_t *out) {
    RefDb *db = (RefDb *)h;
    CKMCFileInfo info;
    if (!db->f.Info(info)) return -1;
    memset(out, 0, sizeof(*out));
    out->kmer_length = info.kmer_length;
    out->mode = info.mode;
    out->counter_size = info.counter_size;
    out->lut_prefix_length = info.lut_prefix_length;
    out->signature_len = info.signature_len;
    out->min_count = info.min_count;
    out->max_count = info.max_count;
    out->total_kmers = info.total_kmers;
    out->both_strands = info.both_strands ? 1 : 0;
    return 0;
}

int pfref_kmc_set_min_count(void *h, uint32_t x) { return ((RefDb *)h)->f.SetMinCount(x) ? 0 : -1; }
int pfref_kmc_set_max_count(void *h, uint32_t x) { return ((RefDb *)h)->f.SetMaxCount(x) ? 0 : -1; }

// counts[] / found[] hold, sequence after sequence, one entry per k-mer window (len-k+1, none if len<k).
// mode == PF_LOOKUP_CANONICAL && use_read_api: CKMCFile::GetCountersForRead (kmc_file.cpp:904).
int pfref_kmc_counts(void *h, const char *bases, const uint64_t *seq_off, uint32_t n_seq, int mode,
                     int use_read_api, int n_threads, uint32_t *counts, uint8_t *found) {
    RefDb *db = (RefDb *)h;
    const uint32_t k = db->k;
    std::vector<uint64_t> koff(n_seq + 1, 0);
    for (uint32_t s = 0; s < n_seq; s++) {
        uint64_t len = seq_off[s + 1] - seq_off[s];
        koff[s + 1] = koff[s] + (len >= k ? len - k + 1 : 0);
    }
    parallel_for(n_seq, n_threads, [&](size_t s) {
        uint64_t len = seq_off[s + 1] - seq_off[s];
        if (len < k) return;
        std::string read(bases + seq_off[s], len);
        uint32_t *c = counts + koff[s];
        uint8_t *f = found ? found + koff[s] : nullptr;
        if (use_read_api && mode == PF_LOOKUP_CANONICAL) {
            std::vector<uint32> v;
            db->f.GetCountersForRead(read, v);
            for (size_t i = 0; i < v.size(); i++) {
                c[i] = v[i];
                if (f) f[i] = v[i] != 0;
            }
            return;
        }
        CKmerAPI km(k);
        for (uint64_t i = 0; i + k <= len; i++) {
            uint32_t cnt;
            bool ok = ref_lookup_window(db, km, read.substr(i, k), mode, cnt);
            c[i] = ok ? cnt : 0;
            if (f) f[i] = ok;
        }
    });
    return 0;
}

// The readCov reductions (CDBG.cpp:29-120) with PloidyFrost's call pattern, one pf_cov_t per sequence.
int pfref_kmc_cov(void *h, const char *bases, const uint64_t *seq_off, uint32_t n_seq, int mode,
                  uint32_t low, uint32_t up, int n_threads, pf_cov_t *out) {
    RefDb *db = (RefDb *)h;
    const uint32_t k = db->k;
    parallel_for(n_seq, n_threads, [&](size_t s) {
        uint64_t len = seq_off[s + 1] - seq_off[s];
        pf_cov_t r;
        r.sum = 0;
        r.min = 10000;
        r.n_kmers = len >= k ? (uint32_t)(len - k + 1) : 0;
        r.first_missing = -1;
        r.first_outside = -1;
        std::string read(bases + seq_off[s], len);
        CKmerAPI km(k);
        for (uint32_t i = 0; i < r.n_kmers; i++) {
            uint32_t cnt;
            bool ok = ref_lookup_window(db, km, read.substr(i, k), mode, cnt);
            if (!ok) {
                if (r.first_missing < 0) r.first_missing = (int32_t)i;
                continue;
            }
            r.sum += cnt;
            if (cnt < r.min) r.min = cnt;
            if (!(cnt > low && cnt < up) && r.first_outside < 0) r.first_outside = (int32_t)i;
        }
        out[s] = r;
    });
    return 0;
}

// SeqAlign::SequenceAlignment over a bubble batch.  Returns an opaque handle owning the arrays *out views.
void *pfref_align(double M, double D, double G, const char *bases, const uint64_t *seq_off,
                  const uint32_t *bubble_off, uint32_t n_bubbles, int n_threads, pf_msa_batch_t *out) {
    std::vector<pforacle::MsaResult> res(n_bubbles);
    parallel_for(n_bubbles, n_threads, [&](size_t b) {
        double m = M, d = D, g = G;
        SeqAlign sa(m, d, g);
        std::vector<std::string> str;
        for (uint32_t s = bubble_off[b]; s < bubble_off[b + 1]; s++)
            str.emplace_back(bases + seq_off[s], seq_off[s + 1] - seq_off[s]);
        std::vector<uint> snp_pos, indel_pos, indel_len;
        std::vector<std::vector<unsigned short>> partition;
        sa.SequenceAlignment(str, snp_pos, indel_pos, partition, indel_len);
        pforacle::MsaResult &r = res[b];
        r.rows = str;
        if (!str.empty()) {
            r.snp_pos.assign(snp_pos.begin(), snp_pos.end());
            r.indel_pos.assign(indel_pos.begin(), indel_pos.end());
            r.indel_len.assign(indel_len.begin(), indel_len.end());
            r.partition = partition;
        }
    });
    pforacle::MsaPacked *p = new pforacle::MsaPacked();
    p->pack(res);
    p->view(out);
    return p;
}

void pfref_msa_free(void *h) { delete (pforacle::MsaPacked *)h; }

// Pairwise stage only: needlemanWunch (fill + traceback, SeqAlign.cpp:480) for one pair.
// Returns the number of co-optimal AlignUnits; writes them as str1\0str2\0 ... into buf (truncated at cap).
int pfref_nw_pair(double M, double D, double G, const char *a, const char *b, char *buf, size_t cap) {
    double m = M, d = D, g = G;
    SeqAlign sa(m, d, g);
    std::vector<AlignUnit> v = sa.needlemanWunch(std::string(a), std::string(b));
    size_t o = 0;
    for (auto &au : v) {
        if (o + au.str1.size() + au.str2.size() + 2 > cap) break;
        memcpy(buf + o, au.str1.c_str(), au.str1.size() + 1);
        o += au.str1.size() + 1;
        memcpy(buf + o, au.str2.c_str(), au.str2.size() + 1);
        o += au.str2.size() + 1;
    }
    return (int)v.size();
}

}  // extern "C"
