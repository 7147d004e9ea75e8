// pf_dropin.hpp -- header-only C++ adapters that present libpfgpu.so (include/pf_gpu.h) with the shapes of the
// two reference classes this path sits behind, so that reference-side code can be switched over call by call:
//
//   pfdropin::KmcFile   ~ CKMCFile  (KMC/kmc_api/kmc_file.h:32-167): OpenForRA :105, Close :117, SetMinCount :120,
//                                    SetMaxCount :126, GetBothStrands :132, KmerLength :138, CheckKmer :149,
//                                    IsKmer :154, Info :163, GetCountersForRead :166
//   pfdropin::SeqAlign  ~ SeqAlign  (src/SeqAlign.hpp:7-22): ctor :10, SequenceAlignment :16
//
// Same member names, argument meaning and error behaviour (bool returns, no exceptions; an alignment that the
// reference drops comes back as an empty `str`).  K-mers are passed as std::string because CKmerAPI is a KMC type;
// a maintainer of the reference calls `kmer.to_string()` (kmer_api.h:433) at the seam -- see INTEGRATION.md.
// Every call is a GPU round trip: use the *Batch members (one call per batch of reads / bubbles) on hot loops.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "pf_gpu.h"

namespace pfdropin {

class Device {  // one pf_ctx per process/GPU, shared by the adapters
  public:
    static pf_ctx *get(int device = -1) {
        static pf_ctx *ctx = nullptr;
        if (!ctx) {
            if (device < 0) device = 0;
            if (pf_init(device, &ctx) != PF_OK) {
                fprintf(stderr, "pfdropin: %s\n", pf_last_error());
                ctx = nullptr;
            }
        }
        return ctx;
    }
};

struct KmcFileInfo {  // CKMCFileInfo, kmc_file.h:19-30
    uint32_t kmer_length, mode, counter_size, lut_prefix_length, signature_len, min_count;
    uint64_t max_count;
    bool both_strands;
    uint64_t total_kmers;
};

class KmcFile {
    pf_kmc *db_ = nullptr;
    pf_kmc_info_t info_{};

    static void flatten(const std::vector<std::string> &seqs, std::string &bases, std::vector<uint64_t> &off) {
        off.assign(1, 0);
        bases.clear();
        for (const std::string &s : seqs) {
            bases += s;
            off.push_back(bases.size());
        }
    }

  public:
    KmcFile() {}
    ~KmcFile() { Close(); }
    KmcFile(const KmcFile &) = delete;
    KmcFile &operator=(const KmcFile &) = delete;

    bool OpenForRA(const std::string &file_name) {
        if (db_) return false;  // kmc_file.cpp:31: already opened
        pf_ctx *ctx = Device::get();
        if (!ctx || pf_kmc_open(ctx, file_name.c_str(), &db_) != PF_OK) return false;
        pf_kmc_info(db_, &info_);
        return true;
    }
    bool Close() {
        if (!db_) return false;
        pf_kmc_close(db_);
        db_ = nullptr;
        return true;
    }
    bool Info(KmcFileInfo &i) const {
        if (!db_) return false;
        i.kmer_length = info_.kmer_length; i.mode = info_.mode; i.counter_size = info_.counter_size;
        i.lut_prefix_length = info_.lut_prefix_length; i.signature_len = info_.signature_len;
        i.min_count = info_.min_count; i.max_count = info_.max_count; i.both_strands = info_.both_strands != 0;
        i.total_kmers = info_.total_kmers;
        return true;
    }
    uint32_t KmerLength() const { return info_.kmer_length; }
    bool GetBothStrands() const { return info_.both_strands != 0; }
    bool SetMinCount(uint32_t x) { return db_ && pf_kmc_set_min_count(db_, x) == PF_OK && pf_kmc_info(db_, &info_) == PF_OK; }
    bool SetMaxCount(uint32_t x) { return db_ && pf_kmc_set_max_count(db_, x) == PF_OK && pf_kmc_info(db_, &info_) == PF_OK; }

    // CheckKmer(CKmerAPI&, uint32&): looks the k-mer up AS WRITTEN (kmc_file.cpp:330)
    bool CheckKmer(const std::string &kmer, uint32_t &count) {
        if (!db_ || kmer.size() != info_.kmer_length) return false;
        const uint64_t off[2] = {0, kmer.size()};
        uint32_t c = 0;
        uint8_t f = 0;
        if (pf_kmc_counts(db_, kmer.data(), off, 1, PF_LOOKUP_FWD, &c, &f) != PF_OK) return false;
        if (f) count = c;
        return f != 0;
    }
    bool IsKmer(const std::string &kmer) {  // kmc_file.cpp:775
        uint32_t c;
        return CheckKmer(kmer, c);
    }
    // GetCountersForRead (kmc_file.cpp:904): false when the read is shorter than k (:909-913)
    bool GetCountersForRead(const std::string &read, std::vector<uint32_t> &counters) {
        if (!db_) return false;
        if (read.size() < info_.kmer_length) { counters.clear(); return false; }
        counters.assign(read.size() - info_.kmer_length + 1, 0);
        const uint64_t off[2] = {0, read.size()};
        return pf_kmc_counts(db_, read.data(), off, 1, PF_LOOKUP_CANONICAL, counters.data(), nullptr) == PF_OK;
    }
    // one GPU call for many reads
    bool GetCountersForReadsBatch(const std::vector<std::string> &reads, std::vector<std::vector<uint32_t>> &counters) {
        if (!db_) return false;
        std::string bases;
        std::vector<uint64_t> off;
        flatten(reads, bases, off);
        std::vector<uint64_t> woff(reads.size() + 1);
        const uint64_t W = pf_window_offsets(off.data(), (uint32_t)reads.size(), info_.kmer_length, woff.data());
        std::vector<uint32_t> flat(W ? W : 1);
        if (pf_kmc_counts(db_, bases.data(), off.data(), (uint32_t)reads.size(), PF_LOOKUP_CANONICAL, flat.data(), nullptr) != PF_OK)
            return false;
        counters.resize(reads.size());
        for (size_t r = 0; r < reads.size(); r++) counters[r].assign(flat.begin() + woff[r], flat.begin() + woff[r + 1]);
        return true;
    }
    // CDBG::readCov semantics for many sequences at once (CDBG.cpp:29-120): see pf_kmc_cov
    bool ReadCovBatch(const std::vector<std::string> &seqs, uint32_t low, uint32_t up, std::vector<pf_cov_t> &out) {
        if (!db_) return false;
        std::string bases;
        std::vector<uint64_t> off;
        flatten(seqs, bases, off);
        out.resize(seqs.size());
        if (seqs.empty()) return true;
        return pf_kmc_cov(db_, bases.data(), off.data(), (uint32_t)seqs.size(), PF_LOOKUP_FWD_THEN_RC, low, up, out.data()) == PF_OK;
    }
};

struct MsaOut {  // the four output vectors of SeqAlign::SequenceAlignment for one bubble
    std::vector<unsigned int> snp_pos, indel_pos, indel_len;
    std::vector<std::vector<unsigned short>> partition;
};

class SeqAlign {
    double M_, D_, G_;

    static void unpack(const pf_msa_batch_t &r, uint32_t b, std::vector<std::string> &str, MsaOut &o) {
        const uint32_t nr = r.n_rows[b], L = r.aln_len[b];
        str.clear();
        o = MsaOut();
        if (!nr) return;  // the reference returns str empty (CDBG.cpp:2051)
        for (uint32_t q = 0; q < nr; q++) str.emplace_back(r.rows + r.rows_off[b] + (uint64_t)q * L, L);
        o.partition.assign(L, std::vector<unsigned short>(nr, 0));
        const uint64_t v0 = r.var_off[b], v1 = r.var_off[b + 1];
        for (uint64_t v = v0; v < v1; v++) {
            const uint32_t col = r.var_col[v];
            if (r.var_kind[v] == 0) o.snp_pos.push_back(col);
            else if (r.var_kind[v] == 1) o.indel_pos.push_back(col);
            for (uint32_t q = 0; q < nr; q++) o.partition[col][q] = r.cls[r.cls_off[b] + (v - v0) * nr + q];
        }
        o.indel_len.assign(r.ilen + r.ilen_off[b], r.ilen + r.ilen_off[b + 1]);
    }

  public:
    SeqAlign(double &match, double &dismatch, double &gap) : M_(match), D_(dismatch), G_(gap) {}  // SeqAlign.hpp:10

    // SeqAlign::SequenceAlignment (SeqAlign.hpp:16): in: raw sequences, out: aligned rows or empty
    void SequenceAlignment(std::vector<std::string> &str, std::vector<unsigned int> &snp_pos, std::vector<unsigned int> &indel_pos,
                           std::vector<std::vector<unsigned short>> &partition, std::vector<unsigned int> &indel_len) {
        std::vector<std::vector<std::string>> one(1, str);
        std::vector<MsaOut> out;
        std::vector<int> status;
        if (!SequenceAlignmentBatch(one, out, status) || status[0] != PF_BUBBLE_OK) { str.clear(); return; }
        str = one[0];
        snp_pos = out[0].snp_pos; indel_pos = out[0].indel_pos; partition = out[0].partition; indel_len = out[0].indel_len;
    }

    // one GPU call for many bubbles; bubbles[b] is replaced by its aligned rows (empty = dropped, as in the reference)
    bool SequenceAlignmentBatch(std::vector<std::vector<std::string>> &bubbles, std::vector<MsaOut> &out, std::vector<int> &status) {
        pf_ctx *ctx = Device::get();
        if (!ctx) return false;
        std::string bases;
        std::vector<uint64_t> off(1, 0);
        std::vector<uint32_t> boff(1, 0);
        for (const auto &b : bubbles) {
            for (const std::string &s : b) { bases += s; off.push_back(bases.size()); }
            boff.push_back((uint32_t)off.size() - 1);
        }
        pf_msa_batch_t r;
        if (pf_align(ctx, M_, D_, G_, bases.data(), off.data(), boff.data(), (uint32_t)bubbles.size(), &r) != PF_OK) return false;
        out.resize(bubbles.size());
        status.resize(bubbles.size());
        for (uint32_t b = 0; b < bubbles.size(); b++) {
            status[b] = r.status[b];
            unpack(r, b, bubbles[b], out[b]);
        }
        return true;
    }
};

}  // namespace pfdropin
