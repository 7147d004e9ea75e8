// pf_caller.hpp -- the per-bubble caller of PloidyFrost over the C ABI, batched (header-only, C++11; link with -pthread).
//
// What the reference does for ONE superbubble inside CDBG::PloidyEstimation -- look the branches up, order them, align them,
// turn every variable column into allele-class coverages and append rows to its output streams (strict bubbles:
// src/CDBG.cpp:1186-1330 = :1998-2189; branching bubbles: :1440-1660 = :2190-2575) -- done here for a whole batch of bubbles
// with three device calls (pf_kmc_cov, pf_align, pf_site_cov).  The graph side stays with the host program: it finds the
// superbubbles on its Bifrost graph and hands over, per bubble, the branch strings and the four numbers the rows need.
// The text produced is the reference's `-t 1` dialect, byte for byte (tests/dropin/caller_test.cpp compares it with the
// unmodified reference's own files, tests/golden/e2e).
#ifndef PF_CALLER_HPP
#define PF_CALLER_HPP

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "pf_gpu.h"

namespace pfdropin {

struct Bubble {
    unsigned entrance_id = 0, exit_id = 0;     // MyUnitig ids as the reference prints them (P_Unitig_Id.txt)
    size_t entrance_size = 0, exit_size = 0;   // UnitigMap::size of entrance / exit (bases); they bound VarDis (CDBG.cpp:2312-2330)
    bool strict = false;                       // MyUnitig::isStrict: every branch is one unitig
    // strict: mappedSequenceToString() of every successor, in successor order (CDBG.cpp:2016-2048);
    // branching: the strings from the last k-mer of the entrance to the first k-mer of the exit (:2226), any order
    std::vector<std::string> branches;
    // strict only, optional: referenceUnitigToString() of every branch -- the tie-break of sortSeq_simple (:482-551) when two
    // branches have exactly the same mean coverage; empty = the branch strings themselves
    std::vector<std::string> sort_keys;
};

struct CallerFiles {   // what the reference appends to its streams (Appendix D of SURVEY.md), `-t 1` dialect
    std::string alignseq;            // P_alignseq.txt
    std::string cov[4], fre[4];      // P_{bi,tri,tetra,penta}{cov,fre}.txt
    std::string allele_frequency;    // P_allele_frequency.txt
    size_t alleles[4] = {0, 0, 0, 0};
    size_t bubbles_called = 0;
    std::vector<unsigned char> called;   // appended per input bubble: 1 = it was aligned and got a VarId
};

class BubbleCaller {
  public:
    BubbleCaller(pf_ctx *ctx, pf_kmc *db, double match, double mismatch, double gap, unsigned lower, unsigned upper)
        : ctx_(ctx), db_(db), M_(match), D_(mismatch), G_(gap), lower_(lower), upper_(upper) {}

    const std::string &error() const { return err_; }

    // false (default): the `-t 1` files (CDBG::ploidyEstimation_ptr).  true: the `-t N` files (ploidyEstimation_multithread_ptr):
    // P_allele_frequency.txt holds, per bubble, the frequencies of its 2-allele sites, then 3-, then 4- (and 5- for strict bubbles)
    // and nothing for sites with more classes (CDBG.cpp:2162, :2550); start `var_id` at 0 there (fetch_add, :2056).  The bubbles
    // come out in the order they went in -- one of the schedules the reference's worker threads can produce.
    void set_thread_dialect(bool multithread) { mt_ = multithread; }
    // host threads that format the rows of a batch (contiguous ranges of its bubbles, joined in order: the text does not depend on it)
    void set_host_threads(unsigned n) { host_threads_ = n ? n : 1; }

    // Calls one batch.  `var_id` is the reference's running variant counter (var_count_all; start it at 1 for the `-t 1` files)
    // and advances by one for every bubble whose alignment is not empty.  Returns false where the reference would have ended
    // the program (a k-mer of a branch or of a site is not in the database, CDBG.cpp:52-56) or on a device error; error() says which,
    // no text of the failed batch is appended to `out` (its `called` flags and `var_id` are not meaningful then).
    bool call(const std::vector<Bubble> &batch, size_t &var_id, CallerFiles &out) { return call(batch.data(), batch.size(), var_id, out); }

    bool call(const Bubble *batch, size_t n_batch, size_t &var_id, CallerFiles &out) {
        err_.clear();
        const size_t called_base = out.called.size();
        out.called.resize(called_base + n_batch, 0);
        // ---- lookup-A: readCov of every branch of the strict bubbles (CDBG.cpp:66-120) ----
        std::string lbases;
        std::vector<uint64_t> loff(1, 0);
        for (size_t bi = 0; bi < n_batch; bi++)
            if (batch[bi].strict)
                for (const std::string &s : batch[bi].branches) { lbases += s; loff.push_back(lbases.size()); }
        std::vector<pf_cov_t> cov(loff.size() - 1);
        if (!cov.empty() && pf_kmc_cov(db_, lbases.data(), loff.data(), (uint32_t)cov.size(), PF_LOOKUP_FWD_THEN_RC, 0, 0xFFFFFFFFu,
                                       cov.data()) != PF_OK)
            return fail(pf_last_error());
        // ---- gate + order the branches; build the alignment batch ----
        struct Kept { size_t src; std::vector<size_t> order; std::vector<double> means; double sum; };
        std::vector<Kept> kept;
        std::string abases;
        std::vector<uint64_t> aoff(1, 0);
        std::vector<uint32_t> boff(1, 0);
        std::vector<uint8_t> skip;
        size_t ci = 0;
        for (size_t bi = 0; bi < n_batch; bi++) {
            const Bubble &b = batch[bi];
            Kept kb;
            kb.src = bi; kb.sum = 0;
            const size_t n = b.branches.size();
            if (b.strict) {
                bool ok = true;
                for (size_t j = 0; j < n; j++, ci++) {
                    const pf_cov_t &c = cov[ci];
                    if (!ok) continue;                                             // the reference stopped reading at the first failure (:2027-2031)
                    if (c.first_missing >= 0) return fail("a k-mer of a branch unitig is not in the database: the reference exits here (CDBG.cpp:94)");
                    if (c.min > lower_ && c.min < upper_) {                       // :2021 (min count of the branch inside the thresholds)
                        const double m = (double)c.sum / (double)c.n_kmers;
                        kb.means.push_back(m);
                        kb.sum += m;                                               // in successor order (:2024)
                    } else ok = false;                                             // :2027-2031: the bubble is dropped
                }
                if (!ok || n < 2) continue;
                kb.order.resize(n);
                for (size_t j = 0; j < n; j++) kb.order[j] = j;
                const std::vector<std::string> &keys = b.sort_keys.size() == n ? b.sort_keys : b.branches;
                std::sort(kb.order.begin(), kb.order.end(), [&](size_t x, size_t y) {          // sortSeq_simple's order (:482-551)
                    if (kb.means[x] != kb.means[y]) return kb.means[x] > kb.means[y];
                    return std::strcmp(keys[x].c_str(), keys[y].c_str()) > 0;
                });
            } else {
                if (n < 2) continue;
                kb.order.resize(n);
                for (size_t j = 0; j < n; j++) kb.order[j] = j;
                std::sort(kb.order.begin(), kb.order.end(), [&](size_t x, size_t y) {          // sortSeq_branching: longer first, then larger
                    if (b.branches[x].size() != b.branches[y].size()) return b.branches[x].size() > b.branches[y].size();
                    return b.branches[x] > b.branches[y];
                });
            }
            for (size_t j : kb.order) { abases += b.branches[j]; aoff.push_back(abases.size()); }
            boff.push_back((uint32_t)(aoff.size() - 1));
            skip.push_back(b.strict ? 1 : 0);
            kept.push_back(std::move(kb));
        }
        if (kept.empty()) return true;
        // ---- SequenceAlignment of every kept bubble, then the site k-mers of the branching ones ----
        pf_msa_batch_t m;
        if (pf_align(ctx_, M_, D_, G_, abases.data(), aoff.data(), boff.data(), (uint32_t)kept.size(), &m) != PF_OK) return fail(pf_last_error());
        pf_site_batch_t sc;
        if (pf_site_cov(db_, lower_, upper_, skip.data(), &sc) != PF_OK) return fail(pf_last_error());
        // ---- ids: the bubbles whose alignment is not empty take consecutive ids in batch order (:2051, :2273) ----
        std::vector<size_t> ids(kept.size(), 0);
        for (size_t q = 0; q < kept.size(); q++) {
            if (m.status[q] != PF_BUBBLE_OK) return fail("a bubble does not fit the device limits (status " + std::to_string(m.status[q]) + ")");
            if (m.n_rows[q] == 0) continue;                                        // str_vec came back empty
            ids[q] = var_id++;
            out.bubbles_called++;
            out.called[called_base + kept[q].src] = 1;
        }
        // ---- rows: contiguous ranges of the kept bubbles, one host thread each, text joined in order ----
        auto emit = [&](size_t q0, size_t q1, CallerFiles &to, std::string &why) -> bool {
            for (size_t q = q0; q < q1; q++) {
                const Kept &kb = kept[q];
                const Bubble &b = batch[kb.src];
                const uint32_t nr = m.n_rows[q], L = m.aln_len[q];
                if (nr == 0) continue;
                const size_t var_count = ids[q];
                const char *rows = m.rows + m.rows_off[q];
                char head[96];
                const int head_len = std::snprintf(head, sizeof head, "%zu\t%d\t%u\t%u\t", var_count, b.strict ? 1 : 0, b.entrance_id, b.exit_id);
                for (uint32_t r = 0; r < nr; r++) {
                    to.alignseq.append(head, (size_t)head_len);
                    to.alignseq.append(rows + (size_t)r * L, L);
                    to.alignseq += "\n";
                }
                const uint64_t v0 = m.var_off[q], v1 = m.var_off[q + 1];
                const size_t n_var = (size_t)(v1 - v0);
                const uint16_t *cls = m.cls + m.cls_off[q];
                const uint32_t *ilen = m.ilen + m.ilen_off[q];
                const size_t n_ilen = (size_t)(m.ilen_off[q + 1] - m.ilen_off[q]);
                size_t indel = 0;
                std::string grouped_fre[4], cov_info, fre_info;
                for (size_t i = 0; i < n_var; i++) {
                    const bool is_indel = m.var_kind[v0 + i] == 1;
                    size_t var_distance;                                           // :2312-2330
                    auto gap_to = [&](size_t a, size_t c) { return (size_t)(m.var_col[v0 + c] - m.var_col[v0 + a] - 1); };
                    if (i == 0) var_distance = n_var > 1 ? std::min(gap_to(0, 1), b.entrance_size) : std::min(b.entrance_size, b.exit_size);
                    else if (i == n_var - 1) var_distance = std::min(gap_to(i - 1, i), b.exit_size);
                    else var_distance = std::min(gap_to(i - 1, i), gap_to(i, i + 1));
                    unsigned maxnum = 0;
                    for (uint32_t r = 0; r < nr; r++) maxnum = std::max<unsigned>(maxnum, cls[i * nr + r]);
                    std::vector<double> tc(maxnum, 0.0);
                    double sum = 0;
                    if (is_indel) indel++;                                         // :2390 / strict :2127, before anything can skip the site
                    if (b.strict) {
                        for (uint32_t r = 0; r < nr; r++) tc[cls[i * nr + r] - 1] += kb.means[kb.order[r]];   // :2105-2108
                        sum = kb.sum;
                    } else {
                        const uint8_t st = sc.status[sc.site_off[q] + i];
                        if (st == PF_SITE_DROPPED) continue;                       // :2415-2418
                        if (st == PF_SITE_MISSING) { why = "a site k-mer is not in the database: the reference exits here (CDBG.cpp:54)"; return false; }
                        if (st != PF_SITE_OK) { why = "a site k-mer cannot be formed (the reference reads outside the aligned row here)"; return false; }
                        const uint64_t *cv = sc.cov + sc.cov_off[q] + i * nr;
                        for (unsigned c = 0; c < maxnum; c++) { tc[c] = (double)cv[c]; sum += tc[c]; }
                    }
                    cov_info.clear();
                    fre_info.clear();
                    for (double c : tc) {                                          // `stream << double`: precision 6, general format == %g
                        put_double(cov_info, c); cov_info += '\t';
                        put_double(fre_info, c / sum); fre_info += '\n';
                    }
                    const uint32_t il = is_indel ? (indel - 1 < n_ilen ? ilen[indel - 1] : 0u) : 0u;
                    char tail[128];
                    const int tail_len = std::snprintf(tail, sizeof tail, "%d\t%u\t%zu\t%zu\t%zu\t\n", b.strict ? 1 : 0, il, var_count, n_var, var_distance);
                    cov_info.append(tail, (size_t)tail_len);
                    if (!mt_) to.allele_frequency += fre_info;                     // -t 1: every site in site order (:1318, :1630)
                    if (maxnum >= 2 && maxnum <= 5) {                              // switch (maxnum), :1319-1340 / :2126-2147
                        to.alleles[maxnum - 2]++;
                        to.cov[maxnum - 2] += cov_info;
                        to.fre[maxnum - 2] += fre_info;
                        if (mt_) grouped_fre[maxnum - 2] += fre_info;
                    }
                }
                if (mt_) {                                                         // -t N: grouped per bubble (:2162 strict, :2550 branching)
                    to.allele_frequency += grouped_fre[0] + grouped_fre[1] + grouped_fre[2];
                    if (b.strict) to.allele_frequency += grouped_fre[3];
                }
            }
            return true;
        };
        const size_t T = std::max<size_t>(1, std::min<size_t>(host_threads_, kept.size() / 64));
        std::vector<CallerFiles> part(T);
        std::vector<std::string> why(T);
        std::vector<char> ok(T, 1);
        if (T == 1) ok[0] = emit(0, kept.size(), part[0], why[0]);
        else {
            std::vector<std::thread> workers;
            for (size_t t = 0; t < T; t++)
                workers.emplace_back([&, t] { ok[t] = emit(kept.size() * t / T, kept.size() * (t + 1) / T, part[t], why[t]); });
            for (std::thread &w : workers) w.join();
        }
        for (size_t t = 0; t < T; t++)
            if (!ok[t]) return fail(why[t]);
        for (size_t t = 0; t < T; t++) {
            out.alignseq += part[t].alignseq;
            out.allele_frequency += part[t].allele_frequency;
            for (int a = 0; a < 4; a++) { out.cov[a] += part[t].cov[a]; out.fre[a] += part[t].fre[a]; out.alleles[a] += part[t].alleles[a]; }
        }
        return true;
    }

  private:
    static void put_double(std::string &to, double v) {
        char buf[40];
        to.append(buf, (size_t)std::snprintf(buf, sizeof buf, "%g", v));
    }
    bool fail(const std::string &why) { err_ = why; return false; }
    pf_ctx *ctx_;
    pf_kmc *db_;
    double M_, D_, G_;
    unsigned lower_, upper_;
    bool mt_ = false;
    unsigned host_threads_ = 1;
    std::string err_;
};

}  // namespace pfdropin
#endif  // PF_CALLER_HPP
