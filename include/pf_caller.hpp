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
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <set>
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

// A batch of bubbles as flat arrays -- what a host that cares about its own speed hands over (no std::string per branch).
// Branch s of the batch is bases[seq_off[s] .. seq_off[s+1]); bubble b owns branches bubble_off[b] .. bubble_off[b+1].
struct FlatBatch {
    std::string bases;
    std::vector<uint64_t> seq_off{0};
    std::vector<uint32_t> bubble_off{0};
    std::vector<uint8_t> fwd;                  // per branch, strict bubbles only: 1 = the string is the unitig's forward strand
                                               // (referenceUnitigToString() == the string), 0 = its reverse complement
    std::vector<uint8_t> strict;               // per bubble
    std::vector<uint32_t> entrance_id, exit_id;
    std::vector<uint64_t> entrance_size, exit_size;
    size_t n_bubbles() const { return bubble_off.size() - 1; }
    size_t n_seq() const { return seq_off.size() - 1; }
    void clear() {
        bases.clear(); seq_off.assign(1, 0); bubble_off.assign(1, 0); fwd.clear(); strict.clear();
        entrance_id.clear(); exit_id.clear(); entrance_size.clear(); exit_size.clear();
    }
    void add_branch(const char *p, size_t n, bool forward) { bases.append(p, n); seq_off.push_back(bases.size()); fwd.push_back(forward ? 1 : 0); }
    void end_bubble(bool is_strict, unsigned ent_id, unsigned ex_id, size_t ent_size, size_t ex_size) {
        bubble_off.push_back((uint32_t)n_seq()); strict.push_back(is_strict ? 1 : 0);
        entrance_id.push_back(ent_id); exit_id.push_back(ex_id); entrance_size.push_back(ent_size); exit_size.push_back(ex_size);
    }
    void append(const FlatBatch &o) {          // batches built by several threads are joined in order
        const uint64_t b0 = bases.size();
        const uint32_t s0 = (uint32_t)n_seq();
        bases += o.bases;
        for (size_t i = 1; i < o.seq_off.size(); i++) seq_off.push_back(o.seq_off[i] + b0);
        for (size_t i = 1; i < o.bubble_off.size(); i++) bubble_off.push_back(o.bubble_off[i] + s0);
        fwd.insert(fwd.end(), o.fwd.begin(), o.fwd.end());
        strict.insert(strict.end(), o.strict.begin(), o.strict.end());
        entrance_id.insert(entrance_id.end(), o.entrance_id.begin(), o.entrance_id.end());
        exit_id.insert(exit_id.end(), o.exit_id.begin(), o.exit_id.end());
        entrance_size.insert(entrance_size.end(), o.entrance_size.begin(), o.entrance_size.end());
        exit_size.insert(exit_size.end(), o.exit_size.begin(), o.exit_size.end());
    }
};

struct CallerFiles {   // what the reference appends to its streams (Appendix D of SURVEY.md), `-t 1` dialect
    std::string alignseq;            // P_alignseq.txt
    std::string cov[4], fre[4];      // P_{bi,tri,tetra,penta}{cov,fre}.txt
    std::string allele_frequency;    // P_allele_frequency.txt
    size_t alleles[4] = {0, 0, 0, 0};
    size_t bubbles_called = 0;
    std::vector<unsigned char> called;   // appended per input bubble: 1 = it was aligned and got a VarId
};

// What SeqAlign::SequenceAlignment hands back for ONE bubble (SeqAlign.hpp:16): the result type of the host aligner a program may
// give the caller for the bubbles that exceed a device limit (more than 64 co-optimal alignments / candidate MSAs, 64 rows, an
// alignment beyond the work area; pf_msa_batch_t status != 0) -- inside the reference program that aligner is the reference's own
// SeqAlign (integration/ploidy_estimation_gpu.cpp).  Without one such a bubble ends the batch with an error, as before.
struct HostMsa {
    std::vector<std::string> rows;                          // `str` after the call: aligned rows, or empty
    std::vector<unsigned> snp_pos, indel_pos, indel_len;
    std::vector<std::vector<unsigned short>> partition;     // per column
};
typedef std::function<void(std::vector<std::string> &str, HostMsa &out)> HostAligner;

// The k-mer every row contributes at variable column c (CDBG.cpp:2338-2388 indel sites, :2433-2472 SNP sites), on the host: used for
// the few bubbles that went through the host aligner (the device builds them from the aligned rows it holds, pf_site_cov).  false
// where the reference itself would read outside a row.
inline bool host_site_kmers(const std::vector<std::string> &rows, size_t c, size_t k, bool is_indel, size_t n_indel_before, std::vector<std::string> &out) {
    const size_t n = rows.size(), L = rows[0].size();
    out.assign(n, "");
    if (is_indel) {
        std::vector<size_t> cur(n, c);
        std::vector<std::string> ext(n);
        for (;;) {                                                   // extend to the right until the rows differ (:2338-2357)
            std::set<char> chars;
            for (size_t r = 0; r < n; r++) {
                while (cur[r] < L && rows[r][cur[r]] == '-') cur[r]++;
                if (cur[r] >= L) return false;
                const char ch = rows[r][cur[r]++];
                ext[r] += ch;
                chars.insert(ch);
            }
            if (chars.size() > 1) break;
        }
        for (size_t r = 0; r < n; r++) {
            const size_t e = ext[r].size();
            if (e > k) return false;
            if (n_indel_before == 0) {                               // left-pad from the aligned row itself (:2358-2365)
                if (c + e < k) return false;
                out[r] = rows[r].substr(c - k + e, k - e) + ext[r];
            } else {                                                 // left-pad from the row with its gaps removed (:2366-2388)
                std::string t;
                for (size_t x = 0; x < c; x++) if (rows[r][x] != '-') t += rows[r][x];
                if (t.size() < k - e) {
                    std::string s = t + ext[r];
                    for (size_t x = cur[r]; s.size() < k; x++) {
                        if (x >= L) return false;
                        if (rows[r][x] != '-') s += rows[r][x];
                    }
                    out[r] = s;
                } else out[r] = t.substr(t.size() - (k - e)) + ext[r];
            }
        }
        return true;
    }
    if (n_indel_before > 0) {                                        // SNP site after an indel site (:2433-2465)
        for (size_t r = 0; r < n; r++) {
            std::string t;
            for (size_t x = 0; x <= c; x++) if (rows[r][x] != '-') t += rows[r][x];
            if (t.size() < k) {
                for (size_t x = c + 1; t.size() < k; x++) {
                    if (x >= L) return false;
                    if (rows[r][x] != '-') t += rows[r][x];
                }
                out[r] = t;
            } else out[r] = t.substr(t.size() - k);
        }
        return true;
    }
    if (c + 1 < k) return false;
    for (size_t r = 0; r < n; r++) out[r] = rows[r].substr(c - k + 1, k);   // the k-mer ending at the site (:2469-2472)
    return true;
}

struct CallerStats {     // where the time of BubbleCaller::call went (seconds, summed over the calls)
    double lookup_s = 0, gate_s = 0, align_s = 0, site_s = 0, emit_s = 0;
    size_t calls = 0, bubbles_in = 0, bubbles_aligned = 0, bubbles_host_aligned = 0;
};

class BubbleCaller {
  public:
    const CallerStats &stats() const { return stats_; }
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
    // the aligner for bubbles beyond a device limit (see HostMsa); PF_CALLER_FORCE_HOST=<n> in the environment sends every n-th aligned
    // bubble through it as well (tests)
    void set_host_aligner(HostAligner f) { host_aligner_ = std::move(f); }

    // Calls one batch.  `var_id` is the reference's running variant counter (var_count_all; start it at 1 for the `-t 1` files)
    // and advances by one for every bubble whose alignment is not empty.  Returns false where the reference would have ended
    // the program (a k-mer of a branch or of a site is not in the database, CDBG.cpp:52-56) or on a device error; error() says which,
    // no text of the failed batch is appended to `out` (its `called` flags and `var_id` are not meaningful then).
    bool call(const std::vector<Bubble> &batch, size_t &var_id, CallerFiles &out) { return call(batch.data(), batch.size(), var_id, out); }

    // the std::string form: flattened and handed to the flat form below
    bool call(const Bubble *batch, size_t n_batch, size_t &var_id, CallerFiles &out) {
        FlatBatch fb;
        size_t total = 0;
        for (size_t bi = 0; bi < n_batch; bi++) for (const std::string &s : batch[bi].branches) total += s.size();
        fb.bases.reserve(total);
        std::vector<std::string> keys;             // explicit sort keys of strict bubbles (referenceUnitigToString), else empty
        bool any_keys = false;
        for (size_t bi = 0; bi < n_batch; bi++) {
            const Bubble &b = batch[bi];
            const bool has_keys = b.strict && b.sort_keys.size() == b.branches.size();
            for (size_t j = 0; j < b.branches.size(); j++) {
                fb.add_branch(b.branches[j].data(), b.branches[j].size(), true);
                keys.push_back(has_keys ? b.sort_keys[j] : std::string());
                any_keys |= has_keys;
            }
            fb.end_bubble(b.strict, b.entrance_id, b.exit_id, b.entrance_size, b.exit_size);
        }
        explicit_keys_ = any_keys ? &keys : nullptr;
        const bool ok = call(fb, var_id, out);
        explicit_keys_ = nullptr;
        return ok;
    }

    // Calls one flat batch (see FlatBatch).
    bool call(const FlatBatch &fb, size_t &var_id, CallerFiles &out) {
        err_.clear();
        const size_t n_batch = fb.n_bubbles(), n_seq = fb.n_seq();
        const size_t called_base = out.called.size();
        out.called.resize(called_base + n_batch, 0);
        if (n_batch == 0) return true;
        auto fail_batch = [&](const std::string &why) { out.called.resize(called_base); return fail(why); };
        auto t_mark = std::chrono::steady_clock::now();
        auto lap = [&](double &acc) {
            const auto now = std::chrono::steady_clock::now();
            acc += std::chrono::duration<double>(now - t_mark).count();
            t_mark = now;
        };
        stats_.calls++; stats_.bubbles_in += n_batch;
        const char *B = fb.bases.data();
        auto seq_ptr = [&](size_t s) { return B + fb.seq_off[s]; };
        auto seq_len = [&](size_t s) { return (size_t)(fb.seq_off[s + 1] - fb.seq_off[s]); };
        // ---- lookup-A: readCov of every branch string (CDBG.cpp:66-120).  The reference reads the branches of STRICT bubbles only;
        //      the records of the other bubbles' paths come back with the same call and are not looked at. ----
        bool any_strict = false;
        for (size_t bi = 0; bi < n_batch && !any_strict; bi++) any_strict = fb.strict[bi] != 0;
        cov_.resize(n_seq);
        if (any_strict && n_seq && pf_kmc_cov(db_, B, fb.seq_off.data(), (uint32_t)n_seq, PF_LOOKUP_FWD_THEN_RC, 0, 0xFFFFFFFFu, cov_.data()) != PF_OK)
            return fail_batch(pf_last_error());
        lap(stats_.lookup_s);
        // ---- gate + order the branches; build the alignment batch ----
        k_src_.clear(); k_first_.assign(1, 0); k_sum_.clear(); sorted_seq_.clear(); sorted_mean_.clear();
        abases_.clear(); aoff_.assign(1, 0); boff_.assign(1, 0); skip_.clear();
        abases_.reserve(fb.bases.size());
        std::vector<double> mean_tmp;
        std::vector<uint32_t> ord;
        std::string key_x, key_y;
        for (size_t bi = 0; bi < n_batch; bi++) {
            const size_t s0 = fb.bubble_off[bi], n = fb.bubble_off[bi + 1] - s0;
            const bool strict = fb.strict[bi] != 0;
            double sum = 0;
            mean_tmp.assign(n, 0.0);
            if (strict) {
                bool ok = true;
                for (size_t j = 0; j < n && ok; j++) {                               // the reference stops reading at the first failure (:2027-2031)
                    const pf_cov_t &c = cov_[s0 + j];
                    if (c.first_missing >= 0) return fail_batch("a k-mer of a branch unitig is not in the database: the reference exits here (CDBG.cpp:94)");
                    if (c.min > lower_ && c.min < upper_) {                         // :2021 (min count of the branch inside the thresholds)
                        mean_tmp[j] = (double)c.sum / (double)c.n_kmers;
                        sum += mean_tmp[j];                                         // in successor order (:2024)
                    } else ok = false;                                              // :2027-2031: the bubble is dropped
                }
                if (!ok || n < 2) continue;
            } else if (n < 2) continue;
            ord.resize(n);
            for (size_t j = 0; j < n; j++) ord[j] = (uint32_t)j;
            if (strict) {
                std::sort(ord.begin(), ord.end(), [&](uint32_t x, uint32_t y) {      // sortSeq_simple's order (:482-551)
                    if (mean_tmp[x] != mean_tmp[y]) return mean_tmp[x] > mean_tmp[y];
                    sort_key(fb, s0 + x, key_x); sort_key(fb, s0 + y, key_y);        // referenceUnitigToString of the two branches
                    return std::strcmp(key_x.c_str(), key_y.c_str()) > 0;
                });
            } else {
                std::sort(ord.begin(), ord.end(), [&](uint32_t x, uint32_t y) {      // sortSeq_branching: longer first, then larger
                    const size_t lx = seq_len(s0 + x), ly = seq_len(s0 + y);
                    if (lx != ly) return lx > ly;
                    return std::memcmp(seq_ptr(s0 + x), seq_ptr(s0 + y), lx) > 0;
                });
            }
            for (uint32_t j : ord) {
                abases_.append(seq_ptr(s0 + j), seq_len(s0 + j));
                aoff_.push_back(abases_.size());
                sorted_seq_.push_back((uint32_t)(s0 + j));
                sorted_mean_.push_back(mean_tmp[j]);
            }
            boff_.push_back((uint32_t)(aoff_.size() - 1));
            skip_.push_back(strict ? 1 : 0);
            k_src_.push_back((uint32_t)bi);
            k_first_.push_back((uint32_t)sorted_seq_.size());
            k_sum_.push_back(sum);
        }
        const size_t n_kept = k_src_.size();
        lap(stats_.gate_s);
        if (n_kept == 0) return true;
        stats_.bubbles_aligned += n_kept;
        // ---- SequenceAlignment of every kept bubble, then the site k-mers of the branching ones ----
        pf_msa_batch_t m;
        if (pf_align(ctx_, M_, D_, G_, abases_.data(), aoff_.data(), boff_.data(), (uint32_t)n_kept, &m) != PF_OK) return fail_batch(pf_last_error());
        lap(stats_.align_s);
        pf_site_batch_t sc;
        if (pf_site_cov(db_, lower_, upper_, skip_.data(), &sc) != PF_OK) return fail_batch(pf_last_error());
        lap(stats_.site_s);
        // ---- bubbles beyond a device limit: the host aligner, if the program gave one (HostMsa) ----
        host_.clear();
        host_of_.assign(n_kept, -1);
        site_kmers_.clear(); site_class_.clear();
        {
            static const unsigned force = std::getenv("PF_CALLER_FORCE_HOST") ? (unsigned)std::atoi(std::getenv("PF_CALLER_FORCE_HOST")) : 0u;
            for (size_t q = 0; q < n_kept; q++) {
                // beyond a device limit: no alignment result, or a branching bubble with more rows than the per-site kernels take (16)
                const bool over = m.status[q] != PF_BUBBLE_OK || (host_aligner_ && !skip_[q] && m.n_rows[q] > 16);
                if (!over && !(force && host_aligner_ && q % force == 0)) continue;
                if (!host_aligner_ || m.status[q] == PF_BUBBLE_BAD_INPUT)
                    return fail_batch("bubble " + std::to_string(k_src_[q]) + " of the batch does not fit the device limits (pf_msa_batch_t status " + std::to_string(m.status[q]) +
                                      ": more than 64 co-optimal alignments / candidate MSAs, 64 rows, or an alignment beyond the work area) and no host aligner is set");
                host_of_[q] = (int)host_.size();
                host_.emplace_back();
                if (!host_bubble(fb, q, sc, host_.back())) return fail_batch(err_);
            }
            stats_.bubbles_host_aligned += host_.size();
            if (!host_.empty() && !host_site_lookups()) return fail_batch(err_);
        }
        // one view per kept bubble: the device's arrays, or the host aligner's
        struct View {
            uint32_t nr, L;
            const char *rows;
            size_t n_var, n_ilen;
            const uint32_t *var_col, *ilen;
            const uint8_t *var_kind, *site_status;
            const uint16_t *cls;
            const uint64_t *site_cov;
        };
        auto view_of = [&](size_t q) {
            View v;
            if (host_of_[q] >= 0) {
                const HostBubble &h = host_[(size_t)host_of_[q]];
                v.nr = h.nr; v.L = h.L; v.rows = h.rows.data(); v.n_var = h.var_col.size(); v.n_ilen = h.ilen.size();
                v.var_col = h.var_col.data(); v.ilen = h.ilen.data(); v.var_kind = h.var_kind.data(); v.site_status = h.site_status.data();
                v.cls = h.cls.data(); v.site_cov = h.site_cov.data();
            } else {
                v.nr = m.n_rows[q]; v.L = m.aln_len[q]; v.rows = m.rows + m.rows_off[q];
                v.n_var = (size_t)(m.var_off[q + 1] - m.var_off[q]); v.n_ilen = (size_t)(m.ilen_off[q + 1] - m.ilen_off[q]);
                v.var_col = m.var_col + m.var_off[q]; v.ilen = m.ilen + m.ilen_off[q]; v.var_kind = m.var_kind + m.var_off[q];
                v.site_status = sc.status + sc.site_off[q]; v.cls = m.cls + m.cls_off[q]; v.site_cov = sc.cov + sc.cov_off[q];
            }
            return v;
        };
        // ---- ids: the bubbles whose alignment is not empty take consecutive ids in batch order (:2051, :2273) ----
        std::vector<size_t> ids(n_kept, 0);
        size_t next_id = var_id, n_called = 0;
        for (size_t q = 0; q < n_kept; q++) {
            if (view_of(q).nr == 0) continue;                                      // str_vec came back empty
            ids[q] = next_id++;
            n_called++;
        }
        // ---- rows: contiguous ranges of the kept bubbles, one host thread each, text joined in order ----
        auto emit = [&](size_t q0, size_t q1, CallerFiles &to, std::string &why) -> bool {
            std::string grouped_fre[4], cov_info, fre_info;
            std::vector<double> tc;
            for (size_t q = q0; q < q1; q++) {
                const size_t bi = k_src_[q];
                const bool strict = fb.strict[bi] != 0;
                const size_t ent_size = (size_t)fb.entrance_size[bi], ex_size = (size_t)fb.exit_size[bi];
                const double *means = sorted_mean_.data() + k_first_[q];
                const View vw = view_of(q);
                const uint32_t nr = vw.nr, L = vw.L;
                if (nr == 0) continue;
                const size_t var_count = ids[q];
                const char *rows = vw.rows;
                char head[96];
                const int head_len = std::snprintf(head, sizeof head, "%zu\t%d\t%u\t%u\t", var_count, strict ? 1 : 0, fb.entrance_id[bi], fb.exit_id[bi]);
                for (uint32_t r = 0; r < nr; r++) {
                    to.alignseq.append(head, (size_t)head_len);
                    to.alignseq.append(rows + (size_t)r * L, L);
                    to.alignseq += "\n";
                }
                const size_t n_var = vw.n_var;
                const uint16_t *cls = vw.cls;
                const uint32_t *ilen = vw.ilen;
                const size_t n_ilen = vw.n_ilen;
                size_t indel = 0;
                for (int a = 0; a < 4; a++) grouped_fre[a].clear();
                for (size_t i = 0; i < n_var; i++) {
                    const bool is_indel = vw.var_kind[i] == 1;
                    size_t var_distance;                                           // :2312-2330
                    auto gap_to = [&](size_t a, size_t c) { return (size_t)(vw.var_col[c] - vw.var_col[a] - 1); };
                    if (i == 0) var_distance = n_var > 1 ? std::min(gap_to(0, 1), ent_size) : std::min(ent_size, ex_size);
                    else if (i == n_var - 1) var_distance = std::min(gap_to(i - 1, i), ex_size);
                    else var_distance = std::min(gap_to(i - 1, i), gap_to(i, i + 1));
                    unsigned maxnum = 0;
                    for (uint32_t r = 0; r < nr; r++) maxnum = std::max<unsigned>(maxnum, cls[i * nr + r]);
                    tc.assign(maxnum, 0.0);
                    double sum = 0;
                    if (is_indel) indel++;                                         // :2390 / strict :2127, before anything can skip the site
                    if (strict) {
                        for (uint32_t r = 0; r < nr; r++) tc[cls[i * nr + r] - 1] += means[r];   // :2105-2108
                        sum = k_sum_[q];
                    } else {
                        const uint8_t st = vw.site_status[i];
                        if (st == PF_SITE_DROPPED) continue;                       // :2415-2418
                        if (st == PF_SITE_MISSING) { why = "a site k-mer is not in the database: the reference exits here (CDBG.cpp:54)"; return false; }
                        if (st != PF_SITE_OK) {
                            why = nr > 16 ? "a branching bubble with more than 16 aligned rows: beyond pf_site_cov's per-site row limit"
                                          : "a site k-mer cannot be formed (the reference reads outside the aligned row here)";
                            return false;
                        }
                        const uint64_t *cv = vw.site_cov + i * nr;
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
                    const int tail_len = std::snprintf(tail, sizeof tail, "%d\t%u\t%zu\t%zu\t%zu\t\n", strict ? 1 : 0, il, var_count, n_var, var_distance);
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
                    to.allele_frequency += grouped_fre[0]; to.allele_frequency += grouped_fre[1]; to.allele_frequency += grouped_fre[2];
                    if (strict) to.allele_frequency += grouped_fre[3];
                }
            }
            return true;
        };
        const size_t T = std::max<size_t>(1, std::min<size_t>(host_threads_, n_kept / 64));
        std::vector<CallerFiles> part(T);
        std::vector<std::string> why(T);
        std::vector<char> ok(T, 1);
        if (T == 1) ok[0] = emit(0, n_kept, part[0], why[0]);
        else {
            std::vector<std::thread> workers;
            for (size_t t = 0; t < T; t++)
                workers.emplace_back([&, t] { ok[t] = emit(n_kept * t / T, n_kept * (t + 1) / T, part[t], why[t]); });
            for (std::thread &w : workers) w.join();
        }
        for (size_t t = 0; t < T; t++)
            if (!ok[t]) return fail_batch(why[t]);
        // ---- commit: only a batch that went through completely changes the caller-visible state ----
        for (size_t q = 0; q < n_kept; q++)
            if (view_of(q).nr) out.called[called_base + k_src_[q]] = 1;
        out.bubbles_called += n_called;
        var_id = next_id;
        for (size_t t = 0; t < T; t++) {
            out.alignseq += part[t].alignseq;
            out.allele_frequency += part[t].allele_frequency;
            for (int a = 0; a < 4; a++) { out.cov[a] += part[t].cov[a]; out.fre[a] += part[t].fre[a]; out.alleles[a] += part[t].alleles[a]; }
        }
        lap(stats_.emit_s);
        return true;
    }

  private:
    // a bubble that went through the host aligner, in the layout the rows are written from (one bubble of pf_msa_batch_t / pf_site_batch_t)
    struct HostBubble {
        uint32_t nr = 0, L = 0;
        std::string rows;                       // nr x L
        std::vector<uint32_t> var_col, ilen;
        std::vector<uint8_t> var_kind, site_status;
        std::vector<uint16_t> cls;              // [var][row]
        std::vector<uint64_t> site_cov;         // [var][row], the first n_class entries are the class coverages
        std::vector<size_t> kmer_first;         // per variable column: first entry in site_kmers_ / site_class_ (branching bubbles)
        bool strict = false;
    };
    // SeqAlign::SequenceAlignment of kept bubble q on the host (the order of its branches is the alignment batch's)
    bool host_bubble(const FlatBatch &fb, size_t q, const pf_site_batch_t &, HostBubble &h) {
        std::vector<std::string> str;
        for (uint32_t s = k_first_[q]; s < k_first_[q + 1]; s++) {
            const size_t src = sorted_seq_[s];
            str.emplace_back(fb.bases.data() + fb.seq_off[src], (size_t)(fb.seq_off[src + 1] - fb.seq_off[src]));
        }
        HostMsa r;
        host_aligner_(str, r);
        h.strict = fb.strict[k_src_[q]] != 0;
        h.nr = (uint32_t)r.rows.size();
        if (h.nr == 0) return true;
        h.L = (uint32_t)r.rows[0].size();
        for (const std::string &row : r.rows) {
            if (row.size() != h.L) return fail("the host aligner returned rows of different lengths");
            h.rows += row;
        }
        for (size_t c = 0; c < r.partition.size(); c++) {                          // the variable columns (:2070-2076)
            if (r.partition[c].empty() || r.partition[c].back() == 0) continue;
            if (r.partition[c].size() != h.nr) return fail("the host aligner returned a partition column of the wrong height");
            h.var_col.push_back((uint32_t)c);
            h.var_kind.push_back(std::find(r.indel_pos.begin(), r.indel_pos.end(), (unsigned)c) != r.indel_pos.end() ? 1 : 0);
            for (unsigned short x : r.partition[c]) h.cls.push_back(x);
        }
        h.ilen.assign(r.indel_len.begin(), r.indel_len.end());
        h.site_status.assign(h.var_col.size(), PF_SITE_SKIPPED);
        h.site_cov.assign(h.var_col.size() * h.nr, 0);
        h.kmer_first.assign(h.var_col.size() + 1, site_kmers_.size());
        if (h.strict) return true;
        // lookup phase B of a branching bubble: distinct k-mers per class in std::set order (:2393-2396)
        if (k_ == 0) {
            pf_kmc_info_t info;
            if (pf_kmc_info(db_, &info) != PF_OK) return fail(pf_last_error());
            k_ = info.kmer_length;
        }
        std::vector<std::string> rows(r.rows), km;
        size_t n_indel = 0;
        for (size_t i = 0; i < h.var_col.size(); i++) {
            const bool is_indel = h.var_kind[i] == 1;
            const bool formed = host_site_kmers(rows, h.var_col[i], k_, is_indel, n_indel, km);
            if (is_indel) n_indel++;
            h.kmer_first[i] = site_kmers_.size();
            if (!formed) { h.site_status[i] = PF_SITE_UNDEFINED; continue; }
            h.site_status[i] = PF_SITE_OK;
            unsigned maxnum = 0;
            for (uint32_t rr = 0; rr < h.nr; rr++) maxnum = std::max<unsigned>(maxnum, h.cls[i * h.nr + rr]);
            std::vector<std::set<std::string>> sets(maxnum);
            for (uint32_t rr = 0; rr < h.nr; rr++) sets[h.cls[i * h.nr + rr] - 1].insert(km[rr]);
            for (unsigned c = 0; c < maxnum; c++)
                for (const std::string &x : sets[c]) { site_kmers_ += x; site_class_.push_back(c); }
        }
        h.kmer_first[h.var_col.size()] = site_kmers_.size();
        return true;
    }
    // one lookup call for the site k-mers of all host-aligned bubbles of the batch, then readCov(s, lower, upper)'s rule per site in the
    // reference's iteration order: a missing k-mer ends the program (:52-56), a counter outside (lower, upper) drops the site (:2415-2418)
    bool host_site_lookups() {
        const size_t n = site_class_.size();
        if (n) {
            std::vector<uint64_t> off(n + 1);
            for (size_t i = 0; i <= n; i++) off[i] = i * k_;
            std::vector<pf_cov_t> cv(n);
            if (pf_kmc_cov(db_, site_kmers_.data(), off.data(), (uint32_t)n, PF_LOOKUP_FWD_THEN_RC, 0, 0xFFFFFFFFu, cv.data()) != PF_OK) return fail(pf_last_error());
            for (HostBubble &h : host_) {
                if (h.strict) continue;
                for (size_t i = 0; i < h.var_col.size(); i++) {
                    if (h.site_status[i] != PF_SITE_OK) continue;
                    for (size_t e = h.kmer_first[i] / k_; e < h.kmer_first[i + 1] / k_; e++) {
                        if (cv[e].first_missing >= 0) { h.site_status[i] = PF_SITE_MISSING; break; }
                        const uint64_t cnt = cv[e].sum;
                        if (!(cnt > lower_ && cnt < upper_)) { h.site_status[i] = PF_SITE_DROPPED; break; }
                        h.site_cov[i * h.nr + site_class_[e]] += cnt;
                    }
                }
            }
        }
        site_kmers_.clear(); site_class_.clear();
        return true;
    }
    static void put_double(std::string &to, double v) {
        char buf[40];
        to.append(buf, (size_t)std::snprintf(buf, sizeof buf, "%g", v));
    }
    bool fail(const std::string &why) { err_ = why; return false; }
    // sort key of a strict bubble's branch: referenceUnitigToString() -- the string itself, or its reverse complement when the
    // branch was read from the unitig's reverse strand; an explicit key list (the std::string form) takes precedence
    void sort_key(const FlatBatch &fb, size_t s, std::string &key) const {
        if (explicit_keys_ && !(*explicit_keys_)[s].empty()) { key = (*explicit_keys_)[s]; return; }
        const char *p = fb.bases.data() + fb.seq_off[s];
        const size_t n = (size_t)(fb.seq_off[s + 1] - fb.seq_off[s]);
        if (fb.fwd[s]) { key.assign(p, n); return; }
        key.resize(n);
        for (size_t i = 0; i < n; i++) {
            const char c = p[n - 1 - i];
            key[i] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c == 'a' ? 't' : c == 'c' ? 'g' : c == 'g' ? 'c' : c == 't' ? 'a' : c;
        }
    }
    CallerStats stats_;
    HostAligner host_aligner_;
    std::vector<HostBubble> host_;
    std::vector<int> host_of_;
    std::string site_kmers_;
    std::vector<unsigned> site_class_;
    unsigned k_ = 0;
    const std::vector<std::string> *explicit_keys_ = nullptr;
    std::vector<pf_cov_t> cov_;
    std::vector<uint32_t> k_src_, k_first_, sorted_seq_;
    std::vector<double> k_sum_, sorted_mean_;
    std::string abases_;
    std::vector<uint64_t> aoff_;
    std::vector<uint32_t> boff_;
    std::vector<uint8_t> skip_;
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
