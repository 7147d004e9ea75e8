// pf_caller_colored.hpp -- the per-bubble caller of PloidyFrost's COLOURED mode over the C ABI, batched (header-only, C++11).
//
// What CCDBG::ploidyEstimation_ptr (`-t 1`, src/CCDBG.cpp:2759) and CCDBG::ploidyEstimation_multithread_ptr (`-t N`, :538) do for
// ONE superbubble of a coloured graph -- one KMC database and one (low, up) gate per colour (sample):
//   strict bubbles    (:674-920 = :2862-3058): readCovUni of every branch unitig in every colour that holds it (:697-720), the
//                      colour-set gate (:714), at least one colour with two covered branches (:722-737), sortSeq_simple by
//                      (number of colours, unitig length, unitig string) (:368-472), SeqAlign, Cramer's V over the per-colour branch
//                      coverages (:330, :794-798), one row per (site, colour) whose class coverages have >= 2 positive entries;
//   branching bubbles (:921-1376 = :3059-3500): path strings, sortSeq_branching, SeqAlign, site k-mers per variable column
//                      (:1057-1110, :1241-1262), per distinct k-mer of a class the colours of the unitig that holds it
//                      (findUnitig + UnitigColors::contains, :1127-1129), readCov in those colours (:1130), every colour must be
//                      seen (:1146), Cramer's V over the per-colour class coverages, one row per colour with >= 2 positive classes
// -- done here for a whole batch with device calls: pf_kmc_cov per colour (branch unitigs + entrances), pf_align, pf_site_kmers,
// pf_kmc_cov per colour (site k-mers).  The graph stays with the host program: it supplies, per strict branch, the colour mask of
// the unitig and whether its colour set is full (UnitigColors::size == colours x k-mers), and a callback that names the colours of
// the unitig holding a k-mer.  Text: the reference's files byte for byte in the `-t 1` dialect; the `-t N` dialect (0-based ids,
// P_allele_frequency grouped per bubble without the 5-allele sites, :873 / :1391) on request.
#ifndef PF_CALLER_COLORED_HPP
#define PF_CALLER_COLORED_HPP

#include <cmath>
#include <functional>
#include <utility>

#include "pf_caller.hpp"

namespace pfdropin {

struct ColoredBatch {
    FlatBatch flat;                    // branches as the strings that get aligned (mappedSequenceToString / path strings)
    std::vector<uint64_t> colors;      // per branch of a strict bubble: bit c = UnitigColors::contains(uu, c) (CCDBG.cpp:699); else 0
    std::vector<uint8_t> full;         // per branch of a strict bubble: UnitigColors::size(uu) == popcount(colors) * uu.len (:714)
    std::string ent_bases;             // per bubble: referenceUnitigToString() of the entrance (readCovUni(u, ...), :656-666)
    std::vector<uint64_t> ent_off{0};
    void clear() { flat.clear(); colors.clear(); full.clear(); ent_bases.clear(); ent_off.assign(1, 0); }
    void add_branch(const char *p, size_t n, bool forward, uint64_t color_mask, bool color_set_full) {
        flat.add_branch(p, n, forward); colors.push_back(color_mask); full.push_back(color_set_full ? 1 : 0);
    }
    void end_bubble(bool is_strict, unsigned ent_id, unsigned ex_id, size_t ent_size, size_t ex_size, const char *ent, size_t ent_len) {
        flat.end_bubble(is_strict, ent_id, ex_id, ent_size, ex_size);
        ent_bases.append(ent, ent_len); ent_off.push_back(ent_bases.size());
    }
    void append_bubble(const ColoredBatch &o, size_t b) {          // copies bubble b of another batch
        for (uint32_t s = o.flat.bubble_off[b]; s < o.flat.bubble_off[b + 1]; s++)
            add_branch(o.flat.bases.data() + o.flat.seq_off[s], (size_t)(o.flat.seq_off[s + 1] - o.flat.seq_off[s]), o.flat.fwd[s] != 0, o.colors[s], o.full[s] != 0);
        end_bubble(o.flat.strict[b] != 0, o.flat.entrance_id[b], o.flat.exit_id[b], (size_t)o.flat.entrance_size[b], (size_t)o.flat.exit_size[b],
                   o.ent_bases.data() + o.ent_off[b], (size_t)(o.ent_off[b + 1] - o.ent_off[b]));
    }
};

class ColoredBubbleCaller {
  public:
    // colours_of(kmer): bit c set iff the unitig that holds the k-mer carries colour c there (cdbg.findUnitig + contains); called
    // from several host threads at once when set_host_threads(n > 1)
    typedef std::function<uint64_t(const char *kmer)> ColorsOf;

    ColoredBubbleCaller(pf_ctx *ctx, const std::vector<pf_kmc *> &dbs, double match, double mismatch, double gap,
                        const std::vector<std::pair<int, int>> &cutoff, unsigned k, ColorsOf colours_of)
        : ctx_(ctx), dbs_(dbs), M_(match), D_(mismatch), G_(gap), cutoff_(cutoff), k_(k), colours_of_(std::move(colours_of)) {
        both_strands_.assign(dbs.size(), 1);
        for (size_t c = 0; c < dbs.size(); c++) {
            pf_kmc_info_t info;
            if (pf_kmc_info(dbs[c], &info) == PF_OK) both_strands_[c] = info.both_strands ? 1 : 0;
        }
    }
    const std::string &error() const { return err_; }
    const CallerStats &stats() const { return stats_; }
    void set_thread_dialect(bool multithread) { mt_ = multithread; }
    void set_host_threads(unsigned n) { host_threads_ = n ? n : 1; }
    // the aligner for bubbles beyond a device limit (HostMsa, pf_caller.hpp): inside the reference program the reference's own SeqAlign;
    // PF_CALLER_FORCE_HOST=<n> sends every n-th aligned bubble through it as well (tests)
    void set_host_aligner(HostAligner f) { host_aligner_ = std::move(f); }
    // sum of (size_t)core.first over the called bubbles and their number -- "Sites' Average Coverage" (:1441)
    size_t core_cov() const { return core_cov_; }
    size_t core_num() const { return core_num_; }

    bool call(const ColoredBatch &cb, size_t &var_id, CallerFiles &out) {
        err_.clear();
        const FlatBatch &fb = cb.flat;
        const size_t n_batch = fb.n_bubbles(), n_seq = fb.n_seq(), C = dbs_.size();
        const size_t called_base = out.called.size();
        out.called.resize(called_base + n_batch, 0);
        if (n_batch == 0) return true;
        if (C == 0 || C > 64) return fail("the coloured caller handles 1 .. 64 colours");
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

        // ---- lookup-A: per colour, readCovUni of the entrances and of the strict branches that carry the colour ----
        // ok = every k-mer present and low < count < up (CCDBG.cpp:139-152); a database counted on one strand only returns (0, true)
        // without lookups (:128, :157)
        ent_mean_.assign(C * n_batch, 0.0); ent_ok_.assign(C * n_batch, 1);
        br_mean_.assign(C * n_seq, 0.0); br_ok_.assign(C * n_seq, 1);
        for (size_t c = 0; c < C; c++) {
            if (!both_strands_[c]) continue;
            lbases_.clear(); loff_.assign(1, 0); lsrc_.clear();
            for (size_t b = 0; b < n_batch; b++) {
                lbases_.append(cb.ent_bases, cb.ent_off[b], cb.ent_off[b + 1] - cb.ent_off[b]);
                loff_.push_back(lbases_.size()); lsrc_.push_back((uint32_t)b);
            }
            for (size_t b = 0; b < n_batch; b++) {
                if (!fb.strict[b]) continue;
                for (size_t s = fb.bubble_off[b]; s < fb.bubble_off[b + 1]; s++)
                    if (cb.colors[s] >> c & 1) {
                        lbases_.append(seq_ptr(s), seq_len(s));
                        loff_.push_back(lbases_.size()); lsrc_.push_back((uint32_t)(n_batch + s));
                    }
            }
            lcov_.resize(lsrc_.size());
            if (!lsrc_.empty() && pf_kmc_cov(dbs_[c], lbases_.data(), loff_.data(), (uint32_t)lsrc_.size(), PF_LOOKUP_FWD_THEN_RC,
                                             (uint32_t)cutoff_[c].first, (uint32_t)cutoff_[c].second, lcov_.data()) != PF_OK)
                return fail_batch(pf_last_error());
            for (size_t i = 0; i < lsrc_.size(); i++) {
                const pf_cov_t &r = lcov_[i];
                const bool ok = r.first_missing < 0 && r.first_outside < 0;
                const double mean = ok ? (double)r.sum / (double)r.n_kmers : 0.0;
                if (lsrc_[i] < n_batch) { ent_ok_[c * n_batch + lsrc_[i]] = ok; ent_mean_[c * n_batch + lsrc_[i]] = mean; }
                else { br_ok_[c * n_seq + (lsrc_[i] - n_batch)] = ok; br_mean_[c * n_seq + (lsrc_[i] - n_batch)] = mean; }
            }
        }
        lap(stats_.lookup_s);

        // ---- gate + order the branches; build the alignment batch ----
        k_src_.clear(); k_first_.assign(1, 0); sorted_seq_.clear();
        abases_.clear(); aoff_.assign(1, 0); boff_.assign(1, 0); skip_.clear();
        abases_.reserve(fb.bases.size());
        std::vector<uint32_t> ord;
        std::vector<unsigned> n_col;
        std::string key_x, key_y;
        for (size_t bi = 0; bi < n_batch; bi++) {
            const size_t s0 = fb.bubble_off[bi], n = fb.bubble_off[bi + 1] - s0;
            const bool strict = fb.strict[bi] != 0;
            if (n < 2) continue;
            ord.resize(n);
            for (size_t j = 0; j < n; j++) ord[j] = (uint32_t)j;
            if (strict) {
                bool ok = true;
                n_col.assign(n, 0);
                for (size_t j = 0; j < n && ok; j++) {
                    for (size_t c = 0; c < C && ok; c++)
                        if (cb.colors[s0 + j] >> c & 1) { n_col[j]++; ok = br_ok_[c * n_seq + s0 + j] != 0; }     // :699-709
                    ok = ok && cb.full[s0 + j];                                                                   // :714
                }
                if (!ok) continue;
                bool two = false;                                                                                 // :722-737
                for (size_t c = 0; c < C && !two; c++) {
                    unsigned nz = 0;
                    for (size_t j = 0; j < n; j++) nz += ((cb.colors[s0 + j] >> c & 1) && br_mean_[c * n_seq + s0 + j] != 0.0) ? 1 : 0;
                    two = nz > 1;
                }
                if (!two) continue;
                std::sort(ord.begin(), ord.end(), [&](uint32_t x, uint32_t y) {      // sortSeq_simple (:368-472)
                    if (n_col[x] != n_col[y]) return n_col[x] > n_col[y];
                    const size_t lx = seq_len(s0 + x), ly = seq_len(s0 + y);
                    if (lx != ly) return lx > ly;
                    sort_key(fb, s0 + x, key_x); sort_key(fb, s0 + y, key_y);        // referenceUnitigToString of the two branches
                    return std::strcmp(key_x.c_str(), key_y.c_str()) > 0;
                });
            } else {
                std::sort(ord.begin(), ord.end(), [&](uint32_t x, uint32_t y) {      // sortSeq_branching (:473-537): longer first, then larger
                    const size_t lx = seq_len(s0 + x), ly = seq_len(s0 + y);
                    if (lx != ly) return lx > ly;
                    return std::memcmp(seq_ptr(s0 + x), seq_ptr(s0 + y), lx) > 0;
                });
            }
            for (uint32_t j : ord) {
                abases_.append(seq_ptr(s0 + j), seq_len(s0 + j));
                aoff_.push_back(abases_.size());
                sorted_seq_.push_back((uint32_t)(s0 + j));
            }
            boff_.push_back((uint32_t)(aoff_.size() - 1));
            skip_.push_back(strict ? 1 : 0);
            k_src_.push_back((uint32_t)bi);
            k_first_.push_back((uint32_t)sorted_seq_.size());
        }
        const size_t n_kept = k_src_.size();
        lap(stats_.gate_s);
        if (n_kept == 0) return true;
        stats_.bubbles_aligned += n_kept;

        // ---- SequenceAlignment, then the site k-mers of the branching bubbles ----
        pf_msa_batch_t m;
        if (pf_align(ctx_, M_, D_, G_, abases_.data(), aoff_.data(), boff_.data(), (uint32_t)n_kept, &m) != PF_OK) return fail_batch(pf_last_error());
        lap(stats_.align_s);
        pf_site_kmers_t sk;
        if (pf_site_kmers(ctx_, k_, skip_.data(), &sk) != PF_OK) return fail_batch(pf_last_error());
        // ---- bubbles beyond a device limit: the host aligner, if the program gave one ----
        host_.clear();
        host_of_.assign(n_kept, -1);
        {
            static const unsigned force = std::getenv("PF_CALLER_FORCE_HOST") ? (unsigned)std::atoi(std::getenv("PF_CALLER_FORCE_HOST")) : 0u;
            std::vector<std::string> str, km;
            for (size_t q = 0; q < n_kept; q++) {
                // beyond a device limit: no alignment result, or a branching bubble with more rows than the per-site kernels take (16)
                const bool over = m.status[q] != PF_BUBBLE_OK || (host_aligner_ && !skip_[q] && m.n_rows[q] > 16);
                if (!over && !(force && host_aligner_ && q % force == 0)) continue;
                if (!host_aligner_ || m.status[q] == PF_BUBBLE_BAD_INPUT)
                    return fail_batch("bubble " + std::to_string(k_src_[q]) + " of the batch does not fit the device limits (pf_msa_batch_t status " + std::to_string(m.status[q]) +
                                      ") and no host aligner is set");
                host_of_[q] = (int)host_.size();
                host_.emplace_back();
                HostBubble &h = host_.back();
                str.clear();
                for (uint32_t s = k_first_[q]; s < k_first_[q + 1]; s++) str.emplace_back(seq_ptr(sorted_seq_[s]), seq_len(sorted_seq_[s]));
                HostMsa r;
                host_aligner_(str, r);
                h.nr = (uint32_t)r.rows.size();
                if (h.nr == 0) continue;
                h.L = (uint32_t)r.rows[0].size();
                for (const std::string &row : r.rows) h.rows += row;
                for (size_t c = 0; c < r.partition.size(); c++) {
                    if (r.partition[c].empty() || r.partition[c].back() == 0) continue;
                    h.var_col.push_back((uint32_t)c);
                    h.var_kind.push_back(std::find(r.indel_pos.begin(), r.indel_pos.end(), (unsigned)c) != r.indel_pos.end() ? 1 : 0);
                    for (unsigned short x : r.partition[c]) h.cls.push_back(x);
                }
                h.ilen.assign(r.indel_len.begin(), r.indel_len.end());
                h.site_status.assign(h.var_col.size(), PF_SITE_SKIPPED);
                h.keys.assign(h.var_col.size() * h.nr, 0);
                if (skip_[q]) continue;
                size_t n_indel = 0;                                            // the site k-mers of a branching bubble, on the host
                for (size_t i = 0; i < h.var_col.size(); i++) {
                    const bool is_indel = h.var_kind[i] == 1;
                    const bool formed = host_site_kmers(r.rows, h.var_col[i], k_, is_indel, n_indel, km);
                    if (is_indel) n_indel++;
                    h.site_status[i] = formed ? PF_SITE_OK : PF_SITE_UNDEFINED;
                    if (!formed) continue;
                    for (uint32_t rr = 0; rr < h.nr; rr++) {
                        uint64_t key = 0;
                        for (char ch : km[rr]) key = key << 2 | (uint64_t)(ch == 'A' || ch == 'a' ? 0 : ch == 'C' || ch == 'c' ? 1 : ch == 'G' || ch == 'g' ? 2 : 3);
                        h.keys[i * h.nr + rr] = key;
                    }
                }
            }
            stats_.bubbles_host_aligned += host_.size();
        }
        // one view per kept bubble: the device's arrays, or the host aligner's
        struct View {
            uint32_t nr, L;
            const char *rows;
            size_t n_var, n_ilen;
            const uint32_t *var_col, *ilen;
            const uint8_t *var_kind, *site_status;
            const uint16_t *cls;
            const uint64_t *keys;
        };
        auto view_of = [&](size_t q) {
            View v;
            if (host_of_[q] >= 0) {
                const HostBubble &h = host_[(size_t)host_of_[q]];
                v.nr = h.nr; v.L = h.L; v.rows = h.rows.data(); v.n_var = h.var_col.size(); v.n_ilen = h.ilen.size();
                v.var_col = h.var_col.data(); v.ilen = h.ilen.data(); v.var_kind = h.var_kind.data(); v.site_status = h.site_status.data();
                v.cls = h.cls.data(); v.keys = h.keys.data();
            } else {
                v.nr = m.n_rows[q]; v.L = m.aln_len[q]; v.rows = m.rows + m.rows_off[q];
                v.n_var = (size_t)(m.var_off[q + 1] - m.var_off[q]); v.n_ilen = (size_t)(m.ilen_off[q + 1] - m.ilen_off[q]);
                v.var_col = m.var_col + m.var_off[q]; v.ilen = m.ilen + m.ilen_off[q]; v.var_kind = m.var_kind + m.var_off[q];
                v.site_status = sk.status + sk.site_off[q]; v.cls = m.cls + m.cls_off[q]; v.keys = sk.keys + sk.key_off[q];
            }
            return v;
        };

        // ---- lookup-B: distinct k-mers per (site, class) in std::set order, their colours from the graph, readCov per colour ----
        // site_first_[col_first_[q] + i] ..: entries of skey_ / sclass_ / smask_ of variable column i of kept bubble q
        col_first_.assign(n_kept + 1, 0);
        for (size_t q = 0; q < n_kept; q++) col_first_[q + 1] = col_first_[q] + view_of(q).n_var;
        const uint64_t n_cols = col_first_[n_kept];
        site_first_.assign(n_cols + 1, 0); skey_.clear(); sclass_.clear();
        {
            std::vector<std::pair<uint32_t, uint64_t>> tmp;
            for (size_t q = 0; q < n_kept; q++) {
                const View vw = view_of(q);
                const uint32_t nr = vw.nr;
                for (size_t i = 0; i < vw.n_var; i++) {
                    if (!skip_[q] && nr && vw.site_status[i] == PF_SITE_OK) {
                        const uint16_t *cls = vw.cls + i * nr;
                        const uint64_t *keys = vw.keys + i * nr;
                        tmp.clear();
                        for (uint32_t r = 0; r < nr; r++) tmp.push_back(std::make_pair((uint32_t)(cls[r] - 1), keys[r]));
                        std::sort(tmp.begin(), tmp.end());                           // class ascending, then k-mer ascending == set<string> order
                        tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
                        for (const auto &e : tmp) { sclass_.push_back(e.first); skey_.push_back(e.second); }
                    }
                    site_first_[col_first_[q] + i + 1] = skey_.size();
                }
            }
        }
        const size_t n_keys = skey_.size();
        smask_.assign(n_keys, 0);
        {
            auto colour = [&](size_t i0, size_t i1) {
                std::string s(k_, 'A');
                for (size_t i = i0; i < i1; i++) { unpack(skey_[i], s); smask_[i] = colours_of_(s.c_str()); }
            };
            const size_t T = std::max<size_t>(1, std::min<size_t>(host_threads_, n_keys / 256));
            if (T == 1) colour(0, n_keys);
            else {
                std::vector<std::thread> w;
                for (size_t t = 0; t < T; t++) w.emplace_back(colour, n_keys * t / T, n_keys * (t + 1) / T);
                for (std::thread &x : w) x.join();
            }
        }
        scount_.assign(C * n_keys, 0.0); sok_.assign(C * n_keys, 1);
        {
            std::string s(k_, 'A');
            for (size_t c = 0; c < C; c++) {
                if (!both_strands_[c]) continue;                                     // readCov returns (0, true) without lookups (:94, :120)
                lbases_.clear(); loff_.assign(1, 0); lsrc_.clear();
                for (size_t i = 0; i < n_keys; i++)
                    if (smask_[i] >> c & 1) { unpack(skey_[i], s); lbases_ += s; loff_.push_back(lbases_.size()); lsrc_.push_back((uint32_t)i); }
                lcov_.resize(lsrc_.size());
                if (!lsrc_.empty() && pf_kmc_cov(dbs_[c], lbases_.data(), loff_.data(), (uint32_t)lsrc_.size(), PF_LOOKUP_FWD_THEN_RC,
                                                 (uint32_t)cutoff_[c].first, (uint32_t)cutoff_[c].second, lcov_.data()) != PF_OK)
                    return fail_batch(pf_last_error());
                for (size_t i = 0; i < lsrc_.size(); i++) {
                    const pf_cov_t &r = lcov_[i];
                    const bool ok = r.first_missing < 0 && r.first_outside < 0;
                    sok_[c * n_keys + lsrc_[i]] = ok;
                    scount_[c * n_keys + lsrc_[i]] = ok ? (double)r.sum / (double)r.n_kmers : 0.0;
                }
            }
        }
        lap(stats_.site_s);

        // ---- ids (:751 fetch_add / :2953 ++var_count) ----
        std::vector<size_t> ids(n_kept, 0);
        size_t next_id = var_id, n_called = 0;
        for (size_t q = 0; q < n_kept; q++) {
            if (view_of(q).nr == 0) continue;
            ids[q] = next_id++;
            n_called++;
        }
        // ---- rows ----
        const uint64_t all_colours = C == 64 ? ~0ull : ((1ull << C) - 1);
        // contiguous ranges of the kept bubbles, one host thread each, text joined in order (it does not depend on the thread count)
        auto emit = [&](size_t q0, size_t q1, CallerFiles &part, std::string &why, size_t &core_cov, size_t &core_num) -> bool {
        std::string grouped_fre[4], cov_info, fre_info, tail;
        std::vector<std::vector<double>> cov_vec(C);
        std::vector<double> tc, res;
        for (size_t q = q0; q < q1; q++) {
            const size_t bi = k_src_[q];
            const bool strict = fb.strict[bi] != 0;
            const size_t ent_size = (size_t)fb.entrance_size[bi], ex_size = (size_t)fb.exit_size[bi];
            const View vw = view_of(q);
            const uint32_t nr = vw.nr, L = vw.L;
            if (nr == 0) continue;
            {   // core: the entrance's mean coverage summed over the colours up to the first one that fails (:656-666)
                double core = 0;
                for (size_t c = 0; c < C; c++) { if (!ent_ok_[c * n_batch + bi]) break; core += ent_mean_[c * n_batch + bi]; }
                core_cov += (size_t)core; core_num++;
            }
            const size_t var_count = ids[q];
            const char *rows = vw.rows;
            char head[96];
            const int head_len = std::snprintf(head, sizeof head, "%zu\t%d\t%u\t%u\t", var_count, strict ? 1 : 0, fb.entrance_id[bi], fb.exit_id[bi]);
            for (uint32_t r = 0; r < nr; r++) {
                part.alignseq.append(head, (size_t)head_len);
                part.alignseq.append(rows + (size_t)r * L, L);
                part.alignseq += "\n";
            }
            const uint64_t v0 = col_first_[q];
            const size_t n_var = vw.n_var;
            const uint16_t *cls = vw.cls;
            const uint32_t *ilen = vw.ilen;
            const size_t n_ilen = vw.n_ilen;
            size_t indel = 0;
            for (int a = 0; a < 4; a++) grouped_fre[a].clear();
            double strict_coef = 0;
            if (strict) {                                                            // per colour, per branch in aligned order (:794-798)
                for (size_t c = 0; c < C; c++) {
                    cov_vec[c].assign(nr, 0.0);
                    for (uint32_t r = 0; r < nr; r++) {
                        const size_t s = sorted_seq_[k_first_[q] + r];
                        if (cb.colors[s] >> c & 1) cov_vec[c][r] = br_mean_[c * n_seq + s];
                    }
                }
                strict_coef = max_cramer(cov_vec);
            }
            for (size_t i = 0; i < n_var; i++) {
                const bool is_indel = vw.var_kind[i] == 1;
                size_t var_distance;                                               // :805-824
                auto gap_to = [&](size_t a, size_t c) { return (size_t)(vw.var_col[c] - vw.var_col[a] - 1); };
                if (i == 0) var_distance = n_var > 1 ? std::min(gap_to(0, 1), ent_size) : std::min(ent_size, ex_size);
                else if (i == n_var - 1) var_distance = std::min(gap_to(i - 1, i), ex_size);
                else var_distance = std::min(gap_to(i - 1, i), gap_to(i, i + 1));
                unsigned maxnum = 0;
                for (uint32_t r = 0; r < nr; r++) maxnum = std::max<unsigned>(maxnum, cls[i * nr + r]);
                if (is_indel) indel++;
                const uint32_t il = is_indel ? (indel - 1 < n_ilen ? ilen[indel - 1] : 0u) : 0u;
                double coef = strict_coef;
                if (!strict) {
                    const uint8_t st = vw.site_status[i];
                    if (st != PF_SITE_OK) {
                        why = nr > 16 ? "a branching bubble with more than 16 aligned rows: beyond pf_site_kmers' per-site row limit"
                                      : "a site k-mer cannot be formed (the reference reads outside the aligned row here)";
                        return false;
                    }
                    for (size_t c = 0; c < C; c++) cov_vec[c].assign(maxnum, 0.0);
                    uint64_t seen = 0;
                    bool ok = true;
                    for (uint64_t e = site_first_[v0 + i]; e < site_first_[v0 + i + 1] && ok; e++) {   // :1119-1142
                        seen |= smask_[e];
                        for (size_t c = 0; c < C && ok; c++)
                            if (smask_[e] >> c & 1) {
                                ok = sok_[c * n_keys + e] != 0;
                                cov_vec[c][sclass_[e]] += scount_[c * n_keys + e];
                            }
                    }
                    if (!ok || seen != all_colours) continue;                      // :1146-1153
                    coef = max_cramer(cov_vec);
                }
                tail.clear();
                {
                    char buf[160];
                    int n = std::snprintf(buf, sizeof buf, "%d\t%u\t%zu\t%zu\t", strict ? 1 : 0, il, var_count, n_var);
                    tail.append(buf, (size_t)n);
                    put_double(tail, coef);
                    n = std::snprintf(buf, sizeof buf, "\t%zu\t\n", var_distance);
                    tail.append(buf, (size_t)n);
                }
                for (size_t c = 0; c < C; c++) {
                    res.clear();
                    double sum = 0;
                    if (strict) {                                                  // :838-851
                        tc.assign(maxnum, 0.0);
                        for (uint32_t r = 0; r < nr; r++) tc[cls[i * nr + r] - 1] += cov_vec[c][r];
                        for (double x : tc) if (x > 0.0) { res.push_back(x); sum += x; }
                    } else {
                        for (double x : cov_vec[c]) if (x > 0.0) { sum += x; res.push_back(x); }
                    }
                    if (res.size() < 2) continue;
                    cov_info.clear(); fre_info.clear();
                    for (double x : res) {
                        put_double(cov_info, x); cov_info += '\t';
                        put_double(fre_info, x / sum); fre_info += '\n';
                    }
                    cov_info += std::to_string(c); cov_info += '\t'; cov_info += tail;
                    if (!mt_) part.allele_frequency += fre_info;                   // -t 1: every row as it is made (:3032, :3312)
                    if (res.size() >= 2 && res.size() <= 5) {
                        const size_t a = res.size() - 2;
                        part.alleles[a]++;
                        part.cov[a] += cov_info;
                        part.fre[a] += fre_info;
                        if (mt_) grouped_fre[a] += fre_info;
                    }
                }
            }
            if (mt_) { part.allele_frequency += grouped_fre[0]; part.allele_frequency += grouped_fre[1]; part.allele_frequency += grouped_fre[2]; }   // :873, :1391
        }
        return true;
        };
        const size_t T = std::max<size_t>(1, std::min<size_t>(host_threads_, n_kept / 64));
        std::vector<CallerFiles> parts(T);
        std::vector<std::string> whys(T);
        std::vector<size_t> cc(T, 0), cn(T, 0);
        std::vector<char> oks(T, 1);
        if (T == 1) oks[0] = emit(0, n_kept, parts[0], whys[0], cc[0], cn[0]);
        else {
            std::vector<std::thread> workers;
            for (size_t t = 0; t < T; t++)
                workers.emplace_back([&, t] { oks[t] = emit(n_kept * t / T, n_kept * (t + 1) / T, parts[t], whys[t], cc[t], cn[t]); });
            for (std::thread &w : workers) w.join();
        }
        for (size_t t = 0; t < T; t++)
            if (!oks[t]) return fail_batch(whys[t]);
        // ---- commit ----
        for (size_t q = 0; q < n_kept; q++)
            if (view_of(q).nr) out.called[called_base + k_src_[q]] = 1;
        out.bubbles_called += n_called;
        var_id = next_id;
        for (size_t t = 0; t < T; t++) {
            core_cov_ += cc[t]; core_num_ += cn[t];
            out.alignseq += parts[t].alignseq;
            out.allele_frequency += parts[t].allele_frequency;
            for (int a = 0; a < 4; a++) { out.cov[a] += parts[t].cov[a]; out.fre[a] += parts[t].fre[a]; out.alleles[a] += parts[t].alleles[a]; }
        }
        lap(stats_.emit_s);
        return true;
    }

    // computeCramerVCoefficient (CCDBG.cpp:330-366), the same operations in the same order
    static double cramer_v(const std::vector<double> &A, const std::vector<double> &Bv) {
        double n = 0, nA = 0, nB = 0, chi = 0;
        uint8_t count = 0;
        double p_small[64];
        std::vector<double> p_big;
        double *p = p_small;
        if (A.size() > 64) { p_big.assign(A.size(), 0); p = p_big.data(); }
        for (size_t i = 0; i < A.size(); i++) {
            nA += A[i]; nB += Bv[i];
            p[i] = A[i] + Bv[i];
            n = n + p[i];
            if (p[i] != 0) ++count;
        }
        if (count < 2) return 0;
        for (size_t i = 0; i < A.size(); i++) {
            if (p[i] == 0) continue;
            const double exA = nA * p[i] / n, exB = nB * p[i] / n;
            chi += std::pow(A[i] - exA, 2) / exA;
            chi += std::pow(Bv[i] - exB, 2) / exB;
        }
        return std::sqrt(chi / n);
    }

  private:
    static double max_cramer(const std::vector<std::vector<double>> &cov_vec) {      // :794-798
        double coefficient = 0;
        for (size_t ci = 0; ci + 1 < cov_vec.size(); ci++)
            for (size_t cj = ci + 1; cj < cov_vec.size(); cj++) coefficient = std::max(coefficient, cramer_v(cov_vec[ci], cov_vec[cj]));
        return coefficient;
    }
    static void put_double(std::string &to, double v) {
        char buf[40];
        to.append(buf, (size_t)std::snprintf(buf, sizeof buf, "%g", v));
    }
    void unpack(uint64_t key, std::string &s) const {
        for (unsigned i = 0; i < k_; i++) s[i] = "ACGT"[(key >> (2 * (k_ - 1 - i))) & 3];
    }
    bool fail(const std::string &why) { err_ = why; return false; }
    static void sort_key(const FlatBatch &fb, size_t s, std::string &key) {
        const char *p = fb.bases.data() + fb.seq_off[s];
        const size_t n = (size_t)(fb.seq_off[s + 1] - fb.seq_off[s]);
        if (fb.fwd[s]) { key.assign(p, n); return; }
        key.resize(n);
        for (size_t i = 0; i < n; i++) {
            const char c = p[n - 1 - i];
            key[i] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c;
        }
    }
    struct HostBubble {      // a bubble that went through the host aligner, in the layout the rows are written from
        uint32_t nr = 0, L = 0;
        std::string rows;
        std::vector<uint32_t> var_col, ilen;
        std::vector<uint8_t> var_kind, site_status;
        std::vector<uint16_t> cls;
        std::vector<uint64_t> keys;
    };
    HostAligner host_aligner_;
    std::vector<HostBubble> host_;
    std::vector<int> host_of_;
    std::vector<uint64_t> col_first_;
    pf_ctx *ctx_;
    std::vector<pf_kmc *> dbs_;
    double M_, D_, G_;
    std::vector<std::pair<int, int>> cutoff_;
    unsigned k_;
    ColorsOf colours_of_;
    std::vector<uint8_t> both_strands_;
    bool mt_ = false;
    unsigned host_threads_ = 1;
    std::string err_;
    CallerStats stats_;
    size_t core_cov_ = 0, core_num_ = 0;
    std::vector<double> ent_mean_, br_mean_, scount_;
    std::vector<uint8_t> ent_ok_, br_ok_, sok_, skip_;
    std::string lbases_, abases_;
    std::vector<uint64_t> loff_, aoff_, site_first_, skey_, smask_;
    std::vector<uint32_t> lsrc_, k_src_, k_first_, sorted_seq_, boff_, sclass_;
    std::vector<pf_cov_t> lcov_;
};

}  // namespace pfdropin
#endif  // PF_CALLER_COLORED_HPP
