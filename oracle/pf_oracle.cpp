// pf_oracle.cpp -- TEST INFRASTRUCTURE.  CPU restatement of the two halves of PloidyFrost's
// per-superbubble hot path, written from the reference's semantics (not its text):
//
//   (1) the KMC database reader + k-mer lookup     KMC/kmc_api/kmc_file.cpp, kmer_api.h, mmer.h
//   (2) SeqAlign (fill, traceback, MSA, site call)  src/SeqAlign.cpp, src/SeqAlign.hpp
//
// Every function cites the reference lines it restates.  PARITY PINNING: the reference ships no tests
// or golden vectors for this path, so this oracle is pinned against the reference
// itself compiled here (oracle/_ref/libpfref.so, see oracle/Makefile) -- tests/test_oracle_vs_ref.py
// diffs every output field on randomised inputs, and tests/golden/*.json holds vectors generated from
// that build (tests/golden/make_golden.py).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load the
// resulting library (oracle/libpforacle.so).  The product never links or calls it.
#include <algorithm>
#include <atomic>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "../include/pf_types.h"
#include "msa_pack.hpp"

namespace {

typedef unsigned long long u64;
typedef unsigned int u32;

// =====================================================================================================
// Part 1: KMC database
// =====================================================================================================

// mmer.h:34-57 -- which m-mers may serve as a signature.
bool mmer_allowed(u32 m, u32 len) {
    if ((m & 0x3f) == 0x3f) return false;  // ...TTT
    if ((m & 0x3f) == 0x3b) return false;  // ...TGT
    if ((m & 0x3c) == 0x3c) return false;  // ...TG*
    for (u32 j = 0; j + 3 < len; j++) {
        if ((m & 0xf) == 0) return false;  // AA inside
        m >>= 2;
    }
    if (m == 0) return false;        // AAA...
    if (m == 0x04) return false;     // ACA...
    if ((m & 0xf) == 0) return false;  // *AA...
    return true;
}

// mmer.h:61-87 -- norm[x] = min(allowed(x) ? x : 4^m, allowed(rc(x)) ? rc(x) : 4^m)
std::vector<u32> build_norm(u32 len) {
    u32 special = 1u << (2 * len);
    std::vector<u32> norm(special);
    for (u32 x = 0; x < special; x++) {
        u32 rc = 0, t = x;
        for (u32 i = 0; i < len; i++) {
            rc = (rc << 2) | (3 - (t & 3));
            t >>= 2;
        }
        u32 a = mmer_allowed(x, len) ? x : special;
        u32 b = mmer_allowed(rc, len) ? rc : special;
        norm[x] = a < b ? a : b;
    }
    return norm;
}

struct OrcDb {
    u32 k = 0, mode = 0, C = 0, p = 0, sig_len = 0, min_count = 0, version = 0;
    u64 max_count = 0, N = 0;
    bool both_strands = true;
    u32 S = 0, R = 0;            // suffix bytes, record bytes (kmc_file.cpp:240-242, 295-297)
    u64 single_lut = 0;          // 4^p (kmc_file.cpp:223)
    std::vector<u64> lut;        // prefix_file_buf up to and including the N+1 sentinel
    std::vector<u32> sigmap;     // KMC2 signature -> bin (kmc_file.cpp:235)
    std::vector<unsigned char> suf;  // records only (markers stripped)
    std::vector<u32> norm;
};

bool read_file(const std::string &path, std::vector<unsigned char> &buf) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize((size_t)sz);
    size_t got = sz ? fread(buf.data(), 1, (size_t)sz, f) : 0;
    fclose(f);
    return got == (size_t)sz;
}

u32 rd32(const unsigned char *p) { u32 v; memcpy(&v, p, 4); return v; }
u64 rd64(const unsigned char *p) { u64 v; memcpy(&v, p, 8); return v; }

// kmc_file.cpp:140-302.  Both files carry a 4-byte marker at either end (:140-181).
bool orc_open(OrcDb &db, const std::string &prefix) {
    std::vector<unsigned char> pre;
    if (!read_file(prefix + ".kmc_pre", pre) || pre.size() < 24) return false;
    if (memcmp(pre.data(), "KMCP", 4) || memcmp(pre.data() + pre.size() - 4, "KMCP", 4)) return false;
    const size_t fs = pre.size();
    db.version = rd32(&pre[fs - 12]);                 // :188-192
    const u64 header_offset = pre[fs - 8];            // only one byte is read (:200, :257)
    if (db.version == 0x200) {                        // :196-245
        const unsigned char *h = &pre[fs - 8 - header_offset];
        db.k = rd32(h); db.mode = rd32(h + 4); db.C = rd32(h + 8); db.p = rd32(h + 12);
        db.sig_len = rd32(h + 16); db.min_count = rd32(h + 20); db.max_count = rd32(h + 24);
        db.N = rd64(h + 28);
        db.both_strands = !h[36];                     // stored inverted (:218-219)
        const u64 sigmap_n = (1ull << (2 * db.sig_len)) + 1;
        const u64 body = fs - 8 - 4;                  // without markers and header_offset word
        const u64 lut_bytes = body - (sigmap_n * 4 + header_offset + 8);
        const u64 lut_n = lut_bytes / 8;              // index of the guard word (:224, :233)
        db.lut.resize(lut_n + 1);
        memcpy(db.lut.data(), &pre[4], (lut_n + 1) * 8);
        db.lut[lut_n] = db.N + 1;
        db.sigmap.resize(sigmap_n);
        memcpy(db.sigmap.data(), &pre[4 + (lut_n + 1) * 8], sigmap_n * 4);
        db.norm = build_norm(db.sig_len);
    } else if (db.version == 0) {                     // :246-300
        const u64 body = fs - 8 - 4;
        const u64 words = body / 8;
        std::vector<u64> w(words);
        memcpy(w.data(), &pre[4], words * 8);
        u64 hi = (body - header_offset) / 8;
        db.k = (u32)w[hi]; db.mode = (u32)(w[hi] >> 32);
        db.C = (u32)w[hi + 1]; db.p = (u32)(w[hi + 1] >> 32);
        db.min_count = (u32)w[hi + 2];
        db.max_count = (w[hi + 2] >> 32) + (w[hi + 4] & 0xFFFFFFFF00000000ull);
        db.N = w[hi + 3];
        db.both_strands = !((w[hi + 4] & 0xF) == 1);
        db.lut.assign(w.begin(), w.begin() + hi + 1);
        db.lut[hi] = db.N + 1;                        // sentinel overwrites first header word (:292)
        db.sig_len = 0;
    } else {
        return false;
    }
    if (db.k == 0 || db.k > 32 || db.p > db.k || (db.k - db.p) % 4 != 0) return false;
    db.S = (db.k - db.p) / 4;
    db.R = db.S + db.C;
    db.single_lut = 1ull << (2 * db.p);
    std::vector<unsigned char> suf;
    if (!read_file(prefix + ".kmc_suf", suf) || suf.size() < 8) return false;
    if (memcmp(suf.data(), "KMCS", 4) || memcmp(suf.data() + suf.size() - 4, "KMCS", 4)) return false;
    db.suf.assign(suf.begin() + 4, suf.end() - 4);
    return true;
}

// kmer_api.h:653-672 on a right-aligned 2-bit k-mer value.
u32 orc_signature(const OrcDb &db, u64 kmer) {
    const u32 m = db.sig_len;
    const u64 mask = (1ull << (2 * m)) - 1;
    u32 best = 0xFFFFFFFFu;
    for (u32 i = 0; i + m <= db.k; i++) {
        u32 v = db.norm[(kmer >> (2 * (db.k - m - i))) & mask];
        if (v < best) best = v;
    }
    return best;
}

// kmc_file.cpp:330-366 (CheckKmer) + :1383-1462 (BinarySearch), on a right-aligned k-mer value.
bool orc_check_kmer(const OrcDb &db, u64 kmer, u32 &count) {
    const u32 sbits = 2 * (db.k - db.p);
    const u64 prefix = sbits >= 64 ? 0 : (kmer >> sbits);
    const u64 suffix = sbits >= 64 ? kmer : (kmer & ((1ull << sbits) - 1));
    if (prefix >= db.lut.size()) return false;                        // :344
    u64 slot = prefix;
    if (db.version == 0x200) slot += (u64)db.sigmap[orc_signature(db, kmer)] * db.single_lut;  // :347-355
    if (slot + 1 >= db.lut.size()) return false;
    long long lo = (long long)db.lut[slot], hi = (long long)db.lut[slot + 1] - 1;
    if (lo >= (long long)db.N) return false;                          // :1385
    if (hi > (long long)db.N - 1) hi = (long long)db.N - 1;           // last bucket's stop is one past the end
    while (lo <= hi) {                                                // :1398-1438
        long long mid = (lo + hi) / 2;
        const unsigned char *r = &db.suf[(size_t)mid * db.R];
        u64 rs = 0;
        for (u32 a = 0; a < db.S; a++) rs = (rs << 8) | r[a];         // MSB-first byte compare == integer compare
        if (rs == suffix) {
            u64 c = 0;
            for (u32 b = 0; b < db.C; b++) c |= (u64)r[db.S + b] << (8 * b);  // :1444-1452
            count = (u32)c;
            return c >= db.min_count && c <= db.max_count;           // :1459 (mode 0)
        }
        if (rs < suffix) lo = mid + 1; else hi = mid - 1;
    }
    return false;
}

int base_code(unsigned char c) {  // kmer_api.h:264-275
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}

u64 revcomp_kmer(u64 v, u32 k) {
    u64 r = 0;
    for (u32 i = 0; i < k; i++) {
        r = (r << 2) | (3 - (v & 3));
        v >>= 2;
    }
    return r;
}

// One window under the three dialects of pf_types.h.  Returns found, count (0 when not found).
bool orc_window(const OrcDb &db, const char *w, int mode, u32 &count) {
    count = 0;
    u64 v = 0;
    for (u32 i = 0; i < db.k; i++) {
        int c = base_code((unsigned char)w[i]);
        if (c < 0) return false;  // from_string refuses (kmer_api.h:502-509); read API zeroes the window (kmc_file.cpp:1036-1047)
        v = (v << 2) | (u64)c;
    }
    u32 c = 0;
    bool ok;
    if (mode == PF_LOOKUP_FWD) {
        ok = orc_check_kmer(db, v, c);
    } else if (mode == PF_LOOKUP_FWD_THEN_RC) {      // CDBG.cpp:38-43
        ok = orc_check_kmer(db, v, c);
        if (!ok) ok = orc_check_kmer(db, revcomp_kmer(v, db.k), c);
    } else {                                         // kmc_file.cpp:1060, :1290
        u64 rc = revcomp_kmer(v, db.k);
        ok = orc_check_kmer(db, v < rc ? v : rc, c);
    }
    if (ok) count = c;
    return ok;
}

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

// =====================================================================================================
// Part 2: SeqAlign
// =====================================================================================================

struct Scoring {
    double match, mismatch, gap;
};

// What survives of an AlignUnit (SeqAlign.hpp:30-68): the aligned pair, where gaps were opened in A,
// and the three keys operator- compares.
struct PairAln {
    std::string a, b;
    std::vector<u32> gap_rows;  // gap_pos: row index of every Left move, in traceback order
    long score = 0;
    size_t n_pos = 0;           // pos.size()
    u32 n_indel = 0;            // indel
};

// AlignUnit::operator- (SeqAlign.hpp:43-67): >0 when `l` ranks above `r`.
long rank_diff(long l_score, size_t l_pos, u32 l_indel, long r_score, size_t r_pos, u32 r_indel) {
    if (l_score != r_score) return l_score > r_score ? 1 : -1;
    if (l_pos != r_pos) return (long)r_pos - (long)l_pos;
    if (l_indel != r_indel) return (long)r_indel - (long)l_indel;
    return 0;
}

// variantAnalyze (SeqAlign.cpp:237-305).  Gap test comes first here (:241), unlike the fill.
void analyze_pair(const Scoring &sc, const std::string &A, const std::string &B, long &score, size_t &n_pos, u32 &n_indel) {
    score = 0;
    n_pos = 0;
    n_indel = 0;
    int run = 0;  // 1: inside a gap run of A, 2: of B
    for (size_t i = 0; i < A.size(); i++) {
        double s = (A[i] == '-' || B[i] == '-') ? sc.gap : (A[i] == B[i] ? sc.match : sc.mismatch);
        score = (long)((double)score + s);  // `long += double` (:255)
        if (A[i] != B[i]) {
            if (A[i] == '-') {
                if (run != 1) { run = 1; n_indel++; n_pos++; }
            } else if (B[i] == '-') {
                if (run != 2) { run = 2; n_indel++; n_pos++; }
            } else {
                run = 0;
                n_pos++;
            }
        } else {
            run = 0;
        }
    }
}

enum { F_UP = 1, F_DIAG = 2, F_LEFT = 4 };

// needlemanWunch fill (SeqAlign.cpp:480-547) -> flags[(m+1)*(n+1)].
void nw_fill(const Scoring &sc, const std::string &A, const std::string &B, std::vector<unsigned char> &flags) {
    const size_t m = A.size(), n = B.size(), W = n + 1;
    std::vector<long> score((m + 1) * W, 0);
    flags.assign((m + 1) * W, 0);
    for (size_t i = 1; i <= m; i++) { score[i * W] = (long)(sc.gap * (double)i); flags[i * W] = F_UP; }
    for (size_t j = 1; j <= n; j++) { score[j] = (long)(sc.gap * (double)j); flags[j] = F_LEFT; }
    for (size_t i = 1; i <= m; i++)
        for (size_t j = 1; j <= n; j++) {
            char a = A[i - 1], b = B[j - 1];
            double sub = (a == b) ? sc.match : ((a == '-' || b == '-') ? sc.gap : sc.mismatch);  // :498-506
            int up = (int)((double)score[(i - 1) * W + j] + sc.gap) + ((flags[(i - 1) * W + j] & F_UP) ? 1 : 0);
            int dg = (int)((double)score[(i - 1) * W + j - 1] + sub) + ((flags[(i - 1) * W + j - 1] & F_DIAG) ? 1 : 0);
            int lf = (int)((double)score[i * W + j - 1] + sc.gap) + ((flags[i * W + j - 1] & F_LEFT) ? 1 : 0);
            int best = std::max(std::max(up, dg), lf);
            if (best == lf && i != m && A[i] == '-') {  // profile rule (:528-532): do not open a gap in front of an old one
                lf = INT_MIN;
                best = up > dg ? up : dg;
            }
            score[i * W + j] = best;
            unsigned char f = 0;
            if (up == best) f |= F_UP;
            if (dg == best) f |= F_DIAG;
            if (lf == best) f |= F_LEFT;
            flags[i * W + j] = f;
        }
}

// traceback (SeqAlign.cpp:306-478).  `base` is the by-value matrix that gets
// permanently pruned, `work` the copy whose flags are consumed and restored.
std::vector<PairAln> nw_traceback(const Scoring &sc, const std::string &A, const std::string &B,
                                  std::vector<unsigned char> base) {
    const size_t W = B.size() + 1;
    std::vector<unsigned char> work = base;
    std::vector<PairAln> out;
    std::vector<std::pair<size_t, size_t>> path;
    path.push_back(std::make_pair(A.size(), B.size()));
    std::string ra, rb;            // built by prepending; '+' marks a gap opened by this traceback
    std::vector<u32> gaps;
    size_t open_a = 0, open_b = 0, cap_a = 5, cap_b = 5;  // size_t on purpose: open_b may wrap (:454-467)
    auto front = [](const std::string &s) -> char { return s.empty() ? '\0' : s[0]; };
    while (!path.empty()) {
        const size_t i = path.back().first, j = path.back().second;
        const size_t cell = i * W + j;
        if (i == 0 && j == 0 && open_a <= cap_a && open_b <= cap_b) {  // :322-355
            PairAln cand;
            cand.a = ra;
            for (char &c : cand.a) if (c == '+') c = '-';
            cand.b = rb;
            cand.gap_rows = gaps;
            analyze_pair(sc, cand.a, cand.b, cand.score, cand.n_pos, cand.n_indel);
            bool keep = true;
            if (!out.empty()) {
                const PairAln &last = out.back();
                int d = (int)rank_diff(last.score, last.n_pos, last.n_indel, cand.score, cand.n_pos, cand.n_indel);
                if (d > 0) keep = false;
                else if (d < 0) out.clear();
            }
            if (keep) { out.push_back(cand); cap_a = open_a; cap_b = open_b; }
        }
        if (work[cell] & F_LEFT) {                                       // :356-392
            bool take;
            if (open_a < cap_a) {
                if (ra.empty() || ra[0] != '+') ++open_a;
                take = true;
            } else if (open_a == cap_a) {
                take = front(ra) == '+';
            } else {
                take = false;
            }
            if (!take) { base[cell] &= ~F_LEFT; work[cell] &= ~F_LEFT; continue; }
            path.push_back(std::make_pair(i, j - 1));
            ra.insert(ra.begin(), '+');
            gaps.push_back((u32)i);
            rb.insert(rb.begin(), B[j - 1]);
            work[cell] &= ~F_LEFT;
        } else if (work[cell] & F_UP) {                                  // :393-424
            bool take;
            if (open_b < cap_b) {
                if (rb.empty() || rb[0] == '-') ++open_b;                // sic: counts extensions (:397)
                take = true;
            } else if (open_b == cap_b) {
                take = front(rb) == '-';
            } else {
                take = false;
            }
            if (!take) { base[cell] &= ~F_UP; work[cell] &= ~F_UP; continue; }
            path.push_back(std::make_pair(i - 1, j));
            ra.insert(ra.begin(), A[i - 1]);
            rb.insert(rb.begin(), '-');
            work[cell] &= ~F_UP;
        } else if (work[cell] & F_DIAG) {                                // :425-431
            path.push_back(std::make_pair(i - 1, j - 1));
            ra.insert(ra.begin(), A[i - 1]);
            rb.insert(rb.begin(), B[j - 1]);
            work[cell] &= ~F_DIAG;
        } else {                                                         // :432-474
            if (ra.empty()) break;
            path.pop_back();
            work[cell] = base[cell];
            if (ra[0] == '+' && (ra.size() < 2 || ra[1] != '+')) --open_a;
            if (rb[0] == '-' && (rb.size() < 2 || rb[1] != '-')) --open_b;
            if (ra[0] == '+') gaps.pop_back();
            ra.erase(ra.begin());
            rb.erase(rb.begin());
        }
    }
    return out;
}

std::vector<PairAln> nw_pair(const Scoring &sc, const std::string &A, const std::string &B) {
    std::vector<unsigned char> flags;
    nw_fill(sc, A, B, flags);
    return nw_traceback(sc, A, B, flags);
}

// compute_dis lambda (SeqAlign.cpp:10-38); L = length of the last candidate's last row.
size_t site_spacing(const std::vector<u32> &v, size_t L) {
    size_t count = 0;
    if (v.empty()) return 0;
    if (v.size() == 1) {
        int left = (int)v[0];
        int right = (int)(L - v[0]) - 1;
        count = left > right ? (size_t)(left + 1) : (size_t)right;
    } else {
        count = v[0];
        for (size_t i = 1; i < v.size(); i++) count = (size_t)std::min((int)(v[i] - v[i - 1] - 1), (int)count);
        count = std::min(count, L - v.back() - 1);
    }
    return count;
}

// compareStrPair (SeqAlign.cpp:8-236): call sites column by column for every candidate MSA, keep the best.
void choose_msa(const std::vector<std::vector<std::string>> &cands, pforacle::MsaResult &res) {
    res = pforacle::MsaResult();
    if (cands.empty()) return;
    const size_t Llast = cands.back().back().size();
    int best_snp_dis = INT_MAX, best_indel_dis = INT_MAX, best_snp = INT_MAX / 2, best_indel = INT_MAX / 2;
    int best_all_dis = INT_MAX, best_l = -1, best_r = -1;
    for (size_t ci = 0; ci < cands.size(); ci++) {
        const std::vector<std::string> &rows = cands[ci];
        const size_t nr = rows.size(), L = rows.back().size();
        std::vector<u32> snp_pos, indel_pos, indel_len;
        std::vector<std::vector<unsigned short>> part;
        bool open = false;
        unsigned char n_indel = 0, n_snp = 0;  // uint8_t in the reference (:54-55)
        for (size_t j = 0; j < L; j++) {
            std::set<char> chars;
            for (size_t r = 0; r < nr; r++) chars.insert(rows[r][j]);
            std::vector<unsigned short> cls(nr, 0);
            bool number = false;
            if (chars.size() > 1) {
                if (!chars.count('-')) {                                  // SNP column (:66-93)
                    if (open) { indel_len.push_back((u32)(j - indel_pos[n_indel - 1])); open = false; }
                    snp_pos.push_back((u32)j);
                    n_snp++;
                    number = true;
                } else {                                                  // column with a gap (:94-146)
                    bool continues = true;
                    if (open) {
                        for (size_t r = 0; r < nr; r++)
                            if ((rows[r][j] == '-') != (rows[r][j - 1] == '-')) { continues = false; break; }
                        if (!continues) {
                            indel_len.push_back((u32)(j - indel_pos[n_indel - 1]));
                            ++n_indel;
                            indel_pos.push_back((u32)j);
                        }
                    } else {
                        continues = false;
                        ++n_indel;
                        indel_pos.push_back((u32)j);
                        open = true;
                    }
                    number = !continues || chars.size() > 2;
                }
            } else if (open) {                                            // :148-155
                indel_len.push_back((u32)(j - indel_pos[n_indel - 1]));
                open = false;
            }
            if (number) {  // class ids in order of first appearance (:75-92, :123-144)
                unsigned short next = 0;
                for (size_t r = 0; r < nr; r++) {
                    bool seen = false;
                    for (size_t q = 0; q < r; q++)
                        if (rows[q][j] == rows[r][j]) { cls[r] = cls[q]; seen = true; break; }
                    if (!seen) cls[r] = ++next;
                }
            }
            part.push_back(cls);
        }
        // 7-level preference (:158-233)
        bool take = false, take_by_rows = false;
        size_t d_indel = 0, d_snp = 0, d_all = 0;
        std::vector<u32> merged(snp_pos.size() + indel_pos.size());
        std::merge(snp_pos.begin(), snp_pos.end(), indel_pos.begin(), indel_pos.end(), merged.begin());
        const int total = (int)n_snp + (int)n_indel, best_total = best_snp + best_indel;
        if (total < best_total) take = true;
        else if (total == best_total) {
            if ((int)n_indel < best_indel) take = true;
            else if ((int)n_indel == best_indel) {
                d_indel = site_spacing(indel_pos, Llast);
                if (d_indel > (size_t)best_indel_dis) take = true;
                else if (d_indel == (size_t)best_indel_dis) {
                    d_snp = site_spacing(snp_pos, Llast);
                    if (d_snp > (size_t)best_snp_dis) take = true;
                    else if (d_snp == (size_t)best_snp_dis) {
                        d_all = site_spacing(merged, Llast);
                        if (d_all > (size_t)best_all_dis) take = true;
                        else if (d_all == (size_t)best_all_dis) {
                            int l = merged.empty() ? 0 : (int)merged[0];
                            int r = merged.empty() ? 0 : (int)merged.back();
                            if (l > best_l || r > best_r) take = true;
                            else if (l == best_l && r == best_r) {
                                for (size_t q = 0; q < nr; q++)
                                    if (strcmp(rows[q].c_str(), res.rows[q].c_str()) > 0) { take_by_rows = true; break; }
                            }
                        }
                    }
                }
            }
        }
        if (take || take_by_rows) {
            // both replacement paths end in the same state: site_l/site_r use max(old,new) on the `flag`
            // path (:222-223) and equal the incumbent's on the strcmp path (:196-198).
            best_all_dis = (int)site_spacing(merged, Llast);
            best_l = std::max(best_l, merged.empty() ? -1 : (int)merged[0]);
            best_r = std::max(best_r, merged.empty() ? -1 : (int)merged.back());
            best_snp = n_snp;
            best_indel = n_indel;
            best_snp_dis = (int)site_spacing(snp_pos, Llast);
            best_indel_dis = (int)site_spacing(indel_pos, Llast);
            res.rows = rows;
            res.snp_pos = snp_pos;
            res.indel_pos = indel_pos;
            res.indel_len = indel_len;
            res.partition = part;
        }
    }
}

// SequenceAlignment (SeqAlign.cpp:550-640): progressive MSA over the candidates of the first pair.
void msa_align(const Scoring &sc, const std::vector<std::string> &seqs, pforacle::MsaResult &res) {
    std::vector<std::vector<std::string>> cands;
    for (const PairAln &pa : nw_pair(sc, seqs[0], seqs[1])) cands.push_back({pa.a, pa.b});
    for (size_t i = 2; i < seqs.size(); i++) {
        std::vector<std::vector<std::string>> prev;
        prev.swap(cands);
        int best_total = INT_MIN;
        for (size_t k = 0; k < prev.size(); k++) {
            std::vector<PairAln> ext = nw_pair(sc, prev[k][0], seqs[i]);   // row 0 (with its gaps) vs new sequence
            std::vector<std::vector<std::string>> built(ext.size());
            std::vector<int> alive;
            for (size_t v = 0; v < ext.size(); v++) { alive.push_back((int)v); built[v].push_back(ext[v].a); }
            int total_k = 0;
            for (size_t j = 1; j < i; j++) {
                int best_j = INT_MIN;
                long inc_score = INT_MIN; size_t inc_pos = 0; u32 inc_indel = 0;  // au_max (:577-579)
                std::vector<int> alive_j;
                for (int v : alive) {
                    // project the gaps this extension opened in row 0 into row j (:583-597)
                    std::string row;
                    const std::vector<u32> &g = ext[v].gap_rows;
                    if (!g.empty()) {
                        u32 from = 0;
                        for (size_t s = g.size(); s-- > 0;) {
                            row += prev[k][j].substr(from, g[s] - from) + "-";
                            from = g[s];
                        }
                        row += prev[k][j].substr(from);
                    } else {
                        row = prev[k][j];
                    }
                    long s; size_t np; u32 ni;
                    analyze_pair(sc, row, ext[v].b, s, np, ni);
                    int d = (int)rank_diff(s, np, ni, inc_score, inc_pos, inc_indel);
                    if (d > 0) {
                        inc_score = s; inc_pos = np; inc_indel = ni;
                        best_j = (int)inc_score;
                        alive_j.clear();
                        alive_j.push_back(v);
                        built[v].push_back(row);
                    } else if (d == 0) {
                        best_j = (int)inc_score;
                        alive_j.push_back(v);
                        built[v].push_back(row);
                    }
                }
                alive = alive_j;
                total_k = (int)((unsigned)total_k + (unsigned)best_j);  // `int +=` may overflow in the reference; wraps on x86-64
            }
            if (total_k > best_total) { best_total = total_k; cands.clear(); }
            if (total_k >= best_total)
                for (int v : alive) { built[v].push_back(ext[v].b); cands.push_back(built[v]); }
        }
    }
    choose_msa(cands, res);
}

}  // namespace

// =====================================================================================================
// C entry points (same shapes as oracle/ref_shim.cpp; loaded by oracle/bindings.py)
// =====================================================================================================
extern "C" {

void *pforc_kmc_open(const char *prefix) {
    OrcDb *db = new OrcDb();
    if (!orc_open(*db, prefix) || db->mode != 0) { delete db; return nullptr; }
    return db;
}
void pforc_kmc_close(void *h) { delete (OrcDb *)h; }

int pforc_kmc_info(void *h, pf_kmc_info_t *o) {
    OrcDb *db = (OrcDb *)h;
    memset(o, 0, sizeof(*o));
    o->kmer_length = db->k; o->mode = db->mode; o->counter_size = db->C; o->lut_prefix_length = db->p;
    o->signature_len = db->sig_len; o->min_count = db->min_count; o->max_count = db->max_count;
    o->total_kmers = db->N; o->both_strands = db->both_strands; o->kmc_version = db->version;
    o->n_bins = db->version == 0x200 ? (uint32_t)((db->lut.size() - 1) / db->single_lut) : 1;
    return 0;
}
int pforc_kmc_set_min_count(void *h, uint32_t x) { ((OrcDb *)h)->min_count = x; return 0; }   // kmc_file.cpp SetMinCount
int pforc_kmc_set_max_count(void *h, uint32_t x) { ((OrcDb *)h)->max_count = x; return 0; }

int pforc_kmc_counts(void *h, const char *bases, const uint64_t *seq_off, uint32_t n_seq, int mode,
                     int /*use_read_api*/, int n_threads, uint32_t *counts, uint8_t *found) {
    OrcDb *db = (OrcDb *)h;
    const u32 k = db->k;
    std::vector<uint64_t> koff(n_seq + 1, 0);
    for (uint32_t s = 0; s < n_seq; s++) {
        uint64_t len = seq_off[s + 1] - seq_off[s];
        koff[s + 1] = koff[s] + (len >= k ? len - k + 1 : 0);
    }
    parallel_for(n_seq, n_threads, [&](size_t s) {
        uint64_t len = seq_off[s + 1] - seq_off[s];
        for (uint64_t i = 0; i + k <= len; i++) {
            u32 c;
            bool ok = orc_window(*db, bases + seq_off[s] + i, mode, c);
            counts[koff[s] + i] = c;
            if (found) found[koff[s] + i] = ok;
        }
    });
    return 0;
}

// readCov reductions (CDBG.cpp:29-120)
int pforc_kmc_cov(void *h, const char *bases, const uint64_t *seq_off, uint32_t n_seq, int mode,
                  uint32_t low, uint32_t up, int n_threads, pf_cov_t *out) {
    OrcDb *db = (OrcDb *)h;
    const u32 k = db->k;
    parallel_for(n_seq, n_threads, [&](size_t s) {
        uint64_t len = seq_off[s + 1] - seq_off[s];
        pf_cov_t r;
        r.sum = 0; r.min = 10000; r.n_kmers = len >= k ? (uint32_t)(len - k + 1) : 0;
        r.first_missing = -1; r.first_outside = -1;
        for (uint32_t i = 0; i < r.n_kmers; i++) {
            u32 c;
            if (!orc_window(*db, bases + seq_off[s] + i, mode, c)) {
                if (r.first_missing < 0) r.first_missing = (int32_t)i;
                continue;
            }
            r.sum += c;
            if (c < r.min) r.min = c;
            if (!(c > low && c < up) && r.first_outside < 0) r.first_outside = (int32_t)i;
        }
        out[s] = r;
    });
    return 0;
}

void *pforc_align(double M, double D, double G, const char *bases, const uint64_t *seq_off,
                  const uint32_t *bubble_off, uint32_t n_bubbles, int n_threads, pf_msa_batch_t *out) {
    std::vector<pforacle::MsaResult> res(n_bubbles);
    Scoring sc = {M, D, G};
    parallel_for(n_bubbles, n_threads, [&](size_t b) {
        std::vector<std::string> seqs;
        for (uint32_t s = bubble_off[b]; s < bubble_off[b + 1]; s++)
            seqs.emplace_back(bases + seq_off[s], seq_off[s + 1] - seq_off[s]);
        msa_align(sc, seqs, res[b]);
    });
    pforacle::MsaPacked *p = new pforacle::MsaPacked();
    p->pack(res);
    p->view(out);
    return p;
}
void pforc_msa_free(void *h) { delete (pforacle::MsaPacked *)h; }

int pforc_nw_pair(double M, double D, double G, const char *a, const char *b, char *buf, size_t cap) {
    Scoring sc = {M, D, G};
    std::vector<PairAln> v = nw_pair(sc, a, b);
    size_t o = 0;
    for (auto &pa : v) {
        if (o + pa.a.size() + pa.b.size() + 2 > cap) break;
        memcpy(buf + o, pa.a.c_str(), pa.a.size() + 1); o += pa.a.size() + 1;
        memcpy(buf + o, pa.b.c_str(), pa.b.size() + 1); o += pa.b.size() + 1;
    }
    return (int)v.size();
}

}  // extern "C"
