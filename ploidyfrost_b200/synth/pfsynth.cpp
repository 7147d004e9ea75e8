// pfsynth.cpp -- synthetic workload generator for tests and bench (data tooling, not the hot path).
//
// i.i.d. uniform ACGT ancestor; each haplotype derived with per-base SNP
// probability p_snp (uniform over the 3 other bases) and indel probability p_indel (50/50 insertion /
// deletion, length geometric with mean 3, capped at 10); optional long indels (config 5).  PRNG =
// splitmix64 keyed by (seed, position) so every base is reproducible independently of threading.
//
// Bubbles are synthesised directly from the known variant map ("flat bubble batches may
// be synthesised directly from the known variant map when a graph build is impractical"): variants whose
// conserved separation is shorter than k merge into one cluster; a cluster is a superbubble whose paths
// are the distinct haplotype sequences from the last k-mer of the entrance unitig to the first k-mer of
// the exit unitig (CDBG.cpp:2226).  A cluster whose paths share no inner k-mer is a strict bubble (each
// branch is one unitig, CDBG.cpp:1998-2050), otherwise a branching one (:2190-2273).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

namespace {

inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
inline double u01(uint64_t r) { return (double)(r >> 11) * (1.0 / 9007199254740992.0); }

struct Event {            // one variant of one haplotype, in ancestor coordinates
    uint64_t L, R;        // last conserved position before / first conserved position after
    uint32_t hap;
};

struct Hap {
    std::vector<uint8_t> seq;             // ASCII
    std::vector<uint64_t> ev_anc;         // indel events: first ancestor position whose shift changed
    std::vector<int64_t> ev_shift;        // shift (hap_pos - anc_pos) valid from ev_anc on
    std::vector<Event> events;
    int64_t shift_at(uint64_t a) const {  // for conserved ancestor position a
        size_t i = std::upper_bound(ev_anc.begin(), ev_anc.end(), a) - ev_anc.begin();
        return i ? ev_shift[i - 1] : 0;
    }
};

const char ALPHA[4] = {'A', 'C', 'G', 'T'};

// `offset`: absolute coordinate of anc[0] -- events are keyed by absolute position, so a region generated on its own carries the
// same variants as the same stretch of the whole genome (away from the region's `margin`s, where no event is placed).
void make_haplotype(uint64_t seed, uint32_t hap_id, const std::vector<uint8_t> &anc, double p_snp, double p_indel,
                    double p_long, uint32_t long_min, uint32_t long_max, Hap &h, uint64_t offset = 0) {
    const uint64_t n = anc.size();
    h.seq.clear();
    h.seq.reserve(n + n / 64);
    int64_t shift = 0;
    uint64_t a = 0;
    const uint64_t key = splitmix64(seed ^ (0xA5A5A5A5ull * (hap_id + 1)));
    const uint64_t margin = 64;
    while (a < n) {
        const uint64_t r0 = splitmix64(key ^ ((a + offset) * 0x9E3779B97F4A7C15ull));
        const double u = u01(r0);
        const bool inner = a > margin && a + margin + long_max + 16 < n;
        if (inner && u < p_snp) {
            const uint64_t r1 = splitmix64(r0);
            const uint8_t c = anc[a];
            uint8_t idx = 0;
            for (uint8_t q = 0; q < 4; q++) if (ALPHA[q] == c) idx = q;
            h.seq.push_back((uint8_t)ALPHA[(idx + 1 + r1 % 3) & 3]);
            h.events.push_back({a - 1, a + 1, hap_id});
            a++;
        } else if (inner && u < p_snp + p_indel + p_long) {
            const uint64_t r1 = splitmix64(r0), r2 = splitmix64(r1);
            uint32_t len;
            if (u < p_snp + p_indel) {
                len = 1;
                uint64_t rr = r2;
                while (len < 10 && u01(rr = splitmix64(rr)) < 2.0 / 3.0) len++;   // geometric, mean 3, cap 10
            } else {
                len = long_min + (uint32_t)(r2 % (uint64_t)(long_max - long_min + 1));
            }
            if (r1 & 1) {  // insertion after ancestor base a
                h.seq.push_back(anc[a]);
                uint64_t rr = r2 ^ 0x1234567ull;
                for (uint32_t q = 0; q < len; q++) { rr = splitmix64(rr); h.seq.push_back((uint8_t)ALPHA[rr & 3]); }
                shift += len;
                h.ev_anc.push_back(a + 1);
                h.ev_shift.push_back(shift);
                h.events.push_back({a, a + 1, hap_id});
                a++;
            } else {       // deletion of [a, a+len)
                shift -= len;
                h.ev_anc.push_back(a + len);
                h.ev_shift.push_back(shift);
                h.events.push_back({a - 1, a + len, hap_id});
                a += len;
            }
        } else {
            h.seq.push_back(anc[a]);
            a++;
        }
    }
}

// CDBG::sortSeq_branching (CDBG.cpp:417-480): the reference's own quicksort, step for step (its result can
// differ from a clean sort when the pivot gets swapped away, so the bench inputs use it verbatim).
void sort_branching(std::vector<std::string> &v, int low, int high) {
    if (high <= low) return;
    int i = low, j = high;
    auto gt = [](const std::string &a, const std::string &b) { return strcmp(a.c_str(), b.c_str()) > 0; };
    auto lt = [](const std::string &a, const std::string &b) { return strcmp(a.c_str(), b.c_str()) < 0; };
    while (true) {
        while (v[i].size() >= v[low].size()) {
            if (v[i].size() > v[low].size()) i++;
            else if (gt(v[i], v[low])) i++;
            else break;
            if (i == high) break;
        }
        while (v[j].size() <= v[low].size()) {
            if (v[j].size() < v[low].size()) j--;
            else if (lt(v[j], v[low])) j--;
            else break;
            if (j == low) break;
        }
        if (i >= j) break;
        std::swap(v[i], v[j]);
    }
    std::swap(v[j], v[low]);
    sort_branching(v, low, j - 1);
    sort_branching(v, j + 1, high);
}

struct Workload {
    std::vector<uint8_t> anc;
    std::vector<Hap> haps;
    // bubble batch
    std::vector<char> bases;
    std::vector<uint64_t> seq_off{0};
    std::vector<uint32_t> bubble_off{0};   // CSR over *branch* sequences
    std::vector<uint8_t> bubble_type;      // 1 strict, 0 branching
    std::vector<uint32_t> ent_size, exit_size;
    std::vector<char> ent_bases;           // entrance unitigs (for the core-coverage lookups, CDBG.cpp:1997)
    std::vector<uint64_t> ent_off{0};
};

}  // namespace

extern "C" {

// root_seed != 0: the ancestor is a SUBGENOME ancestor -- the i.i.d. root genome of `root_seed` with every base substituted
// with probability `divergence` (BASELINE configs[2]: 3 subgenomes, 10 % diverged); root_seed == 0: the ancestor is i.i.d.
// offset: the generated stretch is [offset, offset + genome_len) of the (conceptual) whole genome.
void *pfs_create_sub(uint64_t seed, uint64_t genome_len, uint32_t n_hap, double p_snp, double p_indel, double p_long,
                     uint32_t long_min, uint32_t long_max, int n_threads, uint64_t root_seed, double divergence, uint64_t offset) {
    Workload *w = new Workload();
    w->anc.resize(genome_len);
    const uint64_t gkey = splitmix64(root_seed ? root_seed : seed), skey = splitmix64(seed ^ 0x5B5B5B5B5B5B5B5Bull);
    auto fill = [&](uint64_t b, uint64_t e) {
        for (uint64_t ii = b; ii < e; ii++) {
            const uint64_t i = ii + offset;
            uint32_t c = (uint32_t)(splitmix64(gkey ^ (i * 0xD6E8FEB86659FD93ull)) & 3);
            if (root_seed) {
                const uint64_t r = splitmix64(skey ^ (i * 0x9E3779B97F4A7C15ull));
                if (u01(r) < divergence) c = (c + 1 + (uint32_t)(splitmix64(r) % 3)) & 3;
            }
            w->anc[ii] = (uint8_t)ALPHA[c];
        }
    };
    n_threads = std::max(1, n_threads);
    {
        std::vector<std::thread> th;
        const uint64_t chunk = (genome_len + n_threads - 1) / n_threads;
        for (int t = 0; t < n_threads; t++) {
            uint64_t b = t * chunk, e = std::min(genome_len, b + chunk);
            if (b < e) th.emplace_back(fill, b, e);
        }
        for (auto &t : th) t.join();
    }
    w->haps.resize(n_hap);
    {
        std::vector<std::thread> th;
        for (uint32_t h = 0; h < n_hap; h++)
            th.emplace_back([=] { make_haplotype(seed, h, w->anc, p_snp, p_indel, p_long, long_min, long_max, w->haps[h], offset); });
        for (auto &t : th) t.join();
    }
    return w;
}

void pfs_destroy(void *h) { delete (Workload *)h; }

void *pfs_create(uint64_t seed, uint64_t genome_len, uint32_t n_hap, double p_snp, double p_indel, double p_long,
                 uint32_t long_min, uint32_t long_max, int n_threads) {
    return pfs_create_sub(seed, genome_len, n_hap, p_snp, p_indel, p_long, long_min, long_max, n_threads, 0, 0.0, 0);
}

uint64_t pfs_hap_len(void *h, uint32_t hap) { return ((Workload *)h)->haps[hap].seq.size(); }
const uint8_t *pfs_hap_ptr(void *h, uint32_t hap) { return ((Workload *)h)->haps[hap].seq.data(); }
const uint8_t *pfs_anc_ptr(void *h) { return ((Workload *)h)->anc.data(); }

// Synthesise the bubbles of ancestor region [r0, r1).  Returns the number of bubbles.
uint64_t pfs_make_bubbles(void *hh, uint32_t k, uint64_t r0, uint64_t r1, uint64_t max_bubbles) {
    Workload *w = (Workload *)hh;
    const uint32_t nh = (uint32_t)w->haps.size();
    std::vector<Event> ev;
    for (auto &h : w->haps)
        for (auto &e : h.events)
            if (e.L >= r0 + 4 * k && e.R + 4 * k < r1) ev.push_back(e);
    std::sort(ev.begin(), ev.end(), [](const Event &a, const Event &b) { return a.L < b.L || (a.L == b.L && a.R < b.R); });
    w->bases.clear(); w->seq_off.assign(1, 0); w->bubble_off.assign(1, 0); w->bubble_type.clear();
    w->ent_size.clear(); w->exit_size.clear(); w->ent_bases.clear(); w->ent_off.assign(1, 0);
    size_t i = 0;
    uint64_t prevR = r0 + 2 * k;   // start of the conserved stretch before the current cluster
    uint64_t nb = 0;
    while (i < ev.size() && nb < max_bubbles) {
        uint64_t Lc = ev[i].L, Rc = ev[i].R;
        size_t j = i + 1;
        while (j < ev.size() && (int64_t)ev[j].L - (int64_t)Rc + 1 < (int64_t)k) { Rc = std::max(Rc, ev[j].R); j++; }
        const uint64_t nextL = j < ev.size() ? ev[j].L : Rc + 2 * k;
        if ((int64_t)Lc - (int64_t)prevR + 1 >= (int64_t)k) {
            // paths: haplotype sequence over ancestor [Lc-k+1, Rc+k-1]; both ends lie in conserved stretches
            std::vector<std::string> paths;
            for (uint32_t h = 0; h < nh; h++) {
                const Hap &H = w->haps[h];
                const int64_t s = (int64_t)(Lc - k + 1) + H.shift_at(Lc - k + 1);
                const int64_t e = (int64_t)(Rc + k - 1) + H.shift_at(Rc + k - 1);
                std::string p((const char *)H.seq.data() + s, (size_t)(e - s + 1));
                if (std::find(paths.begin(), paths.end(), p) == paths.end()) paths.push_back(p);
            }
            if (paths.size() >= 2) {
                // strict iff the inner k-mer sets of the paths are pairwise disjoint
                bool strict = true;
                std::unordered_set<std::string> seen;
                for (auto &p : paths) {
                    std::unordered_set<std::string> mine;
                    for (size_t q = 1; q + k + 1 <= p.size(); q++) mine.insert(p.substr(q, k));
                    for (auto &km : mine)
                        if (!seen.insert(km).second) { strict = false; break; }
                    if (!strict) break;
                }
                std::vector<std::string> seqs;
                if (strict) for (auto &p : paths) seqs.push_back(p.substr(1, p.size() - 2));  // branch unitigs
                else { seqs = paths; sort_branching(seqs, 0, (int)seqs.size() - 1); }
                for (auto &s : seqs) {
                    w->bases.insert(w->bases.end(), s.begin(), s.end());
                    w->seq_off.push_back(w->bases.size());
                }
                w->bubble_off.push_back((uint32_t)(w->seq_off.size() - 1));
                w->bubble_type.push_back(strict ? 1 : 0);
                const uint64_t elen = std::min<uint64_t>(Lc - prevR + 1, 2000);
                w->ent_bases.insert(w->ent_bases.end(), w->anc.begin() + (Lc + 1 - elen), w->anc.begin() + Lc + 1);
                w->ent_off.push_back(w->ent_bases.size());
                w->ent_size.push_back((uint32_t)elen);
                w->exit_size.push_back((uint32_t)std::min<uint64_t>(nextL - Rc + 1, 2000));
                nb++;
            }
        }
        prevR = Rc;
        i = j;
    }
    return nb;
}

// BASELINE configs[4] (indel-heavy stress): n bubbles of 2..max_rows branches; the first branch has a length log-uniform in
// [min_len, max_len]; every other branch is a copy with 1 % substitutions, 0.1 % short indels and at least one long indel
// (an insertion or deletion of 5 .. 40 % of the branch, capped so the result stays within [k, max_len]).  Branches are
// ordered like the reference's branching bubbles (sort_branching).  The ancestor / haplotypes of `hh` are not used.
uint64_t pfs_make_long_bubbles(void *hh, uint64_t seed, uint32_t k, uint64_t n_bubbles, uint32_t min_len, uint32_t max_len,
                               uint32_t max_rows) {
    Workload *w = (Workload *)hh;
    w->bases.clear(); w->seq_off.assign(1, 0); w->bubble_off.assign(1, 0); w->bubble_type.clear();
    w->ent_size.clear(); w->exit_size.clear(); w->ent_bases.clear(); w->ent_off.assign(1, 0);
    const double lmin = std::log((double)min_len), lmax = std::log((double)max_len);
    for (uint64_t b = 0; b < n_bubbles; b++) {
        uint64_t r = splitmix64(seed ^ (b * 0xD1B54A32D192ED03ull));
        auto next = [&]() { return r = splitmix64(r); };
        const uint32_t L = (uint32_t)std::min<double>(max_len, std::max<double>(min_len, std::exp(lmin + (lmax - lmin) * u01(next()))));
        const uint32_t nr = 2 + (uint32_t)(next() % (uint64_t)(max_rows - 1));
        std::string first(L, 'A');
        for (auto &c : first) c = ALPHA[next() & 3];
        std::vector<std::string> seqs{first};
        for (uint32_t q = 1; q < nr; q++) {
            std::string s;
            s.reserve(L + L / 2);
            const uint32_t ilen = std::max<uint32_t>(10, (uint32_t)(L * (0.05 + 0.35 * u01(next()))));
            const bool ins = (next() & 1) && L + ilen <= max_len;
            const bool can_del = L > ilen + 2 * k;
            const uint32_t at = k + (uint32_t)(next() % (uint64_t)std::max<int64_t>(1, (int64_t)L - 2 * k - (ins ? 0 : ilen)));
            for (uint32_t i = 0; i < L; i++) {
                if (i == at && ins) for (uint32_t t = 0; t < ilen; t++) s.push_back(ALPHA[next() & 3]);
                if (!ins && can_del && i >= at && i < at + ilen) continue;
                const double u = u01(next());
                if (u < 0.01) { const char c = first[i]; const uint32_t ix = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3; s.push_back(ALPHA[(ix + 1 + next() % 3) & 3]); }
                else if (u < 0.0105) { s.push_back(first[i]); for (uint32_t t = 0, n = 1 + next() % 5; t < n; t++) s.push_back(ALPHA[next() & 3]); }
                else if (u < 0.011) { i += (uint32_t)(next() % 5); }
                else s.push_back(first[i]);
            }
            if (s.size() > max_len) s.resize(max_len);
            if (s.size() < k) s = first.substr(0, std::max<uint32_t>(k, L / 2));
            if (std::find(seqs.begin(), seqs.end(), s) == seqs.end()) seqs.push_back(s);
        }
        if (seqs.size() < 2) { std::string s = first; s[s.size() / 2] = s[s.size() / 2] == 'A' ? 'C' : 'A'; seqs.push_back(s); }
        sort_branching(seqs, 0, (int)seqs.size() - 1);
        for (auto &s : seqs) {
            w->bases.insert(w->bases.end(), s.begin(), s.end());
            w->seq_off.push_back(w->bases.size());
        }
        w->bubble_off.push_back((uint32_t)(w->seq_off.size() - 1));
        w->bubble_type.push_back(0);
        for (uint32_t i = 0; i < 100; i++) w->ent_bases.push_back(ALPHA[next() & 3]);
        w->ent_off.push_back(w->ent_bases.size());
        w->ent_size.push_back(100);
        w->exit_size.push_back(100);
    }
    return n_bubbles;
}

uint64_t pfs_n_seq(void *h) { return ((Workload *)h)->seq_off.size() - 1; }
uint64_t pfs_n_bases(void *h) { return ((Workload *)h)->bases.size(); }
uint64_t pfs_n_ent_bases(void *h) { return ((Workload *)h)->ent_bases.size(); }
const char *pfs_bases(void *h) { return ((Workload *)h)->bases.data(); }
const uint64_t *pfs_seq_off(void *h) { return ((Workload *)h)->seq_off.data(); }
const uint32_t *pfs_bubble_off(void *h) { return ((Workload *)h)->bubble_off.data(); }
const uint8_t *pfs_bubble_type(void *h) { return ((Workload *)h)->bubble_type.data(); }
const uint32_t *pfs_ent_size(void *h) { return ((Workload *)h)->ent_size.data(); }
const uint32_t *pfs_exit_size(void *h) { return ((Workload *)h)->exit_size.data(); }
const char *pfs_ent_bases(void *h) { return ((Workload *)h)->ent_bases.data(); }
const uint64_t *pfs_ent_off(void *h) { return ((Workload *)h)->ent_off.data(); }

}  // extern "C"
