// tests/dropin/abi_over_oracle.cpp -- TEST INFRASTRUCTURE, never shipped, never linked into libpfgpu.so.
//
// A stand-in for the handful of C-ABI entry points include/pf_caller.hpp calls (pf_init, pf_kmc_open, pf_kmc_cov, pf_align,
// pf_site_cov, ...) computed by the CPU oracle (oracle/libpforacle.so) so that the HOST logic of the batched caller -- branch
// gating and ordering, the strict-bubble arithmetic, VarDis, the row text of both dialects -- is exercised by the `-m "not gpu"`
// suite against the reference's own files.  The per-site part below restates the reference's site k-mers (CDBG.cpp:2331-2473),
// class sets and gate (:2393-2418) in C++ (the same restatement as oracle/caller.py, which pins it against those files too).
// The product path never sees this file: tests/test_cpu_hostlogic.py builds it into a scratch directory.
#include <algorithm>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "pf_gpu.h"

extern "C" {
void *pforc_kmc_open(const char *prefix);
void pforc_kmc_close(void *h);
int pforc_kmc_info(void *h, pf_kmc_info_t *o);
int pforc_kmc_counts(void *h, const char *bases, const uint64_t *seq_off, uint32_t n_seq, int mode, int use_read_api, int n_threads,
                     uint32_t *counts, uint8_t *found);
int pforc_kmc_cov(void *h, const char *bases, const uint64_t *seq_off, uint32_t n_seq, int mode, uint32_t low, uint32_t up,
                  int n_threads, pf_cov_t *out);
void *pforc_align(double M, double D, double G, const char *bases, const uint64_t *seq_off, const uint32_t *bubble_off,
                  uint32_t n_bubbles, int n_threads, pf_msa_batch_t *out);
void pforc_msa_free(void *h);
}

struct pf_ctx {
    void *msa_handle = nullptr;
    pf_msa_batch_t msa;
    bool has_msa = false;
    std::vector<uint64_t> sk_site_off, sk_key_off, sk_keys;   // pf_site_kmers
    std::vector<uint8_t> sk_status;
};
struct pf_kmc {
    pf_ctx *ctx = nullptr;
    void *orc = nullptr;
    uint32_t k = 0;
    std::vector<uint64_t> site_off, cov_off, cov;
    std::vector<uint8_t> status, n_class;
};

namespace {
std::string g_error;

// the k-mer every row contributes at variable column c (CDBG.cpp:2338-2388 indel sites, :2433-2472 SNP sites); false where the
// reference itself would read outside a row
bool site_kmers(const std::vector<std::string> &rows, size_t c, size_t k, bool is_indel, size_t n_indel_before, std::vector<std::string> &out) {
    const size_t n = rows.size(), L = rows[0].size();
    out.assign(n, "");
    if (is_indel) {
        std::vector<size_t> cur(n, c);
        std::vector<std::string> ext(n);
        for (;;) {                                                   // :2338-2357
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
            if (n_indel_before == 0) {                               // :2358-2365
                if (c + e < k) return false;
                out[r] = rows[r].substr(c - k + e, k - e) + ext[r];
            } else {                                                 // :2366-2388
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
    if (n_indel_before > 0) {                                        // :2433-2465
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
    for (size_t r = 0; r < n; r++) out[r] = rows[r].substr(c - k + 1, k);   // :2469-2472
    return true;
}
}  // namespace

extern "C" {

const char *pf_last_error(void) { return g_error.c_str(); }

int pf_init(int, pf_ctx **out) { *out = new pf_ctx(); return PF_OK; }

void pf_shutdown(pf_ctx *ctx) {
    if (!ctx) return;
    if (ctx->msa_handle) pforc_msa_free(ctx->msa_handle);
    delete ctx;
}

int pf_kmc_open(pf_ctx *ctx, const char *prefix, pf_kmc **out) {
    void *h = pforc_kmc_open(prefix);
    if (!h) { g_error = std::string("cannot open ") + prefix; return PF_E_IO; }
    pf_kmc *db = new pf_kmc();
    db->ctx = ctx; db->orc = h;
    pf_kmc_info_t info;
    pforc_kmc_info(h, &info);
    db->k = info.kmer_length;
    *out = db;
    return PF_OK;
}

int pf_kmc_info(const pf_kmc *db, pf_kmc_info_t *info) { return pforc_kmc_info(db->orc, info) == 0 ? PF_OK : PF_E_INVALID; }

// the site k-mers themselves (the coloured caller asks the graph about them before any database is read, CCDBG.cpp:1127)
int pf_site_kmers(pf_ctx *ctx, uint32_t k, const uint8_t *skip, pf_site_kmers_t *out) {
    if (!ctx->has_msa) { g_error = "pf_site_kmers: no alignment on this context"; return PF_E_INVALID; }
    const pf_msa_batch_t &m = ctx->msa;
    const uint32_t nb = m.n_bubbles;
    ctx->sk_site_off.assign(m.var_off, m.var_off + nb + 1);
    ctx->sk_key_off.assign(m.cls_off, m.cls_off + nb + 1);
    ctx->sk_status.assign(m.var_off[nb], PF_SITE_SKIPPED);
    ctx->sk_keys.assign(m.cls_off[nb], 0);
    for (uint32_t b = 0; b < nb; b++) {
        const uint32_t nr = m.n_rows[b], L = m.aln_len[b];
        if ((skip && skip[b]) || nr == 0 || m.status[b] != PF_BUBBLE_OK) continue;
        std::vector<std::string> rows(nr);
        for (uint32_t r = 0; r < nr; r++) rows[r].assign(m.rows + m.rows_off[b] + (size_t)r * L, L);
        size_t n_indel = 0;
        for (uint64_t v = m.var_off[b]; v < m.var_off[b + 1]; v++) {
            const bool is_indel = m.var_kind[v] == 1;
            std::vector<std::string> km;
            const bool formed = site_kmers(rows, m.var_col[v], k, is_indel, n_indel, km);
            if (is_indel) n_indel++;
            if (!formed) { ctx->sk_status[v] = PF_SITE_UNDEFINED; continue; }
            ctx->sk_status[v] = PF_SITE_OK;
            for (uint32_t r = 0; r < nr; r++) {
                uint64_t key = 0;
                for (char ch : km[r]) key = key << 2 | (uint64_t)(ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : 3);
                ctx->sk_keys[m.cls_off[b] + (v - m.var_off[b]) * nr + r] = key;
            }
        }
    }
    out->n_bubbles = nb; out->reserved = 0;
    out->site_off = ctx->sk_site_off.data(); out->key_off = ctx->sk_key_off.data(); out->keys = ctx->sk_keys.data(); out->status = ctx->sk_status.data();
    return PF_OK;
}

int pf_kmc_close(pf_kmc *db) {
    if (db) { pforc_kmc_close(db->orc); delete db; }
    return PF_OK;
}

int pf_kmc_cov(pf_kmc *db, const char *bases, const uint64_t *seq_off, uint32_t n_seq, int mode, uint32_t low, uint32_t up, pf_cov_t *out) {
    return pforc_kmc_cov(db->orc, bases, seq_off, n_seq, mode, low, up, 4, out);
}

int pf_align(pf_ctx *ctx, double M, double D, double G, const char *bases, const uint64_t *seq_off, const uint32_t *bubble_off,
             uint32_t n_bubbles, pf_msa_batch_t *out) {
    if (ctx->msa_handle) pforc_msa_free(ctx->msa_handle);
    ctx->msa_handle = pforc_align(M, D, G, bases, seq_off, bubble_off, n_bubbles, 4, &ctx->msa);
    ctx->has_msa = true;
    *out = ctx->msa;
    return PF_OK;
}

int pf_site_cov(pf_kmc *db, uint32_t low, uint32_t up, const uint8_t *skip, pf_site_batch_t *out) {
    pf_ctx *ctx = db->ctx;
    if (!ctx->has_msa) { g_error = "pf_site_cov: no alignment on this context"; return PF_E_INVALID; }
    const pf_msa_batch_t &m = ctx->msa;
    const uint32_t nb = m.n_bubbles;
    const size_t k = db->k;
    db->site_off.assign(m.var_off, m.var_off + nb + 1);
    db->cov_off.assign(m.cls_off, m.cls_off + nb + 1);
    db->status.assign(m.var_off[nb], PF_SITE_SKIPPED);
    db->n_class.assign(m.var_off[nb], 0);
    db->cov.assign(m.cls_off[nb], 0);
    for (uint32_t b = 0; b < nb; b++) {
        const uint32_t nr = m.n_rows[b], L = m.aln_len[b];
        const uint64_t v0 = m.var_off[b], v1 = m.var_off[b + 1];
        const uint16_t *cls = m.cls + m.cls_off[b];
        for (uint64_t v = v0; v < v1; v++) {
            uint16_t mx = 0;
            for (uint32_t r = 0; r < nr; r++) mx = std::max(mx, cls[(v - v0) * nr + r]);
            db->n_class[v] = (uint8_t)mx;
        }
        if ((skip && skip[b]) || nr == 0 || m.status[b] != PF_BUBBLE_OK) continue;
        std::vector<std::string> rows(nr);
        for (uint32_t r = 0; r < nr; r++) rows[r].assign(m.rows + m.rows_off[b] + (size_t)r * L, L);
        size_t n_indel = 0;
        for (uint64_t v = v0; v < v1; v++) {
            const size_t i = (size_t)(v - v0);
            const bool is_indel = m.var_kind[v] == 1;
            std::vector<std::string> km;
            const bool formed = site_kmers(rows, m.var_col[v], k, is_indel, n_indel, km);
            if (is_indel) n_indel++;                                  // :2390
            if (!formed) { db->status[v] = PF_SITE_UNDEFINED; continue; }
            // distinct k-mers per class in std::set order, classes ascending (:2393-2418)
            std::vector<std::set<std::string>> sets(db->n_class[v]);
            for (uint32_t r = 0; r < nr; r++) sets[cls[i * nr + r] - 1].insert(km[r]);
            uint8_t st = PF_SITE_OK;
            uint64_t *cv = db->cov.data() + m.cls_off[b] + i * nr;
            for (size_t q = 0; q < sets.size() && st == PF_SITE_OK; q++)
                for (const std::string &s : sets[q]) {
                    uint64_t off[2] = {0, s.size()};
                    uint32_t cnt = 0;
                    uint8_t found = 0;
                    pforc_kmc_counts(db->orc, s.data(), off, 1, PF_LOOKUP_FWD_THEN_RC, 0, 1, &cnt, &found);
                    if (!found) { st = PF_SITE_MISSING; break; }
                    if (!(cnt > low && cnt < up)) { st = PF_SITE_DROPPED; break; }
                    cv[q] += cnt;
                }
            db->status[v] = st;
        }
    }
    out->n_bubbles = nb; out->reserved = 0;
    out->site_off = db->site_off.data(); out->status = db->status.data(); out->n_class = db->n_class.data();
    out->cov_off = db->cov_off.data(); out->cov = db->cov.data();
    return PF_OK;
}

}  // extern "C"
