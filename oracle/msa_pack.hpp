// msa_pack.hpp -- TEST INFRASTRUCTURE (oracle side).  Packs per-bubble SequenceAlignment results
// (the five output vectors of SeqAlign::SequenceAlignment, SeqAlign.cpp:550) into the flat
// pf_msa_batch_t layout of include/pf_types.h.  Used by the CPU oracle and by the reference shim;
// never linked into the product library.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "../include/pf_types.h"

namespace pforacle {

struct MsaResult {                      // one bubble, in the reference's own output shape
    std::vector<std::string> rows;      // `str` after the call (empty = dropped)
    std::vector<uint32_t> snp_pos, indel_pos, indel_len;
    std::vector<std::vector<unsigned short>> partition;  // [col][row]
};

struct MsaPacked {                      // owns the arrays a pf_msa_batch_t points into
    std::vector<int32_t> status;
    std::vector<uint32_t> n_rows, aln_len, var_col, ilen;
    std::vector<uint64_t> rows_off, var_off, cls_off, ilen_off;
    std::vector<char> rows;
    std::vector<uint8_t> var_kind;
    std::vector<uint16_t> cls;

    void pack(const std::vector<MsaResult> &res) {
        size_t n = res.size();
        status.assign(n, 0);
        n_rows.assign(n, 0);
        aln_len.assign(n, 0);
        rows_off.assign(n + 1, 0);
        var_off.assign(n + 1, 0);
        cls_off.assign(n + 1, 0);
        ilen_off.assign(n + 1, 0);
        rows.clear(); var_col.clear(); var_kind.clear(); cls.clear(); ilen.clear();
        for (size_t b = 0; b < n; b++) {
            const MsaResult &r = res[b];
            n_rows[b] = (uint32_t)r.rows.size();
            aln_len[b] = r.rows.empty() ? 0u : (uint32_t)r.rows[0].size();
            for (const std::string &s : r.rows) rows.insert(rows.end(), s.begin(), s.end());
            if (!r.rows.empty()) {
                size_t is = 0, ii = 0;
                for (size_t c = 0; c < r.partition.size(); c++) {
                    const std::vector<unsigned short> &p = r.partition[c];
                    bool nz = false;
                    for (unsigned short v : p) nz = nz || v != 0;
                    bool in_snp = false, in_ind = false;
                    while (is < r.snp_pos.size() && r.snp_pos[is] < c) is++;
                    while (ii < r.indel_pos.size() && r.indel_pos[ii] < c) ii++;
                    in_snp = is < r.snp_pos.size() && r.snp_pos[is] == c;
                    in_ind = ii < r.indel_pos.size() && r.indel_pos[ii] == c;
                    if (!nz && !in_snp && !in_ind) continue;
                    var_col.push_back((uint32_t)c);
                    var_kind.push_back(in_snp ? 0 : (in_ind ? 1 : 2));
                    for (unsigned short v : p) cls.push_back(v);
                }
                ilen.insert(ilen.end(), r.indel_len.begin(), r.indel_len.end());
            }
            rows_off[b + 1] = rows.size();
            var_off[b + 1] = var_col.size();
            cls_off[b + 1] = cls.size();
            ilen_off[b + 1] = ilen.size();
        }
    }

    void view(pf_msa_batch_t *out) const {
        out->n_bubbles = (uint32_t)n_rows.size();
        out->reserved = 0;
        out->status = status.data();
        out->n_rows = n_rows.data();
        out->aln_len = aln_len.data();
        out->rows_off = rows_off.data();
        out->rows = rows.data();
        out->var_off = var_off.data();
        out->var_col = var_col.data();
        out->var_kind = var_kind.data();
        out->cls_off = cls_off.data();
        out->cls = cls.data();
        out->ilen_off = ilen_off.data();
        out->ilen = ilen.data();
    }
};

}  // namespace pforacle
