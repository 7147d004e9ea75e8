// align_emu.cpp -- TEST INFRASTRUCTURE.  Runs the product's per-bubble alignment state machines
// (ploidyfrost_b200/csrc/pf_align_core.cuh: traceback, progressive-MSA filter, site caller) on the CPU with a
// trivial single-thread execution policy and a plain row-by-row fill, so the device logic can be
// diffed against the oracle without a GPU.  Not part of libpfgpu.so and never used as a fallback.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <thread>
#include <atomic>

#include "../../ploidyfrost_b200/csrc/pf_align_core.cuh"
#include "../../oracle/msa_pack.hpp"

using namespace pfalign;

namespace {

uint32_t g_skew_T = 8;   // lanes of the emulated CTA-wide fill (LAYOUT_SKEW)

template <int LAYOUT>
struct HostExec : SerialHelpers {
    static constexpr int kLayout = LAYOUT;
    static constexpr bool DIAG = LAYOUT == LAYOUT_DIAG;
    uint32_t skew_T() const { return g_skew_T; }
    uint32_t skew_w(uint32_t n) const { return (n + g_skew_T) / g_skew_T; }
    uint64_t fidx(uint32_t i, uint32_t j, uint32_t m, uint32_t n) const {
        return LAYOUT == LAYOUT_SKEW ? skew_index(i, j, skew_w(n), g_skew_T) : (uint64_t)flag_index<DIAG>(i, j, m, n);
    }
    std::vector<int> rowbuf;
    bool leader() const { return true; }
    uint32_t bcast(uint32_t v) const { return v; }
    int bcast_i(int v) const { return v; }
    uint32_t bcast_ld(const uint32_t *p) const { return *p; }
    void sync() const {}
    uint64_t steps = 0;
    void note_steps(uint64_t n) { steps += n; }
    uint32_t pitch_n(uint32_t n) const { return n; }
    void fill(const BV flags, const CBV A, uint32_t m, const CBV B, uint32_t n, const Scoring &sc, int32_t *) {
        rowbuf.assign(2 * (size_t)(n + 1), 0);
        int *prev = rowbuf.data(), *cur = rowbuf.data() + n + 1;
        flags[0] = 0;
        for (uint32_t j = 1; j <= n; j++) { prev[j] = pack_sf(border_score(sc, j), F_LEFT); flags[fidx(0, j, m, n)] = F_LEFT; }
        prev[0] = pack_sf(0, 0);
        for (uint32_t i = 1; i <= m; i++) {
            cur[0] = pack_sf(border_score(sc, i), F_UP);
            flags[fidx(i, 0, m, n)] = F_UP;
            const bool block_left = (i != m) && A[i] == '-';
            for (uint32_t j = 1; j <= n; j++) {
                cur[j] = nw_cell(sc, prev[j], prev[j - 1], cur[j - 1], A[i - 1], B[j - 1], block_left);
                flags[fidx(i, j, m, n)] = (uint8_t)unpack_f(cur[j]);
            }
            std::swap(prev, cur);
        }
    }
};

Limits g_lim = {16, 0, 0, 16, 16, 0, 50000000ull, 0, 0};
uint32_t g_lanes = 1;   // 32: lay the work area out lane-interleaved like msa_lane_kernel and use slot `g_lane`
uint32_t g_lane = 0;
std::vector<uint64_t> g_steps;   // traceback iterations per bubble of the last pfemu_align call

}  // namespace

extern "C" {

// diag = 1: diagonal-major flag bytes (msa_warp_kernel); lanes = 32: lane-interleaved work area (msa_lane_kernel)
// diag = 2: the skewed layout of msa_cta_kernel with `lane` emulated lanes (work area contiguous)
void pfemu_set_layout(uint32_t diag, uint32_t lanes, uint32_t lane) {
    g_lim.diag_flags = diag; g_lanes = lanes; g_lane = lane;
    if (diag == LAYOUT_SKEW) { g_skew_T = lane ? lane : 8; g_lim.pad_ = g_skew_T; g_lanes = 1; g_lane = 0; }
}

void pfemu_set_limits(uint32_t max_rows, uint32_t k_cand, uint32_t k_aln, uint32_t max_alen, uint32_t max_var) {
    g_lim.max_rows = max_rows; g_lim.k_cand = k_cand; g_lim.k_aln = k_aln; g_lim.max_alen = max_alen; g_lim.max_var = max_var;
}

// stats[0] = max co-optimal alignments seen is not tracked here; kept simple.
void *pfemu_align(double M, double D, double G, const char *bases, const uint64_t *seq_off, const uint32_t *bubble_off,
                  uint32_t n_bubbles, int n_threads, pf_msa_batch_t *out) {
    std::vector<pforacle::MsaResult> res(n_bubbles);
    std::vector<int32_t> status(n_bubbles, 0);
    g_steps.assign(n_bubbles, 0);
    const Scoring sc = make_scoring(M, D, G);
    std::atomic<uint32_t> next(0);
    auto worker = [&]() {
        HostExec<LAYOUT_DIAG> xd;
        HostExec<LAYOUT_ROW> xr;
        HostExec<LAYOUT_SKEW> xs;
        std::vector<uint8_t> wbuf, slot;
        for (;;) {
            const uint32_t b = next.fetch_add(1);
            if (b >= n_bubbles) return;
            const uint32_t s0 = bubble_off[b], ns = bubble_off[b + 1] - s0;
            Limits lim = g_lim;
            uint64_t sum = 0, mx = 0;
            for (uint32_t s = 0; s < ns; s++) {
                const uint64_t l = seq_off[s0 + s + 1] - seq_off[s0 + s];
                sum += l;
                if (l > mx) mx = l;
            }
            if (lim.max_alen == 0) lim.max_alen = (uint32_t)sum;
            lim.max_blen = (uint32_t)mx;
            if (lim.max_var == 0) lim.max_var = lim.max_alen;
            wbuf.assign(work_area_bytes(lim, g_lanes) + 64, 0);
            const WorkArea ws = carve_work_area(wbuf.data(), lim, g_lanes, g_lanes > 1 ? (b + g_lane) % g_lanes : 0);
            const SlotLayout lay = slot_layout(ns, sum, lim);
            slot.assign(lay.bytes + 64, 0);
            xd.steps = xr.steps = xs.steps = 0;
            if (lim.diag_flags == LAYOUT_SKEW) msa_run(xs, (const uint8_t *)bases, seq_off, s0, ns, ws, lim, sc, slot.data());
            else if (lim.diag_flags) msa_run(xd, (const uint8_t *)bases, seq_off, s0, ns, ws, lim, sc, slot.data());
            else msa_run(xr, (const uint8_t *)bases, seq_off, s0, ns, ws, lim, sc, slot.data());
            g_steps[b] = xd.steps + xr.steps + xs.steps;
            const SlotHdr *h = (const SlotHdr *)slot.data();
            status[b] = h->status;
            pforacle::MsaResult &r = res[b];
            if (h->status == 0 && h->n_rows) {
                for (uint32_t q = 0; q < h->n_rows; q++)
                    r.rows.emplace_back((const char *)slot.data() + lay.off_rows + (size_t)q * h->alen, h->alen);
                const uint32_t *vc = (const uint32_t *)(slot.data() + lay.off_varcol);
                const uint8_t *vk = slot.data() + lay.off_kind;
                const uint16_t *cl = (const uint16_t *)(slot.data() + lay.off_cls);
                const uint32_t *il = (const uint32_t *)(slot.data() + lay.off_ilen);
                r.partition.assign(h->alen, std::vector<unsigned short>(h->n_rows, 0));
                for (uint32_t v = 0; v < h->n_var; v++) {
                    if (vk[v] == 0) r.snp_pos.push_back(vc[v]);
                    else if (vk[v] == 1) r.indel_pos.push_back(vc[v]);
                    for (uint32_t q = 0; q < h->n_rows; q++) r.partition[vc[v]][q] = cl[(size_t)v * h->n_rows + q];
                }
                r.indel_len.assign(il, il + h->n_ilen);
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < (n_threads < 1 ? 1 : n_threads); t++) th.emplace_back(worker);
    for (auto &t : th) t.join();
    pforacle::MsaPacked *p = new pforacle::MsaPacked();
    p->pack(res);
    p->status = status;
    p->view(out);
    return p;
}

const uint64_t *pfemu_last_steps() { return g_steps.data(); }

void pfemu_msa_free(void *h) { delete (pforacle::MsaPacked *)h; }

}  // extern "C"
