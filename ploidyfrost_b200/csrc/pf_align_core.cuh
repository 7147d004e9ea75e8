// pf_align_core.cuh -- the per-bubble alignment machine of libpfgpu.so (SeqAlign on the GPU).
//
// Everything here is written for ONE warp working on ONE bubble out of a flat work area in HBM:
//   * the DP fill is warp-cooperative (pf_align.cu: lane = matrix row, anti-diagonal wavefront by shuffle);
//   * the parts of SeqAlign that are inherently sequential and data dependent -- the co-optimal
//     traceback DFS (SeqAlign.cpp:306-478), the progressive-MSA filter (:559-638) and the site caller
//     (compareStrPair, :8-236) -- run on the warp's leader lane as the state machines below.
//
// Design notes (what differs from the reference, results identical):
//   * cell flags are one byte: base flags in bits 0-2 (the by-value `matrix`, permanently pruned) and
//     "already tried" marks in bits 4-6 (`matrix_temp` = base & ~tried; the fill writes the 3 base bits only), stored diagonal-major so the wavefront writes coalesce:
//     byte (i,j) lives at (i+j)*(m+1)+i.  1 byte/cell instead of the reference's 16-byte MatrixUnit x 3 copies.
//   * the DFS keeps a move string (L/U/D per step) instead of the prepend-built resA/resB/gap_pos and the
//     (i,j) stack; every test the reference makes on resA[0]/resA[1]/resB[0]/resB[1] is a test on the last
//     two moves because '+' only ever comes from a Left move and '-' in resB only from an Up move.
//   * aligned strings are never materialised during the search: score / #positions / #indels of a
//     candidate (variantAnalyze, :237-305) are computed by replaying the move string.
//   * size_t counters indel1/indel2 are uint64_t here and wrap exactly like the reference's (:454-467).
//
// The functions are __host__ __device__ so that tests/hostemu can run the very same state machines on the
// CPU against the oracle (development aid only; the product library has no CPU path).
#pragma once
#include <limits.h>
#include <stdint.h>

#include "../../include/pf_types.h"

#if defined(__CUDACC__)
#define PF_HD __host__ __device__ __forceinline__
#define PF_HDN __host__ __device__
#else
#define PF_HD inline
#define PF_HDN
#endif

namespace pfalign {

// hint: the flag byte the DFS will most likely want a few steps from now (no-op on the host)
PF_HD void prefetch_byte(const uint8_t *p) {
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.L1 [%0];" ::"l"(p));   // generic address: a no-op when the flags live in shared memory
#else
    (void)p;
#endif
}

enum { F_UP = 1, F_DIAG = 2, F_LEFT = 4 };
enum { MV_L = 0, MV_U = 1, MV_D = 2, MV_NONE = 0xFF };
enum { PF_BUBBLE_OUT_OVERFLOW = 6 };  // more variable columns than the output slot holds

// Strided views.  Every per-bubble array of the work area is addressed through a view so that the same
// state machines run (a) on one warp's private, contiguous area (stride 1: the generic kernel, the host
// emulation) and (b) one bubble per LANE with the 32 lanes' arrays interleaved element by element
// (stride 32: lanes that advance in lock-step touch 32 consecutive bytes = one sector).
struct BV {
    uint8_t *p;
    uint32_t s;
    PF_HD uint8_t &operator[](uint64_t i) const { return p[i * s]; }
    PF_HD BV operator+(uint64_t o) const { BV r; r.p = p + o * s; r.s = s; return r; }
};
struct CBV {
    const uint8_t *p;
    uint32_t s;
    PF_HD uint8_t operator[](uint64_t i) const { return p[i * s]; }
    PF_HD CBV operator+(uint64_t o) const { CBV r; r.p = p + o * s; r.s = s; return r; }
};
struct WV {   // uint32 elements, stride in elements
    uint32_t *p;
    uint32_t s;
    PF_HD uint32_t &operator[](uint64_t i) const { return p[i * s]; }
};
PF_HD BV bv(uint8_t *p, uint32_t s = 1) { BV r; r.p = p; r.s = s; return r; }
PF_HD CBV cbv(const uint8_t *p, uint32_t s = 1) { CBV r; r.p = p; r.s = s; return r; }
PF_HD CBV cbv(const BV &v) { CBV r; r.p = v.p; r.s = v.s; return r; }
PF_HD WV wv(uint32_t *p, uint32_t s = 1) { WV r; r.p = p; r.s = s; return r; }

// Flag-byte address of cell (i,j) of an (m+1) x (n+1) matrix.  Row-major storage may use a row pitch wider than the
// matrix (pass n = pitch - 1): the group kernel gives every bubble of a launch the same pitch (X::pitch_n).
template <bool DIAG>
PF_HD uint32_t flag_index(uint32_t i, uint32_t j, uint32_t m, uint32_t n) {
    return DIAG ? (i + j) * (m + 1) + i : i * (n + 1) + j;
}
// Layouts of the flag bytes (X::kLayout): 0 row-major with a pitch, 1 diagonal-major, 2 SKEWED for the CTA-wide fill:
// T lanes own w-column strips (lane l: columns l*w .. l*w+w-1, column 0 = the border) and lane l is at row t - l at time step t,
// so the byte of cell (i, j) lives at ((i + j/w) * w + j%w) * T + j/w -- the lanes of a warp store 32 consecutive bytes per
// column of their strips.  (0,0) is byte 0 in every layout.
enum { LAYOUT_ROW = 0, LAYOUT_DIAG = 1, LAYOUT_SKEW = 2 };
PF_HD uint64_t skew_index(uint32_t i, uint32_t j, uint32_t w, uint32_t T) {
    const uint32_t l = j / w, q = j - l * w;
    return ((uint64_t)(i + l) * w + q) * T + l;
}

struct Scoring {   // SeqAlign(double&,double&,double&), SeqAlign.hpp:10
    double M, D, G;
    int iM, iD, iG;
    int integral;  // all three are whole numbers: pure INT32 path, bit-identical to the double arithmetic
};

PF_HD Scoring make_scoring(double M, double D, double G) {
    Scoring s;
    s.M = M; s.D = D; s.G = G;
    s.iM = (int)M; s.iD = (int)D; s.iG = (int)G;
    s.integral = ((double)s.iM == M && (double)s.iD == D && (double)s.iG == G && M < 1e6 && M > -1e6 && D < 1e6 &&
                  D > -1e6 && G < 1e6 && G > -1e6) ? 1 : 0;
    return s;
}

// `int x = long + double` (SeqAlign.cpp:512, :517, :522).  INTEGRAL is a compile-time switch so that the
// integer kernels carry no FP64 instructions at all.
template <bool INTEGRAL>
PF_HD int add_trunc(const Scoring &sc, int s, int iv, double dv) {
    return INTEGRAL ? s + iv : (int)((double)s + dv);
}
// border scores `long = GAP * i` (SeqAlign.cpp:489, :494)
PF_HD int border_score(const Scoring &sc, uint32_t i) {
    return sc.integral ? sc.iG * (int)i : (int)(long long)(sc.G * (double)i);
}

PF_HD int pack_sf(int score, int flags) { return score * 8 + flags; }
PF_HD int unpack_s(int p) { return p >> 3; }
PF_HD int unpack_f(int p) { return p & 7; }

// One DP cell (SeqAlign.cpp:512-545).  up/dg/lf are the packed (score,flags) of the three neighbours;
// block_left is `i != m && A[i] == '-'` (the profile rule, :528-532).
template <bool INTEGRAL>
PF_HD int nw_cell_t(const Scoring &sc, int up, int dg, int lf, uint8_t a, uint8_t b, bool block_left) {
    int s_up = add_trunc<INTEGRAL>(sc, unpack_s(up), sc.iG, sc.G) + ((unpack_f(up) & F_UP) ? 1 : 0);
    int s_dg;
    if (a == b) s_dg = add_trunc<INTEGRAL>(sc, unpack_s(dg), sc.iM, sc.M);                    // :498-506, equality first
    else if (a == '-' || b == '-') s_dg = add_trunc<INTEGRAL>(sc, unpack_s(dg), sc.iG, sc.G);
    else s_dg = add_trunc<INTEGRAL>(sc, unpack_s(dg), sc.iD, sc.D);
    s_dg += (unpack_f(dg) & F_DIAG) ? 1 : 0;
    int s_lf = add_trunc<INTEGRAL>(sc, unpack_s(lf), sc.iG, sc.G) + ((unpack_f(lf) & F_LEFT) ? 1 : 0);
    int best = s_up > s_dg ? s_up : s_dg;
    if (s_lf > best) best = s_lf;
    if (best == s_lf && block_left) {
        s_lf = INT_MIN;
        best = s_up > s_dg ? s_up : s_dg;
    }
    int f = (s_up == best ? F_UP : 0) | (s_dg == best ? F_DIAG : 0) | (s_lf == best ? F_LEFT : 0);
    return pack_sf(best, f);
}
PF_HD int nw_cell(const Scoring &sc, int up, int dg, int lf, uint8_t a, uint8_t b, bool block_left) {
    return sc.integral ? nw_cell_t<true>(sc, up, dg, lf, a, b, block_left) : nw_cell_t<false>(sc, up, dg, lf, a, b, block_left);
}

// AlignUnit::operator- (SeqAlign.hpp:43-67) truncated to int like its call sites (SeqAlign.cpp:334, :599).
PF_HD int rank_diff(long long ls, uint32_t lp, uint32_t li, long long rs, uint32_t rp, uint32_t ri) {
    if (ls != rs) return ls > rs ? 1 : -1;
    if (lp != rp) return (int)((long long)rp - (long long)lp);
    if (li != ri) return (int)((long long)ri - (long long)li);
    return 0;
}

struct PairKey {
    long long score;
    uint32_t n_pos, n_indel;
};

// variantAnalyze (SeqAlign.cpp:237-305) over the alignment spelled by a move string.  `row` supplies the
// characters consumed by Up/Diag moves (row 0 of the profile for the pair itself, an earlier MSA row for
// the projection of :583-598); Left moves put a gap into it.  Moves are stored in DFS order, i.e. the
// alignment reads from mv[depth-1] down to mv[0].
// The loops below are latency bound on the GPU (one thread, dependent global loads), so they work in chunks
// of PF_CH elements: all loads of a chunk are issued before any of them is used.
#define PF_CH 8
// dst[0..n) = src[0..n), chunked
PF_HD void copy_bytes(const CBV src, const BV dst, uint32_t n) {
    for (uint32_t c0 = 0; c0 < n; c0 += PF_CH) {
        uint8_t v[PF_CH];
#pragma unroll
        for (uint32_t q = 0; q < PF_CH; q++) v[q] = c0 + q < n ? src[c0 + q] : (uint8_t)0;
#pragma unroll
        for (uint32_t q = 0; q < PF_CH; q++) if (c0 + q < n) dst[c0 + q] = v[q];
    }
}

PF_HD PairKey analyze_moves(const Scoring &sc, const CBV row, const CBV B, const CBV mv, uint32_t depth) {
    PairKey k;
    k.score = 0; k.n_pos = 0; k.n_indel = 0;
    uint32_t ia = 0, jb = 0;
    int run = 0;
    for (uint32_t t0 = depth; t0 > 0;) {
        const uint32_t cnt = t0 < PF_CH ? t0 : PF_CH;
        uint8_t m[PF_CH], a[PF_CH], b[PF_CH];
#pragma unroll
        for (uint32_t q = 0; q < PF_CH; q++) m[q] = q < cnt ? (uint8_t)(mv[t0 - 1 - q] & 3) : (uint8_t)MV_NONE;
        uint32_t ia_q = ia, jb_q = jb;
#pragma unroll
        for (uint32_t q = 0; q < PF_CH; q++) {
            const bool use_a = q < cnt && m[q] != MV_L, use_b = q < cnt && m[q] != MV_U;
            a[q] = use_a ? row[ia_q] : (uint8_t)'-';
            b[q] = use_b ? B[jb_q] : (uint8_t)'-';
            ia_q += use_a; jb_q += use_b;
        }
        ia = ia_q; jb = jb_q;
#pragma unroll
        for (uint32_t q = 0; q < PF_CH; q++) {
            if (q >= cnt) break;
            if (sc.integral) k.score += (a[q] == '-' || b[q] == '-') ? sc.iG : (a[q] == b[q] ? sc.iM : sc.iD);      // :241-246, gap first
            else k.score = (long long)((double)k.score + ((a[q] == '-' || b[q] == '-') ? sc.G : (a[q] == b[q] ? sc.M : sc.D)));  // long += double
            if (a[q] != b[q]) {
                if (a[q] == '-') { if (run != 1) { run = 1; k.n_indel++; k.n_pos++; } }
                else if (b[q] == '-') { if (run != 2) { run = 2; k.n_indel++; k.n_pos++; } }
                else { run = 0; k.n_pos++; }
            } else run = 0;
        }
        t0 -= cnt;
    }
    return k;
}

struct TbResult {
    uint32_t n_aln;
    int status;
    uint64_t steps;
};

// traceback (SeqAlign.cpp:306-478).  flags: the 3 base bits per cell written by the fill.
// Kept alignments go to ext_mv[a*mv_stride ..] with lengths ext_len[a].
//
// How the reference's two matrices map onto this DFS: `matrix` (by value, permanently pruned) is the flag byte in memory, and the
// only writes to it are the prunes.  `matrix_temp` (which directions of a cell were tried already, reset when the search leaves
// the cell, :432) is NOT stored per cell: a cell on the current path is on it exactly once (every move lowers i + j), the
// directions are tried in the fixed order Left, Up, Diag (:356, :393, :425), so "tried" at the cell the search returns to is
// "everything up to the move that was taken from it" -- a function of the stack entry.  The stack entry (one byte of `mv`) is
// move | base_flags << 2: the base flags of a cell cannot change while it is on the stack (pruning only touches the current
// cell), so backtracking needs no flag load at all -- it walks the stack, whose top eight entries live in a register.
// Readers of a move string mask the entry with 3.
template <int LAYOUT, class X>
PF_HDN inline TbResult traceback(X &x, const BV flags, const CBV A, uint32_t m, const CBV B, uint32_t n,
                                 const Scoring &sc, const BV mv, const BV ext_mv, const WV ext_len,
                                 uint32_t mv_stride, uint32_t k_aln, uint64_t step_limit, uint32_t pitch_n) {
    TbResult r;
    r.n_aln = 0; r.status = PF_BUBBLE_OK; r.steps = 0;
    // The forward walk is ONE dependent chain per step (flag byte -> decision -> next cell), i.e. pure latency on a GPU lane, so it
    // is kept to as few instructions as the semantics allow: the cell index moves by a constant per move instead of being
    // recomputed, (0,0) is cell 0 in both layouts, counters are 32 bit.
    // indel1 / indel2 (size_t in the reference, :312-315) may wrap below zero; a 32-bit counter orders against the small caps
    // exactly like the 64-bit one as long as fewer than 2^31 moves are on the stack.
    constexpr bool DIAG = LAYOUT == LAYOUT_DIAG, SKEW = LAYOUT == LAYOUT_SKEW;
    // skewed layout: a Left move inside a strip is -T, across a strip boundary -(T+1) (q = column inside the strip is tracked)
    const uint32_t sk_T = SKEW ? x.skew_T() : 0u, sk_w = SKEW ? x.skew_w(n) : 1u;
    // (cell indices are 32 bit: the launchers keep every flag area below 2^32 bytes)
    const uint32_t dL = SKEW ? sk_T : (DIAG ? m + 1 : 1u);                      // cell(i, j) - cell(i, j-1)
    const uint32_t dU = SKEW ? sk_w * sk_T : (DIAG ? m + 2 : pitch_n + 1);      // cell(i, j) - cell(i-1, j)
    const uint32_t dD = dL + dU;                               // cell(i, j) - cell(i-1, j-1)
    uint32_t cell = SKEW ? (uint32_t)skew_index(m, n, sk_w, sk_T) : flag_index<DIAG>(m, n, m, pitch_n);
    uint32_t sk_q = SKEW ? n % sk_w : 0u;
    uint32_t depth = 0, steps = 0;
    const uint32_t budget = step_limit > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (uint32_t)step_limit;
    uint32_t open_a = 0, open_b = 0, cap_a = 5, cap_b = 5;     // indel1, indel2, indel1_max, indel2_max
    PairKey last;
    last.score = 0; last.n_pos = 0; last.n_indel = 0;
    const bool pf = x.prefetch_flags();
    const uint32_t pf_min = x.prefetch_cells() * dD;           // that many cells up the diagonal: the path of a good alignment
    uint64_t win = 0;                                          // stack entries depth-1 (low byte) .. depth-8
    uint32_t nwin = 0;                                         // valid entries in win; >= min(depth, 2) between steps
    uint32_t c = (uint32_t)flags[cell] & 7u;                   // base flags of the current cell
    uint32_t tried = 0;                                        // directions already searched from the current cell
    for (;;) {
        if (++steps > budget) { r.status = PF_BUBBLE_STEP_LIMIT; r.steps = steps; return r; }
        const uint32_t lastmv = depth ? (uint32_t)win & 3u : (uint32_t)MV_NONE;
        if (cell == 0 && open_a <= cap_a && open_b <= cap_b) {          // :322-355
            const PairKey cand = x.analyze(sc, A, B, cbv(mv), depth);
            bool keep = true;
            if (r.n_aln > 0) {
                const int d = rank_diff(last.score, last.n_pos, last.n_indel, cand.score, cand.n_pos, cand.n_indel);
                if (d > 0) keep = false;
                else if (d < 0) r.n_aln = 0;
            }
            if (keep) {
                if (r.n_aln == k_aln) { r.status = PF_BUBBLE_CAND_OVERFLOW; r.steps = steps; return r; }
                x.copy(cbv(mv), ext_mv + (uint64_t)r.n_aln * mv_stride, depth);
                ext_len[r.n_aln] = depth;
                r.n_aln++;
                last = cand;
                cap_a = open_a;
                cap_b = open_b;
            }
        }
        const uint32_t w = c & ~tried;                                  // still-untried directions of this cell
        uint32_t move;
        if (w & F_LEFT) {                                               // :356-392
            bool take;
            if (open_a < cap_a) { if (lastmv != MV_L) ++open_a; take = true; }
            else if (open_a == cap_a) take = (lastmv == MV_L);
            else take = false;
            if (!take) { c &= ~(uint32_t)F_LEFT; flags[cell] = (uint8_t)c; continue; }   // permanent prune of the base matrix
            move = MV_L;
        } else if (w & F_UP) {                                          // :393-424
            bool take;
            if (open_b < cap_b) { if (depth == 0 || lastmv == MV_U) ++open_b; take = true; }  // sic (:397)
            else if (open_b == cap_b) take = (lastmv == MV_U);
            else take = false;
            if (!take) { c &= ~(uint32_t)F_UP; flags[cell] = (uint8_t)c; continue; }
            move = MV_U;
        } else if (w & F_DIAG) {                                        // :425-431
            move = MV_D;
        } else {                                                        // :432-474: leave the cell, back to where the search came from
            if (depth == 0) break;
            const uint32_t e = (uint32_t)win & 0xFFu;
            const uint32_t pm = e & 3u;                                 // the move that led here
            const uint32_t prevmv = depth >= 2 ? (uint32_t)(win >> 8) & 3u : (uint32_t)MV_NONE;
            if (pm == MV_L) { if (prevmv != MV_L) --open_a; cell += dL; tried = F_LEFT; }
            else if (pm == MV_U) { if (prevmv != MV_U) --open_b; cell += dU; tried = F_LEFT | F_UP; }   // may wrap below zero, as in the reference
            else { cell += dD; tried = F_LEFT | F_UP | F_DIAG; }
            if (SKEW && pm != MV_U) { if (++sk_q == sk_w) { sk_q = 0; cell += 1; } }   // back over a strip boundary
            c = e >> 2;
            win >>= 8; nwin--; depth--;
            if (nwin < 2 && depth > nwin) {                             // refill the register window: independent loads, one latency
                nwin = depth < 8 ? depth : 8;
                win = 0;
#pragma unroll
                for (uint32_t q = 0; q < 8; q++) if (q < nwin) win |= (uint64_t)mv[depth - 1 - q] << (8 * q);
            }
            continue;
        }
        const uint32_t e = move | (c << 2);
        mv[depth++] = (uint8_t)e;
        win = (win << 8) | e;
        nwin = nwin < 8 ? nwin + 1 : 8;
        cell -= move == MV_L ? dL : (move == MV_U ? dU : dD);
        if (SKEW && move != MV_U) { if (sk_q == 0) { sk_q = sk_w - 1; cell -= 1; } else --sk_q; }
        if (SKEW) x.prefetch_ahead(flags, cell, sk_q, sk_w, sk_T);
        else if (pf && cell >= pf_min) prefetch_byte(&flags[cell - pf_min]);
        c = (uint32_t)flags[cell] & 7u;
        tried = 0;
    }
    r.steps = steps;
    return r;
}

// Writes one aligned row: `src` stretched by the gaps of a move string (`gap_move` = MV_L for the rows of the profile,
// MV_U for the new sequence: that move inserts '-').
PF_HD void project_moves(const CBV src, const CBV mv, uint32_t depth, const BV dst, const uint8_t gap_move) {
    uint32_t ia = 0, c = 0;
    for (uint32_t t0 = depth; t0 > 0;) {
        const uint32_t cnt = t0 < PF_CH ? t0 : PF_CH;
        uint8_t m[PF_CH], v[PF_CH];
#pragma unroll
        for (uint32_t q = 0; q < PF_CH; q++) m[q] = q < cnt ? (uint8_t)(mv[t0 - 1 - q] & 3) : gap_move;
        uint32_t ia_q = ia;
#pragma unroll
        for (uint32_t q = 0; q < PF_CH; q++) {
            const bool use = m[q] != gap_move;
            v[q] = use ? src[ia_q] : (uint8_t)'-';
            ia_q += use;
        }
        ia = ia_q;
#pragma unroll
        for (uint32_t q = 0; q < PF_CH; q++) if (q < cnt) dst[c + q] = v[q];
        c += cnt;
        t0 -= cnt;
    }
}
PF_HD void project_row(const CBV src, const CBV mv, uint32_t depth, const BV dst) { project_moves(src, mv, depth, dst, MV_L); }
PF_HD void project_new(const CBV B, const CBV mv, uint32_t depth, const BV dst) { project_moves(B, mv, depth, dst, MV_U); }

// ---- site calling: compareStrPair (SeqAlign.cpp:8-236) ------------------------------------------------

struct PosStats {  // what compute_dis (:10-38) needs from a position list, gathered while streaming
    uint32_t n, first, last;
    int min_gap;   // min over neighbours of v[i]-v[i-1]-1, folded exactly like the reference's int/size_t dance
    PF_HD void init() { n = 0; first = 0; last = 0; min_gap = 0; }
    PF_HD void push(uint32_t v) {
        if (n == 0) { first = v; min_gap = (int)v; }   // count = v[0]
        else { const int g = (int)(v - last - 1); min_gap = g < min_gap ? g : min_gap; }  // min(int(gap), int(count))
        last = v;
        n++;
    }
    PF_HD uint64_t spacing(uint64_t L) const {
        if (n == 0) return 0;
        if (n == 1) {
            const int left = (int)first;
            const int right = (int)(uint32_t)(L - first) - 1;
            return left > right ? (uint64_t)(long long)(left + 1) : (uint64_t)(long long)right;
        }
        const uint64_t cnt = (uint64_t)(long long)min_gap;
        const uint64_t tail = L - last - 1;                 // size_t arithmetic, may wrap (:34)
        return cnt < tail ? cnt : tail;
    }
};

struct MsaKey {
    int n_snp, n_indel;          // the uint8_t counters (:54-55) as promoted ints
    uint64_t d_indel, d_snp, d_all;
    int site_l, site_r;
};

struct SlotView {   // where the winner is written (one bubble's output slot in HBM)
    uint8_t *rows;      // n_rows * alen
    uint32_t *var_col;
    uint8_t *var_kind;
    uint16_t *cls;      // [var][row]
    uint32_t *ilen;
    uint32_t var_cap;
};

// One pass over a candidate MSA (rows r at cand + r*stride, L columns).  Fills `key`; when `emit` also
// writes var columns / classes / indel lengths into `out` and returns their counts.
PF_HDN inline int scan_candidate(const CBV cand, uint32_t stride, uint32_t nr, uint32_t L, uint64_t Llast,
                                 MsaKey &key, bool emit, SlotView *out, uint32_t *n_var_out, uint32_t *n_ilen_out) {
    PosStats snp, ind, all;
    snp.init(); ind.init(); all.init();
    bool open = false;
    uint32_t n_snp = 0, n_indel = 0, last_indel_pos = 0, n_var = 0, n_ilen = 0;
    int status = PF_BUBBLE_OK;
    // Column summaries (more than one symbol? a gap?) are gathered PF_CH columns at a time -- all loads of a chunk are in
    // flight together; the state machine then walks the chunk.  Only variable columns (a few per cent) load anything else.
    for (uint32_t j0 = 0; j0 < L; j0 += PF_CH) {
      uint32_t multi_m = 0, gap_m = 0;
      {
        uint8_t c0[PF_CH];
#pragma unroll
        for (uint32_t q = 0; q < PF_CH; q++) c0[q] = j0 + q < L ? cand[j0 + q] : (uint8_t)0;
#pragma unroll
        for (uint32_t q = 0; q < PF_CH; q++) gap_m |= (uint32_t)(c0[q] == '-') << q;
        for (uint32_t r = 1; r < nr; r++) {
            uint8_t c[PF_CH];
#pragma unroll
            for (uint32_t q = 0; q < PF_CH; q++) c[q] = j0 + q < L ? cand[(uint64_t)r * stride + j0 + q] : (uint8_t)0;
#pragma unroll
            for (uint32_t q = 0; q < PF_CH; q++) { multi_m |= (uint32_t)(c[q] != c0[q]) << q; gap_m |= (uint32_t)(c[q] == '-') << q; }
        }
      }
      for (uint32_t q = 0; q < PF_CH && j0 + q < L; q++) {
        const uint32_t j = j0 + q;
        const bool multi = (multi_m >> q) & 1u, has_gap = (gap_m >> q) & 1u;
        bool number = false;
        int kind = 2;
        if (multi) {
            if (!has_gap) {                                             // SNP column (:66-93)
                if (open) {
                    if (emit) { if (n_ilen < out->var_cap) out->ilen[n_ilen] = j - last_indel_pos; else status = PF_BUBBLE_OUT_OVERFLOW; }
                    n_ilen++;
                    open = false;
                }
                snp.push(j); all.push(j);
                n_snp = (n_snp + 1) & 0xFF;
                number = true;
                kind = 0;
            } else {                                                    // gap column (:94-146)
                bool continues = true;
                if (open) {
                    for (uint32_t r = 0; r < nr; r++) {
                        const CBV row = cand + (uint64_t)r * stride;
                        if ((row[j] == '-') != (row[j - 1] == '-')) { continues = false; break; }
                    }
                    if (!continues) {
                        if (emit) { if (n_ilen < out->var_cap) out->ilen[n_ilen] = j - last_indel_pos; else status = PF_BUBBLE_OUT_OVERFLOW; }
                        n_ilen++;
                        n_indel = (n_indel + 1) & 0xFF;
                        last_indel_pos = j;
                        ind.push(j); all.push(j);
                        kind = 1;
                    }
                } else {
                    continues = false;
                    n_indel = (n_indel + 1) & 0xFF;
                    last_indel_pos = j;
                    ind.push(j); all.push(j);
                    open = true;
                    kind = 1;
                }
                if (!continues) number = true;
                else {  // continued run: numbered only when the column shows more than two symbols (:121)
                    uint32_t distinct = 1;
                    for (uint32_t r = 1; r < nr && distinct <= 2; r++) {
                        const uint8_t c = cand[(uint64_t)r * stride + j];
                        bool seen = false;
                        for (uint32_t q = 0; q < r; q++) if (cand[(uint64_t)q * stride + j] == c) { seen = true; break; }
                        if (!seen) distinct++;
                    }
                    number = distinct > 2;
                }
            }
        } else if (open) {                                              // :148-155
            if (emit) { if (n_ilen < out->var_cap) out->ilen[n_ilen] = j - last_indel_pos; else status = PF_BUBBLE_OUT_OVERFLOW; }
            n_ilen++;
            open = false;
        }
        if (number && emit) {
            if (n_var < out->var_cap) {
                out->var_col[n_var] = j;
                out->var_kind[n_var] = (uint8_t)kind;
                uint16_t *cl = out->cls + (uint64_t)n_var * nr;
                uint16_t next = 0;                                      // ids in order of first appearance (:75-92, :123-144)
                for (uint32_t r = 0; r < nr; r++) {
                    const uint8_t c = cand[(uint64_t)r * stride + j];
                    bool seen = false;
                    for (uint32_t q = 0; q < r; q++)
                        if (cand[(uint64_t)q * stride + j] == c) { cl[r] = cl[q]; seen = true; break; }
                    if (!seen) cl[r] = ++next;
                }
            } else status = PF_BUBBLE_OUT_OVERFLOW;
        }
        if (number) n_var++;
      }
    }
    key.n_snp = (int)n_snp;
    key.n_indel = (int)n_indel;
    key.d_indel = ind.spacing(Llast);
    key.d_snp = snp.spacing(Llast);
    key.d_all = all.spacing(Llast);
    key.site_l = all.n ? (int)all.first : -1;
    key.site_r = all.n ? (int)all.last : -1;
    if (n_var_out) *n_var_out = n_var;
    if (n_ilen_out) *n_ilen_out = n_ilen;
    return status;
}

// strcmp(a, b) > 0 for two rows of possibly different length (no NUL inside rows).
PF_HD bool row_greater(const CBV a, uint32_t la, const CBV b, uint32_t lb) {
    const uint32_t n = la < lb ? la : lb;
    for (uint32_t i = 0; i < n; i++)
        if (a[i] != b[i]) return a[i] > b[i];
    return la > lb;
}

// The 7-level preference of compareStrPair (:158-233) between a candidate and the incumbent state.
struct Incumbent {
    int snp_dis, indel_dis, snp_count, indel_count, all_dis, site_l, site_r;
    int index;  // candidate index of the incumbent, -1 = none
    PF_HD void init() {
        snp_dis = INT_MAX; indel_dis = INT_MAX; snp_count = INT_MAX / 2; indel_count = INT_MAX / 2;
        all_dis = INT_MAX; site_l = -1; site_r = -1; index = -1;
    }
};

// returns 1 = candidate replaces incumbent, 0 = keep, 2 = tie down to the rows (caller does the strcmp scan)
PF_HD int prefer(const MsaKey &k, const Incumbent &b) {
    const int total = k.n_snp + k.n_indel, btotal = b.snp_count + b.indel_count;
    if (total < btotal) return 1;
    if (total != btotal) return 0;
    if (k.n_indel < b.indel_count) return 1;
    if (k.n_indel != b.indel_count) return 0;
    if (k.d_indel > (uint64_t)(long long)b.indel_dis) return 1;   // size_t vs int compare (:167)
    if (k.d_indel != (uint64_t)(long long)b.indel_dis) return 0;
    if (k.d_snp > (uint64_t)(long long)b.snp_dis) return 1;
    if (k.d_snp != (uint64_t)(long long)b.snp_dis) return 0;
    if (k.d_all > (uint64_t)(long long)b.all_dis) return 1;
    if (k.d_all != (uint64_t)(long long)b.all_dis) return 0;
    const int l = k.site_l < 0 ? 0 : k.site_l, r = k.site_r < 0 ? 0 : k.site_r;  // temp_vec[0] on an empty vector is UB in the reference
    if (l > b.site_l || r > b.site_r) return 1;
    if (l == b.site_l && r == b.site_r) return 2;
    return 0;
}

PF_HD void adopt(Incumbent &b, const MsaKey &k, int index) {      // :216-233 (and :196-207, same end state)
    b.all_dis = (int)k.d_all;
    b.site_l = b.site_l > k.site_l ? b.site_l : k.site_l;
    b.site_r = b.site_r > k.site_r ? b.site_r : k.site_r;
    b.snp_count = k.n_snp;
    b.indel_count = k.n_indel;
    b.snp_dis = (int)k.d_snp;
    b.indel_dis = (int)k.d_indel;
    b.index = index;
}

// ---- per-warp work area ---------------------------------------------------------------------------------

struct Limits {          // uniform for one launch
    uint32_t max_rows;   // sequences per bubble
    uint32_t max_alen;   // columns of any intermediate / final MSA, also max rows of a DP matrix
    uint32_t max_blen;   // longest raw sequence (DP matrix columns)
    uint32_t k_cand;     // candidate MSAs carried between rounds (<= 64)
    uint32_t k_aln;      // co-optimal pairwise alignments kept by one traceback (<= 64)
    uint32_t max_var;    // variable columns per output slot
    uint64_t step_limit; // traceback iterations per bubble (summed over its pairwise alignments)
    uint32_t diag_flags; // layout of the flag bytes: LAYOUT_ROW (lane / group kernels, host), LAYOUT_DIAG (generic warp kernel), LAYOUT_SKEW (CTA kernel)
    uint32_t pad_;       // LAYOUT_SKEW: T = lanes of the CTA-wide fill
};

PF_HD uint64_t flag_area_cells(const Limits &l) {
    if (l.diag_flags == LAYOUT_SKEW) {
        const uint64_t T = l.pad_ ? l.pad_ : 1, w = (l.max_blen + T) / T;      // ceil((max_blen + 1) / T)
        return (uint64_t)(l.max_alen + T + 1) * w * T;
    }
    return l.diag_flags ? (uint64_t)(l.max_alen + l.max_blen + 1) * (l.max_alen + 1) : (uint64_t)(l.max_alen + 1) * (l.max_blen + 1);
}


struct WorkArea {
    BV flags;            // flag_area_cells() bytes, see flag_index()
    BV mv;               // max_alen + max_blen
    BV ext_mv;           // k_aln * (max_alen + max_blen)
    WV ext_len;          // k_aln
    BV cand[2];          // k_cand * max_rows * max_alen each
    WV cand_len[2];      // k_cand each
    int32_t *brow;       // 2 * (max_blen + 1)   (generic kernel only: carried row of the 32-row blocks)
};

PF_HD uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

// Bytes of ONE bubble's work area (`lanes` = 1) or of a lane-interleaved group (`lanes` = 32).
PF_HD uint64_t work_area_bytes(const Limits &l, uint32_t lanes = 1) {
    uint64_t b = 0;
    b += align_up(flag_area_cells(l), 16);
    b += align_up(l.max_alen + l.max_blen, 16);
    b += align_up((uint64_t)l.k_aln * (l.max_alen + l.max_blen), 16);
    b += align_up((uint64_t)l.k_aln * 4, 16);
    b += 2 * align_up((uint64_t)l.k_cand * l.max_rows * l.max_alen, 16);
    b += 2 * align_up((uint64_t)l.k_cand * 4, 16);
    b *= lanes;
    if (lanes == 1) b += align_up((uint64_t)2 * (l.max_blen + 1) * 4, 16);
    return b;
}

// lanes == 1: contiguous private area.  lanes == 32: element t of lane L's array sits at array_base + t*32 + L.
// contig (lanes > 1): only the flag bytes are interleaved (the fill's lock-step stores coalesce); the arrays of the sequential
// phases -- move strings, kept alignments, candidate MSAs -- are contiguous per lane, so a lane streaming through its own array
// stays inside one 32-byte sector for 32 elements (L1) instead of touching a new sector per element.
PF_HD WorkArea carve_work_area(uint8_t *base, const Limits &l, uint32_t lanes = 1, uint32_t lane = 0, bool contig = false) {
    WorkArea w;
    uint8_t *p = base;
    const uint32_t st = contig ? 1u : lanes;
    uint64_t sz;
    w.flags = bv(p + lane, lanes); p += lanes * align_up(flag_area_cells(l), 16);
    sz = align_up(l.max_alen + l.max_blen, 16);
    w.mv = bv(p + (contig ? lane * sz : lane), st); p += lanes * sz;
    sz = align_up((uint64_t)l.k_aln * (l.max_alen + l.max_blen), 16);
    w.ext_mv = bv(p + (contig ? lane * sz : lane), st); p += lanes * sz;
    sz = align_up((uint64_t)l.k_aln * 4, 16);
    w.ext_len = wv((uint32_t *)p + (contig ? lane * (sz / 4) : lane), st); p += lanes * sz;
    sz = align_up((uint64_t)l.k_cand * l.max_rows * l.max_alen, 16);
    for (int i = 0; i < 2; i++) { w.cand[i] = bv(p + (contig ? lane * sz : lane), st); p += lanes * sz; }
    sz = align_up((uint64_t)l.k_cand * 4, 16);
    for (int i = 0; i < 2; i++) { w.cand_len[i] = wv((uint32_t *)p + (contig ? lane * (sz / 4) : lane), st); p += lanes * sz; }
    w.brow = (int32_t *)p;
    return w;
}

// ---- output slot ------------------------------------------------------------------------------------------

struct SlotHdr {
    int32_t status;
    uint32_t n_rows, alen, n_var, n_ilen;
    uint32_t pad[3];
};

struct SlotLayout {
    uint32_t alen_cap, var_cap;
    uint64_t off_rows, off_varcol, off_ilen, off_cls, off_kind, bytes;
};

PF_HD SlotLayout slot_layout(uint32_t n_seq, uint64_t sum_len, const Limits &l) {
    SlotLayout s;
    s.alen_cap = (uint32_t)(sum_len < l.max_alen ? sum_len : l.max_alen);
    s.var_cap = s.alen_cap < l.max_var ? s.alen_cap : l.max_var;
    s.off_rows = sizeof(SlotHdr);
    s.off_varcol = align_up(s.off_rows + (uint64_t)n_seq * s.alen_cap, 4);
    s.off_ilen = s.off_varcol + 4ull * s.var_cap;
    s.off_cls = s.off_ilen + 4ull * s.var_cap;
    s.off_kind = s.off_cls + 2ull * s.var_cap * n_seq;
    s.bytes = align_up(s.off_kind + s.var_cap, 16);
    return s;
}

// The helpers a policy offers to the sequential phases.  Policies whose sequential phases run on one thread inherit these;
// the group kernel overrides them with versions that spread the loop over the lanes of the group.
struct SerialHelpers {
    PF_HD bool prefetch_flags() const { return false; }
    PF_HD uint32_t prefetch_cells() const { return 6; }
    PF_HD uint32_t skew_T() const { return 1; }
    PF_HD uint32_t skew_w(uint32_t) const { return 1; }
    PF_HD void prefetch_ahead(const BV, uint32_t, uint32_t, uint32_t, uint32_t) const {}
    PF_HD PairKey analyze(const Scoring &sc, const CBV row, const CBV B, const CBV mv, uint32_t depth) const { return analyze_moves(sc, row, B, mv, depth); }
    PF_HD void copy(const CBV src, const BV dst, uint32_t n) const { copy_bytes(src, dst, n); }
    PF_HD void project(const CBV src, const CBV mv, uint32_t depth, const BV dst, uint8_t gap_move) const { project_moves(src, mv, depth, dst, gap_move); }
};

// ---- the per-bubble driver: SequenceAlignment (SeqAlign.cpp:550-640) --------------------------------------
//
// X is the execution policy:
//   * generic kernel: a warp (leader = lane 0, fill = wavefront by shuffle, bcast = shuffle), contiguous work area;
//   * lane kernel: ONE THREAD per bubble (every lane is its own leader, fill = row-by-row with the score row
//     in shared memory), lane-interleaved work area;
//   * group kernel: G lanes; the fill is cooperative, and everything else runs REDUNDANTLY on all G lanes (leader() is true
//     everywhere: same loads, same decisions, same stores to the same addresses -- no extra time under SIMT), which lets
//     the O(L) helpers (analyze / copy / project) use the G lanes instead of one;
//   * tests/hostemu: a single CPU thread.
// X::kLayout selects the flag-byte layout the policy's fill writes.
template <class X>
PF_HDN inline void msa_run(X &x, const uint8_t *bases, const uint64_t *seq_off, uint32_t s0, uint32_t ns,
                           const WorkArea &ws, const Limits &lim, const Scoring &sc, uint8_t *slot) {
    const uint32_t mv_stride = lim.max_alen + lim.max_blen;
    const uint64_t cand_stride = (uint64_t)lim.max_rows * lim.max_alen;
    const uint64_t sum_len = seq_off[s0 + ns] - seq_off[s0];
    const SlotLayout lay = slot_layout(ns, sum_len, lim);
    SlotHdr *hdr = (SlotHdr *)slot;
    int status = PF_BUBBLE_OK;
    if (ns < 2) status = PF_BUBBLE_BAD_INPUT;
    else if (ns > lim.max_rows) status = PF_BUBBLE_TOO_MANY_ROWS;
    uint32_t ncand = 0;   // leader-owned; other lanes learn it through bcast
    uint64_t steps_left = lim.step_limit;   // traceback budget of the whole bubble (all its pairwise alignments)
    int cur = 0;
    for (uint32_t i = 1; i < ns && status == PF_BUBBLE_OK; i++) {
        const CBV B = cbv(bases + seq_off[s0 + i]);
        const uint32_t n = (uint32_t)(seq_off[s0 + i + 1] - seq_off[s0 + i]);
        if (n > lim.max_blen) { status = PF_BUBBLE_TOO_LONG; break; }
        const uint32_t nk = (i == 1) ? 1u : x.bcast(ncand);
        uint32_t nnext = 0;
        int best_total = INT_MIN;
        for (uint32_t k = 0; k < nk && status == PF_BUBBLE_OK; k++) {
            CBV A;
            uint32_t m;
            if (i == 1) { A = cbv(bases + seq_off[s0]); m = (uint32_t)(seq_off[s0 + 1] - seq_off[s0]); }
            else { A = cbv(ws.cand[cur] + k * cand_stride); m = x.bcast_ld(&ws.cand_len[cur][k]); }
            if (m > lim.max_alen) { status = PF_BUBBLE_TOO_LONG; break; }
            x.fill(ws.flags, A, m, B, n, sc, ws.brow);
            if (x.leader()) {
                const TbResult tb = traceback<X::kLayout>(x, ws.flags, A, m, B, n, sc, ws.mv, ws.ext_mv, ws.ext_len, mv_stride,
                                                             lim.k_aln, steps_left, x.pitch_n(n));
                steps_left -= tb.steps < steps_left ? tb.steps : steps_left;
                x.note_steps(tb.steps);
                status = tb.status;
                if (status == PF_BUBBLE_OK) {
                    uint64_t alive = tb.n_aln >= 64 ? ~0ull : ((1ull << tb.n_aln) - 1);
                    int total_k = 0;
                    for (uint32_t j = 1; j < i; j++) {                  // :575-618
                        PairKey inc;
                        inc.score = INT_MIN; inc.n_pos = 0; inc.n_indel = 0;
                        int best_j = INT_MIN;
                        uint64_t alive_j = 0;
                        const CBV rowj = A + (uint64_t)j * lim.max_alen;
                        for (uint32_t v = 0; v < tb.n_aln; v++) {
                            if (!((alive >> v) & 1)) continue;
                            const PairKey pk = x.analyze(sc, rowj, B, cbv(ws.ext_mv + (uint64_t)v * mv_stride), ws.ext_len[v]);
                            const int d = rank_diff(pk.score, pk.n_pos, pk.n_indel, inc.score, inc.n_pos, inc.n_indel);
                            if (d > 0) { inc = pk; best_j = (int)inc.score; alive_j = 1ull << v; }
                            else if (d == 0) { best_j = (int)inc.score; alive_j |= 1ull << v; }
                        }
                        alive = alive_j;
                        total_k = (int)((unsigned)total_k + (unsigned)best_j);  // the reference's int += (:617) wraps
                    }
                    if (total_k > best_total) { best_total = total_k; nnext = 0; }   // :619-636
                    if (total_k >= best_total) {
                        for (uint32_t v = 0; v < tb.n_aln && status == PF_BUBBLE_OK; v++) {
                            if (!((alive >> v) & 1)) continue;
                            const uint32_t depth = ws.ext_len[v];
                            if (nnext == lim.k_cand) { status = PF_BUBBLE_CAND_OVERFLOW; break; }
                            if (depth > lim.max_alen) { status = PF_BUBBLE_TOO_LONG; break; }
                            const CBV mvv = cbv(ws.ext_mv + (uint64_t)v * mv_stride);
                            const BV dst = ws.cand[cur ^ 1] + nnext * cand_stride;
                            for (uint32_t r = 0; r < i; r++)
                                x.project(A + (uint64_t)r * lim.max_alen, mvv, depth, dst + (uint64_t)r * lim.max_alen, (uint8_t)MV_L);
                            x.project(B, mvv, depth, dst + (uint64_t)i * lim.max_alen, (uint8_t)MV_U);
                            ws.cand_len[cur ^ 1][nnext] = depth;
                            nnext++;
                        }
                    }
                }
            }
            status = x.bcast_i(status);
        }
        if (x.leader()) ncand = nnext;
        cur ^= 1;
        x.sync();
    }
    // ---- compareStrPair over the surviving candidates (leader) ----
    if (x.leader()) {
        hdr->n_rows = 0; hdr->alen = 0; hdr->n_var = 0; hdr->n_ilen = 0;
        hdr->pad[0] = hdr->pad[1] = hdr->pad[2] = 0;
        if (status == PF_BUBBLE_OK && ncand > 0) {
            const CBV cbase = cbv(ws.cand[cur]);
            const WV clen = ws.cand_len[cur];
            const uint64_t Llast = clen[ncand - 1];
            Incumbent inc;
            inc.init();
            if (ncand == 1) inc.index = 0;   // a single candidate always replaces the empty incumbent (:158-233): no need to rank it
            for (uint32_t c = 0; c < ncand && ncand > 1; c++) {
                MsaKey key;
                scan_candidate(cbase + c * cand_stride, lim.max_alen, ns, clen[c], Llast, key, false, nullptr, nullptr, nullptr);
                int p = prefer(key, inc);
                if (p == 2) {   // :190-211: replace as soon as any row compares greater than the incumbent's
                    p = 0;
                    const CBV ib = cbase + (uint64_t)inc.index * cand_stride;
                    for (uint32_t r = 0; r < ns; r++)
                        if (row_greater(cbase + (c * cand_stride + (uint64_t)r * lim.max_alen), clen[c],
                                        ib + (uint64_t)r * lim.max_alen, clen[inc.index])) { p = 1; break; }
                }
                if (p == 1) adopt(inc, key, (int)c);
            }
            if (inc.index >= 0) {
                const uint32_t L = clen[inc.index];
                const CBV win = cbase + (uint64_t)inc.index * cand_stride;
                if (L > lay.alen_cap) status = PF_BUBBLE_TOO_LONG;
                else {
                    SlotView sv;
                    sv.rows = slot + lay.off_rows;
                    sv.var_col = (uint32_t *)(slot + lay.off_varcol);
                    sv.ilen = (uint32_t *)(slot + lay.off_ilen);
                    sv.cls = (uint16_t *)(slot + lay.off_cls);
                    sv.var_kind = slot + lay.off_kind;
                    sv.var_cap = lay.var_cap;
                    MsaKey key;
                    uint32_t nv = 0, nil = 0;
                    status = scan_candidate(win, lim.max_alen, ns, L, Llast, key, true, &sv, &nv, &nil);
                    if (status == PF_BUBBLE_OK) {
                        for (uint32_t r = 0; r < ns; r++) x.copy(win + (uint64_t)r * lim.max_alen, bv(sv.rows + (uint64_t)r * L), L);
                        hdr->n_rows = ns; hdr->alen = L; hdr->n_var = nv; hdr->n_ilen = nil;
                    }
                }
            }
        }
        hdr->status = status;
    }
    x.sync();
}

}  // namespace pfalign
