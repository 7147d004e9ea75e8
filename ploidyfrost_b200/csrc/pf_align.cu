// pf_align.cu -- batched SeqAlign::SequenceAlignment (src/SeqAlign.cpp:550) on sm_100a.
//
// Two kernels run the same per-bubble state machines (pf_align_core.cuh) under two execution policies:
//
//   * msa_lane_kernel -- ONE THREAD PER BUBBLE, the path almost every superbubble takes (branches of 2k-1 ..
//     a few hundred bases).  Bubbles are sorted on the device by (size class, #branches, longest branch) so
//     that the 32 bubbles of a warp are alike and the lanes stay in lock-step.  Each lane fills its own DP
//     matrix row by row (needlemanWunch, SeqAlign.cpp:480-547): the previous score row and the B
//     characters live in shared memory, lane-interleaved ([j*32+lane], bank = lane, conflict free); flag
//     bytes go to a lane-interleaved work area in HBM/L2, so the 32 lanes of a step write 32 consecutive
//     bytes (one sector).  Traceback DFS, progressive-MSA filter and site calling then run on every lane
//     at once instead of on one leader lane of a warp.  INT32 ALU only; the FP64 add+truncate variant is a
//     separate template instantiation used only when -M/-D/-G are not whole numbers.
//   * msa_warp_kernel -- one WARP per bubble for everything the lane kernel cannot hold (branches longer
//     than 256 bases, more than 8 branches) and for its overflow cases: anti-diagonal wavefront inside the
//     warp (lane = matrix row in 32-row blocks, neighbours exchanged by __shfl_up_sync), leader-lane
//     traceback.  Work areas are sized from the batch, so 5 kbp branches fit.
//
// Results land in per-bubble slots; a size pass + CUB exclusive scans + a warp-per-bubble gather compact
// them into the flat pf_msa_batch_t arrays.  A bubble that overflows its tier (too many co-optimal
// alignments, long insertions, traceback step budget) is re-run with the large limits; anything that still
// does not fit is reported per bubble in status[] -- never silently altered.
#include "pf_common.cuh"
#include "pf_align_core.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <type_traits>
#include <cstring>

using namespace pfalign;

namespace {

#ifndef PF_LANE_MINB
#define PF_LANE_MINB 8     // minimum resident CTAs per SM the compiler must allow for (register cap = 65536 / (64 * MINB))
#endif
#ifndef PF_GROUP_MINB
#define PF_GROUP_MINB 8
#endif
constexpr int WARP_BLOCK = 128;  // msa_warp_kernel: 4 warps per CTA
constexpr int LANE_BLOCK = 64;   // msa_lane_kernel: 2 warps per CTA (shared memory is per warp, small CTAs pack SMs tighter)
constexpr uint32_t FULL = 0xffffffffu;

// size classes of the lane kernel: longest branch of the bubble <= LANE_NMAX[c]
constexpr int N_LANE_CLASSES = 5;
constexpr int CLS_BIG = N_LANE_CLASSES;        // longest branch 257 .. BIG_SPLIT: one warp per bubble
constexpr int CLS_HUGE = N_LANE_CLASSES + 1;   // longer branches: one CTA per bubble (msa_cta_kernel)
constexpr int CLS_RETRY_SMEM = N_LANE_CLASSES + 2;  // re-run, warp kernel with the flag bytes in shared memory (long DFS searches)
constexpr int CLS_RETRY = N_LANE_CLASSES + 3;       // re-run with the large limits
constexpr int N_TIERS = N_LANE_CLASSES + 4;
constexpr uint32_t BIG_SPLIT = 1536;           // up to here every score of the default scoring fits the s16x2 fill of a warp
// traceback iterations a lane / group may spend on one bubble before it is handed to the heavy queue (bench workload: median 250,
// p99.9 1 500, p99.99 2 000, a handful per 262 144 bubbles at 3 000 .. 32 000; sweep in profiles/r01_summary.md section 7)
constexpr uint32_t LANE_STEP_LIMIT = 2048;
__host__ __device__ constexpr uint32_t lane_nmax(int c) { return c == 0 ? 64u : c == 1 ? 96u : c == 2 ? 128u : c == 3 ? 192u : 256u; }
constexpr uint32_t LANE_MAX_ROWS = 8;

// ---- warp-cooperative fill (generic kernel) ----------------------------------------------------------------
template <bool INTEGRAL>
__device__ __forceinline__ void warp_fill(uint8_t *__restrict__ flags, const CBV A, const uint32_t m, const uint8_t *B,
                                          const uint32_t n, const Scoring &sc, int32_t *brow, const uint32_t lane) {
    __syncwarp();
    const uint32_t W = m + 1;
    for (uint32_t i = lane; i <= m; i += 32) flags[i * W + i] = i ? (uint8_t)F_UP : (uint8_t)0;   // column 0 (:486-491)
    for (uint32_t j = 1 + lane; j <= n; j += 32) flags[j * W] = (uint8_t)F_LEFT;                  // row 0 (:492-496)
    int32_t *rd = brow, *wr = brow + (n + 1);
    for (uint32_t rb = 0; rb * 32 < m; rb++) {
        const uint32_t i = rb * 32 + lane + 1;
        const bool active = i <= m;
        const uint8_t a = active ? A[i - 1] : (uint8_t)0;
        const bool block_left = active && i != m && A[i] == '-';
        int cur = pack_sf(border_score(sc, i), F_UP);                       // cell (i,0)
        int diag = pack_sf(border_score(sc, rb * 32), rb ? F_UP : 0);     // lane 0 only: cell (rb*32, 0)
        const uint32_t rows_here = min(32u, m - rb * 32);
        const uint32_t nsteps = n + rows_here - 1;
        uint32_t b = 0, bchunk = 0;
        int rdchunk = 0;
        for (uint32_t s = 0; s < nsteps; s++) {
            if ((s & 31) == 0) {
                bchunk = (s + lane < n) ? (uint32_t)B[s + lane] : 0u;       // B[s .. s+31], one coalesced load per 32 steps
                if (rb) rdchunk = (s + 1 + lane <= n) ? rd[s + 1 + lane] : 0;
            }
            const uint32_t b0 = __shfl_sync(FULL, bchunk, s & 31);          // B[s]
            const uint32_t bu = __shfl_up_sync(FULL, b, 1);
            b = lane == 0 ? b0 : bu;                                        // lane L holds B[s-L] = B[j-1]
            int up = __shfl_up_sync(FULL, cur, 1);                          // cell (i-1, j) from the lane above
            const int up0 = rb ? __shfl_sync(FULL, rdchunk, s & 31) : pack_sf(border_score(sc, s + 1), F_LEFT);
            if (lane == 0) up = up0;                                        // row rb*32: carried row / top border
            const int j = (int)s - (int)lane + 1;
            if (active && j >= 1 && j <= (int)n) {
                cur = nw_cell_t<INTEGRAL>(sc, up, diag, cur, a, (uint8_t)b, block_left);
                flags[(i + (uint32_t)j) * W + i] = (uint8_t)unpack_f(cur);
                if (lane == 31) wr[j] = cur;
            }
            diag = up;
        }
        if (lane == 31 && active) wr[0] = 0;  // unused: column 0 of the carried row is rebuilt from border_score
        __syncwarp();
        int32_t *t = rd; rd = wr; wr = t;
    }
    __syncwarp();
}

template <bool INTEGRAL>
struct WarpExec : SerialHelpers {
    static constexpr int kLayout = LAYOUT_DIAG;
    uint32_t lane;
    unsigned long long cells;
    __device__ __forceinline__ bool leader() const { return lane == 0; }
    __device__ __forceinline__ uint32_t bcast(uint32_t v) const { return __shfl_sync(FULL, v, 0); }
    __device__ __forceinline__ int bcast_i(int v) const { return __shfl_sync(FULL, v, 0); }
    __device__ __forceinline__ uint32_t bcast_ld(const uint32_t *p) const { return *(const volatile uint32_t *)p; }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    __device__ __forceinline__ void note_steps(uint64_t) const {}
    __device__ __forceinline__ uint32_t pitch_n(uint32_t n) const { return n; }
    __device__ __forceinline__ void fill(const BV flags, const CBV A, uint32_t m, const CBV B, uint32_t n, const Scoring &sc,
                                         int32_t *brow) {
        warp_fill<INTEGRAL>(flags.p, A, m, B.p, n, sc, brow, lane);
        cells += (unsigned long long)m * n;
    }
};

// ---- per-lane fill (lane kernel) ---------------------------------------------------------------------------
// rowbuf / bs point at THIS lane's element 0 of the warp's lane-interleaved shared arrays (element j at [j*32]).
// VARIANT 0: FP64 add+truncate scoring (-M/-D/-G not whole numbers), 1: INT32, 2: two DP rows per step as s16x2.
constexpr int LANE_FP64 = 0, LANE_I32 = 1, LANE_S16X2 = 2;
constexpr int S16_BLOCK = -16000;   // "Left is blocked" addend of the s16x2 fill: below every reachable score, no int16 wrap

__device__ __forceinline__ uint32_t pack2(int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }

// One step of the two-row fill.  Halves: lo = cell (i, s) of row i, hi = cell (i+1, s-1) of row i+1.
// Carried per half: U = S + [Up flag], D = S + [Diag flag], L = S + [Left flag] of the previous cell of the row
// (what the cell below / diagonally below / to the right adds its own term to, SeqAlign.cpp:512-526).
// `best + flag` is max(best, t + 1) because t <= best: one VIADDMNMX per carried value, no compare/select.
struct S16Step {
    uint32_t curU, curD, curL, upD_prev, b2;
    template <bool LO, bool HI>
    __device__ __forceinline__ void step(uint32_t w, uint32_t b, uint32_t a2, uint32_t M2, uint32_t NE2, uint32_t G2,
                                         uint32_t Grow2, uint8_t *f_lo, uint8_t *f_hi, uint32_t *row_out) {
        const uint32_t ONE2 = 0x00010001u;
        const uint32_t upU = __byte_perm(w, curU, 0x5410);          // lo: U(i-1,s) from the row buffer, hi: U(i,s-1)
        const uint32_t upD = __byte_perm(w, curD, 0x5432);          // lo: D(i-1,s),                     hi: D(i,s-1)
        b2 = __byte_perm(b, b2, 0x5410);                            // lo: B[s-1] (column s), hi: previous column's
        const uint32_t mask = __vminu2(a2 ^ b2, ONE2) * 0xFFFFu;    // per half: 0xFFFF where the characters differ
        const uint32_t sub2 = (mask & NE2) | (~mask & M2);          // :498-506 (B holds no '-', checked by the caller)
        const uint32_t t_up = __vadd2(upU, G2);
        const uint32_t t_dg = __vadd2(upD_prev, sub2);
        const uint32_t t_lf = __vadd2(curL, Grow2);
        const uint32_t best = __vimax3_s16x2(t_up, t_dg, t_lf);
        const uint32_t nU = __viaddmax_s16x2(t_up, ONE2, best);
        const uint32_t nD = __viaddmax_s16x2(t_dg, ONE2, best);
        const uint32_t nL = __viaddmax_s16x2(t_lf, ONE2, best);
        const uint32_t f2 = ((nU ^ best) & ONE2) + (((nD ^ best) & ONE2) << 1) + (((nL ^ best) & ONE2) << 2);
        if (LO) *f_lo = (uint8_t)f2;
        if (HI) { *f_hi = (uint8_t)(f2 >> 16); *row_out = __byte_perm(nU, nD, 0x7632); }
        upD_prev = upD;
        curU = nU; curD = nD; curL = nL;
    }
};

template <int VARIANT>
struct LaneExec : SerialHelpers {
    static constexpr int kLayout = LAYOUT_ROW;
    static constexpr bool INTEGRAL = VARIANT != LANE_FP64;
    int32_t *rowbuf;     // 4 bytes per column: packed score*8+flags (scalar fills) or (U, D) as two int16 (s16x2 fill)
    uint8_t *bs;
    uint32_t pitch;      // flag row pitch shared by the 32 bubbles of the warp (0: every bubble its own n + 1)
    unsigned long long cells;
    __device__ __forceinline__ bool leader() const { return true; }
    __device__ __forceinline__ uint32_t bcast(uint32_t v) const { return v; }
    __device__ __forceinline__ int bcast_i(int v) const { return v; }
    __device__ __forceinline__ uint32_t bcast_ld(const uint32_t *p) const { return *p; }
    __device__ __forceinline__ void sync() const {}
    __device__ __forceinline__ void note_steps(uint64_t) const {}
    __device__ __forceinline__ uint32_t pitch_n(uint32_t n) const { return pitch ? pitch - 1 : n; }
    __device__ __forceinline__ bool prefetch_flags() const { return true; }   // the flag bytes live in HBM / L2

    // rows i = 1..m one at a time
    __device__ __forceinline__ void fill_scalar(const BV flags, const CBV A, uint32_t m, uint32_t n, const Scoring &sc) {
        flags[0] = 0;
        rowbuf[0] = pack_sf(0, 0);
        for (uint32_t j = 1; j <= n; j++) {                                  // row 0 (:492-496)
            rowbuf[j * 32] = pack_sf(border_score(sc, j), F_LEFT);
            flags[j] = (uint8_t)F_LEFT;
        }
        const uint32_t W = pitch ? pitch : n + 1;
        uint8_t a_next = m ? A[0] : (uint8_t)0;
        for (uint32_t i = 1; i <= m; i++) {
            const uint8_t a = a_next;
            a_next = i < m ? A[i] : (uint8_t)0;
            const bool block_left = i != m && a_next == '-';
            int dg = rowbuf[0];
            const int c0 = pack_sf(border_score(sc, i), F_UP);               // column 0 (:486-491)
            rowbuf[0] = c0;
            const BV frow = flags + (uint64_t)i * W;
            frow[0] = (uint8_t)F_UP;
            if (INTEGRAL) {
                // Same cell as nw_cell_t (SeqAlign.cpp:512-545) with the loop-carried part cut to two operations:
                // lf_t = score(i,j-1) + GAP + [Left flag of (i,j-1)] is carried ready-made, and a row whose Left move
                // is blocked by the profile rule (:528-532) carries -2^28 instead, which can never win or tie.
                const int G = sc.iG, Grow = block_left ? -(1 << 28) : sc.iG;
                const int sub_ne = a == '-' ? sc.iG : sc.iD;
                int lf_t = unpack_s(c0) + Grow;                              // (i,0) carries only Up
#pragma unroll 4
                for (uint32_t j = 1; j <= n; j++) {
                    const int up = rowbuf[j * 32];
                    const uint8_t b = bs[(j - 1) * 32];
                    const int t_up = (up >> 3) + (up & 1) + G;
                    const int sub = a == b ? sc.iM : (b == '-' ? sc.iG : sub_ne);
                    const int t_dg = (dg >> 3) + ((dg >> 1) & 1) + sub;
                    const int t = max(t_up, t_dg);
                    const int best = max(t, lf_t);
                    const int fl = lf_t >= t ? 1 : 0;
                    const int f = (t_up == best ? F_UP : 0) | (t_dg == best ? F_DIAG : 0) | (fl ? F_LEFT : 0);
                    lf_t = best + Grow + fl;
                    rowbuf[j * 32] = best * 8 + f;
                    frow[j] = (uint8_t)f;
                    dg = up;
                }
            } else {
                int lf = c0;
                for (uint32_t j = 1; j <= n; j++) {
                    const int up = rowbuf[j * 32];
                    const int cur = nw_cell_t<false>(sc, up, dg, lf, a, bs[(j - 1) * 32], block_left);
                    rowbuf[j * 32] = cur;
                    frow[j] = (uint8_t)unpack_f(cur);
                    dg = up;
                    lf = cur;
                }
            }
        }
    }

    // rows (i, i+1) together, row i+1 one column behind: s16x2 arithmetic (VIMNMX3.S16x2 / VIADDMNMX.S16x2 / VIADD.16x2).
    // Needs n >= 1, integral scoring, every |score| + 1 < 16000 and no '-' in B (the launcher / caller check).
    __device__ __forceinline__ void fill_s16x2(const BV flags, const CBV A, uint32_t m, uint32_t n, const Scoring &sc) {
        uint32_t *rb = (uint32_t *)rowbuf;
        const uint32_t W = pitch ? pitch : n + 1;
        const int G = sc.iG;
        flags[0] = 0;
        rb[0] = pack2(0, 0);                                                 // (0,0): no flags
        for (uint32_t j = 1; j <= n; j++) {                                  // row 0 carries only Left (:492-496)
            rb[j * 32] = pack2(G * (int)j, G * (int)j);
            flags[j] = (uint8_t)F_LEFT;
        }
        const uint32_t M2 = pack2(sc.iM, sc.iM), G2 = pack2(G, G);
        for (uint32_t i = 1; i <= m; i += 2) {
            const bool two = i < m;                                          // is row i+1 real?
            const uint8_t a_lo = A[i - 1], a_hi = two ? A[i] : (uint8_t)0;
            const uint8_t a_after = i + 1 < m ? A[i + 1] : (uint8_t)0;
            const bool blk_lo = two && a_hi == '-';                          // i != m && A[i] == '-'      (:528)
            const bool blk_hi = i + 1 < m && a_after == '-';
            const uint32_t a2 = (uint32_t)a_lo | ((uint32_t)a_hi << 16);
            const uint32_t NE2 = pack2(a_lo == '-' ? G : sc.iD, a_hi == '-' ? G : sc.iD);
            const uint32_t Grow2 = pack2(blk_lo ? S16_BLOCK : G, blk_hi ? S16_BLOCK : G);
            const int b_lo = G * (int)i, b_hi = G * (int)(i + 1);            // border scores (:489)
            S16Step st;
            st.curU = pack2(b_lo + 1, 0); st.curD = pack2(b_lo, 0); st.curL = pack2(b_lo, 0);   // cell (i,0): Up only
            st.upD_prev = rb[0] >> 16;                                       // D(i-1, 0)
            st.b2 = 0;
            uint8_t *f0 = &flags[(uint64_t)i * W], *f1 = &flags[(uint64_t)(i + 1) * W];         // stride-32 rows
            f0[0] = (uint8_t)F_UP;
            uint32_t dummy;
            // step 1: only row i has a cell
            st.step<true, false>(rb[32], bs[0], a2, M2, NE2, G2, Grow2, f0 + 32, nullptr, &dummy);
            if (two) {
                f1[0] = (uint8_t)F_UP;
                rb[0] = pack2(b_hi + 1, b_hi);                               // (i+1,0) for the next pair
                st.curU = (st.curU & 0xFFFFu) | ((uint32_t)(b_hi + 1) << 16);
                st.curD = (st.curD & 0xFFFFu) | ((uint32_t)b_hi << 16);
                st.curL = (st.curL & 0xFFFFu) | ((uint32_t)b_hi << 16);
                // the row-buffer word and the B character of step s+1 are fetched before step s stores (the compiler cannot
                // hoist them itself: the byte loads may alias the row-buffer stores)
                uint32_t w_next = n >= 2 ? rb[2 * 32] : 0u, b_next = n >= 2 ? (uint32_t)bs[32] : 0u;
#pragma unroll 2
                for (uint32_t s = 2; s <= n; s++) {
                    const uint32_t w = w_next, b = b_next;
                    if (s < n) { w_next = rb[(s + 1) * 32]; b_next = bs[s * 32]; }
                    st.step<true, true>(w, b, a2, M2, NE2, G2, Grow2, f0 + (uint64_t)s * 32, f1 + (uint64_t)(s - 1) * 32, rb + (s - 1) * 32);
                }
                // step n+1: only row i+1 has a cell
                st.step<false, true>(0, 0, a2, M2, NE2, G2, Grow2, nullptr, f1 + (uint64_t)n * 32, rb + n * 32);
            } else {
                for (uint32_t s = 2; s <= n; s++)
                    st.step<true, false>(rb[s * 32], bs[(s - 1) * 32], a2, M2, NE2, G2, Grow2, f0 + (uint64_t)s * 32, nullptr, &dummy);
            }
        }
    }

    __device__ __forceinline__ void fill(const BV flags, const CBV A, uint32_t m, const CBV B, uint32_t n, const Scoring &sc,
                                         int32_t *) {
        bool dash = false;
        for (uint32_t j = 0; j < n; j++) {
            const uint8_t b = B[j];
            bs[j * 32] = b;
            dash |= b == '-';
        }
        if (VARIANT == LANE_S16X2 && !dash && n >= 1) fill_s16x2(flags, A, m, n, sc);
        else fill_scalar(flags, A, m, n, sc);
        cells += (unsigned long long)m * n;
    }
};

// ---- G lanes per bubble (group kernel) -------------------------------------------------------------------------
// The thread-per-bubble kernel is throughput-efficient but one long bubble is a long serial job (a bubble with 250-base
// branches is ~3 M dependent instructions on its lane), and a batch holds too few long bubbles to hide that behind other
// warps.  For the long size classes a bubble therefore gets G = 2..32 adjacent lanes.  The fill splits every DP row into G
// column strips; lane g fills strip g of row pair (T - g) at time step T (a software pipeline across the lanes, the same
// two-rows-per-step s16x2 arithmetic as the lane kernel), and hands the last column of its strip -- (U, D, L) of both
// rows -- to lane g+1 with one shuffle per value per time step.  Each lane owns its strip's slice of the bubble's
// shared-memory score row, so the fill needs no barrier at all.  Everything after the fill (traceback DFS, MSA filter,
// site calling) runs on the group's first lane, exactly as in the warp kernel.
//
// Layouts: the NB = 32 / G bubbles of a warp interleave their work-area arrays element by element (stride NB), flag rows
// are row-major with the same pitch for every bubble of the launch (lanes of different bubbles at the same cell share a
// sector), the score row of bubble b sits at rb[j * NB + b] and strips are an odd number of columns wide, which makes
// the G x NB lanes of a step hit 32 different banks.
template <int G>
struct GroupExec {
    static constexpr int kLayout = LAYOUT_ROW;
    static constexpr uint32_t NB = 32 / G;
    uint32_t g;          // lane within the group
    uint32_t gmask;      // the group's lanes
    uint32_t *rb;        // this bubble's score row: element j at rb[j * NB]
    uint8_t *bs;         // this bubble's B characters: element j at bs[j * NB]
    uint32_t pn;         // flag row pitch - 1
    bool pf;             // flag bytes in HBM (prefetch ahead of the DFS) or in shared memory (heavy kernel)
    unsigned long long cells;
    __device__ __forceinline__ bool prefetch_flags() const { return pf; }
    __device__ __forceinline__ uint32_t prefetch_cells() const { return 6; }
    __device__ __forceinline__ uint32_t skew_T() const { return 1; }
    __device__ __forceinline__ uint32_t skew_w(uint32_t) const { return 1; }
    __device__ __forceinline__ void prefetch_ahead(const BV, uint32_t, uint32_t, uint32_t, uint32_t) const {}
    // Everything outside the fill runs redundantly on all G lanes (identical loads, decisions and stores), so every lane
    // "is" the leader; the kernel uses g == 0 where something must happen once (queue, counters).
    __device__ __forceinline__ bool leader() const { return true; }
    __device__ __forceinline__ uint32_t bcast(uint32_t v) const { return __shfl_sync(gmask, v, 0, G); }
    __device__ __forceinline__ int bcast_i(int v) const { return __shfl_sync(gmask, v, 0, G); }
    __device__ __forceinline__ uint32_t bcast_ld(const uint32_t *p) const { return *(const volatile uint32_t *)p; }
    __device__ __forceinline__ void sync() const { __syncwarp(gmask); }
    __device__ __forceinline__ void note_steps(uint64_t) const {}
    __device__ __forceinline__ uint32_t pitch_n(uint32_t) const { return pn; }

    // ---- O(L) helpers spread over the G lanes ----
    __device__ __forceinline__ uint32_t group_bits(uint32_t ballot) const {   // the group's G bits of a ballot, lane g at bit g
        return G == 32 ? ballot : ((ballot >> (__ffs(gmask) - 1)) & ((1u << (G & 31)) - 1u));
    }
    // variantAnalyze over a move string (see analyze_moves): column p of the alignment is move mv[depth-1-p]; the characters it
    // consumes sit at the number of earlier non-Left / non-Up moves (ballot + popc), its score term is local, and whether it
    // opens a gap run depends on the class of the previous column only -- so G columns are handled per step.
    __device__ __forceinline__ PairKey analyze(const Scoring &sc, const CBV row, const CBV B, const CBV mv, uint32_t depth) const {
        if (!sc.integral) return analyze_moves(sc, row, B, mv, depth);       // the double accumulation is order dependent
        const uint32_t ltm = (1u << g) - 1u;
        int score = 0;
        uint32_t n_pos = 0, n_indel = 0, ia = 0, jb = 0, carry = 0;
        for (uint32_t p0 = 0; p0 < depth; p0 += G) {
            const uint32_t p = p0 + g;
            const bool live = p < depth;
            const uint8_t m = live ? (uint8_t)(mv[depth - 1 - p] & 3) : (uint8_t)MV_NONE;
            const bool use_a = live && m != MV_L, use_b = live && m != MV_U;
            const uint32_t ma = group_bits(__ballot_sync(gmask, use_a)), mb = group_bits(__ballot_sync(gmask, use_b));
            const uint8_t a = use_a ? row[ia + __popc(ma & ltm)] : (uint8_t)'-';
            const uint8_t b = use_b ? B[jb + __popc(mb & ltm)] : (uint8_t)'-';
            ia += __popc(ma); jb += __popc(mb);
            const uint32_t c = (!live || a == b) ? 0u : (a == '-' ? 1u : (b == '-' ? 2u : 3u));
            uint32_t prevc = __shfl_up_sync(gmask, c, 1, G);
            if (g == 0) prevc = carry;
            carry = __shfl_sync(gmask, c, G - 1, G);
            int t_score = !live ? 0 : ((a == '-' || b == '-') ? sc.iG : (a == b ? sc.iM : sc.iD));      // :241-246, gap first
            uint32_t t_ind = ((c == 1u && prevc != 1u) || (c == 2u && prevc != 2u)) ? 1u : 0u;
            uint32_t t_pos = (t_ind || c == 3u) ? 1u : 0u;
#pragma unroll
            for (int d = G / 2; d; d >>= 1) {
                t_score += __shfl_xor_sync(gmask, t_score, d, G);
                t_ind += __shfl_xor_sync(gmask, t_ind, d, G);
                t_pos += __shfl_xor_sync(gmask, t_pos, d, G);
            }
            score += t_score; n_indel += t_ind; n_pos += t_pos;
        }
        PairKey k;
        k.score = score; k.n_pos = n_pos; k.n_indel = n_indel;
        return k;
    }
    __device__ __forceinline__ void copy(const CBV src, const BV dst, uint32_t n) const {
        for (uint32_t i = g; i < n; i += G) dst[i] = src[i];
        __syncwarp(gmask);
    }
    // project_moves: dst[p] = the p-th column of `src` stretched by the gaps of the move string
    __device__ __forceinline__ void project(const CBV src, const CBV mv, uint32_t depth, const BV dst, uint8_t gap_move) const {
        const uint32_t ltm = (1u << g) - 1u;
        uint32_t ia = 0;
        for (uint32_t p0 = 0; p0 < depth; p0 += G) {
            const uint32_t p = p0 + g;
            const bool live = p < depth;
            const bool use = live && (mv[depth - 1 - p] & 3) != gap_move;
            const uint32_t mu = group_bits(__ballot_sync(gmask, use));
            if (live) dst[p] = use ? src[ia + __popc(mu & ltm)] : (uint8_t)'-';
            ia += __popc(mu);
        }
        __syncwarp(gmask);
    }

    // the one-row-at-a-time fill (B contains '-': the s16x2 substitution select does not apply)
    __device__ __forceinline__ void fill_scalar(const BV flags, const CBV A, uint32_t m, const CBV B, uint32_t n, const Scoring &sc) {
        int32_t *row = (int32_t *)rb;
        const uint32_t W = pn + 1;
        flags[0] = 0;
        row[0] = pack_sf(0, 0);
        for (uint32_t j = 1; j <= n; j++) {
            row[j * NB] = pack_sf(border_score(sc, j), F_LEFT);
            flags[j] = (uint8_t)F_LEFT;
        }
        for (uint32_t i = 1; i <= m; i++) {
            const bool block_left = i != m && A[i] == '-';
            const uint8_t a = A[i - 1];
            int dg = row[0];
            int lf = pack_sf(border_score(sc, i), F_UP);
            row[0] = lf;
            const BV frow = flags + (uint64_t)i * W;
            frow[0] = (uint8_t)F_UP;
            for (uint32_t j = 1; j <= n; j++) {
                const int up = row[j * NB];
                const int cur = nw_cell_t<true>(sc, up, dg, lf, a, B[j - 1], block_left);
                row[j * NB] = cur;
                frow[j] = (uint8_t)unpack_f(cur);
                dg = up;
                lf = cur;
            }
        }
    }

    __device__ __forceinline__ void fill(const BV flags, const CBV A, uint32_t m, const CBV B, uint32_t n, const Scoring &sc,
                                         int32_t *) {
        __syncwarp(gmask);                                                   // the leader is done reading the previous matrix
        if (g == 0) cells += (unsigned long long)m * n;
        const uint32_t w = ((n + G - 1) / G) | 1u;                           // strip width, odd
        const uint32_t a = g * w + 1;                                        // my strip: columns a .. b (empty when a > n)
        const uint32_t b = min(a + w - 1, n);
        const bool have = a <= n;
        bool dash = false;
        if (have)
            for (uint32_t j0 = a; j0 <= b; j0 += PF_CH) {                    // my B characters, chunked loads
                uint8_t v[PF_CH];
#pragma unroll
                for (uint32_t q = 0; q < PF_CH; q++) v[q] = j0 + q <= b ? B[j0 + q - 1] : (uint8_t)0;
#pragma unroll
                for (uint32_t q = 0; q < PF_CH; q++) if (j0 + q <= b) { bs[(j0 + q) * NB] = v[q]; dash |= v[q] == '-'; }
            }
        if (__any_sync(gmask, dash) || n == 0) {
            if (g == 0) fill_scalar(flags, A, m, B, n, sc);
            __syncwarp(gmask);
            return;
        }
        const uint32_t W = pn + 1;
        const int Gp = sc.iG;
        const uint32_t M2 = pack2(sc.iM, sc.iM), G2 = pack2(Gp, Gp);
        if (g == 0) flags[0] = 0;
        if (have)
            for (uint32_t j = a; j <= b; j++) {                              // row 0 carries only Left (:492-496)
                rb[j * NB] = pack2(Gp * (int)j, Gp * (int)j);
                flags[j] = (uint8_t)F_LEFT;
            }
        const uint32_t R = (m + 1) / 2;                                      // row pairs
        uint32_t inU = 0, inD = 0, inL = 0;                                  // column a-1 of the pair I fill next (from lane g-1)
        uint32_t prev_hi_d = (uint32_t)(Gp * (int)(a - 1)) & 0xFFFFu;        // D(i-1, a-1); row 0: its score
        for (uint32_t T = 0; T < R + G - 1; T++) {
            uint32_t outU = 0, outD = 0, outL = 0;
            const uint32_t r = T - g;
            if (have && T >= g && r < R) {
                const uint32_t i = 2 * r + 1;
                const bool two = i < m;                                      // is row i+1 real?
                const uint8_t a_lo = A[i - 1], a_hi = two ? A[i] : (uint8_t)0;
                const uint8_t a_after = i + 1 < m ? A[i + 1] : (uint8_t)0;
                const bool blk_lo = two && a_hi == '-';                      // i != m && A[i] == '-'      (:528)
                const bool blk_hi = i + 1 < m && a_after == '-';
                const uint32_t a2 = (uint32_t)a_lo | ((uint32_t)a_hi << 16);
                const uint32_t NE2 = pack2(a_lo == '-' ? Gp : sc.iD, a_hi == '-' ? Gp : sc.iD);
                const uint32_t Grow2 = pack2(blk_lo ? S16_BLOCK : Gp, blk_hi ? S16_BLOCK : Gp);
                uint32_t cU, cD, cL;                                         // the column left of my strip
                if (g == 0) {                                                // the border (:486-491): Up only
                    const int b_lo = Gp * (int)i, b_hi = Gp * (int)(i + 1);
                    cU = pack2(b_lo + 1, b_hi + 1); cD = pack2(b_lo, b_hi); cL = pack2(b_lo, b_hi);
                } else { cU = inU; cD = inD; cL = inL; }
                S16Step st;
                st.curU = cU & 0xFFFFu; st.curD = cD & 0xFFFFu; st.curL = cL & 0xFFFFu;
                st.upD_prev = prev_hi_d;
                st.b2 = 0;
                prev_hi_d = cD >> 16;                                        // D(i+1, a-1): the next pair's diagonal
                uint8_t *f0 = &flags[(uint64_t)i * W], *f1 = &flags[(uint64_t)(i + 1) * W];
                if (g == 0) f0[0] = (uint8_t)F_UP;
                uint32_t dummy;
                // step a: only row i has a cell in my strip
                st.step<true, false>(rb[a * NB], bs[a * NB], a2, M2, NE2, G2, Grow2, f0 + (uint64_t)a * NB, nullptr, &dummy);
                uint32_t loU, loD, loL;                                      // (i, b) once row i is through
                if (two) {
                    if (g == 0) f1[0] = (uint8_t)F_UP;
                    st.curU = (st.curU & 0xFFFFu) | (cU & 0xFFFF0000u);
                    st.curD = (st.curD & 0xFFFFu) | (cD & 0xFFFF0000u);
                    st.curL = (st.curL & 0xFFFFu) | (cL & 0xFFFF0000u);
                    uint32_t w_next = b > a ? rb[(a + 1) * NB] : 0u, b_next = b > a ? (uint32_t)bs[(a + 1) * NB] : 0u;
#pragma unroll 2
                    for (uint32_t s = a + 1; s <= b; s++) {
                        const uint32_t wv_ = w_next, bc = b_next;
                        if (s < b) { w_next = rb[(s + 1) * NB]; b_next = bs[(s + 1) * NB]; }
                        st.step<true, true>(wv_, bc, a2, M2, NE2, G2, Grow2, f0 + (uint64_t)s * NB, f1 + (uint64_t)(s - 1) * NB, rb + (s - 1) * NB);
                    }
                    loU = st.curU; loD = st.curD; loL = st.curL;
                    // step b+1: only row i+1 has a cell
                    st.step<false, true>(0, 0, a2, M2, NE2, G2, Grow2, nullptr, f1 + (uint64_t)b * NB, rb + b * NB);
                } else {
                    for (uint32_t s = a + 1; s <= b; s++)
                        st.step<true, false>(rb[s * NB], bs[s * NB], a2, M2, NE2, G2, Grow2, f0 + (uint64_t)s * NB, nullptr, &dummy);
                    loU = st.curU; loD = st.curD; loL = st.curL;
                }
                outU = __byte_perm(loU, st.curU, 0x7610);
                outD = __byte_perm(loD, st.curD, 0x7610);
                outL = __byte_perm(loL, st.curL, 0x7610);
            }
            // my last column becomes the right neighbour's left column for the same row pair, one time step later
            inU = __shfl_up_sync(gmask, outU, 1, G);
            inD = __shfl_up_sync(gmask, outD, 1, G);
            inL = __shfl_up_sync(gmask, outL, 1, G);
        }
        __syncwarp(gmask);                                                   // flags visible to the leader's traceback
    }
};

struct MsaArgs {
    const uint8_t *bases;
    const uint64_t *seq_off;
    const uint32_t *bubble_off;
    const uint32_t *order;     // work item -> bubble id
    uint32_t first;            // first work item of this launch
    uint32_t n_items;
    uint8_t *slot_base;
    const uint64_t *slot_off;  // per work item
    uint64_t *slot_ptr;        // per bubble: address of its slot
    uint8_t *tier;             // per bubble: tier whose Limits laid out its slot
    uint8_t tier_id;
    uint8_t *ws_base;
    uint64_t ws_stride;        // per warp
    uint32_t *counter;
    unsigned long long *stat_cells;  // DP cells filled (m*n per needlemanWunch call), for the roofline figure
    uint32_t *hq;              // heavy queue (see msa_heavy_kernel); nullptr = none
    uint32_t lane_pitch;       // lane kernel: 1 = one flag row pitch per launch (lanes at the same cell share a sector)
    uint32_t warp_dequeue;     // group kernel: 1 = the warp's groups start their bubbles together, 0 = every group on its own
    uint32_t lane_contig;      // lane / group kernels: per-lane contiguous arrays for the sequential phases (carve_work_area)
    Limits lim;
    Scoring sc;
};


// ---- T lanes per bubble: the CTA-wide fill for long branches -----------------------------------------------------------------
// Branches of kilobases (BASELINE configs[4]: up to 5 kbp, 25 M cells per pair) are too long for a warp: one pair is a
// serial job of tens of milliseconds, and beyond ~1.8 kbp the scores leave int16, so the s16x2 fill does not apply.  Here one
// CTA of T = 256 lanes works on one bubble.  The DP row is cut into T column strips of w = ceil((n + 1) / T) columns; lane l fills
// its strip of row t + 1 - l at time step t (an anti-diagonal wavefront over (row, strip)), INT32, with the carried values of the
// s16x2 design (U = S + [Up], D = S + [Diag], L = S + [Left], so a cell is three adds, a max3 and three add-max).  Bands in shared
// memory: the (U, D) pairs of the latest row of every column ([q * T + l]: conflict-free), the B characters, and a double-buffered
// hand-over row through which lane l - 1 passes (L, D) of its last column to lane l; one __syncthreads per time step.  Flag bytes
// go to HBM in the skewed layout (LAYOUT_SKEW): the lanes of a warp store 32 consecutive bytes.  Everything sequential
// (traceback DFS, MSA filter, site calling) runs on warp 0 with the 32-lane cooperative helpers of the group kernel, while the
// DFS's flag reads are covered by a 32-cell prefetch fan up the diagonal (one lane per cell).
constexpr int CTA_THREADS = 256;
struct CtaExec {
    static constexpr int kLayout = LAYOUT_SKEW;
    static constexpr uint32_t T = CTA_THREADS;
    uint32_t tid, lane, warp;
    int2 *rb;            // [w_max * T]
    int2 *hand;          // [2 * T]
    uint8_t *bs;         // [w_max * T]
    uint32_t *bc;        // broadcast word
    GroupExec<32> h;     // warp 0's helpers
    unsigned long long cells;
    __device__ __forceinline__ bool leader() const { return warp == 0; }
    __device__ __forceinline__ uint32_t bcast(uint32_t v) const { __syncthreads(); if (tid == 0) *bc = v; __syncthreads(); return *(volatile uint32_t *)bc; }
    __device__ __forceinline__ int bcast_i(int v) const { return (int)bcast((uint32_t)v); }
    __device__ __forceinline__ uint32_t bcast_ld(const uint32_t *p) const { __syncthreads(); return *(const volatile uint32_t *)p; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ void note_steps(uint64_t) const {}
    __device__ __forceinline__ uint32_t pitch_n(uint32_t n) const { return n; }
    __device__ __forceinline__ bool prefetch_flags() const { return true; }
    __device__ __forceinline__ uint32_t prefetch_cells() const { return 6; }
    __device__ __forceinline__ uint32_t skew_T() const { return T; }
    __device__ __forceinline__ uint32_t skew_w(uint32_t n) const { return (n + T) / T; }
    // lane g asks for the cell g + 1 steps up the diagonal of the DFS's position (the path of a good alignment)
    __device__ __forceinline__ void prefetch_ahead(const BV flags, uint32_t cell, uint32_t q, uint32_t w, uint32_t) const {
        const uint32_t k = lane + 1;
        uint32_t back = k * w * T;                       // k rows up
        if (q >= k) back += k * T;                       // k columns left inside the strip
        else {
            const uint32_t nb = (k - q + w - 1) / w;     // strips crossed
            back += nb * w * T + nb;                     // one row of the skew and one lane per strip
            back -= (nb * w - k) * T;                    // q' - q = nb * w - k columns to the right inside the new strip
        }
        if (back <= cell) prefetch_byte(flags.p + (cell - back));
    }
    __device__ __forceinline__ PairKey analyze(const Scoring &sc, const CBV row, const CBV B, const CBV mv, uint32_t depth) const { return h.analyze(sc, row, B, mv, depth); }
    __device__ __forceinline__ void copy(const CBV src, const BV dst, uint32_t n) const { h.copy(src, dst, n); }
    __device__ __forceinline__ void project(const CBV src, const CBV mv, uint32_t depth, const BV dst, uint8_t gap_move) const { h.project(src, mv, depth, dst, gap_move); }

    __device__ __forceinline__ void fill(const BV flags, const CBV A, uint32_t m, const CBV B, uint32_t n, const Scoring &sc, int32_t *) {
        __syncthreads();                                                     // warp 0 is done with the previous matrix
        const uint32_t w = (n + T) / T;                                      // strip width = ceil((n + 1) / T)
        const uint32_t l = tid, j0 = l * w;                                  // my columns j0 .. j1 (column 0 = the border)
        const uint32_t j1 = min(j0 + w - 1, n);
        const bool have = j0 <= n;
        const uint32_t nl = n / w + 1;                                       // lanes that own a column
        if (tid == 0) cells += (unsigned long long)m * n;
        uint8_t *F = flags.p;
        const int G = sc.iG, Mv = sc.iM, Dv = sc.iD;
        if (have)
            for (uint32_t j = j0; j <= j1; j++) {                            // row 0 carries only Left (:492-496)
                const uint32_t q = j - j0;
                if (j) bs[q * T + l] = B[j - 1];
                rb[q * T + l] = make_int2(G * (int)j, G * (int)j);
                F[((uint64_t)l * w + q) * T + l] = j ? (uint8_t)F_LEFT : (uint8_t)0;
            }
        int prevD = G * ((int)j0 - 1);                                       // D(0, j0 - 1); unused by lane 0
        __syncthreads();
        const uint32_t steps = m + nl - 1;
        for (uint32_t t = 0; t < steps; t++) {
            const uint32_t i = t + 1 - l;                                    // my row at this time step
            if (have && t + 1 > l && i <= m) {
                const uint8_t a = A[i - 1];
                const bool blk = i != m && A[i] == '-';                      // the profile rule (:528-532)
                const int sub_ne = a == '-' ? G : Dv, Grow = blk ? -(1 << 28) : G;
                uint8_t *Fr = F + ((uint64_t)(i + l) * w) * T + l;
                int curL, curD, dgD;
                uint32_t q = 0;
                if (l == 0) {                                                // the border column (:486-491): Up only
                    curL = curD = G * (int)i;
                    dgD = G * ((int)i - 1);
                    Fr[0] = (uint8_t)F_UP;
                    q = 1;
                } else {
                    const int2 hv = hand[((t - 1) & 1) * T + l - 1];         // (L, D) of (i, j0 - 1)
                    curL = hv.x; curD = hv.y;
                    dgD = prevD;                                             // D(i - 1, j0 - 1)
                    prevD = hv.y;
                }
                const uint32_t qn = j1 - j0 + 1;
#pragma unroll 2
                for (; q < qn; q++) {
                    const int2 up = rb[q * T + l];
                    const uint8_t b = bs[q * T + l];
                    const int t_up = up.x + G, t_dg = dgD + (a == b ? Mv : sub_ne), t_lf = curL + Grow;
                    const int best = __vimax3_s32(t_up, t_dg, t_lf);
                    const int nU = __viaddmax_s32(t_up, 1, best), nD = __viaddmax_s32(t_dg, 1, best), nL = __viaddmax_s32(t_lf, 1, best);
                    Fr[(uint64_t)q * T] = (uint8_t)((nU - best) + 2 * (nD - best) + 4 * (nL - best));
                    rb[q * T + l] = make_int2(nU, nD);
                    dgD = up.y; curL = nL; curD = nD;
                }
                hand[(t & 1) * T + l] = make_int2(curL, curD);
            }
            __syncthreads();
        }
    }
};

// ---- bubbles whose co-optimal DFS explodes ------------------------------------------------------------------------------
// One bubble in ~10^5 has a traceback of 10^4 .. 10^6 steps (the DFS enumerates co-optimal paths under the reference's
// pruning rules, it cannot be cut short).  Such a search is a serial job of milliseconds, so the only way to keep it off the
// step's critical path is to start it early and run it beside everything else: the first-pass kernels give a bubble a small
// step budget; a bubble that exceeds it is pushed on a device-side queue, and msa_heavy_kernel -- a handful of one-warp
// CTAs with the flag matrix AND the move string in shared memory, launched before the first pass and polling that queue --
// re-runs it concurrently.  When the first pass is over the heavy CTAs stop taking tickets; whatever is still queued (a batch
// full of explosive bubbles) is re-run by the full-width pass after it.
// Queue layout (u32 words): [0] tail (producers), [1] head (consumer tickets), [2] done flag, [4..5] bytes of the slot pool
// handed out, [HQ_HDR ...] bubble ids, 0xFFFFFFFF = not written yet.
constexpr uint32_t HQ_HDR = 8, HQ_CAP = 4096;
constexpr uint32_t HQ_TRACE = HQ_HDR + HQ_CAP;   // u32 index of the trace area: per ticket {bubble, t_taken, t_done} as u64 (PF_HEAVY_TRACE)

__device__ __forceinline__ void heavy_enqueue(uint32_t *hq, const uint8_t *slot, uint32_t bubble) {
    const int st = ((const SlotHdr *)slot)->status;
    if (st != PF_BUBBLE_STEP_LIMIT && st != PF_BUBBLE_CAND_OVERFLOW) return;
    __threadfence();                                   // slot_ptr / tier of the first pass are written before the hand-over
    const uint32_t idx = atomicAdd(hq, 1u);
    if (idx < HQ_CAP) {
        ((volatile uint32_t *)hq)[HQ_HDR + idx] = bubble;
        __threadfence();
    }
}

template <bool INTEGRAL>
__global__ void __launch_bounds__(WARP_BLOCK) msa_warp_kernel(const MsaArgs a) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    WorkArea ws = carve_work_area(a.ws_base + (uint64_t)warp * a.ws_stride, a.lim);
    WarpExec<INTEGRAL> x;
    x.lane = lane;
    x.cells = 0;
    for (;;) {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(a.counter, 1u);
        w = __shfl_sync(FULL, w, 0);
        if (w >= a.n_items) break;
        w += a.first;
        const uint32_t b = a.order[w];
        const uint32_t s0 = a.bubble_off[b], ns = a.bubble_off[b + 1] - s0;
        uint8_t *slot = a.slot_base + a.slot_off[w];
        msa_run(x, a.bases, a.seq_off, s0, ns, ws, a.lim, a.sc, slot);
        if (lane == 0) { a.slot_ptr[b] = (uint64_t)(uintptr_t)slot; a.tier[b] = a.tier_id; }
    }
    if (lane == 0 && x.cells) atomicAdd(a.stat_cells, x.cells);
}

template <bool I>
__device__ __forceinline__ void heavy_exec_init(WarpExec<I> &x, uint32_t lane, uint8_t *, uint32_t) { x.lane = lane; x.cells = 0; }
__device__ __forceinline__ void heavy_exec_init(GroupExec<32> &x, uint32_t lane, uint8_t *row_smem, uint32_t nmax) {
    x.g = lane; x.gmask = FULL; x.pn = nmax; x.pf = false; x.cells = 0;
    x.rb = (uint32_t *)row_smem;
    x.bs = row_smem + 4 * (nmax + 2);
}

// The consumer side of the heavy queue: one warp per CTA; a.lim = the shared-memory tier's limits, a.slot_base = the slot
// pool (bump-allocated through hq[4..5]), a.first = pool capacity in KB.
// MODE 2: the strip-pipelined s16x2 fill of the group kernel with all 32 lanes on the one bubble (row-major flags);
// MODE 1 / 0: the INT32 / FP64 wavefront of the warp kernel (diagonal-major flags) when the scores do not fit int16.
// Dynamic shared memory: flag matrix | move string | (MODE 2) score row + B characters.
template <int MODE>
__global__ void __launch_bounds__(32) msa_heavy_kernel(const MsaArgs a) {
    extern __shared__ __align__(16) uint8_t smem_flags[];
    const uint32_t lane = threadIdx.x & 31;
    WorkArea ws = carve_work_area(a.ws_base + (uint64_t)blockIdx.x * a.ws_stride, a.lim);
    const uint64_t flag_bytes = align_up(flag_area_cells(a.lim), 16), mv_bytes = align_up(a.lim.max_alen + a.lim.max_blen, 16);
    ws.flags = bv(smem_flags);
    ws.mv = bv(smem_flags + flag_bytes);                                    // the DFS's move string next to the flags
    typename std::conditional<MODE == 2, GroupExec<32>, WarpExec<MODE == 1>>::type x;
    heavy_exec_init(x, lane, smem_flags + flag_bytes + mv_bytes, a.lim.max_blen);
    uint32_t *hq = a.hq;
    const uint64_t pool_cap = (uint64_t)a.first << 10;
    if (!hq) {   // drain mode (the pass after the first one): work items [0, n_items) of a.order, slots laid out by the host
        for (;;) {
            uint32_t w = 0;
            if (lane == 0) w = atomicAdd(a.counter, 1u);
            w = __shfl_sync(FULL, w, 0);
            if (w >= a.n_items) break;
            const uint32_t b = a.order[w];
            const uint32_t s0 = a.bubble_off[b], ns = a.bubble_off[b + 1] - s0;
            uint8_t *slot = a.slot_base + a.slot_off[w];
            msa_run(x, a.bases, a.seq_off, s0, ns, ws, a.lim, a.sc, slot);
            if (lane == 0) { a.slot_ptr[b] = (uint64_t)(uintptr_t)slot; a.tier[b] = a.tier_id; }
            __syncwarp();
        }
        if (lane == 0 && x.cells) atomicAdd(a.stat_cells, x.cells);
        return;
    }
    for (;;) {
        uint32_t b = 0xFFFFFFFFu;
        if (lane == 0 && !*(volatile uint32_t *)(hq + 2)) {                  // the first pass is still running: take a ticket
            const uint32_t t = atomicAdd(hq + 1, 1u);
            if (t < HQ_CAP) {
                volatile uint32_t *e = hq + HQ_HDR + t;
                unsigned long long t0;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                for (;;) {
                    b = *e;
                    if (b != 0xFFFFFFFFu) break;
                    if (*(volatile uint32_t *)(hq + 2)) { b = *e; break; }   // producers are gone: a last look
                    __nanosleep(1000);
                    unsigned long long t1;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                    // nothing for 30 ms: give up (a profiler or sanitizer that serialises kernels would otherwise wait forever
                    // for a first pass that cannot start); whatever is queued later is re-run by the pass after the first one
                    if (t1 - t0 > 30000000ull) break;
                }
            }
        }
        b = __shfl_sync(FULL, b, 0);
        if (b == 0xFFFFFFFFu) break;
        __threadfence();
        const uint32_t s0 = a.bubble_off[b], ns = a.bubble_off[b + 1] - s0;
        const uint64_t bytes = slot_layout(ns, a.seq_off[s0 + ns] - a.seq_off[s0], a.lim).bytes;
        unsigned long long off = 0;
        if (lane == 0) off = atomicAdd((unsigned long long *)(hq + 4), (unsigned long long)bytes);
        off = __shfl_sync(FULL, off, 0);
        if (off + bytes > pool_cap) continue;                                // left to the pass after the first one
        uint8_t *slot = a.slot_base + off;
        unsigned long long tt0, tt1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt0));
        msa_run(x, a.bases, a.seq_off, s0, ns, ws, a.lim, a.sc, slot);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt1));
        if (lane == 0) {
            a.slot_ptr[b] = (uint64_t)(uintptr_t)slot; a.tier[b] = a.tier_id;
            const uint32_t tr = atomicAdd(hq + 3, 1u);
            if (tr < HQ_CAP) {
                unsigned long long *T = (unsigned long long *)(hq + HQ_TRACE) + 3ull * tr;
                T[0] = b; T[1] = tt0; T[2] = tt1;
            }
        }
        __syncwarp();
    }
    if (lane == 0 && x.cells) atomicAdd(a.stat_cells, x.cells);
}

__global__ void heavy_done_kernel(uint32_t *hq) { *(volatile uint32_t *)(hq + 2) = 1u; }

__host__ __device__ constexpr uint32_t lane_smem_per_warp(uint32_t nmax, uint32_t tsize) {
    return 32 * tsize * (nmax + 1) + 32 * nmax;                               // score row + B characters; multiple of 32
}

template <int VARIANT>
__global__ void __launch_bounds__(LANE_BLOCK, PF_LANE_MINB) msa_lane_kernel(const MsaArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nmax = a.lim.max_blen;
    const uint32_t per_warp = lane_smem_per_warp(nmax, 4);
    LaneExec<VARIANT> x;
    x.rowbuf = (int32_t *)(smem + (size_t)wib * per_warp) + lane;
    x.bs = smem + (size_t)wib * per_warp + 32 * 4 * (nmax + 1) + lane;
    x.pitch = a.lane_pitch ? nmax + 1 : 0;
    x.cells = 0;
    const WorkArea ws = carve_work_area(a.ws_base + (uint64_t)warp * a.ws_stride, a.lim, 32, lane, a.lane_contig != 0);
    for (;;) {
        uint32_t g = 0;
        if (lane == 0) g = atomicAdd(a.counter, 1u);
        g = __shfl_sync(FULL, g, 0);
        const uint32_t n_groups = (a.n_items + 31) / 32;
        if (g >= n_groups) break;
        const uint32_t wi = (n_groups - 1 - g) * 32 + lane;                  // sorted ascending by size: biggest groups first
        if (wi < a.n_items) {
            const uint32_t w = a.first + wi;
            const uint32_t b = a.order[w];
            const uint32_t s0 = a.bubble_off[b], ns = a.bubble_off[b + 1] - s0;
            uint8_t *slot = a.slot_base + a.slot_off[w];
            msa_run(x, a.bases, a.seq_off, s0, ns, ws, a.lim, a.sc, slot);
            a.slot_ptr[b] = (uint64_t)(uintptr_t)slot;
            a.tier[b] = a.tier_id;
            if (a.hq) heavy_enqueue(a.hq, slot, b);
        }
        __syncwarp();
    }
    unsigned long long c = x.cells;
#pragma unroll
    for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(FULL, c, d);
    if (lane == 0 && c) atomicAdd(a.stat_cells, c);
}


__host__ __device__ constexpr uint32_t cta_smem_bytes(uint32_t nmax) {
    return (((nmax + CTA_THREADS) / CTA_THREADS) * CTA_THREADS) * 9u + 2u * CTA_THREADS * 8u + 32u;   // rb + bs, hand, broadcast word
}

// One CTA per bubble (CtaExec); work items are taken biggest first.
__global__ void __launch_bounds__(CTA_THREADS) msa_cta_kernel(const MsaArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t wmax = (a.lim.max_blen + CTA_THREADS) / CTA_THREADS;
    CtaExec x;
    x.tid = threadIdx.x; x.lane = threadIdx.x & 31; x.warp = threadIdx.x >> 5;
    x.rb = (int2 *)smem;
    x.hand = x.rb + (size_t)wmax * CTA_THREADS;
    x.bs = (uint8_t *)(x.hand + 2 * CTA_THREADS);
    x.bc = (uint32_t *)(x.bs + (size_t)wmax * CTA_THREADS);
    x.h.g = x.lane; x.h.gmask = FULL; x.h.rb = nullptr; x.h.bs = nullptr; x.h.pn = 0; x.h.pf = true; x.h.cells = 0;
    x.cells = 0;
    const WorkArea ws = carve_work_area(a.ws_base + (uint64_t)blockIdx.x * a.ws_stride, a.lim);
    for (;;) {
        uint32_t q = 0;
        if (threadIdx.x == 0) q = atomicAdd(a.counter, 1u);
        q = x.bcast(q);
        if (q >= a.n_items) break;
        const uint32_t w = a.first + (a.n_items - 1 - q);
        const uint32_t b = a.order[w];
        const uint32_t s0 = a.bubble_off[b], ns = a.bubble_off[b + 1] - s0;
        uint8_t *slot = a.slot_base + a.slot_off[w];
        msa_run(x, a.bases, a.seq_off, s0, ns, ws, a.lim, a.sc, slot);
        if (threadIdx.x == 0) { a.slot_ptr[b] = (uint64_t)(uintptr_t)slot; a.tier[b] = a.tier_id; }
    }
    if (threadIdx.x == 0 && x.cells) atomicAdd(a.stat_cells, x.cells);
}

constexpr int GROUP_BLOCK = 64;
__host__ __device__ constexpr uint32_t group_smem_per_warp(uint32_t nmax, uint32_t nb) {
    return (nb * 4 * (nmax + 2) + nb * (nmax + 2) + 15) / 16 * 16;           // score rows + B characters (index 1..nmax)
}

template <int G>
__global__ void __launch_bounds__(GROUP_BLOCK, PF_GROUP_MINB) msa_group_kernel(const MsaArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr uint32_t NB = 32 / G;
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nmax = a.lim.max_blen;
    const uint32_t per_warp = group_smem_per_warp(nmax, NB);
    GroupExec<G> x;
    x.g = lane % G;
    const uint32_t bi = lane / G;
    x.gmask = G == 32 ? FULL : (((1u << G) - 1u) << (bi * G));
    x.rb = (uint32_t *)(smem + (size_t)wib * per_warp) + bi;
    x.bs = smem + (size_t)wib * per_warp + NB * 4 * (nmax + 2) + bi;
    x.pn = nmax;
    x.pf = true;
    x.cells = 0;
    const WorkArea ws = carve_work_area(a.ws_base + (uint64_t)warp * a.ws_stride, a.lim, NB, bi, a.lane_contig != 0);
    if (a.warp_dequeue) {
        // The warp pulls NB neighbouring bubbles of the size-sorted list at a time and its groups start them together: alike
        // bubbles stay in the same phase (fill with fill, traceback with traceback), i.e. little SIMT divergence between
        // the groups of a warp; a group that finishes early waits for the others.
        for (;;) {
            uint32_t q = 0;
            if (lane == 0) q = atomicAdd(a.counter, NB);
            q = __shfl_sync(FULL, q, 0);
            if (q >= a.n_items) break;
            const uint32_t item = q + bi;
            if (item < a.n_items) {
                const uint32_t w = a.first + (a.n_items - 1 - item);
                const uint32_t b = a.order[w];
                const uint32_t s0 = a.bubble_off[b], ns = a.bubble_off[b + 1] - s0;
                uint8_t *slot = a.slot_base + a.slot_off[w];
                msa_run(x, a.bases, a.seq_off, s0, ns, ws, a.lim, a.sc, slot);
                if (x.g == 0) {
                    a.slot_ptr[b] = (uint64_t)(uintptr_t)slot;
                    a.tier[b] = a.tier_id;
                    if (a.hq) heavy_enqueue(a.hq, slot, b);
                }
            }
            __syncwarp();
        }
    } else {
        for (;;) {   // every group pulls its own bubbles: biggest first
            uint32_t q = 0;
            if (x.g == 0) q = atomicAdd(a.counter, 1u);
            q = __shfl_sync(x.gmask, q, 0, G);
            if (q >= a.n_items) break;
            const uint32_t w = a.first + (a.n_items - 1 - q);
            const uint32_t b = a.order[w];
            const uint32_t s0 = a.bubble_off[b], ns = a.bubble_off[b + 1] - s0;
            uint8_t *slot = a.slot_base + a.slot_off[w];
            msa_run(x, a.bases, a.seq_off, s0, ns, ws, a.lim, a.sc, slot);
            if (x.g == 0) {
                a.slot_ptr[b] = (uint64_t)(uintptr_t)slot;
                a.tier[b] = a.tier_id;
                if (a.hq) heavy_enqueue(a.hq, slot, b);
            }
        }
    }
    if (x.g == 0 && x.cells) atomicAdd(a.stat_cells, x.cells);
}

// ---- planning: size class per bubble, sort key -----------------------------------------------------------
struct TierTable {
    Limits lim[N_TIERS];
};

__global__ void plan_kernel(const uint64_t *__restrict__ seq_off, const uint32_t *__restrict__ bubble_off, uint32_t n,
                            uint32_t big_split, uint32_t *keys, uint32_t *ids) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    const uint32_t s0 = bubble_off[b], ns = bubble_off[b + 1] - s0;
    uint64_t mx = 0;
    for (uint32_t s = 0; s < ns; s++) {
        const uint64_t l = seq_off[s0 + s + 1] - seq_off[s0 + s];
        mx = l > mx ? l : mx;
    }
    uint32_t cls = mx <= big_split ? CLS_BIG : CLS_HUGE;
    if (ns <= LANE_MAX_ROWS) {
#pragma unroll
        for (int c = N_LANE_CLASSES - 1; c >= 0; c--)
            if (mx <= lane_nmax(c)) cls = c;
    }
    // class | rows | longest branch: the 32 bubbles of a warp get alike shapes
    keys[b] = (cls << 28) | (min(ns, 255u) << 20) | (uint32_t)min(mx, (uint64_t)0xFFFFF);
    ids[b] = b;
}

// bounds[c] = first sorted work item of class c (c = 0 .. CLS_HUGE), bounds[CLS_HUGE + 1] = n
__global__ void class_bounds_kernel(const uint32_t *__restrict__ keys, uint32_t n, uint32_t *bounds) {
    const uint32_t c = threadIdx.x;
    if (c > CLS_HUGE + 1) return;
    if (c == CLS_HUGE + 1) { bounds[c] = n; return; }
    uint32_t lo = 0, hi = n;
    const uint32_t want = c << 28;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (keys[mid] < want) lo = mid + 1; else hi = mid;
    }
    bounds[c] = lo;
}

// slot size of every work item; `keys` (sorted) gives the tier of first-pass items, `fixed_tier` >= 0 overrides
__global__ void slot_size_kernel(const uint64_t *__restrict__ seq_off, const uint32_t *__restrict__ bubble_off,
                                 const uint32_t *__restrict__ order, const uint32_t *__restrict__ keys, int fixed_tier,
                                 uint32_t n_items, const TierTable tt, uint64_t *sizes) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w > n_items) return;
    if (w == n_items) { sizes[w] = 0; return; }
    const uint32_t b = order[w];
    const uint32_t s0 = bubble_off[b], ns = bubble_off[b + 1] - s0;
    const uint64_t sum = seq_off[s0 + ns] - seq_off[s0];
    const int t = fixed_tier >= 0 ? fixed_tier : (int)(keys[w] >> 28);
    sizes[w] = slot_layout(ns, sum, tt.lim[t]).bytes;
}

__device__ __forceinline__ bool retryable(int st) {
    return st == PF_BUBBLE_TOO_MANY_ROWS || st == PF_BUBBLE_TOO_LONG || st == PF_BUBBLE_CAND_OVERFLOW ||
           st == PF_BUBBLE_OUT_OVERFLOW || st == PF_BUBBLE_STEP_LIMIT;
}

__global__ void collect_retry_kernel(uint32_t n, const uint64_t *__restrict__ slot_ptr, uint32_t *retry_list, uint32_t *retry_count) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    const SlotHdr *h = (const SlotHdr *)(uintptr_t)slot_ptr[b];
    if (retryable(h->status)) retry_list[atomicAdd(retry_count, 1u)] = b;
}

// A literal '-' in an input sequence is outside the contract: SeqAlign's traceback tells "gap I just opened" from the
// characters of the strings it builds (resB[0] == '-', SeqAlign.cpp:397), so an input dash changes its control flow in a
// way the move-string state machines here do not mirror.  No caller produces one (branch strings are unitig sequences),
// so such a bubble is reported, not guessed at: status PF_BUBBLE_BAD_INPUT, no rows.
__global__ void reject_dash_kernel(const uint8_t *__restrict__ bases, const uint64_t *__restrict__ seq_off,
                                   const uint32_t *__restrict__ bubble_off, uint32_t n, const uint64_t *__restrict__ slot_ptr) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    const uint32_t s0 = bubble_off[b], s1 = bubble_off[b + 1];
    const uint64_t p0 = seq_off[s0], p1 = seq_off[s1];
    bool dash = false;
    uint64_t p = p0;
    for (; p < p1 && ((uintptr_t)(bases + p) & 7); p++) dash |= bases[p] == '-';
    for (; p + 8 <= p1; p += 8) {
        const uint64_t v = *(const uint64_t *)(bases + p) ^ 0x2D2D2D2D2D2D2D2Dull;             // a zero byte where a '-' was
        dash |= ((v - 0x0101010101010101ull) & ~v & 0x8080808080808080ull) != 0;
    }
    for (; p < p1; p++) dash |= bases[p] == '-';
    if (!dash) return;
    SlotHdr *h = (SlotHdr *)(uintptr_t)slot_ptr[b];
    h->status = PF_BUBBLE_BAD_INPUT;
    h->n_rows = 0; h->alen = 0; h->n_var = 0; h->n_ilen = 0;
}

// per bubble: sizes of its four variable-length outputs (+ the scalar outputs)
__global__ void result_size_kernel(const uint64_t *__restrict__ slot_ptr, uint32_t n, int32_t *status, uint32_t *n_rows,
                                   uint32_t *aln_len, uint64_t *sz_rows, uint64_t *sz_var, uint64_t *sz_cls, uint64_t *sz_ilen,
                                   uint32_t *cnt_var, uint32_t *cnt_ilen) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > n) return;
    if (b == n) { sz_rows[b] = sz_var[b] = sz_cls[b] = sz_ilen[b] = 0; return; }
    const SlotHdr *h = (const SlotHdr *)(uintptr_t)slot_ptr[b];
    cnt_var[b] = h->n_var;      // what travels to the host instead of the four offset arrays (align_fetch)
    cnt_ilen[b] = h->n_ilen;
    status[b] = h->status;
    n_rows[b] = h->n_rows;
    aln_len[b] = h->alen;
    sz_rows[b] = (uint64_t)h->n_rows * h->alen;
    sz_var[b] = h->n_var;
    sz_cls[b] = (uint64_t)h->n_var * h->n_rows;
    sz_ilen[b] = h->n_ilen;
}

struct GatherArgs {
    const uint64_t *slot_ptr;
    const uint64_t *seq_off;
    const uint32_t *bubble_off;
    const uint8_t *tier;       // per bubble: which Limits its slot was laid out with
    TierTable tt;
    uint32_t n;
    const uint64_t *off_rows, *off_var, *off_cls, *off_ilen;
    uint8_t *rows;
    uint32_t *var_col;
    uint8_t *var_kind;
    uint16_t *cls;
    uint32_t *ilen;
};

__global__ void gather_kernel(const GatherArgs g) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= g.n) return;
    const uint8_t *slot = (const uint8_t *)(uintptr_t)g.slot_ptr[b];
    const SlotHdr *h = (const SlotHdr *)slot;
    if (h->n_rows == 0) return;
    const uint32_t s0 = g.bubble_off[b], ns = g.bubble_off[b + 1] - s0;
    const SlotLayout lay = slot_layout(ns, g.seq_off[s0 + ns] - g.seq_off[s0], g.tt.lim[g.tier[b]]);
    const uint64_t nrow_bytes = (uint64_t)h->n_rows * h->alen;
    uint8_t *dr = g.rows + g.off_rows[b];
    for (uint64_t i = lane; i < nrow_bytes; i += 32) dr[i] = slot[lay.off_rows + i];
    const uint32_t nv = h->n_var;
    const uint32_t *vc = (const uint32_t *)(slot + lay.off_varcol);
    const uint8_t *vk = slot + lay.off_kind;
    for (uint32_t i = lane; i < nv; i += 32) {
        g.var_col[g.off_var[b] + i] = vc[i];
        g.var_kind[g.off_var[b] + i] = vk[i];
    }
    const uint16_t *cl = (const uint16_t *)(slot + lay.off_cls);
    const uint64_t ncl = (uint64_t)nv * h->n_rows;
    for (uint64_t i = lane; i < ncl; i += 32) g.cls[g.off_cls[b] + i] = cl[i];
    const uint32_t *il = (const uint32_t *)(slot + lay.off_ilen);
    for (uint32_t i = lane; i < h->n_ilen; i += 32) g.ilen[g.off_ilen[b] + i] = il[i];
}

}  // namespace

struct pf_align_state {
    pf::DevBuf ws_warp[3], slots[3], slot_sizes, slot_off, slot_ptr, tier, counter, retry_list, cub_tmp;
    pf::DevBuf keys[2], ids[2];
    pf::DevBuf status, n_rows, aln_len, sz[4], off[4], cnt[2];
    pf::DevBuf rows, var_col, var_kind, cls, ilen;
    pf::PinnedBuf h_cnt[2];
    pf::DevBuf in_bases, in_seq_off, in_bubble_off;
    pf::PinnedBuf h_scalars, h_out[12];
    uint32_t last_retry_count = 0;
    uint32_t last_n = 0;              // bubbles of the last call; its compacted result is still in the buffers below
    uint32_t host_n = 0;              // != 0: the pinned arena (h_out) holds the offsets of that same call (host-pointer forms)
    uint64_t last_tot[4] = {0, 0, 0, 0};   // rows bytes, variable columns, class entries, indel lengths
    uint32_t last_heavy_queued = 0;   // bubbles the first pass pushed on the heavy queue
    uint64_t last_cells = 0;
    uint32_t last_class_count[N_TIERS] = {0};
    int lane_blocks_per_sm[3][N_LANE_CLASSES];   // [variant][class]; variants: LANE_FP64, LANE_I32, LANE_S16X2
    bool lane_attr_done = false;
    cudaStream_t aux[N_LANE_CLASSES + 2] = {nullptr};   // the size classes run concurrently
    cudaEvent_t ev_fork = nullptr, ev_join[N_LANE_CLASSES + 2] = {nullptr};
    pf::DevBuf ws_cta[2];                              // work areas of the CTA kernel (first pass, re-run)
    bool cta_attr_done = false;
    pf::DevBuf ws_cls[N_LANE_CLASSES];
    pf::DevBuf hq, heavy_pool, heavy_ws;               // heavy queue, its slot pool and work areas
    cudaStream_t heavy_stream = nullptr;
    cudaEvent_t ev_heavy = nullptr;
    int heavy_ctas = 4;
    bool heavy_attr_done = false;
    int group_lanes[N_LANE_CLASSES] = {1, 1, 2, 4, 8};   // lanes per bubble of each size class (1 = thread-per-bubble kernel)
    bool group_env_done = false;
};

void pf_align_state_free(pf_align_state *s) {
    if (!s) return;
    for (auto &b : s->ws_cls) b.release();
    for (auto &b : s->ws_cta) b.release();
    s->hq.release(); s->heavy_pool.release(); s->heavy_ws.release();
    if (s->heavy_stream) cudaStreamDestroy(s->heavy_stream);
    if (s->ev_heavy) cudaEventDestroy(s->ev_heavy);
    for (auto &st : s->aux) if (st) cudaStreamDestroy(st);
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    for (auto &e : s->ev_join) if (e) cudaEventDestroy(e);
    pf::DevBuf *d[] = {&s->ws_warp[0], &s->ws_warp[1], &s->ws_warp[2], &s->slots[0], &s->slots[1], &s->slots[2], &s->slot_sizes, &s->slot_off,
                       &s->slot_ptr, &s->tier, &s->counter, &s->retry_list, &s->cub_tmp, &s->keys[0], &s->keys[1], &s->ids[0],
                       &s->ids[1], &s->status, &s->n_rows, &s->aln_len, &s->rows, &s->var_col, &s->var_kind, &s->cls, &s->ilen,
                       &s->in_bases, &s->in_seq_off, &s->in_bubble_off};
    for (auto *b : d) b->release();
    for (auto &b : s->sz) b.release();
    for (auto &b : s->off) b.release();
    s->h_scalars.release();
    for (auto &b : s->h_out) b.release();
    delete s;
}

namespace {

constexpr uint64_t WS_BUDGET = 24ull << 30;  // cap on any one work-area pool

size_t lane_smem_bytes(uint32_t nmax, int) { return (size_t)(LANE_BLOCK / 32) * lane_smem_per_warp(nmax, 4); }

typedef void (*LaneKernel)(const MsaArgs);
LaneKernel lane_kernel(int variant) {
    return variant == LANE_S16X2 ? msa_lane_kernel<LANE_S16X2> : variant == LANE_I32 ? msa_lane_kernel<LANE_I32> : msa_lane_kernel<LANE_FP64>;
}
typedef void (*GroupKernel)(const MsaArgs);
GroupKernel group_kernel(int G) {
    switch (G) {
        case 2: return msa_group_kernel<2>;
        case 4: return msa_group_kernel<4>;
        case 8: return msa_group_kernel<8>;
        case 16: return msa_group_kernel<16>;
        default: return msa_group_kernel<32>;
    }
}
// the s16x2 fill needs every score (border, diagonal, +1 bonuses) to stay well inside int16 even after S16_BLOCK is added
int lane_variant(const Scoring &sc, const Limits &l) {
    if (!sc.integral) return LANE_FP64;
    const long long mag = std::max(std::max(std::llabs((long long)sc.iM), std::llabs((long long)sc.iD)), std::llabs((long long)sc.iG)) + 1;
    return mag * (long long)(l.max_alen + l.max_blen + 2) < 15000 ? LANE_S16X2 : LANE_I32;
}

Limits lane_limits(int c) {
    Limits l;
    l.max_rows = LANE_MAX_ROWS;
    l.max_blen = lane_nmax(c);
    l.max_alen = l.max_blen + 32;
    l.k_cand = 8; l.k_aln = 8; l.max_var = 48;
    static const uint64_t step_env = getenv("PF_LANE_STEP_LIMIT") ? strtoull(getenv("PF_LANE_STEP_LIMIT"), nullptr, 10) : 0;
    l.step_limit = step_env ? step_env : LANE_STEP_LIMIT;
    l.diag_flags = 0; l.pad_ = 0;
    return l;
}

int exclusive_scan_u64(pf_ctx *ctx, pf_align_state *st, const uint64_t *in, uint64_t *out, uint32_t n, cudaStream_t s) {
    size_t tmp = 0;
    PF_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, (int)n, s));
    int rc = st->cub_tmp.reserve(tmp + 16);
    if (rc) return rc;
    tmp = st->cub_tmp.cap;
    PF_CUDA_TRY(cub::DeviceScan::ExclusiveSum(st->cub_tmp.p, tmp, in, out, (int)n, s));
    ctx->launches += 2;
    return PF_OK;
}

void fill_args(MsaArgs &a, pf_align_state *st, int slot_pool, const Limits &lim, const Scoring &sc, const uint8_t *d_bases,
               const uint64_t *d_seq_off, const uint32_t *d_bubble_off, const uint32_t *d_order, uint32_t first, uint32_t n_items,
               int tier_id, uint32_t *counter) {
    a.bases = d_bases; a.seq_off = d_seq_off; a.bubble_off = d_bubble_off; a.order = d_order; a.first = first; a.n_items = n_items;
    a.slot_base = st->slots[slot_pool].as<uint8_t>(); a.slot_off = st->slot_off.as<uint64_t>();
    a.slot_ptr = st->slot_ptr.as<uint64_t>(); a.tier = st->tier.as<uint8_t>(); a.tier_id = (uint8_t)tier_id;
    a.counter = counter; a.lim = lim; a.sc = sc;
    a.stat_cells = (unsigned long long *)(st->counter.as<uint8_t>() + 128);
    a.hq = nullptr;
    static const int wd = getenv("PF_GROUP_WARP_DEQUEUE") ? atoi(getenv("PF_GROUP_WARP_DEQUEUE")) : 0;   // measured: no difference (profiles/r01_summary.md section 9)
    a.warp_dequeue = (uint32_t)wd;
    static const int lp = getenv("PF_LANE_PITCH") ? atoi(getenv("PF_LANE_PITCH")) : 1;
    a.lane_pitch = (uint32_t)lp;
    static const int lc = getenv("PF_LANE_CONTIG") ? atoi(getenv("PF_LANE_CONTIG")) : 0;
    a.lane_contig = (uint32_t)lc;
}

// one launch of the warp-per-bubble kernel over work items [first, first + n_items) of `d_order`
int launch_warp_tier(pf_ctx *ctx, pf_align_state *st, int pool, const Limits &lim, const Scoring &sc, const uint8_t *d_bases,
                     const uint64_t *d_seq_off, const uint32_t *d_bubble_off, const uint32_t *d_order, uint32_t first,
                     uint32_t n_items, int tier_id, uint32_t *counter, cudaStream_t s) {
    const uint64_t ws_bytes = align_up(work_area_bytes(lim), 256);
    uint64_t warps = (uint64_t)ctx->sm_count * 16;
    warps = std::min<uint64_t>(warps, std::max<uint64_t>(1, WS_BUDGET / ws_bytes));
    warps = std::min<uint64_t>(warps, (uint64_t)n_items);
    const uint32_t wpb = WARP_BLOCK / 32;
    const uint32_t blocks = (uint32_t)((warps + wpb - 1) / wpb);
    int rc;
    if ((rc = st->ws_warp[pool].reserve((uint64_t)blocks * wpb * ws_bytes))) return rc;
    MsaArgs a;
    fill_args(a, st, pool, lim, sc, d_bases, d_seq_off, d_bubble_off, d_order, first, n_items, tier_id, counter);
    a.ws_base = st->ws_warp[pool].as<uint8_t>(); a.ws_stride = ws_bytes;
    if (sc.integral) msa_warp_kernel<true><<<blocks, WARP_BLOCK, 0, s>>>(a);
    else msa_warp_kernel<false><<<blocks, WARP_BLOCK, 0, s>>>(a);
    ctx->launches++;
    PF_CUDA_TRY(cudaGetLastError());
    return PF_OK;
}

// the shared-memory tier in drain mode: one warp per CTA over a host-sized list of bubbles (msa_heavy_kernel)
int heavy_mode(const Scoring &sc, const Limits &lim) { return lane_variant(sc, lim) == LANE_S16X2 ? 2 : (sc.integral ? 1 : 0); }
size_t heavy_smem_bytes(const Limits &hl, int hmode) {
    return (size_t)(align_up(flag_area_cells(hl), 16) + align_up(hl.max_alen + hl.max_blen, 16) + (hmode == 2 ? group_smem_per_warp(hl.max_blen, 1) : 0));
}
int heavy_set_attr(bool &done) {   // function attributes are per device: remembered per context, not per process
    if (done) return PF_OK;
    const int big_smem = 200 * 1024;
    PF_CUDA_TRY(cudaFuncSetAttribute(msa_heavy_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem));
    PF_CUDA_TRY(cudaFuncSetAttribute(msa_heavy_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem));
    PF_CUDA_TRY(cudaFuncSetAttribute(msa_heavy_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem));
    done = true;
    return PF_OK;
}
void heavy_launch(int hmode, unsigned grid, size_t smem, cudaStream_t s, const MsaArgs &a) {
    if (hmode == 2) msa_heavy_kernel<2><<<grid, 32, smem, s>>>(a);
    else if (hmode == 1) msa_heavy_kernel<1><<<grid, 32, smem, s>>>(a);
    else msa_heavy_kernel<0><<<grid, 32, smem, s>>>(a);
}

int launch_warp_smem_tier(pf_ctx *ctx, pf_align_state *st, int pool, const Limits &lim, const Scoring &sc, const uint8_t *d_bases,
                          const uint64_t *d_seq_off, const uint32_t *d_bubble_off, const uint32_t *d_order, uint32_t n_items,
                          int tier_id, uint32_t *counter, cudaStream_t s) {
    Limits hl = lim;
    const int hmode = heavy_mode(sc, lim);
    if (hmode == 2) hl.diag_flags = 0;
    int rc;
    if ((rc = heavy_set_attr(st->heavy_attr_done))) return rc;
    const uint64_t ws_bytes = align_up(work_area_bytes(hl), 256);
    const uint32_t blocks = (uint32_t)std::min<uint64_t>((uint64_t)n_items, (uint64_t)ctx->sm_count);
    if ((rc = st->ws_warp[pool].reserve((uint64_t)blocks * ws_bytes))) return rc;
    MsaArgs a;
    fill_args(a, st, pool, hl, sc, d_bases, d_seq_off, d_bubble_off, d_order, 0, n_items, tier_id, counter);
    a.ws_base = st->ws_warp[pool].as<uint8_t>(); a.ws_stride = ws_bytes;
    heavy_launch(hmode, blocks, heavy_smem_bytes(hl, hmode), s, a);
    ctx->launches++;
    PF_CUDA_TRY(cudaGetLastError());
    return PF_OK;
}

// one launch of the G-lanes-per-bubble kernel over work items [first, first + n_items) of `d_order`
int launch_group_tier(pf_ctx *ctx, pf_align_state *st, pf::DevBuf &pool, const Limits &lim, int GL, const Scoring &sc,
                      const uint8_t *d_bases, const uint64_t *d_seq_off, const uint32_t *d_bubble_off, const uint32_t *d_order,
                      uint32_t first, uint32_t n_items, int tier_id, uint32_t *counter, cudaStream_t s, uint32_t *hq = nullptr) {
    const uint32_t NB = 32 / GL, wpb = GROUP_BLOCK / 32;
    const uint64_t ws_bytes = align_up(work_area_bytes(lim, NB), 256);
    const size_t smem = (size_t)wpb * group_smem_per_warp(lim.max_blen, NB);
    if (smem > 200 * 1024) { pf::set_error("group kernel: %u-base branches do not fit shared memory", lim.max_blen); return PF_E_INVALID; }
    int per_sm = 0;
    PF_CUDA_TRY(cudaFuncSetAttribute(group_kernel(GL), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PF_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, group_kernel(GL), GROUP_BLOCK, smem));
    uint64_t warps = (uint64_t)ctx->sm_count * std::max(1, per_sm) * wpb;
    warps = std::min<uint64_t>(warps, std::max<uint64_t>(1, WS_BUDGET / ws_bytes));
    warps = std::min<uint64_t>(warps, ((uint64_t)n_items + NB - 1) / NB);
    const uint32_t grid = (uint32_t)((warps + wpb - 1) / wpb);
    int rc;
    if ((rc = pool.reserve((uint64_t)grid * wpb * ws_bytes))) return rc;
    MsaArgs a;
    fill_args(a, st, 0, lim, sc, d_bases, d_seq_off, d_bubble_off, d_order, first, n_items, tier_id, counter);
    a.ws_base = pool.as<uint8_t>();
    a.ws_stride = ws_bytes;
    a.hq = hq;
    group_kernel(GL)<<<grid, GROUP_BLOCK, smem, s>>>(a);
    ctx->launches++;
    PF_CUDA_TRY(cudaGetLastError());
    return PF_OK;
}


// one launch of the CTA-per-bubble kernel over work items [first, first + n_items) of `d_order`
int launch_cta_tier(pf_ctx *ctx, pf_align_state *st, pf::DevBuf &pool, int slot_pool, const Limits &lim, const Scoring &sc, const uint8_t *d_bases,
                    const uint64_t *d_seq_off, const uint32_t *d_bubble_off, const uint32_t *d_order, uint32_t first, uint32_t n_items,
                    int tier_id, uint32_t *counter, cudaStream_t s) {
    const uint64_t ws_bytes = align_up(work_area_bytes(lim), 256);
    const size_t smem = cta_smem_bytes(lim.max_blen);
    if (smem > 200 * 1024) { pf::set_error("CTA kernel: %u-base branches do not fit shared memory", lim.max_blen); return PF_E_INVALID; }
    if (!st->cta_attr_done) {
        PF_CUDA_TRY(cudaFuncSetAttribute(msa_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        st->cta_attr_done = true;
    }
    int per_sm = 0;
    PF_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, msa_cta_kernel, CTA_THREADS, smem));
    uint64_t ctas = (uint64_t)ctx->sm_count * std::max(1, per_sm);
    ctas = std::min<uint64_t>(ctas, std::max<uint64_t>(1, WS_BUDGET / ws_bytes));
    ctas = std::min<uint64_t>(ctas, (uint64_t)n_items);
    int rc;
    if ((rc = pool.reserve(ctas * ws_bytes))) return rc;
    MsaArgs a;
    fill_args(a, st, slot_pool, lim, sc, d_bases, d_seq_off, d_bubble_off, d_order, first, n_items, tier_id, counter);
    a.ws_base = pool.as<uint8_t>(); a.ws_stride = ws_bytes;
    msa_cta_kernel<<<(unsigned)ctas, CTA_THREADS, smem, s>>>(a);
    ctx->launches++;
    PF_CUDA_TRY(cudaGetLastError());
    return PF_OK;
}

// limits of the CTA kernel: generous from the start (a re-run of a multi-kilobase bubble would cost as much as the pass itself);
// max_alen is capped so that the skewed flag area stays below 2^32 bytes (32-bit cell indices in the traceback)
Limits cta_limits(uint32_t max_blen, uint32_t max_rows, uint32_t alen_factor) {
    Limits l;
    l.max_rows = std::min<uint32_t>(std::max<uint32_t>(max_rows, 2), 64);
    l.max_blen = std::max<uint32_t>(max_blen, 1);
    l.diag_flags = LAYOUT_SKEW; l.pad_ = CTA_THREADS;
    uint64_t alen = (uint64_t)l.max_blen * alen_factor + 64;
    const uint64_t wT = (uint64_t)((l.max_blen + CTA_THREADS) / CTA_THREADS) * CTA_THREADS;
    const uint64_t cap = (0xFFFFFFFFull / wT) - CTA_THREADS - 2;
    l.max_alen = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(alen, cap), 1u << 20);
    l.k_cand = 64; l.k_aln = 64; l.max_var = l.max_alen;
    l.step_limit = 2000000000ull;
    return l;
}

struct DevResult {
    uint64_t tot_rows, tot_var, tot_cls, tot_ilen;
};

// Full device pipeline.  Leaves the compacted arrays in st->{status,n_rows,...}.
int align_device(pf_ctx *ctx, const Scoring &sc, const uint8_t *d_bases, const uint64_t *d_seq_off, uint32_t n_seq,
                 const uint32_t *d_bubble_off, uint32_t n_bubbles, uint32_t max_len, uint32_t max_rows, cudaStream_t s,
                 DevResult &res) {
    if (!ctx->align) ctx->align = new pf_align_state();
    pf_align_state *st = ctx->align;
    const uint32_t n = n_bubbles;
    int rc;
    if ((rc = st->slot_ptr.reserve((uint64_t)n * 8 + 8))) return rc;
    if ((rc = st->tier.reserve((uint64_t)n + 8))) return rc;
    if ((rc = st->retry_list.reserve((uint64_t)n * 4 + 64))) return rc;
    if ((rc = st->counter.reserve(1024))) return rc;
    if ((rc = st->h_scalars.reserve(1024))) return rc;
    for (int i = 0; i < 2; i++) {
        if ((rc = st->keys[i].reserve((uint64_t)n * 4 + 16))) return rc;
        if ((rc = st->ids[i].reserve((uint64_t)n * 4 + 16))) return rc;
    }
    if ((rc = st->slot_sizes.reserve((uint64_t)(n + 1) * 8))) return rc;
    if ((rc = st->slot_off.reserve((uint64_t)(n + 1) * 8))) return rc;

    if (!st->lane_attr_done) {
        for (int v = 0; v < 3; v++) {
            PF_CUDA_TRY(cudaFuncSetAttribute(lane_kernel(v), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)lane_smem_bytes(lane_nmax(N_LANE_CLASSES - 1), v)));
            for (int c = 0; c < N_LANE_CLASSES; c++)
                PF_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&st->lane_blocks_per_sm[v][c], lane_kernel(v), LANE_BLOCK,
                                                                          lane_smem_bytes(lane_nmax(c), v)));
        }
        for (auto &a : st->aux) PF_CUDA_TRY(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
        PF_CUDA_TRY(cudaEventCreateWithFlags(&st->ev_fork, cudaEventDisableTiming));
        for (auto &e : st->ev_join) PF_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        st->lane_attr_done = true;
    }

    if (!st->group_env_done) {   // PF_GROUP_LANES="1,1,4,4,8": lanes per bubble of the five size classes (tuning / A-B runs)
        if (const char *e = getenv("PF_GROUP_LANES")) {
            int v[N_LANE_CLASSES];
            if (sscanf(e, "%d,%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3], &v[4]) == N_LANE_CLASSES)
                for (int c = 0; c < N_LANE_CLASSES; c++)
                    if (v[c] == 1 || v[c] == 2 || v[c] == 4 || v[c] == 8 || v[c] == 16 || v[c] == 32) st->group_lanes[c] = v[c];
        }
        if (const char *e = getenv("PF_HEAVY_CTAS")) { const int v = atoi(e); if (v >= 0 && v <= 64) st->heavy_ctas = v; }
        st->group_env_done = true;
    }
    TierTable tt;
    for (int c = 0; c < N_LANE_CLASSES; c++) tt.lim[c] = lane_limits(c);
    // Branches beyond BIG_SPLIT go to the CTA kernel (integral scoring; the FP64 add+truncate scoring keeps the warp kernel for
    // every long bubble).  PF_BIG_SPLIT overrides the split (diagnostics, A/B runs).
    static const uint32_t split_env = getenv("PF_BIG_SPLIT") ? (uint32_t)strtoul(getenv("PF_BIG_SPLIT"), nullptr, 10) : 0;
    const bool use_cta = sc.integral && cta_smem_bytes(max_len) <= 200 * 1024;
    const uint32_t big_split = use_cta ? (split_env ? split_env : BIG_SPLIT) : 0xFFFFFFFFu;
    Limits &big = tt.lim[CLS_BIG];        // one warp per bubble, first pass: work area sized from the longest branch of the class
    big.max_rows = std::min<uint32_t>(std::max<uint32_t>(max_rows, 2), 8);
    big.max_blen = std::max<uint32_t>(std::min<uint32_t>(max_len, big_split), 1);
    big.max_alen = big.max_blen + std::min<uint32_t>(64, big.max_blen);
    big.k_cand = 8; big.k_aln = 8; big.max_var = 48;
    big.step_limit = 200000000ull; big.diag_flags = 1; big.pad_ = 0;
    // a warp per bubble either way: the strip-pipelined s16x2 fill (group kernel, G = 32, row-major flags) when every score
    // fits int16 and the score row fits shared memory, else the INT32 / FP64 wavefront of the warp kernel
    const bool big_group = lane_variant(sc, big) == LANE_S16X2 && group_smem_per_warp(big.max_blen, 1) * (GROUP_BLOCK / 32) <= 160 * 1024 &&
                           !getenv("PF_BIG_WARP_KERNEL");
    if (big_group) big.diag_flags = 0;
    Limits &heavy = tt.lim[CLS_RETRY_SMEM];   // warp kernel, flag bytes in shared memory: (320+256+1)*321 = 185 KB per CTA
    heavy.max_rows = std::min<uint32_t>(std::max<uint32_t>(max_rows, 2), 64);
    heavy.max_blen = lane_nmax(N_LANE_CLASSES - 1);
    heavy.max_alen = heavy.max_blen + 64;
    heavy.k_cand = 64; heavy.k_aln = 64; heavy.max_var = heavy.max_alen;
    heavy.step_limit = 2000000000ull; heavy.diag_flags = 1; heavy.pad_ = 0;
    tt.lim[CLS_HUGE] = cta_limits(max_len, max_rows, 2);   // CTA kernel, first pass
    Limits &huge = tt.lim[CLS_RETRY];     // second pass: generous (CTA kernel when the scoring is integral, else the warp kernel)
    if (use_cta) huge = cta_limits(max_len, max_rows, std::min<uint32_t>(std::max<uint32_t>(max_rows, 2), 4));
    else {
        huge.max_rows = std::min<uint32_t>(std::max<uint32_t>(max_rows, 2), 64);
        huge.max_blen = std::max<uint32_t>(max_len, 1);
        huge.max_alen = (uint32_t)std::min<uint64_t>((uint64_t)huge.max_blen * std::min<uint32_t>(huge.max_rows, 4) + 64, 1u << 20);
        huge.k_cand = 64; huge.k_aln = 64; huge.max_var = huge.max_alen;
        huge.step_limit = 2000000000ull; huge.diag_flags = 1; huge.pad_ = 0;
    }

    // ---- plan: class + sort ----
    uint32_t *k0 = st->keys[0].as<uint32_t>(), *k1 = st->keys[1].as<uint32_t>();
    uint32_t *i0 = st->ids[0].as<uint32_t>(), *i1 = st->ids[1].as<uint32_t>();
    plan_kernel<<<(n + 255) / 256, 256, 0, s>>>(d_seq_off, d_bubble_off, n, big_split, k0, i0);
    {
        size_t tmp = 0;
        PF_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp, k0, k1, i0, i1, (int)n, 0, 32, s));
        if ((rc = st->cub_tmp.reserve(tmp + 16))) return rc;
        tmp = st->cub_tmp.cap;
        PF_CUDA_TRY(cub::DeviceRadixSort::SortPairs(st->cub_tmp.p, tmp, k0, k1, i0, i1, (int)n, 0, 32, s));
    }
    const uint32_t *d_keys = k1, *d_order = i1;
    uint32_t *d_bounds = st->counter.as<uint32_t>() + 64;   // byte offset 256
    class_bounds_kernel<<<1, 32, 0, s>>>(d_keys, n, d_bounds);
    slot_size_kernel<<<(n + 1 + 255) / 256, 256, 0, s>>>(d_seq_off, d_bubble_off, d_order, d_keys, -1, n, tt, st->slot_sizes.as<uint64_t>());
    ctx->launches += 5;  // plan, sort (2 passes counted as 2), bounds, sizes
    if ((rc = exclusive_scan_u64(ctx, st, st->slot_sizes.as<uint64_t>(), st->slot_off.as<uint64_t>(), n + 1, s))) return rc;
    uint64_t *h_total = st->h_scalars.as<uint64_t>();
    uint32_t *h_bounds = (uint32_t *)(st->h_scalars.as<uint8_t>() + 256);
    PF_CUDA_TRY(cudaMemcpyAsync(h_total, st->slot_off.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, s));
    PF_CUDA_TRY(cudaMemcpyAsync(h_bounds, d_bounds, (CLS_HUGE + 2) * 4, cudaMemcpyDeviceToHost, s));
    PF_CUDA_TRY(cudaMemsetAsync(st->counter.p, 0, 256, s));   // work queues [0..15], stat_cells at byte 128
    PF_CUDA_TRY(pf::stream_sync(s));
    if ((rc = st->slots[0].reserve(*h_total + 64))) return rc;

    // ---- heavy queue: a few one-warp CTAs that re-run exploding DFS bubbles beside the first pass ----
    uint32_t *d_hq = nullptr;
    if (st->heavy_ctas > 0 && h_bounds[CLS_BIG] > 0) {
        Limits hl = heavy;
        const int hmode = heavy_mode(sc, heavy);
        if (hmode == 2) hl.diag_flags = 0;   // row-major flags, one pitch (the group kernel's fill)
        // poll mode asks for (almost) a whole SM's shared memory: the one warp that carries a millisecond-scale serial search
        // should not share its scheduler with first-pass warps
        const size_t hsmem = std::max<size_t>(heavy_smem_bytes(hl, hmode), 190 * 1024);
        const uint64_t hws = align_up(work_area_bytes(hl), 256), pool_bytes = 64ull << 20;
        if ((rc = heavy_set_attr(st->heavy_attr_done))) return rc;
        if (!st->heavy_stream) {
            PF_CUDA_TRY(cudaStreamCreateWithFlags(&st->heavy_stream, cudaStreamNonBlocking));
            PF_CUDA_TRY(cudaEventCreateWithFlags(&st->ev_heavy, cudaEventDisableTiming));
        }
        if ((rc = st->hq.reserve((HQ_HDR + HQ_CAP) * 4 + (uint64_t)HQ_CAP * 24 + 64))) return rc;
        if ((rc = st->heavy_pool.reserve(pool_bytes))) return rc;
        if ((rc = st->heavy_ws.reserve((uint64_t)st->heavy_ctas * hws))) return rc;
        d_hq = st->hq.as<uint32_t>();
        PF_CUDA_TRY(cudaMemsetAsync(d_hq, 0, HQ_HDR * 4, s));
        PF_CUDA_TRY(cudaMemsetAsync(d_hq + HQ_HDR, 0xFF, HQ_CAP * 4, s));
        PF_CUDA_TRY(cudaEventRecord(st->ev_fork, s));
        PF_CUDA_TRY(cudaStreamWaitEvent(st->heavy_stream, st->ev_fork, 0));
        MsaArgs ha;
        fill_args(ha, st, 0, hl, sc, d_bases, d_seq_off, d_bubble_off, d_order, (uint32_t)(pool_bytes >> 10), 0, CLS_RETRY_SMEM, nullptr);
        ha.slot_base = st->heavy_pool.as<uint8_t>();
        ha.ws_base = st->heavy_ws.as<uint8_t>(); ha.ws_stride = hws;
        ha.hq = d_hq;
        heavy_launch(hmode, (unsigned)st->heavy_ctas, hsmem, st->heavy_stream, ha);
        ctx->launches++;
        PF_CUDA_TRY(cudaGetLastError());
        PF_CUDA_TRY(cudaEventRecord(st->ev_heavy, st->heavy_stream));
    }
    // ---- first pass: every non-empty size class on its own stream (they overlap; heaviest first) ----
    PF_CUDA_TRY(cudaEventRecord(st->ev_fork, s));
    {
        const uint32_t cnt = h_bounds[CLS_HUGE + 1] - h_bounds[CLS_HUGE];
        st->last_class_count[CLS_HUGE] = cnt;
        if (cnt) {
            cudaStream_t as = st->aux[N_LANE_CLASSES + 1];
            PF_CUDA_TRY(cudaStreamWaitEvent(as, st->ev_fork, 0));
            if ((rc = launch_cta_tier(ctx, st, st->ws_cta[0], 0, tt.lim[CLS_HUGE], sc, d_bases, d_seq_off, d_bubble_off, d_order, h_bounds[CLS_HUGE], cnt,
                                      CLS_HUGE, st->counter.as<uint32_t>() + CLS_HUGE, as))) return rc;
            PF_CUDA_TRY(cudaEventRecord(st->ev_join[N_LANE_CLASSES + 1], as));
            PF_CUDA_TRY(cudaStreamWaitEvent(s, st->ev_join[N_LANE_CLASSES + 1], 0));
        }
    }
    {
        const uint32_t cnt = h_bounds[CLS_BIG + 1] - h_bounds[CLS_BIG];
        st->last_class_count[CLS_BIG] = cnt;
        if (cnt) {
            cudaStream_t as = st->aux[N_LANE_CLASSES];
            PF_CUDA_TRY(cudaStreamWaitEvent(as, st->ev_fork, 0));
            if (big_group) rc = launch_group_tier(ctx, st, st->ws_warp[0], big, 32, sc, d_bases, d_seq_off, d_bubble_off, d_order, h_bounds[CLS_BIG], cnt,
                                                  CLS_BIG, st->counter.as<uint32_t>() + CLS_BIG, as);
            else rc = launch_warp_tier(ctx, st, 0, big, sc, d_bases, d_seq_off, d_bubble_off, d_order, h_bounds[CLS_BIG], cnt, CLS_BIG,
                                       st->counter.as<uint32_t>() + CLS_BIG, as);
            if (rc) return rc;
            PF_CUDA_TRY(cudaEventRecord(st->ev_join[N_LANE_CLASSES], as));
            PF_CUDA_TRY(cudaStreamWaitEvent(s, st->ev_join[N_LANE_CLASSES], 0));
        }
    }
    for (int c = N_LANE_CLASSES - 1; c >= 0; c--) {
        const uint32_t cnt = h_bounds[c + 1] - h_bounds[c];
        st->last_class_count[c] = cnt;
        if (!cnt) continue;
        const int v = lane_variant(sc, tt.lim[c]);
        const int GL = v == LANE_S16X2 ? st->group_lanes[c] : 1;
        cudaStream_t as = st->aux[c];
        PF_CUDA_TRY(cudaStreamWaitEvent(as, st->ev_fork, 0));
        MsaArgs a;
        fill_args(a, st, 0, tt.lim[c], sc, d_bases, d_seq_off, d_bubble_off, d_order, h_bounds[c], cnt, c, st->counter.as<uint32_t>() + c);
        a.hq = d_hq;
        if (GL > 1) {   // G lanes per bubble
            if ((rc = launch_group_tier(ctx, st, st->ws_cls[c], tt.lim[c], GL, sc, d_bases, d_seq_off, d_bubble_off, d_order, h_bounds[c], cnt, c,
                                        st->counter.as<uint32_t>() + c, as, d_hq))) return rc;
            ctx->launches--;   // counted once below
        } else {
            const uint64_t ws_bytes = align_up(work_area_bytes(tt.lim[c], 32), 256);
            const uint32_t wpb = LANE_BLOCK / 32;
            uint64_t warps = (uint64_t)ctx->sm_count * std::max(1, st->lane_blocks_per_sm[v][c]) * wpb;
            warps = std::min<uint64_t>(warps, std::max<uint64_t>(1, WS_BUDGET / ws_bytes));
            warps = std::min<uint64_t>(warps, ((uint64_t)cnt + 31) / 32);
            const uint32_t grid = (uint32_t)((warps + wpb - 1) / wpb);
            if ((rc = st->ws_cls[c].reserve((uint64_t)grid * wpb * ws_bytes))) return rc;
            a.ws_base = st->ws_cls[c].as<uint8_t>();
            a.ws_stride = ws_bytes;
            lane_kernel(v)<<<grid, LANE_BLOCK, lane_smem_bytes(lane_nmax(c), v), as>>>(a);
        }
        ctx->launches++;
        PF_CUDA_TRY(cudaGetLastError());
        PF_CUDA_TRY(cudaEventRecord(st->ev_join[c], as));
        PF_CUDA_TRY(cudaStreamWaitEvent(s, st->ev_join[c], 0));
    }
    if (d_hq) {   // the first pass is over: the heavy CTAs finish the bubble they hold and stop
        heavy_done_kernel<<<1, 1, 0, s>>>(d_hq);
        ctx->launches++;
        PF_CUDA_TRY(cudaStreamWaitEvent(s, st->ev_heavy, 0));
    }
    // ---- re-runs: whatever overflowed its tier goes to the shared-memory-flags warp kernel (branches <= 256), and what
    //      still does not fit to the warp kernel with the large limits ----
    st->last_retry_count = 0;
    for (int pass = 0; pass < 2; pass++) {
        const int tier = pass == 0 ? CLS_RETRY_SMEM : CLS_RETRY;
        uint32_t *d_retry_cnt = st->counter.as<uint32_t>() + 16 + pass;
        collect_retry_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, st->slot_ptr.as<uint64_t>(), st->retry_list.as<uint32_t>(), d_retry_cnt);
        ctx->launches++;
        uint32_t *h_cnt = (uint32_t *)(st->h_scalars.as<uint8_t>() + 64);
        PF_CUDA_TRY(cudaMemcpyAsync(h_cnt, d_retry_cnt, 4, cudaMemcpyDeviceToHost, s));
        if (pass == 0 && d_hq) PF_CUDA_TRY(cudaMemcpyAsync(h_cnt + 1, d_hq, 8, cudaMemcpyDeviceToHost, s));   // queue tail, tickets
        PF_CUDA_TRY(pf::stream_sync(s));
        const uint32_t nr = *h_cnt;
        st->last_class_count[tier] = nr;
        if (pass == 0) {
            st->last_retry_count = nr;
            st->last_heavy_queued = d_hq ? h_cnt[1] : 0;
            if (d_hq && getenv("PF_HEAVY_TRACE")) {   // diagnostics: when each queued bubble was taken and finished
                std::vector<unsigned long long> T(3 * HQ_CAP);
                uint32_t hdr[HQ_HDR];
                cudaMemcpy(hdr, d_hq, sizeof(hdr), cudaMemcpyDeviceToHost);
                cudaMemcpy(T.data(), d_hq + HQ_TRACE, T.size() * 8, cudaMemcpyDeviceToHost);
                const uint32_t nt = std::min(hdr[3], HQ_CAP);
                unsigned long long t_min = ~0ull;
                for (uint32_t i = 0; i < nt; i++) t_min = std::min(t_min, T[3 * i + 1]);
                fprintf(stderr, "[heavy] queued %u, taken %u, traced %u\n", hdr[0], hdr[1], nt);
                for (uint32_t i = 0; i < nt; i++)
                    fprintf(stderr, "[heavy]   bubble %llu taken +%.3f ms, ran %.3f ms\n", T[3 * i], (T[3 * i + 1] - t_min) * 1e-6,
                            (T[3 * i + 2] - T[3 * i + 1]) * 1e-6);
            }
        }
        if (!nr) break;
        slot_size_kernel<<<(nr + 1 + 255) / 256, 256, 0, s>>>(d_seq_off, d_bubble_off, st->retry_list.as<uint32_t>(), nullptr, tier, nr, tt,
                                                              st->slot_sizes.as<uint64_t>());
        ctx->launches++;
        if ((rc = exclusive_scan_u64(ctx, st, st->slot_sizes.as<uint64_t>(), st->slot_off.as<uint64_t>(), nr + 1, s))) return rc;
        PF_CUDA_TRY(cudaMemcpyAsync(h_total, st->slot_off.as<uint64_t>() + nr, 8, cudaMemcpyDeviceToHost, s));
        PF_CUDA_TRY(pf::stream_sync(s));
        if ((rc = st->slots[1 + pass].reserve(*h_total + 64))) return rc;
        if (pass == 0) rc = launch_warp_smem_tier(ctx, st, 1, heavy, sc, d_bases, d_seq_off, d_bubble_off, st->retry_list.as<uint32_t>(), nr, tier,
                                                  st->counter.as<uint32_t>() + tier, s);
        else if (use_cta) rc = launch_cta_tier(ctx, st, st->ws_cta[1], 2, huge, sc, d_bases, d_seq_off, d_bubble_off, st->retry_list.as<uint32_t>(), 0, nr, tier,
                                               st->counter.as<uint32_t>() + tier, s);
        else rc = launch_warp_tier(ctx, st, 2, huge, sc, d_bases, d_seq_off, d_bubble_off, st->retry_list.as<uint32_t>(), 0, nr, tier,
                                   st->counter.as<uint32_t>() + tier, s);
        if (rc) return rc;
    }
    // ---- sizes -> offsets ----
    const uint32_t n1 = n + 1;
    if ((rc = st->status.reserve((uint64_t)n1 * 4))) return rc;
    if ((rc = st->n_rows.reserve((uint64_t)n1 * 4))) return rc;
    if ((rc = st->aln_len.reserve((uint64_t)n1 * 4))) return rc;
    for (int i = 0; i < 4; i++) {
        if ((rc = st->sz[i].reserve((uint64_t)n1 * 8))) return rc;
        if ((rc = st->off[i].reserve((uint64_t)n1 * 8))) return rc;
    }
    for (int i = 0; i < 2; i++)
        if ((rc = st->cnt[i].reserve((uint64_t)n1 * 4))) return rc;
    reject_dash_kernel<<<(n + 255) / 256, 256, 0, s>>>(d_bases, d_seq_off, d_bubble_off, n, st->slot_ptr.as<uint64_t>());
    ctx->launches++;
    result_size_kernel<<<(n1 + 255) / 256, 256, 0, s>>>(st->slot_ptr.as<uint64_t>(), n, st->status.as<int32_t>(),
                                                        st->n_rows.as<uint32_t>(), st->aln_len.as<uint32_t>(),
                                                        st->sz[0].as<uint64_t>(), st->sz[1].as<uint64_t>(),
                                                        st->sz[2].as<uint64_t>(), st->sz[3].as<uint64_t>(),
                                                        st->cnt[0].as<uint32_t>(), st->cnt[1].as<uint32_t>());
    ctx->launches++;
    uint64_t *h_tot = st->h_scalars.as<uint64_t>() + 16;
    for (int i = 0; i < 4; i++) {
        if ((rc = exclusive_scan_u64(ctx, st, st->sz[i].as<uint64_t>(), st->off[i].as<uint64_t>(), n1, s))) return rc;
        PF_CUDA_TRY(cudaMemcpyAsync(h_tot + i, st->off[i].as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, s));
    }
    PF_CUDA_TRY(cudaMemcpyAsync(h_tot + 4, st->counter.as<uint8_t>() + 128, 8, cudaMemcpyDeviceToHost, s));
    PF_CUDA_TRY(pf::stream_sync(s));
    st->last_cells = h_tot[4];
    res.tot_rows = h_tot[0]; res.tot_var = h_tot[1]; res.tot_cls = h_tot[2]; res.tot_ilen = h_tot[3];
    st->last_n = n;
    st->host_n = 0;
    for (int i = 0; i < 4; i++) st->last_tot[i] = h_tot[i];
    if ((rc = st->rows.reserve(res.tot_rows + 16))) return rc;
    if ((rc = st->var_col.reserve(res.tot_var * 4 + 16))) return rc;
    if ((rc = st->var_kind.reserve(res.tot_var + 16))) return rc;
    if ((rc = st->cls.reserve(res.tot_cls * 2 + 16))) return rc;
    if ((rc = st->ilen.reserve(res.tot_ilen * 4 + 16))) return rc;
    GatherArgs g;
    g.slot_ptr = st->slot_ptr.as<uint64_t>(); g.seq_off = d_seq_off; g.bubble_off = d_bubble_off;
    g.tier = st->tier.as<uint8_t>(); g.tt = tt; g.n = n;
    g.off_rows = st->off[0].as<uint64_t>(); g.off_var = st->off[1].as<uint64_t>();
    g.off_cls = st->off[2].as<uint64_t>(); g.off_ilen = st->off[3].as<uint64_t>();
    g.rows = st->rows.as<uint8_t>(); g.var_col = st->var_col.as<uint32_t>(); g.var_kind = st->var_kind.as<uint8_t>();
    g.cls = st->cls.as<uint16_t>(); g.ilen = st->ilen.as<uint32_t>();
    const uint64_t threads = (uint64_t)n * 32;
    gather_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(g);
    ctx->launches++;
    PF_CUDA_TRY(cudaGetLastError());
    (void)n_seq;
    return PF_OK;
}

}  // namespace


// device -> pinned host: the compacted arrays of the last align_device call.  The four offset arrays (8 bytes per bubble each) do not
// cross the bus: the variable-column and indel-length COUNTS do (4 bytes each), and the offsets are the prefix sums of
// n_rows * aln_len, n_var, n_var * n_rows and n_ilen, taken on the host.
static int align_fetch(pf_align_state *st, uint32_t n_bubbles, const DevResult &res, cudaStream_t s, pf_msa_batch_t *out) {
    int rc;
    const uint64_t n1 = (uint64_t)n_bubbles + 1;
    static const bool copy_offsets = getenv("PF_COPY_OFFSETS") != nullptr;     // A/B switch: the offsets as the device computed them
    const void *src[12] = {st->status.p, st->n_rows.p, st->aln_len.p, copy_offsets ? st->off[0].p : nullptr, st->rows.p, copy_offsets ? st->off[1].p : nullptr,
                           st->var_col.p, st->var_kind.p, copy_offsets ? st->off[2].p : nullptr, st->cls.p, copy_offsets ? st->off[3].p : nullptr, st->ilen.p};
    const uint64_t bytes[12] = {(uint64_t)n_bubbles * 4, (uint64_t)n_bubbles * 4, (uint64_t)n_bubbles * 4, n1 * 8, res.tot_rows,
                                n1 * 8, res.tot_var * 4, res.tot_var, n1 * 8, res.tot_cls * 2, n1 * 8, res.tot_ilen * 4};
    for (int i = 0; i < 12; i++) {
        if ((rc = st->h_out[i].reserve(bytes[i] + 16))) return rc;
        if (bytes[i] && src[i]) PF_CUDA_TRY(cudaMemcpyAsync(st->h_out[i].p, src[i], bytes[i], cudaMemcpyDeviceToHost, s));
    }
    for (int i = 0; i < 2 && !copy_offsets; i++) {
        if ((rc = st->h_cnt[i].reserve((uint64_t)n_bubbles * 4 + 16))) return rc;
        PF_CUDA_TRY(cudaMemcpyAsync(st->h_cnt[i].p, st->cnt[i].p, (uint64_t)n_bubbles * 4, cudaMemcpyDeviceToHost, s));
    }
    PF_CUDA_TRY(pf::stream_sync(s));
    if (!copy_offsets) {
        const uint32_t *nr = st->h_out[1].as<uint32_t>(), *al = st->h_out[2].as<uint32_t>();
        const uint32_t *nv = st->h_cnt[0].as<uint32_t>(), *ni = st->h_cnt[1].as<uint32_t>();
        uint64_t *o_rows = st->h_out[3].as<uint64_t>(), *o_var = st->h_out[5].as<uint64_t>(), *o_cls = st->h_out[8].as<uint64_t>(),
                 *o_ilen = st->h_out[10].as<uint64_t>();
        uint64_t a = 0, b = 0, c = 0, d = 0;
        for (uint32_t q = 0; q < n_bubbles; q++) {
            o_rows[q] = a; o_var[q] = b; o_cls[q] = c; o_ilen[q] = d;
            a += (uint64_t)nr[q] * al[q]; b += nv[q]; c += (uint64_t)nv[q] * nr[q]; d += ni[q];
        }
        o_rows[n_bubbles] = a; o_var[n_bubbles] = b; o_cls[n_bubbles] = c; o_ilen[n_bubbles] = d;
        if (a != res.tot_rows || b != res.tot_var || c != res.tot_cls || d != res.tot_ilen) {
            pf::set_error("pf_align: offsets rebuilt on the host disagree with the device totals (%llu/%llu rows, %llu/%llu columns)",
                          (unsigned long long)a, (unsigned long long)res.tot_rows, (unsigned long long)b, (unsigned long long)res.tot_var);
            return PF_E_CUDA;
        }
    }
    st->host_n = n_bubbles;
    out->n_bubbles = n_bubbles;
    out->status = st->h_out[0].as<int32_t>(); out->n_rows = st->h_out[1].as<uint32_t>(); out->aln_len = st->h_out[2].as<uint32_t>();
    out->rows_off = st->h_out[3].as<uint64_t>(); out->rows = st->h_out[4].as<char>(); out->var_off = st->h_out[5].as<uint64_t>();
    out->var_col = st->h_out[6].as<uint32_t>(); out->var_kind = st->h_out[7].as<uint8_t>(); out->cls_off = st->h_out[8].as<uint64_t>();
    out->cls = st->h_out[9].as<uint16_t>(); out->ilen_off = st->h_out[10].as<uint64_t>(); out->ilen = st->h_out[11].as<uint32_t>();
    return PF_OK;
}

extern "C" {

int pf_align_dev(pf_ctx *ctx, double M, double D, double G, const void *d_bases, uint64_t n_bases, const void *d_seq_off,
                 uint32_t n_seq, const void *d_bubble_off, uint32_t n_bubbles, uint32_t max_len, uint32_t max_rows,
                 pf_msa_batch_t *out_dev, void *cuda_stream) {
    if (!ctx || !out_dev) { pf::set_error("pf_align_dev: null argument"); return PF_E_INVALID; }
    PF_CUDA_TRY(cudaSetDevice(ctx->device));
    memset(out_dev, 0, sizeof(*out_dev));
    if (n_bubbles == 0) return PF_OK;
    (void)n_bases;
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    DevResult res;
    const Scoring sc = make_scoring(M, D, G);
    int rc = align_device(ctx, sc, (const uint8_t *)d_bases, (const uint64_t *)d_seq_off, n_seq, (const uint32_t *)d_bubble_off,
                          n_bubbles, max_len, max_rows, s, res);
    if (rc) return rc;
    pf_align_state *st = ctx->align;
    out_dev->n_bubbles = n_bubbles;
    out_dev->status = st->status.as<int32_t>(); out_dev->n_rows = st->n_rows.as<uint32_t>();
    out_dev->aln_len = st->aln_len.as<uint32_t>(); out_dev->rows_off = st->off[0].as<uint64_t>();
    out_dev->rows = st->rows.as<char>(); out_dev->var_off = st->off[1].as<uint64_t>();
    out_dev->var_col = st->var_col.as<uint32_t>(); out_dev->var_kind = st->var_kind.as<uint8_t>();
    out_dev->cls_off = st->off[2].as<uint64_t>(); out_dev->cls = st->cls.as<uint16_t>();
    out_dev->ilen_off = st->off[3].as<uint64_t>(); out_dev->ilen = st->ilen.as<uint32_t>();
    return PF_OK;
}

int pf_align(pf_ctx *ctx, double M, double D, double G, const char *bases, const uint64_t *seq_off,
             const uint32_t *bubble_off, uint32_t n_bubbles, pf_msa_batch_t *out) {
    if (!ctx || !out || !seq_off || !bubble_off) { pf::set_error("pf_align: null argument"); return PF_E_INVALID; }
    PF_CUDA_TRY(cudaSetDevice(ctx->device));
    memset(out, 0, sizeof(*out));
    if (n_bubbles == 0) return PF_OK;
    if (!ctx->align) ctx->align = new pf_align_state();
    pf_align_state *st = ctx->align;
    const uint32_t n_seq = bubble_off[n_bubbles];
    if (bubble_off[0] != 0) { pf::set_error("pf_align: bubble_off[0] must be 0"); return PF_E_INVALID; }
    const uint64_t base0 = seq_off[0], n_bases = seq_off[n_seq] - base0;
    uint32_t max_len = 0, max_rows = 0;
    int rc;
    const uint64_t *off = seq_off;                          // zero-based offsets are copied to the device as they are
    if (base0) {                                            // otherwise rebased into pinned staging
        if ((rc = ctx->h_stage[0].reserve((uint64_t)(n_seq + 1) * 8))) return rc;
        uint64_t *o = ctx->h_stage[0].as<uint64_t>();
        for (uint32_t s = 0; s <= n_seq; s++) o[s] = seq_off[s] - base0;
        off = o;
    }
    {
        uint64_t mx = 0;
        for (uint32_t s = 0; s < n_seq; s++) { const uint64_t l = off[s + 1] - off[s]; mx = l > mx ? l : mx; }
        max_len = (uint32_t)std::min<uint64_t>(mx, 0xFFFFFFFFull);
    }
    for (uint32_t b = 0; b < n_bubbles; b++) max_rows = std::max(max_rows, bubble_off[b + 1] - bubble_off[b]);
    cudaStream_t s = ctx->stream;
    if ((rc = st->in_bases.reserve(n_bases + 16))) return rc;
    if ((rc = st->in_seq_off.reserve((uint64_t)(n_seq + 1) * 8))) return rc;
    if ((rc = st->in_bubble_off.reserve((uint64_t)(n_bubbles + 1) * 4))) return rc;
    if (n_bases) PF_CUDA_TRY(cudaMemcpyAsync(st->in_bases.p, bases + base0, n_bases, cudaMemcpyHostToDevice, s));
    PF_CUDA_TRY(cudaMemcpyAsync(st->in_seq_off.p, off, (uint64_t)(n_seq + 1) * 8, cudaMemcpyHostToDevice, s));
    PF_CUDA_TRY(cudaMemcpyAsync(st->in_bubble_off.p, bubble_off, (uint64_t)(n_bubbles + 1) * 4, cudaMemcpyHostToDevice, s));
    DevResult res;
    const Scoring sc = make_scoring(M, D, G);
    rc = align_device(ctx, sc, st->in_bases.as<uint8_t>(), st->in_seq_off.as<uint64_t>(), n_seq, st->in_bubble_off.as<uint32_t>(),
                      n_bubbles, max_len, max_rows, s, res);
    if (rc) return rc;
    return align_fetch(st, n_bubbles, res, s, out);
}

// SequenceAlignment of branches that are ALREADY on the device: the sequences first_seq .. of the batch the preceding
// pf_kmc_cov / pf_kmc_cov_async call on `db` staged there (the host looked the branches up anyway, CDBG.cpp:2016-2031; sending
// them a second time for the alignment would be a third of the step's host-to-device bytes).  Only bubble_off travels.
int pf_align_staged(pf_ctx *ctx, pf_kmc *db, double M, double D, double G, uint32_t first_seq, const uint32_t *bubble_off, uint32_t n_bubbles,
                    uint32_t max_len, uint32_t max_rows, pf_msa_batch_t *out) {
    if (!ctx || !db || !out || !bubble_off) { pf::set_error("pf_align_staged: null argument"); return PF_E_INVALID; }
    PF_CUDA_TRY(cudaSetDevice(ctx->device));
    memset(out, 0, sizeof(*out));
    if (n_bubbles == 0) return PF_OK;
    const uint8_t *d_bases = nullptr;
    const uint64_t *d_off = nullptr;
    uint32_t n_staged = 0;
    cudaEvent_t ready;
    if (pf_kmc_staged_dev(db, &d_bases, &d_off, &n_staged, &ready) != PF_OK) {
        pf::set_error("pf_align_staged: no staged batch on this handle (call pf_kmc_cov / pf_kmc_cov_async with zero-based offsets first)");
        return PF_E_INVALID;
    }
    if (bubble_off[0] != 0) { pf::set_error("pf_align_staged: bubble_off[0] must be 0"); return PF_E_INVALID; }
    const uint32_t n_seq = bubble_off[n_bubbles];
    if ((uint64_t)first_seq + n_seq > n_staged) { pf::set_error("pf_align_staged: sequences %u .. %u are outside the staged batch of %u", first_seq, first_seq + n_seq, n_staged); return PF_E_INVALID; }
    if (!ctx->align) ctx->align = new pf_align_state();
    pf_align_state *st = ctx->align;
    cudaStream_t s = ctx->stream;
    int rc;
    if ((rc = st->in_bubble_off.reserve((uint64_t)(n_bubbles + 1) * 4))) return rc;
    PF_CUDA_TRY(cudaMemcpyAsync(st->in_bubble_off.p, bubble_off, (uint64_t)(n_bubbles + 1) * 4, cudaMemcpyHostToDevice, s));
    PF_CUDA_TRY(cudaStreamWaitEvent(s, ready, 0));
    DevResult res;
    const Scoring sc = make_scoring(M, D, G);
    rc = align_device(ctx, sc, d_bases, d_off + first_seq, n_seq, st->in_bubble_off.as<uint32_t>(), n_bubbles, max_len, max_rows, s, res);
    if (rc) return rc;
    return align_fetch(st, n_bubbles, res, s, out);
}

// diagnostics: how many bubbles of the last pf_align* call needed the second (large-limit) pass
uint32_t pf_align_last_retry_count(const pf_ctx *ctx) { return (ctx && ctx->align) ? ctx->align->last_retry_count : 0; }
// diagnostics: DP cells (m*n summed over every needlemanWunch fill) of the last pf_align* call
uint64_t pf_align_last_cells(const pf_ctx *ctx) { return (ctx && ctx->align) ? ctx->align->last_cells : 0; }
// diagnostics: bubbles per tier of the last call: [0..4] lane-kernel size classes (<=64/96/128/192/256), [5] warp kernel,
// [6] re-runs in the shared-memory-flags warp kernel, [7] re-runs with the large limits
}  // extern "C"

// internal (pf_kmc.cu: lookup phase B): device pointers of the last alignment result of this context
int pf_align_last_dev(pf_ctx *ctx, pf_msa_batch_t *out_dev, uint64_t totals[4]) {
    if (!ctx || !ctx->align || !ctx->align->last_n) return PF_E_INVALID;
    pf_align_state *st = ctx->align;
    memset(out_dev, 0, sizeof(*out_dev));
    out_dev->n_bubbles = st->last_n;
    out_dev->status = st->status.as<int32_t>(); out_dev->n_rows = st->n_rows.as<uint32_t>();
    out_dev->aln_len = st->aln_len.as<uint32_t>(); out_dev->rows_off = st->off[0].as<uint64_t>();
    out_dev->rows = st->rows.as<char>(); out_dev->var_off = st->off[1].as<uint64_t>();
    out_dev->var_col = st->var_col.as<uint32_t>(); out_dev->var_kind = st->var_kind.as<uint8_t>();
    out_dev->cls_off = st->off[2].as<uint64_t>(); out_dev->cls = st->cls.as<uint16_t>();
    out_dev->ilen_off = st->off[3].as<uint64_t>(); out_dev->ilen = st->ilen.as<uint32_t>();
    for (int i = 0; i < 4; i++) totals[i] = st->last_tot[i];
    return PF_OK;
}

// internal (pf_kmc.cu: pf_site_cov): var_off / cls_off of the last alignment as they sit in the context's pinned arena -- the site
// batch's site_off / cov_off are the same numbers, so the host form of pf_site_cov does not copy them a second time
int pf_align_last_host_offsets(pf_ctx *ctx, uint32_t n, const uint64_t **var_off, const uint64_t **cls_off) {
    if (!ctx || !ctx->align || ctx->align->host_n != n || !n) return PF_E_INVALID;
    *var_off = ctx->align->h_out[5].as<uint64_t>();
    *cls_off = ctx->align->h_out[8].as<uint64_t>();
    return PF_OK;
}

extern "C" {

uint32_t pf_align_last_heavy_queued(const pf_ctx *ctx) { return (ctx && ctx->align) ? ctx->align->last_heavy_queued : 0; }

int pf_align_last_tier_counts(const pf_ctx *ctx, uint32_t *out, int n) {
    if (!ctx || !out) return PF_E_INVALID;
    for (int i = 0; i < n; i++) out[i] = (ctx->align && i < N_TIERS) ? ctx->align->last_class_count[i] : 0;
    return PF_OK;
}

}  // extern "C"
