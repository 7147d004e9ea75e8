// pf_align.cu -- batched SeqAlign::SequenceAlignment (src/SeqAlign.cpp:550) on sm_100a.
//
// Execution model: persistent warps pull bubbles from an atomic queue; each warp owns a private work
// area in HBM (DP flag bytes, DFS move string, co-optimal alignment store, two candidate-MSA buffers).
//   * DP fill (needlemanWunch, SeqAlign.cpp:480-547): anti-diagonal wavefront inside the warp --
//     lane = matrix row (32-row blocks), one column per step, neighbours exchanged by __shfl_up_sync,
//     the block's bottom row carried to the next block through a small row buffer.  INT32 ALU only
//     (FP64 add+truncate per term only when -M/-D/-G are not whole numbers).  Flag bytes are written
//     diagonal-major so the 32 lanes of a step store 32 consecutive bytes.
//   * traceback / progressive-MSA filter / site calling: leader lane, pf_align_core.cuh.
//   * results land in per-bubble slots, then a size pass + CUB exclusive scans + a warp-per-bubble
//     gather compact them into the flat pf_msa_batch_t arrays.
//   * two capacity tiers: bubbles that overflow the small tier-1 work area (many co-optimal
//     alignments, long insertions) are re-run with the large tier-2 limits; anything that still does
//     not fit is reported per bubble in status[] -- never silently altered.
#include "pf_common.cuh"
#include "pf_align_core.cuh"

#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstring>

using namespace pfalign;

namespace {

constexpr int MSA_BLOCK = 128;  // 4 warps per CTA
constexpr uint32_t FULL = 0xffffffffu;

// ---- warp-cooperative fill ---------------------------------------------------------------------------------
__device__ __forceinline__ void warp_fill(uint8_t *__restrict__ flags, const uint8_t *A, const uint32_t m,
                                          const uint8_t *B, const uint32_t n, const Scoring &sc, int32_t *brow,
                                          const uint32_t lane) {
    __syncwarp();
    const uint32_t W = m + 1;
    for (uint32_t i = lane; i <= m; i += 32) flags[i * W + i] = i ? (uint8_t)(F_UP * 0x11) : (uint8_t)0;   // column 0 (:486-491)
    for (uint32_t j = 1 + lane; j <= n; j += 32) flags[j * W] = (uint8_t)(F_LEFT * 0x11);                  // row 0 (:492-496)
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
                cur = nw_cell(sc, up, diag, cur, a, (uint8_t)b, block_left);
                flags[(i + (uint32_t)j) * W + i] = (uint8_t)(unpack_f(cur) * 0x11);
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

struct WarpExec {
    uint32_t lane;
    unsigned long long cells;
    __device__ __forceinline__ bool leader() const { return lane == 0; }
    __device__ __forceinline__ uint32_t bcast(uint32_t v) const { return __shfl_sync(FULL, v, 0); }
    __device__ __forceinline__ int bcast_i(int v) const { return __shfl_sync(FULL, v, 0); }
    __device__ __forceinline__ uint32_t bcast_ld(const uint32_t *p) const { return *(const volatile uint32_t *)p; }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    __device__ __forceinline__ void fill(uint8_t *flags, const uint8_t *A, uint32_t m, const uint8_t *B, uint32_t n,
                                         const Scoring &sc, int32_t *brow) {
        warp_fill(flags, A, m, B, n, sc, brow, lane);
        cells += (unsigned long long)m * n;
    }
};

struct MsaArgs {
    const uint8_t *bases;
    const uint64_t *seq_off;
    const uint32_t *bubble_off;
    const uint32_t *order;     // work item -> bubble id (nullptr = identity)
    uint32_t n_items;
    uint8_t *slot_base;
    const uint64_t *slot_off;  // per work item
    uint64_t *slot_ptr;        // per bubble: address of its slot
    uint8_t *ws_base;
    uint64_t ws_stride;
    uint32_t *counter;
    unsigned long long *stat_cells;  // DP cells filled (m*n per needlemanWunch call), for the roofline figure
    Limits lim;
    Scoring sc;
};

__global__ void __launch_bounds__(MSA_BLOCK) msa_kernel(const MsaArgs a) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const WorkArea ws = carve_work_area(a.ws_base + (uint64_t)warp * a.ws_stride, a.lim);
    WarpExec x;
    x.lane = lane;
    x.cells = 0;
    for (;;) {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(a.counter, 1u);
        w = __shfl_sync(FULL, w, 0);
        if (w >= a.n_items) break;
        const uint32_t b = a.order ? a.order[w] : w;
        const uint32_t s0 = a.bubble_off[b], ns = a.bubble_off[b + 1] - s0;
        uint8_t *slot = a.slot_base + a.slot_off[w];
        msa_run(x, a.bases, a.seq_off, s0, ns, ws, a.lim, a.sc, slot);
        if (lane == 0) a.slot_ptr[b] = (uint64_t)(uintptr_t)slot;
    }
    if (lane == 0 && x.cells) atomicAdd(a.stat_cells, x.cells);
}

__global__ void slot_size_kernel(const uint64_t *__restrict__ seq_off, const uint32_t *__restrict__ bubble_off,
                                 const uint32_t *__restrict__ order, uint32_t n_items, Limits lim, uint64_t *sizes) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w > n_items) return;
    if (w == n_items) { sizes[w] = 0; return; }
    const uint32_t b = order ? order[w] : w;
    const uint32_t s0 = bubble_off[b], ns = bubble_off[b + 1] - s0;
    const uint64_t sum = seq_off[s0 + ns] - seq_off[s0];
    sizes[w] = slot_layout(ns, sum, lim).bytes;
}

__device__ __forceinline__ bool retryable(int st) {
    return st == PF_BUBBLE_TOO_MANY_ROWS || st == PF_BUBBLE_TOO_LONG || st == PF_BUBBLE_CAND_OVERFLOW || st == PF_BUBBLE_OUT_OVERFLOW;
}

__global__ void collect_retry_kernel(const uint32_t *__restrict__ order, uint32_t n_items, const uint64_t *__restrict__ slot_ptr,
                                     uint32_t *retry_list, uint32_t *retry_count) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_items) return;
    const uint32_t b = order ? order[w] : w;
    const SlotHdr *h = (const SlotHdr *)(uintptr_t)slot_ptr[b];
    if (retryable(h->status)) retry_list[atomicAdd(retry_count, 1u)] = b;
}

// per bubble: sizes of its four variable-length outputs (+ the scalar outputs)
__global__ void result_size_kernel(const uint64_t *__restrict__ slot_ptr, uint32_t n, int32_t *status, uint32_t *n_rows,
                                   uint32_t *aln_len, uint64_t *sz_rows, uint64_t *sz_var, uint64_t *sz_cls, uint64_t *sz_ilen) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > n) return;
    if (b == n) { sz_rows[b] = sz_var[b] = sz_cls[b] = sz_ilen[b] = 0; return; }
    const SlotHdr *h = (const SlotHdr *)(uintptr_t)slot_ptr[b];
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
    Limits lim[2];
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
    const SlotLayout lay = slot_layout(ns, g.seq_off[s0 + ns] - g.seq_off[s0], g.lim[g.tier[b]]);
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

__global__ void fill_tier_kernel(const uint32_t *__restrict__ order, uint32_t n_items, uint8_t *tier, uint8_t value) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w < n_items) tier[order ? order[w] : w] = value;
}

}  // namespace

struct pf_align_state {
    pf::DevBuf ws[2], slots[2], slot_sizes, slot_off, slot_ptr, tier, counter, retry_list, cub_tmp;
    pf::DevBuf status, n_rows, aln_len, sz[4], off[4];
    pf::DevBuf rows, var_col, var_kind, cls, ilen;
    pf::DevBuf in_bases, in_seq_off, in_bubble_off;
    pf::PinnedBuf h_scalars, h_out[12];
    uint32_t last_retry_count = 0;
    uint64_t last_cells = 0;
};

void pf_align_state_free(pf_align_state *s) {
    if (!s) return;
    for (auto &b : s->ws) b.release();
    for (auto &b : s->slots) b.release();
    pf::DevBuf *d[] = {&s->slot_sizes, &s->slot_off, &s->slot_ptr, &s->tier, &s->counter, &s->retry_list, &s->cub_tmp,
                       &s->status, &s->n_rows, &s->aln_len, &s->rows, &s->var_col, &s->var_kind, &s->cls, &s->ilen,
                       &s->in_bases, &s->in_seq_off, &s->in_bubble_off};
    for (auto *b : d) b->release();
    for (auto &b : s->sz) b.release();
    for (auto &b : s->off) b.release();
    s->h_scalars.release();
    for (auto &b : s->h_out) b.release();
    delete s;
}

namespace {

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

// One tier: lay out slots for the work items, run the MSA kernel.  `order` is a device list of bubble ids
// (nullptr = all bubbles 0..n_items-1).
int run_tier(pf_ctx *ctx, pf_align_state *st, int tier, const Limits &lim, const Scoring &sc, const uint8_t *d_bases,
             const uint64_t *d_seq_off, const uint32_t *d_bubble_off, const uint32_t *d_order, uint32_t n_items,
             cudaStream_t s) {
    int rc;
    if ((rc = st->slot_sizes.reserve((uint64_t)(n_items + 1) * 8))) return rc;
    if ((rc = st->slot_off.reserve((uint64_t)(n_items + 1) * 8))) return rc;
    if ((rc = st->counter.reserve(256))) return rc;
    if ((rc = st->h_scalars.reserve(256))) return rc;
    slot_size_kernel<<<(n_items + 1 + 255) / 256, 256, 0, s>>>(d_seq_off, d_bubble_off, d_order, n_items, lim,
                                                               st->slot_sizes.as<uint64_t>());
    ctx->launches++;
    if ((rc = exclusive_scan_u64(ctx, st, st->slot_sizes.as<uint64_t>(), st->slot_off.as<uint64_t>(), n_items + 1, s))) return rc;
    uint64_t *h_total = st->h_scalars.as<uint64_t>();
    PF_CUDA_TRY(cudaMemcpyAsync(h_total, st->slot_off.as<uint64_t>() + n_items, 8, cudaMemcpyDeviceToHost, s));
    PF_CUDA_TRY(cudaMemsetAsync(st->counter.p, 0, 64, s));
    PF_CUDA_TRY(cudaStreamSynchronize(s));
    if ((rc = st->slots[tier].reserve(*h_total + 64))) return rc;
    // warps: as many as the SMs can hold, bounded by a work-area budget
    const uint64_t ws_bytes = align_up(work_area_bytes(lim), 256);
    const uint64_t budget = 24ull << 30;
    uint64_t warps = (uint64_t)ctx->sm_count * 16;
    warps = std::min<uint64_t>(warps, std::max<uint64_t>(1, budget / ws_bytes));
    warps = std::min<uint64_t>(warps, (uint64_t)n_items);
    const uint32_t wpb = MSA_BLOCK / 32;
    const uint32_t blocks = (uint32_t)((warps + wpb - 1) / wpb);
    if ((rc = st->ws[tier].reserve((uint64_t)blocks * wpb * ws_bytes))) return rc;
    MsaArgs a;
    a.bases = d_bases; a.seq_off = d_seq_off; a.bubble_off = d_bubble_off; a.order = d_order; a.n_items = n_items;
    a.slot_base = st->slots[tier].as<uint8_t>(); a.slot_off = st->slot_off.as<uint64_t>();
    a.slot_ptr = st->slot_ptr.as<uint64_t>(); a.ws_base = st->ws[tier].as<uint8_t>(); a.ws_stride = ws_bytes;
    a.counter = st->counter.as<uint32_t>(); a.lim = lim; a.sc = sc;
    a.stat_cells = (unsigned long long *)(st->counter.as<uint8_t>() + 128);
    msa_kernel<<<blocks, MSA_BLOCK, 0, s>>>(a);
    fill_tier_kernel<<<(n_items + 255) / 256, 256, 0, s>>>(d_order, n_items, st->tier.as<uint8_t>(), (uint8_t)tier);
    ctx->launches += 2;
    PF_CUDA_TRY(cudaGetLastError());
    return PF_OK;
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
    int rc;
    if ((rc = st->slot_ptr.reserve((uint64_t)n_bubbles * 8 + 8))) return rc;
    if ((rc = st->tier.reserve((uint64_t)n_bubbles + 8))) return rc;
    if ((rc = st->retry_list.reserve((uint64_t)n_bubbles * 4 + 64))) return rc;
    Limits lim[2];
    // tier 1: small per-warp work area (keeps more of it L2-resident)
    lim[0].max_rows = std::min<uint32_t>(std::max<uint32_t>(max_rows, 2), 8);
    lim[0].max_blen = std::max<uint32_t>(max_len, 1);
    lim[0].max_alen = lim[0].max_blen + std::min<uint32_t>(64, lim[0].max_blen);
    lim[0].k_cand = 8; lim[0].k_aln = 8; lim[0].max_var = 48;
    lim[0].step_limit = 200000000ull;
    // tier 2: generous
    lim[1].max_rows = std::min<uint32_t>(std::max<uint32_t>(max_rows, 2), 64);
    lim[1].max_blen = lim[0].max_blen;
    lim[1].max_alen = (uint32_t)std::min<uint64_t>((uint64_t)lim[1].max_blen * std::min<uint32_t>(lim[1].max_rows, 4) + 64, 1u << 20);
    lim[1].k_cand = 64; lim[1].k_aln = 64; lim[1].max_var = lim[1].max_alen;
    lim[1].step_limit = 2000000000ull;

    if ((rc = st->counter.reserve(256))) return rc;
    PF_CUDA_TRY(cudaMemsetAsync(st->counter.as<uint8_t>() + 128, 0, 8, s));
    if ((rc = run_tier(ctx, st, 0, lim[0], sc, d_bases, d_seq_off, d_bubble_off, nullptr, n_bubbles, s))) return rc;
    // retry list
    uint32_t *d_retry_cnt = st->counter.as<uint32_t>() + 8;
    PF_CUDA_TRY(cudaMemsetAsync(d_retry_cnt, 0, 4, s));
    collect_retry_kernel<<<(n_bubbles + 255) / 256, 256, 0, s>>>(nullptr, n_bubbles, st->slot_ptr.as<uint64_t>(),
                                                                 st->retry_list.as<uint32_t>(), d_retry_cnt);
    ctx->launches++;
    uint32_t *h_cnt = (uint32_t *)(st->h_scalars.as<uint8_t>() + 64);
    PF_CUDA_TRY(cudaMemcpyAsync(h_cnt, d_retry_cnt, 4, cudaMemcpyDeviceToHost, s));
    PF_CUDA_TRY(cudaStreamSynchronize(s));
    st->last_retry_count = *h_cnt;
    if (*h_cnt) {
        if ((rc = run_tier(ctx, st, 1, lim[1], sc, d_bases, d_seq_off, d_bubble_off, st->retry_list.as<uint32_t>(), *h_cnt, s))) return rc;
    }
    // sizes -> offsets
    const uint32_t n1 = n_bubbles + 1;
    if ((rc = st->status.reserve((uint64_t)n1 * 4))) return rc;
    if ((rc = st->n_rows.reserve((uint64_t)n1 * 4))) return rc;
    if ((rc = st->aln_len.reserve((uint64_t)n1 * 4))) return rc;
    for (int i = 0; i < 4; i++) {
        if ((rc = st->sz[i].reserve((uint64_t)n1 * 8))) return rc;
        if ((rc = st->off[i].reserve((uint64_t)n1 * 8))) return rc;
    }
    result_size_kernel<<<(n1 + 255) / 256, 256, 0, s>>>(st->slot_ptr.as<uint64_t>(), n_bubbles, st->status.as<int32_t>(),
                                                        st->n_rows.as<uint32_t>(), st->aln_len.as<uint32_t>(),
                                                        st->sz[0].as<uint64_t>(), st->sz[1].as<uint64_t>(),
                                                        st->sz[2].as<uint64_t>(), st->sz[3].as<uint64_t>());
    ctx->launches++;
    uint64_t *h_tot = st->h_scalars.as<uint64_t>() + 16;
    for (int i = 0; i < 4; i++) {
        if ((rc = exclusive_scan_u64(ctx, st, st->sz[i].as<uint64_t>(), st->off[i].as<uint64_t>(), n1, s))) return rc;
        PF_CUDA_TRY(cudaMemcpyAsync(h_tot + i, st->off[i].as<uint64_t>() + n_bubbles, 8, cudaMemcpyDeviceToHost, s));
    }
    PF_CUDA_TRY(cudaMemcpyAsync(h_tot + 4, st->counter.as<uint8_t>() + 128, 8, cudaMemcpyDeviceToHost, s));
    PF_CUDA_TRY(cudaStreamSynchronize(s));
    st->last_cells = h_tot[4];
    res.tot_rows = h_tot[0]; res.tot_var = h_tot[1]; res.tot_cls = h_tot[2]; res.tot_ilen = h_tot[3];
    if ((rc = st->rows.reserve(res.tot_rows + 16))) return rc;
    if ((rc = st->var_col.reserve(res.tot_var * 4 + 16))) return rc;
    if ((rc = st->var_kind.reserve(res.tot_var + 16))) return rc;
    if ((rc = st->cls.reserve(res.tot_cls * 2 + 16))) return rc;
    if ((rc = st->ilen.reserve(res.tot_ilen * 4 + 16))) return rc;
    GatherArgs g;
    g.slot_ptr = st->slot_ptr.as<uint64_t>(); g.seq_off = d_seq_off; g.bubble_off = d_bubble_off;
    g.tier = st->tier.as<uint8_t>(); g.lim[0] = lim[0]; g.lim[1] = lim[1]; g.n = n_bubbles;
    g.off_rows = st->off[0].as<uint64_t>(); g.off_var = st->off[1].as<uint64_t>();
    g.off_cls = st->off[2].as<uint64_t>(); g.off_ilen = st->off[3].as<uint64_t>();
    g.rows = st->rows.as<uint8_t>(); g.var_col = st->var_col.as<uint32_t>(); g.var_kind = st->var_kind.as<uint8_t>();
    g.cls = st->cls.as<uint16_t>(); g.ilen = st->ilen.as<uint32_t>();
    const uint64_t threads = (uint64_t)n_bubbles * 32;
    gather_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(g);
    ctx->launches++;
    PF_CUDA_TRY(cudaGetLastError());
    (void)n_seq;
    return PF_OK;
}

}  // namespace

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
    std::vector<uint64_t> off(n_seq + 1);
    uint32_t max_len = 0, max_rows = 0;
    for (uint32_t s = 0; s <= n_seq; s++) off[s] = seq_off[s] - base0;
    for (uint32_t s = 0; s < n_seq; s++) max_len = std::max<uint32_t>(max_len, (uint32_t)(off[s + 1] - off[s]));
    for (uint32_t b = 0; b < n_bubbles; b++) max_rows = std::max(max_rows, bubble_off[b + 1] - bubble_off[b]);
    cudaStream_t s = ctx->stream;
    int rc;
    if ((rc = st->in_bases.reserve(n_bases + 16))) return rc;
    if ((rc = st->in_seq_off.reserve((uint64_t)(n_seq + 1) * 8))) return rc;
    if ((rc = st->in_bubble_off.reserve((uint64_t)(n_bubbles + 1) * 4))) return rc;
    if (n_bases) PF_CUDA_TRY(cudaMemcpyAsync(st->in_bases.p, bases + base0, n_bases, cudaMemcpyHostToDevice, s));
    PF_CUDA_TRY(cudaMemcpyAsync(st->in_seq_off.p, off.data(), (uint64_t)(n_seq + 1) * 8, cudaMemcpyHostToDevice, s));
    PF_CUDA_TRY(cudaMemcpyAsync(st->in_bubble_off.p, bubble_off, (uint64_t)(n_bubbles + 1) * 4, cudaMemcpyHostToDevice, s));
    DevResult res;
    const Scoring sc = make_scoring(M, D, G);
    rc = align_device(ctx, sc, st->in_bases.as<uint8_t>(), st->in_seq_off.as<uint64_t>(), n_seq, st->in_bubble_off.as<uint32_t>(),
                      n_bubbles, max_len, max_rows, s, res);
    if (rc) return rc;
    // device -> pinned host
    const uint64_t n1 = (uint64_t)n_bubbles + 1;
    const void *src[12] = {st->status.p, st->n_rows.p, st->aln_len.p, st->off[0].p, st->rows.p, st->off[1].p,
                           st->var_col.p, st->var_kind.p, st->off[2].p, st->cls.p, st->off[3].p, st->ilen.p};
    const uint64_t bytes[12] = {(uint64_t)n_bubbles * 4, (uint64_t)n_bubbles * 4, (uint64_t)n_bubbles * 4, n1 * 8, res.tot_rows,
                                n1 * 8, res.tot_var * 4, res.tot_var, n1 * 8, res.tot_cls * 2, n1 * 8, res.tot_ilen * 4};
    for (int i = 0; i < 12; i++) {
        if ((rc = st->h_out[i].reserve(bytes[i] + 16))) return rc;
        if (bytes[i]) PF_CUDA_TRY(cudaMemcpyAsync(st->h_out[i].p, src[i], bytes[i], cudaMemcpyDeviceToHost, s));
    }
    PF_CUDA_TRY(cudaStreamSynchronize(s));
    out->n_bubbles = n_bubbles;
    out->status = st->h_out[0].as<int32_t>(); out->n_rows = st->h_out[1].as<uint32_t>(); out->aln_len = st->h_out[2].as<uint32_t>();
    out->rows_off = st->h_out[3].as<uint64_t>(); out->rows = st->h_out[4].as<char>(); out->var_off = st->h_out[5].as<uint64_t>();
    out->var_col = st->h_out[6].as<uint32_t>(); out->var_kind = st->h_out[7].as<uint8_t>(); out->cls_off = st->h_out[8].as<uint64_t>();
    out->cls = st->h_out[9].as<uint16_t>(); out->ilen_off = st->h_out[10].as<uint64_t>(); out->ilen = st->h_out[11].as<uint32_t>();
    return PF_OK;
}

// diagnostics: how many bubbles of the last pf_align* call needed the large (tier-2) work area
uint32_t pf_align_last_retry_count(const pf_ctx *ctx) { return (ctx && ctx->align) ? ctx->align->last_retry_count : 0; }
// diagnostics: DP cells (m*n summed over every needlemanWunch fill) of the last pf_align* call
uint64_t pf_align_last_cells(const pf_ctx *ctx) { return (ctx && ctx->align) ? ctx->align->last_cells : 0; }

}  // extern "C"
