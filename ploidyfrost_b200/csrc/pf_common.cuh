// pf_common.cuh -- shared host-side plumbing of libpfgpu.so: error channel, CUDA checks, the context.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/pf_gpu.h"

namespace pf {

void set_error(const char *fmt, ...);
// The host's wait for a stream.  Default: cudaStreamSynchronize (the runtime's spin wait, lowest latency).  PF_BLOCKING_SYNC=1 in the
// environment: record a cudaEventBlockingSync event and sleep on it -- for hosts where every GPU's worker threads spinning at once
// (N ranks x T threads) leaves no core for the threads that have work to enqueue.
cudaError_t stream_sync(cudaStream_t s);

#define PF_CUDA_TRY(expr)                                                                     \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            pf::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return PF_E_CUDA;                                                                 \
        }                                                                                     \
    } while (0)

// Grow-only device / pinned-host buffers (work areas are reused across calls).
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
    template <class T> T *as() const { return (T *)p; }
};
struct PinnedBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
    template <class T> T *as() const { return (T *)p; }
};

}  // namespace pf

struct pf_align_state;  // defined in pf_align.cu
// device pointers + totals {row bytes, variable columns, class entries, indel lengths} of the context's last alignment result
int pf_align_last_dev(pf_ctx *ctx, pf_msa_batch_t *out_dev, uint64_t totals[4]);
// var_off / cls_off of that result in the context's pinned arena (host-pointer forms of pf_align only; PF_E_INVALID otherwise)
int pf_align_last_host_offsets(pf_ctx *ctx, uint32_t n, const uint64_t **var_off, const uint64_t **cls_off);

struct pf_kmc;
// device copy of the sequences of the handle's last pf_kmc_cov / pf_kmc_cov_async call (zero-based offsets) + the event after which they are there
extern "C" int pf_kmc_staged_dev(pf_kmc *db, const uint8_t **bases, const uint64_t **seq_off, uint32_t *n_seq, cudaEvent_t *ready);

struct pf_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;
    pf_align_state *align = nullptr;
    // scratch shared by the host-pointer entry points
    pf::DevBuf d_in[4];
    pf::DevBuf d_out[4];
    pf::PinnedBuf h_stage[2];
    // SM partition of the lookups (pf_lookup_partition): a green-context stream confined to part_sms SMs
    cudaStream_t part_stream = nullptr;
    uint32_t part_sms = 0;
    void *part_green = nullptr;
};
