// pf_api.cu -- context lifetime, error channel and buffer helpers of libpfgpu.so.
#include "pf_common.cuh"

#include <cstring>

namespace pf {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return PF_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("device allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
        p = nullptr;
        return PF_E_NOMEM;
    }
    cap = want;
    return PF_OK;
}
void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}
int PinnedBuf::reserve(size_t bytes) {
    if (bytes <= cap) return PF_OK;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("pinned host allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
        p = nullptr;
        return PF_E_NOMEM;
    }
    cap = want;
    return PF_OK;
}
void PinnedBuf::release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
}

}  // namespace pf

void pf_align_state_free(pf_align_state *);  // pf_align.cu

extern "C" {

const char *pf_last_error(void) { return pf::g_err; }

const char *pf_version(void) { return "libpfgpu 0.1 (sm_100a; KMC lookup + SeqAlign kernels; no CPU fallback)"; }

int pf_init(int device, pf_ctx **out) {
    if (!out) { pf::set_error("pf_init: null output"); return PF_E_INVALID; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        pf::set_error("pf_init: no CUDA device available (%s); this library has no CPU fallback",
                      e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        cudaGetLastError();
        return PF_E_CUDA;
    }
    if (device < 0 || device >= n) { pf::set_error("pf_init: device %d out of range (0..%d)", device, n - 1); return PF_E_INVALID; }
    PF_CUDA_TRY(cudaSetDevice(device));
    pf_ctx *ctx = new pf_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    PF_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    PF_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    *out = ctx;
    return PF_OK;
}

void pf_shutdown(pf_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    pf_align_state_free(ctx->align);
    for (auto &b : ctx->d_in) b.release();
    for (auto &b : ctx->d_out) b.release();
    for (auto &b : ctx->h_stage) b.release();
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

uint64_t pf_launch_count(const pf_ctx *ctx) { return ctx ? ctx->launches : 0; }

int pf_sync(pf_ctx *ctx) {
    if (!ctx) return PF_E_INVALID;
    PF_CUDA_TRY(cudaSetDevice(ctx->device));
    PF_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PF_OK;
}

}  // extern "C"
