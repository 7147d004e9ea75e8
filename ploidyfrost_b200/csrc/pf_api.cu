// pf_api.cu -- context lifetime, error channel and buffer helpers of libpfgpu.so.
#include <cstdlib>
#include "pf_common.cuh"
#include <cuda.h>   // driver-API TYPES only (green contexts); the entry points are fetched at run time, libcuda is not linked
#include <algorithm>

#include <cstring>

namespace pf {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

cudaError_t stream_sync(cudaStream_t s) {
    static const bool blocking = getenv("PF_BLOCKING_SYNC") != nullptr;
    if (!blocking) return cudaStreamSynchronize(s);
    thread_local cudaEvent_t ev = nullptr;      // one per host thread (a pf_ctx is used by one thread at a time)
    cudaError_t e;
    if (!ev && (e = cudaEventCreateWithFlags(&ev, cudaEventBlockingSync | cudaEventDisableTiming)) != cudaSuccess) return e;
    if ((e = cudaEventRecord(ev, s)) != cudaSuccess) return e;
    return cudaEventSynchronize(ev);
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
    // The context's own stream carries the short helper kernels between the big ones (plan, sort, scans, compaction) and the
    // copies: it gets the highest priority, so that when several contexts share the GPU (one per host thread, pf_kmc_share) a
    // helper does not queue behind the pending CTAs of another context's alignment / lookup kernels (those streams stay at the
    // default, lowest priority).  Measured with the CUPTI timeline of bench.py --e2e-profile (profiles/r02_summary.md).
    {
        int least = 0, greatest = 0;
        PF_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        PF_CUDA_TRY(cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, getenv("PF_FLAT_PRIORITY") ? least : greatest));
    }
    if (const char *e = getenv("PF_L2_FETCH_GRANULARITY")) {   // 32 / 64 / 128: a hint to the L2 (random-sector workloads)
        const int g = atoi(e);
        if (g == 32 || g == 64 || g == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)g);
        cudaGetLastError();
    }
    *out = ctx;
    return PF_OK;
}

static void pf_partition_release(pf_ctx *ctx);

void pf_shutdown(pf_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    pf_align_state_free(ctx->align);
    pf_partition_release(ctx);
    for (auto &b : ctx->d_in) b.release();
    for (auto &b : ctx->d_out) b.release();
    for (auto &b : ctx->h_stage) b.release();
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

uint64_t pf_launch_count(const pf_ctx *ctx) { return ctx ? ctx->launches : 0; }

// ---- an SM partition for the k-mer lookups -------------------------------------------------------------------------------
// The lookup kernel is bound by the memory system's random-sector rate, which a fraction of the SMs saturates, while the
// alignment kernels are bound by instruction issue and want every SM they can get.  Run one after the other they each leave
// the other's resource idle; run side by side on ordinary streams whichever is launched first fills every SM with its
// persistent CTAs.  A green context (CUDA 12.4+) gives the lookups a FIXED set of n_sm SMs: their CTAs stay there, and the
// alignment kernels of the same step fill the rest of the device at the same time.  Driver entry points are resolved through
// cudaGetDriverEntryPoint, so the library carries no link-time dependency on libcuda.
static void pf_partition_release(pf_ctx *ctx) {
    if (ctx->part_stream) { cudaStreamSynchronize(ctx->part_stream); cudaStreamDestroy(ctx->part_stream); ctx->part_stream = nullptr; }
    if (ctx->part_green) {
        typedef CUresult (*fn_gdestroy)(CUgreenCtx);
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuGreenCtxDestroy", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && p)
            ((fn_gdestroy)p)((CUgreenCtx)ctx->part_green);
        else cudaGetLastError();
        ctx->part_green = nullptr;
    }
    ctx->part_sms = 0;
}

int pf_lookup_partition(pf_ctx *ctx, uint32_t n_sm, void **stream_out) {
    if (!ctx || !stream_out) { pf::set_error("pf_lookup_partition: null argument"); return PF_E_INVALID; }
    *stream_out = nullptr;
    PF_CUDA_TRY(cudaSetDevice(ctx->device));
    if (ctx->part_stream && ctx->part_sms == n_sm) { *stream_out = ctx->part_stream; return PF_OK; }
    pf_partition_release(ctx);                                  // a different size: the old partition goes first
    if (n_sm < 8 || n_sm + 8 > (uint32_t)ctx->sm_count) { pf::set_error("pf_lookup_partition: %u SMs outside 8 .. %d", n_sm, ctx->sm_count - 8); return PF_E_INVALID; }
    typedef CUresult (*fn_get_res)(CUdevice, CUdevResource *, CUdevResourceType);
    typedef CUresult (*fn_split)(CUdevResource *, unsigned int *, const CUdevResource *, CUdevResource *, unsigned int, unsigned int);
    typedef CUresult (*fn_desc)(CUdevResourceDesc *, CUdevResource *, unsigned int);
    typedef CUresult (*fn_gctx)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned int);
    typedef CUresult (*fn_gstream)(CUstream *, CUgreenCtx, unsigned int, int);
    typedef CUresult (*fn_devget)(CUdevice *, int);
    void *p[6] = {nullptr};
    const char *names[6] = {"cuDeviceGetDevResource", "cuDevSmResourceSplitByCount", "cuDevResourceGenerateDesc", "cuGreenCtxCreate",
                            "cuGreenCtxStreamCreate", "cuDeviceGet"};
    for (int i = 0; i < 6; i++) {
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint(names[i], &p[i], cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p[i]) {
            cudaGetLastError();
            pf::set_error("pf_lookup_partition: the driver does not offer %s (green contexts need CUDA 12.4+)", names[i]);
            return PF_E_UNSUPPORTED;
        }
    }
    PF_CUDA_TRY(cudaFree(0));                                   // the primary context exists
    CUdevice dev;
    CUdevResource sm, part, rest;
    unsigned int groups = 1;
    CUdevResourceDesc desc;
    CUgreenCtx g;
    CUstream st;
#define PF_DRV(call, what) do { CUresult _r = (call); if (_r != CUDA_SUCCESS) { pf::set_error("pf_lookup_partition: %s failed (CUresult %d)", what, (int)_r); return PF_E_CUDA; } } while (0)
    PF_DRV(((fn_devget)p[5])(&dev, ctx->device), "cuDeviceGet");
    PF_DRV(((fn_get_res)p[0])(dev, &sm, CU_DEV_RESOURCE_TYPE_SM), "cuDeviceGetDevResource");
    PF_DRV(((fn_split)p[1])(&part, &groups, &sm, &rest, 0, n_sm), "cuDevSmResourceSplitByCount");
    PF_DRV(((fn_desc)p[2])(&desc, &part, 1), "cuDevResourceGenerateDesc");
    PF_DRV(((fn_gctx)p[3])(&g, desc, dev, CU_GREEN_CTX_DEFAULT_STREAM), "cuGreenCtxCreate");
    PF_DRV(((fn_gstream)p[4])(&st, g, CU_STREAM_NON_BLOCKING, 0), "cuGreenCtxStreamCreate");
#undef PF_DRV
    ctx->part_stream = (cudaStream_t)st;
    ctx->part_sms = part.sm.smCount;
    ctx->part_green = (void *)g;
    *stream_out = ctx->part_stream;
    return PF_OK;
}

uint32_t pf_lookup_partition_sms(const pf_ctx *ctx) { return ctx ? ctx->part_sms : 0; }

int pf_sync(pf_ctx *ctx) {
    if (!ctx) return PF_E_INVALID;
    PF_CUDA_TRY(cudaSetDevice(ctx->device));
    PF_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PF_OK;
}

}  // extern "C"

// ---- roofline denominators ------------------------------------------------------------------------------
namespace {

// every thread chases independent pseudo-random 32-byte sectors of a big table (8 loads in flight)
__global__ void gather_bench_kernel(const uint4 *__restrict__ table, uint64_t n_sectors, uint32_t iters, uint32_t *sink) {
    uint64_t x = (uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
    uint32_t acc = 0;
    for (uint32_t it = 0; it < iters; it++) {
        uint32_t v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            v[u] = __ldg(&table[(x % n_sectors) * 2]).x;   // 2 x uint4 = one 32-byte sector
        }
#pragma unroll
        for (int u = 0; u < 8; u++) acc += v[u];
    }
    if (acc == 0x12345678u) *sink = acc;
}

// The ceiling of a one-sector-per-lookup index: every thread keeps ILP independent random loads of WIDTH bytes (32 = one sector,
// 64 = two adjacent sectors, 16 = half a sector) in flight against a table of 2^bits units, addresses from a xorshift stream
// (shift + mask: no division), L1 no-allocate like the lookup kernel's bucket loads.
template <int WIDTH, int ILP>
__global__ void __launch_bounds__(256) gather_sweep_kernel(const unsigned long long *__restrict__ table, uint32_t unit_bits, uint32_t iters,
                                                           unsigned long long *sink) {
    uint64_t x = ((uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) + 1) * 0x9E3779B97F4A7C15ull;
    const uint64_t mask = (1ull << unit_bits) - 1;
    unsigned long long acc = 0;
    for (uint32_t it = 0; it < iters; it++) {
        unsigned long long v[ILP][WIDTH / 8];
#pragma unroll
        for (int u = 0; u < ILP; u++) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            const unsigned long long *p = table + ((x >> 11) & mask) * (WIDTH / 8);
            if (WIDTH == 16) asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(v[u][0]), "=l"(v[u][1]) : "l"(p));
            else {
                asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[u][0]), "=l"(v[u][1]), "=l"(v[u][2]), "=l"(v[u][3]) : "l"(p));
                if (WIDTH == 64)
                    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[u][4]), "=l"(v[u][5]), "=l"(v[u][6]), "=l"(v[u][7]) : "l"(p + 4));
            }
        }
#pragma unroll
        for (int u = 0; u < ILP; u++)
#pragma unroll
            for (int q = 0; q < WIDTH / 8; q++) acc += v[u][q];
    }
    if (acc == 0x123456789ull) *sink = acc;
}

__global__ void int32_bench_kernel(uint32_t iters, int *sink) {
    int a = threadIdx.x, b = blockIdx.x, c = 7, d = 3;
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            a = max(a + b, c) ^ d;      // IADD + IMNMX + LOP: 3 ops
            b = min(b + c, d) + a;      // 3 ops
            c = (c + a) ^ b;            // 2 ops
            d = max(d + b, a);          // 2 ops
        }
    }
    if ((a ^ b ^ c ^ d) == 0x7fffffff) *sink = a;
}

}  // namespace

extern "C" {

int pf_bench_random_gather(pf_ctx *ctx, uint64_t bytes, double *gb_per_s) {
    if (!ctx || !gb_per_s || bytes < (1u << 20)) { pf::set_error("pf_bench_random_gather: bad argument"); return PF_E_INVALID; }
    PF_CUDA_TRY(cudaSetDevice(ctx->device));
    void *tab = nullptr, *sink = nullptr;
    PF_CUDA_TRY(cudaMalloc(&tab, bytes));
    PF_CUDA_TRY(cudaMalloc(&sink, 4));
    PF_CUDA_TRY(cudaMemsetAsync(tab, 1, bytes, ctx->stream));
    const uint64_t n_sectors = bytes / 32;
    const unsigned blocks = ctx->sm_count * 8, threads = 256;
    const uint32_t iters = 64;
    cudaEvent_t e0, e1;
    PF_CUDA_TRY(cudaEventCreate(&e0));
    PF_CUDA_TRY(cudaEventCreate(&e1));
    gather_bench_kernel<<<blocks, threads, 0, ctx->stream>>>((const uint4 *)tab, n_sectors, 8, (uint32_t *)sink);
    PF_CUDA_TRY(cudaEventRecord(e0, ctx->stream));
    gather_bench_kernel<<<blocks, threads, 0, ctx->stream>>>((const uint4 *)tab, n_sectors, iters, (uint32_t *)sink);
    PF_CUDA_TRY(cudaEventRecord(e1, ctx->stream));
    PF_CUDA_TRY(cudaEventSynchronize(e1));
    ctx->launches += 2;
    float ms = 0;
    PF_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    *gb_per_s = (double)blocks * threads * iters * 8 * 32.0 / (ms * 1e-3) / 1e9;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(tab); cudaFree(sink);
    return PF_OK;
}

// diagnostics: random-access ceiling.  width 16 / 32 / 64 bytes per access, ilp 4 / 8 / 16 loads in flight per thread,
// ctas_per_sm resident 256-thread CTAs; returns accesses per second (G/s).
int pf_bench_gather_sweep(pf_ctx *ctx, uint64_t bytes, uint32_t width, uint32_t ilp, uint32_t ctas_per_sm, double *g_access_per_s) {
    if (!ctx || !g_access_per_s || bytes < (1u << 20)) { pf::set_error("pf_bench_gather_sweep: bad argument"); return PF_E_INVALID; }
    PF_CUDA_TRY(cudaSetDevice(ctx->device));
    uint32_t bits = 0;
    while ((2ull << bits) * width <= bytes) bits++;
    const uint64_t tab_bytes = (1ull << bits) * width;
    void *tab = nullptr, *sink = nullptr;
    PF_CUDA_TRY(cudaMalloc(&tab, tab_bytes));
    PF_CUDA_TRY(cudaMalloc(&sink, 8));
    PF_CUDA_TRY(cudaMemsetAsync(tab, 1, tab_bytes, ctx->stream));
    const unsigned blocks = ctx->sm_count * std::max(1u, ctas_per_sm), threads = 256;
    const uint32_t iters = 256 / ilp * 4;
    auto launch = [&](uint32_t it) {
        const unsigned long long *t = (const unsigned long long *)tab;
        unsigned long long *sk = (unsigned long long *)sink;
#define PF_GS(W, I) gather_sweep_kernel<W, I><<<blocks, threads, 0, ctx->stream>>>(t, bits, it, sk)
        if (width == 16) { if (ilp <= 4) PF_GS(16, 4); else if (ilp <= 8) PF_GS(16, 8); else PF_GS(16, 16); }
        else if (width == 64) { if (ilp <= 4) PF_GS(64, 4); else if (ilp <= 8) PF_GS(64, 8); else PF_GS(64, 16); }
        else { if (ilp <= 4) PF_GS(32, 4); else if (ilp <= 8) PF_GS(32, 8); else PF_GS(32, 16); }
#undef PF_GS
    };
    const uint32_t ilp_eff = ilp <= 4 ? 4 : ilp <= 8 ? 8 : 16;
    cudaEvent_t e0, e1;
    PF_CUDA_TRY(cudaEventCreate(&e0));
    PF_CUDA_TRY(cudaEventCreate(&e1));
    launch(4);
    PF_CUDA_TRY(cudaEventRecord(e0, ctx->stream));
    launch(iters);
    PF_CUDA_TRY(cudaEventRecord(e1, ctx->stream));
    PF_CUDA_TRY(cudaEventSynchronize(e1));
    PF_CUDA_TRY(cudaGetLastError());
    ctx->launches += 2;
    float ms = 0;
    PF_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    *g_access_per_s = (double)blocks * threads * iters * ilp_eff / (ms * 1e-3) / 1e9;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(tab); cudaFree(sink);
    return PF_OK;
}

int pf_bench_int32(pf_ctx *ctx, double *gop_per_s) {
    if (!ctx || !gop_per_s) { pf::set_error("pf_bench_int32: bad argument"); return PF_E_INVALID; }
    PF_CUDA_TRY(cudaSetDevice(ctx->device));
    void *sink = nullptr;
    PF_CUDA_TRY(cudaMalloc(&sink, 4));
    const unsigned blocks = ctx->sm_count * 8, threads = 256;
    const uint32_t iters = 4096;
    cudaEvent_t e0, e1;
    PF_CUDA_TRY(cudaEventCreate(&e0));
    PF_CUDA_TRY(cudaEventCreate(&e1));
    int32_bench_kernel<<<blocks, threads, 0, ctx->stream>>>(64, (int *)sink);
    PF_CUDA_TRY(cudaEventRecord(e0, ctx->stream));
    int32_bench_kernel<<<blocks, threads, 0, ctx->stream>>>(iters, (int *)sink);
    PF_CUDA_TRY(cudaEventRecord(e1, ctx->stream));
    PF_CUDA_TRY(cudaEventSynchronize(e1));
    ctx->launches += 2;
    float ms = 0;
    PF_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    *gop_per_s = (double)blocks * threads * iters * 16 * 10.0 / (ms * 1e-3) / 1e9;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(sink);
    return PF_OK;
}

}  // extern "C"
