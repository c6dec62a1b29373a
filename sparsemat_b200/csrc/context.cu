// Context, events, host memory, error reporting.
#include "common.cuh"
#include "cg_sr.cuh"

namespace smb {

static thread_local char g_err[1024] = "";
thread_local uint64_t g_launches = 0;
thread_local LaunchRedirect g_redirect;
thread_local int g_ring_reserve_sms = 0;
thread_local int g_ring_grid_cap = 0;
thread_local bool g_spmv_accumulate = false;
thread_local bool g_spmv_band = false;
thread_local bool g_x_unpadded = false;
thread_local HaloLaunch g_halo;
thread_local const ArDev* g_dot_ar = nullptr;
thread_local CgSrLaunch g_cgsr;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

void ctx_teardown(smb200_ctx* c);
void ctx_retain(smb200_ctx* ctx) { if (ctx) ++ctx->refs; }
void ctx_release(smb200_ctx* ctx) {
    if (!ctx) return;
    if (--ctx->refs <= 0 && ctx->destroy_requested) ctx_teardown(ctx);
}

smb200_status dev_alloc(void** p, size_t bytes) {
    *p = nullptr;
    SMB_CUDA(cudaMalloc(p, bytes + kPadBytes));
    return SMB200_OK;
}

smb200_status ensure_reduction_scratch(smb200_ctx* ctx, size_t n_partials) {
    if (n_partials <= ctx->red_cap) return SMB200_OK;
    SMB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->red_partials) cudaFree(ctx->red_partials);
    ctx->red_partials = nullptr;
    ctx->red_cap = 0;
    size_t cap = n_partials + n_partials / 2 + 1024;
    SMB_CUDA(cudaMalloc(&ctx->red_partials, cap * sizeof(double)));
    ctx->red_cap = cap;
    return SMB200_OK;
}

smb200_status fetch_result(smb200_ctx* ctx, int slot, double* out) {
    SMB_CUDA(cudaMemcpyAsync(ctx->red_result_host + slot, ctx->red_result + slot, sizeof(double),
                             cudaMemcpyDeviceToHost, ctx->stream));
    SMB_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = ctx->red_result_host[slot];
    return SMB200_OK;
}

__global__ void flush_kernel(uint4* __restrict__ buf, size_t n16, unsigned tag) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n16; i += stride) buf[i] = make_uint4(tag, tag + 1, tag + 2, (unsigned)i);
}

}  // namespace smb

using namespace smb;

extern "C" {

int32_t smb200_version(void) { return SMB200_VERSION; }
const char* smb200_last_error(void) { return get_error(); }
uint64_t smb200_launch_count(void) { return g_launches; }

smb200_status smb200_ctx_create(int32_t device, void* stream, smb200_ctx** out) {
    SMB_REQUIRE(out != nullptr, SMB200_ERR_INVALID, "ctx_create: out is NULL");
    *out = nullptr;
    int count = 0;
    SMB_CUDA(cudaGetDeviceCount(&count));
    SMB_REQUIRE(device >= 0 && device < count, SMB200_ERR_CUDA, "ctx_create: device %d not present (%d CUDA devices)",
                device, count);
    SMB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SMB_CUDA(cudaGetDeviceProperties(&prop, device));
    SMB_REQUIRE(prop.major >= 10, SMB200_ERR_UNSUPPORTED,
                "ctx_create: %s is sm_%d%d; libsmb200 carries sm_100a code only (no fallback path)", prop.name,
                prop.major, prop.minor);
    smb200_ctx* c = new smb200_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->l2_bytes = (size_t)prop.l2CacheSize;
    c->l2_persist_max = (size_t)prop.persistingL2CacheMaxSize;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete c; SMB_CUDA(e); }
        c->own_stream = true;
    }
    // side stream at the highest priority: the few boundary CTAs of a distributed SpMV slip in between the
    // interior kernel's waves instead of queueing behind them
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    cudaError_t e = cudaStreamCreateWithPriority(&c->aux_stream, cudaStreamNonBlocking, prio_hi);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_a, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_b, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&c->red_ticket, 4 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(c->red_ticket, 0, 4 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMalloc(&c->red_result, 16 * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(c->red_result, 0, 16 * sizeof(double));
    if (e == cudaSuccess) e = cudaHostAlloc(&c->red_result_host, 16 * sizeof(double), cudaHostAllocDefault);
    if (e != cudaSuccess) { smb200_ctx_destroy(c); SMB_CUDA(e); }
    smb200_status s = ensure_reduction_scratch(c, 1 << 16);
    if (s != SMB200_OK) { smb200_ctx_destroy(c); return s; }
    *out = c;
    return SMB200_OK;
}

smb200_status smb200_comm_destroy(smb200_ctx* ctx);

smb200_status smb200_ctx_destroy(smb200_ctx* c) {
    if (!c) return SMB200_OK;
    if (c->refs > 0) {                 // handles still alive: finish when the last one is freed
        c->destroy_requested = true;
        if (c->stream) cudaStreamSynchronize(c->stream);
        return SMB200_OK;
    }
    ctx_teardown(c);
    return SMB200_OK;
}

}  // extern "C"

namespace smb {
void ctx_teardown(smb200_ctx* c) {
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm) smb200_comm_destroy(c);
    if (c->red_partials) cudaFree(c->red_partials);
    if (c->red_partials_aux) cudaFree(c->red_partials_aux);
    if (c->red_ticket) cudaFree(c->red_ticket);
    if (c->red_result) cudaFree(c->red_result);
    if (c->red_result_host) cudaFreeHost(c->red_result_host);
    if (c->flush_buf) cudaFree(c->flush_buf);
    if (c->stage_x) cudaFree(c->stage_x);
    if (c->stage_y) cudaFree(c->stage_y);
    if (c->ev_a) cudaEventDestroy(c->ev_a);
    if (c->ev_b) cudaEventDestroy(c->ev_b);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    if (c->copy_in) cudaStreamDestroy(c->copy_in);
    if (c->copy_out) cudaStreamDestroy(c->copy_out);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}
}  // namespace smb

extern "C" {

smb200_status smb200_ctx_sync(smb200_ctx* ctx) {
    SMB_REQUIRE(ctx, SMB200_ERR_INVALID, "ctx_sync: NULL context");
    SMB_CUDA(cudaStreamSynchronize(ctx->stream));
    return SMB200_OK;
}

smb200_status smb200_ctx_devinfo(smb200_ctx* ctx, smb200_devinfo* out) {
    SMB_REQUIRE(ctx && out, SMB200_ERR_INVALID, "ctx_devinfo: NULL argument");
    cudaDeviceProp prop;
    SMB_CUDA(cudaGetDeviceProperties(&prop, ctx->device));
    memset(out, 0, sizeof *out);
    out->device = ctx->device;
    out->sm_count = prop.multiProcessorCount;
    out->cc_major = prop.major;
    out->cc_minor = prop.minor;
    out->l2_bytes = prop.l2CacheSize;
    out->l2_persist_max_bytes = prop.persistingL2CacheMaxSize;
    out->hbm_bytes = (int64_t)prop.totalGlobalMem;
    snprintf(out->name, sizeof out->name, "%.*s", (int)sizeof out->name - 1, prop.name);
    return SMB200_OK;
}

smb200_status smb200_ctx_flush_l2(smb200_ctx* ctx) {
    SMB_REQUIRE(ctx, SMB200_ERR_INVALID, "flush_l2: NULL context");
    if (!ctx->flush_buf) {
        ctx->flush_bytes = ctx->l2_bytes * 2 + (64u << 20);
        SMB_CUDA(cudaMalloc(&ctx->flush_buf, ctx->flush_bytes));
    }
    static unsigned tag = 0;
    flush_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((uint4*)ctx->flush_buf, ctx->flush_bytes / 16, ++tag);
    SMB_CUDA(cudaGetLastError());
    return SMB200_OK;
}

smb200_status smb200_event_create(smb200_ctx* ctx, smb200_event** out) {
    SMB_REQUIRE(ctx && out, SMB200_ERR_INVALID, "event_create: NULL argument");
    smb200_event* e = new smb200_event();
    e->ctx = ctx;
    cudaError_t err = cudaEventCreate(&e->ev);
    if (err != cudaSuccess) { delete e; SMB_CUDA(err); }
    ctx_retain(ctx);
    *out = e;
    return SMB200_OK;
}
smb200_status smb200_event_record(smb200_event* ev) {
    SMB_REQUIRE(ev, SMB200_ERR_INVALID, "event_record: NULL event");
    SMB_CUDA(cudaEventRecord(ev->ev, ev->ctx->stream));
    return SMB200_OK;
}
smb200_status smb200_event_elapsed_ms(smb200_event* start, smb200_event* stop, float* ms) {
    SMB_REQUIRE(start && stop && ms, SMB200_ERR_INVALID, "event_elapsed_ms: NULL argument");
    SMB_CUDA(cudaEventSynchronize(stop->ev));
    SMB_CUDA(cudaEventElapsedTime(ms, start->ev, stop->ev));
    return SMB200_OK;
}
smb200_status smb200_event_destroy(smb200_event* ev) {
    if (!ev) return SMB200_OK;
    cudaEventDestroy(ev->ev);
    smb200_ctx* ctx = ev->ctx;
    delete ev;
    ctx_release(ctx);
    return SMB200_OK;
}

smb200_status smb200_host_alloc(size_t bytes, void** out) {
    SMB_REQUIRE(out, SMB200_ERR_INVALID, "host_alloc: out is NULL");
    *out = nullptr;
    SMB_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return SMB200_OK;
}
smb200_status smb200_host_free(void* p) {
    if (p) SMB_CUDA(cudaFreeHost(p));
    return SMB200_OK;
}

}  // extern "C"
