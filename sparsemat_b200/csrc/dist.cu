// K9 — one process per GPU: row-block distributed SpMV and CG.
//
// Partitioning is SparseMatPar's model (sparsemat_par.rs:20-35): contiguous row blocks, every block
// needs the slice of x its columns touch.  The reference's sketch replicates x (`Arc<rhs>`,
// sparsemat_par.rs:39-67); here each rank keeps only [owned | pad | ghosts] and the ghosts are refreshed per product.
//
// Data path (default): PEER MEMORY over NVLink / NVSwitch, no NCCL kernel anywhere in a product or a CG iteration.
// At creation every rank allocates one "window" (ghost double buffer, flag words, all-reduce slots), the CUDA IPC
// handles are exchanged once (NCCL all-gather, set-up only) and mapped.  A product is then ONE launch of the ring kernel
// (spmv.cu): its consumer warps first store the entries the neighbours need straight into the neighbours' ghost buffers
// and raise their flags; its TMA producer walks the row blocks rotated so that the rows referencing ghosts come last,
// and spin-waits on this rank's flags only when it reaches them (halo.cuh has the protocol and its hazards argument).
// Matrices whose local plan is not the ring kernel use the same protocol from two small kernels around the interior
// launch.  The CG scalars are all-reduced by a 1-CTA kernel that stores every rank's partial into every peer's slot and
// sums them in rank order (bit-identical on all ranks).
// Fallback (SMB200_DIST_P2P=0, IPC unavailable, > 16 ranks): NCCL send/recv on a side stream + ncclAllReduce.
//
// NCCL is loaded with dlopen at smb200_comm_init time, so single-GPU users carry no NCCL dependency
// and the library loads on machines without it.
#include "common.cuh"
#include "cg_sr.cuh"
#include "halo.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace smb {

// ---- minimal NCCL surface (ABI-stable since 2.x) ---------------------------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { kNcclUint8 = 1, kNcclUint64 = 5, kNcclFloat32 = 7, kNcclFloat64 = 8 };
enum { kNcclSum = 0 };

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static smb200_status nccl_load() {
    if (g_nccl.lib) return SMB200_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) {
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    SMB_REQUIRE(lib, SMB200_ERR_NCCL, "NCCL not found (dlopen libnccl.so.2): %s", dlerror());
#define LOAD(field, sym)                                                                 \
    do {                                                                                 \
        *(void**)(&g_nccl.field) = dlsym(lib, sym);                                      \
        if (!g_nccl.field) { dlclose(lib); SMB_FAIL(SMB200_ERR_NCCL, "NCCL symbol %s missing", sym); } \
    } while (0)
    LOAD(GetUniqueId, "ncclGetUniqueId");
    LOAD(CommInitRank, "ncclCommInitRank");
    LOAD(CommDestroy, "ncclCommDestroy");
    LOAD(AllReduce, "ncclAllReduce");
    LOAD(AllGather, "ncclAllGather");
    LOAD(Broadcast, "ncclBroadcast");
    LOAD(Send, "ncclSend");
    LOAD(Recv, "ncclRecv");
    LOAD(GroupStart, "ncclGroupStart");
    LOAD(GroupEnd, "ncclGroupEnd");
    LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
    g_nccl.lib = lib;
    return SMB200_OK;
}

#define SMB_NCCL(expr)                                                                                   \
    do {                                                                                                 \
        ncclResult_t r__ = (expr);                                                                       \
        if (r__ != 0) SMB_FAIL(SMB200_ERR_NCCL, "NCCL error at %s:%d: %s", __FILE__, __LINE__, g_nccl.GetErrorString(r__)); \
    } while (0)

static int nccl_vtype(int vt) { return vt == SMB200_F64 ? kNcclFloat64 : kNcclFloat32; }

// for par.cu (SparseMatPar's gather of y)
smb200_status nccl_group_begin() { SMB_TRY(nccl_load()); SMB_NCCL(g_nccl.GroupStart()); return SMB200_OK; }
smb200_status nccl_group_end() { SMB_NCCL(g_nccl.GroupEnd()); return SMB200_OK; }
smb200_status nccl_bcast_bytes(smb200_ctx* ctx, void* buf, size_t bytes, int root, cudaStream_t stream) {
    SMB_REQUIRE(ctx->comm, SMB200_ERR_INVALID, "call smb200_comm_init first");
    SMB_NCCL(g_nccl.Broadcast(buf, buf, bytes, kNcclUint8, root, (ncclComm_t)ctx->comm, stream));
    return SMB200_OK;
}

template <class T>
__global__ void pack_kernel(const T* __restrict__ x, const uint64_t* __restrict__ idx, uint64_t n, T* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < n; k += stride) out[k] = x[idx[k]];
}

// ---- peer-memory kernels (protocol: halo.cuh) ------------------------------------------------------------
// Generic (non-ring) local plans: push | interior launch | wait + land the ghosts behind x's owned part | boundary rows.
template <class T>
__global__ void __launch_bounds__(256) halo_push_kernel(const HaloDev* __restrict__ h, const T* __restrict__ x) {
    const unsigned long long e = __ldcg(h->epoch) + 1ull;
    halo_push<T>(*h, x, e, blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, (uint64_t)gridDim.x * blockDim.x);
    __syncthreads();
    if (threadIdx.x == 0) halo_arrive(*h, e, gridDim.x);
}

template <class T>
__global__ void __launch_bounds__(256) halo_wait_kernel(const HaloDev* __restrict__ h, T* __restrict__ x) {
    const unsigned long long e = __ldcg(h->epoch) + 1ull;
    if (threadIdx.x == 0) halo_wait(*h, e);
    __syncthreads();
    const T* gx = (const T*)h->ghost + (e & 1ull) * h->ghost_stride;
    T* dst = x + h->g0;
    const uint64_t n = h->n_ghost, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < n; k += stride) dst[k] = __ldcg(gx + k);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(h->ctr + 1, 1u) == gridDim.x - 1) { h->ctr[1] = 0u; *h->epoch = e; }
    }
}

// All-reduce (sum) of `count` <= kArSlots doubles: lane q stores this rank's values into rank q's slot array and raises its
// flag, then waits for rank q's values here.  Thread 0 adds the world's values in rank order, so every rank computes
// the same bits.  One CTA of 32 threads; epochs alternate between two slot sets (a rank can be one all-reduce ahead).
__global__ void __launch_bounds__(32) p2p_allreduce_kernel(const ArDev* __restrict__ a, const double* in, double* out, int count,
                                                           int round_f32) {
    __shared__ double sv[kMaxPeers][kArSlots];
    __shared__ double mine[kArSlots], res[kArSlots];
    if (threadIdx.x < (unsigned)count) mine[threadIdx.x] = in[threadIdx.x];
    __syncwarp();
    ar_warp_allreduce(*a, mine, res, count, sv);
    if (threadIdx.x < (unsigned)count) out[threadIdx.x] = round_f32 ? (double)(float)res[threadIdx.x] : res[threadIdx.x];
}

}  // namespace smb

// Default number of SMs a ring interior leaves free for the exchange: 0 = the halo is awaited first.  Measured on B200
// (256^3 slab per GPU, profiles/dist_schedules_r01.log): at 2 GPUs leaving 8 SMs and overlapping ties with halo-first
// (0.149 vs 0.150 ms), fewer SMs lose (the boundary rows crawl through the few free slots); at 4 GPUs halo-first wins
// clearly (0.160 vs 0.195 ms).  SMB200_DIST_RESERVE_SMS / SMB200_DIST_OVERLAP keep the other schedules reachable.
namespace smb { constexpr int kDistReserveSms = 0; }

struct smb200_dist {
    smb200_ctx* ctx = nullptr;
    int vt = SMB200_F32, it = SMB200_U32;
    uint64_t n_global = 0, row_lo = 0, n_local = 0, n_ghost = 0;
    std::vector<uint64_t> bounds;
    smb200_crs* local = nullptr;                 // columns in local numbering [owned | ghosts]
    // receive side: ghosts are sorted by global id, hence grouped by owner
    std::vector<uint64_t> recv_count, recv_off;  // per peer, in elements, offset into the ghost region
    // send side
    std::vector<uint64_t> send_count, send_off;  // per peer, offset into send_idx / send_buf
    std::vector<int> send_contig;                // 1: the peer wants a contiguous run starting at send_first
    std::vector<uint64_t> send_first;
    uint64_t* send_idx = nullptr;                // device: local row ids to pack, grouped by peer
    void* send_buf = nullptr;                    // device: packed values
    uint64_t total_send = 0;
    // rows [int_begin, int_end) reference no ghost: multiplied while the exchange is in flight
    uint64_t int_begin = 0, int_end = 0;
    smb::SpmvPlan plan_int, plan_lo, plan_hi;
    // local column numbering: [0, n_local) owned, [g0, g0 + n_ghost) ghosts, g0 = n_local rounded up to 64 (so that a
    // bulk copy never straddles the two at an unaligned address)
    uint64_t g0 = 0;
    // ---- peer memory (halo.cuh) ----
    bool p2p = false;
    void* win = nullptr;                          // this rank's window: flags | all-reduce flags | all-reduce slots | ghost double buffer
    std::vector<void*> peer_base;                 // the peers' windows mapped here (nullptr: self / not mapped)
    void* misc = nullptr;                         // local words: halo epoch, all-reduce epoch, arrival counters, error
    smb::HaloDev* h_dev = nullptr;                // device copy (the small push / wait kernels read it)
    smb::HaloDev h_host;                          // host copy (the ring kernel takes it by value)
    smb::ArDev* ar_dev = nullptr;
    double* ar_scratch = nullptr;                 // [2 * kArSlots] operands of barriers / dots
    uint64_t rot = 0;                             // ring block order: first block without lower-ghost rows
    int n_nbr = 0;
    bool self_halo = false;                       // SMB200_DIST_SELF: world == 1 through the distributed kernel path
};

namespace smb {

static smb200_status dist_build_plans(smb200_dist* d) {
    smb200_crs* m = d->local;
    SMB_TRY(plan_build_range(m, d->plan_int, m->want_variant, m->want_lanes, m->want_flags, d->int_begin, d->int_end));
    SMB_TRY(plan_build_range(m, d->plan_lo, m->want_variant, m->want_lanes, m->want_flags, 0, d->int_begin));
    SMB_TRY(plan_build_range(m, d->plan_hi, m->want_variant, m->want_lanes, m->want_flags, d->int_end, d->n_local));
    if (!m->plan.built && m->n_rows) SMB_TRY(plan_build(m));          // whole local block: the ONE launch of a ring product
    // ring block order: start at the first block whose rows are all interior or upper-boundary rows; the blocks of the
    // lower-boundary rows wrap around to the end, next to the upper-boundary ones
    d->rot = 0;
    if (m->plan.built && m->plan.variant == SMB200_SPMV_RING && d->int_begin > 0 && m->plan.n_blocks > 1) {
        const size_t is = isize(m->it);
        std::vector<unsigned char> rows((m->plan.n_blocks + 1) * is);
        SMB_CUDA(cudaMemcpyAsync(rows.data(), m->plan.blk_rows, rows.size(), cudaMemcpyDeviceToHost, d->ctx->stream));
        SMB_CUDA(cudaStreamSynchronize(d->ctx->stream));
        auto row_of = [&](uint64_t b) { return is == 8 ? ((const uint64_t*)rows.data())[b] : (uint64_t)((const uint32_t*)rows.data())[b]; };
        uint64_t lo = 0, hi = m->plan.n_blocks;            // smallest b with blk_rows[b] >= int_begin
        while (lo < hi) { const uint64_t mid = (lo + hi) / 2; if (row_of(mid) < d->int_begin) lo = mid + 1; else hi = mid; }
        d->rot = lo < m->plan.n_blocks ? lo : 0;
    }
    return SMB200_OK;
}

static bool dist_exchanges(const smb200_dist* d) {
    if (d->self_halo) return true;
    return !(d->ctx->world == 1 || (d->n_ghost == 0 && d->total_send == 0));
}

// ---- peer-memory set-up: collective, called once from smb200_dist_create / smb200_dist_laplace ----------------------
struct P2pRecord {
    unsigned char handle[64];                 // cudaIpcMemHandle_t of the rank's window
    uint64_t want;                            // 1: the rank can take the peer-memory path
    uint64_t stride;                          // its ghost_stride (elements)
    uint64_t recv_off[kMaxPeers];             // where rank q's entries start inside its ghost buffer
    uint64_t recv_count[kMaxPeers];
};
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");

struct WinLayout { size_t flags, ar_flags, ar_vals, ghost, total; };
static WinLayout win_layout(int world, uint64_t stride, size_t es) {
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    WinLayout L;
    L.flags = 0;
    L.ar_flags = up(L.flags + (size_t)world * 8);
    L.ar_vals = up(L.ar_flags + 2 * (size_t)world * 8);
    L.ghost = up(L.ar_vals + 2 * (size_t)world * kArSlots * 8);
    L.total = up(L.ghost + 2 * (size_t)stride * es) + kPadBytes;
    return L;
}

static void dist_p2p_release(smb200_dist* d) {
    for (void*& pb : d->peer_base) if (pb) { cudaIpcCloseMemHandle(pb); pb = nullptr; }
    if (d->win) { cudaFree(d->win); d->win = nullptr; }
    if (d->misc) { cudaFree(d->misc); d->misc = nullptr; }
    if (d->h_dev) { cudaFree(d->h_dev); d->h_dev = nullptr; }
    if (d->ar_dev) { cudaFree(d->ar_dev); d->ar_dev = nullptr; }
    if (d->ar_scratch) { cudaFree(d->ar_scratch); d->ar_scratch = nullptr; }
    d->p2p = false;
}

static smb200_status dist_p2p_setup(smb200_dist* d) {
    smb200_ctx* ctx = d->ctx;
    const int world = ctx->world, me = ctx->rank;
    d->p2p = false;
    if (world == 1) {
        // SMB200_DIST_SELF=1 (measurements only): run the distributed kernel path with zero neighbours on one GPU, which
        // isolates what the halo code itself costs from what the coupling of the ranks costs
        const char* self = getenv("SMB200_DIST_SELF");
        if (!(self && self[0] == '1')) return SMB200_OK;
        SMB_CUDA(cudaMalloc(&d->misc, 256));
        SMB_CUDA(cudaMemset(d->misc, 0, 256));
        SMB_CUDA(cudaMalloc(&d->win, 4096));
        SMB_CUDA(cudaMemset(d->win, 0, 4096));
        HaloDev h;
        memset(&h, 0, sizeof h);
        unsigned char* misc = (unsigned char*)d->misc;
        h.epoch = (unsigned long long*)(misc + 0);
        h.ctr = (unsigned*)(misc + 16);
        h.error = (unsigned*)(misc + 24);
        h.wait_stats = (unsigned long long*)(misc + 32);
        h.flags = (unsigned long long*)d->win;
        h.ghost = (unsigned char*)d->win + 256;
        h.ghost_stride = 64;
        h.g0 = d->g0;
        h.timeout_ns = 1000000000ull;
        SMB_CUDA(cudaMalloc(&d->h_dev, sizeof h));
        SMB_CUDA(cudaMemcpy(d->h_dev, &h, sizeof h, cudaMemcpyHostToDevice));
        d->h_host = h;
        d->p2p = true;
        d->self_halo = true;
        return SMB200_OK;
    }
    ncclComm_t comm = (ncclComm_t)ctx->comm;
    const size_t es = vsize(d->vt);
    const char* env = getenv("SMB200_DIST_P2P");
    bool want = !(env && env[0] == '0') && world <= kMaxPeers;
    std::vector<int> nbr;
    for (int q = 0; q < world; ++q)
        if (q != me && (d->send_count[q] || d->recv_count[q])) nbr.push_back(q);
    if ((int)nbr.size() > kMaxNbr) want = false;
    d->n_nbr = (int)nbr.size();
    const uint64_t stride = ((d->n_ghost + 63) & ~(uint64_t)63) + 64;
    const WinLayout L = win_layout(world, stride, es);
    P2pRecord mine;
    memset(&mine, 0, sizeof mine);
    if (want) {
        if (cudaMalloc(&d->win, L.total) != cudaSuccess || cudaMalloc(&d->misc, 256) != cudaSuccess) { cudaGetLastError(); want = false; }
    }
    if (want) {
        cudaMemsetAsync(d->win, 0, L.total, ctx->stream);
        cudaMemsetAsync(d->misc, 0, 256, ctx->stream);
        cudaIpcMemHandle_t hnd;
        if (cudaIpcGetMemHandle(&hnd, d->win) != cudaSuccess) { cudaGetLastError(); want = false; }
        else memcpy(mine.handle, &hnd, 64);
    }
    mine.want = want ? 1 : 0;
    mine.stride = stride;
    for (int q = 0; q < world && q < kMaxPeers; ++q) { mine.recv_off[q] = d->recv_off[q]; mine.recv_count[q] = d->recv_count[q]; }
    // every rank takes part in the exchanges below whatever it wants: all ranks must end up on the same path
    unsigned char* d_rec = nullptr;
    SMB_CUDA(cudaMalloc(&d_rec, (size_t)world * sizeof(P2pRecord)));
    SMB_CUDA(cudaMemcpyAsync(d_rec + (size_t)me * sizeof(P2pRecord), &mine, sizeof mine, cudaMemcpyHostToDevice, ctx->stream));
    SMB_NCCL(g_nccl.AllGather(d_rec + (size_t)me * sizeof(P2pRecord), d_rec, sizeof(P2pRecord), kNcclUint8, comm, ctx->stream));
    std::vector<P2pRecord> recs(world);
    SMB_CUDA(cudaMemcpyAsync(recs.data(), d_rec, (size_t)world * sizeof(P2pRecord), cudaMemcpyDeviceToHost, ctx->stream));
    SMB_CUDA(cudaStreamSynchronize(ctx->stream));      // (also: every window is zeroed before any peer can learn its handle)
    cudaFree(d_rec);
    bool all = true;
    for (int q = 0; q < world; ++q) all = all && recs[q].want != 0;
    double fails = 0.0;
    if (all) {
        d->peer_base.assign(world, nullptr);
        for (int q = 0; q < world; ++q) {
            if (q == me) continue;
            cudaIpcMemHandle_t hnd;
            memcpy(&hnd, recs[q].handle, 64);
            if (cudaIpcOpenMemHandle(&d->peer_base[q], hnd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                d->peer_base[q] = nullptr;
                fails += 1.0;
            }
        }
        for (int q : nbr)
            if (recs[q].recv_count[me] != d->send_count[q]) fails += 1.0;     // the two sides' plans must agree
    }
    // agree on the outcome (a rank that could not map a peer takes everyone back to NCCL)
    double* d_f = nullptr;
    SMB_CUDA(cudaMalloc(&d_f, sizeof(double)));
    SMB_CUDA(cudaMemcpyAsync(d_f, &fails, sizeof fails, cudaMemcpyHostToDevice, ctx->stream));
    SMB_NCCL(g_nccl.AllReduce(d_f, d_f, 1, kNcclFloat64, kNcclSum, comm, ctx->stream));
    SMB_CUDA(cudaMemcpyAsync(&fails, d_f, sizeof fails, cudaMemcpyDeviceToHost, ctx->stream));
    SMB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_f);
    if (!all || fails != 0.0) { dist_p2p_release(d); return SMB200_OK; }

    const char* tenv = getenv("SMB200_P2P_TIMEOUT_MS");
    const unsigned long long timeout_ns = (unsigned long long)(tenv ? atof(tenv) : 30000.0) * 1000000ull;     // 30 s: host-side skew between ranks is not an error
    unsigned char* misc = (unsigned char*)d->misc;
    HaloDev h;
    memset(&h, 0, sizeof h);
    h.epoch = (unsigned long long*)(misc + 0);
    h.ctr = (unsigned*)(misc + 16);
    h.error = (unsigned*)(misc + 24);
    h.wait_stats = (unsigned long long*)(misc + 32);
    h.flags = (unsigned long long*)((unsigned char*)d->win + L.flags);
    h.ghost = (unsigned char*)d->win + L.ghost;
    h.ghost_stride = stride;
    h.n_ghost = d->n_ghost;
    h.g0 = d->g0;
    h.timeout_ns = timeout_ns;
    h.send_idx = (const unsigned long long*)d->send_idx;
    h.total_send = d->total_send;
    h.n_nbr = (int)nbr.size();
    for (int i = 0; i < h.n_nbr; ++i) {
        const int q = nbr[i];
        const WinLayout Lq = win_layout(world, recs[q].stride, es);
        unsigned char* base = (unsigned char*)d->peer_base[q];
        h.nbr_rank[i] = q;
        h.peer_flag[i] = (unsigned long long*)(base + Lq.flags) + me;
        h.peer_ghost[i] = base + Lq.ghost + (size_t)recs[q].recv_off[me] * es;
        h.peer_stride[i] = recs[q].stride;
        h.send_off[i] = d->send_off[q];
        h.send_count[i] = d->send_count[q];
        h.send_first[i] = d->send_first[q];
        h.send_contig[i] = d->send_contig[q];
    }
    ArDev a;
    memset(&a, 0, sizeof a);
    a.epoch = (unsigned long long*)(misc + 8);
    a.error = h.error;
    a.timeout_ns = timeout_ns;
    a.world = world;
    a.me = me;
    a.flags = (unsigned long long*)((unsigned char*)d->win + L.ar_flags);
    a.vals = (double*)((unsigned char*)d->win + L.ar_vals);
    for (int q = 0; q < world; ++q) {
        unsigned char* base = q == me ? (unsigned char*)d->win : (unsigned char*)d->peer_base[q];
        a.peer_flags[q] = (unsigned long long*)(base + L.ar_flags);      // the fixed part of the layout is the same on every rank
        a.peer_vals[q] = (double*)(base + L.ar_vals);
    }
    SMB_CUDA(cudaMalloc(&d->h_dev, sizeof h));
    SMB_CUDA(cudaMalloc(&d->ar_dev, sizeof a));
    SMB_CUDA(cudaMalloc(&d->ar_scratch, 2 * kArSlots * sizeof(double)));
    SMB_CUDA(cudaMemcpy(d->h_dev, &h, sizeof h, cudaMemcpyHostToDevice));
    d->h_host = h;
    SMB_CUDA(cudaMemcpy(d->ar_dev, &a, sizeof a, cudaMemcpyHostToDevice));
    SMB_CUDA(cudaMemset(d->ar_scratch, 0, 2 * kArSlots * sizeof(double)));
    d->p2p = true;
    return SMB200_OK;
}

// Sum `count` doubles over the ranks, in place or out of place, on the context stream.
static smb200_status dist_allreduce(smb200_dist* d, const double* in, double* out, int count, bool round_f32) {
    smb200_ctx* ctx = d->ctx;
    if (ctx->world == 1) {
        if (in != out) SMB_CUDA(cudaMemcpyAsync(out, in, count * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        return SMB200_OK;
    }
    if (d->p2p) {
        p2p_allreduce_kernel<<<1, 32, 0, ctx->stream>>>(d->ar_dev, in, out, count, round_f32 ? 1 : 0);
        count_launch();
        SMB_CUDA(cudaGetLastError());
        return SMB200_OK;
    }
    SMB_NCCL(g_nccl.AllReduce(in, out, count, kNcclFloat64, kNcclSum, (ncclComm_t)ctx->comm, ctx->stream));
    return SMB200_OK;
}

// Did a peer wait run into its timeout?  (host-synchronous; called where the host waits for the stream anyway)
static smb200_status dist_check_peers(smb200_dist* d) {
    if (!d->p2p) return SMB200_OK;
    unsigned err = 0;
    SMB_CUDA(cudaMemcpy(&err, (unsigned char*)d->misc + 24, sizeof err, cudaMemcpyDeviceToHost));
    SMB_REQUIRE(err == 0, SMB200_ERR_NCCL, "dist: a peer did not reach the same halo / all-reduce epoch within the timeout "
                "(SMB200_P2P_TIMEOUT_MS); the ranks have diverged or a peer died");
    return SMB200_OK;
}

// Refresh the ghost entries of x: pack on the main stream, send/recv on the side stream.
// On return the side stream has recorded ctx->ev_b when the ghosts are in place.
static smb200_status dist_exchange_begin(smb200_dist* d, void* x) {
    smb200_ctx* ctx = d->ctx;
    const int world = ctx->world, me = ctx->rank;
    if (!dist_exchanges(d)) return SMB200_OK;
    const size_t es = vsize(d->vt);
    for (int q = 0; q < world; ++q) {
        if (q == me || d->send_count[q] == 0 || d->send_contig[q]) continue;
        const unsigned g = (unsigned)std::min<uint64_t>((d->send_count[q] + 255) / 256, (uint64_t)ctx->sm_count * 4);
        if (d->vt == SMB200_F64) pack_kernel<double><<<g, 256, 0, ctx->stream>>>((const double*)x, d->send_idx + d->send_off[q], d->send_count[q], (double*)d->send_buf + d->send_off[q]);
        else pack_kernel<float><<<g, 256, 0, ctx->stream>>>((const float*)x, d->send_idx + d->send_off[q], d->send_count[q], (float*)d->send_buf + d->send_off[q]);
        count_launch();
    }
    SMB_CUDA(cudaEventRecord(ctx->ev_a, ctx->stream));
    SMB_CUDA(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_a, 0));
    ncclComm_t comm = (ncclComm_t)ctx->comm;
    SMB_NCCL(g_nccl.GroupStart());
    for (int q = 0; q < world; ++q) {
        if (q == me) continue;
        if (d->send_count[q]) {
            const char* src = d->send_contig[q] ? (const char*)x + d->send_first[q] * es : (const char*)d->send_buf + d->send_off[q] * es;
            SMB_NCCL(g_nccl.Send(src, d->send_count[q], nccl_vtype(d->vt), q, comm, ctx->aux_stream));
        }
        if (d->recv_count[q]) {
            char* dst = (char*)x + (d->g0 + d->recv_off[q]) * es;
            SMB_NCCL(g_nccl.Recv(dst, d->recv_count[q], nccl_vtype(d->vt), q, comm, ctx->aux_stream));
        }
    }
    SMB_NCCL(g_nccl.GroupEnd());
    return SMB200_OK;
}

static smb200_status dist_exchange_end(smb200_dist* d) {
    smb200_ctx* ctx = d->ctx;
    if (!dist_exchanges(d)) return SMB200_OK;
    SMB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_b, 0));
    return SMB200_OK;
}

// y = (A x)_local.  S == nullptr: plain product.  S != nullptr: CG kernel A with p.Ap partials in S.
static smb200_status dist_spmv_impl(smb200_dist* d, void* x, void* y, double* S) {
    smb200_crs* m = d->local;
    smb200_ctx* ctx = d->ctx;
    const bool exchange = dist_exchanges(d);
    if (exchange && d->p2p) {
        if (m->plan.built && m->plan.variant == SMB200_SPMV_RING) {
            // ONE launch: push, interior blocks, wait, boundary blocks (spmv_ring_kernel's halo path)
            g_halo.host = &d->h_host;
            g_halo.rot = d->rot;
            const smb200_status st = S ? spmv_launch_cg(m, m->plan, 0, d->n_local, x, y, x, S, 0, true)
                                       : spmv_launch_plan(m, m->plan, 0, d->n_local, x, y, nullptr, 0);
            g_halo = HaloLaunch();
            return st;
        }
        // any other local plan: push | interior rows | wait + land the ghosts behind x's owned part | boundary rows
        const unsigned gp = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((d->total_send + 255) / 256, (uint64_t)ctx->sm_count * 2));
        const unsigned gw = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((d->n_ghost + 1023) / 1024, (uint64_t)ctx->sm_count));
        if (d->vt == SMB200_F64) halo_push_kernel<double><<<gp, 256, 0, ctx->stream>>>(d->h_dev, (const double*)x);
        else halo_push_kernel<float><<<gp, 256, 0, ctx->stream>>>(d->h_dev, (const float*)x);
        count_launch();
        const bool no_int = d->int_end == d->int_begin;
        if (S) SMB_TRY(spmv_launch_cg(m, d->plan_int, d->int_begin, d->int_end, x, y, x, S, 0, true));
        else SMB_TRY(spmv_launch_plan(m, d->plan_int, d->int_begin, d->int_end, x, y, nullptr, 0));
        if (d->vt == SMB200_F64) halo_wait_kernel<double><<<gw, 256, 0, ctx->stream>>>(d->h_dev, (double*)x);
        else halo_wait_kernel<float><<<gw, 256, 0, ctx->stream>>>(d->h_dev, (float*)x);
        count_launch();
        SMB_CUDA(cudaGetLastError());
        if (S) {
            SMB_TRY(spmv_launch_cg(m, d->plan_lo, 0, d->int_begin, x, y, x, S, 1, no_int));
            SMB_TRY(spmv_launch_cg(m, d->plan_hi, d->int_end, d->n_local, x, y, x, S, 2, no_int && d->int_begin == 0));
        } else {
            SMB_TRY(spmv_launch_plan(m, d->plan_lo, 0, d->int_begin, x, y, nullptr, 0));
            SMB_TRY(spmv_launch_plan(m, d->plan_hi, d->int_end, d->n_local, x, y, nullptr, 0));
        }
        return SMB200_OK;
    }
    SMB_TRY(dist_exchange_begin(d, x));
    // Boundary rows: queued on the (high-priority) side stream right behind the halo receive, so they run as soon
    // as the ghosts have landed, in between the waves of the interior kernel, instead of after it.
    // Two schedules.  Non-persistent interior kernels (many small CTAs): the boundary launches ride the high-priority side
    // stream and slip in between the interior's waves.  Persistent ring interior (it holds every SM until it is done, so a
    // concurrent kernel only perturbs it — measured 298 us vs 269 us on 2 GPUs): boundary rows follow on the main stream.
    static const int overlap_env = []{ const char* e = getenv("SMB200_DIST_OVERLAP"); return e ? (e[0] == '0' ? 0 : 1) : -1; }();
    // SMs the persistent ring interior leaves to the NCCL send/recv kernel (set NCCL_MAX_P2P_NCHANNELS to the same number:
    // one CTA per channel); 0 = the ring takes every SM and the halo is awaited first
    static const int reserve_env = []{ const char* e = getenv("SMB200_DIST_RESERVE_SMS"); return e ? atoi(e) : kDistReserveSms; }();
    const bool overlap = overlap_env >= 0 ? overlap_env == 1 : d->plan_int.variant != SMB200_SPMV_RING;
    const bool ring_int = d->plan_int.variant == SMB200_SPMV_RING && d->int_end > d->int_begin;
    if (exchange && !overlap && !(ring_int && reserve_env > 0) && m->plan.built && m->plan.variant == SMB200_SPMV_RING) {
        // Persistent ring on every SM: a NCCL send/recv kernel that arrives meanwhile finds no SM to run on (measured: the
        // exchange then completes only after the interior, +60 us on ranks with two neighbours).
        // So: halo first (~20 us over NVLink), then ONE launch over all local rows.
        SMB_CUDA(cudaEventRecord(ctx->ev_b, ctx->aux_stream));
        SMB_TRY(dist_exchange_end(d));
        if (S) return spmv_launch_cg(m, m->plan, 0, d->n_local, x, y, x, S, 0, true);
        return spmv_launch_plan(m, m->plan, 0, d->n_local, x, y, nullptr, 0);
    }
    if (exchange && !overlap) {
        // Interior on the main stream while the exchange runs on the side stream (a ring interior leaves it a few SMs),
        // then the boundary rows once the ghosts are in.
        SMB_CUDA(cudaEventRecord(ctx->ev_b, ctx->aux_stream));
        g_ring_reserve_sms = ring_int ? reserve_env : 0;
        smb200_status sti;
        if (S) sti = spmv_launch_cg(m, d->plan_int, d->int_begin, d->int_end, x, y, x, S, 0, true);
        else sti = spmv_launch_plan(m, d->plan_int, d->int_begin, d->int_end, x, y, nullptr, 0);
        g_ring_reserve_sms = 0;
        SMB_TRY(sti);
        SMB_TRY(dist_exchange_end(d));
        if (S) {
            SMB_TRY(spmv_launch_cg(m, d->plan_lo, 0, d->int_begin, x, y, x, S, 1, d->int_end == d->int_begin));
            SMB_TRY(spmv_launch_cg(m, d->plan_hi, d->int_end, d->n_local, x, y, x, S, 2, d->int_end == d->int_begin && d->int_begin == 0));
        } else {
            SMB_TRY(spmv_launch_plan(m, d->plan_lo, 0, d->int_begin, x, y, nullptr, 0));
            SMB_TRY(spmv_launch_plan(m, d->plan_hi, d->int_end, d->n_local, x, y, nullptr, 0));
        }
        return SMB200_OK;
    }
    if (exchange) {
        if (S && ctx->red_cap_aux < d->plan_lo.n_blocks + d->plan_hi.n_blocks + 16) {
            SMB_CUDA(cudaStreamSynchronize(ctx->aux_stream));
            if (ctx->red_partials_aux) cudaFree(ctx->red_partials_aux);
            ctx->red_partials_aux = nullptr;
            ctx->red_cap_aux = 2 * (d->plan_lo.n_blocks + d->plan_hi.n_blocks) + 1024;
            SMB_CUDA(cudaMalloc(&ctx->red_partials_aux, ctx->red_cap_aux * sizeof(double)));
        }
        g_redirect.stream = ctx->aux_stream;
        g_redirect.partials = S ? ctx->red_partials_aux : nullptr;
        // beside a ring interior the boundary launches only get the slots it leaves free: two CTAs per reserved SM
        if (ring_int && reserve_env > 0) g_ring_grid_cap = 2 * reserve_env;
    }
    smb200_status st = SMB200_OK;
    if (S) {
        // rr <- rr_new is rolled once per product: by the interior launch, or by a boundary one if there is no interior
        const bool no_int = d->int_end == d->int_begin;
        st = spmv_launch_cg(m, d->plan_lo, 0, d->int_begin, x, y, x, S, 1, no_int);
        if (st == SMB200_OK) st = spmv_launch_cg(m, d->plan_hi, d->int_end, d->n_local, x, y, x, S, 2, no_int && d->int_begin == 0);
    } else {
        st = spmv_launch_plan(m, d->plan_lo, 0, d->int_begin, x, y, nullptr, 0);
        if (st == SMB200_OK) st = spmv_launch_plan(m, d->plan_hi, d->int_end, d->n_local, x, y, nullptr, 0);
    }
    g_redirect = LaunchRedirect();
    g_ring_grid_cap = 0;
    SMB_TRY(st);
    if (exchange) SMB_CUDA(cudaEventRecord(ctx->ev_b, ctx->aux_stream));
    g_ring_reserve_sms = (exchange && ring_int) ? reserve_env : 0;
    if (S) st = spmv_launch_cg(m, d->plan_int, d->int_begin, d->int_end, x, y, x, S, 0, true);
    else st = spmv_launch_plan(m, d->plan_int, d->int_begin, d->int_end, x, y, nullptr, 0);
    g_ring_reserve_sms = 0;
    g_redirect = LaunchRedirect();
    SMB_TRY(st);
    SMB_TRY(dist_exchange_end(d));
    return SMB200_OK;
}

// Exchange "who needs what": recv_count is known locally; learn send lists from the peers.
static smb200_status dist_setup_exchange(smb200_dist* d, const std::vector<uint64_t>& ghosts) {
    smb200_ctx* ctx = d->ctx;
    const int world = ctx->world, me = ctx->rank;
    d->recv_off.assign(world, 0);
    d->send_count.assign(world, 0);
    d->send_off.assign(world, 0);
    d->send_contig.assign(world, 0);
    d->send_first.assign(world, 0);
    uint64_t run = 0;
    for (int q = 0; q < world; ++q) { d->recv_off[q] = run; run += d->recv_count[q]; }
    if (world == 1) return SMB200_OK;
    ncclComm_t comm = (ncclComm_t)ctx->comm;
    // all-gather the world x world matrix of counts
    uint64_t* d_counts = nullptr;
    SMB_CUDA(cudaMalloc(&d_counts, (size_t)world * world * sizeof(uint64_t)));
    SMB_CUDA(cudaMemcpyAsync(d_counts + (size_t)me * world, d->recv_count.data(), world * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    SMB_NCCL(g_nccl.AllGather(d_counts + (size_t)me * world, d_counts, world, kNcclUint64, comm, ctx->stream));
    std::vector<uint64_t> counts((size_t)world * world);
    SMB_CUDA(cudaMemcpyAsync(counts.data(), d_counts, counts.size() * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    SMB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_counts);
    run = 0;
    for (int q = 0; q < world; ++q) {
        d->send_count[q] = (q == me) ? 0 : counts[(size_t)q * world + me];   // what q receives from me
        d->send_off[q] = run;
        run += d->send_count[q];
    }
    d->total_send = run;
    // ship the ghost id lists to their owners
    uint64_t *d_ghosts = nullptr, *d_want = nullptr;
    SMB_CUDA(cudaMalloc(&d_ghosts, (ghosts.size() + 1) * sizeof(uint64_t)));
    SMB_CUDA(cudaMalloc(&d_want, (d->total_send + 1) * sizeof(uint64_t)));
    if (!ghosts.empty()) SMB_CUDA(cudaMemcpyAsync(d_ghosts, ghosts.data(), ghosts.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    SMB_NCCL(g_nccl.GroupStart());
    for (int q = 0; q < world; ++q) {
        if (q == me) continue;
        if (d->recv_count[q]) SMB_NCCL(g_nccl.Send(d_ghosts + d->recv_off[q], d->recv_count[q], kNcclUint64, q, comm, ctx->stream));
        if (d->send_count[q]) SMB_NCCL(g_nccl.Recv(d_want + d->send_off[q], d->send_count[q], kNcclUint64, q, comm, ctx->stream));
    }
    SMB_NCCL(g_nccl.GroupEnd());
    std::vector<uint64_t> want(d->total_send);
    if (d->total_send) SMB_CUDA(cudaMemcpyAsync(want.data(), d_want, d->total_send * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    SMB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_ghosts);
    for (uint64_t k = 0; k < d->total_send; ++k) {
        if (want[k] < d->row_lo || want[k] >= d->row_lo + d->n_local) {
            cudaFree(d_want);
            SMB_FAIL(SMB200_ERR_INVALID, "dist: a peer asked for row %llu which this rank does not own", (unsigned long long)want[k]);
        }
        want[k] -= d->row_lo;
    }
    for (int q = 0; q < world; ++q) {
        const uint64_t c = d->send_count[q], o = d->send_off[q];
        bool contig = c > 0;
        for (uint64_t k = 1; k < c && contig; ++k) contig = want[o + k] == want[o + k - 1] + 1;
        d->send_contig[q] = contig ? 1 : 0;
        d->send_first[q] = c ? want[o] : 0;
    }
    d->send_idx = d_want;
    if (d->total_send) SMB_CUDA(cudaMemcpyAsync(d->send_idx, want.data(), d->total_send * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    SMB_TRY(dev_alloc(&d->send_buf, (d->total_send + 1) * vsize(d->vt)));
    SMB_CUDA(cudaStreamSynchronize(ctx->stream));
    return SMB200_OK;
}

}  // namespace smb

using namespace smb;

extern "C" {

smb200_status smb200_comm_unique_id(void* out128) {
    SMB_REQUIRE(out128, SMB200_ERR_INVALID, "comm_unique_id: NULL argument");
    SMB_TRY(nccl_load());
    ncclUniqueId id;
    SMB_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, sizeof id);
    return SMB200_OK;
}

smb200_status smb200_comm_init(smb200_ctx* ctx, int32_t rank, int32_t world, const void* uid128) {
    SMB_REQUIRE(ctx && (uid128 || world == 1), SMB200_ERR_INVALID, "comm_init: NULL argument");
    SMB_REQUIRE(world >= 1 && rank >= 0 && rank < world, SMB200_ERR_INVALID, "comm_init: bad rank %d / world %d", rank, world);
    SMB_REQUIRE(!ctx->comm, SMB200_ERR_INVALID, "comm_init: communicator already initialised");
    ctx->rank = rank;
    ctx->world = world;
    if (world == 1) return SMB200_OK;
    SMB_TRY(nccl_load());
    SMB_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, uid128, sizeof id);
    ncclComm_t comm = nullptr;
    SMB_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
    ctx->comm = comm;
    return SMB200_OK;
}

smb200_status smb200_comm_destroy(smb200_ctx* ctx) {
    if (!ctx || !ctx->comm) return SMB200_OK;
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->aux_stream);
    g_nccl.CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr;
    ctx->world = 1;
    ctx->rank = 0;
    return SMB200_OK;
}

smb200_status smb200_dist_free(smb200_dist* d) {
    if (!d) return SMB200_OK;
    cudaSetDevice(d->ctx->device);
    cudaStreamSynchronize(d->ctx->stream);
    cudaStreamSynchronize(d->ctx->aux_stream);
    plan_free(d->plan_int);
    plan_free(d->plan_lo);
    plan_free(d->plan_hi);
    dist_p2p_release(d);
    if (d->send_idx) cudaFree(d->send_idx);
    if (d->send_buf) cudaFree(d->send_buf);
    if (d->local) smb200_crs_free(d->local);
    smb200_ctx* ctx = d->ctx;
    delete d;
    ctx_release(ctx);
    return SMB200_OK;
}

smb200_status smb200_dist_create(smb200_ctx* ctx, smb200_vtype vt, smb200_itype it, uint64_t n_global, const uint64_t* bounds,
                                 uint64_t nnz_local, const void* values, const void* columns_global,
                                 const void* offset_rows_local, smb200_dist** out) {
    SMB_REQUIRE(ctx && bounds && out, SMB200_ERR_INVALID, "dist_create: NULL argument");
    SMB_REQUIRE(ctx->world == 1 || ctx->comm, SMB200_ERR_INVALID, "dist_create: call smb200_comm_init first");
    const int world = ctx->world, me = ctx->rank;
    SMB_REQUIRE(bounds[world] == n_global && bounds[0] == 0, SMB200_ERR_INVALID, "dist_create: bounds must span [0, n_global]");
    *out = nullptr;
    smb200_dist* d = new smb200_dist();
    d->ctx = ctx; d->vt = vt; d->it = it; d->n_global = n_global;
    ctx_retain(ctx);
    d->bounds.assign(bounds, bounds + world + 1);
    d->row_lo = bounds[me];
    d->n_local = bounds[me + 1] - bounds[me];
    // ghost plan on the host (partition.cpp)
    uint64_t n_ghost = 0;
    d->recv_count.assign(world, 0);
    smb200_status s = smb200_ghost_plan(it, nnz_local, columns_global, (uint32_t)world, (uint32_t)me, bounds, nullptr, nullptr, &n_ghost, d->recv_count.data());
    std::vector<uint64_t> ghosts(n_ghost);
    std::vector<unsigned char> cols_local(nnz_local * isize(it));
    if (s == SMB200_OK)
        s = smb200_ghost_plan(it, nnz_local, columns_global, (uint32_t)world, (uint32_t)me, bounds, cols_local.data(), ghosts.data(), &n_ghost, d->recv_count.data());
    if (s != SMB200_OK) { smb200_dist_free(d); return s; }
    d->n_ghost = n_ghost;
    d->recv_count[me] = 0;
    // ghosts start at a multiple of 64 columns: [owned | pad | ghosts]
    d->g0 = (d->n_local + 63) & ~(uint64_t)63;
    if (d->g0 != d->n_local) {
        const uint64_t shift = d->g0 - d->n_local;
        if (it == SMB200_U64) { uint64_t* c = (uint64_t*)cols_local.data(); for (uint64_t k = 0; k < nnz_local; ++k) if (c[k] >= d->n_local) c[k] += shift; }
        else {
            if (d->g0 + n_ghost > 0xFFFFFFFFull) { smb200_dist_free(d); SMB_FAIL(SMB200_ERR_UNSUPPORTED, "dist_create: local block too large for u32 columns"); }
            uint32_t* c = (uint32_t*)cols_local.data();
            for (uint64_t k = 0; k < nnz_local; ++k) if (c[k] >= d->n_local) c[k] += (uint32_t)shift;
        }
    }
    s = smb200_crs_upload(ctx, vt, it, d->n_local, d->g0 + n_ghost, nnz_local, values, cols_local.data(), offset_rows_local, &d->local);
    if (s != SMB200_OK) { smb200_dist_free(d); return s; }
    d->local->x_extra = d->g0 + n_ghost - d->n_local;
    // interior = longest run of rows without ghost references
    {
        uint64_t best_b = 0, best_e = 0, run_b = 0;
        auto off = [&](uint64_t r) { return it == SMB200_U64 ? ((const uint64_t*)offset_rows_local)[r] : (uint64_t)((const uint32_t*)offset_rows_local)[r]; };
        auto col = [&](uint64_t k) { return it == SMB200_U64 ? ((const uint64_t*)cols_local.data())[k] : (uint64_t)((const uint32_t*)cols_local.data())[k]; };
        for (uint64_t r = 0; r < d->n_local; ++r) {
            bool ghost = false;
            for (uint64_t k = off(r); k < off(r + 1) && !ghost; ++k) ghost = col(k) >= d->g0;
            if (ghost) { if (r - run_b > best_e - best_b) { best_b = run_b; best_e = r; } run_b = r + 1; }
        }
        if (d->n_local - run_b > best_e - best_b) { best_b = run_b; best_e = d->n_local; }
        d->int_begin = best_b; d->int_end = best_e;
    }
    s = dist_setup_exchange(d, ghosts);
    if (s == SMB200_OK) s = dist_build_plans(d);
    if (s == SMB200_OK) s = dist_p2p_setup(d);
    if (s != SMB200_OK) { smb200_dist_free(d); return s; }
    *out = d;
    return SMB200_OK;
}

smb200_status smb200_dist_laplace(smb200_ctx* ctx, smb200_vtype vt, smb200_itype it, uint64_t nx, uint64_t ny, uint64_t nz,
                                  smb200_dist** out) {
    SMB_REQUIRE(ctx && out, SMB200_ERR_INVALID, "dist_laplace: NULL argument");
    SMB_REQUIRE(ctx->world == 1 || ctx->comm, SMB200_ERR_INVALID, "dist_laplace: call smb200_comm_init first");
    SMB_REQUIRE(nx && ny && nz, SMB200_ERR_INVALID, "dist_laplace: empty grid");
    const int world = ctx->world, me = ctx->rank;
    const uint64_t plane = nx * ny, N = plane * nz;
    SMB_REQUIRE(nz >= (uint64_t)world, SMB200_ERR_INVALID, "dist_laplace: fewer z-planes (%llu) than ranks (%d)", (unsigned long long)nz, world);
    *out = nullptr;
    smb200_dist* d = new smb200_dist();
    d->ctx = ctx; d->vt = vt; d->it = it; d->n_global = N;
    ctx_retain(ctx);
    d->bounds.resize(world + 1);
    // whole z-planes per rank, as even as possible
    for (int q = 0; q <= world; ++q) d->bounds[q] = (nz * (uint64_t)q / (uint64_t)world) * plane;
    d->row_lo = d->bounds[me];
    d->n_local = d->bounds[me + 1] - d->bounds[me];
    const uint64_t n_lo = (nz > 1 && me > 0) ? plane : 0, n_hi = (nz > 1 && me + 1 < world) ? plane : 0;
    d->n_ghost = n_lo + n_hi;
    d->g0 = (d->n_local + 63) & ~(uint64_t)63;
    if (it != SMB200_U64 && d->g0 + d->n_ghost > 0xFFFFFFFFull) { smb200_dist_free(d); SMB_FAIL(SMB200_ERR_UNSUPPORTED, "dist_laplace: local block too large for u32 columns"); }
    smb200_status s = gen_laplace_block(ctx, vt, it, nx, ny, nz, d->row_lo, d->row_lo + d->n_local, 1, n_lo, d->g0 + d->n_ghost, &d->local, d->g0);
    if (s != SMB200_OK) { smb200_dist_free(d); return s; }
    d->local->x_extra = d->g0 + d->n_ghost - d->n_local;
    d->recv_count.assign(world, 0);
    d->recv_off.assign(world, 0);
    d->send_count.assign(world, 0);
    d->send_off.assign(world, 0);
    d->send_contig.assign(world, 0);
    d->send_first.assign(world, 0);
    if (n_lo) { d->recv_count[me - 1] = plane; d->recv_off[me - 1] = 0; d->send_count[me - 1] = plane; d->send_contig[me - 1] = 1; d->send_first[me - 1] = 0; }
    if (n_hi) { d->recv_count[me + 1] = plane; d->recv_off[me + 1] = n_lo; d->send_count[me + 1] = plane; d->send_contig[me + 1] = 1; d->send_first[me + 1] = d->n_local - plane; }
    d->total_send = n_lo + n_hi;
    d->int_begin = std::min<uint64_t>(n_lo, d->n_local);
    d->int_end = std::max<uint64_t>(d->int_begin, d->n_local - std::min<uint64_t>(n_hi, d->n_local));
    s = dist_build_plans(d);
    if (s == SMB200_OK) s = dist_p2p_setup(d);
    if (s != SMB200_OK) { smb200_dist_free(d); return s; }
    *out = d;
    return SMB200_OK;
}

smb200_status smb200_dist_dims(const smb200_dist* d, uint64_t* out4) {
    SMB_REQUIRE(d && out4, SMB200_ERR_INVALID, "dist_dims: NULL argument");
    out4[0] = d->n_local; out4[1] = d->n_ghost; out4[2] = d->local->nnz; out4[3] = d->row_lo;
    return SMB200_OK;
}

smb200_status smb200_dist_local(smb200_dist* d, smb200_crs** out) {
    SMB_REQUIRE(d && out, SMB200_ERR_INVALID, "dist_local: NULL argument");
    *out = d->local;
    return SMB200_OK;
}

smb200_status smb200_dist_vec_create(smb200_dist* d, smb200_vec** out) {
    SMB_REQUIRE(d && out, SMB200_ERR_INVALID, "dist_vec_create: NULL argument");
    return vec_create_cap(d->ctx, d->vt, d->n_local, d->g0 + d->n_ghost, out);
}

smb200_status smb200_dist_spmv(smb200_dist* d, smb200_vec* x, smb200_vec* y) {
    SMB_REQUIRE(d && x && y, SMB200_ERR_INVALID, "dist_spmv: NULL argument");
    SMB_REQUIRE(x->vt == d->vt && y->vt == d->vt, SMB200_ERR_INVALID, "dist_spmv: value types differ");
    SMB_REQUIRE(x->n >= d->n_local && x->cap >= d->g0 + d->n_ghost, SMB200_ERR_DIM,
                "Dimension mismatch: x must come from smb200_dist_vec_create (room for %llu ghosts)", (unsigned long long)d->n_ghost);
    SMB_REQUIRE(y->n >= d->n_local, SMB200_ERR_DIM, "Dimension mismatch");
    SMB_REQUIRE(x->d != y->d, SMB200_ERR_INVALID, "dist_spmv: x and y alias");
    return dist_spmv_impl(d, x->d, y->d, nullptr);
}

smb200_status smb200_dist_dot(smb200_dist* d, const smb200_vec* x, const smb200_vec* y, double* out) {
    SMB_REQUIRE(d && x && y && out, SMB200_ERR_INVALID, "dist_dot: NULL argument");
    SMB_REQUIRE(x->vt == d->vt && y->vt == d->vt, SMB200_ERR_INVALID, "dist_dot: value types differ");
    smb200_ctx* ctx = d->ctx;
    const uint64_t n = std::min(x->n, y->n);
    SMB_TRY(dot_launch(ctx, d->vt, x->d, y->d, n, 2));
    SMB_TRY(dist_allreduce(d, ctx->red_result + 2, ctx->red_result + 2, 1, d->vt == SMB200_F32));
    SMB_TRY(fetch_result(ctx, 2, out));
    if (d->vt == SMB200_F32) *out = (double)(float)*out;
    return dist_check_peers(d);
}

smb200_status smb200_dist_barrier(smb200_dist* d) {
    SMB_REQUIRE(d, SMB200_ERR_INVALID, "dist_barrier: NULL argument");
    smb200_ctx* ctx = d->ctx;
    if (ctx->world == 1) return SMB200_OK;
    if (d->p2p) return dist_allreduce(d, d->ar_scratch, d->ar_scratch + kArSlots, 1, false);
    SMB_TRY(ensure_reduction_scratch(ctx, 16));
    SMB_NCCL(g_nccl.AllReduce(ctx->red_partials, ctx->red_partials + 8, 1, kNcclFloat64, kNcclSum, (ncclComm_t)ctx->comm, ctx->stream));
    return SMB200_OK;
}

smb200_status smb200_dist_info(smb200_dist* d, uint64_t* out6) {
    SMB_REQUIRE(d && out6, SMB200_ERR_INVALID, "dist_info: NULL argument");
    out6[0] = d->p2p ? 1 : 0;
    out6[1] = (uint64_t)d->n_nbr;
    out6[2] = out6[3] = out6[4] = out6[5] = 0;
    if (d->p2p) {
        SMB_CUDA(cudaStreamSynchronize(d->ctx->stream));
        unsigned long long words[6] = {0, 0, 0, 0, 0, 0};            // misc: epoch, all-reduce epoch, counters, error, wait ns, waits
        SMB_CUDA(cudaMemcpy(words, d->misc, sizeof words, cudaMemcpyDeviceToHost));
        out6[2] = words[0];
        out6[3] = (uint64_t)(words[3] & 0xffffffffull);
        out6[4] = words[4];
        out6[5] = words[5];
    }
    return SMB200_OK;
}

}  // extern "C"

// sr: the single-reduction rearrangement of the loop (cg_sr.cuh) instead of the reference's
static smb200_status dist_cg_impl(smb200_dist* d, const smb200_vec* b, smb200_vec* x, double tol, int32_t relative,
                                  uint64_t iter_max, smb200_cg_stats* stats, bool sr) {
    SMB_REQUIRE(d && b && x, SMB200_ERR_INVALID, "dist_cg_solve: NULL argument");
    SMB_REQUIRE(b->vt == d->vt && x->vt == d->vt, SMB200_ERR_INVALID, "dist_cg_solve: value types differ");
    SMB_REQUIRE(d->n_local == b->n && d->n_local == x->n, SMB200_ERR_SIZE_MISMATCH, "Matrix and vector size mismatch");
    SMB_REQUIRE(x->cap >= d->g0 + d->n_ghost, SMB200_ERR_DIM, "Dimension mismatch: x must come from smb200_dist_vec_create");
    smb200_ctx* ctx = d->ctx;
    smb200_crs* a = d->local;
    CgWork& w = a->cg;
    const uint64_t n = d->n_local;
    const uint64_t launches0 = g_launches;
    const bool multi = ctx->world > 1;
    const bool f32 = d->vt == SMB200_F32;
    if (stats) memset(stats, 0, sizeof *stats);
    SMB_CUDA(cudaSetDevice(ctx->device));
    SMB_TRY(cg_prepare(ctx, w, d->vt, n, d->g0 + d->n_ghost, iter_max));
    if (sr) SMB_TRY(cgsr_prepare(ctx, w, d->vt));
    double* S = w.scalars;
    enum { S_RR = 0, S_PAP = 1, S_RR_NEW = 4, S_THRESH = 5, S_ITER = 6, S_DONE = 7, S_RR_LOCAL = 8, S_RES2 = 9, S_COUNT = 16 };   // cg.cu

    double threshold = tol;
    if (relative) {
        double bb = 0.0;
        SMB_TRY(smb200_dist_dot(d, b, b, &bb));
        threshold = tol * sqrt(bb);
    }
    struct Events {                       // destroyed on every return path
        cudaEvent_t ev0 = nullptr, ev1 = nullptr, poll[2] = {nullptr, nullptr};
        ~Events() { for (cudaEvent_t e : {ev0, ev1, poll[0], poll[1]}) if (e) cudaEventDestroy(e); }
    } evs;
    SMB_CUDA(cudaEventCreate(&evs.ev0));
    SMB_CUDA(cudaEventCreate(&evs.ev1));
    SMB_CUDA(cudaEventCreateWithFlags(&evs.poll[0], cudaEventDisableTiming));
    SMB_CUDA(cudaEventCreateWithFlags(&evs.poll[1], cudaEventDisableTiming));
    cudaEvent_t ev0 = evs.ev0, ev1 = evs.ev1;
    cudaEvent_t* poll_ev = evs.poll;
    SMB_CUDA(cudaEventRecord(ev0, ctx->stream));
    double init[S_COUNT] = {0};
    init[S_THRESH] = threshold;
    memcpy(w.scalars_host + 3 * S_COUNT, init, sizeof init);
    SMB_CUDA(cudaMemcpyAsync(S, w.scalars_host + 3 * S_COUNT, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
    // r = b - A x ; p = r ; rr = r.r
    SMB_TRY(dist_spmv_impl(d, x->d, w.ap, nullptr));
    SMB_TRY(cg_init_launch(ctx, w, d->vt, b->d, n, nullptr, sr ? S_RR_NEW : -1));
    if (multi && !sr) SMB_TRY(dist_allreduce(d, S + S_RR_LOCAL, S + S_RR_NEW, 1, f32));

    const char* benv = getenv("SMB200_CG_BATCH");
    int batch = benv ? atoi(benv) : 8;
    if (batch < 1) batch = 1;
    const char* genv = getenv("SMB200_CG_GRAPH");
    const bool use_graph = !(genv && genv[0] == '0');
    smb200_status st = SMB200_OK;
    uint64_t launched = 0, rounds = 0;
    bool finished = false;
    // One iteration: halo exchange + SpMV with p.Ap fused (three launches) | all-reduce | r update + r.r | all-reduce | x, p update.
    // Every rank runs the same sequence: the stop flag derives from all-reduced values, hence is identical everywhere, and
    // iterations past it exit early on the device.
    // With the ring kernel as the local product (ONE launch, one p.Ap slot) the two all-reduces ride inside the kernels that
    // produce their operands — the fused dot's finalize kernel and the last CTA of the r update — so a distributed
    // iteration is the same four launches as a single-GPU one.  Other local plans keep the two 1-CTA all-reduce kernels.
    const bool fused_ar = multi && d->p2p && dist_exchanges(d) && a->plan.built && a->plan.variant == SMB200_SPMV_RING;
    auto iteration = [&]() -> smb200_status {
        if (multi && !fused_ar && cudaMemsetAsync(S + S_PAP, 0, 3 * sizeof(double), ctx->stream) != cudaSuccess) { set_error("dist_cg_solve: memset failed"); return SMB200_ERR_CUDA; }
        g_dot_ar = fused_ar ? d->ar_dev : nullptr;
        const smb200_status sp = dist_spmv_impl(d, w.p, w.ap, S);
        g_dot_ar = nullptr;
        SMB_TRY(sp);
        if (multi && !fused_ar) SMB_TRY(dist_allreduce(d, S + S_PAP, S + S_PAP, 3, false));
        SMB_TRY(cg_r_launch(ctx, w, d->vt, n, nullptr, fused_ar ? d->ar_dev : nullptr));
        if (multi && !fused_ar) SMB_TRY(dist_allreduce(d, S + S_RR_LOCAL, S + S_RR_NEW, 1, f32));     // r.r is a T in the reference: rounded like the single-GPU path
        return cg_xp_launch(ctx, w, d->vt, x->d, n);
    };
    // Single-reduction iteration (cg_sr.cuh): vector update with the rank's r.r | halo exchange + w = A r with w.r fused |
    // ONE all-reduce of the pair + the scalar step — inside the fused dot's finalize kernel when the product is one launch
    // (three launches per iteration), else in a one-warp kernel behind it (NCCL fallback: ncclAllReduce in between).
    const bool sr_fused = dist_exchanges(d) ? multi && d->p2p && a->plan.built && a->plan.variant == SMB200_SPMV_RING
                                            : (!multi || d->p2p) && d->int_begin == 0 && d->int_end == d->n_local;   // no halo: plan_int is the product
    auto sr_product = [&]() -> smb200_status {
        if (!sr_fused && cudaMemsetAsync(S + S_PAP, 0, 3 * sizeof(double), ctx->stream) != cudaSuccess) { set_error("dist_cg_solve: memset failed"); return SMB200_ERR_CUDA; }
        g_cgsr = CgSrLaunch{sr_fused ? S : nullptr, w.history, w.hist_cap, true};
        g_dot_ar = sr_fused && multi ? d->ar_dev : nullptr;
        const smb200_status sp = dist_spmv_impl(d, w.r, w.ap, S);
        g_dot_ar = nullptr;
        g_cgsr = CgSrLaunch();
        SMB_TRY(sp);
        if (sr_fused) return SMB200_OK;
        if (multi && !d->p2p) SMB_TRY(dist_allreduce(d, S + S_PAP, S + S_PAP, 4, false));      // S_PAP..S_PAP+2, S_RR_NEW: contiguous
        return cgsr_scalar_launch(ctx, w, d->vt, multi && d->p2p ? d->ar_dev : nullptr);
    };
    auto sr_iteration = [&]() -> smb200_status {
        SMB_TRY(cgsr_update_launch(ctx, w, d->vt, x->d, n));
        return sr_product();
    };
    if (sr) SMB_TRY(sr_product());          // w0 = A r0, alpha0 = r0.r0 / w0.r0, beta0 = 0
    auto step = [&]() -> smb200_status { return sr ? sr_iteration() : iteration(); };
    // the first iteration runs eagerly: it performs every lazy allocation / connection set-up outside of stream capture
    if (iter_max > 0) { st = step(); launched = 1; }
    while (st == SMB200_OK && !finished) {
        const int slot = (int)(rounds & 1);
        cudaError_t e = cudaMemcpyAsync(w.scalars_host + slot * S_COUNT, S, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaEventRecord(poll_ev[slot], ctx->stream);
        if (e != cudaSuccess) { set_error("dist_cg_solve: %s", cudaGetErrorString(e)); st = SMB200_ERR_CUDA; break; }
        const uint64_t nb = std::min<uint64_t>(iter_max - launched, (uint64_t)batch);
        if (nb == (uint64_t)batch && use_graph) {
            // a batch of iterations — kernels, NCCL send/recv and all-reduces on both streams — replayed from one CUDA graph,
            // so the host enqueues one node per batch instead of ~20 calls per iteration
            if (!w.graph || w.graph_batch != batch || w.graph_x != x->d || w.graph_partials != ctx->red_partials || w.graph_plan != (const void*)d ||
                w.graph_kind != (sr ? 1 : 0)) {
                if (w.graph) { cudaGraphExecDestroy(w.graph); w.graph = nullptr; }
                cudaGraph_t graph = nullptr;
                e = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal);
                if (e == cudaSuccess) {
                    for (int k = 0; k < batch && st == SMB200_OK; ++k) st = step();
                    cudaError_t e2 = cudaStreamEndCapture(ctx->stream, &graph);
                    if (st == SMB200_OK && e2 != cudaSuccess) e = e2;
                }
                if (st == SMB200_OK && e == cudaSuccess) e = cudaGraphInstantiate(&w.graph, graph, 0);
                if (graph) cudaGraphDestroy(graph);
                if (e != cudaSuccess && st == SMB200_OK) { set_error("dist_cg_solve: graph capture failed: %s", cudaGetErrorString(e)); st = SMB200_ERR_CUDA; }
                if (st != SMB200_OK) break;
                w.graph_batch = batch; w.graph_x = x->d; w.graph_partials = ctx->red_partials; w.graph_plan = (const void*)d;
                w.graph_kind = sr ? 1 : 0;
            }
            e = cudaGraphLaunch(w.graph, ctx->stream);
            if (e != cudaSuccess) { set_error("dist_cg_solve: graph launch failed: %s", cudaGetErrorString(e)); st = SMB200_ERR_CUDA; break; }
        } else {
            for (uint64_t k = 0; k < nb && st == SMB200_OK; ++k) st = step();
        }
        if (st != SMB200_OK) break;
        launched += nb;
        e = cudaEventSynchronize(poll_ev[slot]);
        if (e != cudaSuccess) { set_error("dist_cg_solve: %s", cudaGetErrorString(e)); st = SMB200_ERR_CUDA; break; }
        if (w.scalars_host[slot * S_COUNT + S_DONE] != 0.0 || nb == 0) finished = true;
        ++rounds;
    }
    if (st == SMB200_OK) {
        cudaError_t e = cudaEventRecord(ev1, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(w.scalars_host, S, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { set_error("dist_cg_solve: %s", cudaGetErrorString(e)); st = SMB200_ERR_CUDA; }
    }
    if (st == SMB200_OK) {
        const double* H = w.scalars_host;
        const uint64_t iters = (uint64_t)H[S_ITER];
        w.history_host.assign(iters < w.hist_cap ? iters : w.hist_cap, 0.0);      // smb200_cg_history(smb200_dist_local(d)) reads it
        if (!w.history_host.empty())
            cudaMemcpy(w.history_host.data(), w.history, w.history_host.size() * sizeof(double), cudaMemcpyDeviceToHost);
        if (stats) {
            stats->iterations = iters;
            stats->final_residual = sqrt(sr ? H[S_RES2] : H[S_RR_NEW]);
            stats->converged = H[S_DONE] != 0.0;
            cudaEventElapsedTime(&stats->device_ms, ev0, ev1);
            stats->launches = g_launches - launches0;
        }
        st = dist_check_peers(d);
    }
    if (st != SMB200_OK) cudaStreamSynchronize(ctx->stream);
    return st;
}

extern "C" {

smb200_status smb200_dist_cg_solve(smb200_dist* d, const smb200_vec* b, smb200_vec* x, double tol, int32_t relative,
                                   uint64_t iter_max, smb200_cg_stats* stats) {
    return dist_cg_impl(d, b, x, tol, relative, iter_max, stats, false);
}

smb200_status smb200_dist_cg_solve_sr(smb200_dist* d, const smb200_vec* b, smb200_vec* x, double tol, int32_t relative,
                                      uint64_t iter_max, smb200_cg_stats* stats) {
    return dist_cg_impl(d, b, x, tol, relative, iter_max, stats, true);
}

}  // extern "C"
