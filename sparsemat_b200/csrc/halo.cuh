// Peer-memory plumbing of the row-block distributed path (dist.cu): ghost entries of x and the CG scalars travel by
// plain stores into the neighbour's HBM over NVLink / NVSwitch (CUDA IPC mappings exchanged once, when the distributed
// matrix is created) instead of NCCL send/recv and all-reduce kernels.
//
// Halo protocol, one "epoch" per distributed product (SparseMatPar's model, sparsemat_par.rs:39-67: every row block needs
// the entries of x its columns touch; here only those entries move):
//   1. x is final when the product's first kernel starts (stream order).  Its threads copy the entries the neighbours need
//      into the neighbours' ghost buffers (half `epoch & 1` of a double buffer), fence, and the last CTA to finish its share
//      stores the epoch number into every neighbour's flag word (release, system scope).
//   2. Whoever needs ghost entries — the TMA producer of a block with a ghost window, or the wait kernel in front of the
//      boundary rows — spins (acquire, system scope) until all neighbours' flags have reached the epoch.
//   3. A product ends only after all neighbours' flags of its epoch were seen, so a rank can be at most one epoch ahead of
//      a neighbour: the half it writes next is never the half that neighbour may still be reading.
// No kernel of another rank has to be resident for a store to land, so nothing here can dead-lock on SM occupancy; waits
// are bounded by `timeout_ns` (a lost peer sets the error word instead of hanging the GPU).
#pragma once
#include <cstdint>

namespace smb {

constexpr int kMaxNbr = 8;       // neighbours in the halo exchange of one rank
constexpr int kMaxPeers = 16;    // ranks of a peer-memory communicator (one NVSwitch domain)
constexpr int kArSlots = 4;      // doubles per all-reduce

struct HaloDev {
    unsigned long long* epoch;            // [1] last completed epoch (local)
    unsigned* ctr;                        // [2] arrivals: push finished / product finished; zero between launches (local)
    unsigned long long* flags;            // [world] flags[q] = last epoch rank q has pushed into this rank's buffer (peers write)
    unsigned* error;                      // [1] set when a wait ran into the timeout
    unsigned long long* wait_stats;       // [2] evidence: {sum of the waiting threads' spin times in ns, number of waits}
    void* ghost;                          // this rank's ghost buffer: 2 halves of ghost_stride elements (peers write)
    unsigned long long ghost_stride;      // elements per half
    unsigned long long n_ghost;
    unsigned long long g0;                // first ghost column in local numbering (a multiple of 64)
    unsigned long long timeout_ns;
    const unsigned long long* send_idx;   // local row ids to pack, grouped by neighbour
    unsigned long long total_send;
    int n_nbr;
    int nbr_rank[kMaxNbr];
    unsigned long long* peer_flag[kMaxNbr];     // &flags[me] inside neighbour i's window
    void* peer_ghost[kMaxNbr];                  // neighbour i's ghost buffer at this rank's receive offset
    unsigned long long peer_stride[kMaxNbr];    // neighbour i's ghost_stride
    unsigned long long send_off[kMaxNbr], send_count[kMaxNbr], send_first[kMaxNbr];
    int send_contig[kMaxNbr];
};

struct ArDev {
    unsigned long long* epoch;            // [1] local
    unsigned long long* flags;            // [2][world] (peers write)
    double* vals;                         // [2][world][kArSlots] (peers write)
    unsigned* error;
    unsigned long long timeout_ns;
    int world, me;
    unsigned long long* peer_flags[kMaxPeers];
    double* peer_vals[kMaxPeers];
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// Step 1, data: thread `tid` of `nthreads` copies its share of the entries the neighbours need (plain stores; see halo_arrive).
template <class T>
__device__ __forceinline__ void halo_push(const HaloDev& h, const T* __restrict__ x, unsigned long long e, uint64_t tid, uint64_t nthreads) {
    const int n = h.n_nbr;
    for (int i = 0; i < n; ++i) {
        const unsigned long long cnt = h.send_count[i];
        if (cnt == 0) continue;
        T* dst = (T*)h.peer_ghost[i] + (e & 1ull) * h.peer_stride[i];
        if (h.send_contig[i]) {
            const T* src = x + h.send_first[i];
            for (unsigned long long k = tid; k < cnt; k += nthreads) dst[k] = src[k];
        } else {
            const unsigned long long* idx = h.send_idx + h.send_off[i];
            for (unsigned long long k = tid; k < cnt; k += nthreads) dst[k] = x[idx[k]];
        }
    }
}
// The caller then joins its pushing threads with a CTA-level barrier and ONE of them calls halo_arrive: its system-scope
// fence is cumulative over the stores the barrier made visible to it (the pattern of a grid-wide barrier), so the 100k
// pushing threads need no fence of their own.
__device__ __forceinline__ void halo_signal(const HaloDev& h, unsigned long long e);
__device__ __forceinline__ void halo_arrive(const HaloDev& h, unsigned long long e, unsigned n_ctas) {
    __threadfence_system();
    if (atomicAdd(h.ctr, 1u) == n_ctas - 1) { h.ctr[0] = 0u; halo_signal(h, e); }
}

// Step 1, flags: one thread, after every pushing thread of the grid has fenced and arrived.
__device__ __forceinline__ void halo_signal(const HaloDev& h, unsigned long long e) {
    __threadfence_system();
    for (int i = 0; i < h.n_nbr; ++i) st_release_sys(h.peer_flag[i], e);
}

// Step 2: one thread.  Returns false after a timeout (the error word is set; the caller carries on with stale ghosts so
// that the kernel terminates and the host can report the failure).
__device__ __forceinline__ bool halo_wait(const HaloDev& h, unsigned long long e) {
    const unsigned long long t0 = global_timer_ns();
    for (int i = 0; i < h.n_nbr; ++i) {
        const unsigned long long* f = h.flags + h.nbr_rank[i];
        unsigned spins = 0;
        while (ld_acquire_sys(f) < e) {
            if ((++spins & 255u) == 0 && global_timer_ns() - t0 > h.timeout_ns) { atomicExch(h.error, 1u); return false; }
            __nanosleep(40);
        }
    }
    if (h.wait_stats) { atomicAdd(h.wait_stats, global_timer_ns() - t0); atomicAdd(h.wait_stats + 1, 1ull); }
    return true;
}
// All-reduce (sum) of `count` <= kArSlots doubles, executed by ONE converged warp inside any kernel: lane q < world stores
// this rank's values (`mine[k]`, the same in every lane) into rank q's slot array and raises its flag, then waits for rank
// q's values here; lane 0 adds the world's values in rank order (every rank computes the same bits) and returns them in
// out[k].  sv: shared scratch [kMaxPeers][kArSlots].  Epochs alternate between two slot sets (a rank can be one all-reduce
// ahead).  dist.cu's p2p_allreduce_kernel is this in a kernel of its own; the CG kernels call it in their last block.
__device__ __forceinline__ void ar_warp_allreduce(const ArDev& a, const double* mine, double* out, int count, double (*sv)[kArSlots]) {
    const unsigned long long e = __ldcg(a.epoch) + 1ull;
    const int world = a.world, me = a.me, q = (int)(threadIdx.x & 31u);
    const size_t par = (size_t)(e & 1ull);
    if (q < world) {
        double* dst = a.peer_vals[q] + (par * world + me) * kArSlots;
        for (int k = 0; k < count; ++k) dst[k] = mine[k];
        __threadfence_system();
        st_release_sys(a.peer_flags[q] + par * world + me, e);
        const unsigned long long* f = a.flags + par * world + q;
        const unsigned long long t0 = global_timer_ns();
        unsigned spins = 0;
        while (ld_acquire_sys(f) < e) {
            if ((++spins & 255u) == 0 && global_timer_ns() - t0 > a.timeout_ns) { atomicExch(a.error, 1u); break; }
        }
        const double* src = a.vals + (par * world + q) * kArSlots;
        for (int k = 0; k < count; ++k) sv[q][k] = __ldcg(src + k);
    }
    __syncwarp();
    if (q == 0) {
        for (int k = 0; k < count; ++k) {
            double sum = 0.0;
            for (int r = 0; r < world; ++r) sum += sv[r][k];
            out[k] = sum;
        }
        *a.epoch = e;
    }
    __syncwarp();
}
#endif

}  // namespace smb
