// libsmb200 internals shared by all translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/smb200.h"

namespace smb {

// ---- error plumbing: the C layer never throws or aborts ------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define SMB_FAIL(code, ...)            \
    do {                               \
        ::smb::set_error(__VA_ARGS__); \
        return (code);                 \
    } while (0)

#define SMB_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            ::smb::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__, \
                             cudaGetErrorString(e__));                                              \
            return e__ == cudaErrorMemoryAllocation ? SMB200_ERR_OOM : SMB200_ERR_CUDA;             \
        }                                                                                           \
    } while (0)

#define SMB_TRY(expr)                      \
    do {                                   \
        smb200_status s__ = (expr);        \
        if (s__ != SMB200_OK) return s__;  \
    } while (0)

#define SMB_REQUIRE(cond, code, ...)            \
    do {                                        \
        if (!(cond)) SMB_FAIL(code, __VA_ARGS__); \
    } while (0)

// kernels launched by this library (bench evidence: gpu_launches)
extern thread_local uint64_t g_launches;
inline void count_launch(uint64_t n = 1) { g_launches += n; }

inline size_t vsize(int vt) { return vt == SMB200_F64 ? 8 : 4; }
inline size_t isize(int it) { return it == SMB200_U64 ? 8 : 4; }

// All device arrays are padded so that 128-bit loads and 16-byte bulk copies that start at an
// aligned-down address and end at an aligned-up one never leave the allocation.
constexpr size_t kPadBytes = 256;

}  // namespace smb

struct smb200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t aux_stream = nullptr;   // halo traffic / overlap
    cudaStream_t copy_in = nullptr, copy_out = nullptr;   // smb200_spmv_host: H2D / D2H run beside the compute stream
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    int sm_count = 0;
    size_t l2_bytes = 0, l2_persist_max = 0;
    // scratch for reductions: partial sums + ticket + result slots
    double* red_partials = nullptr;      // [red_cap]
    size_t red_cap = 0;
    double* red_partials_aux = nullptr;  // [red_cap_aux] dot partials of launches on the side stream (dist boundary rows)
    size_t red_cap_aux = 0;
    unsigned int* red_ticket = nullptr;  // zero between uses
    double* red_result = nullptr;        // device [8]
    double* red_result_host = nullptr;   // pinned [8]
    void* flush_buf = nullptr;
    size_t flush_bytes = 0;
    // staging for *_host calls
    void* stage_x = nullptr; size_t stage_x_bytes = 0;
    void* stage_y = nullptr; size_t stage_y_bytes = 0;
    // NCCL (dist.cu); opaque here
    void* comm = nullptr;
    int rank = 0, world = 1;
    // lifetime: every vector / matrix / event / dist handle created from the context holds a reference; a
    // smb200_ctx_destroy issued while handles are still alive is deferred until the last one is freed
    // (garbage-collected language bindings finalise objects in arbitrary order).
    int refs = 0;
    bool destroy_requested = false;
};

struct smb200_vec {
    smb200_ctx* ctx = nullptr;
    int vt = SMB200_F32;
    uint64_t n = 0;        // dim()
    uint64_t cap = 0;      // allocated elements (>= n; dist vectors keep ghost room behind n)
    void* d = nullptr;
    bool owned = true;
};

struct smb200_event {
    smb200_ctx* ctx = nullptr;
    cudaEvent_t ev = nullptr;
};

namespace smb {

// ---- SpMV plan: built once per matrix (and per variant) -------------------------------------------
}  // namespace smb
struct smb200_crs;
namespace smb {
struct SpmvPlan {
    int variant = SMB200_SPMV_AUTO;     // resolved
    int lanes = 0;
    uint32_t flags = 0;
    uint64_t n_blocks = 0;              // STREAM*: CTAs
    void* blk_rows = nullptr;           // STREAM*: [n_blocks+1] row split points, index type
    void* blk_nnz = nullptr;            // STREAM*: [n_blocks+1] offset_rows[blk_rows[k]], index type
    unsigned char* blk_flags = nullptr; // PIPE: per block, kBlkNoFit | kBlkLong
    unsigned long long* seg_lo = nullptr; // RING: [4 * n_blocks] first column of each x segment of the block (aligned down)
    unsigned* seg_len = nullptr;        // RING: [4 * n_blocks] segment lengths in elements (0 = unused; all 0 = no window)
    unsigned ocap = 0, xcap = 0;        // RING: row-offset / x-window capacity of a stage (elements)
    // RING, value-indexed plans in sliced-ELLPACK stage order (spmv_sell.cuh): the compressed entries of every block laid out
    // 32 rows at a time the way a warp reads them; replaces vcodes / lcols / loffs when the padding is small
    void* sell_blocks = nullptr;        //   [n_blocks] SellBlock
    uint8_t* sell_codes = nullptr;      //   [sell_entries] value codes
    uint16_t* sell_cols = nullptr;      //   [sell_entries] window positions
    uint8_t* sell_rowlen = nullptr;     //   row lengths, each block padded to 16
    uint32_t* sell_soff = nullptr;      //   slice offsets, n_slices + 1 per block, padded to 4
    uint64_t sell_entries = 0, sell_rowbytes = 0, sell_soffwords = 0;
    unsigned sell_ecap = 0, sell_rcap = 0, sell_scap = 0;
    uint8_t* vcodes = nullptr;          // RING, value indexing: 8-bit code of every non-zero into its block's dictionary
    void* vdict = nullptr;              //       [256 * n_blocks] dictionaries (T): the sorted distinct values of each block
    uint64_t n_v8 = 0;                  //       non-zeros whose value the kernel reads through a code
    uint16_t* loffs = nullptr;          // RING, packed plans: 16-bit row offsets relative to each block's slice start
    uint64_t loffs_row_begin = 0;       //       first row of the plan's range (the kernel derives a block's position from it)
    uint64_t n_o16 = 0;                 //       rows whose offsets the kernel reads from loffs
    unsigned colb = 0;                  // RING: bytes per element of a stage's column area (2 = packed: every block compressed)
    uint64_t n_xwin = 0;                // RING: blocks whose columns fit <= 4 windows
    uint16_t* lcols = nullptr;          // RING: 16-bit window-relative columns of the windowed blocks (index compression);
    uint64_t lcols_base = 0;            //       lcols[k - lcols_base] belongs to non-zero k (lcols_base is a multiple of 8)
    uint64_t n_c16 = 0;                 //       non-zeros whose column the kernel reads from lcols instead of `columns`
    unsigned cap = 0, target = 0;       // STREAM*: staging capacity / merge target the plan was cut for (elements)
    void* blk_win = nullptr;            // BANDED: [2*n_blocks] (cmin, cmax+1) per block, index type
    uint64_t max_win = 0;               // BANDED: widest window (elements)
    // BANDSPLIT: the matrix cut into column bands of band_width elements, each an ordinary CRS matrix (same rows, columns
    // rebased to its band) with its own stream plan; owned by the plan
    std::vector<smb200_crs*> parts;
    uint64_t band_width = 0;
    double build_ms = 0.0;              // host wall-clock time of the build
    bool built = false;
};

// Host-buffer pipeline of smb200_spmv_host: the rows are cut into chunks; chunk c starts as soon as the pieces of x its
// columns reach have arrived (banded matrices: a few pieces), and its slice of y goes back while later chunks compute.
struct HostPipe {
    bool built = false;
    int n_chunks = 0;
    std::vector<uint64_t> row_bounds;    // [n_chunks + 1]
    std::vector<uint64_t> x_bounds;      // [n_chunks + 1] pieces of x (elements)
    std::vector<int> last_piece;         // per chunk: the last piece of x it reads
    std::vector<SpmvPlan> plans;         // per chunk
    std::vector<cudaEvent_t> ev_x, ev_y;
};

// CG workspace kept with the matrix so repeated solves do not reallocate.
struct CgWork {
    uint64_t n = 0;
    void* r = nullptr; void* p = nullptr; void* ap = nullptr;
    void* s = nullptr;                  // single-reduction variant (cg_sr.cuh): s = A p by recurrence
    uint64_t cap = 0;                   // allocated elements of r, p, s (>= n: ghost room in distributed solves)
    double* scalars = nullptr;          // device: see cg.cu
    double* scalars_host = nullptr;     // pinned mirror
    double* history = nullptr;          // device [hist_cap]
    uint64_t hist_cap = 0;
    std::vector<double> history_host;
    cudaGraphExec_t graph = nullptr;    // batch of iterations
    int graph_batch = 0;
    const void* graph_x = nullptr;      // pointers the graph was captured with
    const void* graph_partials = nullptr;
    const void* graph_plan = nullptr;
    int graph_kind = 0;                 // 0: the reference's loop, 1: single-reduction
    const void* graph_dinv = nullptr;   // Jacobi-preconditioned batch: the inverse diagonal it was captured with
    void* dinv = nullptr;               // 1 / diag(A) of the last preconditioned solve (n elements)
    uint64_t dinv_n = 0;
};

}  // namespace smb

struct smb200_crs {
    smb200_ctx* ctx = nullptr;
    int vt = SMB200_F32, it = SMB200_U32;
    uint64_t n_rows = 0, n_cols = 0, nnz = 0;
    void* values = nullptr;
    void* columns = nullptr;
    void* offsets = nullptr;            // [n_rows+1] (nullptr for the 0x0 matrix)
    uint64_t max_row_len = 0;
    int want_variant = SMB200_SPMV_AUTO;
    int want_lanes = 0;
    uint32_t want_flags = 0;
    smb::SpmvPlan plan;
    smb::CgWork cg;
    smb::HostPipe hp;
    uint64_t x_extra = 0;               // dist: number of ghost entries appended to x (n_cols counts them)
};

namespace smb {

// SpMV launches normally go to ctx->stream with ctx->red_partials; dist.cu redirects the boundary-row launches to
// the side stream (behind the halo receive) with their own partials buffer.
struct LaunchRedirect { cudaStream_t stream = nullptr; double* partials = nullptr; };
extern thread_local LaunchRedirect g_redirect;
// SMs a persistent (ring) launch leaves free: dist.cu sets it for the interior product so that the NCCL send/recv kernel of
// the halo exchange, queued on the side stream, finds an SM while the interior is still running.
extern thread_local int g_ring_reserve_sms;
// Upper bound on the CTAs of a persistent (ring) launch (0 = none): boundary-row launches that run beside the interior
// kernel use as many CTAs as it left slots, each walking several blocks through its ring.
extern thread_local int g_ring_grid_cap;
// Set while the second and later bands of a band-split product are launched: the stream kernel adds to y instead of writing it.
extern thread_local bool g_spmv_accumulate;
// Set while the bands of a band-split product are launched: the stream plan of a part runs spmv_band_kernel (L2 eviction
// hints that keep the band of x resident, prefetched y).
extern thread_local bool g_spmv_band;
// Set by dist.cu around the ONE ring launch of a distributed product: the kernel then also runs the halo protocol of
// halo.cuh (push the neighbours' ghost entries, wait for this rank's) and walks the row blocks rotated by `rot`.
struct HaloDev;
struct ArDev;
// Set by dist.cu around the product of a distributed CG iteration: the fused dot's finalize kernel then all-reduces the
// ranks' totals through peer memory (halo.cuh) instead of leaving that to a kernel of its own.
extern thread_local const ArDev* g_dot_ar;
struct HaloLaunch { const HaloDev* host = nullptr; uint64_t rot = 0; };     // host copy: passed to the kernel by value
extern thread_local HaloLaunch g_halo;
// Set while an SpMV reads a caller-owned (smb200_vec_wrap) vector: such memory has no padding behind its last element,
// so kernels must not round bulk copies of it up to 16 bytes.
extern thread_local bool g_x_unpadded;

// implemented across the .cu files
void ctx_retain(smb200_ctx* ctx);
void ctx_release(smb200_ctx* ctx);   // may tear the context down if its destruction was deferred
smb200_status dev_alloc(void** p, size_t bytes);
smb200_status crs_alloc(smb200_ctx* ctx, int vt, int it, uint64_t n_rows, uint64_t n_cols, uint64_t nnz,
                        smb200_crs** out);
smb200_status crs_finalize(smb200_crs* m, bool validate);   // row stats (+ validation), invalidates plan
smb200_status exclusive_scan_inplace(smb200_ctx* ctx, int it, void* data, uint64_t n, uint64_t* total);
smb200_status plan_build(smb200_crs* m);
smb200_status plan_build_range(smb200_crs* m, SpmvPlan& p, int want_variant, int want_lanes, uint32_t flags,
                               uint64_t row_begin, uint64_t row_end);
void plan_free(SpmvPlan& p);
smb200_status bandsplit_build(smb200_crs* m, SpmvPlan& p, int fallback_variant);      // bandsplit.cu
uint64_t bandsplit_stream_bytes(const smb200_crs* m, const SpmvPlan& p);
void hostpipe_free(HostPipe& hp);
void cg_free(CgWork& w);
smb200_status spmv_launch_plan(smb200_crs* m, const SpmvPlan& p, uint64_t row_begin, uint64_t row_end, const void* x,
                               void* y, const void* w, int dot_slot);
smb200_status spmv_launch_cg(smb200_crs* m, const SpmvPlan& p, uint64_t row_begin, uint64_t row_end, const void* x,
                             void* y, const void* w, double* S, int slot, bool roll);
smb200_status vec_create_cap(smb200_ctx* ctx, int vt, uint64_t n, uint64_t cap, smb200_vec** out);
smb200_status gen_laplace_block(smb200_ctx* ctx, int vt, int it, uint64_t nx, uint64_t ny, uint64_t nz, uint64_t row_lo,
                                uint64_t row_hi, int local_cols, uint64_t n_lo_ghost, uint64_t n_cols_out,
                                smb200_crs** out, uint64_t ghost_base = 0);   // ghost_base 0: ghosts right behind the owned columns
smb200_status cg_prepare(smb200_ctx* ctx, CgWork& w, int vt, uint64_t n, uint64_t p_cap, uint64_t iter_max);
smb200_status cg_init_launch(smb200_ctx* ctx, CgWork& w, int vt, const void* b, uint64_t n, const void* dinv = nullptr, int rr_slot = -1);
smb200_status cgsr_prepare(smb200_ctx* ctx, CgWork& w, int vt);
smb200_status cgsr_update_launch(smb200_ctx* ctx, CgWork& w, int vt, void* x, uint64_t n);
smb200_status cgsr_scalar_launch(smb200_ctx* ctx, CgWork& w, int vt, const ArDev* ar);
smb200_status cg_r_launch(smb200_ctx* ctx, CgWork& w, int vt, uint64_t n, const void* dinv = nullptr, const ArDev* ar = nullptr);
smb200_status cg_xp_launch(smb200_ctx* ctx, CgWork& w, int vt, void* x, uint64_t n, const void* dinv = nullptr);
smb200_status cg_solve_impl(smb200_crs* a, const smb200_vec* b, smb200_vec* x, double tol, int32_t relative,
                            uint64_t iter_max, smb200_cg_stats* stats, const void* dinv);

// y = A x on m->ctx->stream.  If dot_out != nullptr, also accumulates sum_r w[r]*y[r] into the
// reduction scratch and leaves the result in ctx->red_result[slot] (fused SpMV + dot, K7a).
smb200_status spmv_launch(smb200_crs* m, const void* x, void* y, const void* w, int dot_slot,
                          uint64_t row_begin = 0, uint64_t row_end = UINT64_MAX);

// reductions leave their value (as double, already rounded to T) in ctx->red_result[slot]
smb200_status dot_launch(smb200_ctx* ctx, int vt, const void* x, const void* y, uint64_t n, int slot);
smb200_status fetch_result(smb200_ctx* ctx, int slot, double* out);   // D2H + sync
smb200_status ensure_reduction_scratch(smb200_ctx* ctx, size_t n_partials);

}  // namespace smb
