// N4 — SparseMatPar (sparsemat_par.rs:12-35) as a device container with the `mvp_par` the reference left unfinished
// (sparsemat_par.rs:37-68: one worker per block over a shared `Arc<rhs>`, the blocks' results gathered at offset b * R).
//
// Block b holds global rows [b R, (b + 1) R), R = max_n_rows / n_blocks, with LOCAL row ids and GLOBAL column ids — the
// reference's layout, kept as is.  Here a worker is a GPU: with a communicator of `world` ranks, block b lives on rank
// b * world / n_blocks (contiguous runs of blocks per rank), x is replicated (every rank passes the whole right-hand side,
// the `Arc<rhs>`), each rank multiplies its own blocks into y at offset b R — concurrently, on a few streams — and the
// ranks then exchange their slices of y (one NCCL broadcast per rank, grouped), which is the channel gather of the sketch.
// On one GPU this is simply all blocks side by side.  This is the reference's replicated-x model and costs an exchange of
// the whole y per product; the row-block path with ghost entries (dist.cu) is the one to use for speed.
//
// Quirks kept: n_rows() stops at the first empty block; a short block in front of later rows makes the
// default mvp walk past the block's last row (indexlist.rs:88 panics) -> SMB200_ERR_INVALID "index out of bounds".
#include "common.cuh"

#include <algorithm>

namespace smb {
smb200_status nccl_group_begin();                                                   // dist.cu
smb200_status nccl_group_end();
smb200_status nccl_bcast_bytes(smb200_ctx* ctx, void* buf, size_t bytes, int root, cudaStream_t stream);
}  // namespace smb

struct smb200_par {
    smb200_ctx* ctx = nullptr;
    int vt = SMB200_F32, it = SMB200_U32;
    uint64_t n_blocks = 0, max_n_rows = 0, R = 0;
    std::vector<smb200_crs*> blocks;                 // device blocks owned by this rank (nullptr: empty / elsewhere)
    std::vector<uint64_t> rows_b, cols_b, nnz_b;     // dimensions of every block, known on every rank
    static constexpr int kStreams = 4;
    cudaStream_t streams[kStreams] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t fork = nullptr, join[kStreams] = {nullptr, nullptr, nullptr, nullptr};
};

using namespace smb;

static int par_owner(const smb200_par* p, uint64_t b) {
    return (int)((b * (uint64_t)p->ctx->world) / p->n_blocks);
}

// sparsemat_par.rs:95-107: scan the blocks until the first empty one
static uint64_t par_n_rows(const smb200_par* p) {
    uint64_t last = 0;
    for (uint64_t b = 0; b < p->n_blocks; ++b) {
        if (p->rows_b[b] == 0) break;
        last = b;
    }
    return last * p->R + p->rows_b[last];
}

extern "C" {

smb200_status smb200_par_create(smb200_ctx* ctx, uint64_t n_blocks, uint64_t max_n_rows, smb200_vtype vt, smb200_itype it,
                                smb200_par** out) {
    SMB_REQUIRE(ctx && out, SMB200_ERR_INVALID, "par_create: NULL argument");
    SMB_REQUIRE(vt == SMB200_F32 || vt == SMB200_F64, SMB200_ERR_INVALID, "par_create: bad value type %d", (int)vt);
    SMB_REQUIRE(it == SMB200_U32 || it == SMB200_U64, SMB200_ERR_INVALID, "par_create: bad index type %d", (int)it);
    SMB_REQUIRE(n_blocks > 0, SMB200_ERR_INVALID, "par_create: n_blocks == 0 (the reference divides by zero here, sparsemat_par.rs:21)");
    SMB_REQUIRE(max_n_rows / n_blocks > 0, SMB200_ERR_INVALID, "par_create: rows per block == 0 (the reference divides by zero on first access)");
    *out = nullptr;
    SMB_CUDA(cudaSetDevice(ctx->device));
    smb200_par* p = new smb200_par();
    p->ctx = ctx; p->vt = vt; p->it = it; p->n_blocks = n_blocks; p->max_n_rows = max_n_rows; p->R = max_n_rows / n_blocks;
    p->blocks.assign(n_blocks, nullptr);
    p->rows_b.assign(n_blocks, 0); p->cols_b.assign(n_blocks, 0); p->nnz_b.assign(n_blocks, 0);
    ctx_retain(ctx);
    cudaError_t e = cudaEventCreateWithFlags(&p->fork, cudaEventDisableTiming);
    for (int i = 0; i < smb200_par::kStreams && e == cudaSuccess; ++i) {
        e = cudaStreamCreateWithFlags(&p->streams[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->join[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) { smb200_par_free(p); SMB_CUDA(e); }
    *out = p;
    return SMB200_OK;
}

smb200_status smb200_par_free(smb200_par* p) {
    if (!p) return SMB200_OK;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    for (int i = 0; i < smb200_par::kStreams; ++i) {
        if (p->streams[i]) { cudaStreamSynchronize(p->streams[i]); cudaStreamDestroy(p->streams[i]); }
        if (p->join[i]) cudaEventDestroy(p->join[i]);
    }
    if (p->fork) cudaEventDestroy(p->fork);
    for (smb200_crs* b : p->blocks) if (b) smb200_crs_free(b);
    smb200_ctx* ctx = p->ctx;
    delete p;
    ctx_release(ctx);
    return SMB200_OK;
}

smb200_status smb200_par_owner(const smb200_par* p, uint64_t block, int32_t* rank) {
    SMB_REQUIRE(p && rank, SMB200_ERR_INVALID, "par_owner: NULL argument");
    SMB_REQUIRE(block < p->n_blocks, SMB200_ERR_INVALID, "index out of bounds: block %llu of %llu", (unsigned long long)block, (unsigned long long)p->n_blocks);
    *rank = par_owner(p, block);
    return SMB200_OK;
}

// Block b from the assembly format's arrays (SparseMatIndexList -> to_crs on the device, sparsemat_crs.rs:24-50).  Every
// rank makes the same call; the arrays are only read on the block's owner (they may be NULL elsewhere).
smb200_status smb200_par_set_block_indexlist(smb200_par* p, uint64_t block, uint64_t n_rows, uint64_t n_cols, uint64_t nnz,
                                             const void* columns, const void* values, const void* pos_start, const void* index_list) {
    SMB_REQUIRE(p, SMB200_ERR_INVALID, "par_set_block: NULL argument");
    SMB_REQUIRE(block < p->n_blocks, SMB200_ERR_INVALID, "index out of bounds: block %llu of %llu", (unsigned long long)block, (unsigned long long)p->n_blocks);
    SMB_REQUIRE(n_rows <= p->R, SMB200_ERR_INVALID, "par_set_block: block %llu has %llu rows, a block holds at most %llu", (unsigned long long)block,
                (unsigned long long)n_rows, (unsigned long long)p->R);
    if (p->blocks[block]) { cudaStreamSynchronize(p->ctx->stream); smb200_crs_free(p->blocks[block]); p->blocks[block] = nullptr; }
    if (nnz == 0) n_rows = n_cols = 0;                               // to_crs of an assembly without entries is 0 x 0 (sparsemat_crs.rs:25)
    p->rows_b[block] = n_rows; p->cols_b[block] = n_cols; p->nnz_b[block] = nnz;
    if (par_owner(p, block) != p->ctx->rank || nnz == 0) return SMB200_OK;
    return smb200_crs_from_indexlist(p->ctx, (smb200_vtype)p->vt, (smb200_itype)p->it, n_rows, n_cols, nnz, columns, values, pos_start,
                                     index_list, &p->blocks[block]);
}

// Block b from finished CRS arrays (local row offsets, global columns).
smb200_status smb200_par_set_block_crs(smb200_par* p, uint64_t block, uint64_t n_rows, uint64_t n_cols, uint64_t nnz,
                                       const void* values, const void* columns, const void* offset_rows) {
    SMB_REQUIRE(p, SMB200_ERR_INVALID, "par_set_block: NULL argument");
    SMB_REQUIRE(block < p->n_blocks, SMB200_ERR_INVALID, "index out of bounds: block %llu of %llu", (unsigned long long)block, (unsigned long long)p->n_blocks);
    SMB_REQUIRE(n_rows <= p->R, SMB200_ERR_INVALID, "par_set_block: block %llu has %llu rows, a block holds at most %llu", (unsigned long long)block,
                (unsigned long long)n_rows, (unsigned long long)p->R);
    if (p->blocks[block]) { cudaStreamSynchronize(p->ctx->stream); smb200_crs_free(p->blocks[block]); p->blocks[block] = nullptr; }
    p->rows_b[block] = n_rows; p->cols_b[block] = n_cols; p->nnz_b[block] = nnz;
    if (par_owner(p, block) != p->ctx->rank || n_rows == 0) return SMB200_OK;
    return smb200_crs_upload(p->ctx, (smb200_vtype)p->vt, (smb200_itype)p->it, n_rows, n_cols, nnz, values, columns, offset_rows, &p->blocks[block]);
}

// out3 = {n_rows (sparsemat_par.rs:95-107: up to the first empty block), n_cols (largest block), non-zeros (all blocks)}
smb200_status smb200_par_dims(const smb200_par* p, uint64_t* out3) {
    SMB_REQUIRE(p && out3, SMB200_ERR_INVALID, "par_dims: NULL argument");
    out3[0] = par_n_rows(p);
    out3[1] = *std::max_element(p->cols_b.begin(), p->cols_b.end());
    out3[2] = 0;
    for (uint64_t z : p->nnz_b) out3[2] += z;
    return SMB200_OK;
}

smb200_status smb200_par_block(smb200_par* p, uint64_t block, smb200_crs** out) {
    SMB_REQUIRE(p && out, SMB200_ERR_INVALID, "par_block: NULL argument");
    SMB_REQUIRE(block < p->n_blocks, SMB200_ERR_INVALID, "index out of bounds: block %llu of %llu", (unsigned long long)block, (unsigned long long)p->n_blocks);
    *out = p->blocks[block];                                         // borrowed; NULL for an empty block or one that lives on another rank
    return SMB200_OK;
}

// y = A x: the completed mvp_par.  x: the whole right-hand side on every rank (dim >= n_cols); y: dim >= n_rows, complete on
// every rank afterwards.
smb200_status smb200_par_mvp(smb200_par* p, const smb200_vec* x, smb200_vec* y) {
    SMB_REQUIRE(p && x && y, SMB200_ERR_INVALID, "par_mvp: NULL argument");
    SMB_REQUIRE(x->vt == p->vt && y->vt == p->vt, SMB200_ERR_INVALID, "par_mvp: value types differ");
    SMB_REQUIRE(x->d != y->d, SMB200_ERR_INVALID, "par_mvp: x and y alias");
    smb200_ctx* ctx = p->ctx;
    const uint64_t n = par_n_rows(p);
    const uint64_t n_cols = *std::max_element(p->cols_b.begin(), p->cols_b.end());
    SMB_REQUIRE(x->n >= n_cols, SMB200_ERR_DIM, "Dimension mismatch: x has %llu entries, matrix has %llu columns", (unsigned long long)x->n,
                (unsigned long long)n_cols);
    SMB_REQUIRE(y->n >= n, SMB200_ERR_DIM, "Dimension mismatch: y has %llu entries, matrix has %llu rows", (unsigned long long)y->n, (unsigned long long)n);
    if (n == 0) return SMB200_OK;
    // the blocks the default mvp visits: 0 .. last (the one n_rows() stopped at); a short block in front of later rows is
    // the reference's IndexList::iter_row past the end
    uint64_t last = 0;
    for (uint64_t b = 0; b < p->n_blocks && p->rows_b[b] != 0; ++b) last = b;
    for (uint64_t b = 0; b < last; ++b)
        SMB_REQUIRE(p->rows_b[b] == p->R, SMB200_ERR_INVALID, "index out of bounds: block %llu holds %llu of its %llu rows but later blocks have rows "
                    "(indexlist.rs:88)", (unsigned long long)b, (unsigned long long)p->rows_b[b], (unsigned long long)p->R);
    SMB_CUDA(cudaSetDevice(ctx->device));
    const size_t es = vsize(p->vt);
    // own blocks side by side on a few streams (the sketch's thread per block)
    SMB_CUDA(cudaEventRecord(p->fork, ctx->stream));
    bool used[smb200_par::kStreams] = {false, false, false, false};
    smb200_status st = SMB200_OK;
    int k = 0;
    for (uint64_t b = 0; b <= last && st == SMB200_OK; ++b) {
        smb200_crs* blk = p->blocks[b];
        if (!blk) continue;
        const int s = k++ % smb200_par::kStreams;
        if (!used[s]) { SMB_CUDA(cudaStreamWaitEvent(p->streams[s], p->fork, 0)); used[s] = true; }
        g_redirect.stream = p->streams[s];
        st = spmv_launch(blk, x->d, (char*)y->d + (size_t)(b * p->R) * es, nullptr, 0);
        g_redirect = LaunchRedirect();
    }
    for (int s = 0; s < smb200_par::kStreams; ++s)
        if (used[s]) {
            SMB_CUDA(cudaEventRecord(p->join[s], p->streams[s]));
            SMB_CUDA(cudaStreamWaitEvent(ctx->stream, p->join[s], 0));
        }
    SMB_TRY(st);
    if (ctx->world > 1) {
        // the gather: every rank broadcasts the rows of its blocks
        SMB_TRY(nccl_group_begin());
        for (int q = 0; q < ctx->world && st == SMB200_OK; ++q) {
            uint64_t b0 = p->n_blocks, b1 = 0;
            for (uint64_t b = 0; b <= last; ++b) if (par_owner(p, b) == q) { b0 = std::min(b0, b); b1 = std::max(b1, b + 1); }
            if (b0 >= b1) continue;
            const uint64_t lo = b0 * p->R, hi = std::min(n, (b1 - 1) * p->R + p->rows_b[b1 - 1]);
            if (hi > lo) st = nccl_bcast_bytes(ctx, (char*)y->d + lo * es, (hi - lo) * es, q, ctx->stream);
        }
        const smb200_status ge = nccl_group_end();
        SMB_TRY(st);
        SMB_TRY(ge);
    }
    return SMB200_OK;
}

}  // extern "C"
