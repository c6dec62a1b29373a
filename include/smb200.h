/*
 * smb200.h — C ABI of libsmb200, the B200 (sm_100a) engine behind lostinc0de/sparsemat's hot path.
 *
 * This is the drop-in boundary: exactly the calls an FFI binding of the reference crate would make
 * for the CRS sparse matrix-vector product, the dense-vector kernels and the CG solver.  Each entry
 * point cites the reference item it replaces (paths relative to /root/reference/src/).
 * INTEGRATION.md shows the Rust `extern "C"` block and the safe wrappers that sit on top of it.
 *
 * Conventions
 *   - every function returns an smb200_status (0 = ok); on failure smb200_last_error() describes it;
 *   - the library never aborts or throws; the reference's panics ("Dimension mismatch", "Matrix is
 *     not symmetric", "Matrix and vector size mismatch") are reported as status codes and re-raised
 *     by the language wrappers with the reference's own messages;
 *   - handles are opaque, owned by the caller and freed explicitly; uploads copy, the caller keeps
 *     its host arrays; a context and everything created from it belong to one host thread at a time
 *     (the reference's &mut discipline; handles are not Sync);
 *   - all device work is asynchronous on the context's stream unless the call returns a host value;
 *   - value types {f32, f64} (types.rs:70-77 FloatType), index types {u32, u64} (types.rs:48-49);
 *     row offsets are stored in the index type exactly like `offset_rows: Vec<I>` (sparsemat_crs.rs:14);
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry point fails with
 *     SMB200_ERR_CUDA.
 */
#ifndef SMB200_H
#define SMB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMB200_VERSION 100 /* 0.1.0 */

typedef int32_t smb200_status;
enum {
    SMB200_OK = 0,
    SMB200_ERR_INVALID = 1,       /* bad argument / malformed CRS arrays                              */
    SMB200_ERR_DIM = 2,           /* densevec.rs:52-54,61-63 "Dimension mismatch"; x shorter than n_cols */
    SMB200_ERR_CUDA = 3,          /* CUDA runtime error (incl. no device)                             */
    SMB200_ERR_NOT_SQUARE = 4,    /* linearsolver.rs:30-32 "Matrix is not symmetric" (n_rows != n_cols) */
    SMB200_ERR_SIZE_MISMATCH = 5, /* linearsolver.rs:33-36 "Matrix and vector size mismatch"          */
    SMB200_ERR_NCCL = 6,
    SMB200_ERR_UNSUPPORTED = 7,
    SMB200_ERR_OOM = 8,
    SMB200_ERR_IO = 9             /* smb200_crsfile_* / crs_save / crs_load: open/read/write failure, bad or damaged file */
};

typedef enum { SMB200_F32 = 0, SMB200_F64 = 1 } smb200_vtype;   /* types.rs:70-77 */
typedef enum { SMB200_U32 = 0, SMB200_U64 = 1 } smb200_itype;   /* types.rs:48-49 */

/* SpMV kernel families.  AUTO picks from the row-length statistics and, for short rows, from
 * the column structure. */
typedef enum {
    SMB200_SPMV_AUTO = 0,
    SMB200_SPMV_SCALAR = 1,  /* one thread per row, storage-order sum (bit-exact vs the reference)   */
    SMB200_SPMV_VECTOR = 2,  /* K1/K2: `lanes` threads per row (2..32), shuffle reduction            */
    SMB200_SPMV_STREAM = 3,  /* K3: nnz+row balanced CTAs, products staged in shared memory,
                                storage-order per-row sums (bit-exact for rows <= 64 nnz)            */
    SMB200_SPMV_STREAM_TMA = 4, /* K3 with values/columns brought in by cp.async.bulk (TMA)          */
    SMB200_SPMV_BANDED = 5,  /* K4: STREAM_TMA + the x window of the block staged in shared memory   */
    SMB200_SPMV_STREAM_PIPE = 6, /* K3 persistent: one CTA per SM slot walks its row blocks through a
                                multi-stage shared-memory ring filled by cp.async.bulk (TMA) + mbarrier,
                                so the HBM stream never waits for the gather / row-sum phases         */
    SMB200_SPMV_RING = 7,    /* K4 persistent, short rows (<= 32 entries: stencils, FEM): values, columns, row
                                offsets and the block's x segments (up to 4 windows found at plan time) are
                                ALL staged by TMA into a shared-memory ring by a producer warp; consumer warps
                                sum one row per thread in storage order (bit-exact) with no block barrier.
                                Blocks without windows gather x from global memory.  Falls back to STREAM
                                when the matrix has longer rows.  AUTO picks it when >= 80 % of the blocks
                                have windows                                                           */
    SMB200_SPMV_BANDSPLIT = 8 /* x larger than L2, scattered columns (power-law / graph matrices): at plan time the
                                matrix is cut into the fewest equal column bands of at most 0.53 L2 of x (same rows,
                                storage order kept inside a band, columns rebased to the band -> u32); one launch
                                per band, the first writes y, the others add to it, so the gathers of a launch
                                hit a cache-resident band.  Rows are summed band-major: tolerance-exact, not
                                bit-exact.  Whole-matrix products only (row ranges fall back to STREAM).
                                AUTO picks it when x is more than twice the L2, the ring kernel was not kept and
                                the plan's second copy of the matrix fits (SMB200_BANDSPLIT_AUTO=0 disables)      */
} smb200_spmv_variant;

/* flags for smb200_crs_configure */
#define SMB200_FLAG_L2_PERSIST_X 1u  /* K4: L2 access-policy window (persisting) over x              */

typedef struct smb200_ctx smb200_ctx;
typedef struct smb200_vec smb200_vec;
typedef struct smb200_crs smb200_crs;
typedef struct smb200_event smb200_event;
typedef struct smb200_dist smb200_dist;

typedef struct {
    int32_t device;
    int32_t sm_count;
    int32_t cc_major, cc_minor;
    int64_t l2_bytes;
    int64_t l2_persist_max_bytes;
    int64_t hbm_bytes;
    char name[128];
} smb200_devinfo;

typedef struct {
    int32_t variant;          /* resolved smb200_spmv_variant                                         */
    int32_t lanes;            /* VECTOR: threads per row                                              */
    uint32_t flags;
    uint64_t n_blocks;        /* STREAM*: number of row blocks (CTAs)                                 */
    uint64_t n_rows, n_cols, nnz;
    uint64_t max_row_len;
    double mean_row_len;
    uint64_t algorithmic_bytes; /* nnz*(sizeof T + sizeof I) + (n_rows+1)*sizeof I + n_cols*sizeof T + n_rows*sizeof T */
    uint64_t launches_per_spmv;
    uint64_t n_xwin_blocks;   /* RING: row blocks whose x segments are staged by TMA (the rest gather from global) */
    uint64_t nnz_c16;         /* RING: non-zeros whose column is streamed as a 16-bit window position (plan-time index
                               * compression; the CRS arrays themselves are untouched)                                */
    uint64_t stream_bytes;    /* bytes the planned kernel moves per product:
                               * algorithmic_bytes - (nnz_c16 + rows_o16)*(sizeof I - 2)                              */
    uint64_t rows_o16;        /* RING: rows whose offsets are streamed as 16-bit block-relative numbers               */
    uint64_t plan_bytes;      /* device memory the plan holds beside the three CRS arrays (split points, x windows,
                               * 16-bit columns / offsets, band parts)                                                */
    double plan_ms;           /* host wall-clock time the plan took to build (once per matrix and variant)            */
    uint64_t nnz_v8;          /* RING: non-zeros whose value is streamed as an 8-bit code into the block's dictionary of
                               * distinct values (plan-time value indexing: every block holds <= 256 distinct values —
                               * constant-coefficient stencils, unit-weight graphs; same operands, bit-identical results);
                               * stream_bytes then counts 1 byte per value plus one 256-entry dictionary per block      */
    uint64_t sell_entries;    /* RING, value-indexed: padded entries of the sliced-ELLPACK stage order (0: CRS order).  The
                               * compressed entries of a block are stored 32 rows at a time, entry j of those rows side by
                               * side, so a warp reads consecutive bytes; kept when the padding is below 20 %            */
} smb200_plan_info;

typedef struct {
    uint64_t iterations;      /* loop bodies executed (linearsolver.rs:41)                            */
    double final_residual;    /* sqrt(r.r) at exit, in f64 like linearsolver.rs:52                    */
    int32_t converged;        /* 1 if the stop test fired, 0 if iter_max was exhausted                */
    float device_ms;          /* CUDA-event time of the whole solve on the context stream             */
    uint64_t launches;        /* kernels launched by the solve                                        */
} smb200_cg_stats;

/* ---- library / context ------------------------------------------------------------------------ */
int32_t smb200_version(void);
/* Message of the last failing call on this thread. */
const char* smb200_last_error(void);
/* Number of kernels launched by this library on this thread's contexts since load (bench evidence). */
uint64_t smb200_launch_count(void);

/* device: CUDA ordinal.  stream: an existing cudaStream_t to run on (e.g. torch's current stream),
 * or NULL to let the context create its own non-blocking stream. */
smb200_status smb200_ctx_create(int32_t device, void* stream, smb200_ctx** out);
smb200_status smb200_ctx_destroy(smb200_ctx* ctx);
smb200_status smb200_ctx_sync(smb200_ctx* ctx);
smb200_status smb200_ctx_devinfo(smb200_ctx* ctx, smb200_devinfo* out);
/* Evict L2 by streaming a scratch buffer larger than the cache (benchmark hygiene). */
smb200_status smb200_ctx_flush_l2(smb200_ctx* ctx);

smb200_status smb200_event_create(smb200_ctx* ctx, smb200_event** out);
smb200_status smb200_event_record(smb200_event* ev);          /* on the context stream */
smb200_status smb200_event_elapsed_ms(smb200_event* start, smb200_event* stop, float* ms); /* syncs stop */
smb200_status smb200_event_destroy(smb200_event* ev);

/* Page-locked host memory for the end-to-end path. */
smb200_status smb200_host_alloc(size_t bytes, void** out);
smb200_status smb200_host_free(void* p);

/* ---- dense vectors: DenseVec<T> (densevec.rs:5-140) + trait Vector (vector.rs:5-64) ----------- */
/* Vector::with_capacity + zero fill; `n` is dim(). */
smb200_status smb200_vec_create(smb200_ctx* ctx, smb200_vtype vt, uint64_t n, smb200_vec** out);
/* Borrow caller-owned device memory (e.g. a torch tensor) as a vector; not freed by vec_free. */
smb200_status smb200_vec_wrap(smb200_ctx* ctx, smb200_vtype vt, uint64_t n, void* device_ptr, smb200_vec** out);
smb200_status smb200_vec_free(smb200_vec* v);
smb200_status smb200_vec_dim(const smb200_vec* v, uint64_t* n);                    /* densevec.rs:36-38 */
smb200_status smb200_vec_device_ptr(const smb200_vec* v, void** out);
/* DenseVec::from_vec (densevec.rs:30-34): copies n elements host -> device (n <= dim). */
smb200_status smb200_vec_upload(smb200_vec* v, const void* host, uint64_t n);
/* iter_ref().as_slice() (densevec.rs:10-12): copies n elements device -> host and synchronises. */
smb200_status smb200_vec_download(const smb200_vec* v, void* host, uint64_t n);
smb200_status smb200_vec_clone(const smb200_vec* v, smb200_vec** out);             /* #[derive(Clone)] */
smb200_status smb200_vec_copy(smb200_vec* dst, const smb200_vec* src);             /* dst[..src.dim] = src */
smb200_status smb200_vec_fill(smb200_vec* v, double value);
/* v[i] = (T)(2*u01(rng1(seed,i)) - 1): the benches' uniform [-1,1) inputs, generated on device. */
smb200_status smb200_vec_fill_uniform(smb200_vec* v, uint64_t seed);

/* densevec.rs:51-58  x += y   (ERR_DIM if x.dim < y.dim; only the first y.dim entries change). */
smb200_status smb200_vec_add(smb200_vec* x, const smb200_vec* y);
/* densevec.rs:60-67  x -= y. */
smb200_status smb200_vec_sub(smb200_vec* x, const smb200_vec* y);
/* densevec.rs:69-73  x *= s  (s is rounded to T first). */
smb200_status smb200_vec_scale(smb200_vec* x, double s);
/* y += (x * alpha): the reference's `*y += x.clone() * alpha` (linearsolver.rs:47,49) — the product
 * is rounded before the add (two roundings, no FMA) so results equal the reference bit for bit. */
smb200_status smb200_vec_axpy(smb200_vec* y, double alpha, const smb200_vec* x);
/* p = (p * beta) + r: `p.scale(beta); p.add(&r)` (linearsolver.rs:58-59), two roundings. */
smb200_status smb200_vec_scale_add(smb200_vec* p, double beta, const smb200_vec* r);
/* vector.rs:50-53  sum_i x_i*y_i over min(dim) entries.  Products are rounded to T, the sum is a
 * fixed-order tree (per-thread partials in T, combined in f64, result rounded to T): deterministic,
 * within 1e-5 (f32) / 1e-12 (f64) of the reference's sequential fold relative to sum|x_i*y_i|. */
smb200_status smb200_vec_dot(const smb200_vec* x, const smb200_vec* y, double* out);
smb200_status smb200_vec_norm2sq(const smb200_vec* x, double* out);                /* vector.rs:56-58 */
smb200_status smb200_vec_norm(const smb200_vec* x, double* out);                   /* vector.rs:61-63 */

/* ---- CRS matrix: SparseMatCRS<T,I> (sparsemat_crs.rs:9-17) ------------------------------------ */
/* Copy a finished CRS matrix to the device (values[nnz], columns[nnz], offset_rows[n_rows+1] in the
 * index type).  Validates what the reference would trip over at run time: offsets start at 0, are
 * non-decreasing and end at nnz; every column < n_cols.  Within-row order is kept as given. */
smb200_status smb200_crs_upload(smb200_ctx* ctx, smb200_vtype vt, smb200_itype it, uint64_t n_rows,
                                uint64_t n_cols, uint64_t nnz, const void* values, const void* columns,
                                const void* offset_rows, smb200_crs** out);
/* SparseMatIndexList::to_crs (sparsemat_indexlist.rs:61-63 -> sparsemat_crs.rs:24-50) on the device:
 * rows ascending, chain (= insertion) order inside a row, empty rows repeat the offset, nnz == 0
 * gives the 0x0 matrix.  Inputs are the IndexList arrays: columns[nnz], values[nnz],
 * pos_start[n_rows] and index_list[nnz] with I::MAX as the UNSET sentinel (indexlist.rs:26-33). */
smb200_status smb200_crs_from_indexlist(smb200_ctx* ctx, smb200_vtype vt, smb200_itype it, uint64_t n_rows,
                                        uint64_t n_cols, uint64_t nnz, const void* columns, const void* values,
                                        const void* pos_start, const void* index_list, smb200_crs** out);
smb200_status smb200_crs_free(smb200_crs* m);
/* Binary CRS container (additive; the reference only writes text/PBM, sparsematrix.rs:304-338, and reads nothing):
 * a 56-byte header (magic "SMBCRS01", value/index type, n_rows, n_cols, nnz, FNV-1a 64 of the payload) followed by
 * offset_rows[n_rows+1], columns[nnz], values[nnz] exactly as SparseMatCRS holds them (sparsemat_crs.rs:9-17), each
 * zero-padded to 8 bytes; little endian.  ERR_IO on open/read/write failures, foreign or truncated files and checksum
 * mismatches.  crsfile_* are host-only (no device needed): info fills vt, it and out3 = {n_rows, n_cols, nnz} so the
 * caller can size the buffers for crsfile_read, which takes their capacities in bytes and refuses (ERR_INVALID) a file
 * that needs more — the file may have changed in between.  The header's sizes are checked against the file's size
 * before anything is allocated or read.  crs_save downloads and writes; crs_load reads (one open) and uploads (with
 * the validation of smb200_crs_upload). */
smb200_status smb200_crsfile_write(const char* path, smb200_vtype vt, smb200_itype it, uint64_t n_rows, uint64_t n_cols,
                                   uint64_t nnz, const void* values, const void* columns, const void* offset_rows);
smb200_status smb200_crsfile_info(const char* path, int32_t* vt, int32_t* it, uint64_t* out3);
smb200_status smb200_crsfile_read(const char* path, void* values, uint64_t values_cap_bytes, void* columns,
                                  uint64_t columns_cap_bytes, void* offset_rows, uint64_t offsets_cap_bytes);
smb200_status smb200_crs_save(const smb200_crs* m, const char* path);
smb200_status smb200_crs_load(smb200_ctx* ctx, const char* path, smb200_crs** out);
/* out3 = {n_rows, n_cols, n_non_zero_entries} (sparsemat_crs.rs:124-134). */
smb200_status smb200_crs_dims(const smb200_crs* m, uint64_t* out3);
smb200_status smb200_crs_types(const smb200_crs* m, int32_t* vt, int32_t* it);
/* Copy the device arrays back (bit-exact layout checks).  Any pointer may be NULL. */
smb200_status smb200_crs_download(const smb200_crs* m, void* values, void* columns, void* offset_rows);
smb200_status smb200_crs_scale(smb200_crs* m, double s);                           /* sparsemat_crs.rs:153-157 */
/* Choose the kernel family (benchmarks / tests); lanes is used by VECTOR (0 = from mean row length). */
smb200_status smb200_crs_configure(smb200_crs* m, smb200_spmv_variant variant, int32_t lanes, uint32_t flags);
smb200_status smb200_crs_plan_info(const smb200_crs* m, smb200_plan_info* out);

/* Synthetic BASELINE.json workloads generated directly on the device.
 * Dirichlet Laplacian on an nx*ny*nz grid (nz == 1: 2-D 5-point, else 3-D 7-point), rows
 * [row_lo,row_hi) of the global operator with global columns; ascending columns inside a row. */
smb200_status smb200_gen_laplace(smb200_ctx* ctx, smb200_vtype vt, smb200_itype it, uint64_t nx, uint64_t ny,
                                 uint64_t nz, uint64_t row_lo, uint64_t row_hi, smb200_crs** out);
/* Power-law rows: L_i = clamp(floor(8/sqrt(u_i)), 1, max_len), uniform random columns and values. */
smb200_status smb200_gen_powerlaw(smb200_ctx* ctx, smb200_vtype vt, smb200_itype it, uint64_t n_rows,
                                  uint64_t n_cols, uint64_t seed_len, uint64_t seed_col, uint64_t seed_val,
                                  uint64_t max_len, smb200_crs** out);

/* ---- the hot path ------------------------------------------------------------------------------ */
/* SparseMatrix::mvp (sparsematrix.rs:146-158): y[0..n_rows) = A x.  ERR_DIM unless x.dim >= n_cols
 * (the reference would panic in rhs.get) and y.dim >= n_rows. */
smb200_status smb200_spmv(smb200_crs* a, const smb200_vec* x, smb200_vec* y);
/* Same through host buffers (pinned or pageable): H2D x, SpMV, D2H y, synchronised — the call the
 * `Mul<DenseVec<T>>` operator (sparsematrix.rs:435-443) maps to when the vectors live on the host. */
smb200_status smb200_spmv_host(smb200_crs* a, const void* x_host, uint64_t nx, void* y_host);
/* SparseMatrix::transpose (sparsematrix.rs:174-183) on the device: row j of the result holds the entries of column j
 * ordered by source row (the order `ret.set(col, i, val)` appends them in on the assembly format, then `to_crs()`);
 * n_rows = largest column + 1, n_cols = last non-empty row + 1, 0 x 0 without entries.  Bit-exact (integer / copy
 * work: a stable radix sort of the entries by column).  A duplicate (i, j) — which the reference's assembly cannot
 * produce — stays two entries. */
smb200_status smb200_crs_transpose(const smb200_crs* a, smb200_crs** out);
/* SparseMatrix::inner_prod (sparsematrix.rs:161-171): lhs^T A rhs in one pass. */
smb200_status smb200_bilinear(smb200_crs* a, const smb200_vec* lhs, const smb200_vec* rhs, double* out);

/* ConjugateGradient::solve (linearsolver.rs:27-61).  tol/iter_max are the struct's private fields
 * (defaults 1e-12 / 10000, :17-24).  relative = 0 is the reference's absolute test sqrt(r.r) < tol;
 * relative = 1 (additive) tests sqrt(r.r) < tol * ||b||.  x is updated in place. */
smb200_status smb200_cg_solve(smb200_crs* a, const smb200_vec* b, smb200_vec* x, double tol, int32_t relative,
                              uint64_t iter_max, smb200_cg_stats* stats);
/* Optional: copy out up to `cap` residual norms (one per iteration) of the last solve on this matrix. */
smb200_status smb200_cg_history(const smb200_crs* a, double* out, uint64_t cap, uint64_t* n);
/* Additive, not in the reference (its LinearSolver trait, linearsolver.rs:6-10, has the unpreconditioned solver only).
 * crs_diagonal: d[i] = the first stored entry of row i with column i — what get(i, i) returns (sparsemat_crs.rs:136-143) —
 * or 0.  pcg_jacobi_solve: CG preconditioned with the inverse diagonal; same argument meaning, checks and panics as
 * smb200_cg_solve; ERR_INVALID when a row has no (or a zero) diagonal entry. */
smb200_status smb200_crs_diagonal(const smb200_crs* a, smb200_vec* d);
smb200_status smb200_pcg_jacobi_solve(smb200_crs* a, const smb200_vec* b, smb200_vec* x, double tol, int32_t relative,
                                      uint64_t iter_max, smb200_cg_stats* stats);

/* ---- host-side partitioning (no GPU needed): SparseMatPar's contract (sparsemat_par.rs:20-35) -- */
/* R = max_n_rows / n_blocks; row -> (min(row / R, n_blocks), row - block*R).  ERR_INVALID on R == 0. */
smb200_status smb200_par_locate(uint64_t n_blocks, uint64_t max_n_rows, uint64_t row, uint64_t* block,
                                uint64_t* local_row);
/* SparseMatPar (sparsemat_par.rs:12-35) with device-resident blocks and the `mvp_par` the reference left unfinished
 * (sparsemat_par.rs:37-68).  Block b = global rows [b R, (b+1) R), R = max_n_rows / n_blocks, local row ids, GLOBAL column
 * ids.  With a communicator (smb200_comm_init) block b lives on rank b * world / n_blocks and every rank makes the same
 * calls; set_block_* read their arrays only on the block's owner.  ERR_INVALID for n_blocks == 0 or R == 0 (the reference
 * divides by zero there).  par_mvp: x is the whole right-hand side on every rank (the sketch's `Arc<rhs>`), y is complete
 * on every rank afterwards (the channel gather); a short block in front of later rows is ERR_INVALID "index out of
 * bounds" (the reference's default mvp panics in IndexList::iter_row, indexlist.rs:88).  out3 of par_dims = {n_rows
 * (scan up to the first empty block, sparsemat_par.rs:95-107), n_cols, non-zeros}.  par_block lends the device block
 * (NULL when it is empty or lives on another rank). */
typedef struct smb200_par smb200_par;
smb200_status smb200_par_create(smb200_ctx* ctx, uint64_t n_blocks, uint64_t max_n_rows, smb200_vtype vt, smb200_itype it,
                                smb200_par** out);
smb200_status smb200_par_free(smb200_par* p);
smb200_status smb200_par_owner(const smb200_par* p, uint64_t block, int32_t* rank);
smb200_status smb200_par_set_block_indexlist(smb200_par* p, uint64_t block, uint64_t n_rows, uint64_t n_cols, uint64_t nnz,
                                             const void* columns, const void* values, const void* pos_start,
                                             const void* index_list);
smb200_status smb200_par_set_block_crs(smb200_par* p, uint64_t block, uint64_t n_rows, uint64_t n_cols, uint64_t nnz,
                                       const void* values, const void* columns, const void* offset_rows);
smb200_status smb200_par_dims(const smb200_par* p, uint64_t* out3);
smb200_status smb200_par_block(smb200_par* p, uint64_t block, smb200_crs** out);
smb200_status smb200_par_mvp(smb200_par* p, const smb200_vec* x, smb200_vec* y);
/* Contiguous row ranges for `world` ranks: rows_per_rank = ceil(n_rows / world) rounded up to a
 * multiple of `align` (e.g. one z-plane); out_bounds has world+1 entries. */
smb200_status smb200_partition_rows(uint64_t n_rows, uint32_t world, uint64_t align, uint64_t* out_bounds);
/* nnz-balanced split points from the offsets array (binary search), for irregular matrices. */
smb200_status smb200_partition_rows_by_nnz(smb200_itype it, uint64_t n_rows, const void* offset_rows,
                                           uint32_t world, uint64_t* out_bounds);
/* Ghost plan of one rank: given its local block (local row offsets, GLOBAL columns) and the row
 * bounds of all ranks, produce (a) columns remapped to local numbering [owned | ghosts sorted by
 * global id], (b) the sorted ghost list, (c) per-owner counts.  Two-call protocol: call with
 * ghosts == NULL to get *n_ghosts, then again with storage. */
smb200_status smb200_ghost_plan(smb200_itype it, uint64_t nnz, const void* columns_global, uint32_t world,
                                uint32_t rank, const uint64_t* bounds, void* columns_local_out,
                                uint64_t* ghosts, uint64_t* n_ghosts, uint64_t* ghosts_per_owner);

/* ---- one process per GPU: row-block distributed SpMV / CG ------------------------
 * Replaces the reference's unfinished thread-per-block mvp_par (sparsemat_par.rs:37-68): same partition
 * contract (sparsemat_par.rs:20-35), one rank per GPU instead of one thread per block.  NCCL carries the
 * set-up only; products and CG scalars move through peer memory over NVLink (CUDA IPC), with NCCL
 * send/recv + all-reduce as the fallback (SMB200_DIST_P2P=0). */
/* 128-byte NCCL unique id, created on rank 0 and broadcast by the host (torch.distributed, MPI, ...). */
smb200_status smb200_comm_unique_id(void* out128);
smb200_status smb200_comm_init(smb200_ctx* ctx, int32_t rank, int32_t world, const void* uid128);
smb200_status smb200_comm_destroy(smb200_ctx* ctx);
/* Local block of a row-partitioned matrix: local offsets, GLOBAL columns; bounds[world+1] as above.
 * Builds the ghost plan, remaps columns and uploads.  Collective over the communicator. */
smb200_status smb200_dist_create(smb200_ctx* ctx, smb200_vtype vt, smb200_itype it, uint64_t n_global,
                                 const uint64_t* bounds, uint64_t nnz_local, const void* values,
                                 const void* columns_global, const void* offset_rows_local, smb200_dist** out);
/* z-slab partition of the global nx*ny*nz Laplacian generated on the device (no host arrays). */
smb200_status smb200_dist_laplace(smb200_ctx* ctx, smb200_vtype vt, smb200_itype it, uint64_t nx, uint64_t ny,
                                  uint64_t nz, smb200_dist** out);
smb200_status smb200_dist_free(smb200_dist* d);
/* out4 = {n_local_rows, n_ghosts, nnz_local, row_lo}. */
smb200_status smb200_dist_dims(const smb200_dist* d, uint64_t* out4);
smb200_status smb200_dist_local(smb200_dist* d, smb200_crs** out);   /* borrowed local matrix handle */
/* A vector slice of this rank: dim = n_local_rows, with hidden room for the ghost entries. */
smb200_status smb200_dist_vec_create(smb200_dist* d, smb200_vec** out);
/* y_local = (A x)_local.  Every rank must issue the same sequence of distributed calls.  The ghost
 * entries of x are stored by the neighbours' kernels into this rank's ghost buffer while the interior
 * rows are multiplied; the boundary rows follow in the same launch (ring kernel) or right behind. */
smb200_status smb200_dist_spmv(smb200_dist* d, smb200_vec* x, smb200_vec* y);
smb200_status smb200_dist_dot(smb200_dist* d, const smb200_vec* x, const smb200_vec* y, double* out);
/* Stream-ordered barrier over the ranks (no host synchronisation): work queued behind it on the
 * context stream starts only after every rank's stream has reached its own barrier. */
smb200_status smb200_dist_barrier(smb200_dist* d);
/* out6 = {1 if the peer-memory path is active (0: NCCL fallback), halo neighbours of this rank,
 * distributed products completed, 1 if a peer wait timed out, sum of the spin times of all threads that
 * waited for a neighbour's flag in ns, number of such waits}.  Synchronises the context stream. */
smb200_status smb200_dist_info(smb200_dist* d, uint64_t* out6);
/* ConjugateGradient::solve (linearsolver.rs:27-61) over the row blocks: the reference's loop, two all-reduces
 * (p.Ap, r.r) per iteration, both inside the kernels that produce their operands on the peer-memory path. */
smb200_status smb200_dist_cg_solve(smb200_dist* d, const smb200_vec* b, smb200_vec* x, double tol, int32_t relative,
                                   uint64_t iter_max, smb200_cg_stats* stats);
/* Additive: the same solve with the loop rearranged (Chronopoulos-Gear) so that r.r and (A r).r come out of
 * ONE all-reduce per iteration; s = A p and alpha follow by recurrence.  Same arguments, stop test, history
 * and vector traffic; the iteration count may differ from the reference's loop by a few. */
smb200_status smb200_dist_cg_solve_sr(smb200_dist* d, const smb200_vec* b, smb200_vec* x, double tol, int32_t relative,
                                      uint64_t iter_max, smb200_cg_stats* stats);

#ifdef __cplusplus
}
#endif
#endif /* SMB200_H */
