// Host-only partitioning logic (no CUDA calls): SparseMatPar's row-block contract
// (sparsemat_par.rs:20-35) and the ghost plan of the row-partitioned multi-GPU path.
// Lives in the same library so that world_size > 1 logic can be tested on CPU boxes.
#include <algorithm>
#include <cstdint>
#include <vector>

#include "../../include/smb200.h"

namespace smb {
void set_error(const char* fmt, ...);
}

#define PART_REQUIRE(cond, ...)              \
    do {                                     \
        if (!(cond)) {                       \
            ::smb::set_error(__VA_ARGS__);   \
            return SMB200_ERR_INVALID;       \
        }                                    \
    } while (0)

namespace {

template <class I>
smb200_status ghost_plan_impl(uint64_t nnz, const I* cols, uint32_t world, uint32_t rank, const uint64_t* bounds,
                              I* cols_local, uint64_t* ghosts, uint64_t* n_ghosts, uint64_t* per_owner) {
    const uint64_t lo = bounds[rank], hi = bounds[rank + 1], n_local = hi - lo, n_global = bounds[world];
    std::vector<uint64_t> ext;
    for (uint64_t k = 0; k < nnz; ++k) {
        const uint64_t c = (uint64_t)cols[k];
        if (c >= n_global) {
            ::smb::set_error("ghost_plan: column %llu >= global size %llu", (unsigned long long)c, (unsigned long long)n_global);
            return SMB200_ERR_INVALID;
        }
        if (c < lo || c >= hi) ext.push_back(c);
    }
    std::sort(ext.begin(), ext.end());
    ext.erase(std::unique(ext.begin(), ext.end()), ext.end());
    *n_ghosts = ext.size();
    if (per_owner) {
        for (uint32_t q = 0; q < world; ++q) {
            auto b = std::lower_bound(ext.begin(), ext.end(), bounds[q]);
            auto e = std::lower_bound(ext.begin(), ext.end(), bounds[q + 1]);
            per_owner[q] = (uint64_t)(e - b);
        }
    }
    if (!ghosts) return SMB200_OK;
    std::copy(ext.begin(), ext.end(), ghosts);
    if (cols_local) {
        for (uint64_t k = 0; k < nnz; ++k) {
            const uint64_t c = (uint64_t)cols[k];
            if (c >= lo && c < hi) cols_local[k] = (I)(c - lo);
            else cols_local[k] = (I)(n_local + (uint64_t)(std::lower_bound(ext.begin(), ext.end(), c) - ext.begin()));
        }
    }
    return SMB200_OK;
}

template <class I>
void split_by_nnz(uint64_t n_rows, const I* off, uint32_t world, uint64_t* out) {
    const uint64_t nnz = n_rows ? (uint64_t)off[n_rows] : 0;
    out[0] = 0;
    for (uint32_t q = 1; q < world; ++q) {
        const uint64_t want = (uint64_t)((__uint128_t)nnz * q / world);
        const I* it = std::lower_bound(off, off + n_rows + 1, (I)want);
        uint64_t r = (uint64_t)(it - off);
        if (r > n_rows) r = n_rows;
        out[q] = std::max(r, out[q - 1]);
    }
    out[world] = n_rows;
}

}  // namespace

extern "C" {

smb200_status smb200_par_locate(uint64_t n_blocks, uint64_t max_n_rows, uint64_t row, uint64_t* block, uint64_t* local_row) {
    PART_REQUIRE(block && local_row, "par_locate: NULL output");
    PART_REQUIRE(n_blocks > 0, "par_locate: n_blocks == 0 (the reference divides by zero here)");
    const uint64_t r = max_n_rows / n_blocks;                       // sparsemat_par.rs:21
    PART_REQUIRE(r > 0, "par_locate: rows per block == 0 (the reference divides by zero here)");
    const uint64_t b = std::min<uint64_t>(row / r, n_blocks);       // :32 — clamps to n_blocks, as written
    *block = b;
    *local_row = row - b * r;                                        // :33
    return SMB200_OK;
}

smb200_status smb200_partition_rows(uint64_t n_rows, uint32_t world, uint64_t align, uint64_t* out_bounds) {
    PART_REQUIRE(out_bounds && world > 0, "partition_rows: bad arguments");
    if (align == 0) align = 1;
    uint64_t per = (n_rows + world - 1) / world;
    per = (per + align - 1) / align * align;
    for (uint32_t q = 0; q <= world; ++q) out_bounds[q] = std::min<uint64_t>(n_rows, (uint64_t)q * per);
    out_bounds[world] = n_rows;
    return SMB200_OK;
}

smb200_status smb200_partition_rows_by_nnz(smb200_itype it, uint64_t n_rows, const void* offset_rows, uint32_t world,
                                           uint64_t* out_bounds) {
    PART_REQUIRE(out_bounds && world > 0 && (offset_rows || n_rows == 0), "partition_rows_by_nnz: bad arguments");
    if (n_rows == 0) { for (uint32_t q = 0; q <= world; ++q) out_bounds[q] = 0; return SMB200_OK; }
    if (it == SMB200_U64) split_by_nnz<uint64_t>(n_rows, (const uint64_t*)offset_rows, world, out_bounds);
    else split_by_nnz<uint32_t>(n_rows, (const uint32_t*)offset_rows, world, out_bounds);
    return SMB200_OK;
}

smb200_status smb200_ghost_plan(smb200_itype it, uint64_t nnz, const void* columns_global, uint32_t world, uint32_t rank,
                                const uint64_t* bounds, void* columns_local_out, uint64_t* ghosts, uint64_t* n_ghosts,
                                uint64_t* ghosts_per_owner) {
    PART_REQUIRE(bounds && n_ghosts && world > 0 && rank < world && (columns_global || nnz == 0), "ghost_plan: bad arguments");
    for (uint32_t q = 0; q < world; ++q) PART_REQUIRE(bounds[q] <= bounds[q + 1], "ghost_plan: bounds not monotone");
    if (it == SMB200_U64)
        return ghost_plan_impl<uint64_t>(nnz, (const uint64_t*)columns_global, world, rank, bounds, (uint64_t*)columns_local_out, ghosts, n_ghosts, ghosts_per_owner);
    return ghost_plan_impl<uint32_t>(nnz, (const uint32_t*)columns_global, world, rank, bounds, (uint32_t*)columns_local_out, ghosts, n_ghosts, ghosts_per_owner);
}

}  // extern "C"
