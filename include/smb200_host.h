/*
 * smb200_host.h — host-side assembly helper for language mirrors without the reference crate
 * (Python, C).  NOT part of the drop-in boundary: a Rust binding keeps the crate's own
 * SparseMatIndexList and hands its arrays to smb200_crs_from_indexlist (smb200.h).  This is the same
 * assembly format (sparsemat_indexlist.rs:14-21, indexlist.rs:26-29) implemented in
 * sparsemat_b200/host/sparsemat.hpp, exported so that ctypes callers can assemble on the host and
 * convert on the device.  No GPU is needed except for smb200_il_to_crs.
 */
#ifndef SMB200_HOST_H
#define SMB200_HOST_H
#include "smb200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct smb200_il smb200_il;

smb200_status smb200_il_create(smb200_vtype vt, smb200_itype it, smb200_il** out);
smb200_status smb200_il_free(smb200_il* il);
/* For k in 0..n (in order): op == 0 -> set(i[k], j[k], v[k]) (sparsematrix.rs:226-228),
 * op == 1 -> add_to (sparsematrix.rs:231-233).  v has the matrix value type. */
smb200_status smb200_il_apply(smb200_il* il, uint64_t n, const uint64_t* i, const uint64_t* j, const void* v, int32_t op);
smb200_status smb200_il_get(const smb200_il* il, uint64_t i, uint64_t j, double* out);
/* out3 = {n_rows, n_cols, n_non_zero_entries} */
smb200_status smb200_il_dims(const smb200_il* il, uint64_t* out3);
/* Raw arrays: columns[nnz], values[nnz], pos_start[n_rows], index_list[nnz] (I::MAX = UNSET). */
smb200_status smb200_il_export(const smb200_il* il, void* columns, void* values, void* pos_start, void* index_list);
/* SparseMatIndexList::to_crs (sparsemat_indexlist.rs:61-63) through smb200_crs_from_indexlist. */
smb200_status smb200_il_to_crs(const smb200_il* il, smb200_ctx* ctx, smb200_crs** out);

#ifdef __cplusplus
}
#endif
#endif
