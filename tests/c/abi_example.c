/* Plain C99 against include/smb200.h only: what any FFI binding (Rust `extern "C"`, cgo, JNI, ctypes) does.
 * Replays the reference's known answers (lib.rs:36-52 CG -> 0.0909; lib.rs:80-82 mvp -> 34.544; lib.rs:150-152 -> 20.16)
 * through the raw C ABI, including the IndexList -> CRS conversion entry point.  Exit code 0 = all checks passed. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "smb200.h"

static int fails = 0;
#define CHECK(c) do { if (!(c)) { ++fails; printf("FAIL %s:%d %s  (%s)\n", __FILE__, __LINE__, #c, smb200_last_error()); } } while (0)
#define OK(call) CHECK((call) == SMB200_OK)

int main(void) {
    smb200_ctx* ctx = NULL;
    if (smb200_ctx_create(0, NULL, &ctx) != SMB200_OK) { printf("no device: %s\n", smb200_last_error()); return 2; }

    /* lib.rs:57-66 assembled in an IndexList; arrays as indexlist.rs:26-29 lays them out (UNSET = u32::MAX) */
    const uint32_t U = 0xFFFFFFFFu;
    const uint32_t columns[6] = {1, 2, 2, 1, 2, 0};
    const float values[6] = {4.2f, 4.12f, 2.12f, 1.12f + 1.12f, 0.12f, 7.12f};
    const uint32_t pos_start[3] = {0, 1, 2};
    const uint32_t index_list[6] = {4, 3, U, U, 5, U};
    smb200_crs* a = NULL;
    OK(smb200_crs_from_indexlist(ctx, SMB200_F32, SMB200_U32, 3, 3, 6, columns, values, pos_start, index_list, &a));
    uint32_t oc[6], oo[4];
    float ov[6];
    OK(smb200_crs_download(a, ov, oc, oo));
    const uint32_t want_c[6] = {1, 2, 0, 2, 1, 2}, want_o[4] = {0, 3, 5, 6};
    CHECK(memcmp(oc, want_c, sizeof oc) == 0 && memcmp(oo, want_o, sizeof oo) == 0);      /* chain order kept */
    const float xh[3] = {2.0f, 4.8f, 1.2f};
    float yh[3] = {0, 0, 0};
    OK(smb200_spmv_host(a, xh, 3, yh));
    CHECK(yh[0] == 34.544f);                                                              /* lib.rs:80-82 */
    smb200_vec *x = NULL, *y = NULL;
    OK(smb200_vec_create(ctx, SMB200_F32, 3, &x));
    OK(smb200_vec_create(ctx, SMB200_F32, 3, &y));
    OK(smb200_vec_upload(x, xh, 3));
    OK(smb200_spmv(a, x, y));
    OK(smb200_vec_download(y, yh, 3));
    CHECK(yh[0] == 34.544f);
    smb200_vec* shorty = NULL;
    OK(smb200_vec_create(ctx, SMB200_F32, 2, &shorty));
    CHECK(smb200_spmv(a, shorty, y) == SMB200_ERR_DIM);                                   /* rhs.get(2) would panic */
    double d = 0.0;
    OK(smb200_vec_dot(x, x, &d));
    CHECK(fabs(d - (double)(2.0f * 2.0f + 4.8f * 4.8f + 1.2f * 1.2f)) < 1e-5);

    /* lib.rs:116-128 state of the direct CRS test; mvp row 0 == 20.16 */
    const float v2[5] = {4.2f, 4.12f, 2.12f, 5.12f, 1.12f};
    const uint32_t c2[5] = {1, 2, 2, 3, 2}, o2[5] = {0, 1, 2, 3, 5};
    smb200_crs* b = NULL;
    OK(smb200_crs_upload(ctx, SMB200_F32, SMB200_U32, 4, 4, 5, v2, c2, o2, &b));
    const float x4[4] = {2.0f, 4.8f, 1.2f, 3.4f};
    float y4[4];
    OK(smb200_spmv_host(b, x4, 4, y4));
    CHECK(y4[0] == 20.16f);                                                               /* lib.rs:150-152 */
    const uint32_t bad_o[5] = {0, 2, 1, 3, 5};
    smb200_crs* bad = NULL;
    CHECK(smb200_crs_upload(ctx, SMB200_F32, SMB200_U32, 4, 4, 5, v2, c2, bad_o, &bad) == SMB200_ERR_INVALID);

    /* lib.rs:36-52: [[4,1],[1,3]] x = [1,2], x0 = [2,1], default tolerances */
    const double v3[4] = {4.0, 1.0, 1.0, 3.0};
    const uint32_t c3[4] = {0, 1, 0, 1}, o3[3] = {0, 2, 4};
    smb200_crs* m = NULL;
    OK(smb200_crs_upload(ctx, SMB200_F64, SMB200_U32, 2, 2, 4, v3, c3, o3, &m));
    smb200_vec *rhs = NULL, *sol = NULL;
    const double bh[2] = {1.0, 2.0};
    double sh[2] = {2.0, 1.0};
    OK(smb200_vec_create(ctx, SMB200_F64, 2, &rhs));
    OK(smb200_vec_create(ctx, SMB200_F64, 2, &sol));
    OK(smb200_vec_upload(rhs, bh, 2));
    OK(smb200_vec_upload(sol, sh, 2));
    smb200_cg_stats st;
    OK(smb200_cg_solve(m, rhs, sol, 1e-12, 0, 10000, &st));
    OK(smb200_vec_download(sol, sh, 2));
    CHECK(floor(sh[0] * 10000.0) / 10000.0 == 0.0909);
    CHECK(st.converged == 1 && st.iterations <= 3);
    CHECK(smb200_cg_solve(b, rhs, sol, 1e-12, 0, 10, &st) != SMB200_OK);                  /* value types differ */
    uint64_t blk = 0, loc = 0;
    OK(smb200_par_locate(4, 16, 9, &blk, &loc));
    CHECK(blk == 2 && loc == 1);                                                          /* sparsemat_par.rs:31-35 */

    smb200_vec_free(x); smb200_vec_free(y); smb200_vec_free(shorty); smb200_vec_free(rhs); smb200_vec_free(sol);
    smb200_crs_free(a); smb200_crs_free(b); smb200_crs_free(m);
    smb200_ctx_destroy(ctx);
    printf("abi_example: %s (%llu kernels launched)\n", fails ? "FAILED" : "ok", (unsigned long long)smb200_launch_count());
    return fails ? 1 : 0;
}
