// Binary CRS container: the reference only has text/PBM writers (sparsematrix.rs:304-338) and no
// reader, so a matrix assembled once (IndexList -> to_crs) cannot be reused by a later run.  This is the smallest
// format that can: the three arrays of SparseMatCRS (sparsemat_crs.rs:9-17) byte for byte, in-row order untouched,
// behind a fixed header with a checksum.  Host-only halves (smb200_crsfile_*) need no device; smb200_crs_save / _load
// put smb200_crs_download / smb200_crs_upload (which validates the arrays like any other upload) around them.
//
//   offset  size  field
//   0       8     magic "SMBCRS01"
//   8       4     value type  (smb200_vtype, little endian)
//   12      4     index type  (smb200_itype)
//   16      8     n_rows
//   24      8     n_cols
//   32      8     nnz
//   40      8     FNV-1a 64 of the payload (offsets, columns, values as stored, padding excluded)
//   48      8     reserved, 0
//   56      ...   offset_rows[n_rows + 1] (index type; absent when n_rows == 0), zero-padded to a multiple of 8 bytes
//                 columns[nnz]            (index type), zero-padded to 8
//                 values[nnz]             (value type), zero-padded to 8
#include "common.cuh"

#include <sys/stat.h>

#include <cerrno>
#include <cstdio>
#include <memory>

namespace smb {
namespace {

constexpr char kMagic[8] = {'S', 'M', 'B', 'C', 'R', 'S', '0', '1'};
constexpr size_t kHeaderBytes = 56;

struct FileCloser { void operator()(FILE* f) const { if (f) fclose(f); } };
using File = std::unique_ptr<FILE, FileCloser>;

struct Header {
    uint32_t vt = 0, it = 0;
    uint64_t n_rows = 0, n_cols = 0, nnz = 0, checksum = 0;
};

uint64_t fnv1a(uint64_t h, const void* data, size_t n) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 0x100000001b3ull; }
    return h;
}
constexpr uint64_t kFnvBasis = 0xcbf29ce484222325ull;

void put_u32(unsigned char* p, uint32_t v) { for (int i = 0; i < 4; ++i) p[i] = (unsigned char)(v >> (8 * i)); }
void put_u64(unsigned char* p, uint64_t v) { for (int i = 0; i < 8; ++i) p[i] = (unsigned char)(v >> (8 * i)); }
uint32_t get_u32(const unsigned char* p) { uint32_t v = 0; for (int i = 0; i < 4; ++i) v |= (uint32_t)p[i] << (8 * i); return v; }
uint64_t get_u64(const unsigned char* p) { uint64_t v = 0; for (int i = 0; i < 8; ++i) v |= (uint64_t)p[i] << (8 * i); return v; }

size_t pad8(size_t n) { return (8 - (n & 7)) & 7; }

smb200_status read_header(FILE* f, const char* path, Header* h) {
    unsigned char raw[kHeaderBytes];
    SMB_REQUIRE(fread(raw, 1, kHeaderBytes, f) == kHeaderBytes, SMB200_ERR_IO, "crsfile: %s is shorter than a header", path);
    SMB_REQUIRE(memcmp(raw, kMagic, 8) == 0, SMB200_ERR_IO, "crsfile: %s is not a SMBCRS01 file", path);
    h->vt = get_u32(raw + 8);
    h->it = get_u32(raw + 12);
    h->n_rows = get_u64(raw + 16);
    h->n_cols = get_u64(raw + 24);
    h->nnz = get_u64(raw + 32);
    h->checksum = get_u64(raw + 40);
    SMB_REQUIRE(h->vt == SMB200_F32 || h->vt == SMB200_F64, SMB200_ERR_IO, "crsfile: %s: unknown value type %u", path, h->vt);
    SMB_REQUIRE(h->it == SMB200_U32 || h->it == SMB200_U64, SMB200_ERR_IO, "crsfile: %s: unknown index type %u", path, h->it);
    // sizes a 64-bit byte count can hold, and an index type that can hold them (offset_rows is Vec<I>)
    SMB_REQUIRE(h->nnz < (1ull << 59) && h->n_rows < (1ull << 59), SMB200_ERR_IO, "crsfile: %s: implausible dimensions", path);
    SMB_REQUIRE(h->it == SMB200_U64 || h->nnz <= 0xFFFFFFFFull, SMB200_ERR_IO, "crsfile: %s: nnz does not fit the u32 offsets", path);
    // the sizes the header claims must be the sizes the file has: nothing is allocated or read on the word of a corrupt header
    struct stat sb;
    SMB_REQUIRE(fstat(fileno(f), &sb) == 0, SMB200_ERR_IO, "crsfile: cannot stat %s: %s", path, strerror(errno));
    auto padded = [](uint64_t n) { return (n + 7) & ~(uint64_t)7; };
    const uint64_t is = h->it == SMB200_U64 ? 8 : 4, vs = h->vt == SMB200_F64 ? 8 : 4;
    const uint64_t want = kHeaderBytes + (h->n_rows ? padded((h->n_rows + 1) * is) : 0) + padded(h->nnz * is) + padded(h->nnz * vs);
    SMB_REQUIRE((uint64_t)sb.st_size == want, SMB200_ERR_IO, "crsfile: %s holds %llu bytes, its header describes %llu (truncated or corrupt)",
                path, (unsigned long long)sb.st_size, (unsigned long long)want);
    return SMB200_OK;
}

// Body of an opened file whose header was parsed and checked against the file size; capacities in bytes.
smb200_status read_body(FILE* f, const char* path, const Header& h, void* values, uint64_t values_cap, void* columns,
                        uint64_t columns_cap, void* offset_rows, uint64_t offsets_cap);

smb200_status write_block(FILE* f, const char* path, const void* data, size_t n, uint64_t* sum) {
    static const unsigned char zeros[8] = {0};
    if (n) {
        SMB_REQUIRE(fwrite(data, 1, n, f) == n, SMB200_ERR_IO, "crsfile: short write to %s: %s", path, strerror(errno));
        *sum = fnv1a(*sum, data, n);
    }
    const size_t p = pad8(n);
    if (p) SMB_REQUIRE(fwrite(zeros, 1, p, f) == p, SMB200_ERR_IO, "crsfile: short write to %s: %s", path, strerror(errno));
    return SMB200_OK;
}

smb200_status read_block(FILE* f, const char* path, void* data, size_t n, uint64_t* sum) {
    unsigned char skip[8];
    if (n) {
        SMB_REQUIRE(fread(data, 1, n, f) == n, SMB200_ERR_IO, "crsfile: %s is truncated", path);
        *sum = fnv1a(*sum, data, n);
    }
    const size_t p = pad8(n);
    if (p) SMB_REQUIRE(fread(skip, 1, p, f) == p, SMB200_ERR_IO, "crsfile: %s is truncated", path);
    return SMB200_OK;
}

smb200_status read_body(FILE* f, const char* path, const Header& h, void* values, uint64_t values_cap, void* columns,
                        uint64_t columns_cap, void* offset_rows, uint64_t offsets_cap) {
    const uint64_t ob = h.n_rows ? (h.n_rows + 1) * isize((int)h.it) : 0, cb = h.nnz * isize((int)h.it), vb = h.nnz * vsize((int)h.vt);
    SMB_REQUIRE((values && columns) || h.nnz == 0, SMB200_ERR_INVALID, "crsfile_read: NULL values/columns");
    SMB_REQUIRE(offset_rows || h.n_rows == 0, SMB200_ERR_INVALID, "crsfile_read: NULL offset_rows");
    SMB_REQUIRE(values_cap >= vb && columns_cap >= cb && offsets_cap >= ob, SMB200_ERR_INVALID,
                "crsfile_read: %s needs %llu / %llu / %llu bytes (values / columns / offset_rows), the buffers hold %llu / %llu / %llu",
                path, (unsigned long long)vb, (unsigned long long)cb, (unsigned long long)ob, (unsigned long long)values_cap,
                (unsigned long long)columns_cap, (unsigned long long)offsets_cap);
    uint64_t sum = kFnvBasis;
    if (h.n_rows) SMB_TRY(read_block(f, path, offset_rows, (size_t)ob, &sum));
    SMB_TRY(read_block(f, path, columns, (size_t)cb, &sum));
    SMB_TRY(read_block(f, path, values, (size_t)vb, &sum));
    SMB_REQUIRE(sum == h.checksum, SMB200_ERR_IO, "crsfile_read: %s: checksum mismatch (file %016llx, data %016llx)", path,
                (unsigned long long)h.checksum, (unsigned long long)sum);
    return SMB200_OK;
}

}  // namespace
}  // namespace smb

using namespace smb;

extern "C" {

smb200_status smb200_crsfile_write(const char* path, smb200_vtype vt, smb200_itype it, uint64_t n_rows, uint64_t n_cols,
                                   uint64_t nnz, const void* values, const void* columns, const void* offset_rows) {
    SMB_REQUIRE(path, SMB200_ERR_INVALID, "crsfile_write: NULL path");
    SMB_REQUIRE(vt == SMB200_F32 || vt == SMB200_F64, SMB200_ERR_INVALID, "crsfile_write: bad value type %d", (int)vt);
    SMB_REQUIRE(it == SMB200_U32 || it == SMB200_U64, SMB200_ERR_INVALID, "crsfile_write: bad index type %d", (int)it);
    SMB_REQUIRE((values && columns) || nnz == 0, SMB200_ERR_INVALID, "crsfile_write: NULL values/columns");
    SMB_REQUIRE(offset_rows || n_rows == 0, SMB200_ERR_INVALID, "crsfile_write: NULL offset_rows");
    SMB_REQUIRE(n_rows > 0 || nnz == 0, SMB200_ERR_INVALID, "crsfile_write: entries without rows");
    File f(fopen(path, "wb"));
    SMB_REQUIRE(f, SMB200_ERR_IO, "crsfile_write: cannot create %s: %s", path, strerror(errno));
    unsigned char raw[kHeaderBytes] = {0};
    memcpy(raw, kMagic, 8);
    put_u32(raw + 8, (uint32_t)vt);
    put_u32(raw + 12, (uint32_t)it);
    put_u64(raw + 16, n_rows);
    put_u64(raw + 24, n_cols);
    put_u64(raw + 32, nnz);
    SMB_REQUIRE(fwrite(raw, 1, kHeaderBytes, f.get()) == kHeaderBytes, SMB200_ERR_IO, "crsfile_write: short write to %s", path);
    uint64_t sum = kFnvBasis;
    if (n_rows) SMB_TRY(write_block(f.get(), path, offset_rows, (size_t)(n_rows + 1) * isize(it), &sum));
    SMB_TRY(write_block(f.get(), path, columns, (size_t)nnz * isize(it), &sum));
    SMB_TRY(write_block(f.get(), path, values, (size_t)nnz * vsize(vt), &sum));
    // the checksum is known only now: patch it into the header
    unsigned char cs[8];
    put_u64(cs, sum);
    SMB_REQUIRE(fseek(f.get(), 40, SEEK_SET) == 0 && fwrite(cs, 1, 8, f.get()) == 8, SMB200_ERR_IO,
                "crsfile_write: cannot finish %s: %s", path, strerror(errno));
    FILE* raw_f = f.release();
    SMB_REQUIRE(fclose(raw_f) == 0, SMB200_ERR_IO, "crsfile_write: closing %s failed: %s", path, strerror(errno));
    return SMB200_OK;
}

smb200_status smb200_crsfile_info(const char* path, int32_t* vt, int32_t* it, uint64_t* out3) {
    SMB_REQUIRE(path, SMB200_ERR_INVALID, "crsfile_info: NULL path");
    File f(fopen(path, "rb"));
    SMB_REQUIRE(f, SMB200_ERR_IO, "crsfile_info: cannot open %s: %s", path, strerror(errno));
    Header h;
    SMB_TRY(read_header(f.get(), path, &h));
    if (vt) *vt = (int32_t)h.vt;
    if (it) *it = (int32_t)h.it;
    if (out3) { out3[0] = h.n_rows; out3[1] = h.n_cols; out3[2] = h.nnz; }
    return SMB200_OK;
}

smb200_status smb200_crsfile_read(const char* path, void* values, uint64_t values_cap_bytes, void* columns,
                                  uint64_t columns_cap_bytes, void* offset_rows, uint64_t offsets_cap_bytes) {
    SMB_REQUIRE(path, SMB200_ERR_INVALID, "crsfile_read: NULL path");
    File f(fopen(path, "rb"));
    SMB_REQUIRE(f, SMB200_ERR_IO, "crsfile_read: cannot open %s: %s", path, strerror(errno));
    Header h;
    SMB_TRY(read_header(f.get(), path, &h));
    // the file may have changed since the caller sized its buffers with crsfile_info: the capacities decide
    return read_body(f.get(), path, h, values, values_cap_bytes, columns, columns_cap_bytes, offset_rows, offsets_cap_bytes);
}

smb200_status smb200_crs_save(const smb200_crs* m, const char* path) {
    SMB_REQUIRE(m && path, SMB200_ERR_INVALID, "crs_save: NULL argument");
    SMB_REQUIRE(m->x_extra == 0, SMB200_ERR_UNSUPPORTED, "crs_save: the local block of a distributed matrix has remapped columns");
    std::vector<unsigned char> values((size_t)m->nnz * vsize(m->vt)), columns((size_t)m->nnz * isize(m->it)),
        offsets(m->n_rows ? (size_t)(m->n_rows + 1) * isize(m->it) : 0);
    SMB_TRY(smb200_crs_download(m, values.data(), columns.data(), offsets.data()));
    return smb200_crsfile_write(path, (smb200_vtype)m->vt, (smb200_itype)m->it, m->n_rows, m->n_cols, m->nnz, values.data(),
                                columns.data(), m->n_rows ? offsets.data() : nullptr);
}

smb200_status smb200_crs_load(smb200_ctx* ctx, const char* path, smb200_crs** out) {
    SMB_REQUIRE(ctx && path && out, SMB200_ERR_INVALID, "crs_load: NULL argument");
    // one open: the header that sizes the buffers is the header whose body is read into them
    File f(fopen(path, "rb"));
    SMB_REQUIRE(f, SMB200_ERR_IO, "crs_load: cannot open %s: %s", path, strerror(errno));
    Header h;
    SMB_TRY(read_header(f.get(), path, &h));
    std::vector<unsigned char> values, columns, offsets;
    try {
        values.resize((size_t)h.nnz * vsize((int)h.vt));
        columns.resize((size_t)h.nnz * isize((int)h.it));
        offsets.resize(h.n_rows ? (size_t)(h.n_rows + 1) * isize((int)h.it) : 0);
    } catch (const std::exception& e) {
        SMB_FAIL(SMB200_ERR_OOM, "crs_load: %s: cannot hold %llu entries in host memory (%s)", path, (unsigned long long)h.nnz, e.what());
    }
    SMB_TRY(read_body(f.get(), path, h, values.data(), values.size(), columns.data(), columns.size(), offsets.data(), offsets.size()));
    // the upload validates the arrays (monotone offsets ending at nnz, columns < n_cols) like any other matrix
    return smb200_crs_upload(ctx, (smb200_vtype)h.vt, (smb200_itype)h.it, h.n_rows, h.n_cols, h.nnz, values.data(), columns.data(),
                             h.n_rows ? offsets.data() : nullptr, out);
}

}  // extern "C"
