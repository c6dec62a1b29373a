// C surface over the host-side SparseMatIndexList of sparsemat.hpp (see include/smb200_host.h).
#include <cstring>

#include "../../include/smb200_host.h"
#include "sparsemat.hpp"

namespace smb {
void set_error(const char* fmt, ...);
}

struct smb200_il {
    int vt, it;
    virtual ~smb200_il() = default;
    virtual void apply(uint64_t n, const uint64_t* i, const uint64_t* j, const void* v, int op) = 0;
    virtual double get(uint64_t i, uint64_t j) const = 0;
    virtual void dims(uint64_t* out3) const = 0;
    virtual void export_arrays(void* columns, void* values, void* pos_start, void* index_list) const = 0;
    virtual smb200_status to_crs(smb200_ctx* ctx, smb200_crs** out) const = 0;
};

namespace {
template <class T, class I>
struct IlImpl final : smb200_il {
    sparsemat::SparseMatIndexList<T, I> m;
    void apply(uint64_t n, const uint64_t* i, const uint64_t* j, const void* v, int op) override {
        const T* vals = static_cast<const T*>(v);
        if (op) for (uint64_t k = 0; k < n; ++k) m.add_to(i[k], j[k], vals[k]);
        else for (uint64_t k = 0; k < n; ++k) m.set(i[k], j[k], vals[k]);
    }
    double get(uint64_t i, uint64_t j) const override { return (double)m.get(i, j); }
    void dims(uint64_t* o) const override { o[0] = m.n_rows(); o[1] = m.n_cols(); o[2] = m.n_non_zero_entries(); }
    void export_arrays(void* columns, void* values, void* pos_start, void* index_list) const override {
        const uint64_t nz = m.n_non_zero_entries(), nr = m.n_rows();
        if (columns && nz) std::memcpy(columns, m.columns().data(), nz * sizeof(I));
        if (values && nz) std::memcpy(values, m.values().data(), nz * sizeof(T));
        if (pos_start && nr) std::memcpy(pos_start, m.chains().pos_start.data(), nr * sizeof(I));
        if (index_list && nz) std::memcpy(index_list, m.chains().index_list.data(), nz * sizeof(I));
    }
    smb200_status to_crs(smb200_ctx* ctx, smb200_crs** out) const override {
        return smb200_crs_from_indexlist(ctx, (smb200_vtype)vt, (smb200_itype)it, m.n_rows(), m.n_cols(), m.n_non_zero_entries(),
                                         m.columns().data(), m.values().data(), m.chains().pos_start.data(),
                                         m.chains().index_list.data(), out);
    }
};
}  // namespace

#define IL_REQUIRE(cond, msg) do { if (!(cond)) { ::smb::set_error(msg); return SMB200_ERR_INVALID; } } while (0)

extern "C" {

smb200_status smb200_il_create(smb200_vtype vt, smb200_itype it, smb200_il** out) {
    IL_REQUIRE(out, "il_create: out is NULL");
    IL_REQUIRE((vt == SMB200_F32 || vt == SMB200_F64) && (it == SMB200_U32 || it == SMB200_U64), "il_create: bad type");
    smb200_il* p;
    if (vt == SMB200_F64) p = it == SMB200_U64 ? (smb200_il*)new IlImpl<double, uint64_t>() : (smb200_il*)new IlImpl<double, uint32_t>();
    else p = it == SMB200_U64 ? (smb200_il*)new IlImpl<float, uint64_t>() : (smb200_il*)new IlImpl<float, uint32_t>();
    p->vt = vt; p->it = it;
    *out = p;
    return SMB200_OK;
}
smb200_status smb200_il_free(smb200_il* il) { delete il; return SMB200_OK; }
smb200_status smb200_il_apply(smb200_il* il, uint64_t n, const uint64_t* i, const uint64_t* j, const void* v, int32_t op) {
    IL_REQUIRE(il && ((i && j && v) || n == 0), "il_apply: NULL argument");
    try { il->apply(n, i, j, v, op); }
    catch (const std::exception& e) { ::smb::set_error("il_apply: %s", e.what()); return SMB200_ERR_INVALID; }
    return SMB200_OK;
}
smb200_status smb200_il_get(const smb200_il* il, uint64_t i, uint64_t j, double* out) {
    IL_REQUIRE(il && out, "il_get: NULL argument");
    *out = il->get(i, j);
    return SMB200_OK;
}
smb200_status smb200_il_dims(const smb200_il* il, uint64_t* out3) {
    IL_REQUIRE(il && out3, "il_dims: NULL argument");
    il->dims(out3);
    return SMB200_OK;
}
smb200_status smb200_il_export(const smb200_il* il, void* columns, void* values, void* pos_start, void* index_list) {
    IL_REQUIRE(il, "il_export: NULL argument");
    il->export_arrays(columns, values, pos_start, index_list);
    return SMB200_OK;
}
smb200_status smb200_il_to_crs(const smb200_il* il, smb200_ctx* ctx, smb200_crs** out) {
    IL_REQUIRE(il && ctx && out, "il_to_crs: NULL argument");
    return il->to_crs(ctx, out);
}

}  // extern "C"
