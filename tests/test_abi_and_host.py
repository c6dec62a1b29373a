"""CPU: the C-ABI library loads and exports every symbol include/*.h declares; the host-side logic that needs no
GPU (IndexList assembly, SparseMatPar contract, row partitioning, ghost plan) agrees with the oracle; compute
entry points fail loudly without a device (there is no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import cases
from conftest import n_gpus

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for h in ("smb200.h", "smb200_host.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(smb200_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol(smb):
    lib = C.CDLL(smb._ffi.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 60
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"
    assert declared == set(smb._ffi.PROTOTYPES), "the ctypes prototypes and the headers list different entry points"
    assert lib.smb200_version() == 100


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: the headers compile as C99 with no C++ or torch types."""
    src = tmp_path / "t.c"
    src.write_text('#include "smb200.h"\n#include "smb200_host.h"\nint main(void){return SMB200_OK;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o",
                    str(tmp_path / "t.o")], check=True)


def test_sm100a_code_only(smb):
    out = subprocess.run(["cuobjdump", "--list-elf", smb._ffi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


@pytest.mark.skipif(n_gpus() > 0, reason="checks the no-device behaviour")
def test_compute_fails_loudly_without_a_device(smb):
    with pytest.raises(smb.SmbError) as e:
        smb.Context(0)
    assert e.value.status == 3 and "CUDA" in str(e.value)                 # SMB200_ERR_CUDA: no fallback path


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sparsemat_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or fn == "Makefile":
                text = open(os.path.join(dirpath, fn), errors="replace").read()
                assert "liboracle" not in text and "oracle_py" not in text and "sparsemat_oracle" not in text, fn


# ---- host assembler (sparsemat_indexlist.rs:29-53,158-164 + indexlist.rs:62-83) against the oracle -------------
@pytest.mark.parametrize("vdt,idt", [(np.float32, np.uint32), (np.float64, np.uint64)])
def test_indexlist_assembly_matches_oracle(smb, orc, vdt, idt):
    rng = np.random.default_rng(5)
    n = 4000
    i = rng.integers(0, 300, n)
    j = rng.integers(0, 200, n)
    v = rng.uniform(-1, 1, n).astype(vdt)
    ops = rng.integers(0, 2, n)
    a, b = smb.SparseMatIndexList(vdt, idt), orc.IndexListMat(vdt, idt)
    for lo in range(0, n, 500):                                           # mixed set / add_to with many repeats
        sl = slice(lo, lo + 500)
        for op in (0, 1):
            m = ops[sl] == op
            (a.set if op == 0 else a.add_to)(i[sl][m], j[sl][m], v[sl][m])
            (b.set if op == 0 else b.add_to)(i[sl][m], j[sl][m], v[sl][m])
    assert a._dims() == b.dims()
    for got, want in zip(a.raw_arrays(), b.raw_arrays()):
        assert got.dtype == want.dtype and np.array_equal(got, want)
    assert a.density() == a.n_non_zero_entries() / (a.n_rows() * a.n_cols())
    # what to_crs will produce on the device is what the oracle produces from the same arrays
    cols, vals, pos, nxt = a.raw_arrays()
    ov, oc, oo = orc.to_crs_raw(a.n_rows(), cols, vals, pos, nxt)
    _, _, wv, wc, wo = b.to_crs()
    assert np.array_equal(ov, wv) and np.array_equal(oc, wc) and np.array_equal(oo, wo)


def test_reference_indexlist_script_on_the_host_mirror(smb):
    sp = smb.SparseMatIndexList(np.float32, np.uint32)                    # lib.rs:57-66
    sp.add_to(0, 1, 4.2); sp.add_to(1, 2, 4.12); sp.add_to(2, 2, 2.12); sp.add_to(1, 1, 1.12)
    sp.add_to(1, 1, 1.12); sp.add_to(0, 2, 0.12); sp.set(0, 0, 8.12); sp.set(0, 0, 7.12)
    assert sp.get(0, 0) == np.float32(7.12) and sp.density() == 6.0 / 9.0
    cols, vals, pos, nxt = sp.raw_arrays()
    assert list(cols) == [1, 2, 2, 1, 2, 0] and list(pos) == [0, 1, 2]
    unset = np.iinfo(np.uint32).max                                       # indexlist.rs:33 UNSET = I::MAX
    assert list(nxt) == [4, 3, unset, unset, 5, unset]


# ---- partitioning (sparsemat_par.rs:20-35 generalised to ranks) -------------------------------------------------
def test_partition_rows(smb):
    assert list(smb.partition_rows(10, 3)) == [0, 4, 8, 10]
    assert list(smb.partition_rows(512 ** 3, 8, 512 ** 2)) == [k * 64 * 512 ** 2 for k in range(9)]
    assert list(smb.partition_rows(5, 8)) == [0, 1, 2, 3, 4, 5, 5, 5, 5]
    b = smb.partition_rows(1000, 7, 16)
    assert b[0] == 0 and b[-1] == 1000 and np.all(np.diff(b.astype(np.int64)) >= 0) and np.all(b[1:-1] % 16 == 0)


def test_partition_rows_by_nnz(smb):
    _, _, _, _, offs = cases.powerlaw(2, 50000, 50000, 20000, np.float64, np.uint64)
    for world in (1, 2, 4, 8):
        b = smb.partition_rows_by_nnz(offs, world).astype(np.int64)
        assert b[0] == 0 and b[-1] == 50000 and np.all(np.diff(b) >= 0)
        nnz = np.diff(offs.astype(np.int64)[b])
        longest = int(np.max(np.diff(offs.astype(np.int64))))
        assert nnz.max() - nnz.min() <= 2 * longest + 1                   # balanced up to one row either side


@pytest.mark.parametrize("idt", [np.uint32, np.uint64])
def test_ghost_plan_simulated_ranks(smb, orc, idt):
    """P simulated ranks: local block + ghost plan + 'exchange' (plain indexing) reproduces the global product."""
    vdt = np.float64
    n_rows, n_cols, vals, cols, offs = cases.ragged(3, 4000, 4000, 15, vdt, idt, empty_frac=0.1)
    x = orc.uniform(vdt, 4, n_cols)
    want = orc.mvp(vals, cols, offs, x)
    o64 = offs.astype(np.int64)
    for world in (1, 2, 4, 8):
        bounds = smb.partition_rows(n_rows, world)
        for rank in range(world):
            lo, hi = int(bounds[rank]), int(bounds[rank + 1])
            lv, lc = vals[o64[lo]:o64[hi]], cols[o64[lo]:o64[hi]]
            lofs = (o64[lo:hi + 1] - o64[lo]).astype(idt)
            local_cols, ghosts, per_owner = smb.ghost_plan(lc, world, rank, bounds)
            assert np.all(np.diff(ghosts.astype(np.int64)) > 0)           # sorted, unique
            assert not np.any((ghosts >= lo) & (ghosts < hi)) and per_owner[rank] == 0 and per_owner.sum() == ghosts.size
            owners = np.searchsorted(bounds.astype(np.int64), ghosts.astype(np.int64), side="right") - 1
            assert np.array_equal(np.bincount(owners, minlength=world).astype(np.uint64), per_owner)
            x_local = np.concatenate([x[lo:hi], x[ghosts.astype(np.int64)]])    # [owned | ghosts]
            assert np.array_equal(orc.mvp(lv, local_cols, lofs, x_local), want[lo:hi])
    with pytest.raises(smb.SmbError):                                      # a column outside the global matrix
        smb.ghost_plan(np.array([5000], idt), 2, 0, smb.partition_rows(n_rows, 2))


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the oracle port on the host cores) honours the bench contract: exactly one JSON line on
    stdout with the metric / config / cpu_baseline / e2e keys.  Runs the full C2 workload on the CPU (a few seconds)."""
    import json
    import sys
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "crs_spmv_effective_hbm_gbs" and d["unit"] == "GB/s" and d["higher_is_better"]
    assert d["value"] > 0 and d["config"]["nnz"] == 117_047_296 and d["dtype"] == "f32"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["serial_value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_binary_crs_container_round_trip_and_damage(tmp_path):
    """smb200_crsfile_* (host only): the three arrays of SparseMatCRS (sparsemat_crs.rs:9-17) survive byte for byte for every
    type combination, ragged and empty shapes included; foreign, truncated and bit-flipped files are refused with ERR_IO."""
    import sparsemat_b200 as smb
    from sparsemat_b200 import _ffi as F
    rng = np.random.default_rng(5)
    for k, (vdt, idt) in enumerate([(np.float32, np.uint32), (np.float64, np.uint32), (np.float32, np.uint64), (np.float64, np.uint64)]):
        for n_rows, n_cols, max_len in [(37, 23, 6), (1, 1, 1), (5, 9, 0), (1000, 1000, 40)]:
            lens = rng.integers(0, max_len + 1, n_rows)
            offs = np.zeros(n_rows + 1, idt)
            offs[1:] = np.cumsum(lens)
            nnz = int(offs[-1])
            cols = rng.integers(0, n_cols, nnz).astype(idt)                   # unsorted inside a row, duplicates allowed
            vals = rng.uniform(-1, 1, nnz).astype(vdt)
            path = tmp_path / f"m{k}_{n_rows}.smbcrs"
            smb.crsfile_write(path, n_rows, n_cols, vals, cols, offs)
            assert os.path.getsize(path) == 56 + sum((a.nbytes + 7) // 8 * 8 for a in (offs, cols, vals))
            r, c, v2, c2, o2 = smb.crsfile_read(path)
            assert (r, c) == (n_rows, n_cols)
            assert v2.dtype == vals.dtype and c2.dtype == cols.dtype and o2.dtype == offs.dtype
            assert v2.tobytes() == vals.tobytes() and np.array_equal(c2, cols) and np.array_equal(o2, offs)
    # the 0 x 0 matrix (SparseMatCRS::new(), what to_crs gives for an empty IndexList): header only
    empty = tmp_path / "empty.smbcrs"
    smb.crsfile_write(empty, 0, 0, np.empty(0, np.float64), np.empty(0, np.uint32), np.empty(0, np.uint32))
    assert os.path.getsize(empty) == 56
    r, c, v2, c2, o2 = smb.crsfile_read(empty)
    assert (r, c, v2.size, c2.size, o2.size) == (0, 0, 0, 0, 0)
    # damage
    good = (tmp_path / "m0_37.smbcrs").read_bytes()

    def refused(data, what):
        bad = tmp_path / "bad.smbcrs"
        bad.write_bytes(data)
        with pytest.raises(smb.SmbError) as e:
            smb.crsfile_read(bad)
        assert e.value.status == F.ERR_IO and what in str(e.value), (what, str(e.value))

    refused(b"NOTACRS!" + good[8:], "not a SMBCRS01 file")
    refused(good[:40], "shorter than a header")
    refused(good[:-9], "truncated or corrupt")
    flipped = bytearray(good)
    flipped[-13] ^= 0x10                                                      # one bit inside the values (the last 4 bytes are padding)
    refused(bytes(flipped), "checksum mismatch")
    # a header that claims 2^31 entries (24 GB): refused on the file's size, before anything is allocated on its word
    huge = bytearray(good)
    huge[32:40] = (1 << 31).to_bytes(8, "little")
    refused(bytes(huge), "truncated or corrupt")
    refused(good + b"\0" * 8, "truncated or corrupt")                        # trailing bytes the header does not describe
    with pytest.raises(smb.SmbError) as e:
        smb.crsfile_read(tmp_path / "does_not_exist.smbcrs")
    assert e.value.status == F.ERR_IO
    # buffers smaller than the file needs are refused (the capacities are part of the call)
    vt, it = C.c_int32(), C.c_int32()
    d = (C.c_uint64 * 3)()
    gp = os.fsencode(tmp_path / "m0_37.smbcrs")
    assert F.lib.smb200_crsfile_info(gp, C.byref(vt), C.byref(it), d) == F.OK
    v = np.empty(int(d[2]), np.float32); c = np.empty(int(d[2]), np.uint32); o = np.empty(int(d[0]) + 1, np.uint32)
    assert F.lib.smb200_crsfile_read(gp, F.ptr(v), v.nbytes, F.ptr(c), c.nbytes, F.ptr(o), o.nbytes) == F.OK
    assert F.lib.smb200_crsfile_read(gp, F.ptr(v), v.nbytes - 4, F.ptr(c), c.nbytes, F.ptr(o), o.nbytes) == F.ERR_INVALID


def test_rust_build_script_compiles_the_same_sources_as_the_makefile():
    """rust/sparsemat-b200-sys/build.rs (the `build.rs invoking nvcc` of the north star) and csrc/Makefile must list the same
    translation units: a stale list links a library with missing symbols."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    mk = open(os.path.join(root, "sparsemat_b200", "csrc", "Makefile")).read()
    srcs = set()
    for var in ("SRCS_CU", "SRCS_CPP", "SRCS_HOST"):
        m = re.search(rf"^{var}\s*:=\s*(.*)$", mk, re.M)
        assert m, var
        srcs |= set(m.group(1).split())
    rs = open(os.path.join(root, "rust", "sparsemat-b200-sys", "build.rs")).read()
    m = re.search(r"let sources = \[(.*?)\];", rs, re.S)
    assert m
    rust_srcs = set(re.findall(r'"([^"]+)"', m.group(1)))
    assert rust_srcs == srcs, (sorted(rust_srcs - srcs), sorted(srcs - rust_srcs))
    for s_ in srcs:
        assert os.path.exists(os.path.join(root, "sparsemat_b200", "csrc", s_)), s_


def test_rust_sys_crate_declares_exactly_the_header_functions():
    """rust/sparsemat-b200-sys/src/lib.rs mirrors include/smb200.h one to one (names and argument counts).  smb200_host.h
    is for language mirrors WITHOUT the reference crate (Python, C): the Rust overlay keeps the crate's own SparseMatIndexList."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def c_decls(path):
        txt = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
        out = {}
        for m in re.finditer(r"\b(smb200_\w+)\s*\(([^;{]*?)\)\s*;", txt, re.S):
            args = m.group(2).strip()
            out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
        return out

    hdr = {}
    hdr.update(c_decls(os.path.join(root, "include", "smb200.h")))
    rs = open(os.path.join(root, "rust", "sparsemat-b200-sys", "src", "lib.rs")).read()
    rust = {}
    for m in re.finditer(r"pub fn (smb200_\w+)\s*\(([^)]*)\)", rs, re.S):
        args = m.group(2).strip()
        rust[m.group(1)] = 0 if not args else args.count(":")
    missing = sorted(set(hdr) - set(rust))
    extra = sorted(set(rust) - set(hdr))
    assert not missing and not extra, {"not bound in the sys crate": missing, "bound but not declared": extra}
    wrong = {k: (hdr[k], rust[k]) for k in hdr if hdr[k] != rust[k]}
    assert not wrong, wrong
