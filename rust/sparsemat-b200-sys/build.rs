// build.rs — compiles libsmb200 with nvcc for sm_100a and links it.  No cuSPARSE, no Triton, no CPU fallback.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let root = manifest.join("../..").canonicalize().unwrap();
    let csrc = root.join("sparsemat_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    // the same translation units as sparsemat_b200/csrc/Makefile (SRCS_CU, SRCS_CPP, SRCS_HOST); tests/test_abi_and_host.py
    // fails when the two lists drift apart
    let sources = ["context.cu", "vector_ops.cu", "crs.cu", "generators.cu", "spmv.cu", "bandsplit.cu", "cg.cu", "pcg.cu", "dist.cu",
                   "transpose.cu", "par.cu", "partition.cpp", "crs_io.cpp", "../host/assembler_capi.cpp"];
    let mut objects = Vec::new();
    for s in sources.iter() {
        let obj = out.join(format!("{}.o", s.replace("../", "").replace("/", "_")));
        let st = Command::new(&nvcc)
            .args(&["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr",
                    "-Xcompiler", "-fPIC", "-c"])
            .arg(csrc.join(s)).arg("-o").arg(&obj)
            .status().expect("nvcc not found (set NVCC)");
        assert!(st.success(), "nvcc failed on {}", s);
        objects.push(obj);
        println!("cargo:rerun-if-changed={}", csrc.join(s).display());
    }
    let st = Command::new(&nvcc).args(&["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o"])
        .arg(out.join("libsmb200.so")).args(&objects).arg("-ldl").status().unwrap();
    assert!(st.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=smb200");
    println!("cargo:rerun-if-changed={}", root.join("include/smb200.h").display());
}
