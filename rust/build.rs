// build.rs -- compiles the CUDA library for sm_100a with nvcc and links it (same commands as
// grayshift_b200/csrc/Makefile).  NOT COMPILED HERE (no Rust toolchain in the build environment).
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let out = PathBuf::from(std::env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from(std::env::var("CARGO_MANIFEST_DIR").unwrap()).join("../grayshift_b200/csrc");
    let units = [("abi.cu", false), ("scene_layout.cu", false), ("render_kernels.cu", false), ("wavefront_kernels.cu", false), ("trace_closest.cu", true)];
    let mut objects = Vec::new();
    for (file, no_fma) in units {
        let obj = out.join(file.replace(".cu", ".o"));
        let mut cmd = Command::new("nvcc");
        cmd.args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-c"]);
        if no_fma {
            cmd.arg("-fmad=false"); // parity kernel: f64 arithmetic without FMA contraction, like rustc
        }
        cmd.arg(csrc.join(file)).arg("-o").arg(&obj);
        assert!(cmd.status().expect("nvcc not found").success(), "nvcc failed on {file}");
        println!("cargo:rerun-if-changed={}", csrc.join(file).display());
        objects.push(obj);
    }
    let lib = out.join("libgrayshift_cuda.so");
    let mut link = Command::new("nvcc");
    link.args(["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o"]).arg(&lib).args(&objects);
    assert!(link.status().unwrap().success(), "nvcc link failed");
    for h in ["gs_device.cuh", "gs_shade.cuh", "gs_kernel_util.cuh", "gs_internal.h", "gs_host.h", "trace_closest_body.cuh"] {
        println!("cargo:rerun-if-changed={}", csrc.join(h).display());
    }
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=grayshift_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", out.display());
}
