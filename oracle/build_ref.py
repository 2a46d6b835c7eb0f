"""Build the UNMODIFIED reference CUDA extension for sm_100a -- TEST / BASELINE INFRASTRUCTURE.

    python oracle/build_ref.py [--ref /root/reference]

Compiles FBTT/tt_embeddings.cpp + FBTT/tt_embeddings_cuda.cu and
Efficient_TT/efficient_kernel_wrap.cpp + efficient_tt_cuda.cu where they lie under the reference
tree (nothing is copied into the repo) with nvcc + g++ against this image's torch headers, the
only change being the flags its setup.py hard-codes (FBTT/setup.py:24-31: compute_86 and a
private CUB include path).  Output: oracle/_ref/tt_embeddings*.so (git-ignored; it travels to
the GPU box with the snapshot).  The source names its module tt_embeddings
(FBTT/tt_embeddings.cpp:131), the same name as the product's Python shim, so it is never put on
sys.path: load it with oracle.ref_ext.load().  Used by tests/test_gpu_vs_reference_ext.py (parity
of the CUDA path against the reference's own kernels) and by bench.py's `reference_gpu` entry.
"""
import argparse
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def _compile(ref, sub, cpp, cu, modname, stem):
    import torch
    from torch.utils import cpp_extension as ce
    src_cpp = os.path.join(ref, sub, cpp)
    src_cu = os.path.join(ref, sub, cu)
    for f in (src_cpp, src_cu):
        if not os.path.exists(f):
            raise RuntimeError("reference source missing: %s" % f)
    os.makedirs(OUT, exist_ok=True)
    ext = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    so = os.path.join(OUT, stem + ext)
    stamp = so + ".stamp"
    sig = "%s|%s|%s" % (torch.__version__, os.path.getmtime(src_cpp), os.path.getmtime(src_cu))
    if os.path.exists(so) and os.path.exists(stamp) and open(stamp).read() == sig:
        return so
    inc = ["-I" + p for p in ce.include_paths("cuda")] + ["-I" + sysconfig.get_paths()["include"],
                                                          "-I" + os.path.join(ref, sub)]
    abi = "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)
    common = ["-DTORCH_EXTENSION_NAME=" + modname, "-DTORCH_API_INCLUDE_EXTENSION_H", abi]
    obj_cpp, obj_cu = os.path.join(OUT, stem + "_cpp.o"), os.path.join(OUT, stem + "_cu.o")
    subprocess.run(["g++", "-O3", "-fPIC", "-std=c++17", "-w", "-c", src_cpp, "-o", obj_cpp] + inc + common,
                   check=True)
    subprocess.run(["nvcc", "-O3", "--expt-relaxed-constexpr", "-D__CUDA_NO_HALF_OPERATORS__",
                    "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-Xcompiler", "-fPIC",
                    "-w", "-c", src_cu, "-o", obj_cu] + inc + common, check=True)
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    subprocess.run(["g++", "-shared", obj_cpp, obj_cu, "-o", so, "-L" + libdir,
                    "-Wl,-rpath," + libdir, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda",
                    "-ltorch", "-ltorch_python", "-L/usr/local/cuda/lib64", "-lcudart", "-lcublas"],
                   check=True)
    for o in (obj_cpp, obj_cu):
        os.remove(o)
    with open(stamp, "w") as f:
        f.write(sig)
    return so


def build(ref):
    """FBTT extension (module name fixed by the source: tt_embeddings) and the Efficient_TT one
    (module name = TORCH_EXTENSION_NAME, built as efficient_tt_ref)."""
    a = _compile(ref, "FBTT", "tt_embeddings.cpp", "tt_embeddings_cuda.cu", "tt_embeddings",
                 "tt_embeddings")
    b = _compile(ref, "Efficient_TT", "efficient_kernel_wrap.cpp", "efficient_tt_cuda.cu",
                 "efficient_tt_ref", "efficient_tt_ref")
    return a, b


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=os.environ.get("TTG_REFERENCE", "/root/reference"))
    a = ap.parse_args()
    for so in build(a.ref):
        print(so)
