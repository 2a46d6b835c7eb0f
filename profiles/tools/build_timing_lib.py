"""Build falcon-ttdforgnns_b200/lib/libttg_timing.so: the library with -DTTG_R_TIMING (in-kernel time marks of the
right-grouped row kernels).  Never loaded by the package; profiles/tools/r_timing.py and rm_timing.py point at it."""
import os
import subprocess
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = os.path.join(ROOT, "falcon-ttdforgnns_b200")
sys.path.insert(0, PKG)
import build as b
objdir = os.path.join(PKG, "build", "timing")
os.makedirs(objdir, exist_ok=True)
objs = []
procs = []
for src in b.SOURCES:
    obj = os.path.join(objdir, src.replace(".cu", ".o"))
    objs.append(obj)
    procs.append(subprocess.Popen([b._nvcc()] + [f for f in b.NVCC_FLAGS if f not in ("-Xptxas", "-v")] +
                                  ["-DTTG_R_TIMING", "-c", os.path.join(b.CSRC, src), "-o", obj]))
assert all(p.wait() == 0 for p in procs)
out = os.path.join(PKG, "lib", "libttg_timing.so")
subprocess.check_call([b._nvcc(), "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                                  "-cudart", "static"])
print(out)
