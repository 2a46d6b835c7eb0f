"""fp32 GEMMs of the dense GNN layers on cuBLAS 12.9's BF16x9 emulation (fp32-accurate, tensor cores).

The dense GraphSAGE / GCN / GAT layers are torch library GEMMs -- not part of the TT path, but half of a
GraphSAGE step's device time as fp32 SIMT SGEMM.  torch 2.11+cu128 ships cuBLAS 12.8; the image's CUDA toolkit
has cuBLAS 12.9, which can emulate fp32 GEMMs with nine bf16 tensor-core products (CUBLAS_EMULATE_SINGLE_PRECISION,
strategy "performant": only where it is faster).  Loaded BEFORE torch, the toolkit's libcublasLt.so.12 /
libcublas.so.12 satisfy torch's dependencies on those sonames, so every torch GEMM of the process runs on them.
Measured on B200 (profiles/r2d_cublas_*.txt, r2d_sage_emulated_performant.txt): maximum error against fp64
1.6e-7 - 1.9e-7 of the largest element (native fp32 SGEMM: 5e-7 - 8e-7), forward GEMMs 1.5 - 1.7x faster, the
weight-gradient GEMMs unchanged, GraphSAGE epoch 0.80 -> 0.73 s, the loss after three epochs equal to five digits.

    import _cublas_emulation      # first, before `import torch`;  TTG_CUBLAS_EMULATION=0 turns it off

ACTIVE tells the caller what happened (the bench records say which GEMMs they timed)."""
import ctypes
import os
import sys

ACTIVE = False
WHY = "TTG_CUBLAS_EMULATION=0"

if os.environ.get("TTG_CUBLAS_EMULATION", "1") != "0":
    _dir = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "lib64")
    _libs = [os.path.join(_dir, n) for n in ("libcublasLt.so.12", "libcublas.so.12")]
    if "torch" in sys.modules:
        WHY = "torch was imported first (its bundled cuBLAS is already loaded)"
    elif not all(os.path.exists(p) for p in _libs):
        WHY = "no cuBLAS under %s" % _dir
    else:
        try:
            _real = os.path.realpath(_libs[1])                      # libcublas.so.12.9.1.4
            _ver = tuple(int(x) for x in _real.rsplit(".so.", 1)[1].split(".")[:2])
            if _ver < (12, 9):
                WHY = "cuBLAS %d.%d has no fp32 emulation" % _ver
            else:
                for _p in _libs:
                    ctypes.CDLL(_p, mode=ctypes.RTLD_GLOBAL)
                os.environ.setdefault("CUBLAS_EMULATE_SINGLE_PRECISION", "1")
                os.environ.setdefault("CUBLAS_EMULATION_STRATEGY", "performant")
                ACTIVE = True
                WHY = "cuBLAS %s, BF16x9 emulation of fp32 (strategy %s)" % (
                    _real.rsplit(".so.", 1)[1], os.environ["CUBLAS_EMULATION_STRATEGY"])
        except Exception as ex:   # noqa: BLE001 -- an optimisation of library GEMMs, never a reason to stop
            WHY = "preload failed: %s" % ex
