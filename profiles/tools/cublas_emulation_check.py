"""fp32 GEMMs of torch under cuBLAS 12.9's BF16x9 emulation (CUBLAS_EMULATE_SINGLE_PRECISION=1 with the system
cuBLAS preloaded): error against fp64 and speed against the native fp32 SGEMM, at the GraphSAGE layer shapes."""
import os
import torch

print("cublas emulate env:", os.environ.get("CUBLAS_EMULATE_SINGLE_PRECISION"), os.environ.get("CUBLAS_EMULATION_STRATEGY"),
      "LD_PRELOAD:", os.environ.get("LD_PRELOAD"))
print("torch", torch.__version__, "cublas version seen by torch:", torch._C._cuda_getCompiledVersion() if hasattr(torch._C, "_cuda_getCompiledVersion") else "?")
dev = "cuda:0"
torch.manual_seed(0)
for (m, k, n, tag) in ((150000, 100, 256, "layer0 fwd"), (150000, 256, 256, "layer1-ish fwd"), (256, 150000, 100, "dW (reduction over rows)"),
                       (150000, 256, 100, "dX")):
    a = torch.randn(m, k, device=dev)
    b = torch.randn(k, n, device=dev)
    ref = (a.double() @ b.double())
    out = a @ b
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        a @ b
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("%-28s m=%d k=%d n=%d  max err / max |ref| = %.2e   %.3f ms  %.1f TFLOP/s" % (tag, m, k, n, err, ms, 2.0 * m * k * n / ms / 1e9))
lin = torch.nn.Linear(256, 256).to(dev)
x = torch.randn(150000, 256, device=dev)
y = lin(x)
ref = x.double() @ lin.weight.double().t() + lin.bias.double()
print("Linear(256,256) with bias: err %.2e" % float((y.double() - ref).abs().max() / ref.abs().max()))
