"""FFN GEMM shapes of the bench workload under the BLAS back ends PyTorch can use (GPU box)."""
import torch, time
T, C, F = 145920, 128, 1024
torch.backends.cuda.matmul.allow_tf32 = True
x = torch.randn(T, C, device="cuda"); w1 = torch.randn(F, C, device="cuda"); h = torch.randn(T, F, device="cuda")
w2 = torch.randn(C, F, device="cuda"); gy = torch.randn(T, C, device="cuda"); gh = torch.randn(T, F, device="cuda")
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
for lib in ("cublas", "cublaslt"):
    torch.backends.cuda.preferred_blas_library(lib)
    print(lib, "fwd1 x@w1.T %.3f | fwd2 h@w2.T %.3f | dW1 gh.T@x %.3f | dW2 gy.T@h %.3f | dX1 gh@w1 %.3f | dH gy@w2 %.3f ms" % (
        t(lambda: x @ w1.t()), t(lambda: h @ w2.t()), t(lambda: gh.t() @ x), t(lambda: gy.t() @ h), t(lambda: gh @ w1), t(lambda: gy @ w2)))
