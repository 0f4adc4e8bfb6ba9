"""Development aid: where does the 3xTF32 GEMM lose accuracy?  Separates the fp32 accumulation inside tcgen05.mma
(operands exact in TF32 => every product is exact) from the operand split (truncated lo)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from d3p_b200 import _native as _n
dev = torch.device('cuda')
g = torch.Generator(device='cuda').manual_seed(1)
def trunc(x): return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)
def rn_lo(lo):   # pre-round lo to nearest TF32 so that the hardware truncation is a no-op
    i = lo.view(torch.int32); return ((i + 0x1000) & ~0x1FFF).view(torch.float32)
def run(a_hi, a_lo, b_hi, b_lo, M, N, K):
    out = torch.empty((1, M, N), device=dev)
    _n.check(_n.lib().d3p_gemm_tf32x3(_n.ptr(a_hi), _n.ptr(a_lo), 0, K, _n.ptr(b_hi), _n.ptr(b_lo), 0, K, M, N, K, 1, 224, _n.ptr(out), N, M * N, 0, _n.stream_ptr()))
    torch.cuda.synchronize(); return out[0].double()
def rep(name, got, want):
    rms = want.pow(2).mean().sqrt().item()
    d = (got - want)
    print(f"{name:42s} rms-rel {d.pow(2).mean().sqrt().item() / rms:.3e}  max/rms {d.abs().max().item() / rms:.3e}  mean signed/rms {d.mean().item() / rms:+.3e}")
for K in (128, 1024, 4096):
    for positive in (False, True):
        M, N = 256, 224
        A = torch.randn((M, K), device=dev, generator=g); B = torch.randn((N, K), device=dev, generator=g)
        if positive: A, B = A.abs(), B.abs()
        print(f"--- K={K} positive={positive}")
        At, Bt = trunc(A), trunc(B)
        want_t = At.double() @ Bt.double().T
        rep("exact-TF32 operands, hi only (accum only)", run(At, None, Bt, None, M, N, K), want_t)
        rep("  torch fp32 matmul (no tf32) same operands", (At @ Bt.T).double(), want_t)
        want = A.double() @ B.double().T
        rep("3xTF32 lo = x - trunc (hardware truncs lo)", run(At, A - At, Bt, B - Bt, M, N, K), want)
        rep("3xTF32 lo pre-rounded to nearest", run(At, rn_lo(A - At), Bt, rn_lo(B - Bt), M, N, K), want)
        rep("  torch fp32 matmul", (A @ B.T).double(), want)
