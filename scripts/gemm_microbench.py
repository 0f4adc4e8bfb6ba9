"""Times d3p_gemm_tf32x3 (pre-split operands) and d3p_gemm_f32x3 (split in shared memory by the kernel) on a few
shapes (CUDA events, 20 reps) — development aid, not a bench line."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from d3p_b200 import _native as _n

def split(x, lo=True):
    hi = torch.empty_like(x); l = torch.empty_like(x) if lo else None
    _n.check(_n.lib().d3p_split_tf32(_n.ptr(x), None, x.shape[-1], _n.ptr(hi), _n.ptr(l), x.numel(), _n.stream_ptr()))
    return hi, l

def run(M, N, K, a_mn, b_mn, split_k, tile_n, a_lo=True, b_lo=True, reps=20, unsplit=False):
    dev = torch.device("cuda")
    A = torch.randn((K, M) if a_mn else (M, K), device=dev)
    B = torch.randn((K, N) if b_mn else (N, K), device=dev)
    a_hi, a_l = split(A, a_lo); b_hi, b_l = split(B, b_lo)
    out = torch.empty((split_k, M, N), device=dev)
    def call():
        if unsplit:
            _n.check(_n.lib().d3p_gemm_f32x3(_n.ptr(A), a_mn, A.stride(0), _n.ptr(B), b_mn, B.stride(0), M, N, K, split_k, tile_n,
                                             _n.ptr(out), N, M * N, 0, _n.stream_ptr()))
            return
        _n.check(_n.lib().d3p_gemm_tf32x3(_n.ptr(a_hi), _n.ptr(a_l), a_mn, A.stride(0), _n.ptr(b_hi), _n.ptr(b_l), b_mn, B.stride(0),
                                          M, N, K, split_k, tile_n, _n.ptr(out), N, M * N, 0, _n.stream_ptr()))
    for _ in range(3): call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): call()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nm = 1 + (1 if a_lo else 0) + (1 if b_lo else 0)
    mt = (M + 127) // 128; nt = (N + tile_n - 1) // tile_n
    kb = (K + 31) // 32
    print(f"{'in-kernel split' if unsplit else 'pre-split      '} M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} split={split_k} bn={tile_n} lo=({int(a_lo)},{int(b_lo)}): {ms*1e3:8.1f} us  "
          f"ctas={mt*nt*split_k} kblocks/cta={kb/split_k:.1f} us/kblock={ms*1e3/(kb/split_k):.2f} exec TF/s={2*M*N*K*nm/ms/1e9:.0f}", flush=True)

if __name__ == "__main__":
    if "--compare" in sys.argv:      # the VAE shapes with both operand forms
        for args in [(128, 224, 32 * 64, 0, 0, 1, 224), (128 * 148, 224, 32 * 64, 0, 0, 1, 224),
                     (128 * 148, 224, 32 * 64, 1, 1, 1, 224), (128 * 148, 128, 32 * 64, 0, 0, 1, 128),
                     (784, 400, 4096, 1, 1, 10, 224), (4096, 784, 400, 0, 1, 1, 224), (4096, 400, 784, 0, 0, 1, 128),
                     (4096, 400, 784, 0, 1, 1, 128), (400, 40, 4096, 1, 1, 10, 128)]:
            run(*args)
            run(*args, unsplit=True)
            if args[0] >= 128 * 148 or args[0] == 128:
                run(*args, a_lo=False, b_lo=False)
        sys.exit(0)
    run(128, 224, 32 * 64, 0, 0, 1, 224)            # one CTA, 64 k-blocks: per-k-block latency
    run(128, 224, 32 * 64, 1, 1, 1, 224)
    run(128, 224, 32 * 64, 0, 0, 1, 224, a_lo=False, b_lo=False)
    run(128, 128, 32 * 64, 0, 0, 1, 128)
    run(128 * 148, 224, 32 * 64, 0, 0, 1, 224)      # all SMs busy
    run(128 * 148, 224, 32 * 64, 1, 1, 1, 224)
    run(128 * 148, 224, 32 * 64, 0, 0, 1, 224, a_lo=False, b_lo=False)
    run(128 * 148, 128, 32 * 64, 0, 0, 1, 128)
    run(785, 400, 4096, 1, 1, 10, 224, a_lo=False)  # GW1
    run(784, 401, 4096, 1, 1, 10, 224)              # GW5
    run(4096, 400, 784, 0, 1, 1, 224, a_lo=False)   # G1
    run(4096, 784, 400, 0, 1, 1, 224)               # G5
    run(4096, 400, 784, 0, 0, 1, 224)               # G5b
    run(401, 40, 4096, 1, 1, 10, 128)               # thin
