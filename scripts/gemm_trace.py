"""Development aid (needs a D3P_NVCC_DEFINES=D3P_GEMM_TRACE build): per-CTA %globaltimer stamps of the tcgen05 GEMMs of
one VAE step -> where each kernel's time goes (setup / first stage / main loop / drain / epilogue) and the gaps between
kernels.  Not part of the product; numbers are from an instrumented build."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import d3p_b200.random as rng
from d3p_b200 import _native as _n, minibatch as mb, models, optimizers, svi as dsvi

dev = torch.device("cuda", 0)
N, B = 60000, 4096
kx, kw = rng.split(rng.PRNGKey(123), 2)
X = (rng.uniform(kx, (N, 28, 28)) < rng.uniform(kw, (N, 28, 28))).to(torch.float32)
fam = models.VAE(784, 400, 20)
s = dsvi.DPSVI(fam.model, fam.guide, optimizers.Adam(1e-3), models.Trace_ELBO(), 10.0, 1.0, num_obs_total=N)
s.donate_state = True
init, get = mb.subsample_batchify_data((X,), batch_size=B, return_mask=True)
key, k_init, k_fetch = rng.split(rng.PRNGKey(0), 3)
_, bst = init(k_fetch)
batch, mask = get(0, bst)
st = s.init(k_init, *batch)
for i in range(5):
    batch, mask = get(i, bst)
    st, loss = s.update(st, *batch, mask=mask)
torch.cuda.synchronize()
lib = _n.lib()._handle if hasattr(_n.lib(), "_handle") else None
fn = _n.lib().d3p_dev_gemm_trace
fn.restype, fn.argtypes = C.c_uint32, [C.c_void_p, C.c_uint32]
cap = 4096
buf = torch.zeros(cap * 8, dtype=torch.int64, device=dev)
fn(C.c_void_p(buf.data_ptr()), cap)
batch, mask = get(7, bst)
torch.cuda.synchronize()
st, loss = s.update(st, *batch, mask=mask)
torch.cuda.synchronize()
used = fn(None, 0)
rec = buf.cpu().numpy().astype(np.uint64).reshape(cap, 8)[:used]
i = 0
t00 = None
print(f"{'launch (M N K BN S EW)':34s} ctas  start   dur | setup 1st-stage mainloop drain epilogue (median us per CTA) | last-cta-end")
prev_end = None
while i < used:
    hdr = rec[i]; ctas = int(hdr[1]); body = rec[i + 1:i + 1 + ctas].astype(np.int64)
    i += ctas + 1
    if t00 is None: t00 = body[:, 0].min()
    st_, en = body[:, 0].min(), body[:, 5].max()
    ph = [(body[:, k + 1] - body[:, k]) / 1e3 for k in range(5)]
    med = [float(np.median(p)) for p in ph]
    gap = "" if prev_end is None else f" gap {(st_ - prev_end) / 1e3:6.1f}"
    prev_end = en
    print(f"M={int(hdr[2]):5d} N={int(hdr[3]):4d} K={int(hdr[4]):5d} BN={int(hdr[5]):3d} S={int(hdr[6])} EW={int(hdr[7]):2d} {ctas:4d} {(st_ - t00) / 1e3:7.1f} {(en - st_) / 1e3:6.1f} | "
          + " ".join(f"{m:7.2f}" for m in med) + f" | cta-start spread {(body[:, 0].max() - st_) / 1e3:5.1f}{gap}")
