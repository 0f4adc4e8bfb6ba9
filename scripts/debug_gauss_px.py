import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import test_gpu_svi as T
from oracle import chacha, threefry
cuda = torch.device("cuda", 0)
kind, d, guide = "gauss", 1024, "hand"
N, B = 10000, 19
s, o, fam = T._pair(kind, d, N, guide)
args = T._data(kind, B, d)
p = T._rand_params(fam)
key = chacha.PRNGKey(5)
st, ost = s.init(key, *[torch.as_tensor(a).to(cuda) for a in args], params=p), o.init(key, *args, params=p)
mask = np.arange(B) < B - 5
_, losses, grads, n, f = s._compute_per_example_gradients(st, st.rng_key, *[torch.as_tensor(a).to(cuda) for a in args], mask=torch.as_tensor(mask).to(cuda))
o.grad_dtype = np.float64
_, olosses, ograds, on, of = o._compute_per_example_gradients(ost, ost.rng_key, *args, mask=mask)
jk = chacha.convert_to_jax_rng_key(ost.rng_key)
eps = o.model.sample_eps(threefry.split(jk, B))["mu"]
for k in grads:
    got = grads[k].cpu().numpy().astype(np.float64); ref = ograds[k]
    rms = np.sqrt(np.mean(ref ** 2))
    e = np.abs(got - ref) / np.maximum(np.abs(ref), rms)
    idx = np.argsort(e.ravel())[::-1][:6]
    print(k, "rms", rms)
    for i in idx:
        b, j = divmod(i, d)
        mu, rho, x, ep = p["mu_loc"][j], p["mu_std_log"][j], args[0][b, j], eps[b, j]
        sc = np.exp(np.float64(rho)); z = mu + ep * sc
        print(f"  b={b} j={j} err={e.ravel()[i]:.3e} got={got[b,j]:.9e} ref={ref[b,j]:.9e} eps={ep:.7f} s={sc:.6f} mu={mu:.5f} x={x:.5f} z-x={z-x:.6f}")
