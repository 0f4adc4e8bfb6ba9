"""Development aid: replay one trial of tests/test_gpu_fuzz.py::test_fuzz_meanfield_shapes and print where it deviates."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import chacha
from test_gpu_svi import _data, _pair, _rand_params
from d3p_b200.minibatch import BatchView
cuda = torch.device("cuda", 0)
want = int(sys.argv[1]) if len(sys.argv) > 1 else 12
rs = np.random.RandomState(808)
ds = [2, 13, 100, 255, 256, 257, 511, 512, 513, 1023, 1024, 1025, 2000]
for trial in range(16):
    kind = ["logreg", "gauss"][rs.randint(2)]
    guide = ["hand", "auto"][rs.randint(2)]
    d = int(ds[rs.randint(len(ds))])
    B = int(rs.choice([1, 2, 31, 32, 33, 100]))
    N = 20000
    C = 1.0 if kind == "logreg" else 40.0
    n_valid = int(rs.randint(1, B + 1))
    use_view = rs.rand() < .5
    perm = rs.permutation(B).astype(np.int32) if use_view else None
    if trial != want:
        continue
    print(trial, kind, guide, d, B, n_valid, use_view)
    s, o, fam = _pair(kind, d, N, guide, optim="sgd", C=C, dp_scale=0.0)
    o.grad_dtype = np.float64
    args = _data(kind, B, d, seed=trial)
    p = _rand_params(fam, seed=trial, scale=.2)
    key = chacha.PRNGKey(trial)
    mask = np.arange(B) < n_valid
    ost = o.init(key, *args, params=p)
    targs = [torch.as_tensor(a).to(cuda) for a in args]
    tmask = torch.as_tensor(mask).to(cuda)
    st = s.init(key, *targs, params=p)
    # per-example quantities
    st1, keys = s._split_rng_key(st, 2)
    ost1, okeys = o._split_rng_key(ost, 2)
    _, losses, grads, n, f = s._compute_per_example_gradients(st1, keys[0], *targs, mask=tmask)
    _, olosses, ograds, on, of = o._compute_per_example_gradients(ost1, okeys[0], *args, mask=mask)
    for k in grads:
        g, og = grads[k].cpu().numpy().reshape(B, -1), np.asarray(ograds[k]).reshape(B, -1)
        err = np.abs(g - og).max(axis=1) / np.maximum(np.abs(og).max(axis=1), 1e-30)
        print(k, "worst example", int(err.argmax()), float(err.max()))
    gn = np.sqrt(sum((grads[k].cpu().numpy().reshape(B, -1).astype(np.float64) ** 2).sum(1) for k in grads))
    on_ = np.sqrt(sum((np.asarray(ograds[k]).reshape(B, -1).astype(np.float64) ** 2).sum(1) for k in ograds))
    print("norms rel err", np.abs(gn - on_)[mask].max() / on_[mask].max(), "norm range", on_[mask].min(), on_[mask].max())
    ost2, oloss = o.update(ost, *args, mask=mask)
    for variant in ("plain", "view"):
        if variant == "view":
            pm = perm if perm is not None else np.random.RandomState(1).permutation(B).astype(np.int32)
            inv = np.argsort(pm).astype(np.int32)
            srcs = [torch.as_tensor(a[pm]).to(cuda) for a in args]
            idx = torch.as_tensor(inv).to(cuda)
            nv = torch.tensor([n_valid], dtype=torch.int32, device=cuda)
            ta = [BatchView(src, idx, nv) for src in srcs]
        else:
            ta = targs
        st2, loss = s.update(st, *ta, mask=tmask)
        got, ref = s.get_params(st2), o.get_params(ost2)
        print(variant, "loss", float(loss), float(oloss))
        for k in ref:
            gk, rk = got[k].cpu().numpy().astype(np.float64), np.asarray(ref[k], np.float64)
            dp = np.asarray(p[k], np.float64)
            print("   ", k, "param err", float(np.abs(gk - rk).max()), "step size", float(np.abs(rk - dp).max()),
                  "rel to step", float(np.abs(gk - rk).max() / max(np.abs(rk - dp).max(), 1e-30)))
