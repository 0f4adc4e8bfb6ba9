"""Development aid: fixed host cost of one DPSVI.run_epoch call (the part a short timed region pays once)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import d3p_b200.random as rng
from d3p_b200 import minibatch as mb, models, optimizers, svi as dsvi
dev = torch.device("cuda", 0)
VAE = "--vae" in sys.argv
if VAE:       # the C5 workload of bench.py
    N = 60_000
    X = (rng.uniform(rng.PRNGKey(1), (N, 28, 28)) < 0.3).float()
    fam = models.VAE(784, 400, 20)
    s = dsvi.DPSVI(fam.model, fam.guide, optimizers.Adam(1e-3), models.Trace_ELBO(), 10.0, 1.0, num_obs_total=N)
    s.donate_state = True
    init, get = mb.subsample_batchify_data((X,), batch_size=4096, return_mask=True)
else:
    N, d = (10_000_000, 1024) if "--c2" in sys.argv else (200_000, 1024)
    X = rng.normal(rng.PRNGKey(1), (N, d)); y = (rng.uniform(rng.PRNGKey(2), (N,)) < 0.5).to(torch.int32)
    fam = models.LogisticRegression(d)
    s = dsvi.DPSVI(fam.model, fam.guide, optimizers.Adam(1e-3), models.Trace_ELBO(), 1.0, 1.0, num_obs_total=N)
    s.donate_state = True
    init, get = mb.poisson_batchify_data((X, y), 0.01, .99)
key, k_init, k_fetch = rng.split(rng.PRNGKey(0), 3)
_, bst = init(k_fetch)
batch, mask = get(0, bst)
st = s.init(k_init, *batch)
st, _ = s.run_epoch(st, get, bst, 3)
torch.cuda.synchronize()
for K in (1, 2, 3, 5, 10, 20, 50, 100):
    ts = []
    for rep in range(5):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        st, _ = s.run_epoch(st, get, bst, K, first_step=10)
        t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
        ts.append((t1 - t0, e0.elapsed_time(e1)))
    h, g = min(t[0] for t in ts), min(t[1] for t in ts)
    print(f"K={K:4d}: host call {h*1e3:.3f} ms, device span {g:.3f} ms ({g/K:.4f} per step)", flush=True)
pr = cProfile.Profile(); pr.enable()
for _ in range(50):
    st, _ = s.run_epoch(st, get, bst, 1, first_step=10)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
