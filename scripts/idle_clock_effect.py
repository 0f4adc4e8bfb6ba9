"""Development aid: does an idle gap before a short timed region cost anything (clock / power-state ramp)?
K = 20 steps of the C2 workload timed right after other work, and after idle gaps of 1, 10, 50, 200 ms."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import d3p_b200.random as rng
from d3p_b200 import minibatch as mb, models, optimizers, svi as dsvi
N, d = 10_000_000, 1024
X = rng.normal(rng.PRNGKey(1), (N, d)); y = (rng.uniform(rng.PRNGKey(2), (N,)) < 0.5).to(torch.int32)
fam = models.LogisticRegression(d)
s = dsvi.DPSVI(fam.model, fam.guide, optimizers.Adam(1e-3), models.Trace_ELBO(), 1.0, 1.0, num_obs_total=N)
s.donate_state = True
init, get = mb.poisson_batchify_data((X, y), 0.01, .99)
key, k_init, k_fetch = rng.split(rng.PRNGKey(0), 3)
_, bst = init(k_fetch)
batch, mask = get(0, bst)
st = s.init(k_init, *batch)
st, _ = s.run_epoch(st, get, bst, 5)
torch.cuda.synchronize()
for gap_ms in (0, 1, 10, 50, 200, 0, 50):
    ts = []
    for rep in range(5):
        st, _ = s.run_epoch(st, get, bst, 5, first_step=3)       # warm-up, as bench.py does
        torch.cuda.synchronize()
        if gap_ms:
            time.sleep(gap_ms / 1000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st, _ = s.run_epoch(st, get, bst, 20, first_step=10)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"idle gap {gap_ms:4d} ms: 20 steps take min {min(ts):.3f} ms, median {sorted(ts)[2]:.3f} ms ({min(ts)/20:.4f} per step)", flush=True)
