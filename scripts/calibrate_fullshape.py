"""Development aid: measured element-wise errors of the C5 (B=4096) and C4 (K=64,d=128) paths vs the oracles."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import d3p_b200.random as rng
from d3p_b200 import models, optimizers, svi as dsvi
from oracle import chacha, svi as osvi, vae as ovae, threefry, gmm as ogmm

def pct(name, e):
    e = np.asarray(e).ravel()
    print(f"  {name}: max {e.max():.3e}  p99.9 {np.quantile(e, .999):.3e}  p99 {np.quantile(e, .99):.3e}  median {np.median(e):.3e}", flush=True)

def vae():
    D, H, Z, B, N, C = 784, 400, 20, 4096, 60000, 10.0
    rs = np.random.RandomState(0)
    X = (rs.rand(B, 28, 28) < 0.35).astype(np.float32)
    mask = np.ones(B, bool); mask[B // 3::5] = False
    for init_std in (0.03, 0.05):
        ofam = ovae.VAE(D, H, Z, N); p0 = ofam.init_params(0, init_std)
        fam = models.VAE(D, H, Z)
        s = dsvi.DPSVI(fam.model, fam.guide, optimizers.SGD(1.0), models.Trace_ELBO(), C, 0.0, num_obs_total=N)
        key = chacha.PRNGKey(11)
        st = s.init(key, torch.as_tensor(X).cuda(), params=p0)
        st1, keys = s._split_rng_key(st, 2)
        norms = torch.zeros(B, device="cuda"); loss = torch.zeros(B, device="cuda")
        ws, n_part, _, P = s._run_step(st1, keys[0], (torch.as_tensor(X).cuda(),), torch.as_tensor(mask).cuda(), px_norms=norms, px_loss=loss)
        part = ws.view(-1)[: n_part * (P + 2)].view(n_part, P + 2).double().sum(0).cpu().numpy()
        jk = chacha.convert_to_jax_rng_key(np.asarray(keys[0]))
        eps = ofam.sample_eps(threefry.split(jk, B))["z"]
        t = time.time()
        l, nrm, cs = ovae.explicit_clipped_sum(p0, X, eps, C, mask, ofam.site_scale, st.observation_scale)
        print("VAE init_std", init_std, "explicit oracle", time.time() - t, "s; clipped fraction", float(np.mean(nrm > C)), "n_part", n_part)
        gn = norms.cpu().numpy().astype(np.float64)
        pct("norms rel", np.abs(gn - nrm)[mask] / nrm[mask])
        gl = loss.cpu().numpy().astype(np.float64)
        ref_l = l * st.observation_scale
        pct("loss rel", np.abs(gl - ref_l)[mask] / np.abs(ref_l[mask]))
        off = 0
        for name, o, shape in fam.layout():
            size = int(np.prod(shape)) if len(shape) else 1
            got = part[o:o + size]; ref = cs[name].ravel()
            rms = np.sqrt(np.mean(ref ** 2)); mx = np.abs(ref).max()
            print(f" leaf {name} size {size} rms {rms:.3e} max {mx:.3e}")
            pct("abs/rms", np.abs(got - ref) / rms)
            pct("abs/max", np.abs(got - ref) / mx)
            pct("elementwise rel floor=rms", np.abs(got - ref) / np.maximum(np.abs(ref), rms))
            pct("elementwise rel floor=1e-2 rms", np.abs(got - ref) / np.maximum(np.abs(ref), 1e-2 * rms))
        print(" loss col", part[P], ref_l.sum(), " count col", part[P + 1], mask.sum())

def gmm():
    K, d, N, B, C = 64, 128, 2000, 96, 20.0
    rs = np.random.RandomState(0)
    centers = rs.randn(K, d).astype(np.float32) * 3
    X = (centers[rs.randint(0, K, B)] + rs.randn(B, d)).astype(np.float32)
    p0 = {"alpha_log": (rs.randn(K) * 0.4).astype(np.float32), "mus_loc": (centers + rs.randn(K, d) * 0.5).astype(np.float32)}
    o = osvi.DPSVI(ogmm.GaussianMixture(K, d, N), None, osvi.Adam(1e-3), None, C, 1.0)
    fam = models.GaussianMixture(K, d)
    s = dsvi.DPSVI(fam.model, fam.guide, optimizers.Adam(1e-3), models.Trace_ELBO(), C, 1.0, num_obs_total=N)
    key = chacha.PRNGKey(21)
    ost = o.init(key, X, params=p0); st = s.init(key, torch.as_tensor(X).cuda(), params=p0)
    ost1, okeys = o._split_rng_key(ost, 2)
    _, opx_loss, opx_grads, n, f = o._compute_per_example_gradients(ost1, okeys[0], X)
    st1, keys = s._split_rng_key(st, 2)
    _, px_loss, px_grads, n2, f2 = s._compute_per_example_gradients(st1, keys[0], torch.as_tensor(X).cuda())
    print("GMM K=64 d=128 B=96")
    pct("loss rel", np.abs(px_loss.cpu().numpy() - opx_loss) / np.abs(opx_loss))
    for k in opx_grads:
        got, ref = px_grads[k].cpu().numpy().astype(np.float64), opx_grads[k].astype(np.float64)
        rms = np.sqrt(np.mean(ref.reshape(B, -1) ** 2, axis=1)).reshape((B,) + (1,) * (ref.ndim - 1))
        mx = np.abs(ref).reshape(B, -1).max(axis=1).reshape((B,) + (1,) * (ref.ndim - 1))
        pct(k + " abs/rowmax", np.abs(got - ref) / mx)
        pct(k + " abs/rowrms", np.abs(got - ref) / rms)
        pct(k + " elementwise floor=rowrms", np.abs(got - ref) / np.maximum(np.abs(ref), rms))
        pct(k + " elementwise floor=1e-2 rowrms", np.abs(got - ref) / np.maximum(np.abs(ref), 1e-2 * rms))

if __name__ == "__main__":
    vae(); gmm()
