import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import d3p_b200.random as rng
from d3p_b200 import _native as _n, minibatch as mb, parallel
cuda = torch.device("cuda", 0)
for N, world, max_b in ((20000, 2, 1100), (300000, 3, 3200), (20000, 2, 1100)):
    wins = parallel.PeerWindow.local_group(world, 16, max_records=N)
    for w in wins: w.set_timeout_ms(1500)
    need = _n.lib().d3p_poisson_workspace_bytes(N)
    streams = [torch.cuda.Stream() for _ in range(world)]
    for it in range(3):
        key = rng.fold_in(rng.PRNGKey(11), it)
        ref_idx, ref_counts, ref_mask = mb.poisson_sample_idxs(key, 0.05, N, cutoff_size=max_b)
        a = np.ascontiguousarray(np.asarray(key, dtype=np.uint32).reshape(16))
        outs = []
        torch.cuda.synchronize()
        t0 = time.time()
        for r in range(world):
            pb, pe = parallel.position_range(max_b, r, world)
            with torch.cuda.stream(streams[r]):
                ws = torch.empty(need, dtype=torch.uint8, device=cuda)
                idx = torch.full((max_b,), -1, dtype=torch.int32, device=cuda)
                counts = torch.empty(2, dtype=torch.int32, device=cuda)
                mask = torch.empty(max_b, dtype=torch.uint8, device=cuda)
                _n.check(_n.lib().d3p_poisson_sample_sharded(wins[r].ptr, a.ctypes.data_as(C.POINTER(C.c_uint32)), float(np.float32(0.05)), N, max_b, 0, pb, pe,
                    _n.ptr(idx), _n.ptr(counts), _n.ptr(mask), _n.ptr(ws), need, _n.stream_ptr()), "x")
            outs.append((pb, pe, idx, counts, mask, ws))
        torch.cuda.synchronize()
        print(N, world, it, "dt", round(time.time() - t0, 3), [w.timeout_detail() for w in wins], [o[3].tolist() for o in outs], ref_counts.tolist(), flush=True)
    for w in wins: w.close()
