import sys, time, os, faulthandler
faulthandler.dump_traceback_later(50, repeat=False, exit=True)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import test_gpu_comm_loopback as T
dev = torch.device("cuda", 0)
for world in (2, 4):
    for epoch in (False, True):
        for name, make_fam, data, clip in T._families(dev):
            print('start', world, epoch, name, flush=True)
            try:
                t0 = time.time()
                (p1,), (l1,), (k1,) = T._run(make_fam, data, clip, 1, False)
                flats, losses, keys = T._run(make_fam, data, clip, world, epoch)
                print(world, epoch, name, "ok", T._rel_err(flats[0], p1, 1e-3), losses[0], l1, time.time() - t0, flush=True)
            except Exception as e:
                print(world, epoch, name, "FAILED", repr(e)[:300], time.time() - t0, flush=True)
                torch.cuda.synchronize()
