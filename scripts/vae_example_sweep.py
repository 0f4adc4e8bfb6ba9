import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples")); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vae as ex
for argv in (["--num-epochs", "30", "--epsilon", "8.0", "-batch-size", "256", "-lr", "3e-3"],
             ["--num-epochs", "30", "--epsilon", "8.0", "-batch-size", "1024", "-lr", "1e-2"],
             ["--num-epochs", "60", "--epsilon", "8.0", "-batch-size", "256", "-lr", "1e-3"]):
    try:
        out = ex.main(ex.parse(argv), verbose=False)
        print(argv, "dp_scale", round(out["dp_scale"], 3), [round(h[1], 1) for h in out["history"]][::3], flush=True)
    except Exception as e:
        print(argv, "ERROR", e, flush=True)
