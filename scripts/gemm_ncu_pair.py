import sys; import os; sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gemm_microbench as g
g.run(128*148, 224, 32*64, 1, 1, 1, 224, reps=2)
g.run(128*148, 224, 32*64, 1, 1, 1, 224, reps=2, unsplit=True)
