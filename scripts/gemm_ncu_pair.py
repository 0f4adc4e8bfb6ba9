import sys; sys.path.insert(0,'/root/repo/scripts'); sys.path.insert(0,'/root/repo')
import gemm_microbench as g
g.run(128*148, 224, 32*64, 1, 1, 1, 224, reps=2)
g.run(128*148, 224, 32*64, 1, 1, 1, 224, reps=2, unsplit=True)
