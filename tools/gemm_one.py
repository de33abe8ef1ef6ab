import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.gemm_bench import bench
which = sys.argv[1] if len(sys.argv) > 1 else 'qkv'
M = 21312
if which == 'qkv': bench(M, 1536, 512, reps=3)
elif which == 'ff2': bench(M, 512, 2048, reps=3)
elif which == 'ff1': bench(M, 2048, 512, act=1, reps=3)
elif which == 'out': bench(M, 512, 512, reps=3)
