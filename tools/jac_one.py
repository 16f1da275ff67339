import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from quick_perf import run
n = int(sys.argv[1]); sw = int(sys.argv[2])
out, tr = run(n, 1, sw, reps=2)
print(n, out["pressure"])
