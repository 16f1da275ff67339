import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from quick_perf import run
out, tr = run(int(sys.argv[1]), 0, int(sys.argv[2]), reps=2)
print(out)
