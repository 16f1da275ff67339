import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from quick_perf import run
n = int(sys.argv[1]); order = int(sys.argv[2])
for sw in [int(x) for x in sys.argv[3].split(",")]:
    out, tr = run(n, order, sw, reps=3)
    print(f"n={n} order={order} sweeps={sw}: pressure {out['pressure'][1]*1e3:.1f} us  momentum {out['momentum'][1]*1e3:.1f} us", flush=True)
