import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["step_breakdown_ms"].items()}, d["inner_sweeps_per_step"])
