import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(sys.argv[1], round(d["value"]/1e6,2), "M/s  step", round(d["ms_per_step"],4), "ingest", round(d["kernels"]["ingest"]["ms"],4), "observe", round(d["kernels"]["observe"]["ms"],4))
