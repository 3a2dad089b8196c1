import json,sys
d=json.load(open(sys.argv[1]))
print(d["value"], d["ms_per_step"], d["clocks"])
for k,v in d["kernels"].items(): print(k, v["ms"], round(v["achieved"]), v["unit"], round(v["frac"],3))
