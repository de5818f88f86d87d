import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], round(d["value"]), "clips/s", round(d["ms_per_step"],3), "ms  e2e", round(d["e2e"]["value"]), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["per_class"].items()}, "frac", round(d["roofline"]["frac"],3), "step_frac", round(d["roofline"]["whole_step"]["frac"],3))
