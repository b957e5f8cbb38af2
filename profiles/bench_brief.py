import json,sys
d=json.loads(open(sys.argv[1]).read())
print("value %.0f e2e %.0f ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
print({k: round(v,3) for k,v in d["stages_ms"]["prefilter"].items()})
print({k: round(v,3) for k,v in d["stages_ms"]["align"].items()})
