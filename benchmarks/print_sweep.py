import json,sys
d=json.loads(sys.stdin.read())
print([(s["batch"], round(s["ms_per_step"],3), round(s["tensor_tflops"])) for s in d["sweep"]])
