import json
d = json.loads(open("gpurun_out/r2h_api.json").read().strip().splitlines()[-1])
print(d["fit_api"])
