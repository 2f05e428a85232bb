import json, sys
d = json.load(open(sys.argv[1]))
print("ms/step", round(d["ms_per_step"], 3), "device", round(d.get("device_ms_per_step", 0), 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "value %.4g" % d["value"], "gpus", d["n_gpus"])
for k, v in sorted(d["stages"].items(), key=lambda kv: -kv[1]["ms_per_step"]):
    print("%-18s %7.3f ms  n=%5.0f  %7.1f us/launch  %s" % (k, v["ms_per_step"], v["launches_per_step"], v["us_per_launch"], ("%.0f GB/s" % v["achieved_GBs"]) if "achieved_GBs" in v else ""))
print(d["stage_ms_last_step"])
