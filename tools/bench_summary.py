import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("train ms", round(d["ms_per_step"], 2), "value", round(d["value"]), "inf ms", round(d["inference"]["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]),
      "launches", d["gpu_launches"], "clocks", d["clocks"])
for k in (d["kernels"] or [])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print("%8.3f ms %5.1f%% n=%5.1f %s tf=%s gbs=%s" % (k["ms_per_step"], 100 * k["share"], k["launches_per_step"], k["site"], round(k.get("tflops", 0), 1), round(k.get("gbs", 0), 1)))
print("cpu", d.get("cpu_baseline"))
