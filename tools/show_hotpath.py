import json, sys
for l in open(sys.argv[1]):
    d = json.loads(l)
    print(d["name"], {k: round(v, 1) for k, v in d["gpu_us"].items()}, "frac", round(d["roofline"]["frac"], 3),
          "cpu", d.get("cpu_baseline", {}).get("value"), "gpu", round(d["value"], 1), d["unit"])
