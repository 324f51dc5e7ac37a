import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
print("C2 value %.1f e2e %.1f warm %.1f launches/frame %.1f ms/step %.4f"%(d["value"], d["e2e"]["value"], d["value_warm_l2"], d["gpu_launches"]/d["steps"], d["ms_per_step"]))
for k,v in d["kernels"].items(): print("  %-26s n=%5.1f us=%8.1f share=%.3f"%(k,v["launches_per_frame"],v["us_per_frame"],v["share"]))
print(d["roofline"]); print(d.get("cpu_baseline")); print(d.get("clocks"))
for name,c in d.get("workloads",{}).items():
    print(name.upper(), {k:v for k,v in c.items() if k not in ("kernels","trace")})
    for k,v in c.get("kernels",{}).items(): print("  %-26s n=%4d ms=%8.3f share=%.3f"%(k,v["launches"],v["ms"],v["share"]))
