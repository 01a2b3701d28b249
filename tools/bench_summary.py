#!/usr/bin/env python
"""One-screen summary of a bench.py JSON line:  python bench.py | python tools/bench_summary.py"""
import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print("value", round(d["value"]), d["unit"], "| ms/step", round(d["ms_per_step"], 4), "| e2e", round(d["e2e"]["value"]),
      "| launches", d.get("gpu_launches"))
print("kernels ms:", {k: round(v["ms_per_step"], 3) for k, v in d.get("kernels", {}).items()})
print("roofline fracs:", {k: round(v["frac"], 3) for k, v in d.get("rooflines", {}).items() if v.get("frac")})
o = d.get("other_workloads") or {}
if "cfg2" in o and "error" not in o["cfg2"]:
    print("cfg2 ms", round(o["cfg2"]["ms_per_step"], 4), "gemm frac", round(o["cfg2"]["roofline"]["frac"], 3))
if "cfg5" in o and "error" not in o["cfg5"]:
    c = o["cfg5"]
    print("cfg5 ms first/reference", round(c["first"]["ms_per_solve"], 2), round(c["reference"]["ms_per_solve"], 2),
          "eval frac", round(c["roofline"]["frac"], 3), round(c["roofline"]["frac_of_register_operand_peak"], 3))
if "cfg4" in o and "error" not in o["cfg4"]:
    for m in ("adaptive", "fixed_H"):
        c = o["cfg4"][m]
        print("cfg4", m, round(c["frames_per_s"]), "frames/s", {k: round(v["ms_per_step"], 3) for k, v in c["kernels"].items()})
if "cfg1" in o and "error" not in o["cfg1"]:
    print("cfg1 ms/pair", {k: round(v["ms_per_pair_device_resident"], 4) for k, v in o["cfg1"].items() if isinstance(v, dict)})
if "dr_ye" in o and "error" not in o["dr_ye"]:
    c = o["dr_ye"]
    print("dr_ye", round(c["pairs_per_s"]), "pairs/s", {k: round(v["ms_per_step"], 3) for k, v in c["kernels"].items()})
if "frames" in o and "error" not in o["frames"]:
    c = o["frames"]
    print("frames maps", round(c["maps"]["frames_per_s"]), "frames/s, HBM frac", round(c["maps"]["roofline"]["frac"], 3),
          "| fused", round(c["fused_features"]["frames_per_s"]), "frames/s, HBM frac",
          round(c["fused_features"]["roofline"]["frac"], 3))
if "ekf_update" in o and "error" not in o["ekf_update"]:
    c = o["ekf_update"]
    print("ekf_update", round(c["frames_per_s"]), "frames/s", round(c["fp64_tflops_executed"], 2), "TFLOP/s fp64 executed |",
          "cpu", c.get("cpu"))
for k, v in o.items():
    if isinstance(v, dict) and "error" in v:
        print("ERROR in", k, v["error"])
