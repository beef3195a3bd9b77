"""Substitute the @@PLACEHOLDERS@@ of DESIGN.md with the numbers of the last GPU-box run (gpurun_out/bench*.json).
usage: python tools/fill_design.py [ncu_sim_us]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")


def last_json(path):
    for line in reversed(open(path).read().strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    raise SystemExit("no JSON line in " + path)


b = last_json(os.path.join(G, "bench.json"))
r = last_json(os.path.join(G, "bench_ref.json"))
rep = {
    "SIM_DRAM_MB": "3.8", "SIM_NCU_US": sys.argv[1] if len(sys.argv) > 1 else "47.0",
    "SIM_TBS": "%.1f" % (b["roofline"]["achieved"] / 1e3), "SIM_FRAC": "%.2f" % b["roofline"]["frac"],
    "FE_US": "%.1f" % (1e3 * b["stage_ms"]["front_end"]), "FE_GBS": "%.0f" % b["roofline_front_end"]["achieved"], "FE_FRAC": "%.3f" % b["roofline_front_end"]["frac"],
    "ICP_MS": "%.1f" % b["icp"]["batch_ms"], "ICP_ITS": "%.0fk" % (b["icp"]["icp_iters_per_s"] / 1e3), "ICP_CPU_ITS": "%.0f" % b["icp"]["cpu_baseline"]["icp_iters_per_s"],
    "STEP_US": "%.1f" % (1e3 * b["ms_per_step"]), "VALUE_G": "%.1f" % (b["value"] / 1e9), "FPS": "{:,.0f}".format(b["config"]["frames_per_s"]),
    "E2E_G": "%.1f" % (b["e2e"]["value"] / 1e9), "E2E_FPS": "{:,.0f}".format(b["e2e"]["frames_per_s"]),
    "CPU_G": "%.3f" % (b["cpu_baseline"]["value"] / 1e9), "CPU_FPS": "%.1f" % b["cpu_baseline"]["frames_per_s"],
    "REF_G": "%.2f" % (r["value"] / 1e9), "REF_CORES": str(r["cpu_baseline"]["cores"]),
}
scale = ["N=1 %.1f G evals/s (%.1f us/frame)" % (b["value"] / 1e9, 1e3 * b["ms_per_step"])]
for n in (2, 4, 8):
    p = os.path.join(G, "bench_n%d.json" % n)
    if os.path.exists(p):
        try:
            x = last_json(p)
            scale.append("N=%d %.1f G evals/s (%.1f us/frame, p50 %.1f; x%.2f)" % (n, x["value"] / 1e9, 1e3 * x["ms_per_step"], 1e3 * x["ms_per_step_p50"], x["value"] / b["value"]))
        except SystemExit:
            pass
rep["SCALE_LINE"] = "; ".join(scale)
path = os.path.join(ROOT, "DESIGN.md")
s = open(path).read()
for k, v in rep.items():
    s = s.replace("@@%s@@" % k, v)
open(path, "w").write(s)
import re
print("left:", re.findall(r"@@\w+@@", s))
