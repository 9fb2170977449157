"""Print the interesting parts of bench.py JSON lines. usage: python scripts/show_bench.py file.json [...]"""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e)
        print(open(f).read()[-1500:])
        continue
    print("==", f)
    for k in ["metric", "value", "iterations", "vcycle_ms", "vcycle_frac_of_hbm_peak", "setup_ms", "gpu_launches", "nccl_ops"]:
        if k in d:
            print("  ", k, ":", d[k])
    if "e2e" in d and d["e2e"]:
        print("   e2e:", d["e2e"]["value"], "setup", d["e2e"].get("setup_ms"))
    if "roofline" in d:
        r = d["roofline"]
        print("   roofline:", r["kernel"] if "kernel" in r else "", "%.0f GB/s frac %.3f" % (r["achieved"], r["frac"]))
    if "gauss_seidel" in d and d["gauss_seidel"]:
        print("   gs:", d["gauss_seidel"]["solve_ms"])
    for k, v in (d.get("fine_level_kernels") or {}).items():
        print("     %-16s %8.1f us x%d  %.0f GB/s  frac %.3f" % (k, v["us_per_launch"], v["launches_per_vcycle"], v["algorithmic_gbs"], v["frac_of_hbm_peak"]))
    for l, row in (d.get("kernels_by_level") or {}).items():
        per = "solve" if any("ms_per_solve" in v for v in row.values()) else "vcycle"
        parts = ["%s %.3f/%d" % (k[:8], v["ms_per_" + per], v["launches_per_" + per]) for k, v in row.items() if v.get("launches_per_" + per)]
        if parts:
            print("     L%s: %s" % (l, "  ".join(parts)))
