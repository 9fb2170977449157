"""profiles/traffic.json from ncu --set full reports (run HERE, no GPU needed).
usage: python scripts/ncu_traffic.py workload=report.ncu-rep [workload=report ...]
Per kernel class: dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the class's FINE-LEVEL launches of the
report (the launches with the largest grid of that kernel family)."""
import csv
import io
import json
import os
import subprocess
import sys

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
CLASSES = [("k_stencil<0", "jacobi_interior"), ("k_stencil<3", "jacobi_interior"), ("k_stencil<1", "apply_poisson"), ("k_stencil<2", "residual"),
           ("k_stencil_tma<0", "jacobi_interior"), ("k_stencil_tma<1", "apply_poisson"), ("k_stencil_tma<2", "residual"), ("k_band", "band_jacobi"),
           ("k_restrict", "restrict"), ("k_prolong", "prolong_add"), ("k_zero", "zero_fill"), ("k_vec<3", "reduce"), ("k_vec<4", "reduce"),
           ("k_vec", "blas1"), ("k_gauss_seidel", "gauss_seidel")]


def klass(name):
    name = name.replace("void ", "").replace("gmg::", "")
    for prefix, c in CLASSES:
        if name.startswith(prefix):
            return c
    return None


def main():
    out = {}
    path = os.environ.get("GMG_TRAFFIC_JSON") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
    if os.path.exists(path):
        out = json.load(open(path))
    for arg in sys.argv[1:]:
        workload, rep = arg.split("=", 1)
        txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        ir, iw, ig, ik = col["dram__bytes_read.sum"], col["dram__bytes_write.sum"], col["Grid Size"], col["Kernel Name"]
        per = {}
        for r in rows[2:]:
            c = klass(r[ik])
            if not c:
                continue
            grid = int(r[ig].strip("()").split(",")[0])
            b = float(r[ir].replace(",", "")) * SCALE[units[ir]] + float(r[iw].replace(",", "")) * SCALE[units[iw]]
            per.setdefault(c, []).append((grid, b))
        res = {}
        for c, v in per.items():
            gmax = max(g for g, _ in v)
            fine = [b for g, b in v if g >= 0.9 * gmax]
            res[c] = round(sum(fine) / len(fine))
        out[workload] = {"report": os.path.basename(rep), "how": "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the fine-level launches", "bytes_per_launch": res}
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
