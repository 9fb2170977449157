"""Per-launch table from ncu --set full reports (run HERE, no GPU needed): duration, DRAM bytes, stall picture.
usage: python scripts/ncu_table.py rep.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "us", 1e-3), ("dram__bytes_read.sum", "rdMB", None), ("dram__bytes_write.sum", "wrMB", None),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1), ("lts__t_sector_hit_rate.pct", "L2hit%", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%", 1), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1),
        ("launch__registers_per_thread", "regs", 1), ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "ldMsect", 1e-6),
        ("smsp__inst_executed.sum", "Minst", 1e-6)]
SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def main():
    for rep in sys.argv[1:]:
        txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]

        def col(name):
            for i, h in enumerate(hdr):
                if h == name or h.endswith("." + name):
                    return i
            return None

        print(f"## {rep.split('/')[-1]}")
        print("| # | kernel | grid | " + " | ".join(k[1] for k in KEYS) + " | top stall |")
        print("|" + "---|" * (len(KEYS) + 4))
        for n, r in enumerate(rows[2:]):
            cells = []
            for key, name, scale in KEYS:
                i = col(key)
                if i is None or r[i] == "":
                    cells.append("-")
                    continue
                v = float(r[i].replace(",", ""))
                if scale is None:
                    v *= SCALE.get(units[i], 1.0)
                else:
                    v *= scale
                    if name == "us" and units[i] == "us":
                        v *= 1e3
                    if name == "us" and units[i] == "ms":
                        v *= 1e6
                cells.append(f"{v:.1f}")
            stalls = []
            for i, h in enumerate(hdr):
                if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                    try:
                        stalls.append((float(r[i]), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
                    except ValueError:
                        pass
            top = ", ".join(f"{nm} {v:.0f}" for v, nm in sorted(stalls, reverse=True)[:2])
            print(f"| {n} | {r[4][:40]} | {r[8]} | " + " | ".join(cells) + f" | {top} |")


main()
