"""Summarise ncu --set full reports (run HERE, no GPU needed): per launch duration, DRAM bytes, utilisation.
usage: python scripts/ncu_summary.py out.md rep1.ncu-rep [rep2 ...]"""
import csv
import io
import subprocess
import sys

WANT = [
    ("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram rd MB"),
    ("dram__bytes_write.sum", "dram wr MB"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"),
    ("launch__registers_per_thread", "regs"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
]


def to_unit(v, unit, want):
    v = float(v.replace(",", ""))
    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3}
    if want in ("dram rd MB", "dram wr MB", "us") and unit in scale:
        return v * scale[unit]
    return v


def main():
    out, reps = sys.argv[1], sys.argv[2:]
    lines = ["| report | # | kernel | grid | us | dram rd MB | dram wr MB | dram % | L2 % | warps act % | regs | L1 hit % | L2 hit % |", "|" + "---|" * 13]
    for rep in reps:
        txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        for n, r in enumerate(rows[2:]):
            cells = []
            for key, name in WANT:
                if key not in hdr:
                    cells.append("-")
                    continue
                i = hdr.index(key)
                v = r[i]
                if name in ("kernel", "grid"):
                    cells.append(v.replace("|", "/")[:48])
                else:
                    cells.append(f"{to_unit(v, units[i], name):.2f}")
            lines.append(f"| {rep.split('/')[-1]} | {n} | " + " | ".join(cells) + " |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


main()
