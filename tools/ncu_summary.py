"""`ncu -i x.ncu-rep --page raw --csv` -> one line per launch with the counters the docs quote (profiles/*_summary.csv).

    python tools/ncu_summary.py gpurun_out/r2_full_raw.csv profiles/r2b_ncu_full_b32_summary.csv "header comment"
"""
import csv
import re
import sys

WANT = [  # (column title, substrings the metric name must contain)
    ("dur [us]", ("gpu__time_duration.sum",)),
    ("tensor%", ("sm__pipe_tensor", "pct_of_peak_sustained_active")),
    ("dram_rd [MB]", ("dram__bytes_read.sum",)),
    ("dram_wr [MB]", ("dram__bytes_write.sum",)),
    ("l2hit%", ("lts__t_sector_hit_rate.pct",)),
    ("occ%", ("sm__warps_active.avg.pct_of_peak_sustained_active",)),
    ("regs", ("launch__registers_per_thread",)),
    ("xu%", ("sm__inst_executed_pipe_xu", "pct_of_peak_sustained_active")),
    ("issue%", ("smsp__issue_active.avg.pct_of_peak_sustained_active",)),
    ("dram%", ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",)),
    ("sm%", ("sm__throughput.avg.pct_of_peak_sustained_elapsed",)),
    ("grid", ("launch__grid_size",)),
]
SCALE = {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
         "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def main():
    src, dst = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = list(csv.reader(l for l in open(src) if not l.startswith("==")))
    head, units, data = rows[0], rows[1], rows[2:]
    kcol = head.index("Kernel Name")
    cols = []
    for title, keys in WANT:
        idx = next((i for i, h in enumerate(head) if all(k in h for k in keys)), None)
        cols.append((title, idx))
    with open(dst, "w") as f:
        if note:
            f.write(f"# {note}\n")
        f.write("id,kernel," + ",".join(t for t, _ in cols) + "\n")
        for n, r in enumerate(data):
            name = re.sub(r"^void |pnp::|\(anonymous namespace\)::|<unnamed>::|\(.*$", "", r[kcol]).replace("unnamed>::", "")
            vals = []
            for title, idx in cols:
                if idx is None or idx >= len(r) or r[idx] == "":
                    vals.append("")
                    continue
                try:
                    v = float(r[idx].replace(",", ""))
                except ValueError:  # "no data" / "n/a"
                    vals.append("")
                    continue
                v *= SCALE.get(units[idx], 1.0) if ("[us]" in title or "[MB]" in title) else 1.0
                vals.append(f"{v:.3f}" if v != int(v) or "%" in title else str(int(v)))
            f.write(f'{n},"{name}",' + ",".join(vals) + "\n")
    print("wrote", dst, len(data), "launches; missing columns:", [t for t, i in cols if i is None])


if __name__ == "__main__":
    main()
