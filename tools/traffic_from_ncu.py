"""ncu launch list (csv with gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch; one UNet forward in
its real order, `--cache-control none`) -> profiles/r2_traffic.json {unet_batch: {gemm_dram_bytes_per_launch, ...}} and a
per-kernel table on stdout.

    python tools/traffic_from_ncu.py gpurun_out/r2c_launches_b32.csv 32 "tools/sessions/r2_session20.sh"
"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    path, batch, src = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    h = rows[0]
    ik, im, iv, iu, iid = (h.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    launches = {}
    for r in rows[1:]:
        d = launches.setdefault(int(r[iid]), {"k": r[ik]})
        v = float(r[iv].replace(",", ""))
        if r[im].startswith("dram__bytes"):
            v *= UNIT[r[iu]]
        d[r[im]] = v
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for d in launches.values():
        name = re.sub(r"^void |pnp::|\(anonymous namespace\)::|<unnamed>::|\(.*$", "", d["k"])
        fam = "gemm_tcgen05_kernel" if "gemm_tcgen05" in name else name
        for key in (name, "FAMILY " + fam) if fam != name else (name,):
            a = agg[key]
            a[0] += 1
            a[1] += d["gpu__time_duration.sum"] / 1e3
            a[2] += d["dram__bytes_read.sum"]
            a[3] += d["dram__bytes_write.sum"]
    total = sum(a[1] for k, a in agg.items() if not k.startswith("FAMILY"))
    print(f"one UNet forward at B = {batch}: {len(launches)} launches, sum of kernel durations {total / 1e3:.3f} ms")
    for k, a in sorted(agg.items(), key=lambda t: -t[1][1]):
        print(f"{k:48s} n={a[0]:3d} {a[1] / 1e3:7.3f} ms {100 * a[1] / total:5.1f} %  DRAM rd {a[2] / 1e9:6.2f} GB wr {a[3] / 1e9:6.2f} GB "
              f"-> {(a[2] + a[3]) / a[1] / 1e6:5.2f} TB/s")
    g = agg["FAMILY gemm_tcgen05_kernel"]
    out = os.path.join(ROOT, "profiles", "r2_traffic.json")
    data = json.load(open(out)) if os.path.exists(out) else {}
    data[str(batch)] = {"gemm_dram_bytes_per_launch": (g[2] + g[3]) / g[0], "gemm_launches": g[0],
                        "gemm_dram_read_bytes": g[2], "gemm_dram_write_bytes": g[3], "gemm_time_us": g[1],
                        "gemm_share_of_forward": g[1] / total,
                        "source": f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                                  f"--cache-control none over one B={batch} UNet forward ({src}); warm caches, launches in their real order"}
    with open(out, "w") as f:
        json.dump(data, f, indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
