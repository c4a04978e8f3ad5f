"""profiles/traffic.json from an ncu metrics pass over one eager train step:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/step_metrics.csv python scripts/prof_step.py        (STEPS=1)
    python scripts/make_traffic.py gpurun_out/step_metrics.csv profiles/traffic.json

Per kernel function: launches, DRAM bytes (read + write) per launch and in total, device time.  bench.py reads the file
to fill roofline.traffic for the dominant kernel.
"""
import collections
import csv
import json
import re
import sys


def main(src, dst):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    ids = collections.defaultdict(set)
    for row in csv.DictReader(lines):
        name = re.sub(r"^void ", "", row["Kernel Name"])
        name = re.sub(r"^semb::", "", name)
        name = re.sub(r"[<(].*", "", name)
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
        per[name][row["Metric Name"]] += v * scale
        ids[name].add(row["ID"])
    out = {}
    for name, m in per.items():
        n = len(ids[name])
        total = m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        out[name] = {"launches": n, "dram_bytes_total": total, "dram_bytes_per_launch": total / n,
                     "dram_read_bytes": m.get("dram__bytes_read.sum", 0.0), "dram_write_bytes": m.get("dram__bytes_write.sum", 0.0),
                     "time_us_total": m.get("gpu__time_duration.sum", 0.0),
                     "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum, one eager train step (batch 32, 256x256), {src}"}
    with open(dst, "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    for k, v in sorted(out.items(), key=lambda kv: -kv[1]["dram_bytes_total"])[:12]:
        print(f"{k:40s} {v['launches']:4d} launches  {v['dram_bytes_total'] / 1e9:7.2f} GB  {v['time_us_total'] / 1e3:7.2f} ms")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
