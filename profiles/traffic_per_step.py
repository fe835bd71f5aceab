"""Per-step DRAM traffic and serialised kernel time from an ncu launch list:
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline
  python profiles/traffic_per_step.py launches.csv 8 > profiles/<round>_traffic_per_step.json
A step ends with the fused upsample+argmax kernel; launches after the last complete step are dropped."""
import collections
import csv
import json
import sys


def main(path, batch):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    # whole steps only: a step ends with the fused upsample+argmax kernel
    last = max((int(r["ID"]) for r in rows if "upsample_argmax" in r["Kernel Name"]), default=None)
    if last is not None:
        rows = [r for r in rows if int(r["ID"]) <= last]
    per = collections.OrderedDict()
    for row in rows:
        k = row["Kernel Name"][:40]
        d = per.setdefault(k, {"launches": 0, "us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        name = row["Metric Name"]
        if name == "gpu__time_duration.sum":
            d["launches"] += 1
            d["us"] += v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        else:
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            d["dram_read_bytes" if "read" in name else "dram_write_bytes"] += v * scale
    steps = max(1, sum(d["launches"] for k, d in per.items() if "upsample_argmax" in k))
    out = collections.OrderedDict()
    for k, d in per.items():
        out[k] = {"launches": d["launches"] / steps, "us": d["us"] / steps,
                  "dram_read_bytes": d["dram_read_bytes"] / steps, "dram_write_bytes": d["dram_write_bytes"] / steps}
    json.dump({"batch": batch, "steps_in_capture": steps,
               "note": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over bench.py steps "
                       "(1024x2048); cold-cache serialised launches, averaged per step",
               "per_step": out}, sys.stdout, indent=1)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 8)
