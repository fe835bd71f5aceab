"""Joins an ncu per-launch CSV of ONE bench step (tools/profile_step.py) with the launch-order -> layer-name table:
one row per kernel launch with the layer's name, time, DRAM bytes vs the layer's algorithmic bytes, tensor-pipe %.
  python profiles/per_layer_table.py gpurun_out/step.csv gpurun_out/step_layers.json > profiles/r2_per_layer.txt"""
import collections
import csv
import json
import sys

H, W = 1024, 2048


def res_of(name):
    if name.startswith("model.base."):
        i = int(name.split(".")[2])
        return {0: 2, 1: 2, 2: 4, 3: 4, 4: 4, 5: 4, 7: 8, 8: 8, 10: 16, 11: 16, 13: 32, 14: 32, 16: 64, 17: 64}[i]
    if name.startswith("model.conv1x1_up.") or name.startswith("model.denseBlocksUp."):
        return [32, 16, 8, 4][int(name.split(".")[2])]
    return 4


def main(csv_path, layers_path):
    meta = json.load(open(layers_path))
    batch = meta["batch"]
    lines = [l for l in open(csv_path) if l.startswith('"')]
    launches = collections.OrderedDict()
    for row in csv.DictReader(lines):
        d = launches.setdefault(int(row["ID"]), {"kernel": row["Kernel Name"]})
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        n = row["Metric Name"]
        if n == "gpu__time_duration.sum":
            d["us"] = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        elif n.startswith("dram__bytes"):
            d[n] = v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        else:
            d[n] = v
    # launch order: every ConvLayer is one conv_halo launch, except the four conv1x1_up layers, which (fused with their
    # TransitionUp) are a low-resolution launch over the upsampled slices (fp32 partial sums out)
    # followed by a high-resolution launch over the skip slices that interpolates the partial in its epilogue
    skip_ch = [214, 160, 78, 48]
    convs = []
    for l in meta["layers"]:
        if l["type"] not in ("conv", "head"):
            continue
        if l["name"].startswith("model.conv1x1_up."):
            j = int(l["name"].split(".")[2])
            convs.append(dict(l, part="low", cin=l["cin"] - skip_ch[j]))
            convs.append(dict(l, part="high", cin=skip_ch[j]))
        else:
            convs.append(l)
    ci = 0
    pad = lambda c: (c + 15) // 16 * 16
    print("%-4s %-28s %-34s %-12s %9s %9s %9s %7s %8s %7s" % ("id", "kernel", "layer", "shape", "us", "dram MB", "algo MB", "x algo", "TFLOP/s", "tensor%"))
    tot = collections.Counter()
    for i, d in launches.items():
        kern = d["kernel"].split("(")[0].replace("void ", "")[:28]
        layer, shape, algo, tfl = "", "", None, None
        if "conv_halo" in d["kernel"] and ci < len(convs):
            c = convs[ci]; ci += 1
            r = res_of(c["name"])
            part = c.get("part")
            if part == "low":
                r *= 2
            px = batch * (H // r) * (W // r)
            layer, shape = c["name"] + (" [%s]" % part if part else ""), "%d->%d k%d 1/%d" % (c["cin"], c["cout"], c["k"], r)
            out_b = pad(c["cout"]) * 4 * px / (4 if (c["k"] == 1 and c["name"].startswith("model.base.") and r < 64) else 1)
            if c["type"] == "head":
                out_b = 16 * 4 * px
            algo = pad(c["cin"]) * 4 * px + out_b                 # split-bf16 storage: 4 B per element in and out
            if part == "high":
                algo += pad(c["cout"]) * 4 * px / 4               # + the fp32 partial sums of the low-resolution launch
            tfl = 2.0 * c["k"] * c["k"] * c["cin"] * c["cout"] * px / (d["us"] * 1e-6) / 1e12
        dram = d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
        tp = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
        print("%-4d %-28s %-34s %-12s %9.1f %9.1f %9s %7s %8s %7s" % (
            i, kern, layer, shape, d.get("us", 0), dram / 1e6, "%.1f" % (algo / 1e6) if algo else "-",
            "%.2f" % (dram / algo) if algo else "-", "%.0f" % tfl if tfl else "-", "%.1f" % tp if tp is not None else "-"))
        tot[kern + " us"] += d.get("us", 0)
        tot[kern + " MB"] += dram / 1e6
    print()
    for k in sorted(tot):
        print("%-40s %12.1f" % (k, tot[k]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
