#!/usr/bin/env python
"""profiles/traffic.json from the CSV export of an `ncu --set full` capture (scripts/gpu_round2.sh): DRAM bytes per launch
(dram__bytes_read.sum + dram__bytes_write.sum) of every captured kernel, mean over its launches with L = 2 (the mean-field
iterations; the L = 1 norm launches are listed separately).  bench.py quotes the dominant kernel's figure as
roofline.traffic when the capture was taken at the step size it times.
Usage: ncu_traffic.py <prof_raw.csv[.gz]> <points per step of the capture> <out.json> <source note>"""
import collections, csv, gzip, io, json, re, sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "B": 1.0, "KB": 1e3, "MB": 1e6, "GB": 1e9}


def main(path, points, out, note):
    f = io.TextIOWrapper(gzip.open(path), newline="") if path.endswith(".gz") else open(path, newline="")
    rows = list(csv.DictReader(l for l in f if not l.startswith("==")))
    units, data = rows[0], rows[1:]
    agg = collections.OrderedDict()
    for r in data:
        full = r["Kernel Name"]
        base = re.sub(r"<.*", "", full.split("(")[0]).split("::")[-1].strip()
        m = re.search(r"<([^>]*)>", full)
        tag = base + (" <%s>" % m.group(1) if m else "")
        b = sum(float(r[k].replace(",", "")) * UNIT[units[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        t = float(r["gpu__time_duration.sum"].replace(",", "")) * {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "second": 1e6, "s": 1e6}[units["gpu__time_duration.sum"]]
        agg.setdefault(tag, []).append((b, t))
    res = {"_source": note, "_captured_points_per_step": int(points), "_per_variant": {}}
    by_base = collections.OrderedDict()
    for tag, lst in agg.items():
        res["_per_variant"][tag] = {"launches": len(lst), "dram_bytes_per_launch": sum(b for b, _ in lst) / len(lst),
                                    "mean_us_under_ncu": sum(t for _, t in lst) / len(lst)}
        by_base.setdefault(tag.split(" <")[0], []).extend(lst)
    for base, lst in by_base.items():
        res[base] = sum(b for b, _ in lst) / len(lst)
    json.dump(res, open(out, "w"), indent=1)
    for k, v in res.items():
        if not k.startswith("_"):
            print("%-24s %10.1f MB per launch" % (k, v / 1e6))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4])
