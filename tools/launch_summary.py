"""ncu launch list (--csv, gpu__time_duration.sum [+ more metrics]) -> markdown table of kernel, launches, total us, share.
usage: python tools/launch_summary.py <launches.csv> [first_launch last_launch]"""
import csv, io, sys, collections

txt = open(sys.argv[1]).read()
rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
per = collections.OrderedDict()
for r in rows:
    per.setdefault(r["ID"], {"name": r["Kernel Name"]})[r["Metric Name"]] = r["Metric Value"]
launches = list(per.values())
if len(sys.argv) > 3:
    launches = launches[int(sys.argv[2]):int(sys.argv[3])]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for l in launches:
    name = l["name"].replace("void ", "")
    name = name.split("(")[0][:64]
    a = agg[name]
    a[0] += 1
    a[1] += float(l["gpu__time_duration.sum"]) / 1000.0
    a[2] += float(l.get("dram__bytes_read.sum", 0) or 0) / 1e6
    a[3] += float(l.get("dram__bytes_write.sum", 0) or 0) / 1e6
tot = sum(a[1] for a in agg.values())
has_dram = any(a[2] or a[3] for a in agg.values())
print("| kernel | launches | total us | share |" + (" dram rd MB | dram wr MB |" if has_dram else ""))
print("|---|---:|---:|---:|" + ("---:|---:|" if has_dram else ""))
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    line = f"| `{name}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f}% |"
    if has_dram:
        line += f" {a[2]:.1f} | {a[3]:.1f} |"
    print(line)
print(f"\ntotal {tot:.1f} us over {len(launches)} launches")
