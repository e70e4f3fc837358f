"""Aggregate an `ncu --csv --metrics ...` launch list by kernel: count, total time, share (+ extra metrics)."""
import csv, collections, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
rows = collections.OrderedDict()
for row in csv.DictReader(lines):
    key = row['ID']
    d = rows.setdefault(key, {'name': row['Kernel Name']})
    try:
        v = float(row['Metric Value'].replace(',', ''))
    except ValueError:
        continue
    unit = row['Metric Unit']
    if row['Metric Name'] == 'gpu__time_duration.sum':
        v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
    d[row['Metric Name']] = v
agg = collections.OrderedDict()
for d in rows.values():
    name = d['name'].split('(')[0].replace('void ', '')[:64]
    a = agg.setdefault(name, collections.defaultdict(list))
    for k, v in d.items():
        if k != 'name':
            a[k].append(v)
tot = sum(sum(a['gpu__time_duration.sum']) for a in agg.values())
metrics = [m for m in next(iter(agg.values())).keys() if m != 'gpu__time_duration.sum']
print("| kernel | launches | total us | share |" + "".join(f" {m} (mean per launch) |" for m in metrics))
print("|---|---:|---:|---:|" + "---:|" * len(metrics))
for name, a in sorted(agg.items(), key=lambda kv: -sum(kv[1]['gpu__time_duration.sum'])):
    t = a['gpu__time_duration.sum']
    extra = "".join(f" {sum(a[m]) / max(1, len(a[m])):.3g} |" for m in metrics)  # mean per launch
    print(f"| `{name}` | {len(t)} | {sum(t):.1f} | {sum(t) / tot:.1%} |{extra}")
print(f"\ntotal {tot:.1f} us over {sum(len(a['gpu__time_duration.sum']) for a in agg.values())} launches")
