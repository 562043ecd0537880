"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel for the LAST full step."""
import collections, csv, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [x["Kernel Name"] for x in rows]
idx = [i for i, n in enumerate(names) if "embedding_fwd" in n]
start, end = (idx[-2], idx[-1]) if len(idx) >= 2 else (idx[-1], len(rows))
sub = rows[start:end]
agg, tot = collections.OrderedDict(), 0.0
for x in sub:
    n = x["Kernel Name"].split("(")[0]
    v = float(x["Metric Value"].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(x["Metric Unit"], 1)
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print(f"launches in step: {len(sub)}  total {tot/1e6:.3f} ms")
print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{n[:90]}` | {c} | {v/1e6:.3f} | {100*v/tot:.1f}% |")
