#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and share of the last step.
usage: summarize_launches.py launches.csv launches_per_step [out.md]"""
import csv, sys, re, collections
path, per_step = sys.argv[1], int(sys.argv[2])
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
rows = rows[-per_step:]
def short(name):
    name = re.sub(r"^void ", "", name)
    m = re.match(r"([A-Za-z0-9_]+)(<[^(]*>)?\(", name)
    base = m.group(1) if m else name[:60]
    if "for_each_kernel" in base or "compact" in base:
        lam = re.search(r"\[lambda\(unsigned long(?:, unsigned int)?\) \(instance (\d+)\)\]", name)
        fn = re.search(r"(?:T\d+=|<)([A-Za-z_0-9:<>, ]*?)::?(?:operator|\[lambda)", name)
        where = re.search(r"((?:Engine<[^>]*>::|kc_|run_)[A-Za-z_0-9]+)", name[len(base):])
        base += " @" + (where.group(1) if where else "?") + (f"#{lam.group(1)}" if lam else "")
    targs = re.match(r"[A-Za-z0-9_]+<([0-9, ]+)>", name)
    if targs: base += f"<{targs.group(1)}>"
    return base
agg = collections.OrderedDict()
tot = 0.0
for r in rows:
    ns = float(r["Metric Value"].replace(",", ""))
    k = short(r["Kernel Name"])
    a = agg.setdefault(k, [0, 0.0, r["Grid Size"], r["Block Size"]])
    a[0] += 1; a[1] += ns; tot += ns
out = ["| kernel | launches | total us | share | grid (first) | block |", "|---|---:|---:|---:|---|---|"]
for k, (n, ns, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {n} | {ns/1000:.1f} | {100*ns/tot:.1f}% | {g} | {b} |")
out.append(f"| **total ({len(rows)} launches)** | | {tot/1000:.1f} | 100% | | |")
text = "\n".join(out)
print(text)
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write(text + "\n")
