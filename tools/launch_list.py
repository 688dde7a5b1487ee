"""ncu launch list (csv from `ncu --metrics gpu__time_duration.sum --csv --log-file X`) -> per-kernel totals as a markdown table.
Usage: launch_list.py gpurun_out/r02_launches.csv [top]"""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
rows = [r for r in csv.reader(open(path, errors="replace")) if r]
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]; kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(",", ""))
    except ValueError: continue
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[mu].strip().replace("usecond", "us").replace("nsecond", "ns").replace("msecond", "ms"), 1e-6)
    tot[r[kn]] += v * scale; cnt[r[kn]] += 1
allms = sum(tot.values())
print("| launches | total ms | share | kernel |\n|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:top]:
    print(f"| {cnt[k]} | {v:.3f} | {100 * v / allms:.1f} % | `{k[:110]}` |")
own = {k: v for k, v in tot.items() if "mpcg::" in k}
print("\nown kernels:", ", ".join(f"{k.split('(')[0].split('::')[-1]} {100 * v / sum(own.values()):.1f} %" for k, v in sorted(own.items(), key=lambda kv: -kv[1])))
