"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import collections, csv, sys

def summarise(path, skip_prefix=("void at::",)):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            t = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        t = t / 1000.0 if row["Metric Unit"] == "ns" else (t * 1000.0 if row["Metric Unit"] == "ms" else t)
        a = agg.setdefault(row["Kernel Name"][:90], [0, 0.0, row.get("Grid Size"), row.get("Block Size")])
        a[0] += 1; a[1] += t
    tot = sum(v[1] for k, v in agg.items() if not k.startswith(skip_prefix))
    out = [f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.1f} us in our kernels (torch fill = L2 flush, excluded from shares)"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        share = "" if k.startswith(skip_prefix) else f"{100 * v[1] / tot:5.1f}%"
        out.append(f"{v[1]:10.1f} us  {v[0]:4d} calls  {v[1] / v[0]:9.1f} us/call  {share:>6}  grid {v[2]} block {v[3]}  {k}")
    return "\n".join(out)

if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(summarise(p)); print()
