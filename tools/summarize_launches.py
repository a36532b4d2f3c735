"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel name -> markdown table."""
import collections, csv, re, sys

def main(path, norm_kernel=None):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("<unnamed>::", "").strip()[:70]
        agg[name][0] += 1; agg[name][1] += v; tot += v
    ne = agg[norm_kernel][0] if norm_kernel and norm_kernel in agg else 1
    n_k = sum(a[0] for a in agg.values())
    print(f"launches in window: {n_k}; network evaluations in window: {ne}; kernels per evaluation: {n_k / ne:.0f}; "
          f"GPU time per evaluation (serialised, cold cache): {tot / ne / 1e3:.2f} ms\n")
    print("| share | us / evaluation | launches / evaluation | kernel |\n|---:|---:|---:|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| {t / tot * 100:.2f} % | {t / ne:.0f} | {n / ne:.1f} | `{k}` |")

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
