"""Summarise an `ncu --metrics ... --csv` capture (tools/gpu_ncu_all.sh) -> markdown tables for profiles/.

usage: python tools/summarize_ncu.py gpurun_out/ncu_eval.csv [--eval head_bridge_kernel] [--hbm 6550.7]

Per kernel NAME (all launches of one network evaluation): launches, total time, DRAM bytes read+written, achieved DRAM GB/s and the
fraction of the measured copy bandwidth, tensor-pipe active %, shared-memory pipe % (tensor-core operand reads / LSU).  With
--eval KERNEL only the launches between the first and the second launch of KERNEL (one complete evaluation) are used.
ncu runs every kernel alone with a cold cache: the absolute times are upper bounds of what the graph achieves; the DRAM byte counts
and the pipe percentages are the evidence."""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    launches = collections.OrderedDict()
    for row in csv.DictReader(lines):
        i = int(row["ID"])
        d = launches.setdefault(i, {"name": re.sub(r"\(.*", "", row["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "").strip(),
                                    "grid": row.get("Grid Size", ""), "block": row.get("Block Size", "")})
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        u = row["Metric Unit"]
        name = row["Metric Name"]
        if name == "gpu__time_duration.sum":
            v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else (v * 1e6 if u in ("s", "second") else v))
        if name.startswith("dram__bytes"):
            mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "B": 1, "KB": 1e3, "MB": 1e6, "GB": 1e9}.get(u, 1)
            v *= mul
        d[name] = v
    return list(launches.values())


def main():
    path = sys.argv[1]
    ev = sys.argv[sys.argv.index("--eval") + 1] if "--eval" in sys.argv else None
    hbm = float(sys.argv[sys.argv.index("--hbm") + 1]) if "--hbm" in sys.argv else 6550.7
    L = load(path)
    if ev:
        marks = [i for i, d in enumerate(L) if d["name"].startswith(ev)]
        if len(marks) >= 2:
            L = L[marks[0] + 1: marks[1] + 1]
    T = "gpu__time_duration.sum"
    tot = sum(d.get(T, 0.0) for d in L)
    print(f"launches: {len(L)}; serialised cold-cache GPU time: {tot / 1e3:.2f} ms; DRAM traffic: "
          f"{sum(d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0) for d in L) / 1e9:.2f} GB\n")
    agg = collections.OrderedDict()
    for d in L:
        a = agg.setdefault(d["name"], {"n": 0, "t": 0.0, "rd": 0.0, "wr": 0.0, "tensor": 0.0, "tc": 0.0, "lsu": 0.0, "lts": 0.0, "regs": 0})
        t = d.get(T, 0.0)
        a["n"] += 1
        a["t"] += t
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
        a["tensor"] += t * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0)
        a["tc"] += t * d.get("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", 0.0)
        a["lsu"] += t * d.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", 0.0)
        a["lts"] += t * d.get("lts__throughput.avg.pct_of_peak_sustained_elapsed", 0.0)
        a["regs"] = max(a["regs"], int(d.get("launch__registers_per_thread", 0)))
    print("| kernel | launches | time (us) | share | DRAM read (MB) | DRAM write (MB) | DRAM GB/s | frac of measured %.0f GB/s | tensor pipe active %% | smem pipe %% (tensor-core operands / LSU) | L2 throughput %% | regs |" % hbm)
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
        t = a["t"]
        gbs = (a["rd"] + a["wr"]) / t / 1e3 if t > 0 else 0.0
        w = (lambda key: a[key] / t if t > 0 else 0.0)
        print(f"| `{k[:58]}` | {a['n']} | {t:.0f} | {t / tot * 100:.1f} % | {a['rd'] / 1e6:.1f} | {a['wr'] / 1e6:.1f} | {gbs:.0f} | {gbs / hbm:.2f} | "
              f"{w('tensor'):.1f} | {w('tc'):.1f} / {w('lsu'):.1f} | {w('lts'):.1f} | {a['regs']} |")
    if "--launches" in sys.argv:
        pat = re.compile(sys.argv[sys.argv.index("--launches") + 1])
        print("\n| # | kernel | grid | time (us) | DRAM read (MB) | DRAM write (MB) | GB/s | tensor % |\n|---:|---|---|---:|---:|---:|---:|---:|")
        for i, d in enumerate(L):
            if pat.search(d["name"]):
                t = d.get(T, 0.0)
                rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
                print(f"| {i} | `{d['name'][:48]}` | {d['grid']} | {t:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {(rd + wr) / t / 1e3 if t else 0:.0f} | "
                      f"{d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 0.0):.1f} |")


if __name__ == "__main__":
    main()
