"""Dev tool: in-situ kernel timeline of the bench engine (PVDS, 64 patches, T = 30) from CUPTI via torch.profiler.

Unlike an ncu launch list (every kernel alone, cold caches, serialised) this records the kernels as they really run: warm L2,
back to back, streams overlapping.  Prints per-kernel totals of ONE network evaluation (the window between two head_bridge launches)
and how much of the evaluation's wall time the main stream is idle.
usage: python tools/prof_timeline.py [--graph] [--out gpurun_out/timeline.md]"""
import argparse, collections, json, os, re, sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
from p2pb_b200 import engine as ENG
from p2pb_b200.config import Config
from p2pb_b200.model_loader import seeded_state_dict
from p2pb_b200.p2pb import P2PB
from p2pb_b200.unet_pvc import PVCNN2Unet

ap = argparse.ArgumentParser()
ap.add_argument("--graph", action="store_true")
ap.add_argument("--out", default="gpurun_out/timeline.md")
ap.add_argument("--steps", type=int, default=6)
args = ap.parse_args()
ENG.OPTIONS.no_graph = not args.graph
dev = torch.device("cuda:0")
cfg = Config.wrap(bench.load_cfg_dict()); cfg.gpu = str(dev); cfg.model.ema = False; cfg.backend = "engine"
net = PVCNN2Unet(cfg); net.load_state_dict(seeded_state_dict(net, seed=0), strict=True)
model = P2PB(cfg, net.to(dev)).eval()
x = bench.synth_patches(64, bench.NPTS, seed=1000).to(dev)
run = lambda T: model.sample(x_start=x, steps=T, log_count=1, verbose=False, use_ema=False)["x_pred"]
run(bench.TSTEPS); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run(args.steps); torch.cuda.synchronize()
os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
trace = args.out.replace(".md", ".json")
prof.export_chrome_trace(trace)
ev = [e for e in json.load(open(trace))["traceEvents"] if e.get("cat") == "kernel"]
os.remove(trace)
ev.sort(key=lambda e: e["ts"])
heads = [i for i, e in enumerate(ev) if "head_bridge" in e["name"]]
assert len(heads) >= 3, f"only {len(heads)} evaluations in the trace ({len(ev)} kernels)"
lo, hi = heads[-3] + 1, heads[-2] + 1          # one complete evaluation in the middle of the call
win = ev[lo:hi]
t0, t1 = ev[lo - 1]["ts"] + ev[lo - 1]["dur"], win[-1]["ts"] + win[-1]["dur"]
streams = collections.Counter(e["args"].get("stream") for e in win)
main = streams.most_common(1)[0][0]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for e in win:
    n = re.sub(r"\(.*", "", e["name"].replace("(anonymous namespace)::", "").replace("void ", "")).strip()[:64]
    a = agg[n]; a[0] += 1; a[1] += e["dur"]; a[2] += e["dur"] if e["args"].get("stream") == main else 0.0
busy_main = sum(e["dur"] for e in win if e["args"].get("stream") == main)
# union of all kernel intervals = time the GPU runs anything
iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in win)
union, cs, ce = 0.0, iv[0][0], iv[0][1]
for s, e_ in iv[1:]:
    if s > ce: union += ce - cs; cs, ce = s, e_
    else: ce = max(ce, e_)
union += ce - cs
lines = [f"one evaluation: {len(win)} kernels, wall {t1 - t0:.0f} us, GPU busy (union of kernels) {union:.0f} us, "
         f"main-stream kernels {busy_main:.0f} us, streams {dict(streams)}; graph={args.graph}", "",
         "| kernel | launches | us (sum) | us on main stream | share of wall |", "|---|---:|---:|---:|---:|"]
for n, (c, t, tm) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"| `{n}` | {c} | {t:.0f} | {tm:.0f} | {100 * t / (t1 - t0):.1f} % |")
lines += ["", "per launch, in order (main stream = *):", ""]
for e in win:
    n = re.sub(r"\(.*", "", e["name"].replace("(anonymous namespace)::", "").replace("void ", ""))[:60]
    lines.append(f"{'*' if e['args'].get('stream') == main else ' '} {e['ts'] - t0:8.0f} {e['dur']:7.1f}  {n}  grid={e['args'].get('grid')}")
open(args.out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:45]))
