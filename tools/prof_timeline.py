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
ap.add_argument("--pvdl", action="store_true", help="BASELINE config 3: PVDL, 32 patches of 8192 points")
args = ap.parse_args()
ENG.OPTIONS.no_graph = not args.graph
dev = torch.device("cuda:0")
if args.pvdl:
    import yaml
    from tests.helpers import patch_input
    cd = yaml.safe_load(open(os.path.join(bench.ROOT, "p2pb_b200", "configs", "PVDL_SNPP.yaml")))
    cd["data"]["npoints"] = bench.PVDL_N; cd["model"]["extra_feature_channels"] = 0
    cfg = Config.wrap(cd)
    x = patch_input(bench.PVDL_B_PER_GPU, bench.PVDL_N, seed=7).to(dev)
else:
    cfg = Config.wrap(bench.load_cfg_dict())
    x = bench.synth_patches(64, bench.NPTS, seed=1000).to(dev)
cfg.gpu = str(dev); cfg.model.ema = False; cfg.backend = "engine"
net = PVCNN2Unet(cfg); net.load_state_dict(seeded_state_dict(net, seed=0), strict=True)
model = P2PB(cfg, net.to(dev)).eval()
run = lambda T: model.sample(x_start=x, steps=T, log_count=1, verbose=False, use_ema=False)["x_pred"]
run(args.steps); run(args.steps); torch.cuda.synchronize()       # steady state of THIS step count (tables, graph)
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
# step-to-step drift: per-kernel totals of the fastest and the slowest evaluation of the call
def window_totals(i):
    d = collections.defaultdict(float)
    for e in ev[heads[i] + 1:heads[i + 1] + 1]:
        d[re.sub(r"\(.*", "", e["name"].replace("(anonymous namespace)::", "").replace("void ", "")).strip()[:64]] += e["dur"]
    return d
if len(heads) > 3:
    per = [(ev[heads[i + 1]]["ts"] + ev[heads[i + 1]]["dur"]) - (ev[heads[i]]["ts"] + ev[heads[i]]["dur"]) for i in range(len(heads) - 1)]
    ia, ib = per.index(min(per)), per.index(max(per))
    da, db = window_totals(ia), window_totals(ib)
    lines += ["", f"fastest evaluation (#{ia + 1}, {per[ia]:.0f} us) vs slowest (#{ib + 1}, {per[ib]:.0f} us), kernels that differ by more than 5 us:"]
    for n in sorted(da, key=lambda n: -(db.get(n, 0) - da[n])):
        if abs(db.get(n, 0) - da[n]) > 5: lines.append(f"  {n:64s} {da[n]:8.0f} -> {db.get(n, 0):8.0f}")
    for nm in ("voxelize_sparse", "voxel_prep"):
        la = [e["dur"] for e in ev[heads[ia] + 1:heads[ia + 1] + 1] if nm in e["name"]]
        lb = [e["dur"] for e in ev[heads[ib] + 1:heads[ib + 1] + 1] if nm in e["name"]]
        lines.append(f"  {nm} per launch: " + " ".join(f"{x:.0f}->{y:.0f}" for x, y in zip(la, lb)))
# what one sample() call does outside its network evaluations
first_sv = next(i for i, e in enumerate(ev) if "step_vectors" in e["name"])
span = ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]
per_eval = [(ev[heads[i + 1]]["ts"] + ev[heads[i + 1]]["dur"]) - (ev[heads[i]]["ts"] + ev[heads[i]]["dur"]) for i in range(len(heads) - 1)]
lines += ["", f"whole sample() call ({args.steps} steps): first kernel -> last kernel {span:.0f} us; evaluation periods (head_bridge end to "
          f"head_bridge end): {', '.join(f'{p:.0f}' for p in per_eval)} us", "kernels before the first evaluation:"]
for e in ev[:first_sv]:
    lines.append(f"  {e['ts'] - ev[0]['ts']:8.0f} {e['dur']:7.1f}  {e['name'][:90]}")
lines.append("kernels after the last evaluation:")
for e in ev[heads[-1] + 1:]:
    lines.append(f"  {e['ts'] - ev[0]['ts']:8.0f} {e['dur']:7.1f}  {e['name'][:90]}")
open(args.out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:45]))
