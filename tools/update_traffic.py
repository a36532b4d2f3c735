"""Writes profiles/roofline_traffic.json from an `ncu --set full` report of the dominant kernel (tools/gpu_prof_conv.sh 100 ->
gpurun_out/prof_halo_r2_G100.ncu-rep), keyed by the sha256 of the kernel source the capture was taken from.  Run it HERE (no GPU
needed) right after the capture, before touching csrc/conv_halo.cu again.
usage: python tools/update_traffic.py [report]"""
import csv, hashlib, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "prof_halo_r2_G100.ncu-rep")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2]
def metric(name):
    i = hdr.index(name)
    v, u = float(vals[i].replace(",", "")), units[i]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
rd, wr = metric("dram__bytes_read.sum"), metric("dram__bytes_write.sum")
kname = vals[hdr.index("Kernel Name")]
assert "conv_halo_kernel" in kname, kname
sha = hashlib.sha256(open(os.path.join(ROOT, "p2pb_b200", "csrc", "conv_halo.cu"), "rb").read()).hexdigest()
path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
d = json.load(open(path)) if os.path.exists(path) else {}
d["conv_halo_64x64_r32_B64_f16"] = {
    "dram_bytes": int(rd + wr), "source_sha256": sha,
    "capture": f"ncu --set full, {os.path.relpath(rep, ROOT)} (profiles/r02_ncu_conv.md): dram__bytes_read.sum {rd / 1e6:.1f} MB + "
               f"dram__bytes_write.sum {wr / 1e6:.1f} MB"}
json.dump(d, open(path, "w"), indent=1)
print(json.dumps(d, indent=1))
