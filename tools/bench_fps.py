"""Dev tool: FPS kernels, cluster (DSMEM) vs one CTA per cloud."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from p2pb_b200 import ops
from p2pb_b200._lib import lib


def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for B, N, M in [(16, 8192, 2048), (64, 2048, 512)]:
    x = torch.randn(B, 3, N, device="cuda")
    ref = None
    for shape in (0, 1, 2):
        lib().p2pb_fps_set_shape(shape)
        t = timeit(lambda: ops.furthest_point_sampling(x, M))
        idx = ops.furthest_point_sampling(x, M)
        ref = idx if ref is None else ref
        print(f"B={B} N={N} M={M} shape {shape}: {t:8.3f} ms ({t / M * 1e3:5.2f} us/iter) same={bool(torch.equal(idx, ref))}")
    lib().p2pb_fps_set_shape(0)
for B, N, M in [(1, 28672, 10000), (1, 149504, 50000)]:
    x = torch.randn(B, 3, N, device="cuda")
    res = []
    for on in (1, 0):
        lib().p2pb_fps_set_cluster(on)
        res.append(timeit(lambda: ops.furthest_point_sampling(x, M), n=1 if M >= 10000 else 3))
    lib().p2pb_fps_set_cluster(1)
    print(f"B={B} N={N} M={M}: cluster {res[0]:9.3f} ms ({res[0] / M * 1e3:6.2f} us/iter) | one CTA per cloud {res[1]:9.3f} ms ({res[1] / M * 1e3:6.2f} us/iter)")
