import sys, os, torch, time
sys.path.insert(0, os.getcwd())
from p2pb_b200 import ops
from p2pb_b200._lib import lib
for N, M in ((2000000, 980), (400000, 2000), (149504, 50000), (28672, 10000), (60000, 5000)):
    x = torch.randn(1, 3, N, device="cuda")
    for mode in (1, 3, 0):
        lib().p2pb_fps_set_cluster(mode)
        ops.furthest_point_sampling(x, 8); torch.cuda.synchronize()
        t0 = time.time(); ops.furthest_point_sampling(x, M); torch.cuda.synchronize()
        print(f"N={N} M={M} mode={ {0: 'one CTA', 1: 'default (cluster <= 196608 < grid)', 3: 'grid'}[mode] }: {(time.time()-t0)*1e3:.1f} ms = {(time.time()-t0)*1e6/M:.2f} us/iteration", flush=True)
lib().p2pb_fps_set_cluster(1)
