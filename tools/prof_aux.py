"""Dev tool (run under ncu): one launch each of the kernels AROUND the network -- kNN-2048 patch extraction, Chamfer, EMD pass,
radius query, room patch creation (pad / ragged cluster FPS), patch normalisation, fixed-point reassembly, P2F, big FPS -- at the
sizes of the entry points (denoise_object: 50k-point cloud, 73 patches; denoise_room: 2 M-point room)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from p2pb_b200 import ops
from p2pb_b200 import room as R
from tools.room_sweep import synth_room

dev = "cuda"
g = torch.Generator().manual_seed(0)
cloud = torch.randn(50000, 3, generator=g)
cloud = (cloud / cloud.norm(dim=1, keepdim=True) + 0.02 * torch.randn(50000, 3, generator=g)).to(dev).contiguous()
seeds = cloud[ops.furthest_point_sampling(cloud.t().contiguous()[None], 73)[0].long()].contiguous()
idx = ops.knn_points(seeds, cloud, 2048)                                          # knn_select_kernel
a = torch.randn(64, 2048, 3, device=dev); b = a + 0.01 * torch.randn_like(a)
ops.chamfer_forward(a, b)                                                         # nm_distance_kernel
ops.emd_approx(a[:8].contiguous(), b[:8].contiguous())                            # emd_pass_kernel
big = torch.randn(1, 3, 150000, device=dev)
ops.furthest_point_sampling(big, 2000)                                            # fps_cluster_kernel<16>
pts, _ = synth_room(2_000_000, 0)
room = torch.from_numpy(pts).to(dev)
plan = R.plan_jobs(room, 8192, 4, 0.5, 32)                                        # fps_global + radius_count
pt = R.create_patches(room, plan, 0, 64, 8192, 0.5, 42)                           # radius_fill, room_pad / room_fps
x, c, s = ops.patch_normalize(pt.xyz)                                             # patch_normalize_kernel
acc = R.RoomAccumulator(room.shape[0], dev)
acc.add(x, c, s, pt.idx, pt.cut)                                                  # room_accumulate_kernel
v = torch.randn(20000, 3, device=dev); v = v / v.norm(dim=1, keepdim=True)
f = torch.randint(0, 20000, (40000, 3), device=dev)
ops.point_face_dist(cloud[:20000].contiguous(), v, f, normalize=False)            # p2f_kernel
torch.cuda.synchronize()
print("ok")
