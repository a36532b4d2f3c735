"""Drop-in for the reference's JIT module ``_pvcnn_backend`` (``third_party/pvcnn/functional/backend.py:9-31``,
exports ``third_party/pvcnn/functional/src/bindings.cpp:10-37``): same kernels as ``pointnet2_batch_cuda`` with
one renamed export, ``furthest_point_sampling`` (used by ``denoise_room.py:21,462,533``)."""
from .pointnet2_batch_cuda import *  # noqa: F401,F403
from .pointnet2_batch_cuda import furthest_point_sampling_forward as furthest_point_sampling  # noqa: F401
