/*
 * p2pb_b200.h -- C ABI of libp2pb_b200.so: hand-written sm_100a CUDA kernels for the P2P-Bridge denoising hot path.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless named host_*; tensors are dense, fp32 / int32, in the layout stated;
 *   - `stream` is a cudaStream_t (passed as void* so that no CUDA header is needed to bind this file);
 *   - every function returns 0 on success, <0 on error (-1 invalid argument, -2 CUDA error, -3 unsupported) and
 *     never calls exit() (the reference does on launch failure, cuda_utils.cuh:30-40); p2pb_last_error() returns the
 *     message of the last failure on the calling thread;
 *   - no global state besides a cached device-attribute lookup; all scratch buffers are caller-provided.
 * Reference paths are relative to /root/reference/third_party/openpoints/cpp/pointnet2_batch/src/ unless stated.
 */
#ifndef P2PB_B200_H
#define P2PB_B200_H

#ifdef __cplusplus
extern "C" {
#endif

const char* p2pb_last_error(void);
int p2pb_abi_version(void);
int p2pb_device_sm_count(void);
/* kernels launched (or captured into a CUDA graph) through this library since load */
unsigned long long p2pb_launch_count(void);

/* ---- point ops, reference layout (channel-first [B,C,N]) ------------------------------------------------- */

/* replaces furthest_point_sampling_forward (pvcnn_sampling.cpp:45-61, kernel pvcnn_sampling_gpu.cu:92-184) and,
 * when `centers` != NULL, the following gather_features_forward of the coordinates (sampling.py:35-42).
 * coords [B,3,N] -> idx int32 [B,M]; centers [B,3,M] optional; scratch [B,N] fp32 only needed when N > 16384. */
int p2pb_furthest_point_sampling(const float* coords, int B, int N, int M, int* idx, float* centers, float* scratch,
                                 void* stream);

/* replaces gather_features_forward (pvcnn_sampling.cpp:6-24): out[b,c,j] = feat[b,c,idx[b,j]] */
int p2pb_gather_features(const float* feat, const int* idx, float* out, int B, int C, int N, int M, void* stream);

/* replaces grouping_forward (pvcnn_grouping.cpp:6-25): out[b,c,j,k] = feat[b,c,idx[b,j,k]], idx [B,M,U] */
int p2pb_grouping(const float* feat, const int* idx, float* out, int B, int C, int N, int M, int U, void* stream);

/* replaces ball_query_forward (pvcnn_ball_query.cpp:6-31): centers [B,3,M], points [B,3,N] -> idx int32 [B,M,U] */
int p2pb_ball_query(const float* centers, const float* points, int B, int M, int N, float radius, int U, int* out,
                    void* stream);

/* replaces three_nearest_neighbors_interpolate_forward (pvcnn_neighbor_interpolate.cpp:6-41):
 * points [B,3,N], centers [B,3,M], cfeat [B,C,M] -> out [B,C,N], idx int32 [B,3,N], w [B,3,N] */
int p2pb_three_nn(const float* points, const float* centers, int B, int N, int M, int* idx, float* w, void* stream);
int p2pb_three_nn_interpolate(const float* points, const float* centers, const float* cfeat, int B, int C, int N, int M,
                              float* out, int* idx, float* w, void* stream);

/* ---- voxel ops --------------------------------------------------------------------------------------------- */

/* fused Voxelization.forward coordinate prep (/root/reference/models/pvcnn.py:215-231) + grid_stats (vox_gpu.cu:18-36)
 * + CSR build.  coords [B,3,N] -> norm_coords [B,3,N], ind [B,N], order [B,N], start [B,r^3], cnt [B,r^3] (int32) */
int p2pb_voxel_prep(const float* coords, int B, int N, int r, int normalize, float eps, float* norm_coords, int* ind,
                    int* order, int* start, int* cnt, void* stream);

/* replaces avg_voxelize_forward (vox.cpp:17-44): feat [B,C,N], int coords [B,3,N] -> out [B,C,r^3], ind [B,N],
 * cnt [B,r^3]; scratch_order [B,N] int32, scratch_start [B,r^3] int32 */
int p2pb_avg_voxelize(const float* feat, const int* coords, int B, int C, int N, int r, float* out, int* ind, int* cnt,
                      int* scratch_order, int* scratch_start, void* stream);

/* replaces trilinear_devoxelize_forward(is_training=false) (trilinear_devox.cpp:18-59):
 * coords [B,3,N] in [0,r-1], grid [B,C,r^3] -> out [B,C,N] */
int p2pb_trilinear_devoxelize(const float* coords, const float* grid, int B, int C, int N, int r, float* out,
                              void* stream);

/* ---- parity metric ----------------------------------------------------------------------------------------- */

/* replaces one NmDistanceKernel launch of chamfer_cuda_forward (/root/reference/metrics/chamfer3D/chamfer3D.cu:12-146):
 * xyz1 [B,n,3], xyz2 [B,m,3] -> dist [B,n] (squared), idx int32 [B,n]; scratch: B*n 64-bit words */
int p2pb_nm_distance(const float* xyz1, const float* xyz2, int B, int n, int m, float* dist, int* idx,
                     unsigned long long* scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* P2PB_B200_H */
