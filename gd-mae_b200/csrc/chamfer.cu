// Generative-decoder head: ground-truth point grouping (+ pillar-centre normalisation) and the
// weighted chamfer loss with its gradient w.r.t. the predicted points.
//
// Replaces (reference file:line, relative to /root/reference):
//   sst_ops_utils.group_inner_inds + points[group_inds]     pcdet/ops/sst_ops/sst_ops_utils.py:15-27
//   common_utils.get_voxel_centers, gt - centre             pcdet/utils/common_utils.py:130-145, spt_backbone_mae.py:67-72
//   pytorch3d.loss.chamfer_distance(pred, gt, weights=mask) pcdet/models/backbones_3d/spt_backbone_mae.py:88 (third party)
//
// One warp per pillar: the 16 predicted points sit in registers of every lane, each lane owns
// P2/32 ground-truth points.  The forward kernel also emits d(loss)/d(pred) (up to the global
// 1/sum(w) factor), so the backward pass launches nothing.
#include "common.cuh"

struct CenterParams { float r0, r1, r2, v0, v1, v2; int n_cols; };

// gt[m, k, :] = xyz[ k-th point of pillar m (cyclic) ] - centre(m);  centre = (c + 0.5) * voxel + min
__global__ void group_points_kernel(const float* __restrict__ pts, const int* __restrict__ seg_off, const int* __restrict__ seg_pts,
                                    const long long* __restrict__ vcoords, CenterParams p, long long M, int K, float* __restrict__ gt) {
  long long total = M * K;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    long long m = t / K;
    int k = (int)(t % K);
    int s = seg_off[m], cnt = seg_off[m + 1] - s;
    float* o = gt + 3 * t;
    if (cnt == 0) { o[0] = o[1] = o[2] = 0.f; continue; }
    const float* row = pts + (long long)seg_pts[s + (k < cnt ? k : k % cnt)] * p.n_cols;
    float cx = __fadd_rn(__fmul_rn(__fadd_rn((float)vcoords[4 * m + 3], 0.5f), p.v0), p.r0);
    float cy = __fadd_rn(__fmul_rn(__fadd_rn((float)vcoords[4 * m + 2], 0.5f), p.v1), p.r1);
    float cz = __fadd_rn(__fmul_rn(__fadd_rn((float)vcoords[4 * m + 1], 0.5f), p.v2), p.r2);
    o[0] = __fsub_rn(row[1], cx);
    o[1] = __fsub_rn(row[2], cy);
    o[2] = __fsub_rn(row[3], cz);
  }
}

extern "C" int gdmae_group_points_centered(const float* points, int n_cols, const int32_t* seg_offsets, const int32_t* seg_points,
                                           const int64_t* voxel_coords, const float* pc_range, const float* voxel, int64_t M, int K,
                                           float* out_gt, void* stream_) {
  GDMAE_CHECK_ARG(M >= 0 && K > 0 && n_cols >= 4);
  if (M == 0) return GDMAE_OK;
  CenterParams p{pc_range[0], pc_range[1], pc_range[2], voxel[0], voxel[1], voxel[2], n_cols};
  group_points_kernel<<<gdmae_grid(M * K, 256, 32), 256, 0, (cudaStream_t)stream_>>>(points, seg_offsets, seg_points,
                                                                                   (const long long*)voxel_coords, p, M, K, out_gt);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

template <int P1>
__global__ void __launch_bounds__(256) chamfer_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                      const float* __restrict__ w, long long N, int P2,
                                                      float* __restrict__ per_item, float* __restrict__ dpred) {
  int lane = threadIdx.x & 31;
  for (long long n = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; n < N;
       n += ((long long)gridDim.x * blockDim.x) >> 5) {
    float wn = w ? w[n] : 1.f;
    if (wn == 0.f) {  // visible pillar: contributes nothing (weights are the MAE mask)
      if (lane == 0) per_item[n] = 0.f;
      for (int i = lane; i < P1 * 3; i += 32) dpred[n * P1 * 3 + i] = 0.f;
      continue;
    }
    float px[P1], py[P1], pz[P1];
    const float* pr = pred + n * P1 * 3;
#pragma unroll
    for (int i = 0; i < P1; ++i) { px[i] = __ldg(pr + 3 * i); py[i] = __ldg(pr + 3 * i + 1); pz[i] = __ldg(pr + 3 * i + 2); }
    float best_x[P1];   // min over this lane's gt points, per pred point
    int arg_x[P1];
#pragma unroll
    for (int i = 0; i < P1; ++i) { best_x[i] = INFINITY; arg_x[i] = 0; }
    float gx_acc[P1], gy_acc[P1], gz_acc[P1];  // gradient of the gt->pred term, per pred point (lane partial)
#pragma unroll
    for (int i = 0; i < P1; ++i) { gx_acc[i] = 0.f; gy_acc[i] = 0.f; gz_acc[i] = 0.f; }
    float sum_y = 0.f;
    const float* g = gt + n * (long long)P2 * 3;
    for (int j = lane; j < P2; j += 32) {
      float x = __ldg(g + 3 * j), y = __ldg(g + 3 * j + 1), z = __ldg(g + 3 * j + 2);
      float bmin = INFINITY;
      int bi = 0;
#pragma unroll
      for (int i = 0; i < P1; ++i) {
        float dx = px[i] - x, dy = py[i] - y, dz = pz[i] - z;
        float d = dx * dx + dy * dy + dz * dz;
        if (d < best_x[i]) { best_x[i] = d; arg_x[i] = j; }
        if (d < bmin) { bmin = d; bi = i; }
      }
      sum_y += bmin;
#pragma unroll
      for (int i = 0; i < P1; ++i) {
        if (i == bi) { gx_acc[i] += px[i] - x; gy_acc[i] += py[i] - y; gz_acc[i] += pz[i] - z; }
      }
    }
    sum_y = warp_sum(sum_y);
    float sum_x = 0.f;
    float sx = 2.f * wn / (float)P1, sy = 2.f * wn / (float)P2;
#pragma unroll
    for (int i = 0; i < P1; ++i) {
      // warp arg-min over lanes (ties: lowest gt index)
      float v = best_x[i];
      int idx = arg_x[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, v, o);
        int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
      }
      sum_x += v;
      float ax = warp_sum(gx_acc[i]), ay = warp_sum(gy_acc[i]), az = warp_sum(gz_acc[i]);
      if (lane == 0) {
        float x = __ldg(g + 3 * idx), y = __ldg(g + 3 * idx + 1), z = __ldg(g + 3 * idx + 2);
        float* d = dpred + n * P1 * 3 + 3 * i;
        d[0] = sx * (px[i] - x) + sy * ax;
        d[1] = sx * (py[i] - y) + sy * ay;
        d[2] = sx * (pz[i] - z) + sy * az;
      }
    }
    if (lane == 0) per_item[n] = wn * (sum_x / (float)P1 + sum_y / (float)P2);
  }
}

// per_item[n] = w_n * ( mean_i min_j |p_i - g_j|^2 + mean_j min_i |g_j - p_i|^2 )
// dpred[n,i,:] = d per_item[n] / d pred[n,i,:].   loss = sum(per_item) / sum(w)  (done by the caller).
extern "C" int gdmae_chamfer_fwd(const float* pred, const float* gt, const float* weights, int64_t N, int P1, int P2,
                                 float* per_item, float* dpred, void* stream_) {
  GDMAE_CHECK_ARG(N >= 0 && P1 == 16 && P2 >= 1);
  if (N == 0) return GDMAE_OK;
  chamfer_kernel<16><<<gdmae_grid(N * 32, 256, 8), 256, 0, (cudaStream_t)stream_>>>(pred, gt, weights, N, P2, per_item, dpred);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}
