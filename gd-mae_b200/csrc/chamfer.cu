// Generative-decoder head: ground-truth point grouping (+ pillar-centre normalisation) and the
// weighted chamfer loss with its gradient w.r.t. the predicted points.
//
// Replaces (reference file:line, relative to /root/reference):
//   sst_ops_utils.group_inner_inds + points[group_inds]     pcdet/ops/sst_ops/sst_ops_utils.py:15-27
//   common_utils.get_voxel_centers, gt - centre             pcdet/utils/common_utils.py:130-145, spt_backbone_mae.py:67-72
//   pytorch3d.loss.chamfer_distance(pred, gt, weights=mask) pcdet/models/backbones_3d/spt_backbone_mae.py:88 (third party)
//
// One warp per pillar.  Lane l < 16 owns predicted point l, every lane owns P2/32 ground-truth
// points.  The 16 x 64 distance table is never stored: predicted points are broadcast with
// shuffles, the nearest ground-truth point of each prediction is ONE integer warp reduction
// (distance bits with the point index in the low bits, REDUX.MIN), and the gt->pred assignments
// go through 1 KB of shared memory so that the owner lanes accumulate their gradient in a fixed
// order (deterministic, no float atomics).  ~40 registers per thread -> full occupancy; the
// forward kernel also emits d(loss)/d(pred) (up to the global 1/sum(w) factor), so the backward
// pass launches nothing.
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

#define CH_P1 16
#define CH_MAXG 4  // ground-truth points per lane: P2 <= 128

// P2 <= 128 ground-truth points, 16 predicted points per item.
template <int MAXG>     // ground-truth points per lane: P2 <= 32 * MAXG (2 for the 64 points of the GD-MAE configs)
__global__ void __launch_bounds__(256) chamfer_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                      const float* __restrict__ w, long long N, int P2,
                                                      float* __restrict__ per_item, float* __restrict__ dpred) {
  __shared__ float4 sg[8][32 * MAXG];  // per warp: (gx, gy, gz, nearest pred index as float bits) per gt point
  __shared__ float4 sp[8][CH_P1];         // per warp: the pillar's predicted points (one broadcast read per point in pass 1)
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float4* mine = sg[wib];
  const int G = (P2 + 31) >> 5;  // gt points per lane (lane owns j = lane + 32 k)
  for (long long n = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; n < N;
       n += ((long long)gridDim.x * blockDim.x) >> 5) {
    float wn = w ? w[n] : 1.f;
    if (wn == 0.f) {  // visible pillar: contributes nothing (weights are the MAE mask)
      if (lane == 0) per_item[n] = 0.f;
      for (int i = lane; i < CH_P1 * 3; i += 32) dpred[n * CH_P1 * 3 + i] = 0.f;
      continue;
    }
    float px = 0.f, py = 0.f, pz = 0.f;
    if (lane < CH_P1) {
      const float* pr = pred + n * CH_P1 * 3 + 3 * lane;
      px = __ldg(pr); py = __ldg(pr + 1); pz = __ldg(pr + 2);
    }
    __syncwarp();
    if (lane < CH_P1) sp[wib][lane] = make_float4(px, py, pz, 0.f);
    __syncwarp();
    float gx[MAXG], gy[MAXG], gz[MAXG], gbest[MAXG];
    int gidx[MAXG];
    const float* g = gt + n * (long long)P2 * 3;
#pragma unroll
    for (int k = 0; k < MAXG; ++k) {
      int j = lane + 32 * k;
      bool ok = k < G && j < P2;
      gx[k] = ok ? __ldg(g + 3 * j) : 0.f;
      gy[k] = ok ? __ldg(g + 3 * j + 1) : 0.f;
      gz[k] = ok ? __ldg(g + 3 * j + 2) : 0.f;
      gbest[k] = INFINITY;
      gidx[k] = 0;
    }
    // pass 1: every predicted point against every gt point
    unsigned my_key = 0xffffffffu;  // lane i < 16 ends up with the packed arg-min of prediction i
#pragma unroll
    for (int i = 0; i < CH_P1; ++i) {
      const float4 qp = sp[wib][i];
      const float qx = qp.x, qy = qp.y, qz = qp.z;
      unsigned dbits = 0xffffffffu, jmin = 0xffffffffu;  // this lane's nearest gt point to prediction i
#pragma unroll
      for (int k = 0; k < MAXG; ++k) {
        int j = lane + 32 * k;
        if (k < G && j < P2) {
          float dx = qx - gx[k], dy = qy - gy[k], dz = qz - gz[k];
          float d = dx * dx + dy * dy + dz * dz;
          if (d < gbest[k]) { gbest[k] = d; gidx[k] = i; }
          unsigned b = __float_as_uint(d);  // distance >= 0: the bit pattern orders like the value
          if (b < dbits) { dbits = b; jmin = (unsigned)j; }
        }
      }
      // exact warp arg-min in two integer reductions (REDUX): the minimum, then the lowest index attaining it
      unsigned wmin = __reduce_min_sync(0xffffffffu, dbits);
      unsigned wj = __reduce_min_sync(0xffffffffu, dbits == wmin ? jmin : 0xffffffffu);
      if (lane == i) my_key = wj;
    }
    // gt -> pred term and hand-over of the assignments through shared memory
    float sum_y = 0.f;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < MAXG; ++k) {
      int j = lane + 32 * k;
      if (k < G && j < P2) {
        sum_y += gbest[k];
        mine[j] = make_float4(gx[k], gy[k], gz[k], __int_as_float(gidx[k]));
      }
    }
    sum_y = warp_sum(sum_y);
    __syncwarp();
    float sum_x = 0.f;
    {
      // gt -> pred gradient: prediction (lane & 15) collects the gt points assigned to it; the two half-warps scan one half of
      // the gt points each (fixed order inside a half, halves added at the end -> deterministic)
      static_assert(CH_P1 == 16, "half-warp split assumes 16 predicted points");
      const int pi = lane & 15, hf = lane >> 4;
      const float4 pp = sp[wib][pi];
      const int jh = (P2 + 1) >> 1, j0 = hf * jh, j1 = hf ? P2 : jh;
      float ax = 0.f, ay = 0.f, az = 0.f;
      for (int j = j0; j < j1; ++j) {
        const float4 e = mine[j];
        if (__float_as_int(e.w) == pi) { ax += pp.x - e.x; ay += pp.y - e.y; az += pp.z - e.z; }
      }
      ax += __shfl_xor_sync(0xffffffffu, ax, 16);
      ay += __shfl_xor_sync(0xffffffffu, ay, 16);
      az += __shfl_xor_sync(0xffffffffu, az, 16);
      if (lane < CH_P1) {
        // pred -> gt term: exact distance to the arg-min found above (the packed key dropped 7 mantissa bits)
        const float4 gm = mine[my_key & 127u];
        const float dx = px - gm.x, dy = py - gm.y, dz = pz - gm.z;
        sum_x = dx * dx + dy * dy + dz * dz;
        const float sx = 2.f * wn / (float)CH_P1, sy = 2.f * wn / (float)P2;
        float* d = dpred + n * CH_P1 * 3 + 3 * lane;
        d[0] = sx * dx + sy * ax;
        d[1] = sx * dy + sy * ay;
        d[2] = sx * dz + sy * az;
      }
    }
    sum_x = warp_sum(sum_x);
    if (lane == 0) per_item[n] = wn * (sum_x / (float)CH_P1 + sum_y / (float)P2);
    __syncwarp();
  }
}

// per_item[n] = w_n * ( mean_i min_j |p_i - g_j|^2 + mean_j min_i |g_j - p_i|^2 )
// dpred[n,i,:] = d per_item[n] / d pred[n,i,:].   loss = sum(per_item) / sum(w)  (done by the caller).
extern "C" int gdmae_chamfer_fwd(const float* pred, const float* gt, const float* weights, int64_t N, int P1, int P2,
                                 float* per_item, float* dpred, void* stream_) {
  GDMAE_CHECK_ARG(N >= 0 && P1 == CH_P1 && P2 >= 1 && P2 <= 32 * CH_MAXG);
  if (N == 0) return GDMAE_OK;
  if (P2 <= 64) chamfer_kernel<2><<<gdmae_grid(N * 32, 256, 8), 256, 0, (cudaStream_t)stream_>>>(pred, gt, weights, N, P2, per_item, dpred);
  else chamfer_kernel<CH_MAXG><<<gdmae_grid(N * 32, 256, 8), 256, 0, (cudaStream_t)stream_>>>(pred, gt, weights, N, P2, per_item, dpred);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}
