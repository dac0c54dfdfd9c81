// CenterHead training targets and heat-map loss on the device (BASELINE config 4).
//
// Replaces (reference file:line, relative to /root/reference):
//   CenterHead.assign_target_of_single_head / assign_targets   pcdet/models/dense_heads/center_head.py:105-231
//        - a Python loop over frames and boxes on CPU tensors (gt_boxes.cpu(), draw_gaussian_to_heatmap with a numpy
//          gaussian per box, five .to(device) copies per frame)
//   centernet_utils.gaussian_radius / gaussian2D / draw_gaussian_to_heatmap   pcdet/models/model_utils/centernet_utils.py:9-72
//   loss_utils.neg_loss_cornernet (FocalLossCenterNet)          pcdet/utils/loss_utils.py:273-309, with the clamped sigmoid of
//        CenterHead.sigmoid (center_head.py:233-235)
//
// One CTA per frame assigns its targets: the boxes of the head's classes are compacted in order (the reference appends
// them to a list, center_head.py:196-208), slot k of the first `max_objs` of them receives index / mask / regression
// target / IoU box, and its gaussian is merged into the class heat map with atomicMax on the float bits (all values are
// >= 0, so the integer order is the float order and the result does not depend on the order of the boxes).  The gaussian
// is evaluated in double precision and rounded to float, like numpy's float64 gaussian2D followed by .float().
#include "common.cuh"
#include <math.h>
#include "../../include/gdmae_b200.h"

namespace {

struct AssignArgs {
  const float* gt;          // (B, M, 8) [x, y, z, dx, dy, dz, heading, class id (1-based, 0 = padding)]
  const int* class_map;     // (num_class_total + 1): class id -> 1-based id inside this head, 0 = not in this head
  int M, n_class_total, C, H, W, max_objs, min_radius, stride;
  float x0, y0, vx, vy, overlap;
  float* heat;              // (B, C, H, W), zeroed
  float* target;            // (B, max_objs, 8), zeroed
  float* iou_boxes;         // (B, max_objs, 7), zeroed
  long long* inds;          // (B, max_objs), zeroed
  long long* mask;          // (B, max_objs), zeroed
};

// centernet_utils.gaussian_radius (float32, as torch evaluates it on the CPU tensors)
__device__ inline float gaussian_radius(float h, float w, float ov) {
  const float b1 = h + w, c1 = w * h * (1.f - ov) / (1.f + ov);
  const float r1 = (b1 + sqrtf(b1 * b1 - 4.f * c1)) / 2.f;
  const float b2 = 2.f * (h + w), c2 = (1.f - ov) * w * h;
  const float r2 = (b2 + sqrtf(b2 * b2 - 16.f * c2)) / 2.f;
  const float a3 = 4.f * ov, b3 = -2.f * ov * (h + w), c3 = (ov - 1.f) * w * h;
  const float r3 = (b3 + sqrtf(b3 * b3 - 4.f * a3 * c3)) / 2.f;
  return fminf(fminf(r1, r2), r3);
}

__global__ void __launch_bounds__(256) center_assign_kernel(AssignArgs a) {
  const int b = blockIdx.x;
  const float* gt = a.gt + (long long)b * a.M * 8;
  __shared__ int s_base;
  __shared__ int s_warp[8];
  extern __shared__ int s_slot[];      // slot of every box of the frame (-1: not a target)
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ---- ordered compaction of the boxes that belong to this head
  for (int m0 = 0; m0 < a.M; m0 += 256) {
    const int m = m0 + threadIdx.x;
    int cls = 0;
    if (m < a.M) {
      const int c = (int)gt[m * 8 + 7];
      cls = (c >= 1 && c <= a.n_class_total) ? a.class_map[c] : 0;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, cls > 0);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int before = s_base;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    const int k = before + __popc(bal & ((1u << lane) - 1u));
    if (m < a.M) s_slot[m] = (cls > 0 && k < a.max_objs) ? k : -1;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < 8; ++w) t += s_warp[w];
      s_base += t;
    }
    __syncthreads();
  }
  // ---- per-box targets, then its gaussian (one warp per box)
  for (int m = warp; m < a.M; m += 8) {
    const int k = s_slot[m];
    if (k < 0) continue;
    const float* g = gt + m * 8;
    const float dxm = g[3], dym = g[4];
    if (dxm <= 0.f || dym <= 0.f) continue;                       // center_head.py:141-142 (the slot stays empty)
    float cx = ((g[0] - a.x0) / a.vx) / (float)a.stride, cy = ((g[1] - a.y0) / a.vy) / (float)a.stride;
    cx = fminf(fmaxf(cx, 0.f), (float)a.W - 0.5f);
    cy = fminf(fmaxf(cy, 0.f), (float)a.H - 0.5f);
    const int ix = (int)cx, iy = (int)cy;
    const float dx = (dxm / a.vx) / (float)a.stride, dy = (dym / a.vy) / (float)a.stride;
    int radius = (int)gaussian_radius(dx, dy, a.overlap);
    radius = radius < a.min_radius ? a.min_radius : radius;
    const int cls = a.class_map[(int)g[7]] - 1;
    const long long slot = (long long)b * a.max_objs + k;
    if (lane == 0) {
      a.inds[slot] = (long long)iy * a.W + ix;
      a.mask[slot] = 1;
      float* t = a.target + slot * 8;
      t[0] = cx - (float)ix;
      t[1] = cy - (float)iy;
      t[2] = g[2];
      t[3] = logf(g[3]);
      t[4] = logf(g[4]);
      t[5] = logf(g[5]);
      t[6] = cosf(g[6]);
      t[7] = sinf(g[6]);
      float* ib = a.iou_boxes + slot * 7;
      for (int j = 0; j < 7; ++j) ib[j] = g[j];
    }
    // draw_gaussian_to_heatmap: diameter 2r+1, sigma = diameter / 6, clipped to the map
    const int left = min(ix, radius), right = min(a.W - ix, radius + 1), top = min(iy, radius), bottom = min(a.H - iy, radius + 1);
    const int gw = left + right, gh = top + bottom;
    const double sigma = (double)(2 * radius + 1) / 6.0;
    float* heat = a.heat + ((long long)b * a.C + cls) * a.H * a.W;
    for (int i = lane; i < gw * gh; i += 32) {
      const int oy = i / gw - top, ox = i % gw - left;
      const float v = (float)exp(-(double)(ox * ox + oy * oy) / (2.0 * sigma * sigma));
      atomicMax(reinterpret_cast<int*>(heat + (long long)(iy + oy) * a.W + (ix + ox)), __float_as_int(v));
    }
  }
}

// Focal loss of CenterNet over logits x and targets gt (both (n)):  p = clamp(sigmoid(x), 1e-4, 1 - 1e-4)
//   pos (gt == 1): log(p) (1 - p)^2        neg (gt < 1): log(1 - p) p^2 (1 - gt)^4
// sums[0] += pos terms, sums[1] += neg terms, sums[2] += number of positives (double atomics, one per CTA);
// graw = d(pos + neg) / dx (0 where the clamp is active), so that the backward pass is one scaling.
__global__ void __launch_bounds__(256) center_focal_kernel(const float* __restrict__ x, const float* __restrict__ gt, long long n,
                                                           float* __restrict__ graw, double* __restrict__ sums) {
  double pos = 0.0, neg = 0.0, cnt = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float xv = x[i], g = gt[i];
    const float s = 1.f / (1.f + expf(-xv));
    const float p = fminf(fmaxf(s, 1e-4f), 1.f - 1e-4f);
    const float dpdx = (s >= 1e-4f && s <= 1.f - 1e-4f) ? s * (1.f - s) : 0.f;   // torch.clamp passes the gradient on [min, max]
    float dldp;
    if (g == 1.f) {
      const float q = 1.f - p, lp = logf(p);
      pos += (double)(lp * q * q);
      cnt += 1.0;
      dldp = q * q / p - 2.f * q * lp;
    } else {
      const float w1 = 1.f - g, w = (w1 * w1) * (w1 * w1), l1p = logf(1.f - p);
      neg += (double)(l1p * p * p * w);
      dldp = (2.f * p * l1p - p * p / (1.f - p)) * w;
    }
    graw[i] = dldp * dpdx;
  }
  __shared__ double sh[3][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    pos += __shfl_xor_sync(0xffffffffu, pos, o);
    neg += __shfl_xor_sync(0xffffffffu, neg, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if (lane == 0) { sh[0][warp] = pos; sh[1][warp] = neg; sh[2][warp] = cnt; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[threadIdx.x][w];
    atomicAdd(sums + threadIdx.x, t);
  }
}

}  // namespace

extern "C" int gdmae_center_assign_targets(const float* gt_boxes, int B, int M, const int* class_map, int n_class_total, int C, int H, int W,
                                           int max_objs, int min_radius, int stride, const float* range_xy_voxel_xy, float overlap,
                                           float* heatmap, float* target_boxes, float* iou_boxes, int64_t* inds, int64_t* mask,
                                           void* stream) {
  GDMAE_CHECK_ARG(B >= 0 && M >= 0 && C > 0 && H > 0 && W > 0 && max_objs > 0 && stride > 0 && class_map && range_xy_voxel_xy);
  cudaStream_t st = (cudaStream_t)stream;
  GDMAE_CHECK_CUDA(cudaMemsetAsync(heatmap, 0, (size_t)B * C * H * W * sizeof(float), st));
  GDMAE_CHECK_CUDA(cudaMemsetAsync(target_boxes, 0, (size_t)B * max_objs * 8 * sizeof(float), st));
  GDMAE_CHECK_CUDA(cudaMemsetAsync(iou_boxes, 0, (size_t)B * max_objs * 7 * sizeof(float), st));
  GDMAE_CHECK_CUDA(cudaMemsetAsync(inds, 0, (size_t)B * max_objs * sizeof(int64_t), st));
  GDMAE_CHECK_CUDA(cudaMemsetAsync(mask, 0, (size_t)B * max_objs * sizeof(int64_t), st));
  if (B == 0 || M == 0) return GDMAE_OK;
  GDMAE_CHECK_ARG((size_t)M * sizeof(int) <= 40 * 1024);
  AssignArgs a;
  a.gt = gt_boxes; a.class_map = class_map; a.M = M; a.n_class_total = n_class_total; a.C = C; a.H = H; a.W = W;
  a.max_objs = max_objs; a.min_radius = min_radius; a.stride = stride;
  a.x0 = range_xy_voxel_xy[0]; a.y0 = range_xy_voxel_xy[1]; a.vx = range_xy_voxel_xy[2]; a.vy = range_xy_voxel_xy[3]; a.overlap = overlap;
  a.heat = heatmap; a.target = target_boxes; a.iou_boxes = iou_boxes; a.inds = (long long*)inds; a.mask = (long long*)mask;
  center_assign_kernel<<<B, 256, (size_t)M * sizeof(int), st>>>(a);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

extern "C" int gdmae_center_focal_loss(const float* logits, const float* gt, int64_t n, float* grad_raw, double* sums3, void* stream) {
  GDMAE_CHECK_ARG(n >= 0 && sums3);
  cudaStream_t st = (cudaStream_t)stream;
  GDMAE_CHECK_CUDA(cudaMemsetAsync(sums3, 0, 3 * sizeof(double), st));
  if (n == 0) return GDMAE_OK;
  center_focal_kernel<<<gdmae_grid(n, 256, 8), 256, 0, st>>>(logits, gt, n, grad_raw, sums3);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}
