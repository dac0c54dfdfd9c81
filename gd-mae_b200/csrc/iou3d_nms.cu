// Rotated BEV overlap / IoU and NMS on sm_100a.
//
// Replaces (reference file:line, relative to /root/reference):
//   box_overlap / iou_bev device functions      pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu:104-236
//   boxes_overlap_kernel / boxes_iou_bev_kernel  .../iou3d_nms_kernel.cu:238-266, launchers :372-394
//   nms_kernel / nms_normal_kernel               .../iou3d_nms_kernel.cu:267-366
//   nms_gpu / nms_normal_gpu host loops          pcdet/ops/iou3d_nms/src/iou3d_nms.cpp:88-187
//   boxes_iou_bev_cpu                            pcdet/ops/iou3d_nms/src/iou3d_cpu.cpp:222-252
//
// The geometry is the reference's: the intersection polygon of two rotated rectangles is the set of edge-edge
// intersection points plus the corners of either box inside the other (with the reference's 1e-2 margin), ordered by
// angle around their centroid and summed as a triangle fan.  It is written once as __host__ __device__ code, so the
// CPU entry point (the reference ships one, iou3d_nms_utils.py:12-28) and the kernels share it.
//
// What differs from the reference: no cudaMalloc / cudaMemcpy / host loop inside the NMS call - the N x ceil(N/64)
// suppression matrix stays on the device (caller-owned workspace) and the sequential sweep runs as one CTA: a 64-box
// block is resolved by one thread from the diagonal words held in shared memory, then all threads OR the kept rows into
// the running removal mask of the later blocks.  Only the count is read back by the caller (the reference returns it).
#include "common.cuh"
#include <math.h>
#include "../../include/gdmae_b200.h"

namespace {

constexpr int NMS_BLOCK = 64;   // boxes per suppression word (unsigned long long)
constexpr float IOU_EPS = 1e-8f;

struct P2 {
  float x, y;
};

__host__ __device__ inline P2 mk(float x, float y) {
  P2 p;
  p.x = x;
  p.y = y;
  return p;
}
__host__ __device__ inline float cross3(const P2& a, const P2& b, const P2& o) { return (a.x - o.x) * (b.y - o.y) - (b.x - o.x) * (a.y - o.y); }

// intersection point of segments p0p1 and q0q1 (proper crossings only), reference iou3d_nms_kernel.cu:58-88
__host__ __device__ inline bool seg_intersection(const P2& p1, const P2& p0, const P2& q1, const P2& q0, P2& out) {
  // bounding boxes of the two segments must overlap
  if (!(fminf(p0.x, p1.x) <= fmaxf(q0.x, q1.x) && fminf(q0.x, q1.x) <= fmaxf(p0.x, p1.x) && fminf(p0.y, p1.y) <= fmaxf(q0.y, q1.y) &&
        fminf(q0.y, q1.y) <= fmaxf(p0.y, p1.y)))
    return false;
  const float s1 = cross3(q0, p1, p0), s2 = cross3(p1, q1, p0), s3 = cross3(p0, q1, q0), s4 = cross3(q1, p1, q0);
  if (!(s1 * s2 > 0.f && s3 * s4 > 0.f)) return false;
  const float s5 = cross3(q1, p1, p0);
  if (fabsf(s5 - s1) > IOU_EPS) {
    out.x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
    out.y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
  } else {
    const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
    const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
    const float D = a0 * b1 - a1 * b0;
    out.x = (b0 * c1 - b1 * c0) / D;
    out.y = (a1 * c0 - a0 * c1) / D;
  }
  return true;
}

// point inside the rotated rectangle, with the reference's 1e-2 margin (iou3d_nms_kernel.cu:46-56)
__host__ __device__ inline bool in_box(const float* box, const P2& p) {
  const float c = cosf(-box[6]), s = sinf(-box[6]);
  const float rx = (p.x - box[0]) * c + (p.y - box[1]) * (-s);
  const float ry = (p.x - box[0]) * s + (p.y - box[1]) * c;
  return fabsf(rx) < box[3] / 2 + 1e-2f && fabsf(ry) < box[4] / 2 + 1e-2f;
}

__host__ __device__ inline void box_corners(const float* box, P2 (&c)[5]) {
  const float hx = box[3] / 2, hy = box[4] / 2;
  const float ca = cosf(box[6]), sa = sinf(box[6]);
  const float lx[4] = {-hx, hx, hx, -hx}, ly[4] = {-hy, -hy, hy, hy};
  for (int k = 0; k < 4; ++k) {
    // the reference forms the axis-aligned corner first and rotates it around the centre (iou3d_nms_kernel.cu:90-94)
    const float px = box[0] + lx[k], py = box[1] + ly[k];
    c[k] = mk((px - box[0]) * ca + (py - box[1]) * (-sa) + box[0], (px - box[0]) * sa + (py - box[1]) * ca + box[1]);
  }
  c[4] = c[0];
}

// area of the intersection of two rotated BEV rectangles [x, y, z, dx, dy, dz, heading]
__host__ __device__ inline float bev_overlap(const float* a, const float* b) {
  P2 ca[5], cb[5];
  box_corners(a, ca);
  box_corners(b, cb);
  P2 pts[16];
  int n = 0;
  float cx = 0.f, cy = 0.f;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      if (seg_intersection(ca[i + 1], ca[i], cb[j + 1], cb[j], pts[n])) {
        cx += pts[n].x;
        cy += pts[n].y;
        ++n;
      }
  for (int k = 0; k < 4; ++k) {
    if (in_box(a, cb[k])) {
      cx += cb[k].x;
      cy += cb[k].y;
      pts[n++] = cb[k];
    }
    if (in_box(b, ca[k])) {
      cx += ca[k].x;
      cy += ca[k].y;
      pts[n++] = ca[k];
    }
  }
  cx /= n;   // n == 0: the loops below do not run (the reference divides by zero here as well)
  cy /= n;
  // bubble sort by angle around the centroid (descending comparison, iou3d_nms_kernel.cu:96-98,198-208)
  for (int j = 0; j < n - 1; ++j)
    for (int i = 0; i < n - j - 1; ++i)
      if (atan2f(pts[i].y - cy, pts[i].x - cx) > atan2f(pts[i + 1].y - cy, pts[i + 1].x - cx)) {
        const P2 t = pts[i];
        pts[i] = pts[i + 1];
        pts[i + 1] = t;
      }
  float area = 0.f;
  for (int k = 0; k < n - 1; ++k) {
    const float ax = pts[k].x - pts[0].x, ay = pts[k].y - pts[0].y, bx = pts[k + 1].x - pts[0].x, by = pts[k + 1].y - pts[0].y;
    area += ax * by - ay * bx;
  }
  return fabsf(area) / 2.0f;
}

__host__ __device__ inline float bev_iou(const float* a, const float* b) {
  const float sa = a[3] * a[4], sb = b[3] * b[4];
  const float so = bev_overlap(a, b);
  return so / fmaxf(sa + sb - so, IOU_EPS);
}

// axis-aligned IoU (headings ignored), iou3d_nms_kernel.cu:316-327
__host__ __device__ inline float normal_iou(const float* a, const float* b) {
  const float left = fmaxf(a[0] - a[3] / 2, b[0] - b[3] / 2), right = fminf(a[0] + a[3] / 2, b[0] + b[3] / 2);
  const float top = fmaxf(a[1] - a[4] / 2, b[1] - b[4] / 2), bottom = fminf(a[1] + a[4] / 2, b[1] + b[4] / 2);
  const float inter = fmaxf(right - left, 0.f) * fmaxf(bottom - top, 0.f);
  return inter / fmaxf(a[3] * a[4] + b[3] * b[4] - inter, IOU_EPS);
}

// one thread per (a, b) pair; b varies fastest so that a warp writes 32 consecutive outputs and reads one a box
template <bool IOU>
__global__ void __launch_bounds__(256) pair_kernel(int na, const float* __restrict__ boxes_a, int nb, const float* __restrict__ boxes_b,
                                                   float* __restrict__ out) {
  const long long total = (long long)na * nb;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ia = (int)(i / nb), ib = (int)(i % nb);
    float a[7], b[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      a[k] = __ldg(boxes_a + (long long)ia * 7 + k);
      b[k] = __ldg(boxes_b + (long long)ib * 7 + k);
    }
    out[i] = IOU ? bev_iou(a, b) : bev_overlap(a, b);
  }
}

// suppression words: mask[i, cb] bit j = IoU(box i, box 64 cb + j) > thresh, only for j > i (boxes are sorted by score)
template <bool ROTATED>
__global__ void __launch_bounds__(NMS_BLOCK) nms_mask_kernel(int n, float thresh, const float* __restrict__ boxes,
                                                             unsigned long long* __restrict__ mask) {
  const int rb = blockIdx.y, cb = blockIdx.x;
  const int rows = min(n - rb * NMS_BLOCK, NMS_BLOCK), cols = min(n - cb * NMS_BLOCK, NMS_BLOCK);
  __shared__ float sb[NMS_BLOCK * 7];
  for (int k = threadIdx.x; k < cols * 7; k += NMS_BLOCK) sb[k] = boxes[(long long)cb * NMS_BLOCK * 7 + k];
  __syncthreads();
  if ((int)threadIdx.x < rows) {
    const int i = rb * NMS_BLOCK + threadIdx.x;
    float a[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) a[k] = boxes[(long long)i * 7 + k];
    unsigned long long t = 0;
    const int start = rb == cb ? threadIdx.x + 1 : 0;
    for (int j = start; j < cols; ++j) {
      const float v = ROTATED ? bev_iou(a, sb + j * 7) : normal_iou(a, sb + j * 7);
      if (v > thresh) t |= 1ull << j;
    }
    mask[(long long)i * gridDim.x + cb] = t;
  }
}

// sequential sweep over the sorted boxes (iou3d_nms.cpp:113-128) on the device, one CTA.
__global__ void __launch_bounds__(256) nms_sweep_kernel(int n, int col_blocks, const unsigned long long* __restrict__ mask,
                                                        long long* __restrict__ keep, int* __restrict__ num_out) {
  extern __shared__ unsigned long long sh[];
  unsigned long long* remv = sh;                 // [col_blocks]
  __shared__ unsigned long long diag[NMS_BLOCK];
  __shared__ unsigned long long kept_bits;
  __shared__ int n_keep;
  for (int j = threadIdx.x; j < col_blocks; j += blockDim.x) remv[j] = 0ull;
  if (threadIdx.x == 0) n_keep = 0;
  __syncthreads();
  for (int b = 0; b < col_blocks; ++b) {
    const int rows = min(n - b * NMS_BLOCK, NMS_BLOCK);
    if ((int)threadIdx.x < rows) diag[threadIdx.x] = mask[(long long)(b * NMS_BLOCK + threadIdx.x) * col_blocks + b];
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long word = remv[b], kept = 0ull;
      int k = n_keep;
      for (int i = 0; i < rows; ++i)
        if (!((word >> i) & 1ull)) {
          keep[k++] = b * NMS_BLOCK + i;
          kept |= 1ull << i;
          word |= diag[i];
        }
      n_keep = k;
      kept_bits = kept;
    }
    __syncthreads();
    const unsigned long long kept = kept_bits;
    // later blocks: OR the rows of the boxes kept in this block into the running removal mask
    for (int j = b + 1 + threadIdx.x; j < col_blocks; j += blockDim.x) {
      unsigned long long acc = remv[j];
      unsigned long long m = kept;
      while (m) {
        const int i = __ffsll((long long)m) - 1;
        m &= m - 1;
        acc |= mask[(long long)(b * NMS_BLOCK + i) * col_blocks + j];
      }
      remv[j] = acc;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *num_out = n_keep;
}

int check_boxes(const float* a, int na) {
  GDMAE_CHECK_ARG(na >= 0 && (na == 0 || a != nullptr));
  return GDMAE_OK;
}

template <bool ROTATED>
int nms_impl(const float* boxes, int n, float thresh, void* workspace, size_t ws_bytes, int64_t* keep, int* num_out, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  GDMAE_CHECK_ARG(n >= 0 && keep && num_out);
  if (n == 0) {
    GDMAE_CHECK_CUDA(cudaMemsetAsync(num_out, 0, sizeof(int), st));
    return GDMAE_OK;
  }
  const int cbs = gdmae_div_up(n, NMS_BLOCK);
  if (ws_bytes < gdmae_nms_workspace_bytes(n)) {
    gdmae_set_error("nms: workspace too small (gdmae_nms_workspace_bytes)");
    return GDMAE_ERR_WORKSPACE;
  }
  GDMAE_CHECK_ARG((size_t)cbs * 8 <= 160 * 1024);          // removal mask in shared memory: up to 1.3 M boxes
  unsigned long long* mask = (unsigned long long*)workspace;
  nms_mask_kernel<ROTATED><<<dim3(cbs, cbs), NMS_BLOCK, 0, st>>>(n, thresh, boxes, mask);
  GDMAE_LAUNCH_CHECK();
  const size_t sh = (size_t)cbs * 8;
  if (sh > 48 * 1024) GDMAE_CHECK_CUDA(cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
  nms_sweep_kernel<<<1, 256, sh, st>>>(n, cbs, mask, (long long*)keep, num_out);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

}  // namespace

extern "C" int gdmae_boxes_overlap_bev(const float* boxes_a, int na, const float* boxes_b, int nb, float* ans_overlap, void* stream) {
  if (check_boxes(boxes_a, na) || check_boxes(boxes_b, nb)) return GDMAE_ERR_ARG;
  if (na == 0 || nb == 0) return GDMAE_OK;
  pair_kernel<false><<<gdmae_grid((long long)na * nb, 256, 8), 256, 0, (cudaStream_t)stream>>>(na, boxes_a, nb, boxes_b, ans_overlap);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

extern "C" int gdmae_boxes_iou_bev(const float* boxes_a, int na, const float* boxes_b, int nb, float* ans_iou, void* stream) {
  if (check_boxes(boxes_a, na) || check_boxes(boxes_b, nb)) return GDMAE_ERR_ARG;
  if (na == 0 || nb == 0) return GDMAE_OK;
  pair_kernel<true><<<gdmae_grid((long long)na * nb, 256, 8), 256, 0, (cudaStream_t)stream>>>(na, boxes_a, nb, boxes_b, ans_iou);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

extern "C" size_t gdmae_nms_workspace_bytes(int n) {
  const size_t cbs = (size_t)((n + NMS_BLOCK - 1) / NMS_BLOCK);
  return gdmae_align((size_t)(n > 0 ? n : 1) * cbs * sizeof(unsigned long long));
}

extern "C" int gdmae_nms_bev(const float* boxes, int n, float thresh, void* workspace, size_t ws_bytes, int64_t* keep, int* num_out,
                             void* stream) {
  return nms_impl<true>(boxes, n, thresh, workspace, ws_bytes, keep, num_out, stream);
}

extern "C" int gdmae_nms_normal(const float* boxes, int n, float thresh, void* workspace, size_t ws_bytes, int64_t* keep, int* num_out,
                                void* stream) {
  return nms_impl<false>(boxes, n, thresh, workspace, ws_bytes, keep, num_out, stream);
}

// host entry point: the reference ships a CPU function of its own (boxes_iou_bev_cpu), this is its counterpart - HOST
// pointers, same geometry code as the kernels
extern "C" int gdmae_boxes_iou_bev_cpu(const float* boxes_a, int na, const float* boxes_b, int nb, float* ans_iou) {
  if (check_boxes(boxes_a, na) || check_boxes(boxes_b, nb)) return GDMAE_ERR_ARG;
  for (int i = 0; i < na; ++i)
    for (int j = 0; j < nb; ++j) ans_iou[(long long)i * nb + j] = bev_iou(boxes_a + (long long)i * 7, boxes_b + (long long)j * 7);
  return GDMAE_OK;
}
