// Pillar feature encoder kernels: point feature build, segment max (+argmax) and its backward.
//
// Replaces (reference file:line, relative to /root/reference):
//   DynVFE point feature build (f_center, xyz+feat, f_cluster)   pcdet/models/backbones_3d/vfe/dyn_vfe.py:86-105
//   torch_scatter.scatter_max(x, inverse, dim=0)                 pcdet/models/backbones_3d/vfe/dyn_vfe.py:109-111
//
// The pillar's points are contiguous in the CSR order produced by gdmae_dynvox, so the max is a
// segmented reduction with one warp per pillar and 128-bit loads of whole 512-byte rows - no
// atomics, no second arg pass (torch_scatter does an atomicMax pass plus an argmax pass).
#include "common.cuh"

struct FeatParams {
  float r0, r1, r2, v0, v1, v2;
  int n_cols;   // 1 + n_feat
};

// x[p] = [ xyz - centre(cell) , points[p,1:] , xyz - mean_xyz[inverse[p]] ]   (no FMA contraction:
// the reference evaluates (c + 0.5) * voxel + min with separate fp32 roundings)
__global__ void vfe_feat_kernel(const float* __restrict__ pts, const long long* __restrict__ coords,
                                const long long* __restrict__ inverse, const float* __restrict__ mean, int mean_stride,
                                long long Np, FeatParams p, float* __restrict__ out) {
  int nf = p.n_cols - 1;
  int C = nf + 6;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < Np; i += (long long)gridDim.x * blockDim.x) {
    const float* row = pts + i * p.n_cols;
    float x = row[1], y = row[2], z = row[3];
    long long cz = coords[4 * i + 1], cy = coords[4 * i + 2], cx = coords[4 * i + 3];
    float* o = out + i * C;
    o[0] = __fsub_rn(x, __fadd_rn(__fmul_rn(__fadd_rn((float)cx, 0.5f), p.v0), p.r0));
    o[1] = __fsub_rn(y, __fadd_rn(__fmul_rn(__fadd_rn((float)cy, 0.5f), p.v1), p.r1));
    o[2] = __fsub_rn(z, __fadd_rn(__fmul_rn(__fadd_rn((float)cz, 0.5f), p.v2), p.r2));
    for (int k = 0; k < nf; ++k) o[3 + k] = row[1 + k];
    const float* mu = mean + inverse[i] * mean_stride;
    o[3 + nf] = __fsub_rn(x, mu[0]);
    o[4 + nf] = __fsub_rn(y, mu[1]);
    o[5 + nf] = __fsub_rn(z, mu[2]);
  }
}

extern "C" int gdmae_vfe_point_features(const float* points, const int64_t* point_coords, const int64_t* inverse,
                                        const float* mean, int mean_stride, int64_t Np, int n_cols, const float* pc_range,
                                        const float* voxel, float* out, void* stream_) {
  GDMAE_CHECK_ARG(Np >= 0 && n_cols >= 4 && mean_stride >= 3);
  if (Np == 0) return GDMAE_OK;
  FeatParams p;
  p.r0 = pc_range[0]; p.r1 = pc_range[1]; p.r2 = pc_range[2];
  p.v0 = voxel[0]; p.v1 = voxel[1]; p.v2 = voxel[2];
  p.n_cols = n_cols;
  vfe_feat_kernel<<<gdmae_grid(Np, 256), 256, 0, (cudaStream_t)stream_>>>(points, (const long long*)point_coords,
                                                                        (const long long*)inverse, mean, mean_stride, Np, p, out);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// One warp per pillar; lane handles float4 column groups.  Ties keep the lowest point index.
__global__ void __launch_bounds__(256) segment_max_fwd_kernel(const float* __restrict__ src, int C,
                                                              const int* __restrict__ seg_off, const int* __restrict__ seg_pts,
                                                              int M, float* __restrict__ out, int* __restrict__ arg) {
  int lane = threadIdx.x & 31;
  int C4 = C >> 2;
  for (int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < M; m += (gridDim.x * blockDim.x) >> 5) {
    int s = seg_off[m], e = seg_off[m + 1];
    for (int c4 = lane; c4 < C4; c4 += 32) {
      float4 best = make_float4(0.f, 0.f, 0.f, 0.f);
      int4 bi = make_int4(-1, -1, -1, -1);
      for (int k = s; k < e; ++k) {
        int pnt = seg_pts[k];
        float4 v = __ldg(reinterpret_cast<const float4*>(src + (long long)pnt * C) + c4);
        if (k == s || v.x > best.x) { best.x = v.x; bi.x = pnt; }
        if (k == s || v.y > best.y) { best.y = v.y; bi.y = pnt; }
        if (k == s || v.z > best.z) { best.z = v.z; bi.z = pnt; }
        if (k == s || v.w > best.w) { best.w = v.w; bi.w = pnt; }
      }
      reinterpret_cast<float4*>(out + (long long)m * C)[c4] = best;
      if (arg) reinterpret_cast<int4*>(arg + (long long)m * C)[c4] = bi;
    }
  }
}

extern "C" int gdmae_segment_max_fwd(const float* src, int C, const int32_t* seg_offsets, const int32_t* seg_points, int64_t M,
                                     float* out, int32_t* out_argmax, void* stream_) {
  GDMAE_CHECK_ARG(M >= 0 && C > 0 && (C % 4) == 0);
  if (M == 0) return GDMAE_OK;
  segment_max_fwd_kernel<<<gdmae_grid(M * 32, 256), 256, 0, (cudaStream_t)stream_>>>(src, C, seg_offsets, seg_points, (int)M, out,
                                                                                   out_argmax);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// dsrc[p, c] = dout[m, c] if argmax[m, c] == p else 0; every point row is written exactly once.
__global__ void __launch_bounds__(256) segment_max_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ arg, int C,
                                                              const int* __restrict__ seg_off, const int* __restrict__ seg_pts,
                                                              int M, float* __restrict__ dsrc) {
  int lane = threadIdx.x & 31;
  int C4 = C >> 2;
  for (int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < M; m += (gridDim.x * blockDim.x) >> 5) {
    int s = seg_off[m], e = seg_off[m + 1];
    for (int c4 = lane; c4 < C4; c4 += 32) {
      float4 g = __ldg(reinterpret_cast<const float4*>(dout + (long long)m * C) + c4);
      int4 a = __ldg(reinterpret_cast<const int4*>(arg + (long long)m * C) + c4);
      for (int k = s; k < e; ++k) {
        int pnt = seg_pts[k];
        float4 v = make_float4(a.x == pnt ? g.x : 0.f, a.y == pnt ? g.y : 0.f, a.z == pnt ? g.z : 0.f, a.w == pnt ? g.w : 0.f);
        reinterpret_cast<float4*>(dsrc + (long long)pnt * C)[c4] = v;
      }
    }
  }
}

extern "C" int gdmae_segment_max_bwd(const float* dout, const int32_t* argmax, int C, const int32_t* seg_offsets,
                                     const int32_t* seg_points, int64_t M, float* dsrc, void* stream_) {
  GDMAE_CHECK_ARG(M >= 0 && C > 0 && (C % 4) == 0);
  if (M == 0) return GDMAE_OK;
  segment_max_bwd_kernel<<<gdmae_grid(M * 32, 256), 256, 0, (cudaStream_t)stream_>>>(dout, argmax, C, seg_offsets, seg_points, (int)M,
                                                                                   dsrc);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}
