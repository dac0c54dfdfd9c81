// (Modulated) deformable convolution: sampling kernels on sm_100a (SURVEY.md 8f rank 4 - the pcdet/ops/dcn API surface;
// no GD-MAE config executes it, `DLASeg` image backbones do).
//
// Replaces (reference file:line, relative to /root/reference):
//   deformable_im2col / modulated_deformable_im2col            pcdet/ops/dcn/src/deform_conv_cuda_kernel.cu:84-262, 570-695
//   deformable_col2im / modulated_deformable_col2im            .../deform_conv_cuda_kernel.cu:264-352, 697-780
//   deformable_col2im_coord / modulated_..._col2im_coord       .../deform_conv_cuda_kernel.cu:354-470, 782-866
// The five pybind entry points of deform_conv_cuda.cpp:687-701 are rebuilt on these three launchers plus library GEMMs in
// gd-mae_b200/pcdet/ops/dcn/deform_conv.py.
//
// Semantics kept: a tap samples its input plane at (h_in + i*dil + dh, w_in + j*dil + dw) bilinearly; a sample whose centre lies
// outside (-1, H) x (-1, W) is zero, corners outside the plane contribute zero; modulation multiplies the sample by the mask.
// Offsets: channel g*2*kh*kw + 2*(i*kw + j) (+1) = dh (dw) of deformable group g; masks: channel g*kh*kw + i*kw + j.
// Layout here: columns (B, C*kh*kw, Ho*Wo) per image (the reference interleaves `im2col_step` images; the GEMMs are batched
// instead).  One thread per (b, c, ho, wo); consecutive threads walk wo, so offsets / masks / columns are coalesced.
#include "common.cuh"
#include "../../include/gdmae_b200.h"

namespace {

struct DcnGeom {
  int B, C, H, W, kh, kw, pad_h, pad_w, stride_h, stride_w, dil_h, dil_w, dg, Ho, Wo;
};

__device__ __forceinline__ float bilinear(const float* __restrict__ plane, int H, int W, float h, float w) {
  const int h0 = (int)floorf(h), w0 = (int)floorf(w), h1 = h0 + 1, w1 = w0 + 1;
  const float lh = h - h0, lw = w - w0, hh = 1.f - lh, hw = 1.f - lw;
  const float v1 = (h0 >= 0 && w0 >= 0) ? __ldg(plane + h0 * W + w0) : 0.f;
  const float v2 = (h0 >= 0 && w1 <= W - 1) ? __ldg(plane + h0 * W + w1) : 0.f;
  const float v3 = (h1 <= H - 1 && w0 >= 0) ? __ldg(plane + h1 * W + w0) : 0.f;
  const float v4 = (h1 <= H - 1 && w1 <= W - 1) ? __ldg(plane + h1 * W + w1) : 0.f;
  return hh * hw * v1 + hh * lw * v2 + lh * hw * v3 + lh * lw * v4;
}

__global__ void __launch_bounds__(256) dcn_im2col_kernel(DcnGeom g, const float* __restrict__ im, const float* __restrict__ offset,
                                                         const float* __restrict__ mask, float* __restrict__ col) {
  const long long n = (long long)g.B * g.C * g.Ho * g.Wo;
  const int cpg = g.C / g.dg, HW = g.Ho * g.Wo, K2 = g.kh * g.kw;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int wo = (int)(t % g.Wo), ho = (int)((t / g.Wo) % g.Ho), c = (int)((t / HW) % g.C), b = (int)(t / ((long long)HW * g.C));
    const int grp = c / cpg;
    const float* plane = im + ((long long)b * g.C + c) * g.H * g.W;
    const float* off = offset + ((long long)b * g.dg + grp) * 2 * K2 * HW + ho * g.Wo + wo;
    const float* msk = mask ? mask + ((long long)b * g.dg + grp) * K2 * HW + ho * g.Wo + wo : nullptr;
    float* out = col + ((long long)b * g.C + c) * K2 * HW + ho * g.Wo + wo;
    const int h_in = ho * g.stride_h - g.pad_h, w_in = wo * g.stride_w - g.pad_w;
    for (int i = 0; i < g.kh; ++i)
      for (int j = 0; j < g.kw; ++j) {
        const int tap = i * g.kw + j;
        const float h = h_in + i * g.dil_h + __ldg(off + (long long)(2 * tap) * HW), w = w_in + j * g.dil_w + __ldg(off + (long long)(2 * tap + 1) * HW);
        float v = 0.f;
        if (h > -1.f && w > -1.f && h < g.H && w < g.W) v = bilinear(plane, g.H, g.W, h, w);
        if (msk) v *= __ldg(msk + (long long)tap * HW);
        out[(long long)tap * HW] = v;
      }
  }
}

// gradient w.r.t. the input: every column entry scatters to its (at most) four corners (atomics: taps overlap)
__global__ void __launch_bounds__(256) dcn_col2im_kernel(DcnGeom g, const float* __restrict__ dcol, const float* __restrict__ offset,
                                                         const float* __restrict__ mask, float* __restrict__ dim) {
  const int cpg = g.C / g.dg, HW = g.Ho * g.Wo, K2 = g.kh * g.kw;
  const long long n = (long long)g.B * g.C * K2 * HW;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(t % HW), tap = (int)((t / HW) % K2), c = (int)((t / ((long long)HW * K2)) % g.C), b = (int)(t / ((long long)HW * K2 * g.C));
    const int wo = p % g.Wo, ho = p / g.Wo, i = tap / g.kw, j = tap % g.kw, grp = c / cpg;
    const float* off = offset + ((long long)b * g.dg + grp) * 2 * K2 * HW + p;
    const float h = ho * g.stride_h - g.pad_h + i * g.dil_h + __ldg(off + (long long)(2 * tap) * HW);
    const float w = wo * g.stride_w - g.pad_w + j * g.dil_w + __ldg(off + (long long)(2 * tap + 1) * HW);
    if (!(h > -1.f && w > -1.f && h < g.H && w < g.W)) continue;
    float gv = dcol[t];
    if (mask) gv *= __ldg(mask + ((long long)b * g.dg + grp) * K2 * HW + (long long)tap * HW + p);
    const int h0 = (int)floorf(h), w0 = (int)floorf(w);
    const float lh = h - h0, lw = w - w0;
    float* plane = dim + ((long long)b * g.C + c) * g.H * g.W;
    if (h0 >= 0 && w0 >= 0) atomicAdd(plane + h0 * g.W + w0, (1.f - lh) * (1.f - lw) * gv);
    if (h0 >= 0 && w0 + 1 <= g.W - 1) atomicAdd(plane + h0 * g.W + w0 + 1, (1.f - lh) * lw * gv);
    if (h0 + 1 <= g.H - 1 && w0 >= 0) atomicAdd(plane + (h0 + 1) * g.W + w0, lh * (1.f - lw) * gv);
    if (h0 + 1 <= g.H - 1 && w0 + 1 <= g.W - 1) atomicAdd(plane + (h0 + 1) * g.W + w0 + 1, lh * lw * gv);
  }
}

// gradient w.r.t. offsets (and masks): one thread per (b, group, tap, ho, wo), summing over the group's channels - no atomics
__global__ void __launch_bounds__(256) dcn_col2im_coord_kernel(DcnGeom g, const float* __restrict__ dcol, const float* __restrict__ im,
                                                               const float* __restrict__ offset, const float* __restrict__ mask,
                                                               float* __restrict__ doffset, float* __restrict__ dmask) {
  const int cpg = g.C / g.dg, HW = g.Ho * g.Wo, K2 = g.kh * g.kw;
  const long long n = (long long)g.B * g.dg * K2 * HW;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(t % HW), tap = (int)((t / HW) % K2), grp = (int)((t / ((long long)HW * K2)) % g.dg), b = (int)(t / ((long long)HW * K2 * g.dg));
    const int wo = p % g.Wo, ho = p / g.Wo, i = tap / g.kw, j = tap % g.kw;
    const long long obase = ((long long)b * g.dg + grp) * 2 * K2 * HW + p;
    const float h = ho * g.stride_h - g.pad_h + i * g.dil_h + __ldg(offset + obase + (long long)(2 * tap) * HW);
    const float w = wo * g.stride_w - g.pad_w + j * g.dil_w + __ldg(offset + obase + (long long)(2 * tap + 1) * HW);
    const bool inside = h > -1.f && w > -1.f && h < g.H && w < g.W;
    const float m = mask ? __ldg(mask + ((long long)b * g.dg + grp) * K2 * HW + (long long)tap * HW + p) : 1.f;
    float gh = 0.f, gw = 0.f, gm = 0.f;
    if (inside) {
      const int h0 = (int)floorf(h), w0 = (int)floorf(w), h1 = h0 + 1, w1 = w0 + 1;
      const float lh = h - h0, lw = w - w0;
      for (int cc = 0; cc < cpg; ++cc) {
        const int c = grp * cpg + cc;
        const float* plane = im + ((long long)b * g.C + c) * g.H * g.W;
        const float v1 = (h0 >= 0 && w0 >= 0) ? __ldg(plane + h0 * g.W + w0) : 0.f;
        const float v2 = (h0 >= 0 && w1 <= g.W - 1) ? __ldg(plane + h0 * g.W + w1) : 0.f;
        const float v3 = (h1 <= g.H - 1 && w0 >= 0) ? __ldg(plane + h1 * g.W + w0) : 0.f;
        const float v4 = (h1 <= g.H - 1 && w1 <= g.W - 1) ? __ldg(plane + h1 * g.W + w1) : 0.f;
        const float gc = dcol[(((long long)b * g.C + c) * K2 + tap) * HW + p];
        // d sample / dh and / dw of the bilinear form
        gh += gc * ((1.f - lw) * (v3 - v1) + lw * (v4 - v2));
        gw += gc * ((1.f - lh) * (v2 - v1) + lh * (v4 - v3));
        gm += gc * ((1.f - lh) * (1.f - lw) * v1 + (1.f - lh) * lw * v2 + lh * (1.f - lw) * v3 + lh * lw * v4);
      }
    }
    doffset[obase + (long long)(2 * tap) * HW] = gh * m;
    doffset[obase + (long long)(2 * tap + 1) * HW] = gw * m;
    if (dmask) dmask[((long long)b * g.dg + grp) * K2 * HW + (long long)tap * HW + p] = gm;
  }
}

int fill_geom(DcnGeom& g, const int* geom) {
  GDMAE_CHECK_ARG(geom != nullptr);
  g.B = geom[0]; g.C = geom[1]; g.H = geom[2]; g.W = geom[3]; g.kh = geom[4]; g.kw = geom[5]; g.pad_h = geom[6]; g.pad_w = geom[7];
  g.stride_h = geom[8]; g.stride_w = geom[9]; g.dil_h = geom[10]; g.dil_w = geom[11]; g.dg = geom[12];
  GDMAE_CHECK_ARG(g.B >= 0 && g.C > 0 && g.H > 0 && g.W > 0 && g.kh > 0 && g.kw > 0 && g.stride_h > 0 && g.stride_w > 0 && g.dil_h > 0 &&
                  g.dil_w > 0 && g.dg > 0 && g.C % g.dg == 0);
  g.Ho = (g.H + 2 * g.pad_h - (g.dil_h * (g.kh - 1) + 1)) / g.stride_h + 1;
  g.Wo = (g.W + 2 * g.pad_w - (g.dil_w * (g.kw - 1) + 1)) / g.stride_w + 1;
  GDMAE_CHECK_ARG(g.Ho > 0 && g.Wo > 0);
  return GDMAE_OK;
}

}  // namespace

// geom: HOST int[13] = {B, C, H, W, kh, kw, pad_h, pad_w, stride_h, stride_w, dil_h, dil_w, deformable_groups}
extern "C" int gdmae_deform_im2col(const float* input, const float* offset, const float* mask, const int* geom, float* columns, void* stream) {
  DcnGeom g;
  if (int rc = fill_geom(g, geom)) return rc;
  if (g.B == 0) return GDMAE_OK;
  dcn_im2col_kernel<<<gdmae_grid((long long)g.B * g.C * g.Ho * g.Wo, 256, 8), 256, 0, (cudaStream_t)stream>>>(g, input, offset, mask, columns);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

extern "C" int gdmae_deform_col2im(const float* grad_columns, const float* offset, const float* mask, const int* geom, float* grad_input,
                                   void* stream) {
  DcnGeom g;
  if (int rc = fill_geom(g, geom)) return rc;
  if (g.B == 0) return GDMAE_OK;
  GDMAE_CHECK_CUDA(cudaMemsetAsync(grad_input, 0, (size_t)g.B * g.C * g.H * g.W * sizeof(float), (cudaStream_t)stream));
  dcn_col2im_kernel<<<gdmae_grid((long long)g.B * g.C * g.kh * g.kw * g.Ho * g.Wo, 256, 8), 256, 0, (cudaStream_t)stream>>>(g, grad_columns, offset,
                                                                                                                         mask, grad_input);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

extern "C" int gdmae_deform_col2im_coord(const float* grad_columns, const float* input, const float* offset, const float* mask, const int* geom,
                                         float* grad_offset, float* grad_mask, void* stream) {
  DcnGeom g;
  if (int rc = fill_geom(g, geom)) return rc;
  if (g.B == 0) return GDMAE_OK;
  GDMAE_CHECK_ARG((mask == nullptr) == (grad_mask == nullptr));
  dcn_col2im_coord_kernel<<<gdmae_grid((long long)g.B * g.dg * g.kh * g.kw * g.Ho * g.Wo, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      g, grad_columns, input, offset, mask, grad_offset, grad_mask);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}
