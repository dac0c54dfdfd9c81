// Input side of the step: the world augmentation of DataAugmentor and the point shuffle of
// DataProcessor applied to a whole collated batch on the GPU, instead of frame by frame in numpy dataset workers.
//
// Replaces (reference file:line, relative to /root/reference):
//   DataAugmentor.random_world_flip / random_world_rotation / random_world_scaling   pcdet/datasets/augmentor/data_augmentor.py:55-143
//   common_utils.rotate_points_along_z                                                pcdet/utils/common_utils.py:99-121
//   DataProcessor.shuffle_points                                                      pcdet/datasets/processor/data_processor.py:92-102
// The random draws stay on the host (same numpy stream as the reference, gd-mae_b200/pcdet/datasets/augmentor/
// data_augmentor.py); the kernel applies them: HBM-bound, one pass, rows of 1 + C floats (24 bytes at Waymo).
#include "common.cuh"

#define AUG_MAX_COLS 16

// params (B, 6) per frame: flip_x (negate y), flip_y (negate x), cos, sin, scale, unused
__device__ __forceinline__ void aug_apply(float* v, const float* __restrict__ params, int B) {
  const int b = min(max((int)v[0], 0), B - 1);
  const float* p = params + 6 * b;
  float x = v[1], y = v[2];
  if (__ldg(p + 0) != 0.f) y = -y;
  if (__ldg(p + 1) != 0.f) x = -x;
  const float cs = __ldg(p + 2), sn = __ldg(p + 3), sc = __ldg(p + 4);
  // p @ [[c, s, 0], [-s, c, 0], [0, 0, 1]] in fp32 like the reference's matmul, then xyz *= scale
  const float xr = __fadd_rn(__fmul_rn(x, cs), __fmul_rn(y, -sn));
  const float yr = __fadd_rn(__fmul_rn(x, sn), __fmul_rn(y, cs));
  v[1] = __fmul_rn(xr, sc);
  v[2] = __fmul_rn(yr, sc);
  v[3] = __fmul_rn(v[3], sc);
}

// one thread per row; PAIRS: rows of an even number of floats move as 8-byte words (24-byte rows at Waymo: 3 per row)
template <bool PAIRS>
__global__ void __launch_bounds__(256) world_augment_kernel(const float* __restrict__ points, long long N, int n_cols,
                                                            const float* __restrict__ params, int B,
                                                            const int* __restrict__ src_index, float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    const long long src = src_index ? (long long)src_index[i] : i;
    float v[AUG_MAX_COLS];
    if (PAIRS) {
      const float2* row = reinterpret_cast<const float2*>(points + src * n_cols);
#pragma unroll
      for (int c = 0; c < AUG_MAX_COLS / 2; ++c) {
        const float2 t = 2 * c < n_cols ? __ldg(row + c) : make_float2(0.f, 0.f);
        v[2 * c] = t.x; v[2 * c + 1] = t.y;
      }
    } else {
      const float* row = points + src * n_cols;
#pragma unroll
      for (int c = 0; c < AUG_MAX_COLS; ++c) v[c] = c < n_cols ? __ldg(row + c) : 0.f;
    }
    aug_apply(v, params, B);
    if (PAIRS) {
      float2* dst = reinterpret_cast<float2*>(out + i * n_cols);
#pragma unroll
      for (int c = 0; c < AUG_MAX_COLS / 2; ++c)
        if (2 * c < n_cols) dst[c] = make_float2(v[2 * c], v[2 * c + 1]);
    } else {
      float* dst = out + i * n_cols;
#pragma unroll
      for (int c = 0; c < AUG_MAX_COLS; ++c)
        if (c < n_cols) dst[c] = v[c];
    }
  }
}

// out (N, n_cols) <- world-augmented rows of points (N, n_cols), column 0 = frame index, columns 1..3 = x, y, z.
// params (B, 6) fp32 on the device: flip_x, flip_y, cos(angle), sin(angle), scale, 0.  src_index (N) int32, nullable:
// out row i is built from points row src_index[i] (the shuffle; frames must stay contiguous).  out may not alias points
// when src_index is given.
extern "C" int gdmae_world_augment(const float* points, int64_t N, int n_cols, const float* params, int B, const int32_t* src_index,
                                   float* out, void* stream_) {
  GDMAE_CHECK_ARG(N >= 0 && n_cols >= 4 && n_cols <= AUG_MAX_COLS && B >= 1);
  GDMAE_CHECK_ARG(src_index == nullptr || points != out);
  if (N == 0) return GDMAE_OK;
  const bool pairs = (n_cols % 2) == 0 && ((uintptr_t)points % 8) == 0 && ((uintptr_t)out % 8) == 0;
  if (pairs) world_augment_kernel<true><<<gdmae_grid(N, 256, 16), 256, 0, (cudaStream_t)stream_>>>(points, N, n_cols, params, B, src_index, out);
  else world_augment_kernel<false><<<gdmae_grid(N, 256, 16), 256, 0, (cudaStream_t)stream_>>>(points, N, n_cols, params, B, src_index, out);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}
