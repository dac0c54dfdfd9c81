// Dynamic voxelisation, pillar segments (CSR), pillar mean, MAE random mask and the sst_ops
// index operators.  Integer / byte work, HBM(L2)-bound; no sort of 32-byte rows, no host sync.
//
// Replaces (reference file:line, relative to /root/reference):
//   common_utils.get_in_range_mask            pcdet/utils/common_utils.py:66-76
//   DynVFE.forward voxelise + unique          pcdet/models/backbones_3d/vfe/dyn_vfe.py:60-68
//   torch_scatter.scatter(..., 'mean')        pcdet/models/backbones_3d/vfe/dyn_vfe.py:81
//   common_utils.random_masking               pcdet/utils/common_utils.py:49-63
//   ingroup_inds_wrapper / group_inner_inds   pcdet/ops/sst_ops/src/sst_ops.cpp:21-48, sst_ops_gpu.cu:14-39
//
// B200-first design: instead of torch.unique(dim=0) (a thrust sort of Np 32-byte rows + sync)
// the pillar grid (B*Z*Y*X cells, 7 MB at Waymo B=8) lives in the 126 MB L2: points count
// themselves into their cell, one exclusive scan over the cells yields the lexicographically
// sorted pillar rank AND the CSR segment offsets, and a 21-bit key radix sort of point indices
// gives the stable point order inside each pillar (= the canonical outcome of the reference's
// atomic race, see oracle/gdmae_oracle.py).
#include "common.cuh"
#include <cub/cub.cuh>

struct VoxParams {
  float r0, r1, r2;   // range min x,y,z
  float v0, v1, v2;   // voxel size x,y,z
  int X, Y, Z, B;
  int n_cols;
};

struct NonZeroOp {
  __host__ __device__ __forceinline__ int operator()(const int& c) const { return c > 0 ? 1 : 0; }
};

// cell[i] = linear cell id ((b*Z+z)*Y+y)*X+x, or -1 when dropped.  fp32 sub, fp32 IEEE divide,
// truncate toward zero on the int64 value, then the range test - exactly common_utils.py:74-75.
__global__ void vox_mark_kernel(const float* __restrict__ pts, long long n, VoxParams p, int* __restrict__ cell,
                                int* __restrict__ keep, int* __restrict__ sort_key, int* __restrict__ cellcnt,
                                int n_cells, int* __restrict__ counts) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float* row = pts + i * p.n_cols;
    float fb = row[0], x = row[1], y = row[2], z = row[3];
    long long cx = (long long)__fdiv_rn(__fsub_rn(x, p.r0), p.v0);
    long long cy = (long long)__fdiv_rn(__fsub_rn(y, p.r1), p.v1);
    long long cz = (long long)__fdiv_rn(__fsub_rn(z, p.r2), p.v2);
    long long b = (long long)fb;
    bool ok = cx >= 0 && cx < p.X && cy >= 0 && cy < p.Y && cz >= 0 && cz < p.Z;
    if (ok && (b < 0 || b >= p.B)) {  // frame index outside [0, batch_size): flag and drop
      atomicOr(&counts[2], 1);
      ok = false;
    }
    int c = -1;
    if (ok) {
      c = (int)(((b * p.Z + cz) * p.Y + cy) * p.X + cx);
      atomicAdd(&cellcnt[c], 1);
    }
    cell[i] = c;
    keep[i] = ok ? 1 : 0;
    sort_key[i] = ok ? c : n_cells;
  }
}

__global__ void vox_emit_kernel(const float* __restrict__ pts, long long n, VoxParams p, const int* __restrict__ cell,
                                const int* __restrict__ keep_scan, const int* __restrict__ rank,
                                float* __restrict__ out_pts, long long* __restrict__ out_coords,
                                long long* __restrict__ out_inverse, int* __restrict__ sort_val,
                                int* __restrict__ counts) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int c = cell[i];
    int pos = keep_scan[i];
    sort_val[i] = c >= 0 ? pos : -1;
    if (i == n - 1) counts[0] = pos + (c >= 0 ? 1 : 0);
    if (c < 0) continue;
    const float* row = pts + i * p.n_cols;
    float* dst = out_pts + (long long)pos * p.n_cols;
    for (int k = 0; k < p.n_cols; ++k) dst[k] = row[k];
    int x = c % p.X;
    int t = c / p.X;
    int y = t % p.Y;
    t /= p.Y;
    int z = t % p.Z;
    int b = t / p.Z;
    longlong4 v = make_longlong4(b, z, y, x);
    *reinterpret_cast<longlong4*>(out_coords + 4ll * pos) = v;
    out_inverse[pos] = rank[c];
  }
}

__global__ void vox_pillar_kernel(int n_cells, VoxParams p, const int* __restrict__ cellcnt, const int* __restrict__ rank,
                                  const int* __restrict__ cntscan, long long* __restrict__ voxel_coords,
                                  int* __restrict__ cell2pillar, int* __restrict__ seg_off, int* __restrict__ counts) {
  int cells_per_batch = p.X * p.Y * p.Z;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += gridDim.x * blockDim.x) {
    int cnt = cellcnt[c];
    int m = rank[c];
    if (c % cells_per_batch == 0) counts[4 + c / cells_per_batch] = m;
    if (c == n_cells - 1) {
      int M = m + (cnt > 0 ? 1 : 0);
      counts[1] = M;
      counts[4 + p.B] = M;
      seg_off[M] = cntscan[c] + cnt;
    }
    if (cnt > 0) {
      int x = c % p.X;
      int t = c / p.X;
      int y = t % p.Y;
      t /= p.Y;
      int z = t % p.Z;
      int b = t / p.Z;
      *reinterpret_cast<longlong4*>(voxel_coords + 4ll * m) = make_longlong4(b, z, y, x);
      cell2pillar[c] = m;
      seg_off[m] = cntscan[c];
    } else {
      cell2pillar[c] = -1;
    }
  }
}

static int bits_for(long long max_value) {
  int b = 1;
  while ((1ll << b) <= max_value) ++b;
  return b;
}

extern "C" size_t gdmae_dynvox_workspace_bytes(int64_t n_in, int64_t n_cells) {
  size_t scan_tmp = 0, sort_tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (int*)nullptr, (int*)nullptr, (int)(n_in > n_cells ? n_in : n_cells));
  cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (int*)nullptr, (int*)nullptr, (int*)nullptr, (int*)nullptr, (int)n_in);
  size_t tmp = scan_tmp > sort_tmp ? scan_tmp : sort_tmp;
  return gdmae_align(tmp) + 6 * gdmae_align((size_t)n_in * 4) + 3 * gdmae_align((size_t)n_cells * 4) + 4096;
}

extern "C" int gdmae_dynvox(const float* points, int64_t n_in, int n_cols, const float* pc_range, const float* voxel,
                            const int* grid_xyz, int batch_size, float* out_points, int64_t* out_point_coords,
                            int64_t* out_inverse, int64_t* out_voxel_coords, int32_t* out_cell2pillar,
                            int32_t* out_seg_offsets, int32_t* out_seg_points, int32_t* out_counts, void* workspace,
                            size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GDMAE_CHECK_ARG(n_in >= 0 && n_in < (1ll << 31) && n_cols >= 4 && batch_size >= 1);
  VoxParams p;
  p.r0 = pc_range[0]; p.r1 = pc_range[1]; p.r2 = pc_range[2];
  p.v0 = voxel[0]; p.v1 = voxel[1]; p.v2 = voxel[2];
  p.X = grid_xyz[0]; p.Y = grid_xyz[1]; p.Z = grid_xyz[2]; p.B = batch_size; p.n_cols = n_cols;
  long long n_cells_ll = (long long)p.X * p.Y * p.Z * p.B;
  GDMAE_CHECK_ARG(n_cells_ll > 0 && n_cells_ll < (1ll << 30));
  int n_cells = (int)n_cells_ll;
  if (ws_bytes < gdmae_dynvox_workspace_bytes(n_in, n_cells)) { gdmae_set_error("dynvox: workspace too small"); return GDMAE_ERR_WORKSPACE; }
  Workspace ws(workspace, ws_bytes);
  size_t scan_tmp = 0, sort_tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (int*)nullptr, (int*)nullptr, (int)(n_in > n_cells ? n_in : n_cells));
  cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (int*)nullptr, (int*)nullptr, (int*)nullptr, (int*)nullptr, (int)n_in);
  size_t tmp_bytes = scan_tmp > sort_tmp ? scan_tmp : sort_tmp;
  char* tmp = ws.take<char>(tmp_bytes);
  int* cell = ws.take<int>(n_in);
  int* keep = ws.take<int>(n_in);
  int* keep_scan = ws.take<int>(n_in);
  int* sort_key = ws.take<int>(n_in);
  int* sort_key_out = ws.take<int>(n_in);
  int* sort_val = ws.take<int>(n_in);
  int* cellcnt = ws.take<int>(n_cells);
  int* rank = ws.take<int>(n_cells);
  int* cntscan = ws.take<int>(n_cells);
  if (!tmp || !cell || !keep || !keep_scan || !sort_key || !sort_key_out || !sort_val || !cellcnt || !rank || !cntscan) {
    gdmae_set_error("dynvox: workspace carve failed");
    return GDMAE_ERR_WORKSPACE;
  }
  GDMAE_CHECK_CUDA(cudaMemsetAsync(cellcnt, 0, (size_t)n_cells * 4, stream));
  GDMAE_CHECK_CUDA(cudaMemsetAsync(out_counts, 0, (size_t)(4 + batch_size + 1) * 4, stream));
  const int T = 256;
  if (n_in > 0) {
    vox_mark_kernel<<<gdmae_grid(n_in, T), T, 0, stream>>>(points, n_in, p, cell, keep, sort_key, cellcnt, n_cells, out_counts);
    GDMAE_LAUNCH_CHECK();
    size_t tb = tmp_bytes;
    GDMAE_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, keep, keep_scan, (int)n_in, stream));
  }
  {
    size_t tb = tmp_bytes;
    cub::TransformInputIterator<int, NonZeroOp, const int*> occ(cellcnt, NonZeroOp());
    GDMAE_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, occ, rank, n_cells, stream));
    tb = tmp_bytes;
    GDMAE_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, cellcnt, cntscan, n_cells, stream));
  }
  if (n_in > 0) {
    vox_emit_kernel<<<gdmae_grid(n_in, T), T, 0, stream>>>(points, n_in, p, cell, keep_scan, rank, out_points,
                                                          (long long*)out_point_coords, (long long*)out_inverse, sort_val,
                                                          out_counts);
    GDMAE_LAUNCH_CHECK();
    size_t tb = tmp_bytes;
    GDMAE_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, sort_key, sort_key_out, sort_val, out_seg_points, (int)n_in, 0,
                                                     bits_for(n_cells), stream));
  }
  vox_pillar_kernel<<<gdmae_grid(n_cells, T), T, 0, stream>>>(n_cells, p, cellcnt, rank, cntscan, (long long*)out_voxel_coords,
                                                             out_cell2pillar, out_seg_offsets, out_counts);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// ---------------------------------------------------------------------------------------------
// pillar mean: one thread per (pillar, channel-slot) walks the pillar's points in ascending point
// index and adds sequentially with __fadd_rn, reproducing torch_scatter's CPU summation order
// bit for bit (sum, then divide by max(count,1)).  src rows are L2 resident (31 MB at B=8).
// ---------------------------------------------------------------------------------------------
__global__ void segment_mean_kernel(const float* __restrict__ src, int src_stride, int col0, int C,
                                    const int* __restrict__ seg_off, const int* __restrict__ seg_pts, int M,
                                    float* __restrict__ out) {
  long long total = (long long)M * C;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int m = (int)(t / C), c = (int)(t % C);
    int s = seg_off[m], e = seg_off[m + 1];
    float acc = 0.f;
    for (int k = s; k < e; ++k) acc = __fadd_rn(acc, src[(long long)seg_pts[k] * src_stride + col0 + c]);
    int cnt = e - s;
    out[t] = __fdiv_rn(acc, (float)(cnt > 1 ? cnt : 1));
  }
}

extern "C" int gdmae_segment_mean(const float* src, int src_stride, int col0, int C, const int32_t* seg_offsets,
                                  const int32_t* seg_points, int64_t M, float* out, void* stream_) {
  GDMAE_CHECK_ARG(M >= 0 && C > 0 && src_stride >= col0 + C);
  if (M == 0) return GDMAE_OK;
  segment_mean_kernel<<<gdmae_grid(M * C, 256), 256, 0, (cudaStream_t)stream_>>>(src, src_stride, col0, C, seg_offsets,
                                                                               seg_points, (int)M, out);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// ---------------------------------------------------------------------------------------------
// MAE random mask.  key = (frame << 32) | bits(noise) (noise >= 0 so the bit pattern orders like
// the value); a stable radix sort gives argsort(noise) per frame with ties broken by index; the
// first int(L * (1 - ratio)) of each frame (double arithmetic, like the Python expression) are visible.
// ---------------------------------------------------------------------------------------------
__global__ void mask_keys_kernel(const float* __restrict__ noise, const int* __restrict__ batch_off, int B, int M,
                                 unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
    int b = 0;
    while (b + 1 < B && i >= batch_off[b + 1]) ++b;
    keys[i] = ((unsigned long long)b << 32) | (unsigned long long)__float_as_uint(noise[i]);
    vals[i] = i;
  }
}

__global__ void mask_write_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ vals,
                                  const int* __restrict__ batch_off, double keep_ratio, int M, float* __restrict__ mask) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
    int b = (int)(keys[i] >> 32);
    int L = batch_off[b + 1] - batch_off[b];
    long long len_keep = (long long)((double)L * keep_ratio);
    mask[vals[i]] = (i - batch_off[b]) < len_keep ? 0.f : 1.f;
  }
}

extern "C" size_t gdmae_random_mask_workspace_bytes(int64_t M) {
  size_t sort_tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (int*)nullptr, (int*)nullptr, (int)M);
  return gdmae_align(sort_tmp) + 2 * gdmae_align((size_t)M * 8) + 2 * gdmae_align((size_t)M * 4) + 1024;
}

extern "C" int gdmae_random_mask(const float* noise, int64_t M, const int32_t* batch_offsets, int batch_size,
                                 double keep_ratio, float* out_mask, void* workspace, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GDMAE_CHECK_ARG(M >= 0 && batch_size >= 1);
  if (M == 0) return GDMAE_OK;
  if (ws_bytes < gdmae_random_mask_workspace_bytes(M)) { gdmae_set_error("random_mask: workspace too small"); return GDMAE_ERR_WORKSPACE; }
  Workspace ws(workspace, ws_bytes);
  size_t sort_tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (int*)nullptr, (int*)nullptr, (int)M);
  char* tmp = ws.take<char>(sort_tmp);
  unsigned long long* keys = ws.take<unsigned long long>(M);
  unsigned long long* keys_out = ws.take<unsigned long long>(M);
  int* vals = ws.take<int>(M);
  int* vals_out = ws.take<int>(M);
  if (!tmp || !keys || !keys_out || !vals || !vals_out) { gdmae_set_error("random_mask: workspace carve failed"); return GDMAE_ERR_WORKSPACE; }
  mask_keys_kernel<<<gdmae_grid(M, 256), 256, 0, stream>>>(noise, batch_offsets, batch_size, (int)M, keys, vals);
  GDMAE_LAUNCH_CHECK();
  GDMAE_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(tmp, sort_tmp, keys, keys_out, vals, vals_out, (int)M, 0,
                                                   32 + bits_for(batch_size), stream));
  mask_write_kernel<<<gdmae_grid(M, 256), 256, 0, stream>>>(keys_out, vals_out, batch_offsets, keep_ratio, (int)M, out_mask);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// ---------------------------------------------------------------------------------------------
// sst_ops operators on arbitrary int64 group ids (the reference's Python-visible API).
// Stable sort by group id, then rank inside the run = position - run start.
// ---------------------------------------------------------------------------------------------
__global__ void iota_kernel(int n, int* __restrict__ v) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = i;
}

__global__ void run_start_kernel(const long long* __restrict__ keys, int n, int* __restrict__ start) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    start[i] = (i == 0 || keys[i] != keys[i - 1]) ? i : 0;
}

__global__ void ingroup_write_kernel(const int* __restrict__ vals, const int* __restrict__ start_max, int n,
                                     long long* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[vals[i]] = (long long)(i - start_max[i]);
}

extern "C" size_t gdmae_ingroup_inds_workspace_bytes(int64_t N) {
  size_t sort_tmp = 0, scan_tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (long long*)nullptr, (long long*)nullptr, (int*)nullptr, (int*)nullptr, (int)N);
  cub::DeviceScan::InclusiveScan(nullptr, scan_tmp, (int*)nullptr, (int*)nullptr, cub::Max(), (int)N);
  size_t tmp = sort_tmp > scan_tmp ? sort_tmp : scan_tmp;
  return gdmae_align(tmp) + gdmae_align((size_t)N * 8) + 4 * gdmae_align((size_t)N * 4) + 1024;
}

// sst_ops.cpp:21-33 ingroup_inds_wrapper(group_inds, out_inds): out[i] = rank of i inside its group.
extern "C" int gdmae_ingroup_inds(const int64_t* group_inds, int64_t N, int64_t* out_inds, void* workspace, size_t ws_bytes,
                                  void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GDMAE_CHECK_ARG(N >= 0 && N < (1ll << 31));
  if (N == 0) return GDMAE_OK;
  if (ws_bytes < gdmae_ingroup_inds_workspace_bytes(N)) { gdmae_set_error("ingroup_inds: workspace too small"); return GDMAE_ERR_WORKSPACE; }
  Workspace ws(workspace, ws_bytes);
  size_t sort_tmp = 0, scan_tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (long long*)nullptr, (long long*)nullptr, (int*)nullptr, (int*)nullptr, (int)N);
  cub::DeviceScan::InclusiveScan(nullptr, scan_tmp, (int*)nullptr, (int*)nullptr, cub::Max(), (int)N);
  size_t tmp_bytes = sort_tmp > scan_tmp ? sort_tmp : scan_tmp;
  char* tmp = ws.take<char>(tmp_bytes);
  long long* keys_out = ws.take<long long>(N);
  int* vals = ws.take<int>(N);
  int* vals_out = ws.take<int>(N);
  int* start = ws.take<int>(N);
  int* start_max = ws.take<int>(N);
  if (!tmp || !keys_out || !vals || !vals_out || !start || !start_max) { gdmae_set_error("ingroup_inds: workspace carve failed"); return GDMAE_ERR_WORKSPACE; }
  int g = gdmae_grid(N, 256);
  iota_kernel<<<g, 256, 0, stream>>>((int)N, vals);
  GDMAE_LAUNCH_CHECK();
  size_t tb = tmp_bytes;
  GDMAE_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, (const long long*)group_inds, keys_out, vals, vals_out, (int)N, 0, 64, stream));
  run_start_kernel<<<g, 256, 0, stream>>>(keys_out, (int)N, start);
  GDMAE_LAUNCH_CHECK();
  tb = tmp_bytes;
  GDMAE_CHECK_CUDA(cub::DeviceScan::InclusiveScan(tmp, tb, start, start_max, cub::Max(), (int)N, stream));
  ingroup_write_kernel<<<g, 256, 0, stream>>>(vals_out, start_max, (int)N, (long long*)out_inds);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// group_inds[m, k] = k-th point (ascending index) of pillar m; slots cnt..K-1 repeat cyclically
// (sst_ops_gpu.cu:22-39).  CSR form: no atomics, no counter allocation.
__global__ void group_fill_kernel(const int* __restrict__ seg_off, const int* __restrict__ seg_pts, int M, int K,
                                  long long* __restrict__ out) {
  long long total = (long long)M * K;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int m = (int)(t / K), k = (int)(t % K);
    int s = seg_off[m], cnt = seg_off[m + 1] - s;
    out[t] = cnt == 0 ? -1ll : (long long)seg_pts[s + (k < cnt ? k : k % cnt)];
  }
}

extern "C" int gdmae_group_inner_inds_csr(const int32_t* seg_offsets, const int32_t* seg_points, int64_t M, int K,
                                          int64_t* out_group_inds, void* stream_) {
  GDMAE_CHECK_ARG(M >= 0 && K > 0);
  if (M == 0) return GDMAE_OK;
  group_fill_kernel<<<gdmae_grid(M * K, 256), 256, 0, (cudaStream_t)stream_>>>(seg_offsets, seg_points, (int)M, K,
                                                                             (long long*)out_group_inds);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

__global__ void hist_kernel(const long long* __restrict__ inv, int n, int M, int* __restrict__ cnt, int* __restrict__ key32,
                            int* __restrict__ vals, int* __restrict__ err) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    long long g = inv[i];
    if (g < 0 || g >= M) { atomicOr(err, 1); g = 0; }
    atomicAdd(&cnt[g], 1);
    key32[i] = (int)g;
    vals[i] = i;
  }
}

extern "C" size_t gdmae_group_inner_inds_workspace_bytes(int64_t Np, int64_t M) {
  size_t sort_tmp = 0, scan_tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (int*)nullptr, (int*)nullptr, (int*)nullptr, (int*)nullptr, (int)Np);
  cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (int*)nullptr, (int*)nullptr, (int)(M + 1));
  size_t tmp = sort_tmp > scan_tmp ? sort_tmp : scan_tmp;
  return gdmae_align(tmp) + 4 * gdmae_align((size_t)Np * 4) + 2 * gdmae_align((size_t)(M + 1) * 4) + 1024;
}

// sst_ops.cpp:35-48 group_inner_inds_wrapper(inverse_inds (Np), group_inds (M,K)).
extern "C" int gdmae_group_inner_inds(const int64_t* inverse_inds, int64_t Np, int64_t M, int K, int64_t* out_group_inds,
                                      void* workspace, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GDMAE_CHECK_ARG(Np >= 0 && Np < (1ll << 31) && M >= 0 && M < (1ll << 31) && K > 0);
  if (M == 0) return GDMAE_OK;
  if (ws_bytes < gdmae_group_inner_inds_workspace_bytes(Np, M)) { gdmae_set_error("group_inner_inds: workspace too small"); return GDMAE_ERR_WORKSPACE; }
  Workspace ws(workspace, ws_bytes);
  size_t sort_tmp = 0, scan_tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (int*)nullptr, (int*)nullptr, (int*)nullptr, (int*)nullptr, (int)Np);
  cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (int*)nullptr, (int*)nullptr, (int)(M + 1));
  size_t tmp_bytes = sort_tmp > scan_tmp ? sort_tmp : scan_tmp;
  char* tmp = ws.take<char>(tmp_bytes);
  int* key32 = ws.take<int>(Np);
  int* key_out = ws.take<int>(Np);
  int* vals = ws.take<int>(Np);
  int* vals_out = ws.take<int>(Np);
  int* cnt = ws.take<int>(M + 1);
  int* off = ws.take<int>(M + 1);
  if (!tmp || !key32 || !key_out || !vals || !vals_out || !cnt || !off) { gdmae_set_error("group_inner_inds: workspace carve failed"); return GDMAE_ERR_WORKSPACE; }
  GDMAE_CHECK_CUDA(cudaMemsetAsync(cnt, 0, (size_t)(M + 1) * 4, stream));
  if (Np > 0) {
    hist_kernel<<<gdmae_grid(Np, 256), 256, 0, stream>>>((const long long*)inverse_inds, (int)Np, (int)M, cnt, key32, vals, cnt + M);
    GDMAE_LAUNCH_CHECK();
    size_t tb = tmp_bytes;
    GDMAE_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, key32, key_out, vals, vals_out, (int)Np, 0, bits_for(M), stream));
  }
  // cnt[M] is an error flag slot, not a count: the scan over M+1 items still yields off[M] = Np
  size_t tb = tmp_bytes;
  GDMAE_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, off, (int)(M + 1), stream));
  group_fill_kernel<<<gdmae_grid(M * K, 256), 256, 0, stream>>>(off, vals_out, (int)M, K, (long long*)out_group_inds);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}
