// Structure (index) kernels of the Sparse Pyramid Transformer: visible-site compaction, strided
// sparse-conv site sets, 3x3 neighbour maps ("rulebooks") and the window tables of the SRA blocks.
// All integer work on dense int32 grids that stay resident in the 126 MB L2 (B*468*468*4 B = 7 MB).
//
// Replaces (reference file:line, relative to /root/reference):
//   SparseConvTensor construction of the visible pillars  pcdet/models/backbones_3d/spt_backbone_mae.py:102-107
//   spconv SparseConv2d(k3,s2,p1)/SubMConv2d indice pairs pcdet/utils/spconv_utils.py:37-56 (third party spconv 2.x)
//   sst_utils.get_window_coors                            pcdet/models/model_utils/sst_utils.py:6-47
//   get_inner_win_inds / drop_single_shift                pcdet/models/backbones_3d/spt_backbone.py:32-51
//   make_continuous_inds / get_flat2win_inds              pcdet/models/model_utils/sst_utils.py:50-96
//
// B200-first design: an 8x8 window is a 64-bit occupancy word.  The rank of a token inside its
// window is popc(word & below(pos)), the window population is popc(word), the drop level follows
// from it - no atomics counters, no unique/sort, no host sync (the reference does 18 unique+sort
// calls and ~250 host syncs per forward for the same information).
#include "common.cuh"
#include <cub/cub.cuh>

// ------------------------------------------------------------------ visible sites (mask == 0)
__global__ void vis_flag_kernel(const float* __restrict__ mask, int M, int* __restrict__ flag) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) flag[i] = mask[i] == 0.f ? 1 : 0;
}

__global__ void vis_emit_kernel(const long long* __restrict__ vcoords, const int* __restrict__ flag, const int* __restrict__ pos,
                                int M, int Y, int X, int* __restrict__ vis_idx, int* __restrict__ indices,
                                int* __restrict__ rank_grid, int* __restrict__ count) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
    if (i == M - 1) count[0] = pos[i] + flag[i];
    if (!flag[i]) continue;
    int p = pos[i];
    int b = (int)vcoords[4ll * i], y = (int)vcoords[4ll * i + 2], x = (int)vcoords[4ll * i + 3];
    vis_idx[p] = i;
    indices[3 * p] = b; indices[3 * p + 1] = y; indices[3 * p + 2] = x;
    rank_grid[(b * Y + y) * X + x] = p;
  }
}

extern "C" size_t gdmae_visible_sites_workspace_bytes(int64_t M) {
  size_t scan_tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (int*)nullptr, (int*)nullptr, (int)M);
  return gdmae_align(scan_tmp) + 2 * gdmae_align((size_t)M * 4) + 1024;
}

// voxel_coords (M,4) int64 [b,z,y,x] sorted; mask (M) float {0 visible, 1 masked}.
// out_rank_grid (B*Y*X) int32 is fully written (-1 at cells without a visible pillar).
extern "C" int gdmae_visible_sites(const int64_t* voxel_coords, const float* mask, int64_t M, int B, int Y, int X,
                                   int32_t* out_vis_idx, int32_t* out_indices, int32_t* out_rank_grid, int32_t* out_count,
                                   void* workspace, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GDMAE_CHECK_ARG(M >= 0 && M < (1ll << 31) && B >= 1 && Y >= 1 && X >= 1);
  if (ws_bytes < gdmae_visible_sites_workspace_bytes(M)) { gdmae_set_error("visible_sites: workspace too small"); return GDMAE_ERR_WORKSPACE; }
  GDMAE_CHECK_CUDA(cudaMemsetAsync(out_rank_grid, 0xff, (size_t)B * Y * X * 4, stream));
  GDMAE_CHECK_CUDA(cudaMemsetAsync(out_count, 0, 4, stream));
  if (M == 0) return GDMAE_OK;
  Workspace ws(workspace, ws_bytes);
  size_t scan_tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (int*)nullptr, (int*)nullptr, (int)M);
  char* tmp = ws.take<char>(scan_tmp);
  int* flag = ws.take<int>(M);
  int* pos = ws.take<int>(M);
  if (!tmp || !flag || !pos) { gdmae_set_error("visible_sites: workspace carve failed"); return GDMAE_ERR_WORKSPACE; }
  int g = gdmae_grid(M, 256);
  vis_flag_kernel<<<g, 256, 0, stream>>>(mask, (int)M, flag);
  GDMAE_LAUNCH_CHECK();
  GDMAE_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(tmp, scan_tmp, flag, pos, (int)M, stream));
  vis_emit_kernel<<<g, 256, 0, stream>>>((const long long*)voxel_coords, flag, pos, (int)M, Y, X, out_vis_idx, out_indices,
                                        out_rank_grid, out_count);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// ------------------------------------------------------------------ rank grid of a user-built sparse tensor
__global__ void rank_grid_kernel(const int* __restrict__ idx, int N, int H, int W, int* __restrict__ grid) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
    grid[(idx[3 * i] * H + idx[3 * i + 1]) * W + idx[3 * i + 2]] = i;
}

// rank_grid (B*H*W) int32: row of the site at each cell, -1 where empty.
extern "C" int gdmae_build_rank_grid(const int32_t* indices, int64_t N, int B, int H, int W, int32_t* out_rank_grid,
                                     void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GDMAE_CHECK_ARG(N >= 0 && N < (1ll << 31) && (long long)B * H * W < (1ll << 31));
  GDMAE_CHECK_CUDA(cudaMemsetAsync(out_rank_grid, 0xff, (size_t)B * H * W * 4, stream));
  if (N == 0) return GDMAE_OK;
  rank_grid_kernel<<<gdmae_grid(N, 256), 256, 0, stream>>>(indices, (int)N, H, W, out_rank_grid);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// ------------------------------------------------------------------ strided conv output sites (k3 s2 p1)
__global__ void down_mark_kernel(const int* __restrict__ idx, int N, int Ho, int Wo, int* __restrict__ occ) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    int b = idx[3 * i], y = idx[3 * i + 1], x = idx[3 * i + 2];
    if (b < 0) continue;  // unused tail of a capacity buffer (rows pre-filled with -1)
    // output oy covers input rows 2*oy-1 .. 2*oy+1
    int oy0 = y >> 1, oy1 = (y + 1) >> 1;
    int ox0 = x >> 1, ox1 = (x + 1) >> 1;
    for (int oy = oy0; oy <= oy1; ++oy) {
      if (oy >= Ho) continue;
      for (int ox = ox0; ox <= ox1; ++ox) {
        if (ox >= Wo) continue;
        occ[(b * Ho + oy) * Wo + ox] = 1;
      }
    }
  }
}

__global__ void down_emit_kernel(const int* __restrict__ occ, const int* __restrict__ rank, int n_cells, int Ho, int Wo,
                                 int* __restrict__ out_idx, int* __restrict__ rank_grid, int* __restrict__ count) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += gridDim.x * blockDim.x) {
    int o = occ[c], r = rank[c];
    if (c == n_cells - 1) count[0] = r + o;
    if (o) {
      int x = c % Wo, t = c / Wo;
      out_idx[3 * r] = t / Ho; out_idx[3 * r + 1] = t % Ho; out_idx[3 * r + 2] = x;
      rank_grid[c] = r;
    } else {
      rank_grid[c] = -1;
    }
  }
}

extern "C" size_t gdmae_down_sites_workspace_bytes(int64_t n_out_cells) {
  size_t scan_tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (int*)nullptr, (int*)nullptr, (int)n_out_cells);
  return gdmae_align(scan_tmp) + 2 * gdmae_align((size_t)n_out_cells * 4) + 1024;
}

// in_indices (N,3) int32 [b,y,x] on an H x W grid -> active output sites of SparseConv2d(3, stride 2, pad 1)
// on the Ho x Wo grid (Ho = (H-1)/2+1), lexicographic order; out_rank_grid (B*Ho*Wo) fully written.
extern "C" int gdmae_down_sites(const int32_t* in_indices, int64_t N, int B, int H, int W, int32_t* out_indices,
                                int32_t* out_rank_grid, int32_t* out_count, void* workspace, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  long long n_cells = (long long)B * Ho * Wo;
  GDMAE_CHECK_ARG(N >= 0 && n_cells > 0 && n_cells < (1ll << 31));
  if (ws_bytes < gdmae_down_sites_workspace_bytes(n_cells)) { gdmae_set_error("down_sites: workspace too small"); return GDMAE_ERR_WORKSPACE; }
  Workspace ws(workspace, ws_bytes);
  size_t scan_tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (int*)nullptr, (int*)nullptr, (int)n_cells);
  char* tmp = ws.take<char>(scan_tmp);
  int* occ = ws.take<int>(n_cells);
  int* rank = ws.take<int>(n_cells);
  if (!tmp || !occ || !rank) { gdmae_set_error("down_sites: workspace carve failed"); return GDMAE_ERR_WORKSPACE; }
  GDMAE_CHECK_CUDA(cudaMemsetAsync(occ, 0, (size_t)n_cells * 4, stream));
  if (N > 0) {
    down_mark_kernel<<<gdmae_grid(N, 256), 256, 0, stream>>>(in_indices, (int)N, Ho, Wo, occ);
    GDMAE_LAUNCH_CHECK();
  }
  GDMAE_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(tmp, scan_tmp, occ, rank, (int)n_cells, stream));
  down_emit_kernel<<<gdmae_grid(n_cells, 256), 256, 0, stream>>>(occ, rank, (int)n_cells, Ho, Wo, out_indices, out_rank_grid, out_count);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// ------------------------------------------------------------------ neighbour maps
// subm: nbr[n, ky*3+kx] = row of the active site at (y+ky-1, x+kx-1) or -1.
__global__ void subm_map_kernel(const int* __restrict__ idx, int N, const int* __restrict__ grid, int H, int W,
                                int* __restrict__ nbr) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < N * 9; t += gridDim.x * blockDim.x) {
    int n = t / 9, k = t % 9;
    int b = idx[3 * n], y = idx[3 * n + 1] + k / 3 - 1, x = idx[3 * n + 2] + k % 3 - 1;
    nbr[t] = (y >= 0 && y < H && x >= 0 && x < W) ? grid[(b * H + y) * W + x] : -1;
  }
}

extern "C" int gdmae_subm_neighbor_map(const int32_t* indices, int64_t N, const int32_t* rank_grid, int B, int H, int W,
                                       int32_t* out_nbr, void* stream_) {
  GDMAE_CHECK_ARG(N >= 0 && N * 9 < (1ll << 31));
  (void)B;
  if (N == 0) return GDMAE_OK;
  subm_map_kernel<<<gdmae_grid(N * 9, 256), 256, 0, (cudaStream_t)stream_>>>(indices, (int)N, rank_grid, H, W, out_nbr);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// down: nbr_down[o, k] = input row at (2*oy-1+ky, 2*ox-1+kx) or -1        (forward gather)
__global__ void down_map_kernel(const int* __restrict__ oidx, int No, const int* __restrict__ in_grid, int H, int W,
                                int* __restrict__ nbr) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < No * 9; t += gridDim.x * blockDim.x) {
    int n = t / 9, k = t % 9;
    int b = oidx[3 * n], y = 2 * oidx[3 * n + 1] - 1 + k / 3, x = 2 * oidx[3 * n + 2] - 1 + k % 3;
    nbr[t] = (y >= 0 && y < H && x >= 0 && x < W) ? in_grid[(b * H + y) * W + x] : -1;
  }
}
// up: nbr_up[i, k] = output row o with (2*oy-1+ky, 2*ox-1+kx) == (y, x) or -1   (backward gather)
__global__ void up_map_kernel(const int* __restrict__ iidx, int N, const int* __restrict__ out_grid, int Ho, int Wo,
                              int* __restrict__ nbr) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < N * 9; t += gridDim.x * blockDim.x) {
    int n = t / 9, k = t % 9;
    int b = iidx[3 * n], ty = iidx[3 * n + 1] + 1 - k / 3, tx = iidx[3 * n + 2] + 1 - k % 3;
    int r = -1;
    if (ty >= 0 && tx >= 0 && !(ty & 1) && !(tx & 1)) {
      int oy = ty >> 1, ox = tx >> 1;
      if (oy < Ho && ox < Wo) r = out_grid[(b * Ho + oy) * Wo + ox];
    }
    nbr[t] = r;
  }
}

extern "C" int gdmae_down_neighbor_maps(const int32_t* in_indices, int64_t N, const int32_t* in_rank_grid, int H, int W,
                                        const int32_t* out_indices, int64_t No, const int32_t* out_rank_grid,
                                        int32_t* out_nbr_down, int32_t* out_nbr_up, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  GDMAE_CHECK_ARG(N >= 0 && No >= 0 && N * 9 < (1ll << 31) && No * 9 < (1ll << 31));
  if (No > 0) {
    down_map_kernel<<<gdmae_grid(No * 9, 256), 256, 0, stream>>>(out_indices, (int)No, in_rank_grid, H, W, out_nbr_down);
    GDMAE_LAUNCH_CHECK();
  }
  if (N > 0) {
    up_map_kernel<<<gdmae_grid(N * 9, 256), 256, 0, stream>>>(in_indices, (int)N, out_rank_grid, Ho, Wo, out_nbr_up);
    GDMAE_LAUNCH_CHECK();
  }
  return GDMAE_OK;
}

// ------------------------------------------------------------------ window tables (8x8x1 windows)
// shift 0 adds a full window (8), shift 1 half a window (4) (sst_utils.py:18-21).  Dense window id
// w = (b*nWx + wx)*nWy + wy, the reference's batch_win_inds equals 2*w (max_num_win_z = 2, win z = 0).
__global__ void win_mark_kernel(const int* __restrict__ idx, int N, int nWx, int nWy, int shift, int* __restrict__ win_of,
                                unsigned char* __restrict__ pos_of, unsigned long long* __restrict__ wmask) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    int b = idx[3 * i], y = idx[3 * i + 1] + shift, x = idx[3 * i + 2] + shift;
    int w = (b * nWx + (x >> 3)) * nWy + (y >> 3);
    int pos = ((y & 7) << 3) | (x & 7);
    win_of[i] = w;
    pos_of[i] = (unsigned char)pos;
    atomicOr(&wmask[w], 1ull << pos);
  }
}

__host__ __device__ __forceinline__ int popc64(unsigned long long m) {
#ifdef __CUDA_ARCH__
  return __popcll(m);
#else
  return __builtin_popcountll(m);
#endif
}

struct LevelCount {  // three 21-bit counters packed into one word so a single scan ranks all levels
  __host__ __device__ __forceinline__ unsigned long long operator()(const unsigned long long& m) const {
    int n = popc64(m);
    if (n == 0) return 0ull;
    return n < 16 ? 1ull : (n < 32 ? (1ull << 21) : (1ull << 42));
  }
};
struct PopCount {
  __host__ __device__ __forceinline__ int operator()(const unsigned long long& m) const { return popc64(m); }
};

__global__ void win_finish_kernel(const int* __restrict__ win_of, const unsigned char* __restrict__ pos_of, int N,
                                  const unsigned long long* __restrict__ wmask, const int* __restrict__ win_off,
                                  int* __restrict__ inner, int* __restrict__ level, int* __restrict__ win_tok,
                                  int4* __restrict__ row_info) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    int w = win_of[i];
    unsigned long long m = wmask[w];
    int pos = pos_of[i];
    int r = __popcll(m & ((1ull << pos) - 1ull));
    int n = __popcll(m);
    inner[i] = r;
    level[i] = n < 16 ? 0 : (n < 32 ? 1 : 2);
    int s = win_off[w];
    win_tok[s + r] = i;
    // everything the SRA kernels need about CSR row (s + r), one 16-byte load: token, window extent, cell
    row_info[s + r] = make_int4(i, s, s + n, pos);
  }
}

__global__ void win_level_kernel(const unsigned long long* __restrict__ wmask, const unsigned long long* __restrict__ lscan,
                                 int nW, int* __restrict__ lvl_rank, int* __restrict__ lvl_counts, int* __restrict__ win_off, int N) {
  for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < nW; w += gridDim.x * blockDim.x) {
    int n = __popcll(wmask[w]);
    unsigned long long s = lscan[w];
    int l = n < 16 ? 0 : (n < 32 ? 1 : 2);
    lvl_rank[w] = n == 0 ? -1 : (int)((s >> (21 * l)) & 0x1fffffull);
    if (w == nW - 1) {
      unsigned long long tot = s + LevelCount()(wmask[w]);
      lvl_counts[0] = (int)(tot & 0x1fffffull);
      lvl_counts[1] = (int)((tot >> 21) & 0x1fffffull);
      lvl_counts[2] = (int)((tot >> 42) & 0x1fffffull);
      win_off[nW] = N;
    }
  }
}

extern "C" size_t gdmae_window_table_workspace_bytes(int64_t n_windows) {
  size_t a = 0, b = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, a, (int*)nullptr, (int*)nullptr, (int)n_windows);
  cub::DeviceScan::ExclusiveSum(nullptr, b, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)n_windows);
  return gdmae_align(a > b ? a : b) + gdmae_align((size_t)n_windows * 8) + 1024;
}

// indices (N,3) int32 [b,y,x] sorted lexicographically on an H x W grid.  nWx = ceil(W/8)+1, nWy = ceil(H/8)+1.
// Outputs: win_of_token (N) dense window id; pos_of_token (N) u8 = yy*8+xx; inner (N) rank in window;
// level (N) drop level; win_mask (nW) u64; win_off (nW+1) CSR offsets over dense windows;
// win_tok (N) tokens grouped by window; lvl_rank (nW) rank of the window among the non-empty
// windows of its level (-1 if empty); lvl_counts (3); row_info (N,4) int32 per CSR row:
// [token, first row of its window, one past its last row, in-window cell].
extern "C" int gdmae_window_table(const int32_t* indices, int64_t N, int B, int H, int W, int shifted,
                                  int32_t* win_of_token, uint8_t* pos_of_token, int32_t* inner, int32_t* level,
                                  uint64_t* win_mask, int32_t* win_off, int32_t* win_tok, int32_t* lvl_rank,
                                  int32_t* lvl_counts, int32_t* row_info, void* workspace, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int nWx = (W + 7) / 8 + 1, nWy = (H + 7) / 8 + 1;
  long long nW = (long long)B * nWx * nWy;
  GDMAE_CHECK_ARG(N >= 0 && N < (1ll << 31) && nW < (1ll << 21));
  if (ws_bytes < gdmae_window_table_workspace_bytes(nW)) { gdmae_set_error("window_table: workspace too small"); return GDMAE_ERR_WORKSPACE; }
  Workspace ws(workspace, ws_bytes);
  size_t a = 0, b = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, a, (int*)nullptr, (int*)nullptr, (int)nW);
  cub::DeviceScan::ExclusiveSum(nullptr, b, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)nW);
  size_t tmp_bytes = a > b ? a : b;
  char* tmp = ws.take<char>(tmp_bytes);
  unsigned long long* lscan = ws.take<unsigned long long>(nW);
  if (!tmp || !lscan) { gdmae_set_error("window_table: workspace carve failed"); return GDMAE_ERR_WORKSPACE; }
  unsigned long long* wm = (unsigned long long*)win_mask;
  GDMAE_CHECK_CUDA(cudaMemsetAsync(wm, 0, (size_t)nW * 8, stream));
  int shift = shifted ? 4 : 8;
  if (N > 0) {
    win_mark_kernel<<<gdmae_grid(N, 256), 256, 0, stream>>>(indices, (int)N, nWx, nWy, shift, win_of_token, pos_of_token, wm);
    GDMAE_LAUNCH_CHECK();
  }
  size_t tb = tmp_bytes;
  cub::TransformInputIterator<int, PopCount, const unsigned long long*> pc(wm, PopCount());
  GDMAE_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, pc, win_off, (int)nW, stream));
  tb = tmp_bytes;
  cub::TransformInputIterator<unsigned long long, LevelCount, const unsigned long long*> lc(wm, LevelCount());
  GDMAE_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, lc, lscan, (int)nW, stream));
  win_level_kernel<<<gdmae_grid(nW, 256), 256, 0, stream>>>(wm, lscan, (int)nW, lvl_rank, lvl_counts, win_off, (int)N);
  GDMAE_LAUNCH_CHECK();
  if (N > 0) {
    win_finish_kernel<<<gdmae_grid(N, 256), 256, 0, stream>>>(win_of_token, pos_of_token, (int)N, wm, win_off, inner, level, win_tok,
                                                              (int4*)row_info);
    GDMAE_LAUNCH_CHECK();
  }
  return GDMAE_OK;
}
