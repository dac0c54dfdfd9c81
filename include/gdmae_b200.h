/* gdmae_b200 - C ABI of the B200 (sm_100a) kernels behind GD-MAE's MAE pre-train hot path.
 *
 * Plain pointers and sizes only: no torch / ATen types.  Every pointer named below is a DEVICE
 * pointer unless marked "host".  Every function takes the cudaStream_t to launch on as `stream`
 * (void*), allocates nothing, never synchronises and returns 0 on success or a negative code
 * (GDMAE_ERR_*); gdmae_last_error() gives the message.  Scratch memory comes from the caller:
 * query gdmae_<op>_workspace_bytes(), pass `workspace` / `ws_bytes`.
 *
 * Each entry cites the reference interface it replaces (file:line relative to the reference
 * checkout, Nightmare-n/GD-MAE @ abd05ce).  INTEGRATION.md shows the reference-side bindings.
 */
#ifndef GDMAE_B200_H
#define GDMAE_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GDMAE_OK 0
#define GDMAE_ERR_ARG (-1)
#define GDMAE_ERR_WORKSPACE (-2)
#define GDMAE_ERR_CUDA (-3)

const char* gdmae_last_error(void);
int gdmae_version(void);
int64_t gdmae_launch_count(void); /* hand-written kernels launched so far by this process */
int gdmae_check_device(void); /* 0 iff the current device is sm_100 class; there is no fallback path */
/* bench-only: CUDA events around the SRA launches issued inside the encoder-layer executor.
 * drain: meta = 4 int64 per span (kind 0 fwd / 1 bwd, d, N, algorithmic bytes), ms = elapsed; returns the count */
void gdmae_timing_enable(int on);
int gdmae_timing_drain(int64_t* meta, float* ms, int cap);

/* ---- a1/a2 dynamic voxelisation --------------------------------------------------------------
 * replaces common_utils.get_in_range_mask (pcdet/utils/common_utils.py:66-76) and the boolean
 * compaction + coords.unique(dim=0, return_inverse=True) of DynVFE.forward
 * (pcdet/models/backbones_3d/vfe/dyn_vfe.py:60-68).
 * points (n_in, n_cols) fp32 rows [frame, x, y, z, feat...]; pc_range[6], voxel[3], grid_xyz[3]: host.
 * Outputs (capacities): out_points (n_in, n_cols) kept points in input order; out_point_coords
 * (n_in, 4) int64 [b,z,y,x]; out_inverse (n_in) int64; out_voxel_coords (min(n_in, n_cells), 4)
 * int64, lexicographically sorted unique rows; out_cell2pillar (n_cells) int32, -1 = empty;
 * out_seg_offsets (cap_m + 1) / out_seg_points (n_in): CSR of the kept-point rows of each pillar,
 * ascending point index inside a pillar; out_counts int32[4 + batch_size + 1]:
 * [0] = kept points, [1] = pillars M, [2] = error flags (1: frame index out of range),
 * [4 + b] = first pillar of frame b, [4 + batch_size] = M.   n_cells = B*Z*Y*X. */
size_t gdmae_dynvox_workspace_bytes(int64_t n_in, int64_t n_cells);
int gdmae_dynvox(const float* points, int64_t n_in, int n_cols, const float* pc_range, const float* voxel,
                 const int* grid_xyz, int batch_size, float* out_points, int64_t* out_point_coords,
                 int64_t* out_inverse, int64_t* out_voxel_coords, int32_t* out_cell2pillar,
                 int32_t* out_seg_offsets, int32_t* out_seg_points, int32_t* out_counts, void* workspace,
                 size_t ws_bytes, void* stream);

/* ---- a3 pillar mean --------------------------------------------------------------------------
 * replaces torch_scatter.scatter(points[:, 1:], inverse, dim=0, reduce='mean') (dyn_vfe.py:81).
 * src rows have src_stride floats; columns [col0, col0+C) are averaged. out (M, C). */
int gdmae_segment_mean(const float* src, int src_stride, int col0, int C, const int32_t* seg_offsets,
                       const int32_t* seg_points, int64_t M, float* out, void* stream);

/* ---- a4 point feature build ------------------------------------------------------------------
 * replaces dyn_vfe.py:86-105: out (Np, n_feat+6) = [xyz - pillar centre, points[:,1:], xyz - mean_xyz]. */
int gdmae_vfe_point_features(const float* points, const int64_t* point_coords, const int64_t* inverse,
                             const float* mean, int mean_stride, int64_t Np, int n_cols, const float* pc_range,
                             const float* voxel, float* out, void* stream);

/* ---- a6 pillar max ---------------------------------------------------------------------------
 * replaces torch_scatter.scatter_max(x, inverse, dim=0) (dyn_vfe.py:109-111) and its backward.
 * src (Np, C), out (M, C), out_argmax (M, C) int32 point row (may be NULL in fwd). C % 4 == 0. */
int gdmae_segment_max_fwd(const float* src, int C, const int32_t* seg_offsets, const int32_t* seg_points, int64_t M,
                          float* out, int32_t* out_argmax, void* stream);
int gdmae_segment_max_bwd(const float* dout, const int32_t* argmax, int C, const int32_t* seg_offsets,
                          const int32_t* seg_points, int64_t M, float* dsrc, void* stream);

/* ---- a7 MAE random mask ----------------------------------------------------------------------
 * replaces common_utils.random_masking per frame (pcdet/utils/common_utils.py:49-63,
 * spt_backbone_mae.py:96-100). noise (M) uniform [0,1); batch_offsets (B+1) int32 (= counts + 4);
 * keep_ratio = 1 - mask_ratio as a double. out_mask (M) float: 0 visible, 1 masked. */
size_t gdmae_random_mask_workspace_bytes(int64_t M);
int gdmae_random_mask(const float* noise, int64_t M, const int32_t* batch_offsets, int batch_size,
                      double keep_ratio, float* out_mask, void* workspace, size_t ws_bytes, void* stream);

/* ---- a11 / a24 sst_ops -----------------------------------------------------------------------
 * gdmae_ingroup_inds replaces sst_ops_cuda.ingroup_inds_wrapper (pcdet/ops/sst_ops/src/sst_ops.cpp:21-33,
 * sst_ops_gpu.cu:14-20); gdmae_group_inner_inds replaces sst_ops_cuda.group_inner_inds_wrapper
 * (sst_ops.cpp:35-48, sst_ops_gpu.cu:22-39). Deterministic: order inside a group = ascending index.
 * The _csr variant reuses the pillar CSR of gdmae_dynvox. */
size_t gdmae_ingroup_inds_workspace_bytes(int64_t N);
int gdmae_ingroup_inds(const int64_t* group_inds, int64_t N, int64_t* out_inds, void* workspace, size_t ws_bytes,
                       void* stream);
size_t gdmae_group_inner_inds_workspace_bytes(int64_t Np, int64_t M);
int gdmae_group_inner_inds(const int64_t* inverse_inds, int64_t Np, int64_t M, int K, int64_t* out_group_inds,
                           void* workspace, size_t ws_bytes, void* stream);
int gdmae_group_inner_inds_csr(const int32_t* seg_offsets, const int32_t* seg_points, int64_t M, int K,
                               int64_t* out_group_inds, void* stream);

/* ---- a8/a9/a21 sparse tensor structure -------------------------------------------------------
 * replaces the SparseConvTensor built from the visible pillars (spt_backbone_mae.py:102-107) and the
 * indice-pair generation of spconv's SparseConv2d(3, stride 2, pad 1) / SubMConv2d(3)
 * (pcdet/utils/spconv_utils.py:37-56; spconv 2.x is not vendored in the reference).
 * indices are (N,3) int32 [b,y,x], lexicographic; rank grids are (B*H*W) int32, -1 = empty. */
size_t gdmae_visible_sites_workspace_bytes(int64_t M);
int gdmae_visible_sites(const int64_t* voxel_coords, const float* mask, int64_t M, int B, int Y, int X,
                        int32_t* out_vis_idx, int32_t* out_indices, int32_t* out_rank_grid, int32_t* out_count,
                        void* workspace, size_t ws_bytes, void* stream);
int gdmae_build_rank_grid(const int32_t* indices, int64_t N, int B, int H, int W, int32_t* out_rank_grid, void* stream);
/* in_indices rows with b < 0 are skipped (capacity buffers pre-filled with -1) */
size_t gdmae_down_sites_workspace_bytes(int64_t n_out_cells);
int gdmae_down_sites(const int32_t* in_indices, int64_t N, int B, int H, int W, int32_t* out_indices,
                     int32_t* out_rank_grid, int32_t* out_count, void* workspace, size_t ws_bytes, void* stream);
int gdmae_subm_neighbor_map(const int32_t* indices, int64_t N, const int32_t* rank_grid, int B, int H, int W,
                            int32_t* out_nbr, void* stream);
int gdmae_down_neighbor_maps(const int32_t* in_indices, int64_t N, const int32_t* in_rank_grid, int H, int W,
                             const int32_t* out_indices, int64_t No, const int32_t* out_rank_grid,
                             int32_t* out_nbr_down, int32_t* out_nbr_up, void* stream);
/* out (N, K*C) = rows of src (.., C) gathered through map (N, K); 0 where map < 0 */
int gdmae_gather_rows(const float* src, const int32_t* map, int64_t N, int K, int C, void* out,
                      int out_dtype /* 0 fp32, 1 bf16 */, void* stream);
/* dsrc (N, C) = sum_k dcol[tmap[i, mirror ? K-1-k : k], k*C:(k+1)*C]; dcol fp32 (0) or bf16 (1) */
int gdmae_gather_rows_transposed(const void* dcol, int in_dtype, const int32_t* tmap, int64_t N, int K, int C,
                                 int mirror, float* dsrc, void* stream);

/* ---- a10-a16 window tables -------------------------------------------------------------------
 * replaces sst_utils.get_window_coors (pcdet/models/model_utils/sst_utils.py:6-47),
 * get_inner_win_inds + drop_single_shift (pcdet/models/backbones_3d/spt_backbone.py:32-51),
 * make_continuous_inds / get_flat2win_inds (sst_utils.py:50-96) and the key-padding mask
 * (spt_backbone.py:184-194) for 8x8x1 windows. nW = B*(ceil(W/8)+1)*(ceil(H/8)+1); the
 * reference's batch_win_inds == 2 * win_of_token. */
size_t gdmae_window_table_workspace_bytes(int64_t n_windows);
int gdmae_window_table(const int32_t* indices, int64_t N, int B, int H, int W, int shifted, int32_t* win_of_token,
                       uint8_t* pos_of_token, int32_t* inner, int32_t* level, uint64_t* win_mask, int32_t* win_off,
                       int32_t* win_tok, int32_t* lvl_rank, int32_t* lvl_counts, int32_t* row_info /* (N,4): token,
                       first row of its window, one past its last row, in-window cell; 16-byte aligned */,
                       void* workspace, size_t ws_bytes, void* stream);

/* ---- a16-a18 SRA attention core --------------------------------------------------------------
 * replaces flat2window/window2flat (sst_utils.py:107-181), the per-level loop of
 * WindowAttention.forward (pcdet/models/model_utils/sst_basic_block.py:22-54) and
 * _scaled_cosine_attention (pcdet/models/model_utils/cosine_msa.py:114-176) incl. the -inf key mask.
 * qkv (N, 3d) = [x Wq^T | x Wk^T | x Wv^T] (no biases); lut (64, 2d) = pos_table [Wq;Wk]^T + [bq;bk];
 * bv (d, nullable) = value bias, added to the output (softmax rows sum to 1);
 * out (N, d) pre out-proj, fp32 or (io_bf16 = 1) bf16; lse (N, 8).  nhead == 8, d in {128, 256}.
 * Backward: out as written by forward; dqkv (N, 3d) fp32 or bf16 (io_bf16); dtau_sum (1) double, caller
 * zeroes, accumulates sum dS*S; work_D (N, 8) scratch. */
int gdmae_sra_attention_fwd(const float* qkv, const float* lut, const int32_t* row_info, int64_t N, int d, int nhead,
                            const float* tau, float tau_min, const float* bv, int io_bf16, void* out, float* lse,
                            void* stream);
/* same operator for bf16 q/k/v (qkv_bf16 (N,3d) bf16): QK^T and PV on the tensor cores (bf16 mma, fp32 accumulate,
 * softmax in fp32).  Flat-layout entry point: a re-layout kernel (+ LUT, per-head L2 norm, log2(e)/tau folded into q,
 * rows moved to CSR order) followed by the window-major kernel below; workspace gdmae_sra_tc_workspace_bytes(N, d);
 * other arguments and outputs as above */
size_t gdmae_sra_tc_workspace_bytes(int64_t N, int d);
int gdmae_sra_attention_fwd_tc(const void* qkv_bf16, const float* lut, const int32_t* row_info, const int32_t* bin_units,
                               int64_t N, int d, int nhead, const float* tau, float tau_min, const float* bv, int io_bf16,
                               void* out, float* lse, void* workspace, size_t ws_bytes, void* stream);
/* work units of the tensor-core kernels, built once per window table: bin_units (gdmae_sra_bin_units_bytes(N) bytes,
 * 16-byte aligned) <- for every 64-row bin of the CSR rows the packed (query tile, key range) units of the windows that
 * start in the bin (a run of whole small windows totalling <= 16 rows, or a 16-row chunk of a larger window), the
 * bin's first row / row count and a copy of its 128 row records (one 2304-byte block per bin = one bulk copy); followed, at int32 offset gdmae_sra_tok_info_offset(N), by tok_info (N): CSR row |
 * in-window cell << 26 per token (read by the in-projection epilogue, gdmae_tc_gemm mode 4) */
/* the kernels' internal waits are bounded; *out <- how many ran out since load (device sync; must be 0) */
int gdmae_sra_wait_timeouts(int* out);
size_t gdmae_sra_bin_units_bytes(int64_t N);
int64_t gdmae_sra_tok_info_offset(int64_t N);
int gdmae_sra_bin_units(const int32_t* row_info, int64_t N, int32_t* bin_units, void* stream);
/* tensor-core backward for bf16 tensors: qkv (N,3d), dout (N,d), dqkv (N,3d) all bf16; needs neither the forward
 * output nor the value bias; workspace as the forward */
int gdmae_sra_attention_bwd_tc(const void* qkv_bf16, const float* lut, const int32_t* row_info, const int32_t* bin_units,
                               int64_t N, int d, int nhead, const float* tau, float tau_min, const float* lse,
                               const void* dout_bf16, void* dqkv_bf16, double* dtau_sum, void* workspace, size_t ws_bytes,
                               void* stream);
/* the same kernels on the WINDOW-MAJOR layout the fused encoder layer uses (r2): qkvw[tensor][d/64 slices][N rows][64]
 * bf16 holding q^ * log2(e)/tau, k^ (positional term added, L2-normalised per head) and v with rows in CSR (window)
 * order - written by the in-projection GEMM's epilogue (gdmae_tc_gemm mode 4), so that a bin of windows is a rectangle
 * per tensor and arrives by TMA (one cp.async.bulk.tensor box per 16 rows) instead of per-row gathers.  out (N, d) token
 * order; lse (N, 8) by token, or (lse_by_row) columns 0..7 of lrr.  Backward: qkvdw = the same array with dO as a fourth
 * tensor (gdmae_tc_gemm mode 5); lrr (N, 24) fp32 per-row records by CSR row = lse | 1/|q| | 1/|k| (8 heads each);
 * dqkv (N, 3d) bf16 in token order. */
/* flat -> window-major: qkv_bf16 (N, 3d, nullable) -> tensors 0..2 of qkvdw and 1/|q|, 1/|k| in lrr; dout_bf16 (N, d,
 * nullable) -> tensor 3; lse_tok (N, 8, token order, nullable) -> lrr.  What modes 4 / 5 do inside the fused layer. */
int gdmae_sra_relayout(const void* qkv_bf16, const float* lut, const int32_t* row_info, const float* tau, float tau_min,
                       int64_t N, int d, const void* dout_bf16, const float* lse_tok, void* qkvdw, float* lrr, void* stream);
int gdmae_sra_fwd_win(const void* qkvw, const int32_t* bin_units, int64_t N, int d, const float* bv, int out_bf16, void* out,
                      float* lse, int lse_by_row, void* stream);
int gdmae_sra_bwd_win(const void* qkvdw, const float* lrr, const int32_t* bin_units, int64_t N, int d, const float* tau,
                      float tau_min, void* dqkv_bf16, double* dtau_sum, void* stream);
int gdmae_sra_attention_bwd(const float* qkv, const float* lut, const int32_t* row_info, int64_t N, int d, int nhead,
                            const float* tau, float tau_min, const float* bv, int io_bf16, const void* out,
                            const float* lse, const float* dout, void* dqkv, double* dtau_sum, float* work_D,
                            void* stream);

/* ---- a19 encoder-layer row kernels ------------------------------------------------------------
 * replace the residual + LayerNorm pairs and the GELU(linear1) of EncoderLayer.forward
 * (pcdet/models/model_utils/sst_basic_block.py:77-84), the bias-gradient reductions of its linears and
 * q = k = feat + pos (sst_basic_block.py:39-46).  d in {128,256}; parameter gradients are written
 * (accumulate=0) or added (accumulate=1).  *_bf16 outputs (nullable) are bf16 copies = GEMM operands of
 * the bf16 configuration.  workspace: gdmae_rowwise_workspace_bytes(max columns). */
size_t gdmae_rowwise_workspace_bytes(int max_cols);
int gdmae_add_layernorm_fwd(const float* x, const float* res, const float* bias, const float* gamma, const float* beta,
                            int64_t N, int d, float eps, float* y, void* y_bf16, float* mean, float* rstd, void* stream);
int gdmae_add_layernorm_bwd(const float* x, const float* res, const float* bias, const float* gamma, const float* mean,
                            const float* rstd, const float* dy, int64_t N, int d, float* dz, void* dz_bf16,
                            float* dgamma, float* dbeta, int accumulate, void* workspace, size_t ws_bytes, void* stream);
int gdmae_bias_gelu_fwd(const float* h, const float* bias, int64_t N, int C, float* out, void* out_bf16, void* stream);
int gdmae_bias_gelu_bwd(const float* h, const float* bias, const float* dg, int64_t N, int C, float* dh, void* dh_bf16,
                        float* dbias, int accumulate, void* workspace, size_t ws_bytes, void* stream);
int gdmae_colsum(const void* x, int dtype /* 0 fp32, 1 bf16 */, int64_t N, int ld, int col0, int C, float* out,
                 int accumulate, void* workspace, size_t ws_bytes, void* stream);
int gdmae_gather_add_rows(const float* x, const float* table, const uint8_t* idx, int64_t N, int C, float* out,
                          void* out_bf16, void* stream);

/* ---- a5/a6 pillar feature encoder MLP ----------------------------------------------------------------
 * One call = DynVFE's dvfe_mlps (Linear(K,64,no bias) -> BatchNorm1d(train) -> ReLU -> Linear(64,128,no bias) ->
 * BatchNorm1d(train) -> ReLU) followed by torch_scatter.scatter_max over the pillars
 * (pcdet/models/backbones_3d/vfe/dyn_vfe.py:105-111, network_utils.py:7-21), forward or backward.
 * x (Np,K) fp32 point features; seg_offsets/seg_points = pillar CSR from gdmae_dynvox (seg_points == NULL: the rows of x
 * are already in pillar order, pillar m owns rows [seg_offsets[m], seg_offsets[m+1]) - the per-pillar kernels then stream
 * contiguous rows and argmax holds row numbers in that order); "op" buffers are fp32
 * (gemm_mode 0/2) or bf16 (gemm_mode 1).  Backward writes (accumulate=0) or adds to (accumulate=1) d_*. */
typedef struct gdmae_vfe_mlp_args {
  int64_t Np, M;
  int K, C1, C2;          /* C1 == 64, C2 == 128, K <= 16 */
  int gemm_mode;          /* as gdmae_gemm's ab_dtype */
  int accumulate;
  float eps, momentum;
  const float* x;
  const int32_t *seg_offsets, *seg_points;
  const float *W1, *g1, *b1, *g2, *b2;      /* W1 (64,K); BatchNorm weights / biases */
  const void* W2_g;                          /* (128,64) op */
  float *running_mean1, *running_var1, *running_mean2, *running_var2;   /* nullable */
  /* written by forward, read by backward */
  void* h1;               /* (Np,64) op */
  void* y2;               /* (Np,128) op */
  float *mean1, *rstd1, *mean2, *rstd2;
  float* out;             /* (M,128) pillar features */
  uint8_t* argmax;        /* (M,128) position of each maximum inside its pillar's segment, saturated at 255 (then recomputed) */
  /* backward */
  const float* dout;      /* (M,128) */
  void* dy2;              /* (Np,128) op scratch */
  void* dh1;              /* (Np,64) op scratch */
  float *tmp_dbeta1, *tmp_dgamma1, *tmp_dbeta2, *tmp_dgamma2;   /* (64),(64),(128),(128) scratch */
  float *d_W1, *d_g1, *d_b1, *d_W2, *d_g2, *d_b2;
  double* moments;        /* (K + K(K+1)/2 <= 152) first / second moments of x: written by forward, read by backward */
  void* ws;               /* gdmae_vfe_mlp_workspace_bytes(K) */
  size_t ws_bytes;
  void* stream;
} gdmae_vfe_mlp_args;
size_t gdmae_vfe_mlp_workspace_bytes(int K);
int gdmae_vfe_mlp_fwd(const gdmae_vfe_mlp_args* args);
int gdmae_vfe_mlp_bwd(const gdmae_vfe_mlp_args* args);

/* ---- a13-a19 encoder-layer executor -------------------------------------------------------------
 * One call = the whole forward (or the whole backward) of an SST EncoderLayer
 * (pcdet/models/model_utils/sst_basic_block.py:60-92 with WindowAttention :22-54 and the
 * CosineMultiheadAttention projections, cosine_msa.py:57-62,380-431): x -> LN2(x1 + W2 gelu(W1 x1 + b1) + b2),
 * x1 = LN1(x + Wo SRA(x) + bo).  All buffers belong to the caller; "op" pointers are in the GEMM
 * operand dtype (fp32 for gemm_mode 0/2, bf16 for gemm_mode 1).  Weights are the torch layouts:
 * w_in (3d,d), w_o (d,d), w1 (dff,d), w2 (d,dff).  Backward writes (accumulate=0) or adds to
 * (accumulate=1) the parameter gradients d_*; dx receives the input gradient. */
typedef struct gdmae_encoder_layer_args {
  int64_t N;
  int d, dff, nhead;
  int gemm_mode;        /* 0: fp32 operands, TF32 math; 1: bf16 operands; 2: fp32 operands, fp32 math */
  int sra_tensor_cores; /* SRA kernels: 0 fp32 SIMT (fp32 qkv), 1 bf16 tensor cores (bf16 qkv; gemm_mode 1 only) */
  int accumulate;
  float tau_min, eps;
  /* inputs */
  const float* x;            /* (N,d) */
  const void* xg_in;         /* (N,d) op copy of x handed over by the producer, or NULL */
  const float* pos_table;    /* (64,d) */
  const int32_t* row_info;   /* (N,4) from gdmae_window_table */
  const int32_t* bin_units;  /* from gdmae_sra_bin_units; required when sra_tensor_cores */
  const uint8_t* pos_of_token;
  /* parameters: fp32 masters and op-dtype copies of the four weights (same pointers when fp32) */
  const float *w_in, *b_in, *tau, *w_o, *b_o, *g1, *be1, *w1, *b1, *w2, *b2, *g2, *be2;
  const void *w_in_g, *w_o_g, *w1_g, *w2_g;
  /* activations written by forward and read by backward */
  void* xg;                  /* (N,d) op; used when xg_in == NULL and gemm_mode == 1 */
  void* qkv;                 /* (N,3d) fp32; when sra_tensor_cores: (4, d/64, N, 64) bf16 window-major q^ | k^ | v | dO (gdmae_sra_fwd_win) */
  float* lut;                /* (64,2d) */
  void* o;                   /* (N,d) op */
  float* lse;                /* (N,8); when sra_tensor_cores: (N,24) per-row records lse | 1/|q| | 1/|k| by CSR row */
  void* a;                   /* (N,d) out-projection output: fp32, bf16 when gemm_mode == 1 */
  float* x1;                 /* (N,d) */
  void* x1g;                 /* (N,d) op (bf16 mode only) */
  float *mean1, *rstd1;      /* (N) */
  void* h;                   /* (N,dff) FFN pre-activation (fp32); when gemm_mode == 1: bf16 gelu'(pre-activation + b1) */
  void* g;                   /* (N,dff) op */
  void* f;                   /* (N,d) FFN output: fp32, bf16 when gemm_mode == 1 */
  float *mean2, *rstd2;      /* (N) */
  float* x2;                 /* (N,d) output */
  void* x2g;                 /* (N,d) op copy of the output (bf16 mode only) */
  /* backward */
  const float* dy;           /* (N,d) gradient w.r.t. x2 */
  float* dx;                 /* (N,d) gradient w.r.t. x */
  float *d_w_in, *d_b_in, *d_tau, *d_w_o, *d_b_o, *d_g1, *d_be1, *d_w1, *d_b1, *d_w2, *d_b2, *d_g2, *d_be2;
  void* ws;                  /* gdmae_encoder_layer_bwd_workspace_bytes(N, d, dff) */
  size_t ws_bytes;
  void* stream;
} gdmae_encoder_layer_args;
int gdmae_encoder_layer_fwd(const gdmae_encoder_layer_args* args);
size_t gdmae_encoder_layer_bwd_workspace_bytes(int64_t N, int d, int dff);
int gdmae_encoder_layer_bwd(const gdmae_encoder_layer_args* args);

/* ---- plain dense GEMM through cuBLAS (library GEMM) ------------------------------------------------
 * row-major C (M,N) = op(A) op(B) + beta*C; A/B fp32 (TF32 math, ab_dtype 0) or bf16 (1); C fp32 or bf16.
 * Owns one cublasLt handle + 64 MB scratch per device (created on first use) and caches the selected
 * algorithm per shape bucket: the heuristic costs ~200 us of host time per call otherwise.
 * replaces F.linear of the projections / FFN / sparse-conv GEMMs (cosine_msa.py:57-62,431, sst_basic_block.py:81). */
int gdmae_gemm(int transa, int transb, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B,
               int64_t ldb, int ab_dtype, void* C, int64_t ldc, int c_dtype /* 0 fp32, 1 bf16 */, float beta, void* stream);

/* ---- own Blackwell GEMM: tcgen05.mma + TMEM accumulators + TMA, fused epilogues (csrc/tc_gemm.cu) -----------------
 * row-major C (M,N) = op(A) (M,K) op(B) (K,N), bf16 operands, fp32 accumulation; transa: A stored (K,M); transb: B
 * stored (N,K).  N a multiple of 64, leading dimensions multiples of 8 elements, 16-byte aligned pointers.
 * c_dtype 0 fp32 (beta 0 or 1) / 1 bf16.  split_k_atomic: the K range is split across the SMs and partial tiles are
 * added to C with fp32 reductions (weight gradients, K = tokens; beta 0 zero-fills C first, needs ldc == N).
 * epilogue (nullable = mode 0):
 *   mode 1  c2 = gelu(acc + bias), C = gelu'(acc + bias) (both bf16; C is what mode 3 reads) - linear1 + GELU, sst_basic_block.py:81
 *   mode 2  z = acc + bias + res; y32/y16 = LayerNorm(z) * gamma + beta_ln, mean, rstd; N in {128, 256}; C (nullable, bf16) = acc
 *                                                                                     - out_proj / linear2 + residual + norm, :78-83
 *   mode 3  acc = gradient w.r.t. gelu(h + bias): C = acc * h16 (bf16; h16 = the derivative saved by mode 1) and colsum (N, fp32) += its column sums
 *                                                                                     - backward of linear1's bias + GELU
 *   mode 4  in-projection of the attention, N = 3d: q / k tiles += lut[cell(token)], L2-normalised per head (q also
 *           * log2(e)/max(tau, tau_min)), 1/|q|, 1/|k| -> lrr; q^, k^, v rows (bf16) go to C = the window-major array
 *           [tensor][d/64][M][64] at the token's CSR row (gdmae_sra_fwd_win's operand)       - cosine_msa.py:57-62,114-140
 *   mode 5  C = window-major array: bf16 rows of the (M, N) result go to planes plane0 + column/64 at the token's CSR row
 *           (dO of the attention, plane0 = 3 d/64)
 * replaces F.linear of cosine_msa.py:57-62,431 / sst_basic_block.py:77-84 and the spconv / deblock / VFE GEMMs. */
typedef struct gdmae_tc_epilogue {
  int mode;
  const float* bias;         /* (N) */
  void* c2;                  /* mode 1: (M, ldc2) bf16 */
  int64_t ldc2;
  const float* res;          /* mode 2: (M, N) fp32 residual */
  const float* gamma;        /* mode 2: (N) */
  const float* beta_ln;      /* mode 2: (N) */
  float eps;
  float* y32;                /* mode 2: (M, N) fp32 */
  void* y16;                 /* mode 2: (M, N) bf16, nullable */
  float* mean;               /* mode 2: (M) */
  float* rstd;               /* mode 2: (M) */
  const void* h16;           /* mode 3: (M, ldh) bf16 gelu'(pre-activation + bias) saved by mode 1 (its C) */
  int64_t ldh;
  float* colsum;             /* mode 3: (N) fp32, accumulated into */
  const int32_t* tok_info;   /* modes 4, 5: (M) CSR row | cell << 26 per token (gdmae_sra_bin_units) */
  const float* lut;          /* mode 4: (64, 2d) positional LUT incl. the q / k biases */
  const float* tau;          /* mode 4: (1) temperature parameter */
  float tau_min;
  float* lrr;                /* mode 4: (M, 24) per-row records: 1/|q| -> columns 8..15, 1/|k| -> 16..23 */
  int plane0;                /* mode 5: first destination plane */
} gdmae_tc_epilogue;
int gdmae_tc_gemm(int transa, int transb, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B,
                  int64_t ldb, void* C, int64_t ldc, int c_dtype, float beta, int split_k_atomic,
                  const gdmae_tc_epilogue* epilogue, void* stream);
int gdmae_tc_gemm_timeouts(int* out);
/* weight gradient of the decoder's dense 3x3 convolution (spt_backbone_mae.py:45-49, Conv2d(384, 128, 3, padding=1)) on the
 * same tcgen05 / TMA machinery: dW (c_out, 3, 3, c_in) fp32 (+)= sum_{b,y,x} dy[b,y,x,co] in[b,y+ky-1,x+kx-1,ci];
 * dy (B,Y,X,c_out), in (B,Y,X,c_in) NHWC bf16; c_in == 384, c_out == 128.  The shifted operand of every tap is a TMA box of
 * the input map (padding = out-of-bounds zero fill), K = pixels is split across the SMs, partial tiles reach dW by
 * TMA reduce-add.  Replaces cuDNN's implicit-GEMM wgrad (r2 profile: 10.8 GB of DRAM traffic for 1.8 GB of operands). */
int gdmae_conv3x3_wgrad(const void* dy_bf16, const void* in_bf16, int B, int Y, int X, int c_in, int c_out, float* dW,
                        int accumulate, void* stream);

/* ---- a5/a9/a21/a22 training-mode BatchNorm (+ReLU) over (N, C) rows ------------------------------
 * replaces norm_fn + nn.ReLU of post_act_block (pcdet/utils/spconv_utils.py:50-54), of make_fc_layers
 * (pcdet/models/model_utils/network_utils.py:7-21) and BatchNorm2d + ReLU of the decoder deblocks evaluated
 * on their sparse rows (spt_backbone_mae.py:31-43).  `count` >= N rows enter the statistics (missing rows
 * are zeros); running buffers (nullable) get the momentum / unbiased-variance update. */
size_t gdmae_batchnorm_workspace_bytes(int C);
int gdmae_batchnorm_relu_fwd(const float* y, const float* gamma, const float* beta, int64_t N, int C, double count,
                             float eps, float momentum, int relu, float* out, float* mean, float* rstd,
                             float* running_mean, float* running_var, void* workspace, size_t ws_bytes, void* stream);
int gdmae_batchnorm_relu_bwd(const float* y, const float* beta, const float* dout, const float* gamma, const float* mean,
                             const float* rstd, int64_t N, int C, double count, int relu, const float* extra_dbeta,
                             const float* extra_dgamma, float* dy /* nullable */, void* dy_bf16 /* nullable: bf16 copy of dy */,
                             float* dgamma, float* dbeta, void* workspace, size_t ws_bytes, void* stream);
/* the same pair with typed tensors (dtype 0 = fp32, 1 = bf16; C % 8 == 0): the decoder deblocks of the bf16 configuration
 * (spt_backbone_mae.py:31-44) keep their output rows and the rows of the map's gradient as bf16 - the map itself is bf16 */
int gdmae_batchnorm_relu_fwd_t(const void* y, int y_dtype, const float* gamma, const float* beta, int64_t N, int C, double count,
                               float eps, float momentum, int relu, void* out, int out_dtype, float* mean, float* rstd,
                               float* running_mean, float* running_var, void* workspace, size_t ws_bytes, void* stream);
int gdmae_batchnorm_relu_bwd_t(const void* y, int y_dtype, const float* beta, const void* dout, int dout_dtype, const float* gamma,
                               const float* mean, const float* rstd, int64_t N, int C, double count, int relu,
                               const float* extra_dbeta, const float* extra_dgamma, void* dy, int dy_dtype, float* dgamma,
                               float* dbeta, void* workspace, size_t ws_bytes, void* stream);

/* ---- a22/a23 decoder dense fill and pillar gather --------------------------------------------
 * replaces SparseConvTensor.dense() + ConvTranspose2d(k=s) + BatchNorm2d + ReLU + torch.cat
 * (spt_backbone_mae.py:125-132) once the per-site GEMM/BN is done on the sparse rows, and the
 * gather at all pillars (spt_backbone_mae.py:141-143).  rows/bg/rank_grids/drows are HOST
 * arrays of 3 device pointers; out (B, Y, X, 3*Cs) NHWC. */
int gdmae_dense_fill(const void* const* rows, int rows_dtype /* 0 fp32, 1 bf16 */, const float* const* bg,
                     const int32_t* const* rank_grids, const int* strides, int B, int Y, int X, int Cs, void* out,
                     int out_dtype /* 0 fp32, 1 bf16 */, void* stream);
/* one pass over dout: covered (cell, scale) packets go to drows[s] (n_sites[s] * k_s^2, Cs), the others are summed into dbg */
int gdmae_dense_fill_bwd(const void* dout, int dtype, const int32_t* const* rank_grids, const int64_t* n_sites,
                         const int* strides, int B, int Y, int X, int Cs, void* const* drows, int drows_dtype,
                         float* dbg /* (3*Cs) */, void* stream);
int gdmae_gather_nhwc(const void* src, int dtype, const int64_t* voxel_coords, int64_t M, int Y, int X, int C, float* out,
                      void* stream);
int gdmae_scatter_nhwc(const float* dout, const int64_t* voxel_coords, int64_t M, int Y, int X, int C, void* dsrc,
                       int dtype, void* stream);

/* ---- a21/a22 decoder tail -------------------------------------------------------------------------
 * BatchNorm2d (training statistics over all B*Y*X cells) + ReLU of decoder_conv_out
 * (pcdet/models/backbones_3d/spt_backbone_mae.py:52-57) evaluated at the pillar cells only, where the MAE head
 * gathers it (spt_backbone_mae.py:141-143): y = conv output NHWC fp32 (dtype 0) or bf16 (1); out (M,C) fp32.
 * Backward: dy (B,Y,X,C) in y's dtype for every cell, dgamma/dbeta (C); cell2pillar from gdmae_dynvox.
 * workspace: gdmae_batchnorm_workspace_bytes(C). */
int gdmae_decoder_tail_fwd(const void* y, int dtype, int B, int Y, int X, int C, const int64_t* voxel_coords, int64_t M,
                           const float* gamma, const float* beta, float eps, float momentum, float* out, float* mean,
                           float* rstd, float* running_mean, float* running_var, void* workspace, size_t ws_bytes, void* stream);
int gdmae_decoder_tail_bwd(const void* y, int dtype, int B, int Y, int X, int C, const int64_t* voxel_coords,
                           const int32_t* cell2pillar, int64_t M, const float* out, const float* dout, const float* gamma,
                           const float* mean, const float* rstd, void* dy, float* dgamma, float* dbeta, void* workspace,
                           size_t ws_bytes, void* stream);

/* ---- a24-a26 chamfer head --------------------------------------------------------------------
 * gdmae_group_points_centered replaces sst_ops_utils.group_inner_inds + points[group_inds]
 * (pcdet/ops/sst_ops/sst_ops_utils.py:15-27) fused with get_voxel_centers
 * (pcdet/utils/common_utils.py:130-145) and the subtraction at spt_backbone_mae.py:67-72.
 * gdmae_chamfer_fwd replaces pytorch3d.loss.chamfer_distance(pred, gt, weights=mask)
 * (spt_backbone_mae.py:88): per_item (N) and dpred (N, P1, 3); loss = sum(per_item)/sum(w). */
int gdmae_group_points_centered(const float* points, int n_cols, const int32_t* seg_offsets, const int32_t* seg_points,
                                const int64_t* voxel_coords, const float* pc_range, const float* voxel, int64_t M, int K,
                                float* out_gt, void* stream);
int gdmae_chamfer_fwd(const float* pred, const float* gt, const float* weights, int64_t N, int P1, int P2,
                      float* per_item, float* dpred, void* stream);

/* ---- optimizer ---------------------------------------------------------------------------------
 * replaces clip_grad_norm_ (tools/train_utils/train_utils.py:52) and OptimWrapper.step + Adam
 * (tools/train_utils/optimization/fastai_optim.py:135-152) on one flat fp32 bucket. */
int gdmae_grad_sumsq(const float* grads, int64_t n, double* out, void* stream);
int gdmae_adam_onecycle_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n_opt,
                             const double* sumsq, float clip, float decay, float mom, float beta2, float eps,
                             float step_size, float bc2_sqrt, float grad_scale, void* stream);

/* ---- input side: world augmentation + shuffle of a collated batch -------------------------
 * replaces DataAugmentor.random_world_flip / random_world_rotation / random_world_scaling
 * (pcdet/datasets/augmentor/data_augmentor.py:55-143, common_utils.rotate_points_along_z common_utils.py:99-121) and
 * DataProcessor.shuffle_points (pcdet/datasets/processor/data_processor.py:92-102) applied to the whole batch on the device.
 * points / out (N, n_cols) fp32, column 0 = frame index, 1..3 = x, y, z; params (B, 6) device fp32 per frame:
 * flip_x (negates y), flip_y (negates x), cos, sin, scale, 0; src_index (N) int32 nullable: out[i] <- T(points[src_index[i]]). */
int gdmae_world_augment(const float* points, int64_t N, int n_cols, const float* params, int B, const int32_t* src_index,
                        float* out, void* stream);

/* ---- rotated BEV overlap / IoU / NMS (csrc/iou3d_nms.cu) ------------------------------------------
 * boxes (N,7) fp32 [x, y, z, dx, dy, dz, heading].
 * gdmae_boxes_overlap_bev / gdmae_boxes_iou_bev: ans (na, nb) fp32, replace iou3d_nms_cuda.boxes_overlap_bev_gpu /
 *   boxes_iou_bev_gpu (pcdet/ops/iou3d_nms/src/iou3d_nms.cpp:47-87, kernels iou3d_nms_kernel.cu:238-266).
 * gdmae_nms_bev / gdmae_nms_normal: boxes sorted by descending score; keep (n) int64 and num_out (1) int32 on the DEVICE,
 *   workspace = gdmae_nms_workspace_bytes(n) (the suppression matrix).  Replace iou3d_nms_cuda.nms_gpu / nms_normal_gpu
 *   (iou3d_nms.cpp:88-187: cudaMalloc + mask copy to the host + host loop; here everything stays on the device).
 * gdmae_boxes_iou_bev_cpu: HOST pointers, the counterpart of iou3d_nms_cuda.boxes_iou_bev_cpu (iou3d_cpu.cpp:222-252). */
int gdmae_boxes_overlap_bev(const float* boxes_a, int na, const float* boxes_b, int nb, float* ans_overlap, void* stream);
int gdmae_boxes_iou_bev(const float* boxes_a, int na, const float* boxes_b, int nb, float* ans_iou, void* stream);
size_t gdmae_nms_workspace_bytes(int n);
int gdmae_nms_bev(const float* boxes, int n, float thresh, void* workspace, size_t ws_bytes, int64_t* keep, int* num_out, void* stream);
int gdmae_nms_normal(const float* boxes, int n, float thresh, void* workspace, size_t ws_bytes, int64_t* keep, int* num_out,
                     void* stream);
int gdmae_boxes_iou_bev_cpu(const float* boxes_a, int na, const float* boxes_b, int nb, float* ans_iou);

/* ---- CenterHead targets and heat-map loss on the device (csrc/center_head.cu) ----------------------
 * gdmae_center_assign_targets: replaces the CPU Python loop of CenterHead.assign_targets / assign_target_of_single_head
 *   (pcdet/models/dense_heads/center_head.py:105-231) and centernet_utils.gaussian_radius / draw_gaussian_to_heatmap
 *   (pcdet/models/model_utils/centernet_utils.py:9-72) for ONE head.  gt_boxes (B, M, 8) fp32, last column = 1-based class id
 *   (0 = padding); class_map (n_class_total + 1) int32 on the device: class id -> 1-based id inside the head, 0 = not in it;
 *   range_xy_voxel_xy = HOST {x_min, y_min, voxel_x, voxel_y}.  Outputs (zero-filled here): heatmap (B, C, H, W),
 *   target_boxes (B, max_objs, 8) [dx, dy, z, log dims, cos, sin], iou_boxes (B, max_objs, 7), inds / mask (B, max_objs) int64.
 * gdmae_center_focal_loss: FocalLossCenterNet on clamp(sigmoid(logits), 1e-4, 1 - 1e-4) (pcdet/utils/loss_utils.py:273-309,
 *   center_head.py:233-235): sums3 = {sum of positive terms, sum of negative terms, number of positives} (device, double),
 *   grad_raw (n) = d(pos + neg)/dlogits. */
int gdmae_center_assign_targets(const float* gt_boxes, int B, int M, const int* class_map, int n_class_total, int C, int H, int W,
                                int max_objs, int min_radius, int stride, const float* range_xy_voxel_xy, float overlap, float* heatmap,
                                float* target_boxes, float* iou_boxes, int64_t* inds, int64_t* mask, void* stream);
int gdmae_center_focal_loss(const float* logits, const float* gt, int64_t n, float* grad_raw, double* sums3, void* stream);

/* ---- (modulated) deformable convolution sampling kernels (csrc/dcn.cu) -------------------------------
 * The building blocks behind the reference's five pybind entry points (pcdet/ops/dcn/src/deform_conv_cuda.cpp:687-701:
 * deform_conv_forward_cuda, deform_conv_backward_input_cuda, deform_conv_backward_parameters_cuda,
 * modulated_deform_conv_cuda_forward, modulated_deform_conv_cuda_backward), which gd-mae_b200/pcdet/ops/dcn/deform_conv.py
 * rebuilds under the same names on these launchers plus library GEMMs.  Kernels replaced: deform_conv_cuda_kernel.cu:84-470
 * and 570-866.  input (B,C,H,W), offset (B, dg*2*kh*kw, Ho, Wo), mask (B, dg*kh*kw, Ho, Wo) or NULL (= plain deformable conv),
 * columns (B, C*kh*kw, Ho*Wo); all fp32.  geom: HOST int[13] {B,C,H,W,kh,kw,pad_h,pad_w,stride_h,stride_w,dil_h,dil_w,dg}. */
int gdmae_deform_im2col(const float* input, const float* offset, const float* mask, const int* geom, float* columns, void* stream);
int gdmae_deform_col2im(const float* grad_columns, const float* offset, const float* mask, const int* geom, float* grad_input, void* stream);
int gdmae_deform_col2im_coord(const float* grad_columns, const float* input, const float* offset, const float* mask, const int* geom,
                              float* grad_offset, float* grad_mask, void* stream);

#ifdef __cplusplus
}
#endif
#endif
