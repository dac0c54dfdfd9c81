"""Python-side operators over the C ABI (include/gdmae_b200.h): thin launch wrappers and the
``torch.autograd.Function``s that wire the hand-written forward/backward kernels into autograd.

Nothing here computes on the CPU and nothing falls back to PyTorch ops for the kernels' work.
"""
import ctypes

import torch
from torch.amp import custom_bwd, custom_fwd

from . import _lib as L

# the kernels compute in fp32: under torch.autocast (bf16 GEMM/conv mode) inputs are cast back up
_fwd = custom_fwd(device_type="cuda", cast_inputs=torch.float32)
_bwd = custom_bwd(device_type="cuda")

I32, I64, F32 = torch.int32, torch.int64, torch.float32
_DT = {torch.float32: 0, torch.bfloat16: 1}  # dtype codes of the C ABI


def _dev(t):
    if not t.is_cuda:
        raise L.GdmaeError("gd-mae_b200 operators need CUDA tensors (no CPU path exists)")
    return t.device


# ----------------------------------------------------------------------------- a1/a2 voxelisation
class PillarSet:
    """Result of dynamic voxelisation (all device tensors, exact sizes)."""
    __slots__ = ("points", "point_coords", "inverse", "voxel_coords", "cell2pillar", "seg_offsets", "seg_points",
                 "batch_offsets", "batch_offsets_dev", "n_points", "n_pillars", "grid", "batch_size")


def dynamic_voxelize(points, pc_range, voxel_size, grid_xyz, batch_size):
    """common_utils.get_in_range_mask + unique(dim=0) of DynVFE.forward (dyn_vfe.py:60-68).
    ONE host sync (reads the 4+B+1 int32 counts) to size the outputs exactly."""
    dev = _dev(points)
    points = points.contiguous().float()
    n_in, n_cols = points.shape
    X, Y, Z = [int(g) for g in grid_xyz]
    n_cells = batch_size * X * Y * Z
    cap_m = max(min(n_in, n_cells), 1)
    out_points = torch.empty((max(n_in, 1), n_cols), dtype=F32, device=dev)
    out_coords = torch.empty((max(n_in, 1), 4), dtype=I64, device=dev)
    out_inverse = torch.empty((max(n_in, 1),), dtype=I64, device=dev)
    out_vcoords = torch.empty((cap_m, 4), dtype=I64, device=dev)
    cell2pillar = torch.empty((n_cells,), dtype=I32, device=dev)
    seg_off = torch.empty((cap_m + 1,), dtype=I32, device=dev)
    seg_pts = torch.empty((max(n_in, 1),), dtype=I32, device=dev)
    counts = torch.empty((4 + batch_size + 1,), dtype=I32, device=dev)
    lib = L.lib()
    nbytes = lib.gdmae_dynvox_workspace_bytes(L.i64(n_in), L.i64(n_cells))
    ws = L.workspace(nbytes, dev)
    L.check(lib.gdmae_dynvox(L.P(points), L.i64(n_in), n_cols, L.farr(pc_range), L.farr(voxel_size), L.iarr([X, Y, Z]),
                             batch_size, L.P(out_points), L.P(out_coords), L.P(out_inverse), L.P(out_vcoords),
                             L.P(cell2pillar), L.P(seg_off), L.P(seg_pts), L.P(counts), L.P(ws),
                             ctypes.c_size_t(ws.numel()), L.stream()), "gdmae_dynvox")
    host = counts.cpu().tolist()  # the single sync of DynVFE.forward
    if host[2] != 0:
        raise L.GdmaeError("dynamic_voxelize: a point carries a frame index outside [0, batch_size)")
    ps = PillarSet()
    ps.n_points, ps.n_pillars = host[0], host[1]
    ps.batch_offsets = host[4:4 + batch_size + 1]
    ps.batch_offsets_dev = counts[4:4 + batch_size + 1]
    ps.points = out_points[:ps.n_points]
    ps.point_coords = out_coords[:ps.n_points]
    ps.inverse = out_inverse[:ps.n_points]
    ps.voxel_coords = out_vcoords[:ps.n_pillars]
    ps.cell2pillar = cell2pillar
    ps.seg_offsets = seg_off[:ps.n_pillars + 1]
    ps.seg_points = seg_pts[:ps.n_points]
    ps.grid, ps.batch_size = (X, Y, Z), batch_size
    return ps


def segment_mean(src, col0, C, seg_offsets, seg_points, M):
    """torch_scatter.scatter(src[:, col0:col0+C], inverse, reduce='mean') (dyn_vfe.py:81)."""
    out = torch.empty((M, C), dtype=F32, device=_dev(src))
    L.check(L.lib().gdmae_segment_mean(L.P(src), src.shape[1], col0, C, L.P(seg_offsets), L.P(seg_points), L.i64(M),
                                       L.P(out), L.stream()), "gdmae_segment_mean")
    return out


def vfe_point_features(ps, mean, pc_range, voxel_size):
    """dyn_vfe.py:86-105."""
    n_cols = ps.points.shape[1]
    out = torch.empty((ps.n_points, n_cols - 1 + 6), dtype=F32, device=ps.points.device)
    L.check(L.lib().gdmae_vfe_point_features(L.P(ps.points), L.P(ps.point_coords), L.P(ps.inverse), L.P(mean),
                                             mean.shape[1], L.i64(ps.n_points), n_cols, L.farr(pc_range),
                                             L.farr(voxel_size), L.P(out), L.stream()), "gdmae_vfe_point_features")
    return out


class SegmentMax(torch.autograd.Function):
    """torch_scatter.scatter_max(x, inverse, dim=0)[0] over the pillar CSR (dyn_vfe.py:109-111)."""

    @staticmethod
    @_fwd
    def forward(ctx, x, seg_offsets, seg_points, M):
        x = x.contiguous()
        out = torch.empty((M, x.shape[1]), dtype=F32, device=_dev(x))
        arg = torch.empty((M, x.shape[1]), dtype=I32, device=x.device)
        # algorithmic bytes (op boundary a6): Np*C*4 + Np*4 + M*C*4
        with L.timed("segment_max_fwd", x.shape[0] * x.shape[1] * 4 + x.shape[0] * 4 + M * x.shape[1] * 4):
            L.check(L.lib().gdmae_segment_max_fwd(L.P(x), x.shape[1], L.P(seg_offsets), L.P(seg_points), L.i64(M), L.P(out),
                                                  L.P(arg), L.stream()), "gdmae_segment_max_fwd")
        ctx.save_for_backward(arg, seg_offsets, seg_points)
        ctx.shape = x.shape
        return out

    @staticmethod
    @_bwd
    def backward(ctx, dout):
        arg, seg_offsets, seg_points = ctx.saved_tensors
        dout = dout.contiguous()
        dx = torch.empty(ctx.shape, dtype=F32, device=dout.device)
        L.check(L.lib().gdmae_segment_max_bwd(L.P(dout), L.P(arg), ctx.shape[1], L.P(seg_offsets), L.P(seg_points),
                                              L.i64(arg.shape[0]), L.P(dx), L.stream()), "gdmae_segment_max_bwd")
        return dx, None, None, None


# ----------------------------------------------------------------------------- a7 mask
def random_mask(noise, batch_offsets_dev, batch_size, mask_ratio):
    """common_utils.random_masking per frame (common_utils.py:49-63). 0 = visible, 1 = masked."""
    dev = _dev(noise)
    M = noise.shape[0]
    out = torch.empty((M,), dtype=F32, device=dev)
    lib = L.lib()
    ws = L.workspace(lib.gdmae_random_mask_workspace_bytes(L.i64(M)), dev)
    L.check(lib.gdmae_random_mask(L.P(noise.contiguous().float()), L.i64(M), L.P(batch_offsets_dev), batch_size,
                                  ctypes.c_double(1 - mask_ratio), L.P(out), L.P(ws), ctypes.c_size_t(ws.numel()),
                                  L.stream()), "gdmae_random_mask")
    return out


# ----------------------------------------------------------------------------- sst_ops
def ingroup_inds(group_inds, out_inds=None):
    dev = _dev(group_inds)
    N = group_inds.shape[0]
    if out_inds is None:
        out_inds = torch.empty((N,), dtype=I64, device=dev)
    lib = L.lib()
    ws = L.workspace(lib.gdmae_ingroup_inds_workspace_bytes(L.i64(N)), dev)
    L.check(lib.gdmae_ingroup_inds(L.P(group_inds), L.i64(N), L.P(out_inds), L.P(ws), ctypes.c_size_t(ws.numel()),
                                   L.stream()), "gdmae_ingroup_inds")
    return out_inds


def group_inner_inds(inverse_inds, M, K, out=None):
    dev = _dev(inverse_inds)
    if out is None:
        out = torch.empty((M, K), dtype=I64, device=dev)
    lib = L.lib()
    Np = inverse_inds.shape[0]
    ws = L.workspace(lib.gdmae_group_inner_inds_workspace_bytes(L.i64(Np), L.i64(M)), dev)
    L.check(lib.gdmae_group_inner_inds(L.P(inverse_inds), L.i64(Np), L.i64(M), K, L.P(out), L.P(ws),
                                       ctypes.c_size_t(ws.numel()), L.stream()), "gdmae_group_inner_inds")
    return out


def group_inner_inds_csr(seg_offsets, seg_points, M, K):
    out = torch.empty((M, K), dtype=I64, device=_dev(seg_offsets))
    L.check(L.lib().gdmae_group_inner_inds_csr(L.P(seg_offsets), L.P(seg_points), L.i64(M), K, L.P(out), L.stream()),
            "gdmae_group_inner_inds_csr")
    return out


# ----------------------------------------------------------------------------- sparse structure
def visible_sites(voxel_coords, mask, n_visible, B, Y, X):
    """-> vis_idx (N1) int32 pillar rows, indices (N1,3) int32, rank_grid (B*Y*X) int32, count (1) int32 (device)."""
    dev = _dev(voxel_coords)
    M = voxel_coords.shape[0]
    vis_idx = torch.empty((max(n_visible, 1),), dtype=I32, device=dev)
    indices = torch.empty((max(n_visible, 1), 3), dtype=I32, device=dev)
    grid = torch.empty((B * Y * X,), dtype=I32, device=dev)
    count = torch.empty((1,), dtype=I32, device=dev)
    lib = L.lib()
    ws = L.workspace(lib.gdmae_visible_sites_workspace_bytes(L.i64(M)), dev)
    L.check(lib.gdmae_visible_sites(L.P(voxel_coords), L.P(mask), L.i64(M), B, Y, X, L.P(vis_idx), L.P(indices), L.P(grid),
                                    L.P(count), L.P(ws), ctypes.c_size_t(ws.numel()), L.stream()), "gdmae_visible_sites")
    return vis_idx[:n_visible], indices[:n_visible], grid, count


def build_rank_grid(indices, B, H, W):
    grid = torch.empty((B * H * W,), dtype=I32, device=_dev(indices))
    L.check(L.lib().gdmae_build_rank_grid(L.P(indices), L.i64(indices.shape[0]), B, H, W, L.P(grid), L.stream()),
            "gdmae_build_rank_grid")
    return grid


def down_sites(in_indices, n_in_rows, B, H, W):
    """SparseConv2d(3, s2, p1) output sites.  ``in_indices`` may be a capacity buffer: rows whose frame
    index is -1 are skipped, and the returned capacity buffer is pre-filled with -1 for the same reason
    (lets two pyramid levels be planned before the single host sync that reads their counts).
    -> (indices (cap,3), rank_grid, count (1) device, Ho, Wo)."""
    dev = _dev(in_indices)
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    cap = max(min(4 * n_in_rows, B * Ho * Wo), 1)
    out_idx = torch.full((cap, 3), -1, dtype=I32, device=dev)
    grid = torch.empty((B * Ho * Wo,), dtype=I32, device=dev)
    count = torch.empty((1,), dtype=I32, device=dev)
    lib = L.lib()
    ws = L.workspace(lib.gdmae_down_sites_workspace_bytes(L.i64(B * Ho * Wo)), dev)
    L.check(lib.gdmae_down_sites(L.P(in_indices), L.i64(n_in_rows), B, H, W, L.P(out_idx), L.P(grid), L.P(count), L.P(ws),
                                 ctypes.c_size_t(ws.numel()), L.stream()), "gdmae_down_sites")
    return out_idx, grid, count, Ho, Wo


def subm_neighbor_map(indices, rank_grid, B, H, W):
    N = indices.shape[0]
    nbr = torch.empty((N, 9), dtype=I32, device=_dev(indices))
    L.check(L.lib().gdmae_subm_neighbor_map(L.P(indices), L.i64(N), L.P(rank_grid), B, H, W, L.P(nbr), L.stream()),
            "gdmae_subm_neighbor_map")
    return nbr


def down_neighbor_maps(in_indices, in_grid, H, W, out_indices, out_grid):
    dev = _dev(in_indices)
    N, No = in_indices.shape[0], out_indices.shape[0]
    down = torch.empty((No, 9), dtype=I32, device=dev)
    up = torch.empty((N, 9), dtype=I32, device=dev)
    L.check(L.lib().gdmae_down_neighbor_maps(L.P(in_indices), L.i64(N), L.P(in_grid), H, W, L.P(out_indices), L.i64(No),
                                             L.P(out_grid), L.P(down), L.P(up), L.stream()), "gdmae_down_neighbor_maps")
    return down, up


def _al(n):
    return (n + 15) & ~15


def world_augment(points, params, src_index=None):
    """World flip / rotation / scaling (+ shuffle) of a collated batch on the device (csrc/augment.cu; reference:
    data_augmentor.py:55-143, data_processor.py:92-102).  points (N, 1 + C) fp32, params (B, 6) fp32 rows
    [flip_x, flip_y, cos, sin, scale, 0], src_index (N) int32 or None -> new (N, 1 + C) tensor."""
    _dev(points)
    points = points.contiguous()
    out = torch.empty_like(points)
    L.check(L.lib().gdmae_world_augment(L.P(points), L.i64(points.shape[0]), points.shape[1], L.P(params.contiguous()),
                                        params.shape[0], L.P(src_index), L.P(out), L.stream()), "gdmae_world_augment")
    return out


class WindowTable:
    """Window bookkeeping of one shift (replaces batch_win_inds / coors_in_win / drop levels /
    flat2win_inds / key masks / pos dict of SSTInputLayer.forward, spt_backbone.py:106-135).
    All arrays live in ONE device buffer; ``row_info`` and ``pos_of_token`` (the two the encoder layers read) are
    views made up front, the others (``win_of_token``, ``inner``, ``level``, ``win_tok``, ``win_off``, ``lvl_rank``,
    ``lvl_counts``, ``win_mask`` - used by the sst_utils mirrors and the tests) are made on first access: a step
    builds six tables and the host was the bottleneck while it did so."""
    __slots__ = ("_buf", "_fields", "_views", "row_info", "pos_of_token", "n_windows", "nWx", "nWy", "N", "_pos_long", "_bin_units")

    def __getattr__(self, name):
        # only reached for names without a slot value: the lazily viewed arrays
        fields = object.__getattribute__(self, "_fields")
        if name not in fields:
            raise AttributeError(name)
        views = object.__getattribute__(self, "_views")
        if name not in views:
            off, nbytes, dtype, shape = fields[name]
            views[name] = self._buf[off:off + nbytes].view(dtype).view(shape)
        return views[name]

    def bin_units(self):
        """work units of the tensor-core SRA kernels (gdmae_sra_bin_units), built on first use and cached: one table
        serves two encoder layers, forward and backward"""
        if getattr(self, "_bin_units", None) is None:
            lib = L.lib()
            n = lib.gdmae_sra_bin_units_bytes(L.i64(self.N))
            u = torch.empty((n // 4,), dtype=I32, device=self.row_info.device)
            L.check(lib.gdmae_sra_bin_units(L.P(self.row_info), L.i64(self.N), L.P(u), L.stream()), "gdmae_sra_bin_units")
            self._bin_units = u
        return self._bin_units

    def pos_long(self):
        """in-window cell per token as int64 (index for torch gathers), cached"""
        if getattr(self, "_pos_long", None) is None:
            self._pos_long = self.pos_of_token.long()
        return self._pos_long


def window_table(indices, B, H, W, shifted):
    dev = _dev(indices)
    N = indices.shape[0]
    nWx, nWy = (W + 7) // 8 + 1, (H + 7) // 8 + 1
    nW = B * nWx * nWy
    t = WindowTable()
    t.N, t.n_windows, t.nWx, t.nWy = N, nW, nWx, nWy
    fields, off = {}, 0
    for name, count, dtype, width, shape in (("row_info", 4 * N, I32, 4, (N, 4)), ("win_of_token", N, I32, 4, (N,)),
                                             ("inner", N, I32, 4, (N,)), ("level", N, I32, 4, (N,)), ("win_tok", N, I32, 4, (N,)),
                                             ("win_off", nW + 1, I32, 4, (nW + 1,)), ("lvl_rank", nW, I32, 4, (nW,)),
                                             ("lvl_counts", 3, I32, 4, (3,)), ("win_mask", nW, I64, 8, (nW,)),
                                             ("pos_of_token", N, torch.uint8, 1, (N,))):
        fields[name] = (off, count * width, dtype, shape)
        off += _al(max(count, 1) * width)
    buf = torch.empty((off,), dtype=torch.uint8, device=dev)
    t._buf, t._fields, t._views = buf, fields, {}
    base = buf.data_ptr()
    ptr = {k: ctypes.c_void_p(base + v[0]) for k, v in fields.items()}
    o, nb, dt, sh = fields["row_info"]
    t.row_info = buf[o:o + nb].view(dt).view(sh)
    o, nb, dt, sh = fields["pos_of_token"]
    t.pos_of_token = buf[o:o + nb]
    t._pos_long = None
    t._bin_units = None
    lib = L.lib()
    ws = L.workspace(lib.gdmae_window_table_workspace_bytes(L.i64(nW)), dev)
    L.check(lib.gdmae_window_table(L.P(indices), L.i64(N), B, H, W, int(bool(shifted)), ptr["win_of_token"],
                                   ptr["pos_of_token"], ptr["inner"], ptr["level"], ptr["win_mask"], ptr["win_off"],
                                   ptr["win_tok"], ptr["lvl_rank"], ptr["lvl_counts"], ptr["row_info"], L.P(ws),
                                   ctypes.c_size_t(ws.numel()), L.stream()), "gdmae_window_table")
    return t


# False: fp32 SIMT kernels (parity / tf32 configurations).  True: bf16 tensor-core kernels (bf16 configuration, bf16 q/k/v).
SRA_TENSOR_CORES = False


def sra_wait_timeouts():
    """bounded waits inside the tensor-core SRA kernels that timed out since the library was loaded (0 in a healthy run)"""
    out = ctypes.c_int(0)
    L.check(L.lib().gdmae_sra_wait_timeouts(ctypes.byref(out)), "gdmae_sra_wait_timeouts")
    return int(out.value)


def sra_fwd(qkv, lut, tau, table, tau_min, nhead, bv=None, out_dtype=torch.float32):
    """raw launch: -> (out (N,d) fp32 or bf16, lse (N,nhead)).  bv (d): value bias added to the output
    (then qkv holds v without bias)."""
    N, d3 = qkv.shape
    d = d3 // 3
    out = torch.empty((N, d), dtype=out_dtype, device=qkv.device)
    lse = torch.empty((N, nhead), dtype=F32, device=qkv.device)
    # algorithmic bytes (SURVEY.md 8d, a18 minus projections): N*d*(3*s_in + s_out) + N*8
    with L.timed(f"sra_fwd_d{d}", N * d * (3 * qkv.element_size() + out.element_size()) + N * 8 * 4):
        if qkv.dtype == torch.bfloat16:
            ws = L.workspace(L.lib().gdmae_sra_tc_workspace_bytes(L.i64(N), d), qkv.device)
            L.check(L.lib().gdmae_sra_attention_fwd_tc(L.P(qkv), L.P(lut), L.P(table.row_info), L.P(table.bin_units()), L.i64(N), d, nhead, L.P(tau),
                                                       L.f32(tau_min), L.P(bv), _DT[out_dtype], L.P(out), L.P(lse), L.P(ws),
                                                       ctypes.c_size_t(ws.numel()), L.stream()),
                    "gdmae_sra_attention_fwd_tc")
        else:
            L.check(L.lib().gdmae_sra_attention_fwd(L.P(qkv), L.P(lut), L.P(table.row_info), L.i64(N), d, nhead, L.P(tau),
                                                    L.f32(tau_min), L.P(bv), _DT[out_dtype], L.P(out), L.P(lse), L.stream()),
                    "gdmae_sra_attention_fwd")
    return out, lse


def sra_bwd(qkv, lut, tau, table, tau_min, nhead, out, lse, dout, bv=None, io_dtype=torch.float32):
    """raw launch: -> (dqkv (N,3d) in io_dtype, dtau_sum (1) float64 = sum dS*S); ``out`` is the forward output.
    bf16 qkv (+ bf16 dout) selects the tensor-core kernel, which needs neither ``out`` nor ``bv``."""
    N, d3 = qkv.shape
    d = d3 // 3
    dtau_sum = torch.zeros((1,), dtype=torch.float64, device=qkv.device)
    if qkv.dtype == torch.bfloat16:
        assert dout.dtype == torch.bfloat16
        dqkv = torch.empty((N, d3), dtype=torch.bfloat16, device=qkv.device)
        with L.timed(f"sra_bwd_d{d}", N * d * (6 + 2 + 6) + N * 8 * 4):
            ws = L.workspace(L.lib().gdmae_sra_tc_workspace_bytes(L.i64(N), d), qkv.device)
            L.check(L.lib().gdmae_sra_attention_bwd_tc(L.P(qkv), L.P(lut), L.P(table.row_info), L.P(table.bin_units()), L.i64(N), d, nhead, L.P(tau),
                                                       L.f32(tau_min), L.P(lse), L.P(dout), L.P(dqkv), L.P(dtau_sum), L.P(ws),
                                                       ctypes.c_size_t(ws.numel()), L.stream()),
                    "gdmae_sra_attention_bwd_tc")
        return dqkv, dtau_sum
    assert out.dtype == io_dtype
    dqkv = torch.empty((N, d3), dtype=io_dtype, device=qkv.device)
    work = torch.empty((N, nhead), dtype=F32, device=qkv.device)
    es = dqkv.element_size()
    # bwd algorithmic bytes: qkv + o + dO in, dqkv out
    with L.timed(f"sra_bwd_d{d}", N * d * (12 + es + 4 + 3 * es) + N * 8):
        L.check(L.lib().gdmae_sra_attention_bwd(L.P(qkv), L.P(lut), L.P(table.row_info), L.i64(N), d, nhead, L.P(tau),
                                                L.f32(tau_min), L.P(bv), _DT[io_dtype], L.P(out), L.P(lse), L.P(dout), L.P(dqkv),
                                                L.P(dtau_sum), L.P(work), L.stream()), "gdmae_sra_attention_bwd")
    return dqkv, dtau_sum


# ----------------------------------------------------------------------------- sparse conv gather
class GatherRows(torch.autograd.Function):
    """col (N, 9*C) for a sparse 3x3 conv; backward is the transposed gather (no atomics)."""

    @staticmethod
    @_fwd
    def forward(ctx, x, fwd_map, bwd_map, mirror):
        col = gather_rows(x.contiguous(), fwd_map, F32)
        ctx.save_for_backward(bwd_map)
        ctx.mirror, ctx.n_src = mirror, x.shape[0]
        return col

    @staticmethod
    @_bwd
    def backward(ctx, dcol):
        (bwd_map,) = ctx.saved_tensors
        return gather_rows_transposed(dcol.contiguous(), bwd_map, ctx.n_src, ctx.mirror), None, None, None


def gather_rows(x, fwd_map, out_dtype):
    """col (N, K*C) = rows of x through the neighbour map (0 where empty), fp32 or bf16."""
    N, K = fwd_map.shape
    C = x.shape[1]
    col = torch.empty((N, K * C), dtype=out_dtype, device=_dev(x))
    L.check(L.lib().gdmae_gather_rows(L.P(x), L.P(fwd_map), L.i64(N), K, C, L.P(col), _DT[out_dtype], L.stream()),
            "gdmae_gather_rows")
    return col


def gather_rows_transposed(dcol, bwd_map, n_src, mirror):
    K = bwd_map.shape[1]
    C = dcol.shape[1] // K
    dx = torch.empty((n_src, C), dtype=F32, device=dcol.device)
    L.check(L.lib().gdmae_gather_rows_transposed(L.P(dcol), _DT[dcol.dtype], L.P(bwd_map), L.i64(n_src), K, C, int(mirror), L.P(dx),
                                                 L.stream()), "gdmae_gather_rows_transposed")
    return dx


# ----------------------------------------------------------------------------- SRA attention core
class SraAttention(torch.autograd.Function):
    """Cosine window attention on flat tokens (cosine_msa.py:114-176 + sst_basic_block.py:22-54)."""

    @staticmethod
    @_fwd
    def forward(ctx, qkv, lut, tau, table, tau_min, nhead):
        qkv, lut = qkv.contiguous(), lut.contiguous()
        _dev(qkv)
        tau_c = tau.detach().reshape(-1).contiguous()
        out, lse = sra_fwd(qkv, lut, tau_c, table, tau_min, nhead)
        ctx.save_for_backward(qkv, lut, tau_c, out, lse)
        ctx.table, ctx.tau_min, ctx.nhead, ctx.tau_shape = table, tau_min, nhead, tau.shape
        return out

    @staticmethod
    @_bwd
    def backward(ctx, dout):
        qkv, lut, tau_c, out, lse = ctx.saved_tensors
        t = ctx.table
        d = qkv.shape[1] // 3
        dqkv, dtau_sum = sra_bwd(qkv, lut, tau_c, t, ctx.tau_min, ctx.nhead, out, lse, dout.contiguous())
        # the LUT rows receive the q/k gradients of the tokens sitting on that in-window cell
        dlut = torch.zeros_like(lut)
        dlut.index_add_(0, t.pos_long(), dqkv[:, :2 * d])
        tau_eff = torch.clamp(tau_c, min=ctx.tau_min)
        dtau = torch.where(tau_c >= ctx.tau_min, -(dtau_sum.float() / tau_eff), torch.zeros_like(tau_c))
        return dqkv, dlut, dtau.reshape(ctx.tau_shape), None, None, None


# ----------------------------------------------------------------------------- decoder
class DenseFill(torch.autograd.Function):
    """(B, Y, X, 3*Cs) NHWC map: scale-s rows at covered cells, bg_s elsewhere (see sparse_feat.cu)."""

    @staticmethod
    @_fwd
    def forward(ctx, r0, r1, r2, bg0, bg1, bg2, grids, indices, strides, B, Y, X, out_dtype=torch.float32):
        # bf16 rows (the deblocks of the bf16 configuration emit them) are read as they are; anything else as fp32
        row_dtype = torch.bfloat16 if all(r.dtype == torch.bfloat16 for r in (r0, r1, r2)) else F32
        rows = [r.contiguous().to(row_dtype) for r in (r0, r1, r2)]
        bgs = [b.contiguous().float() for b in (bg0, bg1, bg2)]
        Cs = bgs[0].shape[0]
        out = torch.empty((B, Y, X, 3 * Cs), dtype=out_dtype, device=_dev(rows[0]))
        with L.timed("dense_fill", out.numel() * out.element_size()):
            L.check(L.lib().gdmae_dense_fill(L.parr(rows), _DT[row_dtype], L.parr(bgs), L.parr(grids), L.iarr(strides), B, Y, X, Cs,
                                             L.P(out), _DT[out_dtype], L.stream()), "gdmae_dense_fill")
        ctx.grids, ctx.strides, ctx.dims = grids, strides, (B, Y, X, Cs)
        ctx.row_shapes = [r.shape for r in rows]
        ctx.in_dtypes = [r.dtype for r in (r0, r1, r2)]
        ctx.row_dtype = row_dtype
        return out

    @staticmethod
    @_bwd
    def backward(ctx, dout):
        B, Y, X, Cs = ctx.dims
        dout = dout.contiguous()
        drows = [torch.empty(s, dtype=ctx.row_dtype, device=dout.device) for s in ctx.row_shapes]
        dbg = torch.empty((3 * Cs,), dtype=F32, device=dout.device)
        k2 = [int(k) * int(k) for k in ctx.strides]
        n_sites = (ctypes.c_int64 * 3)(*[int(s[0]) // k2[i] for i, s in enumerate(ctx.row_shapes)])
        with L.timed("dense_fill_bwd", dout.numel() * dout.element_size()):
            L.check(L.lib().gdmae_dense_fill_bwd(L.P(dout), _DT[dout.dtype], L.parr(ctx.grids), n_sites, L.iarr(ctx.strides), B, Y, X, Cs,
                                                 L.parr(drows), _DT[ctx.row_dtype], L.P(dbg), L.stream()),
                    "gdmae_dense_fill_bwd")
        drows = [d if d.dtype == t else d.to(t) for d, t in zip(drows, ctx.in_dtypes)]
        return (drows[0], drows[1], drows[2], dbg[:Cs], dbg[Cs:2 * Cs], dbg[2 * Cs:], None, None, None, None, None, None, None)


class GatherNHWC(torch.autograd.Function):
    """spatial_features.permute(0,2,3,1)[b, y, x] at all pillars (spt_backbone_mae.py:141-143)."""

    @staticmethod
    def forward(ctx, src_nhwc, voxel_coords):
        src_nhwc = src_nhwc.contiguous()
        B, Y, X, C = src_nhwc.shape
        M = voxel_coords.shape[0]
        out = torch.empty((M, C), dtype=F32, device=_dev(src_nhwc))
        L.check(L.lib().gdmae_gather_nhwc(L.P(src_nhwc), _DT[src_nhwc.dtype], L.P(voxel_coords), L.i64(M), Y, X, C, L.P(out),
                                          L.stream()), "gdmae_gather_nhwc")
        ctx.save_for_backward(voxel_coords)
        ctx.shape, ctx.dtype = (B, Y, X, C), src_nhwc.dtype
        return out

    @staticmethod
    def backward(ctx, dout):
        (voxel_coords,) = ctx.saved_tensors
        B, Y, X, C = ctx.shape
        dsrc = torch.zeros(ctx.shape, dtype=ctx.dtype, device=dout.device)
        L.check(L.lib().gdmae_scatter_nhwc(L.P(dout.contiguous().float()), L.P(voxel_coords), L.i64(voxel_coords.shape[0]), Y, X, C,
                                           L.P(dsrc), _DT[ctx.dtype], L.stream()), "gdmae_scatter_nhwc")
        return dsrc, None


class DecoderTail(torch.autograd.Function):
    """relu(BatchNorm2d(y))[pillar cells] without the second dense map: statistics over all B*Y*X cells,
    values only where the MAE head gathers them (spt_backbone_mae.py:52-57, 141-143).  y: conv output,
    NHWC view (B, Y, X, C), fp32 or bf16."""

    @staticmethod
    def forward(ctx, y_nhwc, gamma, beta, running_mean, running_var, momentum, eps, training, voxel_coords, cell2pillar):
        assert training, "DecoderTail implements the training-mode (batch statistics) path"
        y = y_nhwc.contiguous()
        B, Y, X, C = y.shape
        M = voxel_coords.shape[0]
        dev = y.device
        out = torch.empty((M, C), dtype=F32, device=dev)
        mean = torch.empty((C,), dtype=F32, device=dev)
        rstd = torch.empty((C,), dtype=F32, device=dev)
        lib = L.lib()
        ws = L.workspace(lib.gdmae_batchnorm_workspace_bytes(C), dev)
        L.check(lib.gdmae_decoder_tail_fwd(L.P(y), _DT[y.dtype], B, Y, X, C, L.P(voxel_coords), L.i64(M), L.P(gamma), L.P(beta),
                                           L.f32(eps), L.f32(momentum), L.P(out), L.P(mean), L.P(rstd), L.P(running_mean),
                                           L.P(running_var), L.P(ws), ctypes.c_size_t(ws.numel()), L.stream()), "gdmae_decoder_tail_fwd")
        ctx.save_for_backward(y, gamma, mean, rstd, out, voxel_coords, cell2pillar)
        return out

    @staticmethod
    def backward(ctx, dout):
        y, gamma, mean, rstd, out, voxel_coords, cell2pillar = ctx.saved_tensors
        B, Y, X, C = y.shape
        dev = y.device
        dy = torch.empty_like(y)
        dgamma = torch.empty((C,), dtype=F32, device=dev)
        dbeta = torch.empty((C,), dtype=F32, device=dev)
        lib = L.lib()
        ws = L.workspace(lib.gdmae_batchnorm_workspace_bytes(C), dev)
        L.check(lib.gdmae_decoder_tail_bwd(L.P(y), _DT[y.dtype], B, Y, X, C, L.P(voxel_coords), L.P(cell2pillar),
                                           L.i64(voxel_coords.shape[0]), L.P(out), L.P(dout.contiguous().float()), L.P(gamma),
                                           L.P(mean), L.P(rstd), L.P(dy), L.P(dgamma), L.P(dbeta), L.P(ws),
                                           ctypes.c_size_t(ws.numel()), L.stream()), "gdmae_decoder_tail_bwd")
        return dy, dgamma, dbeta, None, None, None, None, None, None, None


DGRAD_AS_FPROP = __import__("os").environ.get("GDMAE_DGRAD_AS_FPROP", "1") != "0"      # =0: the library's dgrad kernel (A/B)


class DecoderConv3x3(torch.autograd.Function):
    """decoder_conv_out's Conv2d(384, 128, 3, padding=1, bias=False) (spt_backbone_mae.py:45-49) on the NHWC bf16 map.
    Forward and the input gradient are the library convolution (cuDNN runs them at the dense tensor roofline); the WEIGHT
    gradient is the own tcgen05 / TMA kernel (gdmae_conv3x3_wgrad: shifted TMA boxes of the map as the operand of every
    tap, K = pixels split across the SMs, TMA reduce-add) - cuDNN's wgrad moved 6x the operand bytes through DRAM.
    x (B, Y, X, 384) bf16, weight = the fp32 master (128, 384, 3, 3), w_op = its bf16 copy."""

    @staticmethod
    def forward(ctx, x_nhwc, weight, w_op):
        x = x_nhwc.contiguous()
        w_cl = w_op.contiguous(memory_format=torch.channels_last)
        y = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w_cl, padding=1)
        ctx.save_for_backward(x, w_cl)
        return y.permute(0, 2, 3, 1)

    @staticmethod
    def backward(ctx, dy_nhwc):
        x, w_cl = ctx.saved_tensors
        dy = dy_nhwc.contiguous()
        B, Y, X, Ci = x.shape
        Co = dy.shape[3]
        dx = None
        if ctx.needs_input_grad[0] and DGRAD_AS_FPROP:
            # stride 1 / padding 1: the input gradient IS a forward convolution of dy with the transposed, flipped filter
            # W'[ci, co, ky, kx] = W[co, ci, 2 - ky, 2 - kx]; the library's forward kernel for this shape runs at 1.8 PFLOP/s, its
            # dedicated dgrad kernel at 1.3 (r2 timeline: 1.19 ms against 0.85 for the same FLOPs)
            w_t = w_cl.flip(2, 3).transpose(0, 1).contiguous(memory_format=torch.channels_last)
            dx = torch.nn.functional.conv2d(dy.permute(0, 3, 1, 2), w_t, padding=1).permute(0, 2, 3, 1)
        elif ctx.needs_input_grad[0]:
            dx = torch.ops.aten.convolution_backward(dy.permute(0, 3, 1, 2), x.permute(0, 3, 1, 2), w_cl, None, [1, 1], [1, 1], [1, 1],
                                                     False, [0, 0], 1, [True, False, False])[0].permute(0, 2, 3, 1)
        dw = torch.empty((Co, 3, 3, Ci), dtype=F32, device=x.device)
        with L.timed("conv3x3_wgrad", 2 * B * Y * X * 9 * Ci * Co):          # FLOPs, not bytes: the kernel is tensor bound
            L.check(L.lib().gdmae_conv3x3_wgrad(L.P(dy), L.P(x), B, Y, X, Ci, Co, L.P(dw), 0, L.stream()), "gdmae_conv3x3_wgrad")
        return dx, dw.permute(0, 3, 1, 2), None


class TallLinear(torch.autograd.Function):
    """y = x W^T + b for a tall x (M rows = all pillars, 10^5..10^6) and a small W (decoder_pred, spt_backbone_mae.py:57,84).
    Forward and input gradient are plain library GEMMs; the weight gradient dW = dy^T x has K = M: the library picks a kernel
    without split-K for it (r2 timeline: 193 us on 4 SMs' worth of tiles), so it is taken as a batched product over S row
    blocks (one tile per block on every SM) followed by a sum over the blocks."""
    S = 128

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        return torch.addmm(bias, x, weight.t())

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        M = x.shape[0]
        dx = dy @ weight if ctx.needs_input_grad[0] else None
        S = TallLinear.S
        mb = M // S
        if mb >= 8:
            head = torch.bmm(dy[:mb * S].view(S, mb, -1).transpose(1, 2), x[:mb * S].view(S, mb, -1)).sum(0)
            dw = head if mb * S == M else head.addmm_(dy[mb * S:].t(), x[mb * S:])
        else:
            dw = dy.t() @ x
        return dx, dw, dy.sum(0)


# ----------------------------------------------------------------------------- chamfer head
def group_points_centered(ps, pc_range, voxel_size, K):
    """gt_points - voxel_centers of target_assigner (spt_backbone_mae.py:67-72), (M, K, 3)."""
    out = torch.empty((ps.n_pillars, K, 3), dtype=F32, device=ps.points.device)
    L.check(L.lib().gdmae_group_points_centered(L.P(ps.points), ps.points.shape[1], L.P(ps.seg_offsets), L.P(ps.seg_points),
                                                L.P(ps.voxel_coords), L.farr(pc_range[:3]), L.farr(voxel_size), L.i64(ps.n_pillars),
                                                K, L.P(out), L.stream()), "gdmae_group_points_centered")
    return out


class ChamferLoss(torch.autograd.Function):
    """pytorch3d.loss.chamfer_distance(pred, gt, weights=w)[0] with its defaults."""

    @staticmethod
    @_fwd
    def forward(ctx, pred, gt, weights):
        pred, gt, weights = pred.contiguous(), gt.contiguous(), weights.contiguous()
        N, P1, _ = pred.shape
        per_item = torch.empty((N,), dtype=F32, device=_dev(pred))
        dpred = torch.empty_like(pred)
        L.check(L.lib().gdmae_chamfer_fwd(L.P(pred), L.P(gt), L.P(weights), L.i64(N), P1, gt.shape[1], L.P(per_item), L.P(dpred),
                                          L.stream()), "gdmae_chamfer_fwd")
        wsum = weights.sum()
        ctx.save_for_backward(dpred, wsum)
        # weights.sum() == 0 -> 0 (pytorch3d returns zeros in that case); per_item is all zero then
        return per_item.sum() / torch.clamp(wsum, min=1e-30)

    @staticmethod
    @_bwd
    def backward(ctx, g):
        dpred, wsum = ctx.saved_tensors
        return dpred * (g / torch.clamp(wsum, min=1e-30)), None, None


def chamfer_distance(x, y, weights=None):
    """Drop-in for pytorch3d.loss.chamfer_distance on the path's call shape: returns (loss, None)."""
    if weights is None:
        weights = torch.ones((x.shape[0],), dtype=F32, device=x.device)
    return ChamferLoss.apply(x, y, weights), None


# ----------------------------------------------------------------------------- CenterHead on the device
def center_assign_targets(gt_boxes, class_map, num_classes_head, feature_map_hw, pc_range, voxel_size, stride, num_max_objs=500,
                          gaussian_overlap=0.1, min_radius=2):
    """CenterHead.assign_targets for one head (center_head.py:105-231) without leaving the device.
    gt_boxes (B, M, 8) with the 1-based class id last; class_map (n_class_total + 1) int32: class id -> id inside the head.
    -> heatmap (B, C, H, W), target_boxes (B, K, 8), iou_boxes (B, K, 7), inds (B, K) int64, mask (B, K) int64."""
    dev = _dev(gt_boxes)
    gt_boxes = gt_boxes.contiguous().float()
    if gt_boxes.shape[-1] != 8:
        raise L.GdmaeError("center_assign_targets handles (x, y, z, dx, dy, dz, heading, class) boxes (no velocity columns)")
    B, M = gt_boxes.shape[0], gt_boxes.shape[1]
    H, W = int(feature_map_hw[0]), int(feature_map_hw[1])
    K, C = int(num_max_objs), int(num_classes_head)
    heat = torch.empty((B, C, H, W), dtype=F32, device=dev)
    tgt = torch.empty((B, K, 8), dtype=F32, device=dev)
    iou_boxes = torch.empty((B, K, 7), dtype=F32, device=dev)
    inds = torch.empty((B, K), dtype=I64, device=dev)
    mask = torch.empty((B, K), dtype=I64, device=dev)
    L.check(L.lib().gdmae_center_assign_targets(L.P(gt_boxes), B, M, L.P(class_map), class_map.numel() - 1, C, H, W, K, int(min_radius),
                                                int(stride), L.farr([pc_range[0], pc_range[1], voxel_size[0], voxel_size[1]]),
                                                L.f32(gaussian_overlap), L.P(heat), L.P(tgt), L.P(iou_boxes), L.P(inds), L.P(mask),
                                                L.stream()), "gdmae_center_assign_targets")
    return heat, tgt, iou_boxes, inds, mask


class CenterFocalLoss(torch.autograd.Function):
    """FocalLossCenterNet(clamp(sigmoid(logits)), target) as one pass over the map (csrc/center_head.cu)"""

    @staticmethod
    @_fwd
    def forward(ctx, logits, target):
        logits, target = logits.contiguous(), target.contiguous()
        graw = torch.empty_like(logits)
        sums = torch.empty((3,), dtype=torch.float64, device=logits.device)
        L.check(L.lib().gdmae_center_focal_loss(L.P(logits), L.P(target), L.i64(logits.numel()), L.P(graw), L.P(sums), L.stream()),
                "gdmae_center_focal_loss")
        num_pos = sums[2]
        # -neg when there is no positive, else -(pos + neg) / num_pos (loss_utils.py:304-308), without a host sync
        scale = torch.where(num_pos == 0, torch.ones_like(num_pos), 1.0 / torch.clamp_min(num_pos, 1.0))
        pos = torch.where(num_pos == 0, torch.zeros_like(num_pos), sums[0])
        ctx.save_for_backward(graw, scale)
        return (-(pos + sums[1]) * scale).float()

    @staticmethod
    @_bwd
    def backward(ctx, dloss):
        graw, scale = ctx.saved_tensors
        # positions with gt == 1 contribute through `pos`, which is dropped when num_pos == 0 - then there are none
        return graw * (-(dloss.double() * scale)).float(), None
