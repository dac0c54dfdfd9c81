"""Manually differentiated building blocks: one autograd node per SRA encoder layer instead of ~60.

The reference runs every encoder layer as ~45 ATen ops forward and as many autograd nodes
backward; on a B200 that path is launch/host bound.  Here a layer is a fixed
sequence of 9 launches forward (4 cuBLAS GEMMs + 5 hand-written kernels) and its backward is
written out by hand, so Python/autograd overhead is paid once per layer.

Replaces (reference file:line, relative to /root/reference):
  EncoderLayer.forward                  pcdet/models/model_utils/sst_basic_block.py:77-84
  WindowAttention.forward               pcdet/models/model_utils/sst_basic_block.py:22-54
  cosine_multi_head_attention_forward   pcdet/models/model_utils/cosine_msa.py:178-438
"""
import ctypes

import torch

from . import _lib as L
from . import ops as _ops

F32 = torch.float32


def _ws(device, cols):
    lib = L.lib()
    nbytes = lib.gdmae_rowwise_workspace_bytes(int(cols))
    ws = L.workspace(nbytes, device)
    return ws, ctypes.c_size_t(ws.numel())


def add_layernorm_fwd(x, res, gamma, beta, eps=1e-5):
    N, d = x.shape
    y = torch.empty_like(x)
    mean = torch.empty((N,), dtype=F32, device=x.device)
    rstd = torch.empty((N,), dtype=F32, device=x.device)
    L.check(L.lib().gdmae_add_layernorm_fwd(L.P(x), L.P(res), L.P(gamma), L.P(beta), L.i64(N), d, L.f32(eps), L.P(y), L.P(mean),
                                            L.P(rstd), L.stream()), "gdmae_add_layernorm_fwd")
    return y, mean, rstd


def add_layernorm_bwd(x, res, gamma, mean, rstd, dy):
    N, d = x.shape
    dz = torch.empty_like(x)
    dgamma = torch.empty((d,), dtype=F32, device=x.device)
    dbeta = torch.empty((d,), dtype=F32, device=x.device)
    ws, n = _ws(x.device, 512)
    L.check(L.lib().gdmae_add_layernorm_bwd(L.P(x), L.P(res), L.P(gamma), L.P(mean), L.P(rstd), L.P(dy), L.i64(N), d, L.P(dz),
                                            L.P(dgamma), L.P(dbeta), 0, L.P(ws), n, L.stream()), "gdmae_add_layernorm_bwd")
    return dz, dgamma, dbeta


def bias_gelu_fwd(h, bias):
    out = torch.empty_like(h)
    L.check(L.lib().gdmae_bias_gelu_fwd(L.P(h), L.P(bias), L.i64(h.shape[0]), h.shape[1], L.P(out), L.stream()),
            "gdmae_bias_gelu_fwd")
    return out


def bias_gelu_bwd(h, bias, dg):
    dh = torch.empty_like(h)
    dbias = torch.empty_like(bias)
    ws, n = _ws(h.device, 512)
    L.check(L.lib().gdmae_bias_gelu_bwd(L.P(h), L.P(bias), L.P(dg), L.i64(h.shape[0]), h.shape[1], L.P(dh), L.P(dbias), 0, L.P(ws),
                                        n, L.stream()), "gdmae_bias_gelu_bwd")
    return dh, dbias


def colsum(x, col0=0, C=None):
    N, ld = x.shape
    C = ld if C is None else C
    out = torch.empty((C,), dtype=F32, device=x.device)
    ws, n = _ws(x.device, 1024)
    L.check(L.lib().gdmae_colsum(L.P(x), L.i64(N), ld, col0, C, L.P(out), 0, L.P(ws), n, L.stream()), "gdmae_colsum")
    return out


sra_fwd, sra_bwd = _ops.sra_fwd, _ops.sra_bwd

# dtype of the cuBLAS GEMM operands of the encoder layers / sparse convs.  torch.float32 = parity
# (or TF32 when allowed); torch.bfloat16 = bf16 operands, fp32 accumulation and fp32 outputs
# (config.set_precision('bf16')).  Residual stream, LayerNorm, softmax and gradients stay fp32.
GEMM_DTYPE = torch.float32


def _g(t):
    """GEMM operand in the configured dtype."""
    return t if GEMM_DTYPE == torch.float32 else t.to(GEMM_DTYPE)


def _mm(a, b):
    if a.dtype == torch.float32:
        return torch.mm(a, b)
    return torch.mm(a, b, out_dtype=torch.float32)


def _addmm(c, a, b):
    """c + a @ b with fp32 output (c fp32: bias row or a full residual matrix)."""
    if a.dtype == torch.float32:
        return torch.addmm(c, a, b)
    return torch.addmm(c, a, b, out_dtype=torch.float32)


class EncoderLayerFunction(torch.autograd.Function):
    """x -> LN2( x1 + W2 gelu(W1 x1 + b1) + b2 ),  x1 = LN1( x + Wo SRA(x) + bo )."""

    @staticmethod
    @_ops._fwd
    def forward(ctx, x, pos_table, table, tau_min, nhead, w_in, b_in, tau, w_o, b_o, g1, be1, w1, b1, w2, b2, g2, be2):
        x = x.contiguous()
        d = x.shape[1]
        tau_c = tau.reshape(-1).contiguous()
        bias_v = torch.cat([torch.zeros(2 * d, dtype=F32, device=x.device), b_in[2 * d:]])
        # GEMM operands: fp32 (TF32 when allowed) or bf16 copies with fp32 accumulate/output
        xg, w_in_g, w_o_g, w1_g, w2_g = _g(x), _g(w_in), _g(w_o), _g(w1), _g(w2)
        qkv = _addmm(bias_v, xg, w_in_g.t())
        lut = torch.addmm(b_in[:2 * d], pos_table, w_in[:2 * d].t())
        o, lse = sra_fwd(qkv, lut, tau_c, table, tau_min, nhead)
        og = _g(o)
        a = _addmm(b_o, og, w_o_g.t())
        x1, mean1, rstd1 = add_layernorm_fwd(x, a, g1, be1)
        x1g = _g(x1)
        h = _mm(x1g, w1_g.t())
        g = _g(bias_gelu_fwd(h, b1))
        f = _addmm(b2, g, w2_g.t())
        x2, mean2, rstd2 = add_layernorm_fwd(x1, f, g2, be2)
        ctx.save_for_backward(x, pos_table, w_in_g, b_in, tau_c, w_o_g, g1, w1_g, b1, w2_g, g2, qkv, lut, o, lse, a, x1, mean1, rstd1,
                              h, g, f, mean2, rstd2, xg, og, x1g)
        ctx.table, ctx.tau_min, ctx.nhead, ctx.tau_shape = table, tau_min, nhead, tau.shape
        return x2

    @staticmethod
    @_ops._bwd
    def backward(ctx, dx2):
        (x, pos_table, w_in, b_in, tau_c, w_o, g1, w1, b1, w2, g2, qkv, lut, o, lse, a, x1, mean1, rstd1, h, g, f, mean2,
         rstd2, xg, og, x1g) = ctx.saved_tensors           # w_*, g, xg, og, x1g are in the GEMM operand dtype
        t = ctx.table
        d = x.shape[1]
        dx2 = dx2.contiguous()
        # ---- LN2 and the feed-forward
        dz2, dg2, dbe2 = add_layernorm_bwd(x1, f, g2, mean2, rstd2, dx2)     # grad wrt f and (residual) x1
        db2 = colsum(dz2)
        dz2g = _g(dz2)
        dw2 = _mm(dz2g.t(), g)
        dgl = _mm(dz2g, w2)
        dh, db1 = bias_gelu_bwd(h, b1, dgl)
        dhg = _g(dh)
        dw1 = _mm(dhg.t(), x1g)
        dx1 = _addmm(dz2, dhg, w1)                                           # residual + through linear1
        # ---- LN1 and the attention
        dz1, dg1, dbe1 = add_layernorm_bwd(x, a, g1, mean1, rstd1, dx1)      # grad wrt a and (residual) x
        db_o = colsum(dz1)
        dz1g = _g(dz1)
        dw_o = _mm(dz1g.t(), og)
        do = _mm(dz1g, w_o)
        dqkv, dtau_sum = sra_bwd(qkv, lut, tau_c, t, ctx.tau_min, ctx.nhead, o, lse, do)
        # in-projection: q = (x + pos) Wq^T + bq, k likewise, v = x Wv^T + bv
        xpos = pos_table.index_select(0, t.pos_long())
        xpos += x
        dqkvg = _g(dqkv)
        dw_in = torch.cat([_mm(dqkvg[:, :2 * d].t(), _g(xpos)), _mm(dqkvg[:, 2 * d:].t(), xg)])
        db_in = torch.cat([colsum(dqkv, 0, 2 * d), colsum(dqkv, 2 * d, d)])
        dx = _addmm(dz1, dqkvg, w_in)
        tau_eff = torch.clamp(tau_c, min=ctx.tau_min)
        dtau = torch.where(tau_c >= ctx.tau_min, -(dtau_sum.float() / tau_eff), torch.zeros_like(tau_c)).reshape(ctx.tau_shape)
        return (dx, None, None, None, None, dw_in, db_in, dtau, dw_o, db_o, dg1, dbe1, dw1, db1, dw2, db2, dg2, dbe2)


def encoder_layer(layer, x, pos_table, table):
    """Fused forward/backward of an EncoderLayer module (parameters read from the module)."""
    at = layer.win_attn.self_attn
    return EncoderLayerFunction.apply(x, pos_table, table, at.tau_min, at.num_heads, at.in_proj_weight, at.in_proj_bias, at.tau,
                                      at.out_proj.weight, at.out_proj.bias, layer.norm1.weight, layer.norm1.bias,
                                      layer.linear1.weight, layer.linear1.bias, layer.linear2.weight, layer.linear2.bias,
                                      layer.norm2.weight, layer.norm2.bias)


class SparseConvFunction(torch.autograd.Function):
    """3x3 sparse conv as gather -> ONE GEMM (spconv SubMConv2d / SparseConv2d, spconv_utils.py:37-56),
    manual backward: dW = dy^T col, dcol = dy W, dx = transposed gather (no atomics).  In the bf16
    configuration the gathered (N, 9*C_in) operand, dy and dcol are bf16 (written directly by the
    gather kernel / the GEMM), accumulation and outputs are fp32."""

    @staticmethod
    @_ops._fwd
    def forward(ctx, x, weight, fwd_map, bwd_map, mirror):
        x = x.contiguous()
        w = _g(weight.view(weight.shape[0], -1))            # (C_out, 9*C_in)
        col = _ops.gather_rows(x, fwd_map, GEMM_DTYPE)
        y = _mm(col, w.t())
        ctx.save_for_backward(col, w, bwd_map)
        ctx.mirror, ctx.n_src, ctx.wshape = mirror, x.shape[0], weight.shape
        return y

    @staticmethod
    @_ops._bwd
    def backward(ctx, dy):
        col, w, bwd_map = ctx.saved_tensors
        dyg = _g(dy.contiguous())
        dw = _mm(dyg.t(), col).view(ctx.wshape)
        dcol = torch.mm(dyg, w)                              # operand dtype (bf16 in the bf16 configuration)
        dx = _ops.gather_rows_transposed(dcol, bwd_map, ctx.n_src, ctx.mirror)
        return dx, dw, None, None, None


class BatchNormReLUFunction(torch.autograd.Function):
    """Training-mode BatchNorm over the rows of (N, C) + ReLU, statistics over ``count`` >= N rows (the
    missing rows are zeros: the empty cells of the decoder's dense map).  Returns (out, bg) where
    bg (C) = relu(BN(0)) is the value every missing row would take.  Running buffers are updated in
    the kernel.  Two passes forward, two backward (csrc/batchnorm.cu)."""

    @staticmethod
    @_ops._fwd
    def forward(ctx, y, gamma, beta, running_mean, running_var, momentum, eps, count, relu):
        y = y.contiguous()
        N, C = y.shape
        dev = y.device
        out = torch.empty_like(y)
        mean = torch.empty((C,), dtype=F32, device=dev)
        rstd = torch.empty((C,), dtype=F32, device=dev)
        lib = L.lib()
        ws = L.workspace(lib.gdmae_batchnorm_workspace_bytes(C), dev)
        L.check(lib.gdmae_batchnorm_relu_fwd(L.P(y), L.P(gamma), L.P(beta), L.i64(N), C, ctypes.c_double(count), L.f32(eps),
                                             L.f32(momentum), int(relu), L.P(out), L.P(mean), L.P(rstd), L.P(running_mean),
                                             L.P(running_var), L.P(ws), ctypes.c_size_t(ws.numel()), L.stream()),
                "gdmae_batchnorm_relu_fwd")
        shift = beta - mean * rstd * gamma
        bg = torch.relu(shift) if relu else shift
        ctx.save_for_backward(y, out, gamma, mean, rstd, shift)
        ctx.count, ctx.relu = count, relu
        return out, bg

    @staticmethod
    @_ops._bwd
    def backward(ctx, dout, dbg):
        y, out, gamma, mean, rstd, shift = ctx.saved_tensors
        N, C = y.shape
        dev = y.device
        e_db = e_dg = None
        if ctx.count > N and dbg is not None:
            e_db = (dbg * (shift > 0)) if ctx.relu else dbg
            e_db = e_db.contiguous().float()
            e_dg = (e_db * (-mean * rstd)).contiguous()
        dy = torch.empty_like(y)
        dgamma = torch.empty((C,), dtype=F32, device=dev)
        dbeta = torch.empty((C,), dtype=F32, device=dev)
        lib = L.lib()
        ws = L.workspace(lib.gdmae_batchnorm_workspace_bytes(C), dev)
        L.check(lib.gdmae_batchnorm_relu_bwd(L.P(y), L.P(out), L.P(dout.contiguous()), L.P(gamma), L.P(mean), L.P(rstd), L.i64(N), C,
                                             ctypes.c_double(ctx.count), int(ctx.relu), L.P(e_db), L.P(e_dg), L.P(dy), L.P(dgamma),
                                             L.P(dbeta), L.P(ws), ctypes.c_size_t(ws.numel()), L.stream()),
                "gdmae_batchnorm_relu_bwd")
        return dy, dgamma, dbeta, None, None, None, None, None, None


def batchnorm_relu(bn, y, training, relu=True, count=None):
    """BatchNorm1d/2d module ``bn`` (+ ReLU) applied to the rows of y (N, C); -> (out, bg)."""
    if not training:
        scale = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
        shift = bn.bias - bn.running_mean * scale
        out = y * scale + shift
        return (torch.relu(out), torch.relu(shift)) if relu else (out, shift)
    out, bg = BatchNormReLUFunction.apply(y, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum, bn.eps,
                                          float(y.shape[0] if count is None else count), relu)
    bn.num_batches_tracked += 1
    return out, bg
