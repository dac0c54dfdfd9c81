"""Manually differentiated building blocks: one autograd node per SRA encoder layer instead of ~60.

The reference runs every encoder layer as ~45 ATen ops forward and as many autograd nodes
backward; on a B200 that path is launch/host bound.  Here a layer is a fixed
sequence of launches (4 cuBLAS GEMMs + 5 hand-written kernels forward) and its backward is
written out by hand, so Python/autograd overhead is paid once per layer.  All biases live in the
hand-written kernels (out-proj / FFN2 bias inside residual+LayerNorm, FFN1 bias inside GELU, q/k
bias inside the positional LUT, v bias added to the attention output), so every GEMM is a plain
C = A B.  In the bf16 configuration the GEMM operands are bf16 copies that the same kernels emit
next to (or instead of) their fp32 results; GEMM outputs, the residual stream, LayerNorm / softmax /
BatchNorm statistics and all gradients of parameters stay fp32.

Replaces (reference file:line, relative to /root/reference):
  EncoderLayer.forward                  pcdet/models/model_utils/sst_basic_block.py:77-84
  WindowAttention.forward               pcdet/models/model_utils/sst_basic_block.py:22-54
  cosine_multi_head_attention_forward   pcdet/models/model_utils/cosine_msa.py:178-438
  post_act_block (conv + BN1d + ReLU)   pcdet/utils/spconv_utils.py:37-56
"""
import ctypes

import torch

from . import _lib as L
from . import ops as _ops

F32, BF16 = torch.float32, torch.bfloat16

# dtype of the GEMM operands of the encoder layers / sparse convs: torch.float32 (parity; TF32 when
# allowed) or torch.bfloat16 (config.set_precision(model, 'bf16')).
GEMM_DTYPE = torch.float32


def _bf():
    return GEMM_DTYPE == BF16


def _g(t):
    """GEMM operand in the configured dtype (cast only if needed)."""
    return t if t.dtype == GEMM_DTYPE else t.to(GEMM_DTYPE)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _ws(device, cols):
    lib = L.lib()
    nbytes = lib.gdmae_rowwise_workspace_bytes(int(cols))
    ws = L.workspace(nbytes, device)
    return ws, ctypes.c_size_t(ws.numel())


# ----------------------------------------------------------------------------- GEMM
def _layout(t):
    """2-D tensor -> (tensor, transposed?, leading dimension) for the row-major C ABI."""
    if t.stride(1) == 1 and t.stride(0) >= t.shape[1]:
        return t, 0, t.stride(0)
    if t.stride(0) == 1 and t.stride(1) >= t.shape[0]:
        return t, 1, t.stride(1)
    t = t.contiguous()
    return t, 0, t.stride(0)


def gemm(a, b, out=None, beta=0.0, out_dtype=F32):
    """out (M,N) = a (M,K) @ b (K,N) (+ beta*out).  fp32 operands go through torch (cuBLAS, TF32 when
    allowed); bf16 operands through gdmae_gemm (cublasLt bf16 x bf16 -> fp32 or bf16, cached algorithm)."""
    if a.dtype == F32:
        if out is None:
            return torch.mm(a, b)
        return out.addmm_(a, b) if beta == 1.0 else torch.mm(a, b, out=out)
    M, K = a.shape
    N = b.shape[1]
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    a, ta, lda = _layout(a)
    b, tb, ldb = _layout(b)
    L.check(L.lib().gdmae_gemm(ta, tb, L.i64(M), L.i64(N), L.i64(K), _ptr(a), L.i64(lda), _ptr(b), L.i64(ldb), 1, _ptr(out),
                               L.i64(out.stride(0)), _ops._DT[out.dtype], L.f32(beta), L.stream()), "gdmae_gemm")
    return out


# ----------------------------------------------------------------------------- row kernels
def add_layernorm_fwd(x, res, bias, gamma, beta, eps=1e-5):
    """-> y fp32, y in GEMM dtype (same tensor in fp32 mode), mean, rstd"""
    N, d = x.shape
    y = torch.empty_like(x)
    yb = torch.empty((N, d), dtype=BF16, device=x.device) if _bf() else None
    mean = torch.empty((N,), dtype=F32, device=x.device)
    rstd = torch.empty((N,), dtype=F32, device=x.device)
    L.check(L.lib().gdmae_add_layernorm_fwd(L.P(x), L.P(res), L.P(bias), L.P(gamma), L.P(beta), L.i64(N), d, L.f32(eps), L.P(y),
                                            L.P(yb), L.P(mean), L.P(rstd), L.stream()), "gdmae_add_layernorm_fwd")
    return y, (yb if yb is not None else y), mean, rstd


def add_layernorm_bwd(x, res, bias, gamma, mean, rstd, dy):
    """-> dz fp32, dz in GEMM dtype, dgamma, dbeta"""
    N, d = x.shape
    dz = torch.empty_like(x)
    dzb = torch.empty((N, d), dtype=BF16, device=x.device) if _bf() else None
    dgamma = torch.empty((d,), dtype=F32, device=x.device)
    dbeta = torch.empty((d,), dtype=F32, device=x.device)
    ws, n = _ws(x.device, 512)
    L.check(L.lib().gdmae_add_layernorm_bwd(L.P(x), L.P(res), L.P(bias), L.P(gamma), L.P(mean), L.P(rstd), L.P(dy), L.i64(N), d,
                                            L.P(dz), L.P(dzb), L.P(dgamma), L.P(dbeta), 0, L.P(ws), n, L.stream()),
            "gdmae_add_layernorm_bwd")
    return dz, (dzb if dzb is not None else dz), dgamma, dbeta


def bias_gelu_fwd(h, bias):
    """gelu(h + bias) in the GEMM dtype"""
    out = torch.empty(h.shape, dtype=GEMM_DTYPE, device=h.device)
    o32, o16 = (None, out) if _bf() else (out, None)
    L.check(L.lib().gdmae_bias_gelu_fwd(L.P(h), L.P(bias), L.i64(h.shape[0]), h.shape[1], L.P(o32), L.P(o16), L.stream()),
            "gdmae_bias_gelu_fwd")
    return out


def bias_gelu_bwd(h, bias, dg):
    """-> dh in the GEMM dtype, dbias fp32"""
    dh = torch.empty(h.shape, dtype=GEMM_DTYPE, device=h.device)
    d32, d16 = (None, dh) if _bf() else (dh, None)
    dbias = torch.empty_like(bias)
    ws, n = _ws(h.device, 512)
    L.check(L.lib().gdmae_bias_gelu_bwd(L.P(h), L.P(bias), L.P(dg), L.i64(h.shape[0]), h.shape[1], L.P(d32), L.P(d16), L.P(dbias), 0,
                                        L.P(ws), n, L.stream()), "gdmae_bias_gelu_bwd")
    return dh, dbias


def colsum(x, col0=0, C=None):
    N, ld = x.shape
    C = ld if C is None else C
    out = torch.empty((C,), dtype=F32, device=x.device)
    ws, n = _ws(x.device, 1024)
    L.check(L.lib().gdmae_colsum(L.P(x), _ops._DT[x.dtype], L.i64(N), ld, col0, C, L.P(out), 0, L.P(ws), n, L.stream()),
            "gdmae_colsum")
    return out


def gather_add_rows(x, table, idx_u8):
    """x + table[idx] in the GEMM dtype (q = k = feat + pos, sst_basic_block.py:39-46)"""
    out = torch.empty(x.shape, dtype=GEMM_DTYPE, device=x.device)
    o32, o16 = (None, out) if _bf() else (out, None)
    L.check(L.lib().gdmae_gather_add_rows(L.P(x), L.P(table), L.P(idx_u8), L.i64(x.shape[0]), x.shape[1], L.P(o32), L.P(o16),
                                          L.stream()), "gdmae_gather_add_rows")
    return out


# bf16 copy of the previous layer's output (written by its LayerNorm kernel) handed to the next layer
_LAST_OUT = (None, None)


class EncoderLayerFunction(torch.autograd.Function):
    """x -> LN2( x1 + W2 gelu(W1 x1 + b1) + b2 ),  x1 = LN1( x + Wo SRA(x) + bo )."""

    @staticmethod
    @_ops._fwd
    def forward(ctx, x, pos_table, table, tau_min, nhead, w_in, b_in, tau, w_o, b_o, g1, be1, w1, b1, w2, b2, g2, be2):
        global _LAST_OUT
        x = x.contiguous()
        d = x.shape[1]
        tau_c = tau.reshape(-1).contiguous()
        w_in_g, w_o_g, w1_g, w2_g = _g(w_in), _g(w_o), _g(w1), _g(w2)
        xg = _LAST_OUT[1] if _LAST_OUT[0] is x else _g(x)
        qkv = gemm(xg, w_in_g.t())                                               # (N,3d): q, k, v without biases
        lut = torch.addmm(b_in[:2 * d], pos_table, w_in[:2 * d].t())             # (64,2d): pos term + q/k biases
        bv = b_in[2 * d:].contiguous()
        o, lse = _ops.sra_fwd(qkv, lut, tau_c, table, tau_min, nhead, bv=bv, out_dtype=GEMM_DTYPE)
        a = gemm(o, w_o_g.t())
        x1, x1g, mean1, rstd1 = add_layernorm_fwd(x, a, b_o, g1, be1)
        h = gemm(x1g, w1_g.t())
        g = bias_gelu_fwd(h, b1)
        f = gemm(g, w2_g.t())
        x2, x2g, mean2, rstd2 = add_layernorm_fwd(x1, f, b2, g2, be2)
        _LAST_OUT = (x2, x2g)
        ctx.save_for_backward(x, xg, pos_table, w_in_g, tau_c, bv, w_o_g, b_o, g1, w1_g, b1, w2_g, b2, g2, qkv, lut, o, lse, a, x1,
                              x1g, mean1, rstd1, h, g, f, mean2, rstd2)
        ctx.table, ctx.tau_min, ctx.nhead, ctx.tau_shape = table, tau_min, nhead, tau.shape
        return x2

    @staticmethod
    @_ops._bwd
    def backward(ctx, dx2):
        (x, xg, pos_table, w_in, tau_c, bv, w_o, b_o, g1, w1, b1, w2, b2, g2, qkv, lut, o, lse, a, x1, x1g, mean1, rstd1, h, g, f,
         mean2, rstd2) = ctx.saved_tensors                       # w_*, xg, o, x1g, g are in the GEMM operand dtype
        t = ctx.table
        d = x.shape[1]
        dx2 = dx2.contiguous()
        # ---- LN2 and the feed-forward
        dz2, dz2g, dg2, dbe2 = add_layernorm_bwd(x1, f, b2, g2, mean2, rstd2, dx2)   # grad wrt f, b2 and (residual) x1
        db2 = colsum(dz2)
        dw2 = gemm(dz2g.t(), g)
        dgl = gemm(dz2g, w2)
        dh, db1 = bias_gelu_bwd(h, b1, dgl)
        dw1 = gemm(dh.t(), x1g)
        dx1 = gemm(dh, w1, out=dz2, beta=1.0)                                        # residual + through linear1, in place
        # ---- LN1 and the attention
        dz1, dz1g, dg1, dbe1 = add_layernorm_bwd(x, a, b_o, g1, mean1, rstd1, dx1)   # grad wrt a, b_o and (residual) x
        db_o = colsum(dz1)
        dw_o = gemm(dz1g.t(), o)
        do = gemm(dz1g, w_o)
        dqkv, dtau_sum = _ops.sra_bwd(qkv, lut, tau_c, t, ctx.tau_min, ctx.nhead, o, lse, do, bv=bv, io_dtype=GEMM_DTYPE)
        # in-projection: q = (x + pos) Wq^T + bq, k likewise, v = x Wv^T + bv
        xpos = gather_add_rows(x, pos_table, t.pos_of_token)
        dw_in = torch.cat([gemm(dqkv[:, :2 * d].t(), xpos), gemm(dqkv[:, 2 * d:].t(), xg)])
        db_in = torch.cat([colsum(dqkv, 0, 2 * d), colsum(dqkv, 2 * d, d)])
        dx = gemm(dqkv, w_in, out=dz1, beta=1.0)
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
        y = gemm(col, w.t())
        ctx.save_for_backward(col, w, bwd_map)
        ctx.mirror, ctx.n_src, ctx.wshape = mirror, x.shape[0], weight.shape
        return y

    @staticmethod
    @_ops._bwd
    def backward(ctx, dy):
        col, w, bwd_map = ctx.saved_tensors
        dyg = _g(dy.contiguous())
        dw = gemm(dyg.t(), col).view(ctx.wshape)
        dcol = gemm(dyg, w, out_dtype=GEMM_DTYPE)            # operand dtype (bf16 output in the bf16 configuration)
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
