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
import os

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


# True: bf16 contractions run on the own tcgen05 / TMA kernel (csrc/tc_gemm.cu).  False (debug only): cuBLASLt.
TC_GEMM = True


def gemm(a, b, out=None, beta=0.0, out_dtype=F32, wgrad=False):
    """out (M,N) = a (M,K) @ b (K,N) (+ beta*out).  bf16 operands: the own tcgen05 / TMA GEMM (``wgrad``: K = tokens, split
    over the SMs, fp32 reductions into ``out``); fp32 operands (parity configurations) go through torch (cuBLAS, TF32
    when allowed)."""
    if a.dtype == F32:
        if out is None:
            return torch.mm(a, b)
        return out.addmm_(a, b) if beta == 1.0 else torch.mm(a, b, out=out)
    if TC_GEMM and b.shape[1] % 64 == 0:
        return tc_gemm(a, b, out=out, beta=beta, out_dtype=out_dtype, split_k=wgrad)
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
# Parameter gradients of the two C executors (EncoderLayer, VFE) go straight into `.grad` - as a side effect, with None returned
# to autograd - only while this flag is set: MAETrainer sets it around its backward (its `.grad` tensors are views of the flat
# bucket).  Anything else (torch.autograd.grad, hooks, a DDP wrapper, parameters that merely happen to have a .grad) gets the
# gradients returned the normal way.
INPLACE_PARAM_GRADS = False

# bf16 shadows of parameters, keyed by the parameter's data pointer: the trainer keeps one bf16 mirror of its
# flat parameter bucket (refreshed once per step) so no per-layer weight casts are launched.
BF16_SHADOW = {}


def _gw(w):
    """GEMM-dtype copy of a weight: the trainer's bf16 shadow if registered, else a cast."""
    if GEMM_DTYPE == F32:
        return w
    sh = BF16_SHADOW.get(w.data_ptr())
    return sh if sh is not None else w.to(BF16)


class _ShadowCast(torch.autograd.Function):
    """weight -> its bf16 shadow (no cast kernel), gradient routed back to the fp32 master."""

    @staticmethod
    def forward(ctx, w, shadow):
        return shadow

    @staticmethod
    def backward(ctx, g):
        return g.float(), None


def cast_param(w, dtype):
    """autograd-visible cast of a parameter for ATen consumers (the cuDNN decoder conv)."""
    if w.dtype == dtype:
        return w
    sh = BF16_SHADOW.get(w.data_ptr()) if dtype == BF16 else None
    return w.to(dtype) if sh is None else _ShadowCast.apply(w, sh)


def gemm_mode():
    """gdmae_gemm operand mode of the current configuration: 1 bf16, 0 fp32/TF32 math, 2 fp32/fp32 math."""
    if GEMM_DTYPE == BF16:
        return 1
    return 0 if torch.backends.cuda.matmul.allow_tf32 else 2


_VP = ctypes.c_void_p
_EL_PTRS = ["x", "xg_in", "pos_table", "row_info", "bin_units", "pos_of_token",
            "w_in", "b_in", "tau", "w_o", "b_o", "g1", "be1", "w1", "b1", "w2", "b2", "g2", "be2",
            "w_in_g", "w_o_g", "w1_g", "w2_g",
            "xg", "qkv", "lut", "o", "lse", "a", "x1", "x1g", "mean1", "rstd1", "h", "g", "f", "mean2", "rstd2", "x2", "x2g",
            "dy", "dx", "d_w_in", "d_b_in", "d_tau", "d_w_o", "d_b_o", "d_g1", "d_be1", "d_w1", "d_b1", "d_w2", "d_b2", "d_g2",
            "d_be2", "ws"]


class EncoderLayerArgs(ctypes.Structure):
    """mirror of gdmae_encoder_layer_args (include/gdmae_b200.h)"""
    _fields_ = ([("N", ctypes.c_int64), ("d", ctypes.c_int), ("dff", ctypes.c_int), ("nhead", ctypes.c_int),
                 ("gemm_mode", ctypes.c_int), ("sra_tensor_cores", ctypes.c_int), ("accumulate", ctypes.c_int),
                 ("tau_min", ctypes.c_float), ("eps", ctypes.c_float)]
                + [(n, _VP) for n in _EL_PTRS] + [("ws_bytes", ctypes.c_size_t), ("stream", _VP)])


_PARAM_NAMES = ("w_in", "b_in", "tau", "w_o", "b_o", "g1", "be1", "w1", "b1", "w2", "b2", "g2", "be2")


def _r64(n):
    return (n + 63) & ~63


class EncoderLayerFunction(torch.autograd.Function):
    """x -> LN2( x1 + W2 gelu(W1 x1 + b1) + b2 ),  x1 = LN1( x + Wo SRA(x) + bo ).
    Forward and backward are one C call each (csrc/encoder_layer.cu); this class only owns the buffers."""

    @staticmethod
    @_ops._fwd
    def forward(ctx, x, pos_table, table, tau_min, nhead, eps, *params):
        global _LAST_OUT
        x = x.contiguous()
        N, d = x.shape
        dff = params[7].shape[0]
        mode = gemm_mode()
        bf = mode == 1
        dev = x.device
        A = EncoderLayerArgs()
        A.N, A.d, A.dff, A.nhead, A.gemm_mode = N, d, dff, nhead, mode
        tc = bool(_ops.SRA_TENSOR_CORES) and bf
        A.sra_tensor_cores, A.accumulate, A.tau_min, A.eps = int(tc), 0, tau_min, eps
        A.x, A.pos_table, A.row_info, A.pos_of_token = x.data_ptr(), pos_table.data_ptr(), table.row_info.data_ptr(), \
            table.pos_of_token.data_ptr()
        keep = [x, pos_table, table.row_info, table.pos_of_token]
        if tc:
            units = table.bin_units()
            A.bin_units = units.data_ptr()
            keep.append(units)
        for name, p in zip(_PARAM_NAMES, params):
            setattr(A, name, p.data_ptr())
        for name, p in zip(("w_in_g", "w_o_g", "w1_g", "w2_g"), (params[0], params[3], params[7], params[9])):
            pg = _gw(p)
            keep.append(pg)
            setattr(A, name, pg.data_ptr())
        xg_in = _LAST_OUT[1] if (bf and _LAST_OUT[0] is x) else None
        if xg_in is not None:
            keep.append(xg_in)
            A.xg_in = xg_in.data_ptr()
        # ---- activations saved for backward: one fp32 and one bf16 buffer, carved here
        Nd, Nf = _r64(N * d), _r64(N * dff)
        # a, h, f (GEMM outputs that only feed a row kernel) are bf16 in the bf16 configuration: half the 32-bit words
        Na, Nh = (Nd // 2, Nf // 2) if bf else (Nd, Nf)
        Nl = _r64(N * (24 if tc else 8))    # tensor-core SRA: per-row records lse | 1/|q| | 1/|k|
        n32 = (0 if tc else 3 * Nd) + 2 * Na + Nh + Nd + Nl + 4 * _r64(N) + 128 * d + (0 if bf else Nd + Nf)
        save32 = torch.empty((max(n32, 1),), dtype=F32, device=dev)
        b = save32.data_ptr()
        off = 0
        for name, n in (("qkv", 0 if tc else 3 * Nd), ("a", Na), ("x1", Nd), ("h", Nh), ("f", Na), ("lse", Nl), ("mean1", _r64(N)),
                        ("rstd1", _r64(N)), ("mean2", _r64(N)), ("rstd2", _r64(N)), ("lut", 128 * d)):
            setattr(A, name, b + 4 * off)
            off += n
        x2 = torch.empty((N, d), dtype=F32, device=dev)
        A.x2 = x2.data_ptr()
        save16 = x2g = None
        if bf:
            # tensor-core SRA: window-major q^ | k^ | v (forward) | dO (backward) - N * d is a multiple of 64 elements (128 B)
            save16 = torch.empty((max(4 * Nd + Nf + (4 * Nd if tc else 0), 1),), dtype=BF16, device=dev)
            b16 = save16.data_ptr()
            A.xg, A.o, A.x1g, A.x2g, A.g = b16, b16 + 2 * Nd, b16 + 4 * Nd, b16 + 6 * Nd, b16 + 8 * Nd
            if tc:
                A.qkv = b16 + 2 * (4 * Nd + Nf)
            x2g = save16[3 * Nd:3 * Nd + N * d].view(N, d)
        else:
            A.o, A.g = b + 4 * off, b + 4 * (off + Nd)
        A.stream = torch.cuda.current_stream().cuda_stream
        L.check(L.lib().gdmae_encoder_layer_fwd(ctypes.byref(A)), "gdmae_encoder_layer_fwd")
        _LAST_OUT = (x2, x2g)
        ctx.save_for_backward(x, save32, *params)
        ctx.args, ctx.keep, ctx.save16 = A, keep, save16
        ctx.params = params
        return x2

    @staticmethod
    @_ops._bwd
    def backward(ctx, dx2):
        A, params = ctx.args, ctx.params
        x = ctx.saved_tensors[0]
        dx2 = dx2.contiguous()
        dev = x.device
        N, d, dff = A.N, A.d, A.dff
        dx = torch.empty((N, d), dtype=F32, device=dev)
        A.dy, A.dx = dx2.data_ptr(), dx.data_ptr()
        if getattr(ctx, "_consumed", False):
            raise RuntimeError("EncoderLayerFunction.backward called twice: the activation buffers were released after the first call "
                               "(retain_graph is not supported by the fused node)")
        ctx._consumed = True
        # parameter gradients: straight into .grad when the trainer asked for it (its flat bucket)
        inplace = INPLACE_PARAM_GRADS and all(p.grad is not None and p.grad.dtype == F32 and p.grad.is_contiguous() for p in params)
        if inplace:
            grads = [p.grad for p in params]
            A.accumulate = 1
        else:
            sizes = [_r64(p.numel()) for p in params]
            flat = torch.empty((sum(sizes),), dtype=F32, device=dev)
            grads, off = [], 0
            for p, n in zip(params, sizes):
                grads.append(flat[off:off + p.numel()].view(p.shape))
                off += n
            A.accumulate = 0
        for name, g in zip(_PARAM_NAMES, grads):
            setattr(A, "d_" + name, g.data_ptr())
        nbytes = L.lib().gdmae_encoder_layer_bwd_workspace_bytes(L.i64(N), d, dff)
        ws = L.workspace(nbytes, dev)
        A.ws, A.ws_bytes = ws.data_ptr(), ws.numel()
        A.stream = torch.cuda.current_stream().cuda_stream
        L.check(L.lib().gdmae_encoder_layer_bwd(ctypes.byref(A)), "gdmae_encoder_layer_bwd")
        ctx.keep = ctx.save16 = None
        if inplace:
            return (dx,) + (None,) * (5 + len(params))
        return (dx, None, None, None, None, None) + tuple(grads)


def encoder_layer(layer, x, pos_table, table):
    """Fused forward/backward of an EncoderLayer module (parameters read from the module)."""
    at = layer.win_attn.self_attn
    return EncoderLayerFunction.apply(x, pos_table, table, at.tau_min, at.num_heads, layer.norm1.eps, at.in_proj_weight,
                                      at.in_proj_bias, at.tau, at.out_proj.weight, at.out_proj.bias, layer.norm1.weight,
                                      layer.norm1.bias, layer.linear1.weight, layer.linear1.bias, layer.linear2.weight,
                                      layer.linear2.bias, layer.norm2.weight, layer.norm2.bias)


_VFE_PTRS = ["x", "seg_offsets", "seg_points", "W1", "g1", "b1", "g2", "b2", "W2_g", "running_mean1", "running_var1", "running_mean2",
             "running_var2", "h1", "y2", "mean1", "rstd1", "mean2", "rstd2", "out", "argmax", "dout", "dy2", "dh1", "tmp_dbeta1",
             "tmp_dgamma1", "tmp_dbeta2", "tmp_dgamma2", "d_W1", "d_g1", "d_b1", "d_W2", "d_g2", "d_b2", "moments", "ws"]


class VfeMlpArgs(ctypes.Structure):
    """mirror of gdmae_vfe_mlp_args (include/gdmae_b200.h)"""
    _fields_ = ([("Np", ctypes.c_int64), ("M", ctypes.c_int64), ("K", ctypes.c_int), ("C1", ctypes.c_int), ("C2", ctypes.c_int),
                 ("gemm_mode", ctypes.c_int), ("accumulate", ctypes.c_int), ("eps", ctypes.c_float), ("momentum", ctypes.c_float)]
                + [(n, _VP) for n in _VFE_PTRS] + [("ws_bytes", ctypes.c_size_t), ("stream", _VP)])


class VfeMlpFunction(torch.autograd.Function):
    """pillar features (M,128) = scatter_max(relu(bn2(relu(bn1(x W1^T)) W2^T))) - the whole DynVFE MLP in one node
    (csrc/vfe_mlp.cu): point-sized intermediates are recomputed or never formed."""

    @staticmethod
    @_ops._fwd
    def forward(ctx, x, seg_offsets, seg_points, M, eps, momentum, rm1, rv1, rm2, rv2, W1, g1, b1, W2, g2, b2):
        x = x.contiguous()
        Np, K = x.shape
        dev = x.device
        mode = gemm_mode()
        opdt = BF16 if mode == 1 else F32
        A = VfeMlpArgs()
        A.Np, A.M, A.K, A.C1, A.C2, A.gemm_mode, A.accumulate, A.eps, A.momentum = Np, M, K, 64, 128, mode, 0, eps, momentum
        W2g = _gw(W2)
        h1 = torch.empty((Np, 64), dtype=opdt, device=dev)
        y2 = torch.empty((Np, 128), dtype=opdt, device=dev)
        stats = torch.empty((4 * 128 + 4 * 128 + 2 * 160,), dtype=F32, device=dev)   # mean1 rstd1 mean2 rstd2 | 4 scratch rows | moments (fp64)
        out = torch.empty((M, 128), dtype=F32, device=dev)
        arg = torch.empty((M, 128), dtype=torch.uint8, device=dev)   # position inside the pillar's segment (csrc/vfe_mlp.cu)
        for name, t in (("x", x), ("seg_offsets", seg_offsets), ("seg_points", seg_points), ("W1", W1), ("g1", g1), ("b1", b1),
                        ("g2", g2), ("b2", b2), ("W2_g", W2g), ("running_mean1", rm1), ("running_var1", rv1), ("running_mean2", rm2),
                        ("running_var2", rv2), ("h1", h1), ("y2", y2), ("out", out), ("argmax", arg)):
            setattr(A, name, None if t is None else t.data_ptr())      # seg_points None: rows of x already in pillar order
        sb = stats.data_ptr()
        for i, name in enumerate(("mean1", "rstd1", "mean2", "rstd2", "tmp_dbeta1", "tmp_dgamma1", "tmp_dbeta2", "tmp_dgamma2")):
            setattr(A, name, sb + 4 * 128 * i)
        A.moments = sb + 4 * 128 * 8
        lib = L.lib()
        ws = L.workspace(lib.gdmae_vfe_mlp_workspace_bytes(K), dev)
        A.ws, A.ws_bytes, A.stream = ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream
        L.check(lib.gdmae_vfe_mlp_fwd(ctypes.byref(A)), "gdmae_vfe_mlp_fwd")
        ctx.save_for_backward(x, seg_offsets, W1, g1, b1, W2, g2, b2)
        ctx.args, ctx.keep = A, (W2g, h1, y2, stats, out, arg, seg_points)
        ctx.params = (W1, g1, b1, W2, g2, b2)
        return out

    @staticmethod
    @_ops._bwd
    def backward(ctx, dout):
        A, params = ctx.args, ctx.params
        x = ctx.saved_tensors[0]
        dev = x.device
        dout = dout.contiguous().float()
        opdt = BF16 if A.gemm_mode == 1 else F32
        dy2 = torch.empty((A.Np, 128), dtype=opdt, device=dev)
        dh1 = torch.empty((A.Np, 64), dtype=opdt, device=dev)
        if getattr(ctx, "_consumed", False):
            raise RuntimeError("VfeMlpFunction.backward called twice: the activation buffers were released after the first call")
        ctx._consumed = True
        inplace = INPLACE_PARAM_GRADS and all(p.grad is not None and p.grad.dtype == F32 and p.grad.is_contiguous() for p in params)
        grads = [p.grad for p in params] if inplace else [torch.empty_like(p) for p in params]
        A.accumulate = int(inplace)
        A.dout, A.dy2, A.dh1 = dout.data_ptr(), dy2.data_ptr(), dh1.data_ptr()
        for name, g in zip(("d_W1", "d_g1", "d_b1", "d_W2", "d_g2", "d_b2"), grads):
            setattr(A, name, g.data_ptr())
        lib = L.lib()
        ws = L.workspace(lib.gdmae_vfe_mlp_workspace_bytes(A.K), dev)
        A.ws, A.ws_bytes, A.stream = ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream
        L.check(lib.gdmae_vfe_mlp_bwd(ctypes.byref(A)), "gdmae_vfe_mlp_bwd")
        ctx.keep = None
        if inplace:
            return (None,) * 16
        return (None,) * 10 + tuple(grads)


def vfe_mlp_supported(mlp, x):
    """the fused node covers the shipped configs: Linear(K<=16, 64) -> BN -> ReLU -> Linear(64, 128) -> BN -> ReLU, no biases"""
    import torch.nn as nn
    return (len(mlp) == 6 and isinstance(mlp[0], nn.Linear) and isinstance(mlp[3], nn.Linear) and mlp[0].bias is None
            and mlp[3].bias is None and mlp[0].out_features == 64 and mlp[3].in_features == 64 and mlp[3].out_features == 128
            and x.shape[1] <= 16 and mlp[1].eps == mlp[4].eps and mlp[1].momentum == mlp[4].momentum and x.shape[0] > 0)


def vfe_mlp(mlp, x, ps):
    bn1, bn2 = mlp[1], mlp[4]
    # Point rows in pillar (CSR) order: every point-sized tensor of the node (h1, y2, their gradients) is then walked
    # sequentially by the per-pillar kernels (seg_points = None selects their sorted mode): the pillar scatter-max and
    # its backward stream contiguous rows instead of chasing seg_points[k] -> row.  x is 10/11 floats per point, so this
    # one gather is cheap; the per-pillar results
    # do not depend on the row order (BatchNorm sums change in the last bits only).
    x = x.index_select(0, ps.seg_points)
    out = VfeMlpFunction.apply(x, ps.seg_offsets, None, ps.n_pillars, bn1.eps, bn1.momentum, bn1.running_mean,
                               bn1.running_var, bn2.running_mean, bn2.running_var, mlp[0].weight, bn1.weight, bn1.bias,
                               mlp[3].weight, bn2.weight, bn2.bias)
    bn1.num_batches_tracked += 1
    bn2.num_batches_tracked += 1
    return out


class SparseConvFunction(torch.autograd.Function):
    """3x3 sparse conv as gather -> ONE GEMM (spconv SubMConv2d / SparseConv2d, spconv_utils.py:37-56),
    manual backward: dW = dy^T col, dcol = dy W, dx = transposed gather (no atomics).  In the bf16
    configuration the gathered (N, 9*C_in) operand, dy and dcol are bf16 (written directly by the
    gather kernel / the GEMM), accumulation and outputs are fp32."""

    @staticmethod
    @_ops._fwd
    def forward(ctx, x, weight, fwd_map, bwd_map, mirror):
        x = x.contiguous()
        w = _gw(weight).view(weight.shape[0], -1)            # (C_out, 9*C_in)
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
        dw = gemm(dyg.t(), col, wgrad=True).view(ctx.wshape)
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
        ctx.save_for_backward(y, beta, gamma, mean, rstd, shift)    # the ReLU mask is rebuilt from y: `out` is not kept
        ctx.count, ctx.relu = count, relu
        return out, bg

    @staticmethod
    @_ops._bwd
    def backward(ctx, dout, dbg):
        y, beta, gamma, mean, rstd, shift = ctx.saved_tensors
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
        L.check(lib.gdmae_batchnorm_relu_bwd(L.P(y), L.P(beta), L.P(dout.contiguous()), L.P(gamma), L.P(mean), L.P(rstd), L.i64(N), C,
                                             ctypes.c_double(ctx.count), int(ctx.relu), L.P(e_db), L.P(e_dg), L.P(dy), None, L.P(dgamma),
                                             L.P(dbeta), L.P(ws), ctypes.c_size_t(ws.numel()), L.stream()),
                "gdmae_batchnorm_relu_bwd")
        return dy, dgamma, dbeta, None, None, None, None, None, None


_DTC = {torch.float32: 0, torch.bfloat16: 1}      # dtype codes of the C ABI
# pre-BatchNorm deconvolution output u of the deblocks: bf16 straight from the GEMM epilogue, like every other GEMM output of
# the bf16 configuration that only feeds a row kernel (a, h, f, VFE y2): the BatchNorm statistics are taken in fp32 from those
# values.  Five passes over the 135 M-element tensor move half the bytes (r2: -0.14 ms per step).  GDMAE_DEBLOCK_U16=0 keeps
# u in fp32.
DEBLOCK_U_DTYPE = torch.float32 if os.environ.get("GDMAE_DEBLOCK_U16", "1") == "0" else torch.bfloat16


class DeblockRowsFunction(torch.autograd.Function):
    """One decoder deblock on its sparse rows in the bf16 configuration: ConvTranspose2d(k = stride) as ONE GEMM
    (N, C_in) x (C_in, k*k*C_out) on the own tcgen05 kernel + training-mode BatchNorm2d + ReLU over the N*k*k produced rows
    (statistics over ``count`` cells, the others are zeros) - spt_backbone_mae.py:31-44, 125-132.  One autograd node so
    that the BatchNorm backward can hand its gradient to the two backward GEMMs as bf16 (no fp32 round trip, no cast pass):
    dx = du W^T, dW = x^T du (K = rows, split over the SMs).  Returns (rows (N*k*k, C_out) fp32, bg (C_out))."""

    @staticmethod
    def forward(ctx, x, weight, k, gamma, beta, running_mean, running_var, momentum, eps, count):
        x = x.contiguous()
        N, C_in = x.shape
        c_out = weight.shape[1]
        dev = x.device
        xg = _LAST_OUT[1] if _LAST_OUT[0] is x else x.to(BF16)
        wg = _gw(weight).permute(0, 2, 3, 1).reshape(C_in, k * k * c_out).contiguous()       # (C_in, [a, b, c_out]) bf16
        u = tc_gemm(xg, wg, out_dtype=DEBLOCK_U_DTYPE).view(N * k * k, c_out)
        # the rows only feed the bf16 dense map (ops.DenseFill): they leave the BatchNorm pass as bf16 - the one rounding
        # the map's store would have applied anyway
        out = torch.empty(u.shape, dtype=BF16, device=dev)
        mean = torch.empty((c_out,), dtype=F32, device=dev)
        rstd = torch.empty((c_out,), dtype=F32, device=dev)
        lib = L.lib()
        ws = L.workspace(lib.gdmae_batchnorm_workspace_bytes(c_out), dev)
        L.check(lib.gdmae_batchnorm_relu_fwd_t(L.P(u), _DTC[u.dtype], L.P(gamma), L.P(beta), L.i64(u.shape[0]), c_out, ctypes.c_double(count),
                                               L.f32(eps), L.f32(momentum), 1, L.P(out), 1, L.P(mean), L.P(rstd), L.P(running_mean),
                                               L.P(running_var), L.P(ws), ctypes.c_size_t(ws.numel()), L.stream()),
                "gdmae_batchnorm_relu_fwd_t")
        shift = beta - mean * rstd * gamma
        ctx.save_for_backward(xg, wg, u, beta, gamma, mean, rstd, shift)
        ctx.count, ctx.k, ctx.wshape = count, k, weight.shape
        return out, torch.relu(shift)

    @staticmethod
    def backward(ctx, dout, dbg):
        xg, wg, u, beta, gamma, mean, rstd, shift = ctx.saved_tensors
        R, c_out = u.shape
        N, C_in = xg.shape
        k = ctx.k
        dev = u.device
        e_db = e_dg = None
        if ctx.count > R and dbg is not None:
            e_db = (dbg * (shift > 0)).contiguous().float()
            e_dg = (e_db * (-mean * rstd)).contiguous()
        dout = dout.contiguous()          # bf16 rows gathered from d(map) by ops.DenseFill (fp32 from any other consumer)
        du16 = torch.empty((R, c_out), dtype=BF16, device=dev)
        dgamma = torch.empty((c_out,), dtype=F32, device=dev)
        dbeta = torch.empty((c_out,), dtype=F32, device=dev)
        lib = L.lib()
        ws = L.workspace(lib.gdmae_batchnorm_workspace_bytes(c_out), dev)
        L.check(lib.gdmae_batchnorm_relu_bwd_t(L.P(u), _DTC[u.dtype], L.P(beta), L.P(dout), _DTC[dout.dtype], L.P(gamma), L.P(mean), L.P(rstd),
                                               L.i64(R), c_out, ctypes.c_double(ctx.count), 1, L.P(e_db), L.P(e_dg), L.P(du16), 1,
                                               L.P(dgamma), L.P(dbeta), L.P(ws), ctypes.c_size_t(ws.numel()), L.stream()),
                "gdmae_batchnorm_relu_bwd_t")
        du16 = du16.view(N, k * k * c_out)
        dx = tc_gemm(du16, wg.t(), out_dtype=F32) if ctx.needs_input_grad[0] else None
        dw = tc_gemm(xg.t(), du16, out_dtype=F32, split_k=True)                               # (C_in, k*k*c_out)
        dw = dw.view(C_in, k, k, c_out).permute(0, 3, 1, 2)
        return dx, dw, None, dgamma, dbeta, None, None, None, None, None


class SparseConvBNReLUFunction(torch.autograd.Function):
    """post_act_block of the bf16 configuration as ONE autograd node (pcdet/utils/spconv_utils.py:37-56: SubMConv2d /
    SparseConv2d + BatchNorm1d + ReLU): gather -> GEMM (own tcgen05 kernel, pre-BatchNorm output y as bf16) -> BatchNorm
    statistics + apply; backward: the BatchNorm backward hands its gradient to the two backward GEMMs as bf16 (no fp32
    dy, no cast pass), dx by the transposed gather.  y (bf16) is the only activation kept besides the im2col operand."""

    @staticmethod
    @_ops._fwd
    def forward(ctx, x, weight, fwd_map, bwd_map, mirror, gamma, beta, running_mean, running_var, momentum, eps):
        x = x.contiguous()
        w = _gw(weight).view(weight.shape[0], -1)            # (C_out, 9*C_in) bf16
        col = _ops.gather_rows(x, fwd_map, BF16)
        y = gemm(col, w.t(), out_dtype=BF16)                 # (N, C_out)
        N, C = y.shape
        dev = y.device
        out = torch.empty((N, C), dtype=F32, device=dev)     # feeds the fp32 residual stream of the encoder layers
        mean = torch.empty((C,), dtype=F32, device=dev)
        rstd = torch.empty((C,), dtype=F32, device=dev)
        lib = L.lib()
        ws = L.workspace(lib.gdmae_batchnorm_workspace_bytes(C), dev)
        L.check(lib.gdmae_batchnorm_relu_fwd_t(L.P(y), 1, L.P(gamma), L.P(beta), L.i64(N), C, ctypes.c_double(float(N)), L.f32(eps),
                                               L.f32(momentum), 1, L.P(out), 0, L.P(mean), L.P(rstd), L.P(running_mean),
                                               L.P(running_var), L.P(ws), ctypes.c_size_t(ws.numel()), L.stream()),
                "gdmae_batchnorm_relu_fwd_t")
        ctx.save_for_backward(col, w, bwd_map, y, beta, gamma, mean, rstd)
        ctx.mirror, ctx.n_src, ctx.wshape = mirror, x.shape[0], weight.shape
        return out

    @staticmethod
    @_ops._bwd
    def backward(ctx, dout):
        col, w, bwd_map, y, beta, gamma, mean, rstd = ctx.saved_tensors
        N, C = y.shape
        dev = y.device
        dout = dout.contiguous()
        dy16 = torch.empty((N, C), dtype=BF16, device=dev)
        dgamma = torch.empty((C,), dtype=F32, device=dev)
        dbeta = torch.empty((C,), dtype=F32, device=dev)
        lib = L.lib()
        ws = L.workspace(lib.gdmae_batchnorm_workspace_bytes(C), dev)
        L.check(lib.gdmae_batchnorm_relu_bwd_t(L.P(y), 1, L.P(beta), L.P(dout), _DTC[dout.dtype], L.P(gamma), L.P(mean), L.P(rstd), L.i64(N), C,
                                               ctypes.c_double(float(N)), 1, None, None, L.P(dy16), 1, L.P(dgamma), L.P(dbeta), L.P(ws),
                                               ctypes.c_size_t(ws.numel()), L.stream()),
                "gdmae_batchnorm_relu_bwd_t")
        dw = gemm(dy16.t(), col, wgrad=True).view(ctx.wshape)
        dcol = gemm(dy16, w, out_dtype=BF16)
        dx = _ops.gather_rows_transposed(dcol, bwd_map, ctx.n_src, ctx.mirror)
        return dx, dw, None, None, None, dgamma, dbeta, None, None, None, None


def sparse_conv_bn_relu_ok(conv, bn, x, training):
    """the fused node applies: bf16 configuration, training-mode statistics, channel counts the typed BatchNorm kernels take"""
    c = conv.out_channels
    return (GEMM_DTYPE == BF16 and TC_GEMM and training and bn.training and bn.track_running_stats and x.is_cuda and c % 64 == 0
            and 256 % (c // 8) == 0 and os.environ.get("GDMAE_FUSE_SPCONV_BN", "1") != "0")


def sparse_conv_bn_relu(conv, bn, feats, fwd_map, bwd_map, mirror):
    out = SparseConvBNReLUFunction.apply(feats, conv.weight, fwd_map, bwd_map, mirror, bn.weight, bn.bias, bn.running_mean,
                                         bn.running_var, bn.momentum, bn.eps)
    bn.num_batches_tracked += 1
    return out


def deblock_rows(deconv, bn, k, x, count):
    """ConvTranspose2d(k = stride) + BatchNorm2d(train) + ReLU on the sparse rows, bf16 configuration (see DeblockRowsFunction)"""
    out, bg = DeblockRowsFunction.apply(x, deconv.weight, k, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum, bn.eps,
                                        float(count))
    bn.num_batches_tracked += 1
    return out, bg


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


# ----------------------------------------------------------------------------- own tcgen05 GEMM (csrc/tc_gemm.cu)
class TcEpilogue(ctypes.Structure):
    """mirror of gdmae_tc_epilogue (include/gdmae_b200.h)"""
    _fields_ = [("mode", ctypes.c_int), ("bias", _VP), ("c2", _VP), ("ldc2", ctypes.c_int64), ("res", _VP), ("gamma", _VP),
                ("beta_ln", _VP), ("eps", ctypes.c_float), ("y32", _VP), ("y16", _VP), ("mean", _VP), ("rstd", _VP),
                ("h16", _VP), ("ldh", ctypes.c_int64), ("colsum", _VP), ("tok_info", _VP), ("lut", _VP), ("tau", _VP),
                ("tau_min", ctypes.c_float), ("lrr", _VP), ("plane0", ctypes.c_int)]


def tc_gemm(a, b, out=None, beta=0.0, out_dtype=F32, split_k=False, epilogue=None):
    """out (M,N) = a (M,K) @ b (K,N) on the tcgen05 / TMA kernel; a, b bf16 (any of the two row-major layouts each)."""
    assert a.dtype == BF16 and b.dtype == BF16
    M, K = a.shape
    N = b.shape[1]
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    a, ta, lda = _layout(a)
    b, tb, ldb = _layout(b)
    L.check(L.lib().gdmae_tc_gemm(ta, tb, L.i64(M), L.i64(N), L.i64(K), _ptr(a), L.i64(lda), _ptr(b), L.i64(ldb), _ptr(out),
                                  L.i64(out.stride(0)), _ops._DT[out.dtype], L.f32(beta), int(bool(split_k)),
                                  ctypes.byref(epilogue) if epilogue is not None else None, L.stream()), "gdmae_tc_gemm")
    return out
