"""Mirror of pcdet/ops/dcn/deform_conv.py:1-337 over the B200 sampling kernels (csrc/dcn.cu).

``deform_conv_cuda`` keeps the reference's five pybind names and argument lists (pcdet/ops/dcn/src/deform_conv_cuda.cpp:
687-701; callers deform_conv.py:45-95 and :143-165): results are written in place into the tensors the caller passes, the
scratch buffers (``columns`` / ``ones`` / ``bufs``) are accepted and ignored (the columns live in a buffer of the call,
laid out per image), CPU tensors raise NotImplementedError like the reference's Functions do.  Sampling / scatter /
coordinate-gradient run in csrc/dcn.cu; the contractions with the weights are library GEMMs (torch.matmul).  No GD-MAE config
executes this operator; it completes the pcdet/ops API surface that north_star names."""
import math

import torch
import torch.nn as nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

from .... import _lib as L


def _geom(x, kh, kw, ph, pw, sh, sw, dh, dw, dg):
    B, C, H, W = x.shape
    Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
    return L.iarr([B, C, H, W, kh, kw, ph, pw, sh, sw, dh, dw, dg]), Ho, Wo


def _chk(*ts):
    for t in ts:
        if t is not None and (not t.is_cuda or t.dtype != torch.float32):
            raise NotImplementedError("deform_conv_cuda takes float32 CUDA tensors")


def _im2col(x, offset, mask, g, Ho, Wo, K2):
    col = torch.empty((x.shape[0], x.shape[1] * K2, Ho * Wo), dtype=torch.float32, device=x.device)
    L.check(L.lib().gdmae_deform_im2col(L.P(x.contiguous()), L.P(offset.contiguous()), L.P(None if mask is None else mask.contiguous()), g,
                                        L.P(col), L.stream()), "gdmae_deform_im2col")
    return col


def _conv_from_col(col, weight, group):
    """col (B, C*K2, P), weight (Cout, C/group, kh, kw) -> (B, Cout, P)"""
    B, CK, P = col.shape
    Cout = weight.shape[0]
    w = weight.reshape(group, Cout // group, -1)
    return torch.matmul(w.unsqueeze(0), col.view(B, group, CK // group, P)).reshape(B, Cout, P)


def _forward(x, weight, offset, mask, output, kh, kw, sh, sw, ph, pw, dh, dw, group, dg):
    _chk(x, weight, offset, mask, output)
    g, Ho, Wo = _geom(x, kh, kw, ph, pw, sh, sw, dh, dw, dg)
    col = _im2col(x, offset, mask, g, Ho, Wo, kh * kw)
    output.copy_(_conv_from_col(col, weight, group).view(x.shape[0], weight.shape[0], Ho, Wo))
    return col


def _backward(x, weight, offset, mask, grad_output, grad_input, grad_offset, grad_mask, grad_weight, scale, kh, kw, sh, sw, ph, pw, dh, dw,
              group, dg):
    _chk(x, weight, offset, mask, grad_output)
    B, C = x.shape[0], x.shape[1]
    g, Ho, Wo = _geom(x, kh, kw, ph, pw, sh, sw, dh, dw, dg)
    K2, P, Cout = kh * kw, Ho * Wo, weight.shape[0]
    go = grad_output.contiguous().view(B, group, Cout // group, P)
    w = weight.reshape(group, Cout // group, -1)
    lib = L.lib()
    if grad_input is not None or grad_offset is not None:
        dcol = torch.matmul(w.transpose(1, 2).unsqueeze(0), go).reshape(B, C * K2, P).contiguous()
        if grad_offset is not None:
            L.check(lib.gdmae_deform_col2im_coord(L.P(dcol), L.P(x.contiguous()), L.P(offset.contiguous()),
                                                  L.P(None if mask is None else mask.contiguous()), g, L.P(grad_offset),
                                                  L.P(grad_mask if mask is not None else None), L.stream()), "gdmae_deform_col2im_coord")
        if grad_input is not None:
            L.check(lib.gdmae_deform_col2im(L.P(dcol), L.P(offset.contiguous()), L.P(None if mask is None else mask.contiguous()), g,
                                            L.P(grad_input), L.stream()), "gdmae_deform_col2im")
    if grad_weight is not None:
        col = _im2col(x, offset, mask, g, Ho, Wo, K2).view(B, group, C * K2 // group, P)
        gw = torch.matmul(go, col.transpose(2, 3)).sum(0)                       # (group, Cout/group, C/group*K2)
        grad_weight.add_(gw.reshape(grad_weight.shape), alpha=scale)


class _DeformConvCuda:
    """the five entry points of deform_conv_cuda.cpp:687-701 (same names, argument order and in-place outputs)"""

    @staticmethod
    def deform_conv_forward_cuda(input, weight, offset, output, columns, ones, kW, kH, dW, dH, padW, padH, dilationW, dilationH, group,
                                 deformable_group, im2col_step):
        _forward(input, weight, offset, None, output, kH, kW, dH, dW, padH, padW, dilationH, dilationW, group, deformable_group)
        return 1

    @staticmethod
    def deform_conv_backward_input_cuda(input, offset, gradOutput, gradInput, gradOffset, weight, columns, kW, kH, dW, dH, padW, padH,
                                        dilationW, dilationH, group, deformable_group, im2col_step):
        _backward(input, weight, offset, None, gradOutput, gradInput, gradOffset, None, None, 1.0, kH, kW, dH, dW, padH, padW, dilationH,
                  dilationW, group, deformable_group)
        return 1

    @staticmethod
    def deform_conv_backward_parameters_cuda(input, offset, gradOutput, gradWeight, columns, ones, kW, kH, dW, dH, padW, padH, dilationW,
                                             dilationH, group, deformable_group, scale, im2col_step):
        _backward(input, gradWeight, offset, None, gradOutput, None, None, None, gradWeight, float(scale), kH, kW, dH, dW, padH, padW,
                  dilationH, dilationW, group, deformable_group)
        return 1

    @staticmethod
    def modulated_deform_conv_cuda_forward(input, weight, bias, ones, offset, mask, output, columns, kernel_h, kernel_w, stride_h, stride_w,
                                           pad_h, pad_w, dilation_h, dilation_w, group, deformable_group, with_bias):
        _forward(input, weight, offset, mask, output, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group,
                 deformable_group)
        if with_bias:
            output.add_(bias.view(1, -1, 1, 1))

    @staticmethod
    def modulated_deform_conv_cuda_backward(input, weight, bias, ones, offset, mask, columns, grad_input, grad_weight, grad_bias, grad_offset,
                                            grad_mask, grad_output, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h,
                                            dilation_w, group, deformable_group, with_bias):
        _backward(input, weight, offset, mask, grad_output, grad_input, grad_offset, grad_mask, grad_weight, 1.0, kernel_h, kernel_w,
                  stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group, deformable_group)
        if with_bias:
            grad_bias.add_(grad_output.sum(dim=(0, 2, 3)))


deform_conv_cuda = _DeformConvCuda()


class DeformConvFunction(Function):
    @staticmethod
    def forward(ctx, input, offset, weight, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1, im2col_step=64):
        if input is not None and input.dim() != 4:
            raise ValueError("Expected 4D tensor as input, got {}D tensor instead.".format(input.dim()))
        ctx.stride, ctx.padding, ctx.dilation = _pair(stride), _pair(padding), _pair(dilation)
        ctx.groups, ctx.deformable_groups, ctx.im2col_step = groups, deformable_groups, im2col_step
        ctx.save_for_backward(input, offset, weight)
        output = input.new_empty(DeformConvFunction._output_size(input, weight, ctx.padding, ctx.dilation, ctx.stride))
        ctx.bufs_ = [input.new_empty(0), input.new_empty(0)]
        if not input.is_cuda:
            raise NotImplementedError
        cur_im2col_step = min(ctx.im2col_step, input.shape[0])
        assert (input.shape[0] % cur_im2col_step) == 0, 'im2col step must divide batchsize'
        deform_conv_cuda.deform_conv_forward_cuda(input, weight, offset, output, ctx.bufs_[0], ctx.bufs_[1], weight.size(3), weight.size(2),
                                                  ctx.stride[1], ctx.stride[0], ctx.padding[1], ctx.padding[0], ctx.dilation[1],
                                                  ctx.dilation[0], ctx.groups, ctx.deformable_groups, cur_im2col_step)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        input, offset, weight = ctx.saved_tensors
        grad_input = grad_offset = grad_weight = None
        if not grad_output.is_cuda:
            raise NotImplementedError
        cur_im2col_step = min(ctx.im2col_step, input.shape[0])
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            grad_input, grad_offset = torch.zeros_like(input), torch.zeros_like(offset)
            deform_conv_cuda.deform_conv_backward_input_cuda(input, offset, grad_output, grad_input, grad_offset, weight, ctx.bufs_[0],
                                                             weight.size(3), weight.size(2), ctx.stride[1], ctx.stride[0], ctx.padding[1],
                                                             ctx.padding[0], ctx.dilation[1], ctx.dilation[0], ctx.groups,
                                                             ctx.deformable_groups, cur_im2col_step)
        if ctx.needs_input_grad[2]:
            grad_weight = torch.zeros_like(weight)
            deform_conv_cuda.deform_conv_backward_parameters_cuda(input, offset, grad_output, grad_weight, ctx.bufs_[0], ctx.bufs_[1],
                                                                  weight.size(3), weight.size(2), ctx.stride[1], ctx.stride[0],
                                                                  ctx.padding[1], ctx.padding[0], ctx.dilation[1], ctx.dilation[0],
                                                                  ctx.groups, ctx.deformable_groups, 1, cur_im2col_step)
        return grad_input, grad_offset, grad_weight, None, None, None, None, None

    @staticmethod
    def _output_size(input, weight, padding, dilation, stride):
        size = (input.size(0), weight.size(0))
        for d in range(input.dim() - 2):
            kernel = dilation[d] * (weight.size(d + 2) - 1) + 1
            size += ((input.size(d + 2) + 2 * padding[d] - kernel) // stride[d] + 1,)
        if not all(s > 0 for s in size):
            raise ValueError("convolution input is too small (output would be {})".format('x'.join(map(str, size))))
        return size


class ModulatedDeformConvFunction(Function):
    @staticmethod
    def forward(ctx, input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1):
        ctx.stride, ctx.padding, ctx.dilation, ctx.groups, ctx.deformable_groups = stride, padding, dilation, groups, deformable_groups
        ctx.with_bias = bias is not None
        if not ctx.with_bias:
            bias = input.new_empty(1)
        if not input.is_cuda:
            raise NotImplementedError
        if weight.requires_grad or mask.requires_grad or offset.requires_grad or input.requires_grad:
            ctx.save_for_backward(input, offset, mask, weight, bias)
        output = input.new_empty(ModulatedDeformConvFunction._infer_shape(ctx, input, weight))
        ctx._bufs = [input.new_empty(0), input.new_empty(0)]
        deform_conv_cuda.modulated_deform_conv_cuda_forward(input, weight, bias, ctx._bufs[0], offset, mask, output, ctx._bufs[1],
                                                            weight.shape[2], weight.shape[3], ctx.stride, ctx.stride, ctx.padding,
                                                            ctx.padding, ctx.dilation, ctx.dilation, ctx.groups, ctx.deformable_groups,
                                                            ctx.with_bias)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        if not grad_output.is_cuda:
            raise NotImplementedError
        input, offset, mask, weight, bias = ctx.saved_tensors
        grad_input, grad_offset, grad_mask = torch.zeros_like(input), torch.zeros_like(offset), torch.zeros_like(mask)
        grad_weight, grad_bias = torch.zeros_like(weight), torch.zeros_like(bias)
        deform_conv_cuda.modulated_deform_conv_cuda_backward(input, weight, bias, ctx._bufs[0], offset, mask, ctx._bufs[1], grad_input,
                                                             grad_weight, grad_bias, grad_offset, grad_mask, grad_output, weight.shape[2],
                                                             weight.shape[3], ctx.stride, ctx.stride, ctx.padding, ctx.padding,
                                                             ctx.dilation, ctx.dilation, ctx.groups, ctx.deformable_groups, ctx.with_bias)
        if not ctx.with_bias:
            grad_bias = None
        return grad_input, grad_offset, grad_mask, grad_weight, grad_bias, None, None, None, None, None

    @staticmethod
    def _infer_shape(ctx, input, weight):
        n, channels_out = input.size(0), weight.size(0)
        height, width = input.shape[2:4]
        kernel_h, kernel_w = weight.shape[2:4]
        height_out = (height + 2 * ctx.padding - (ctx.dilation * (kernel_h - 1) + 1)) // ctx.stride + 1
        width_out = (width + 2 * ctx.padding - (ctx.dilation * (kernel_w - 1) + 1)) // ctx.stride + 1
        return n, channels_out, height_out, width_out


deform_conv = DeformConvFunction.apply
modulated_deform_conv = ModulatedDeformConvFunction.apply


class DeformConv(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1, bias=False):
        super().__init__()
        assert not bias
        assert in_channels % groups == 0, 'in_channels {} cannot be divisible by groups {}'.format(in_channels, groups)
        assert out_channels % groups == 0, 'out_channels {} cannot be divisible by groups {}'.format(out_channels, groups)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding, self.dilation = _pair(kernel_size), _pair(stride), _pair(padding), _pair(dilation)
        self.groups, self.deformable_groups = groups, deformable_groups
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // self.groups, *self.kernel_size))
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1. / math.sqrt(self.in_channels * self.kernel_size[0] * self.kernel_size[1])
        self.weight.data.uniform_(-stdv, stdv)

    def forward(self, x, offset):
        return deform_conv(x, offset, self.weight, self.stride, self.padding, self.dilation, self.groups, self.deformable_groups)


class DeformConvPack(DeformConv):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.conv_offset = nn.Conv2d(self.in_channels, self.deformable_groups * 2 * self.kernel_size[0] * self.kernel_size[1],
                                     kernel_size=self.kernel_size, stride=_pair(self.stride), padding=_pair(self.padding), bias=True)
        self.init_offset()

    def init_offset(self):
        self.conv_offset.weight.data.zero_()
        self.conv_offset.bias.data.zero_()

    def forward(self, x):
        return deform_conv(x, self.conv_offset(x), self.weight, self.stride, self.padding, self.dilation, self.groups, self.deformable_groups)


class ModulatedDeformConv(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1, bias=True):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, _pair(kernel_size)
        self.stride, self.padding, self.dilation, self.groups = stride, padding, dilation, groups
        self.deformable_groups, self.with_bias = deformable_groups, bias
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // groups, *self.kernel_size))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1. / math.sqrt(self.in_channels * self.kernel_size[0] * self.kernel_size[1])
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.zero_()

    def forward(self, x, offset, mask):
        return modulated_deform_conv(x, offset, mask, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups,
                                     self.deformable_groups)


class ModulatedDeformConvPack(ModulatedDeformConv):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.conv_offset_mask = nn.Conv2d(self.in_channels, self.deformable_groups * 3 * self.kernel_size[0] * self.kernel_size[1],
                                          kernel_size=self.kernel_size, stride=_pair(self.stride), padding=_pair(self.padding), bias=True)
        self.init_offset()

    def init_offset(self):
        self.conv_offset_mask.weight.data.zero_()
        self.conv_offset_mask.bias.data.zero_()

    def forward(self, x):
        o1, o2, mask = torch.chunk(self.conv_offset_mask(x), 3, dim=1)
        return modulated_deform_conv(x, torch.cat((o1, o2), dim=1), torch.sigmoid(mask), self.weight, self.bias, self.stride, self.padding,
                                     self.dilation, self.groups, self.deformable_groups)
