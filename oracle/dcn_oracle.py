"""TEST INFRASTRUCTURE - torch-CPU restatement of (modulated) deformable convolution, differentiable
through autograd.  Only tests/ may import this.

Follows the sampling rule of the reference's kernels (paths relative to /root/reference/pcdet/ops/dcn/src):
  deformable_im2col_bilinear              deform_conv_cuda_kernel.cu:84-117   corners outside the plane contribute zero
  (modulated_)deformable_im2col_gpu_kernel  :195-244, 588-640                 centre outside (-1,H) x (-1,W) -> zero; x mask
  channel layout of offsets / masks       :214-231, 612-622                   group g, tap t: offsets g*2*K2 + 2t (+1), mask g*K2 + t
and the contraction deform_conv_cuda.cpp:190-259, 540-600 (per group: output = W . columns (+ bias)).
PARITY UNPINNED: the reference has no CPU path for this operator (deform_conv.py:38-40 raises NotImplementedError) and no
test or golden vector touches it; the algorithm restated here is the published one (Dai et al. 2017, Zhu et al. 2019)."""
import torch


def deform_conv2d(x, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1):
    B, C, H, W = x.shape
    Cout, _, kh, kw = weight.shape
    sh = sw = stride
    ph = pw = padding
    dh = dw = dilation
    Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
    K2, cpg = kh * kw, C // deformable_groups
    ho = torch.arange(Ho).view(1, Ho, 1).float()
    wo = torch.arange(Wo).view(1, 1, Wo).float()
    cols = []
    for g in range(deformable_groups):
        xg = x[:, g * cpg:(g + 1) * cpg].reshape(B, cpg, H * W)
        taps = []
        for t in range(K2):
            i, j = t // kw, t % kw
            h = ho * sh - ph + i * dh + offset[:, g * 2 * K2 + 2 * t]
            w = wo * sw - pw + j * dw + offset[:, g * 2 * K2 + 2 * t + 1]
            inside = ((h > -1) & (w > -1) & (h < H) & (w < W)).float()
            h0, w0 = torch.floor(h), torch.floor(w)
            lh, lw = h - h0, w - w0
            val = 0
            for (hc, wc, wt) in ((h0, w0, (1 - lh) * (1 - lw)), (h0, w0 + 1, (1 - lh) * lw), (h0 + 1, w0, lh * (1 - lw)), (h0 + 1, w0 + 1, lh * lw)):
                ok = ((hc >= 0) & (hc <= H - 1) & (wc >= 0) & (wc <= W - 1)).float()
                idx = (hc.clamp(0, H - 1) * W + wc.clamp(0, W - 1)).long().view(B, 1, Ho * Wo).expand(B, cpg, Ho * Wo)
                val = val + torch.gather(xg, 2, idx) * (wt * ok * inside).view(B, 1, Ho * Wo)
            if mask is not None:
                val = val * mask[:, g * K2 + t].reshape(B, 1, Ho * Wo)
            taps.append(val)                                  # (B, cpg, P)
        cols.append(torch.stack(taps, 2))                     # (B, cpg, K2, P)
    col = torch.cat(cols, 1).reshape(B, C * K2, Ho * Wo)
    w = weight.reshape(groups, Cout // groups, -1)
    out = torch.matmul(w.unsqueeze(0), col.view(B, groups, C * K2 // groups, Ho * Wo)).reshape(B, Cout, Ho, Wo)
    if bias is not None:
        out = out + bias.view(1, -1, 1, 1)
    return out
