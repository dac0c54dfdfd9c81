"""Mirror of pcdet/ops/iou3d_nms/iou3d_nms_utils.py:12-116 over the B200 kernels (csrc/iou3d_nms.cu).

``iou3d_nms_cuda`` keeps the reference's pybind names and calling convention (pcdet/ops/iou3d_nms/src/iou3d_nms_api.cpp:11-17):
outputs are written in place into caller-allocated tensors, the pair functions return 1, the NMS functions return the
number of kept boxes and fill the CPU ``keep`` tensor the caller passes.  Unlike the reference (exit(-1) on a CPU tensor,
iou3d_nms.cpp:13-24) a wrong device raises."""
import ctypes

import torch

from .... import _lib as L


def _check(t, name, cuda=True):
    if cuda and not t.is_cuda:
        raise L.GdmaeError(f"{name} must be a CUDA tensor (gd-mae_b200 has no CPU fallback for this op)")
    if not t.is_contiguous() or t.dtype != torch.float32:
        raise L.GdmaeError(f"{name} must be a contiguous float32 tensor")


class _Iou3dNmsCuda:
    @staticmethod
    def boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
        for t, n in ((boxes_a, "boxes_a"), (boxes_b, "boxes_b"), (ans_overlap, "ans_overlap")):
            _check(t, n)
        L.check(L.lib().gdmae_boxes_overlap_bev(L.P(boxes_a), boxes_a.shape[0], L.P(boxes_b), boxes_b.shape[0], L.P(ans_overlap),
                                                L.stream()), "gdmae_boxes_overlap_bev")
        return 1

    @staticmethod
    def boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou):
        for t, n in ((boxes_a, "boxes_a"), (boxes_b, "boxes_b"), (ans_iou, "ans_iou")):
            _check(t, n)
        L.check(L.lib().gdmae_boxes_iou_bev(L.P(boxes_a), boxes_a.shape[0], L.P(boxes_b), boxes_b.shape[0], L.P(ans_iou), L.stream()),
                "gdmae_boxes_iou_bev")
        return 1

    @staticmethod
    def _nms(fn_name, boxes, keep, thresh):
        _check(boxes, "boxes")
        n = boxes.shape[0]
        lib = L.lib()
        ws = L.workspace(lib.gdmae_nms_workspace_bytes(n), boxes.device)
        keep_dev = torch.empty((max(n, 1),), dtype=torch.int64, device=boxes.device)
        num = torch.zeros((1,), dtype=torch.int32, device=boxes.device)
        L.check(getattr(lib, fn_name)(L.P(boxes), n, L.f32(thresh), L.P(ws), ctypes.c_size_t(ws.numel()), L.P(keep_dev), L.P(num),
                                      L.stream()), fn_name)
        num_out = int(num.item())                      # the one host read the reference's interface implies (it returns the count)
        keep[:num_out] = keep_dev[:num_out].to(keep.device)
        return num_out

    @staticmethod
    def nms_gpu(boxes, keep, nms_overlap_thresh):
        return _Iou3dNmsCuda._nms("gdmae_nms_bev", boxes, keep, nms_overlap_thresh)

    @staticmethod
    def nms_normal_gpu(boxes, keep, nms_overlap_thresh):
        return _Iou3dNmsCuda._nms("gdmae_nms_normal", boxes, keep, nms_overlap_thresh)

    @staticmethod
    def boxes_iou_bev_cpu(boxes_a, boxes_b, ans_iou):
        for t, n in ((boxes_a, "boxes_a"), (boxes_b, "boxes_b"), (ans_iou, "ans_iou")):
            if t.is_cuda:
                raise L.GdmaeError(f"{n}: boxes_iou_bev_cpu takes CPU tensors")
            _check(t, n, cuda=False)
        L.check(L.lib().gdmae_boxes_iou_bev_cpu(ctypes.c_void_p(boxes_a.data_ptr()), boxes_a.shape[0], ctypes.c_void_p(boxes_b.data_ptr()),
                                                boxes_b.shape[0], ctypes.c_void_p(ans_iou.data_ptr())), "gdmae_boxes_iou_bev_cpu")
        return 1


iou3d_nms_cuda = _Iou3dNmsCuda()


def boxes_bev_iou_cpu(boxes_a, boxes_b):
    """(N,7), (M,7) CPU tensors or numpy arrays -> (N,M) rotated BEV IoU (iou3d_nms_utils.py:12-28)"""
    import numpy as np
    is_numpy = isinstance(boxes_a, np.ndarray)
    a = torch.from_numpy(boxes_a).float() if isinstance(boxes_a, np.ndarray) else boxes_a
    b = torch.from_numpy(boxes_b).float() if isinstance(boxes_b, np.ndarray) else boxes_b
    assert not (a.is_cuda or b.is_cuda), 'Only support CPU tensors'
    assert a.shape[1] == 7 and b.shape[1] == 7
    ans_iou = a.new_zeros(torch.Size((a.shape[0], b.shape[0])))
    iou3d_nms_cuda.boxes_iou_bev_cpu(a.contiguous(), b.contiguous(), ans_iou)
    return ans_iou.numpy() if is_numpy else ans_iou


def boxes_iou_bev(boxes_a, boxes_b):
    """(N,7), (M,7) [x, y, z, dx, dy, dz, heading] -> (N,M) rotated BEV IoU (iou3d_nms_utils.py:31-45)"""
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    ans_iou = torch.zeros((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32, device=boxes_a.device)
    iou3d_nms_cuda.boxes_iou_bev_gpu(boxes_a.contiguous(), boxes_b.contiguous(), ans_iou)
    return ans_iou


def boxes_iou3d_gpu(boxes_a, boxes_b):
    """(N,7), (M,7) -> (N,M) 3-D IoU: BEV overlap x height overlap over the union volume (iou3d_nms_utils.py:48-79)"""
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    a_top, a_bottom = (boxes_a[:, 2] + boxes_a[:, 5] / 2).view(-1, 1), (boxes_a[:, 2] - boxes_a[:, 5] / 2).view(-1, 1)
    b_top, b_bottom = (boxes_b[:, 2] + boxes_b[:, 5] / 2).view(1, -1), (boxes_b[:, 2] - boxes_b[:, 5] / 2).view(1, -1)
    overlaps_bev = torch.zeros((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32, device=boxes_a.device)
    iou3d_nms_cuda.boxes_overlap_bev_gpu(boxes_a.contiguous(), boxes_b.contiguous(), overlaps_bev)
    overlaps_h = torch.clamp(torch.min(a_top, b_top) - torch.max(a_bottom, b_bottom), min=0)
    overlaps_3d = overlaps_bev * overlaps_h
    vol_a = (boxes_a[:, 3] * boxes_a[:, 4] * boxes_a[:, 5]).view(-1, 1)
    vol_b = (boxes_b[:, 3] * boxes_b[:, 4] * boxes_b[:, 5]).view(1, -1)
    return overlaps_3d / torch.clamp(vol_a + vol_b - overlaps_3d, min=1e-6)


def nms_gpu(boxes, scores, thresh, pre_maxsize=None, **kwargs):
    """rotated NMS; -> (indices of the kept boxes in descending score order, None) (iou3d_nms_utils.py:82-98)"""
    assert boxes.shape[1] == 7
    order = scores.sort(0, descending=True)[1]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    boxes = boxes[order].contiguous()
    keep = torch.LongTensor(boxes.size(0))
    num_out = iou3d_nms_cuda.nms_gpu(boxes, keep, thresh)
    return order[keep[:num_out].to(boxes.device)].contiguous(), None


def nms_normal_gpu(boxes, scores, thresh, **kwargs):
    """axis-aligned NMS (headings ignored) (iou3d_nms_utils.py:101-116)"""
    assert boxes.shape[1] == 7
    order = scores.sort(0, descending=True)[1]
    boxes = boxes[order].contiguous()
    keep = torch.LongTensor(boxes.size(0))
    num_out = iou3d_nms_cuda.nms_normal_gpu(boxes, keep, thresh)
    return order[keep[:num_out].to(boxes.device)].contiguous(), None
