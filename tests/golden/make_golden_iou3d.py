"""tests/golden/iou3d_kat.npz: known answers of the rotated BEV IoU written by the REFERENCE's own CPU implementation
(pcdet/ops/iou3d_nms/src/iou3d_cpu.cpp boxes_iou_bev_cpu, compiled from where it lies by oracle/build_oracle.py).
Run in the build container only:  python tests/golden/make_golden_iou3d.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import build_oracle, iou3d_oracle as IO  # noqa: E402


def special_boxes():
    """hand-made cases: identical, nested, edge-sharing, disjoint, crossed at 90 degrees, tiny, corner-touching, 1e-2 margin"""
    b = [[0, 0, 0, 4, 2, 1.5, 0.0], [0, 0, 0, 4, 2, 1.5, 0.0], [0.5, 0.2, 0, 2, 1, 1.5, 0.3], [4, 0, 0, 4, 2, 1.5, 0.0],
         [10, 10, 0, 4, 2, 1.5, 1.0], [0, 0, 0, 4, 2, 1.5, np.pi / 2], [0, 0, 0, 0.1, 0.1, 1.0, 0.7], [4, 2, 0, 4, 2, 1.5, 0.0],
         [4.005, 0, 0, 4, 2, 1.5, 0.0], [0, 0, 0.5, 4, 2, 1.5, np.pi], [1, 1, 0, 3, 3, 2.0, np.pi / 4], [0, 0, 0, 4, 2, 1.5, 1e-4]]
    return np.array(b, dtype=np.float32)


if __name__ == "__main__":
    ref = build_oracle.load_ref()
    assert ref is not None, "needs /root/reference"
    out = {}
    cases = {"special": (special_boxes(), special_boxes()), "rand_a": (IO.random_boxes(150, 1), IO.random_boxes(130, 2)),
             "rand_self": (IO.random_boxes(200, 3, spread=8.0), IO.random_boxes(200, 3, spread=8.0))}
    for name, (a, b) in cases.items():
        ans = torch.zeros(a.shape[0], b.shape[0])
        ref.boxes_iou_bev_cpu(torch.from_numpy(a).contiguous(), torch.from_numpy(b).contiguous(), ans)
        out[name + ".a"], out[name + ".b"], out[name + ".iou"] = a, b, ans.numpy()
        print(name, a.shape, b.shape, "pairs with iou > 0:", int((ans > 0).sum()), "max", float(ans.max()))
    np.savez_compressed(os.path.join(HERE, "iou3d_kat.npz"), **out)
