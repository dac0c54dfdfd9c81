"""Places the UNMODIFIED reference files of the hot path under baseline/_ref/ (git-ignored; it travels to the GPU box with
gpurun, /root/reference does not), keeping their relative paths, so that bench.py's reference arm and cpu_baseline leg time the
reference's own Python on the box's host cores.  Nothing is edited: the files are
byte-for-byte copies, checked by size + sha256 in the manifest.  The third-party packages the files import (torch_scatter,
spconv, pytorch3d, the CUDA extension of sst_ops) stay the stand-ins of tests/golden/ref_harness.py.

Run in the build container (where /root/reference exists); __graft_entry__.build() calls it."""
import hashlib
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = [
    "pcdet/utils/common_utils.py", "pcdet/utils/spconv_utils.py",
    "pcdet/models/backbones_3d/vfe/vfe_template.py", "pcdet/models/backbones_3d/vfe/dyn_vfe.py",
    "pcdet/models/backbones_3d/spt_backbone.py", "pcdet/models/backbones_3d/spt_backbone_mae.py",
    "pcdet/models/model_utils/sst_utils.py", "pcdet/models/model_utils/sst_basic_block.py", "pcdet/models/model_utils/cosine_msa.py",
    "pcdet/models/model_utils/network_utils.py", "pcdet/ops/sst_ops/sst_ops_utils.py",
    "tools/train_utils/optimization/__init__.py", "tools/train_utils/optimization/fastai_optim.py",
    "tools/train_utils/optimization/learning_schedules_fastai.py",
    "tools/cfgs/waymo_models/gd_mae_ssl.yaml", "tools/cfgs/dataset_configs/waymo_dataset.yaml", "LICENSE",
]


def install(verbose=False):
    """-> DST when the reference files are in place (copied now or earlier), None when neither source nor copy exists"""
    if not os.path.isdir(SRC):
        return DST if os.path.exists(os.path.join(DST, "MANIFEST.json")) else None
    manifest = {}
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.exists(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        with open(d, "rb") as f:
            manifest[rel] = {"bytes": os.path.getsize(d), "sha256": hashlib.sha256(f.read()).hexdigest()}
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1)
    if verbose:
        print(f"{len(manifest)} reference files -> {DST}")
    return DST


if __name__ == "__main__":
    print(install(verbose=True))
