"""Host-side logic that needs no GPU: the device DataAugmentor's random draws, the WindowTable's lazily viewed arrays, the
stream-ownership walker of the prefetched index structures, and the bench.py reference arm's contract line."""
import json
import os
import subprocess
import sys

import numpy as np
import torch

from oracle import gdmae_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_device_augmentor_draws_the_reference_stream(golden):
    """the mirror draws flip / rotation / scaling per frame in the reference's order: same numpy seed -> the parameters the
    reference's DataAugmentor drew (tests/golden/augment_kat.npz); the permutation draw is interleaved like a dataset worker"""
    from gd_mae_b200 import config
    from gd_mae_b200.pcdet.datasets.augmentor.data_augmentor import DataAugmentor
    K = golden("augment_kat")
    acfg = config.to_attr({"DISABLE_AUG_LIST": ["placeholder"], "AUG_CONFIG_LIST": [
        {"NAME": "random_world_flip", "PROBABILITY": 0.5, "ALONG_AXIS_LIST": ["x", "y"]},
        {"NAME": "random_world_rotation", "PROBABILITY": 1.0, "WORLD_ROT_ANGLE": [-0.78539816, 0.78539816]},
        {"NAME": "random_world_scaling", "PROBABILITY": 1.0, "WORLD_SCALE_RANGE": [0.95, 1.05]}]})
    aug = DataAugmentor(None, acfg, ["Vehicle"])
    np.random.seed(1234)
    for f in range(3):
        fp = {}
        for a in aug.data_augmentor_queue:
            fp = a(frame_params=fp)
        assert ("x" in fp["random_world_flip"]) == bool(K[f"f{f}.flip_x"]) and ("y" in fp["random_world_flip"]) == bool(K[f"f{f}.flip_y"])
        assert fp["random_world_rotation"] == float(K[f"f{f}.rotation"]) and fp["random_world_scaling"] == float(K[f"f{f}.scaling"])
        assert np.array_equal(np.random.permutation(K[f"f{f}.points_in"].shape[0]), K[f"f{f}.perm"])
    bad = config.to_attr({"DISABLE_AUG_LIST": [], "AUG_CONFIG_LIST": [{"NAME": "gt_sampling"}]})
    try:
        DataAugmentor(None, bad, ["Vehicle"])
        assert False, "gt_sampling must raise: it is not on the pre-train path"
    except NotImplementedError:
        pass


def test_window_table_lazy_views():
    from gd_mae_b200 import ops
    t = ops.WindowTable()
    buf = torch.arange(96, dtype=torch.uint8)
    t._buf, t._views, t.N = buf, {}, 3
    t._fields = {"inner": (16, 12, torch.int32, (3,)), "win_mask": (32, 16, torch.int64, (2,))}
    assert t.inner.dtype == torch.int32 and t.inner.shape == (3,) and t.inner.data_ptr() == buf.data_ptr() + 16
    assert t.win_mask.shape == (2,) and t.inner is t.inner          # cached view
    try:
        t.no_such_field
        assert False
    except AttributeError:
        pass
    assert getattr(t, "_bin_units", None) is None                    # unset slot: falls through to the default


def test_record_stream_walker_visits_nested_structures():
    """GDMAE._record_stream must reach every tensor of the prefetched structures (dicts, lists, slot objects, namespaces)"""
    from types import SimpleNamespace
    from gd_mae_b200 import ops
    from gd_mae_b200.pcdet.models.detectors.gd_mae import GDMAE
    seen_tensors = []


    def mk():
        t = torch.zeros(2)
        seen_tensors.append(t)
        return t
    wt = ops.WindowTable()
    wt._buf, wt._fields, wt._views, wt.row_info, wt.pos_of_token, wt.N = mk(), {}, {"a": mk()}, mk(), mk(), 2
    ps = ops.PillarSet()
    ps.points, ps.n_points = mk(), 5
    struct = {"rank_grid": mk(), "win": [wt], "down": SimpleNamespace(indices=mk(), struct={"rank_grid": mk()})}
    bd = {"points": mk(), "pillar_set": ps, "mae_index": (mk(), SimpleNamespace(_struct=struct, indices=mk())), "batch_size": 2}
    visited = set()
    GDMAE._record_stream(bd, None, visited)        # CPU tensors: record_stream is skipped, the walk itself is what is tested
    assert all(id(t) in visited for t in seen_tensors), len(seen_tensors)


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the reference's own Python from baseline/_ref on the host cores - the oracle port only when
    those copies are absent -, bounded sample) - one step, checks the JSON keys"""
    env = dict(os.environ, OMP_NUM_THREADS="8")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "mae_pretrain_frames_per_sec" and line["unit"] == "frames/s"
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["higher_is_better"] is True
    have_ref = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "MANIFEST.json")) or os.path.isdir("/root/reference")
    assert line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    # other ranks of a torchrun launch exit 0 without work
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=120, env=dict(env, RANK="1", WORLD_SIZE="2"), cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""
