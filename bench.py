"""Benchmark of the MAE pre-train step (BASELINE.json metric: MAE-pretrain frames/sec, Waymo-shape
160k-point scenes; SRA HBM GB/s vs peak).

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--dtype bf16|tf32|fp32]

Own arm: one process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE), each rank trains on its
own B=8 synthetic frames per step (weak scaling, frame sharded, one NCCL all-reduce of the flat
gradient bucket per step).  ``value`` = frames/s with the inputs already resident in HBM;
``e2e`` = the same step driven through the public API with HOST (pinned) input batches, H2D copy
and a D2H read of the loss inside the timed region.
Reference arm (--impl reference): the reference's own Python for the path on the box's host cores -
the unmodified files under baseline/_ref/ (tools/install_reference.py; oracle port only if they are
absent) - on a bounded sample (B=1 frame per step) of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# Real batches differ in size from step to step (the e2e leg augments every batch: random rotation / scaling), so buffer sizes
# vary and the default caching allocator falls back to cudaMalloc / cudaFree - device-wide syncs of ~40 ms each (r2: e2e
# 25.1 ms/step with 2 cudaMallocs in the leg, 20.8 with none).  Expandable segments grow the pool by mapping pages instead.
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")

import numpy as np  # noqa: E402
import torch  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "waymo_gd_mae_ssl_pretrain_synthetic_160k_pt_B8_per_gpu"
B_PER_GPU = 8
# kinds of the CUDA-event spans recorded inside the C executors (csrc/common.cuh GdmaeSpan)
BYTES_DEF = {
    "sra_fwd": "N*d*(3*s_qkv + s_o) + N*32 per launch: q,k,v in, o and lse out (s = 2 bytes in the bf16 configuration; SURVEY.md 8d a18)",
    "sra_bwd": "N*d*(3*s + s + 3*s) + N*32 per launch: q,k,v,dO in, dq,dk,dv out, lse in",
    "pillar_scatter_max": "Np*C*s + Np*4 + M*C*4 per launch: point rows + segment index in, pillar maxima out (SURVEY.md 8d a6)",
    "tc_gemm": "M*K*2 + K*N*2 + M*N*s_out per launch (+ old C for C+=, + M*N*10 for the LayerNorm epilogue's residual in / fp32+bf16 rows "
               "out, + M*N*2 for the GELU modes' pre-activation); activation-bound shapes (K, N <= 768): arithmetic intensity <= 170 "
               "FLOP/B, below the 210 FLOP/B balance of the measured peaks, so the bound is HBM",
}
SETTLE_STEPS = 12
SPAN_NAMES = {0: "sra_fwd_d{d}", 1: "sra_bwd_d{d}", 2: "pillar_scatter_max_c{d}", 3: "tc_gemm_{d}"}
TC_GEMM_MODES = {0: "plain", 1: "gelu", 2: "ln", 3: "gelu_bwd", 4: "qkv_win", 5: "rows_win"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed regions by an in-process NVML thread (pynvml: three cheap
    driver queries every `period` seconds on this rank's own GPU).  r1 spawned one `nvidia-smi -lms 100` process per rank
    for this; eight of them polling the driver slowed the measured loop itself."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, dev, period=0.05):
        self.period, self.rows, self.h, self.nv, self.err = period, [], None, None, None
        self._stop = threading.Event()
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(dev).uuid)
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(dev.index or 0)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def _sample(self):
        nv = self.nv
        try:
            self.rows.append((float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)),
                              int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))))
        except Exception:
            try:
                self.rows.append((float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)),
                                  int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))))
            except Exception as e:  # pragma: no cover
                self.err = repr(e)

    def _run(self):
        while not self._stop.is_set():
            self._sample()
            self._stop.wait(self.period)

    def start(self):
        if self.h is None or os.environ.get("GDMAE_BENCH_NO_SAMPLER"):
            return
        self._stop.clear()
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        if self.thread is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            self.thread = None

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples: " + str(self.err or "sampler disabled")]}
        sm = [r[0] for r in self.rows]
        reasons = sorted({n for _, bits in self.rows for n, m in self.REASONS if bits & m})
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm),
                "how": f"pynvml thread, {int(self.period * 1e3)} ms period, inside both timed regions (value and e2e)"}


def make_batches(n_batches, rank, cfg_o, O):
    """frame i of rank r uses seed 1000*r + i."""
    out = []
    for k in range(n_batches):
        seeds = [1000 * rank + k * B_PER_GPU + i for i in range(B_PER_GPU)]
        out.append(O.synth_batch(seeds, cfg_o))
    return out


class CpuReference:
    """The reference's own Python for the path (DynVFE + SPTBackboneMAE forward, loss, backward, clip_grad_norm_, OptimWrapper
    adam_onecycle step - train_utils.py:34-53) on the host CPUs: the UNMODIFIED files under baseline/_ref/ (byte-identical
    copies placed by tools/install_reference.py; /root/reference itself does not exist on the GPU box) imported through
    tests/golden/ref_harness.py, whose stand-ins cover only the third-party packages missing from the image.  ``kind`` is
    "reference" then; if the copies are absent (a checkout that never ran build() next to the reference) the oracle port
    runs instead and ``kind`` is "port"."""

    def __init__(self, total_steps):
        from oracle import gdmae_oracle as O
        self.O, self.cfg = O, O.make_cfg("waymo_ssl")
        self.cores = len(os.sched_getaffinity(0))
        torch.set_num_threads(self.cores)
        self.g = torch.Generator().manual_seed(666)
        ref_root = os.path.join(ROOT, "baseline", "_ref")
        if os.path.exists(os.path.join(ref_root, "MANIFEST.json")) or os.path.isdir("/root/reference"):
            sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
            import ref_harness as RH
            mcfg = RH.load_model_cfg("tools/cfgs/waymo_models/gd_mae_ssl.yaml")
            torch.manual_seed(666)
            self.model = RH.RefMAE(mcfg.MODEL, self.cfg["n_feat"], self.cfg["voxel"], np.array(self.cfg["pc_range"], dtype=np.float32),
                                   np.array(self.cfg["grid"], dtype=np.int64)).train()
            sys.path.insert(0, RH.REF + "/tools")
            import train_utils.optimization as RO
            self.opt = RO.build_optimizer(self.model, mcfg.OPTIMIZATION)
            self.sched, _ = RO.build_scheduler(self.opt, total_iters_each_epoch=max(total_steps, 10), total_epochs=1, last_epoch=-1,
                                               optim_cfg=mcfg.OPTIMIZATION)
            self.clip = mcfg.OPTIMIZATION.GRAD_NORM_CLIP
            self.kind = "reference"
        else:
            self.P, self.Bf = O.init_params(self.cfg, 0)
            self.opt = O.AdamOneCycle(self.P, self.cfg, max(total_steps, 10))
            self.kind = "port"

    def step(self, pts, it):
        if self.kind == "reference":
            from torch.nn.utils import clip_grad_norm_
            self.sched.step(it)
            self.opt.zero_grad()
            loss, _ = self.model(dict(points=pts, batch_size=1))
            loss.backward()
            clip_grad_norm_(self.model.parameters(), self.clip)
            self.opt.step()
            return float(loss.detach())
        O = self.O
        _, _, _, vc, _ = O.voxelize(pts, self.cfg)
        return O.train_step(self.P, self.Bf, self.opt, pts, 1, self.cfg, torch.rand(vc.shape[0], generator=self.g), it)[0]

    def describe(self):
        what = ("the reference's own Python (baseline/_ref, unmodified) through its DynVFE / SPTBackboneMAE / OptimWrapper"
                if self.kind == "reference" else "torch CPU oracle port")
        return f"B=1 frame (~159k points) per step of the same synthetic Waymo-shape generator; fwd+loss+bwd+clip+adam_onecycle, fp32, {what}"


def run_reference(args):
    """The reference's CPU implementation of the path on the host cores, bounded sample: B=1 frame per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import warnings
    warnings.filterwarnings("ignore")
    total = max(args.steps + args.warmup, 2)
    R = CpuReference(total)
    frames = [torch.from_numpy(R.O.synth_batch([s], R.cfg)) for s in range(min(total, 4))]
    for it in range(args.warmup):
        R.step(frames[it % len(frames)], it)
    t0 = time.perf_counter()
    for it in range(args.steps):
        loss = R.step(frames[(args.warmup + it) % len(frames)], args.warmup + it)
    dt = time.perf_counter() - t0
    v = args.steps / dt
    sample = R.describe()
    print(json.dumps({
        "impl": "reference", "metric": "mae_pretrain_frames_per_sec", "value": v, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": R.cores, "kind": R.kind, "sample": sample},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "final_loss": float(loss)}))


def cpu_baseline_leg(budget_s=25.0):
    """Rank 0, N=1: the reference's CPU implementation on the host cores for a bounded sample of the same workload."""
    import warnings
    warnings.filterwarnings("ignore")
    R = CpuReference(10)
    pts = torch.from_numpy(R.O.synth_batch([0], R.cfg))
    R.step(pts, 0)  # warm-up
    n, t0 = 0, time.perf_counter()
    while n < 3 or (time.perf_counter() - t0 < budget_s and n < 8):
        R.step(pts, n + 1)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "frames/s", "cores": R.cores, "kind": R.kind, "sample": f"{n} iterations of " + R.describe()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--no-augment", action="store_true", help="e2e leg without the device-side world augmentation")
    ap.add_argument("--loss-read", default="lagged", choices=["lagged", "sync"],
                    help="e2e leg: loss of step i-1 read during step i from a pinned slot (default), or a blocking loss.item() per step")
    ap.add_argument("--impl", default="gdmae_b200")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "tf32", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--matmul", default="high", choices=["high", "medium"], help="torch float32 matmul precision in tf32/bf16 mode")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    import gd_mae_b200  # noqa: F401
    from gd_mae_b200 import _lib, config
    from gd_mae_b200.trainer import MAETrainer
    from oracle import gdmae_oracle as O  # synthetic scene generator + cpu_baseline leg only

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL's log is the driver's evidence of the communicator size: leave NCCL_DEBUG / NCCL_DEBUG_FILE as the caller
        # set them.  If the caller asked for INFO without a file, send it to stderr so stdout stays the one JSON line.
        if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
            os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"
    _lib.check(_lib.lib().gdmae_check_device(), "gdmae_check_device")

    torch.backends.cudnn.benchmark = True
    import contextlib
    autocast = contextlib.nullcontext()  # no autocast: encoder GEMMs run TF32, the decoder map/conv are emitted in bf16

    torch.manual_seed(666 + rank)
    cfg = config.builtin_cfg("waymo_ssl")
    model = config.build_mae_model(cfg).to(dev)
    config.set_precision(model, args.dtype, args.matmul, dense_spatial_features=False)  # MAE head reads the map at pillar cells only
    if world > 1:  # identical initial weights on every rank (DDP broadcasts rank 0's, train.py:146)
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, 0)
    total_steps = 2 * (args.steps + args.warmup) + 8 + SETTLE_STEPS + 12
    trainer = MAETrainer(model, cfg.OPTIMIZATION, total_steps=total_steps, world_size=world)

    cfg_o = O.make_cfg("waymo_ssl")
    n_pool = 4
    host_batches = [torch.from_numpy(b).pin_memory() for b in make_batches(n_pool, rank, cfg_o, O)]
    dev_batches = [b.to(dev) for b in host_batches]
    pts_per_batch = float(np.mean([b.shape[0] for b in host_batches]))

    # Both legs hand the trainer the NEXT batch as well: its index structures (voxelisation, mask, site sets, window
    # tables) are built on a side stream while this step's backward is queued, so a step starts without a host sync.
    res_bd = {}

    def step_resident(i):
        bd = res_bd.pop(i, None) or {"points": dev_batches[i % n_pool], "batch_size": B_PER_GPU}
        nxt = res_bd[i + 1] = {"points": dev_batches[(i + 1) % n_pool], "batch_size": B_PER_GPU}
        with autocast:
            return trainer.step(bd, next_batch=nxt)

    # e2e: every step copies its input batch from pinned host memory and reads its loss back.  The copy of step i+1 is
    # issued on a side stream while step i computes (the double buffering a DataLoader with pin_memory gives), so the
    # 30 MB H2D transfer overlaps compute instead of preceding it; all copies lie inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    pending = {}
    # the input side of the SSL config (tools/cfgs/waymo_models/gd_mae_ssl.yaml:18-31: random_world_flip / rotation / scaling,
    # then mask_points_and_boxes_outside_range = the voxeliser's keep mask): the collated batch is augmented ON THE DEVICE,
    # one launch, right behind its H2D copy on the copy stream
    from gd_mae_b200.pcdet.datasets.augmentor.data_augmentor import DataAugmentor
    augmentor = DataAugmentor(None, config.to_attr({"DISABLE_AUG_LIST": ["placeholder"], "AUG_CONFIG_LIST": [
        {"NAME": "random_world_flip", "PROBABILITY": 0.5, "ALONG_AXIS_LIST": ["x", "y"]},
        {"NAME": "random_world_rotation", "PROBABILITY": 1.0, "WORLD_ROT_ANGLE": [-0.78539816, 0.78539816]},
        {"NAME": "random_world_scaling", "PROBABILITY": 1.0, "WORLD_SCALE_RANGE": [0.95, 1.05]}]}), ["Vehicle", "Pedestrian", "Cyclist"])
    np.random.seed(1000 + rank)

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            t = host_batches[i % n_pool].to(dev, non_blocking=True)
            bd = {"points": t, "batch_size": B_PER_GPU}
            if not args.no_augment:
                bd = augmentor.forward(bd)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        pending[i] = ({"points": bd["points"], "batch_size": B_PER_GPU}, ev)

    def step_e2e(i):
        if i not in pending:
            prefetch(i)
        bd, ev = pending.pop(i)
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        bd["points"].record_stream(cur)
        prefetch(i + 1)
        nxt, nev = pending[i + 1]
        with autocast:
            loss = trainer.step(bd, next_batch=nxt, next_ready_event=nev)
        if args.loss_read == "sync":
            return loss.item()          # blocking D2H read of the step's result: the launch queue drains every step
        # D2H read of every step's loss through a pinned slot: the copy of step i is enqueued behind step i, the host reads
        # the value of step i-1 (waiting for that copy only) - the progress-bar read of train_utils.py:68-79 one step late.
        # Every loss of the leg is on the host inside the timed region: timed() ends with trainer.drain_loss().
        return trainer.loss_to_host(loss)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(dev)
    device_allocs = []      # cudaMalloc calls of the caching allocator inside each timed leg (0 in a steady-state loop)

    def timed(fn, steps, warmup, sample=True):
        """W untimed steps, then exactly `steps` steps between barrier+synchronize on both sides, CUDA events on the
        launching stream, MAX over ranks.  The value and the e2e leg run this same loop; nothing else is inside it."""
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.lib().gdmae_launch_count()
        mallocs0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
        if sample:
            sampler.start()
        e0.record()
        for i in range(steps):
            last = fn(warmup + i)
        if fn is step_e2e and args.loss_read == "lagged":
            last = trainer.drain_loss()      # the last step's loss reaches the host before the closing event
        e1.record()
        barrier()
        sampler.stop()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        device_allocs.append(torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - mallocs0)
        return float(ms) / steps, last, _lib.lib().gdmae_launch_count() - l0

    # ---- settle: a fresh box still pages the image in and loads CUDA modules lazily; r2 traces (tools/step_times.py) showed a
    # rare 50-80 ms host stall within the first ~30 steps of a process.  A few untimed steps through every batch of the
    # pool come before the W warm-up steps of each leg (they are not counted anywhere).
    for i in range(SETTLE_STEPS):
        step_resident(i)
    res_bd.clear()
    # after the settle steps: cuDNN's first-call algorithm search (cudnn.benchmark) ends with an emptyCache(), which would
    # hand an earlier reservation back to the driver
    trainer.reserve_memory(main_gb=24.0, side_gb=2.0, extra_streams=[copy_stream])
    import gc
    gc.collect()
    gc.freeze()     # the model / trainer object graph is permanent: keep it out of the cyclic collector's full passes
    # ---- leg 1: device-resident timing (`value`)
    ms_step, last_loss, launches = timed(step_resident, args.steps, args.warmup)
    # ---- leg 2: end-to-end timing through the public API with host inputs (`e2e`)
    ms_e2e, last_e2e, _ = timed(step_e2e, args.steps, max(3, args.warmup))
    clocks = sampler.summary()
    # ---- leg 3 (NOT part of value / e2e): the same resident step with CUDA events around the hand-written kernels the
    # roofline section reports (Python-side events in _lib.timed, C-side spans inside the executors)
    n_prof = min(args.steps, 10)
    _lib.KERNEL_TIMERS = {}
    _lib.lib().gdmae_timing_enable(1)
    ms_prof, _, _ = timed(step_resident, n_prof, 1, sample=False)
    _lib.lib().gdmae_timing_enable(0)
    timers, _lib.KERNEL_TIMERS = _lib.KERNEL_TIMERS or {}, None
    import ctypes
    cap = 512 * max(n_prof + 1, 1)
    meta = (ctypes.c_int64 * (4 * cap))()
    span_ms = (ctypes.c_float * cap)()
    n_span = _lib.lib().gdmae_timing_drain(meta, span_ms, cap)
    c_spans = {}
    for i in range(n_span):
        name = SPAN_NAMES.get(int(meta[4 * i]), "kind%d_{d}" % int(meta[4 * i])).format(d=int(meta[4 * i + 1]))
        if int(meta[4 * i]) == 3:      # own tcgen05 GEMM: one entry per kernel instantiation (tile width BN, epilogue mode)
            dd = int(meta[4 * i + 1])
            name = "tc_gemm_bn%d_%s" % (dd // 10, TC_GEMM_MODES.get(dd % 10, str(dd % 10)))
        c_spans.setdefault(name, []).append((float(span_ms[i]), int(meta[4 * i + 3])))

    frames = B_PER_GPU * world
    value = frames / (ms_step * 1e-3)
    e2e_value = frames / (ms_e2e * 1e-3)

    pk, pk_src = peaks()
    kernels = {}

    def add_kernel(name, ms, nbytes):
        gbs = [nb / (t * 1e-3) / 1e9 for nb, t in zip(nbytes, ms) if t > 0]
        if name.startswith("tc_gemm"):      # launches of very different sizes share a kernel: bytes of all / time of all
            gbs = [float(np.sum(nbytes)) / (float(np.sum(ms)) * 1e-3) / 1e9]
        kernels[name] = {"launches_per_step": len(ms) / (n_prof + 1), "avg_us": 1e3 * float(np.mean(ms)),
                         "achieved_gbs": float(np.mean(gbs)), "frac": float(np.mean(gbs)) / pk["hbm_gbs"],
                         "share_of_step": float(np.sum(ms)) / (ms_prof * (n_prof + 1)),
                         "algorithmic_bytes": float(np.mean(nbytes))}

    for name, evs in timers.items():
        add_kernel(name, [a.elapsed_time(b) for a, b, _ in evs], [nb for _, _, nb in evs])
        if name.startswith("conv3x3_wgrad"):
            # tensor-bound kernel: `algorithmic_bytes` carries FLOPs (2 * pixels * 9 * 384 * 128), the fraction is of the measured
            # cuBLAS bf16 rate sustained under the power cap (the kernel runs inside a long step)
            k = kernels[name]
            tf = k.pop("achieved_gbs") / 1e3
            peak_tf = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
            k.update({"bound": "tensor", "achieved_tflops": tf, "peak_tflops": peak_tf, "frac": tf / peak_tf,
                      "algorithmic_flops": k.pop("algorithmic_bytes")})
    for name, sp in c_spans.items():
        add_kernel(name, [t for t, _ in sp], [nb for _, nb in sp])
    traffic_tab = {}
    for tname in ("r2_ncu_traffic.json", "r1_ncu_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath):
            with open(tpath) as f:
                for k, v in json.load(f).items():
                    traffic_tab.setdefault(k, v)

    def roofline_of(name):
        k = kernels[name]
        if k.get("bound") == "tensor":
            return {"kernel": name, "bound": "tensor", "achieved": k["achieved_tflops"], "peak": k["peak_tflops"], "unit": "TFLOP/s",
                    "frac": k["frac"], "traffic": traffic_tab.get(name, {}).get("dram_bytes_per_launch"),
                    "traffic_unit": "DRAM bytes per launch (ncu --set full, cold cache)", "algorithmic_flops": k["algorithmic_flops"],
                    "peak_source": pk_src + ", bf16_tflops_sustained", "avg_us": k["avg_us"], "share_of_step": k["share_of_step"]}
        return {"kernel": name, "bound": "hbm", "achieved": k["achieved_gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": k["frac"], "traffic": traffic_tab.get(name, {}).get("dram_bytes_per_launch"),
                "traffic_unit": "DRAM bytes per launch (ncu --set full, cold cache)",
                "algorithmic_bytes": k["algorithmic_bytes"], "peak_source": pk_src, "avg_us": k["avg_us"],
                "share_of_step": k["share_of_step"], "bytes_def": BYTES_DEF.get("tc_gemm" if name.startswith("tc_gemm") else name.split("_d")[0].split("_c")[0], "SURVEY.md 8d")}

    # `roofline` = the hand-written kernel with the largest share of the step; the two kernels north_star names
    # (SRA attention, pillar scatter-max) are reported next to it whichever is dominant.
    roofline = None
    rooflines = {}
    if kernels:
        dom = max(kernels, key=lambda n: kernels[n]["share_of_step"])
        roofline = roofline_of(dom)
        for n in kernels:
            if n.startswith(("sra_", "pillar_scatter_max", "conv3x3_wgrad", "tc_gemm")):
                rooflines[n] = roofline_of(n)

    out = {
        "metric": "mae_pretrain_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_gpu": B_PER_GPU, "global_batch": frames, "points_per_batch": pts_per_batch,
                   "grid": "468x468x1", "parallelism": f"dp{world}", "params": trainer.n_params,
                   "settle_steps_before_warmup": SETTLE_STEPS,
                   "l2": "per-step working set (>2 GB of activations) exceeds the 126 MB L2; input batch changes every step"},
        "e2e": {"value": e2e_value, "unit": "frames/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(pts_per_batch * 6 * 4), "d2h_bytes_per_step": 4 + 2 * 4 * (4 + B_PER_GPU + 1),
                "input_pipeline": "pinned host batch of step i+1 copied on a side stream during step i, world flip / rotation / "
                                  "scaling of the SSL config applied to it on the device (one launch), its index structures "
                                  "prefetched (MAETrainer.step(batch, next_batch)); " + (
                                      "loss.item() every step" if args.loss_read == "sync" else
                                      "every step's loss copied to a pinned host slot behind the step and read by the host one "
                                      "step later (MAETrainer.loss_to_host; the last one before the closing event)"),
                "loss_read": args.loss_read},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "rooflines": rooflines, "kernels": kernels,
        "instrumented_pass": {"steps": n_prof + 1, "ms_per_step": ms_prof, "note": "separate pass after both timed legs; CUDA events around the kernels listed in `kernels`"},
        "final_loss": float(last_loss), "final_loss_e2e": float(last_e2e),
        "cuda_mallocs_in_timed_legs": {"value": device_allocs[0], "e2e": device_allocs[1]},
        "peak_allocated_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30, "reserved_gb": torch.cuda.memory_reserved(dev) / 2 ** 30, "host_cores": len(os.sched_getaffinity(0)),
    }
    if world > 1:
        # every rank applied the same all-reduced gradients to the same start: the parameter buckets must be bit-identical
        chk = torch.stack([trainer.flat_params.double().sum(), trainer.flat_params.double().abs().sum()])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        out["ddp_params_in_sync"] = bool((lo == hi).all())
        assert out["ddp_params_in_sync"], "parameter buckets diverged across ranks"
    from gd_mae_b200 import ops as _ops
    n_to = _ops.sra_wait_timeouts()
    assert n_to == 0, f"{n_to} bounded waits inside the SRA kernels timed out: results are invalid"
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_leg()
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
