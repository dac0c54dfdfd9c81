"""The training step around GDMAE.forward: one flat fp32 parameter bucket, one flat gradient
bucket (single NCCL all-reduce over NVLink when world_size > 1), gradient-norm clip and the
adam_onecycle update as ONE fused kernel.

Replaces (reference file:line, relative to /root/reference):
  train_one_epoch body                  tools/train_utils/train_utils.py:34-53
  build_optimizer('adam_onecycle')      tools/train_utils/optimization/__init__.py:19-32
  OptimWrapper.step + torch Adam        tools/train_utils/optimization/fastai_optim.py:135-152
  OneCycle                              tools/train_utils/optimization/learning_schedules_fastai.py:44-77
  DDP gradient all-reduce               tools/train.py:146

Semantics kept: clip norm over ALL parameters' gradients; decoupled weight decay
p *= 1 - wd*lr on every optimised parameter (BN included); Adam betas (mom(t), 0.99), eps 1e-8
with bias corrections using the current momentum; parameters held directly by a module that has
children (CosineMultiheadAttention.in_proj_weight / in_proj_bias / tau) receive gradients, count
in the clip norm and are all-reduced but are NEVER updated (flatten_model() only collects leaf
modules, optimization/__init__.py:26-27)."""
import ctypes
import math
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L


def annealing_cos(start, end, pct):
    return end + (start - end) / 2 * (np.cos(np.pi * pct) + 1)


def onecycle(step, total_steps, lr_max, moms, div_factor, pct_start):
    """-> (lr, mom) at iteration ``step`` (learning_schedules_fastai.py:44-77)."""
    low = lr_max / div_factor
    a1 = int(total_steps * pct_start)
    lr, mom = low, moms[0]
    for s, e, a, b in ((0, a1, low, lr_max), (a1, total_steps, lr_max, low / 1e4)):
        if step >= s:
            lr = annealing_cos(a, b, (step - s) / (e - s))
    for s, e, a, b in ((0, a1, moms[0], moms[1]), (a1, total_steps, moms[1], moms[0])):
        if step >= s:
            mom = annealing_cos(a, b, (step - s) / (e - s))
    return float(lr), float(mom)


def optimised_parameter_names(model):
    """Names of the parameters the reference's flatten_model()/get_layer_groups() hands to Adam:
    those of leaf modules only (optimization/__init__.py:19-32)."""
    names = set()
    for mname, m in model.named_modules():
        if len(list(m.children())) == 0:
            for pname, p in m.named_parameters(recurse=False):
                if p.requires_grad:
                    names.add(f"{mname}.{pname}" if mname else pname)
    return names


def reference_param_groups(model):
    """Parameter NAMES in the order the reference's optimizer indexes them: build_optimizer flattens the model into its leaf
    modules, split_bn_bias puts the non-BatchNorm ones into group 0 and the BatchNorm ones into group 1
    (optimization/__init__.py:19-32, fastai_optim.py:17-29, 115-121); torch numbers the parameters group after group."""
    import torch.nn as nn
    bn_types = (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d, nn.SyncBatchNorm)
    name_of = {id(p): n for n, p in model.named_parameters()}
    leaves = [m for m in model.modules() if len(list(m.children())) == 0]
    groups = []
    for pick_bn in (False, True):
        groups.append([name_of[id(p)] for m in leaves if isinstance(m, bn_types) == pick_bn for p in m.parameters() if p.requires_grad])
    return groups


class MAETrainer:
    def __init__(self, model, optim_cfg, total_steps, world_size=1):
        self.model, self.cfg, self.total_steps, self.world_size = model, optim_cfg, int(total_steps), world_size
        opt_names = optimised_parameter_names(model)
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        ordered = [(n, p) for n, p in named if n in opt_names] + [(n, p) for n, p in named if n not in opt_names]
        # every tensor starts on a 256-byte boundary of the bucket (the kernels read parameters with 128-bit
        # loads); the zero padding has zero gradient, so it neither moves nor contributes to the clip norm
        ALIGN = 64
        pad = lambda k: (k + ALIGN - 1) // ALIGN * ALIGN  # noqa: E731
        self.n_params = sum(p.numel() for _, p in ordered)
        self.n_opt = sum(pad(p.numel()) for n, p in ordered if n in opt_names)
        self.n_all = sum(pad(p.numel()) for _, p in ordered)
        dev = ordered[0][1].device
        self.flat_params = torch.zeros(self.n_all, dtype=torch.float32, device=dev)
        self.flat_grads = torch.zeros(self.n_all, dtype=torch.float32, device=dev)
        off = 0
        self.slices = {}
        for n, p in ordered:
            k = p.numel()
            self.flat_params[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_params[off:off + k].view_as(p)
            p.grad = self.flat_grads[off:off + k].view_as(p)
            self.slices[n] = (off, k)
            off += pad(k)
        self.exp_avg = torch.zeros(self.n_opt, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(self.n_opt, dtype=torch.float32, device=dev)
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        # bf16 mirror of the bucket for the bf16 GEMM configuration: refreshed once per step, its views are the
        # GEMM-operand copies of the weights (fused.BF16_SHADOW), so no per-layer weight casts are launched
        self.flat_bf16 = None
        self._params = [p for _, p in ordered]
        # Parameters whose gradient arrives through autograd's AccumulateGrad (everything outside the two C executors that
        # write into the bucket themselves: EncoderLayer and the VFE): with .grad = bucket view each of them costs one tiny
        # `grad += g` launch per step (42 launches, 0.22 ms at Waymo, r2 timeline).  Their .grad is cleared before backward,
        # so AccumulateGrad keeps the produced tensor by reference, and ONE multi-tensor copy moves them into the bucket.
        self._deferred = []
        if os.environ.get("GDMAE_DEFER_GRADS", "1") != "0" and dev.type == "cuda":
            owned = set()
            for m in model.modules():
                if type(m).__name__ in ("EncoderLayer", "DynVFE"):
                    owned.update(id(q) for q in m.parameters())
            self._deferred = [(p_, p_.grad) for _, p_ in ordered if id(p_) not in owned]
        self.it = 0       # accumulated_iter of train_one_epoch
        self.t = 0        # Adam step count
        # the VFE's gradients are the last ones backward produces: everything after them in the bucket is complete when the
        # gradient of `pillar_features` arrives, and is all-reduced on a side stream while the VFE backward runs
        vfe_end = 0
        for n, p in ordered:
            if n.startswith('vfe.'):
                vfe_end = max(vfe_end, self.slices[n][0] + pad(p.numel()))
        self.early_split = vfe_end if 0 < vfe_end < self.n_opt else 0
        # float buffers (BatchNorm running statistics) in one flat tensor: DDP (tools/train.py:146, broadcast_buffers=True)
        # hands every rank rank 0's buffers at each forward; here that is ONE small broadcast per step
        fbufs = [(n, b) for n, b in model.named_buffers() if b.dtype == torch.float32]
        self.flat_buffers = torch.zeros(sum(pad(b.numel()) for _, b in fbufs), dtype=torch.float32, device=dev)
        off = 0
        for n, b in fbufs:
            k = b.numel()
            self.flat_buffers[off:off + k].copy_(b.data.reshape(-1))
            b.data = self.flat_buffers[off:off + k].view_as(b)
            off += pad(k)
        # the host may enqueue at most `max_steps_in_flight` iterations ahead of the device: without a bound a loop that never
        # reads the loss queues dozens of steps, and the caching allocator - whose blocks shared with the index side stream
        # are only reusable once their recorded events have passed - falls back to cudaMalloc in the middle of the run
        self.max_steps_in_flight = 2
        self._inflight = []
        self._comm_stream = None
        self._loss_ring = None       # loss_to_host(): two pinned slots + their copy-complete events
        self._early_work = None
        self._late_works = []
        # multi-GPU schedule knobs (measured on 2 B200s, r2)
        self.overlap_allreduce = os.environ.get("GDMAE_DDP_OVERLAP", "1") != "0"
        # BatchNorm running statistics: DDP (broadcast_buffers=True) hands every rank rank 0's buffers at each forward.  They are
        # written, never read, while training, so the observable contract is "every rank evaluates / checkpoints with rank 0's
        # statistics".  "lazy" (default) keeps exactly that with no per-step collective: sync_buffers() broadcasts rank 0's
        # buffers when asked (train_utils.checkpoint_state and MAETrainer.eval_mode call it).  Measured at N=2, 30 steps:
        # per-step broadcast after the all-reduce 22.3 ms/step, overlapped with backward 28.9 ms (its kernel spins on SMs the
        # persistent GEMM / SRA kernels need), none 20.7 ms (= the 1-GPU step).
        self.buffer_sync = os.environ.get("GDMAE_DDP_BCAST", "lazy")      # lazy | after_reduce | overlap
        if self.world_size > 1 and dist.is_available() and dist.is_initialized():
            # DDP construction semantics: every rank starts from rank 0's parameters and buffers
            dist.broadcast(self.flat_params, 0)
            if self.flat_buffers.numel():
                dist.broadcast(self.flat_buffers, 0)

    def zero_grad(self):
        self.flat_grads.zero_()

    def refresh_bf16_mirror(self):
        from . import fused
        if fused.GEMM_DTYPE != torch.bfloat16 or not self.flat_params.is_cuda:
            return
        if self.flat_bf16 is None:
            self.flat_bf16 = torch.empty(self.n_all, dtype=torch.bfloat16, device=self.flat_params.device)
            base = self.flat_params.data_ptr()
            for p in self._params:
                off = (p.data_ptr() - base) // 4
                fused.BF16_SHADOW[p.data_ptr()] = self.flat_bf16[off:off + p.numel()].view(p.shape)
        self.flat_bf16.copy_(self.flat_params)

    def release_bf16_mirror(self):
        """Drops this trainer's entries of fused.BF16_SHADOW (keyed by raw data pointers): a later model whose parameters land
        on recycled addresses must not pick up a dead trainer's mirror.  Called by __del__; call it explicitly
        before building another trainer in the same process."""
        from . import fused
        if self.flat_bf16 is not None:
            for p in self._params:
                sh = fused.BF16_SHADOW.get(p.data_ptr())
                if sh is not None and sh.untyped_storage().data_ptr() == self.flat_bf16.untyped_storage().data_ptr():
                    del fused.BF16_SHADOW[p.data_ptr()]
            self.flat_bf16 = None

    def __del__(self):
        try:
            self.release_bf16_mirror()
        except Exception:
            pass

    def _comm(self):
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=self.flat_grads.device, priority=-1)
        return self._comm_stream

    def _flush_deferred(self):
        """gradients AccumulateGrad left on the deferred parameters -> their bucket views (one multi-tensor copy), .grad = view"""
        src, dst = [], []
        for p_, view in self._deferred:
            g = p_.grad
            if g is not None and g.data_ptr() != view.data_ptr():
                src.append(g.detach())
                dst.append(view)
            p_.grad = view
        if src:
            torch._foreach_copy_(dst, src)

    def _reduce_early(self, grad):
        """Tensor hook on `pillar_features`: the backbone's backward is enqueued, its gradients (bucket[early_split:]) are
        final.  Their all-reduce starts on the side stream now and overlaps the VFE backward (DDP's bucket overlap,
        tools/train.py:146)."""
        if self._early_work is None and self.early_split:
            self._flush_deferred()      # the backbone's deferred gradients are part of what is reduced now
            ev = torch.cuda.Event()
            ev.record()
            side = self._comm()
            with torch.cuda.stream(side):
                side.wait_event(ev)
                self._early_work = dist.all_reduce(self.flat_grads[self.early_split:], op=dist.ReduceOp.SUM, async_op=True)
        return grad

    def sync_buffers(self):
        """rank 0's BatchNorm running statistics to every rank (DDP broadcast_buffers).  Call before evaluating or saving on
        a rank other than 0; a no-op on one GPU."""
        if self.world_size > 1 and self.flat_buffers.numel() and dist.is_initialized():
            dist.broadcast(self.flat_buffers, 0)

    def eval_mode(self):
        """train -> eval transition: every rank continues with rank 0's running statistics"""
        self.sync_buffers()
        self.model.eval()

    def _sync_buffers_overlapped(self):
        if self.world_size > 1 and self.flat_buffers.numel() and self.flat_buffers.is_cuda and self.buffer_sync == "overlap":
            ev = torch.cuda.Event()
            ev.record()
            side = self._comm()
            with torch.cuda.stream(side):
                side.wait_event(ev)
                self._late_works.append(dist.broadcast(self.flat_buffers, 0, async_op=True))

    def reduce_gradients(self):
        """The only exchange step of the path: all-reduce(SUM) of the flat gradient bucket (32.4 MB at Waymo), as one
        message, or as two when the early part was already started from the backward hook; the 1/world_size averaging is
        folded into the clip/Adam kernel (grad_scale)."""
        if self.world_size <= 1:
            return
        if self._early_work is not None:
            w = dist.all_reduce(self.flat_grads[:self.early_split], op=dist.ReduceOp.SUM, async_op=True)
            self._early_work.wait()      # stream-level: the current stream waits for NCCL's, the host does not block
            w.wait()
            self._early_work = None
        else:
            dist.all_reduce(self.flat_grads, op=dist.ReduceOp.SUM)
        if self.buffer_sync == "after_reduce" and self.flat_buffers.numel():
            # the ranks have just met in the all-reduce: the small broadcast of rank 0's running statistics (DDP
            # broadcast_buffers) follows on the same stream without anybody spinning for a late peer
            dist.broadcast(self.flat_buffers, 0)

    def state_dict(self, reference_format=False):
        """Optimizer state of the path (train_utils.checkpoint_state: optimizer_state + it): the Adam moments and both
        counters, keyed by parameter name - or, with ``reference_format``, laid out as the ``optimizer_state`` the reference
        saves (the state_dict of the torch Adam inside its OptimWrapper: per-index state + the two param groups), so that a
        checkpoint written here resumes in the reference and the other way round."""
        if reference_format:
            groups = reference_param_groups(self.model)
            template = torch.optim.Adam([{'params': [torch.nn.Parameter(torch.zeros(1))], 'lr': 0}], betas=(0.9, 0.99)).state_dict()
            lr, mom = onecycle(max(self.it - 1, 0), self.total_steps, self.cfg.LR, list(self.cfg.MOMS), self.cfg.DIV_FACTOR,
                               self.cfg.PCT_START)
            state, pgs, idx = {}, [], 0
            for names in groups:
                pg = dict(template['param_groups'][0])
                pg.update(lr=lr, betas=(mom, 0.99), weight_decay=0, params=list(range(idx, idx + len(names))))
                pgs.append(pg)
                for n in names:
                    off, k = self.slices[n]
                    if self.t > 0:
                        shape = dict(self.model.named_parameters())[n].shape
                        state[idx] = {'step': torch.tensor(float(self.t)),
                                      'exp_avg': self.exp_avg[off:off + k].detach().cpu().clone().view(shape),
                                      'exp_avg_sq': self.exp_avg_sq[off:off + k].detach().cpu().clone().view(shape)}
                    idx += 1
            return {'state': state, 'param_groups': pgs}
        out = {'it': self.it, 't': self.t, 'total_steps': self.total_steps, 'exp_avg': {}, 'exp_avg_sq': {}}
        for n, (off, k) in self.slices.items():
            if off < self.n_opt:
                out['exp_avg'][n] = self.exp_avg[off:off + k].detach().cpu().clone()
                out['exp_avg_sq'][n] = self.exp_avg_sq[off:off + k].detach().cpu().clone()
        return out

    def load_state_dict(self, sd):
        """own format (state_dict()) or the reference's ``optimizer_state`` (detected by its 'param_groups' key;
        detector3d_template.py:425-436 / train.py resume path)"""
        if 'param_groups' in sd:
            names = [n for g in reference_param_groups(self.model) for n in g]
            assert sum(len(g['params']) for g in sd['param_groups']) == len(names), "optimizer_state does not belong to this model"
            order = [i for g in sd['param_groups'] for i in g['params']]
            steps = set()
            for i, n in zip(order, names):
                st = sd['state'].get(i, None)
                off, k = self.slices[n]
                if st is None:
                    self.exp_avg[off:off + k].zero_()
                    self.exp_avg_sq[off:off + k].zero_()
                    continue
                self.exp_avg[off:off + k].copy_(st['exp_avg'].reshape(-1))
                self.exp_avg_sq[off:off + k].copy_(st['exp_avg_sq'].reshape(-1))
                steps.add(int(st['step']))
            assert len(steps) <= 1, "parameters with different Adam step counts"
            self.t = steps.pop() if steps else 0
            self.it = self.t
            return
        self.it, self.t = int(sd['it']), int(sd['t'])
        self.total_steps = int(sd.get('total_steps', self.total_steps))
        for n, (off, k) in self.slices.items():
            if off < self.n_opt:
                self.exp_avg[off:off + k].copy_(sd['exp_avg'][n].reshape(-1))
                self.exp_avg_sq[off:off + k].copy_(sd['exp_avg_sq'][n].reshape(-1))

    def optimizer_step(self):
        lr, mom = onecycle(self.it, self.total_steps, self.cfg.LR, list(self.cfg.MOMS), self.cfg.DIV_FACTOR, self.cfg.PCT_START)
        self.reduce_gradients()
        if not self.flat_grads.is_cuda:
            raise L.GdmaeError("MAETrainer.optimizer_step needs CUDA tensors (no CPU optimizer path exists)")
        lib = L.lib()
        self.sumsq.zero_()
        L.check(lib.gdmae_grad_sumsq(L.P(self.flat_grads), L.i64(self.n_all), L.P(self.sumsq), L.stream()), "gdmae_grad_sumsq")
        self.t += 1
        beta2 = 0.99
        bc1 = 1 - mom ** self.t
        bc2 = 1 - beta2 ** self.t
        L.check(lib.gdmae_adam_onecycle_step(
            L.P(self.flat_params), L.P(self.flat_grads), L.P(self.exp_avg), L.P(self.exp_avg_sq), L.i64(self.n_opt),
            L.P(self.sumsq), L.f32(self.cfg.GRAD_NORM_CLIP), L.f32(1 - self.cfg.WEIGHT_DECAY * lr), L.f32(mom), L.f32(beta2),
            L.f32(1e-8), L.f32(lr / bc1), L.f32(math.sqrt(bc2)), L.f32(1.0 / self.world_size), L.stream()),
            "gdmae_adam_onecycle_step")
        self.it += 1
        return lr, mom

    def reserve_memory(self, main_gb=24.0, side_gb=2.0, extra_streams=()):
        """Maps device memory for the caching allocator's pools up front: one block of ``main_gb`` on the current stream and
        one of ``side_gb`` on every side stream of the step (the index-prefetch stream, ``extra_streams`` such as the caller's
        H2D copy stream) is allocated and released again, which leaves the address range mapped and cached.  Batches differ in
        size from step to step, so the pools keep growing by a few blocks long after warm-up; growing a pool inside a step
        costs one cuMemCreate + cuMemMap per 20 MB (r2 measurement: 30-120 ms of host stall per growth of the step's
        GB-sized buffers, 1-6 of them inside a 20-step timed leg on a fresh box).  A B200 has 180 GB; the step peaks far below
        the default reservation."""
        if not self.flat_grads.is_cuda:
            return
        dev = self.flat_grads.device
        if hasattr(self.model, 'prefetch_index') and getattr(self.model, '_side_stream', None) is None:
            self.model._side_stream = torch.cuda.Stream(device=dev, priority=-1)
        plan = [(torch.cuda.current_stream(dev), main_gb)]
        plan += [(st, side_gb) for st in [getattr(self.model, '_side_stream', None), *extra_streams] if st is not None]
        free_b, _ = torch.cuda.mem_get_info(dev)
        for st, gb in plan:
            nbytes = int(min(gb * 2 ** 30, 0.5 * free_b))
            with torch.cuda.stream(st):
                block = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                del block
        torch.cuda.synchronize(dev)

    def loss_to_host(self, loss):
        """The per-iteration `loss.item()` of train_one_epoch (tools/train_utils/train_utils.py:68-79: on rank 0 only, the value
        feeds the progress bar and tensorboard) without draining the launch queue: the loss of THIS iteration is copied to a pinned host slot behind the
        step on the current stream, and the call returns the loss of the PREVIOUS iteration (None on the first call), waiting
        only for that older copy.  One device-to-host read per iteration; the host stays one step ahead of the device.
        `drain_loss()` returns the newest value (blocking)."""
        if self._loss_ring is None:
            self._loss_ring = {"buf": torch.empty(2, dtype=torch.float32).pin_memory(), "ev": [None, None], "k": 0}
        r = self._loss_ring
        k = r["k"]
        r["buf"][k:k + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        r["ev"][k] = ev
        r["k"] = k ^ 1
        prev = r["ev"][k ^ 1]
        if prev is None:
            return None
        prev.synchronize()
        return float(r["buf"][k ^ 1])

    def drain_loss(self):
        r = self._loss_ring
        if r is None or r["ev"][r["k"] ^ 1] is None:
            return None
        r["ev"][r["k"] ^ 1].synchronize()
        return float(r["buf"][r["k"] ^ 1])

    def step(self, batch_dict, next_batch=None, next_ready_event=None):
        """One iteration (train_utils.py:34-53): zero_grad, forward, backward, all-reduce, clip, update.
        Returns the loss tensor (no host sync).  ``next_batch``: the batch_dict of the following iteration (its points already
        on the device, or arriving - ``next_ready_event``); its index structures are built on a side stream once this
        iteration is enqueued (GDMAE.prefetch_index), so the next call to step(next_batch) starts without a host sync."""
        self.model.train()
        if next_batch is not None and next_ready_event is None and self.flat_grads.is_cuda:
            # the producer of next_batch['points'] (an H2D copy, the device augmentor) was enqueued by the caller on the current
            # stream before this call: the side stream that builds the index structures must not start before it
            next_ready_event = torch.cuda.Event()
            next_ready_event.record()
        for w in self._late_works:       # last step's buffer broadcast is complete before this forward updates the buffers
            w.wait()
        self._late_works = []
        self.zero_grad()
        self.refresh_bf16_mirror()
        ret_dict, tb_dict, _ = self.model(batch_dict)
        loss = ret_dict['loss'].mean()
        self._sync_buffers_overlapped()  # only with GDMAE_DDP_BCAST=overlap (measured slower, kept for the record)
        hook = None
        if self.world_size > 1 and self.early_split and self.flat_grads.is_cuda and self.overlap_allreduce:
            pf = batch_dict.get('pillar_features', None)
            if isinstance(pf, torch.Tensor) and pf.requires_grad:
                hook = pf.register_hook(self._reduce_early)
        for p_, _ in self._deferred:
            p_.grad = None
        from . import fused as _fused
        prev_inplace, _fused.INPLACE_PARAM_GRADS = _fused.INPLACE_PARAM_GRADS, True      # the executors write into the bucket
        try:
            loss.backward()
        finally:
            _fused.INPLACE_PARAM_GRADS = prev_inplace
        if hook is not None:
            hook.remove()
        self._flush_deferred()
        self.optimizer_step()
        if next_batch is not None and hasattr(self.model, 'prefetch_index'):
            self.model.prefetch_index(next_batch, next_ready_event)
        if hasattr(self.model, 'update_global_step'):
            self.model.update_global_step()
        if self.flat_grads.is_cuda and self.max_steps_in_flight:
            ev = torch.cuda.Event()
            ev.record()
            self._inflight.append(ev)
            if len(self._inflight) > self.max_steps_in_flight:
                self._inflight.pop(0).synchronize()
        return loss.detach()
