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
        self.it = 0       # accumulated_iter of train_one_epoch
        self.t = 0        # Adam step count

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

    def reduce_gradients(self):
        """The only exchange step of the path: ONE all-reduce(SUM) of the flat gradient bucket (32.4 MB at
        Waymo); the 1/world_size averaging is folded into the clip/Adam kernel (grad_scale)."""
        if self.world_size > 1:
            dist.all_reduce(self.flat_grads, op=dist.ReduceOp.SUM)

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

    def step(self, batch_dict, next_batch=None, next_ready_event=None):
        """One iteration (train_utils.py:34-53): zero_grad, forward, backward, all-reduce, clip, update.
        Returns the loss tensor (no host sync).  ``next_batch``: the batch_dict of the following iteration (its points already
        on the device, or arriving - ``next_ready_event``); its index structures are built on a side stream once this
        iteration is enqueued (GDMAE.prefetch_index), so the next call to step(next_batch) starts without a host sync."""
        self.model.train()
        self.zero_grad()
        self.refresh_bf16_mirror()
        ret_dict, tb_dict, _ = self.model(batch_dict)
        loss = ret_dict['loss'].mean()
        loss.backward()
        self.optimizer_step()
        if next_batch is not None and hasattr(self.model, 'prefetch_index'):
            self.model.prefetch_index(next_batch, next_ready_event)
        if hasattr(self.model, 'update_global_step'):
            self.model.update_global_step()
        return loss.detach()
