"""Fused AdamW over flat parameter / gradient / moment buffers.

Replaces torch.optim.AdamW as configured by VQVAE.configure_optimizers (reference vqvae/model.py:411-438): same
update rule and param-group interface (`param_groups[i]['lr']` is what on_train_batch_start rewrites every step,
model.py:216-218), but all tensors of a group live in ONE contiguous fp32 range so that the update is one kernel
launch per group (vqb_adamw) and the data-parallel gradient all-reduce is one NCCL call on one buffer.

Differences from torch.optim.AdamW, both invisible on the reference's path: (1) a parameter of a group that receives no
gradient in a step is still decayed (its gradient view is zero), where torch skips `grad is None` tensors -- every tensor
the reference hands to its optimizers gets a gradient in every step (frozen tensors are excluded at construction);
(2) parameters become views of the flat buffer: moving or casting the module AFTER the optimizer was built (model.to(...))
breaks the views, exactly as it invalidates a torch optimizer's state.
state_dict() / load_state_dict() speak torch.optim.AdamW's per-parameter format (`state[i] = {step, exp_avg, exp_avg_sq}`),
so the optimizer_states of a reference Lightning checkpoint resume the moments and the bias-correction step.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch

from . import ops


ALIGN = 64          # every tensor starts on a 256-byte boundary of the flat buffers: the kernels read parameters (e.g. the codebook,
                    # a view of flat_param) with 16-byte vector loads and TMA, and a 3-element bias must not shift its successors


def _padded(n: int) -> int:
    return (n + ALIGN - 1) // ALIGN * ALIGN


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, layout_rank=None):
        """`layout_rank` (optional dict id(param) -> int): order of the tensors INSIDE each group's range of the flat buffers
        (ascending).  The data-parallel trainer passes the expected order of gradient completion in the backward pass, so that
        contiguous buckets of the flat gradient become ready -- and are all-reduced -- while the rest of backward still runs.
        The membership and order of `param_groups` (what state_dict() refers to) is unaffected."""
        defaults = dict(lr=float(lr), betas=tuple(float(b) for b in betas), eps=float(eps), weight_decay=float(weight_decay))
        super().__init__(params, defaults)
        self._ranges = []          # per group: (start, end) in the flat buffers
        live: List[List[torch.nn.Parameter]] = []
        total = 0
        device = None
        for g in self.param_groups:
            ps = [p for p in g['params'] if p.requires_grad]
            if layout_rank is not None:
                ps.sort(key=lambda p: layout_rank.get(id(p), 1 << 30))
            for p in ps:
                if p.dtype != torch.float32:
                    raise TypeError('FusedAdamW keeps fp32 master parameters')
                device = device or p.device
            live.append(ps)
            n = sum(_padded(p.numel()) for p in ps)
            self._ranges.append((total, total + n))
            total += n
        if device is None or device.type != 'cuda':
            raise RuntimeError('FusedAdamW needs CUDA parameters (no CPU fallback)')
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=device)      # padding elements stay exactly zero
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=device)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=device)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=device)
        off = 0
        with torch.no_grad():
            for ps in live:
                for p in ps:
                    n = p.numel()
                    view = self.flat_param[off:off + n].view(p.shape)
                    view.copy_(p.data)
                    p.data = view                                   # parameters become views of the flat buffer
                    p.grad = self.flat_grad[off:off + n].view(p.shape)   # autograd accumulates in place into these views
                    off += _padded(n)
        self._live = live
        self.step_count = 0
        # per-step scalars of every group {lr, 1 - beta1^t, sqrt(1 - beta2^t)}: uploaded to device memory by host_step_update() on
        # the current stream (before the step's kernels, or before the replay of a captured graph of the step)
        self._hyper = ops.StepScalars((len(self.param_groups), 4), device)
        self.grad_scale = 1.0       # set to 1/world_size by the data-parallel trainer (after a SUM all-reduce)
        ops.bump_weights_epoch()

    def buckets(self, max_bytes: int = 25 << 20):
        """Contiguous ranges [(start, end, [params])] of the flat buffers of at most ~max_bytes, cut at tensor boundaries (a single
        larger tensor is its own bucket) -- the unit of the overlapped gradient all-reduce (reference: DDP's 25 MB buckets)."""
        out, cur, start, off = [], [], 0, 0
        for p in self._flat_params():
            n = p.numel()
            if cur and (off + n - start) * 4 > max_bytes:
                out.append((start, off, cur)); cur, start = [], off
            cur.append(p); off += _padded(n)
        if cur:
            out.append((start, off, cur))
        return out

    # ---- checkpointing in torch.optim.AdamW's format ---------------------------------------------------------------------
    def _flat_params(self):
        return [p for ps in self._live for p in ps]

    def state_dict(self):
        index, packed_groups, live_ids = {}, [], {id(p) for p in self._flat_params()}
        for g in self.param_groups:
            pg = {k: v for k, v in g.items() if k != 'params'}
            pg['params'] = []
            for p in g['params']:
                index.setdefault(id(p), len(index))
                pg['params'].append(index[id(p)])
            packed_groups.append(pg)
        state, off = {}, 0
        for p in self._flat_params():
            n = p.numel()
            if self.step_count > 0:
                state[index[id(p)]] = {'step': torch.tensor(float(self.step_count)),
                                       'exp_avg': self.exp_avg[off:off + n].view(p.shape).clone(),
                                       'exp_avg_sq': self.exp_avg_sq[off:off + n].view(p.shape).clone()}
            off += _padded(n)
        return {'state': state, 'param_groups': packed_groups}

    @torch.no_grad()
    def load_state_dict(self, state_dict) -> None:
        groups = state_dict['param_groups']
        if len(groups) != len(self.param_groups) or any(len(a['params']) != len(b['params']) for a, b in zip(groups, self.param_groups)):
            raise ValueError('loaded state dict has different parameter groups')
        index = {}
        for saved, g in zip(groups, self.param_groups):
            for k, v in saved.items():
                if k != 'params':
                    g[k] = v
            for i, p in zip(saved['params'], g['params']):
                index[id(p)] = i
        steps, off = set(), 0
        for p in self._flat_params():
            n = p.numel()
            st = state_dict['state'].get(index[id(p)])
            if st is None:
                self.exp_avg[off:off + n].zero_(); self.exp_avg_sq[off:off + n].zero_()
            else:
                self.exp_avg[off:off + n].copy_(st['exp_avg'].reshape(-1).to(self.exp_avg.device, torch.float32))
                self.exp_avg_sq[off:off + n].copy_(st['exp_avg_sq'].reshape(-1).to(self.exp_avg.device, torch.float32))
                steps.add(int(float(st['step'])))
            off += _padded(n)
        if len(steps) > 1:
            raise ValueError(f'FusedAdamW keeps ONE step counter; the loaded state has {sorted(steps)}')
        self.step_count = steps.pop() if steps else 0

    def zero_grad(self, set_to_none: bool = False) -> None:      # grads must stay views of flat_grad
        self.flat_grad.zero_()
        off = 0
        for ps in self._live:
            for p in ps:
                n = p.numel()
                if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * off:
                    p.grad = self.flat_grad[off:off + n].view(p.shape)
                off += _padded(n)

    def host_step_update(self) -> None:
        """Advance the step counter and upload this step's scalars from param_groups.  step() calls it; a trainer replaying a
        captured CUDA graph of the step calls it INSTEAD of step(), before the replay (under capture the upload is skipped: the
        graph holds only the kernels, which read the device copy)."""
        self.step_count += 1
        vals = torch.zeros(len(self.param_groups), 4, dtype=torch.float32)
        for i, g in enumerate(self.param_groups):
            b1, b2 = g['betas']
            vals[i, 0] = float(g['lr'])
            vals[i, 1] = 1.0 - b1 ** self.step_count
            vals[i, 2] = (1.0 - b2 ** self.step_count) ** 0.5
        self._hyper.upload(vals)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self.host_step_update()
        for i, (g, (a, b)) in enumerate(zip(self.param_groups, self._ranges)):
            if b > a:
                ops.adamw_flat_dev(self.flat_param[a:b], self.flat_grad[a:b], self.exp_avg[a:b], self.exp_avg_sq[a:b],
                                   self._hyper.dev[i], g['betas'][0], g['betas'][1], g['eps'], g['weight_decay'], self.grad_scale)
        ops.bump_weights_epoch()
        return loss
