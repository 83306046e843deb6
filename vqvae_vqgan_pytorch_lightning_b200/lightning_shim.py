"""Minimal stand-in for the part of pytorch_lightning that the reference's VQVAE module touches
(vqvae/model.py: self.log, self.optimizers(), self.manual_backward, self.trainer.{num_training_batches,optimizers},
self.current_epoch, automatic_optimization and the fit-loop hooks), plus a one-process-per-GPU data-parallel
Trainer.  This class is ALWAYS the base of VQVAE, also when pytorch_lightning is importable: the fused optimizers, the EMA
statistics all-reduce and the gradient scaling are wired up by Trainer.attach, which a stock pl.Trainer would not do
(INTEGRATION.md shows how to drive the module from a Lightning-style loop).

Data parallelism (reference: Lightning DDPStrategy, vqvae/train.py:128-131): the global batch is sharded across
ranks; the flat gradient buffer of each FusedAdamW is cut into ~25 MB buckets laid out in the order the backward pass
completes them, and every bucket is SUM-all-reduced with NCCL over NVLink (async, on NCCL's stream) as soon as its last
gradient has been accumulated -- overlapped with the rest of backward, like DDP's bucketed reducer; the 1/world scaling
happens inside the AdamW kernel; EMA cluster statistics are all-reduced by the quantizer in ONE [counts | dw] buffer.
"""
from __future__ import annotations

from typing import Any, Dict, Iterable, List, Optional

import torch
import torch.distributed as dist
from torch import nn

HAVE_LIGHTNING = False            # kept for callers that branched on it: the shim never defers to pytorch_lightning


class LightningModule(nn.Module):
    """The subset of pl.LightningModule used by vqvae/model.py."""

    def __init__(self):
        super().__init__()
        self.trainer: Optional['Trainer'] = None
        self.current_epoch = 0
        self.automatic_optimization = True
        self.logged: Dict[str, Any] = {}

    def log(self, name: str, value, **kwargs) -> None:
        # values may be device tensors; nothing is synchronised here (the reference does 7 .item() syncs per step)
        self.logged[name] = value

    def optimizers(self):
        opts = self.trainer.optimizers
        return opts if len(opts) > 1 else opts[0]

    def manual_backward(self, loss: torch.Tensor, optimizer=None) -> None:
        """backward + data-parallel all-reduce of `optimizer`'s gradients (all optimizers when None)"""
        loss.backward()
        if self.trainer is not None:
            self.trainer.sync_gradients(optimizer)

    # hooks (overridden by the model)
    def on_train_start(self): ...
    def on_train_batch_start(self, batch, batch_index): ...
    def on_train_epoch_end(self): ...
    def on_train_end(self): ...


class DevicePrefetcher:
    """Host batches -> device batches one step AHEAD of their use: the pinned-memory H2D copy of batch i+1 runs on a side stream
    while step i computes (what the reference's DataLoader(pin_memory=True) + Lightning's batch transfer give it,
    vqvae/train.py:121-142).  Iterating yields device tensors that are safe to use on the current stream until the next batch is
    requested.  The copies land in a small ring of persistent device buffers (allocating a fresh tensor per batch on the side
    stream costs cudaMalloc stalls: a side stream has its own allocator pool)."""

    def __init__(self, batches: Iterable, device, depth: int = 1):
        self.it = iter(batches)
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self.depth = max(1, depth)
        self.slots = [{'bufs': None, 'ready': torch.cuda.Event(), 'free': None} for _ in range(self.depth + 2)]
        self.queue = []
        self.w = 0
        self.last = None

    @staticmethod
    def _tensors(batch):
        return [batch] if torch.is_tensor(batch) else [t for t in batch if torch.is_tensor(t)]

    def preallocate(self, example) -> 'DevicePrefetcher':
        """allocate the ring for batches shaped like `example` now (outside a timed / latency-critical region)"""
        for slot in self.slots:
            slot['bufs'] = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in self._tensors(example)]
        return self

    def _issue(self) -> bool:
        try:
            host = next(self.it)
        except StopIteration:
            return False
        slot = self.slots[self.w % len(self.slots)]
        self.w += 1
        src = self._tensors(host)
        with torch.cuda.stream(self.stream):
            if slot['free'] is not None:
                self.stream.wait_event(slot['free'])              # the step that read this buffer has finished with it
            if slot['bufs'] is None or len(slot['bufs']) != len(src) or any(b.shape != t.shape or b.dtype != t.dtype for b, t in zip(slot['bufs'], src)):
                slot['bufs'] = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in src]
            for b, t in zip(slot['bufs'], src):
                b.copy_(t, non_blocking=True)
            slot['ready'].record(self.stream)
        it = iter(slot['bufs'])
        out = next(it) if torch.is_tensor(host) else tuple(next(it) if torch.is_tensor(h) else h for h in host)
        self.queue.append((slot, out))
        return True

    def __iter__(self):
        return self

    def __next__(self):
        cur = torch.cuda.current_stream(self.device)
        if self.last is not None:                                  # the consumer asked for the next batch: the previous one is released
            self.last['free'] = torch.cuda.Event()
            self.last['free'].record(cur)
            self.last = None
        while len(self.queue) < self.depth + 1 and self._issue():
            pass
        if not self.queue:
            raise StopIteration
        slot, out = self.queue.pop(0)
        cur.wait_event(slot['ready'])                              # the copy has landed before the step reads it
        self.last = slot
        return out


class Trainer:
    """One-process-per-GPU fit loop (rank / world from torch.distributed when initialised)."""

    def __init__(self, max_epochs: int = 1, num_training_batches: Optional[int] = None, overlap_grad_sync: bool = True,
                 bucket_bytes: int = 25 << 20, cuda_graph: bool = False, graph_warmup: int = 2):
        """cuda_graph: capture one whole optimisation step (forward, backward, gradient reduction, optimizer) per step VARIANT
        (model.graph_variant) after `graph_warmup` eager steps of that variant and replay it afterwards: ~600 kernel launches
        become one graph launch, which removes the host-side launch gaps between the small kernels of the low-resolution
        levels.  Everything that changes from step to step reaches the kernels through device memory (learning rate and Adam
        bias corrections: FusedAdamW's pinned hyper buffer; Gumbel temperature / KL weight: the quantizer's; the batch: a
        static input tensor), so a replay computes exactly what the eager step would."""
        self.cuda_graph = cuda_graph
        self.graph_warmup = graph_warmup
        self._graphs: Dict[Any, dict] = {}
        self.overlap_grad_sync = overlap_grad_sync
        self.bucket_bytes = bucket_bytes
        self._buckets: Dict[int, list] = {}          # id(optimizer) -> [bucket state]
        self.max_epochs = max_epochs
        self.num_training_batches = num_training_batches or 0
        self.optimizers: List[torch.optim.Optimizer] = []
        self.world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world_size > 1 else 0
        self.model: Optional[LightningModule] = None

    # ---- data parallel plumbing ---------------------------------------------------------------------
    def attach(self, model: LightningModule) -> None:
        self.model = model
        model.trainer = self
        opt = model.configure_optimizers()
        if isinstance(opt, tuple):
            opt = opt[0]
        self.optimizers = list(opt) if isinstance(opt, (list, tuple)) else [opt]
        for o in self.optimizers:
            if hasattr(o, 'grad_scale'):
                o.grad_scale = 1.0 / self.world_size
        q = getattr(model, 'quantizer', None)
        if q is not None and self.world_size > 1:
            q.world_size = self.world_size
            q.stats_allreduce = self._allreduce_stats
        if self.world_size > 1 and self.overlap_grad_sync:
            for o in self.optimizers:
                if hasattr(o, 'buckets'):
                    self._install_bucket_hooks(o)
        # Single process: the backward kernels accumulate parameter gradients straight into the flat gradient buffers
        # (ops.grad_sink) -- no AccumulateGrad `add` launch per parameter.  With bucket hooks (post-accumulate hooks) installed the
        # gradients keep going through autograd.
        if self.world_size > 1 and torch.cuda.is_available():
            from . import ops
            # cooperative kernels start only when all their CTAs fit: leave 32 SMs to the NCCL kernels that run beside backward
            ops.coop_cta_limit = max(32, torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count - 32)
        direct = not (self.world_size > 1 and self.overlap_grad_sync)
        for o in self.optimizers:
            if hasattr(o, 'flat_grad'):
                for g in o.param_groups:
                    for p in g['params']:
                        p._vqb_direct_grad = direct

    def _install_bucket_hooks(self, opt) -> None:
        """One post-accumulate hook per parameter: when the last gradient of a bucket has landed in the flat buffer, its
        all-reduce is enqueued (async_op: NCCL's stream waits for the kernels issued so far and runs beside backward)."""
        states = []
        for (a, b, params) in opt.buckets(self.bucket_bytes):
            st = {'range': (a, b), 'n': len(params), 'pending': len(params), 'handle': None}
            states.append(st)
            for p in params:
                def hook(_p, st=st, opt=opt):
                    st['pending'] -= 1
                    if st['pending'] == 0 and st['handle'] is None:
                        a_, b_ = st['range']
                        st['handle'] = dist.all_reduce(opt.flat_grad[a_:b_], op=dist.ReduceOp.SUM, async_op=True)
                p.register_post_accumulate_grad_hook(hook)
        self._buckets[id(opt)] = states

    def _allreduce_stats(self, *tensors: torch.Tensor) -> None:
        for t in tensors:
            if t is not None:
                dist.all_reduce(t, op=dist.ReduceOp.SUM)

    def sync_gradients(self, optimizer=None) -> None:
        if self.world_size <= 1:
            return
        for o in (self.optimizers if optimizer is None else [optimizer]):
            flat = getattr(o, 'flat_grad', None)
            states = self._buckets.get(id(o))
            if flat is not None and states is not None:
                # buckets whose reduction is already in flight: wait; the others (a tensor received no gradient): reduce now
                for st in states:
                    if st['handle'] is None:
                        a_, b_ = st['range']
                        st['handle'] = dist.all_reduce(flat[a_:b_], op=dist.ReduceOp.SUM, async_op=True)
                for st in states:
                    st['handle'].wait()
                    st['handle'], st['pending'] = None, st['n']
            elif flat is not None:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM)       # one NCCL call per optimizer; averaged in vqb_adamw
            else:
                for g in o.param_groups:
                    for p in g['params']:
                        if p.grad is not None:
                            dist.all_reduce(p.grad, op=dist.ReduceOp.SUM)
                            p.grad.div_(self.world_size)

    # ---- one optimisation step (what Lightning's fit loop does around training_step) -------------------
    def run_step(self, batch, batch_index: int):
        if isinstance(batch, tuple):              # (images, labels) of the FFCV loaders: the step reads the images only (model.py:237)
            batch = batch[0]
        if self.cuda_graph and torch.is_tensor(batch) and batch.is_cuda:
            return self._run_step_graphed(batch, batch_index)
        return self._run_step_eager(batch, batch_index)

    def _run_step_graphed(self, batch, batch_index: int):
        m = self.model
        q = getattr(m, 'quantizer', None)
        if q is not None and hasattr(q, 'enable_device_consts') and not q.device_consts:
            q.enable_device_consts(batch.device)
        key = (tuple(batch.shape), batch.dtype, m.graph_variant(batch_index) if hasattr(m, 'graph_variant') else ())
        st = self._graphs.setdefault(key, {'calls': 0, 'graph': None})
        st['calls'] += 1
        if st['graph'] is None and st['calls'] <= self.graph_warmup:
            return self._run_step_eager(batch, batch_index)
        from . import ops
        if st['graph'] is None:
            # capture: the step runs once more under stream capture (nothing executes), on a static copy of the batch
            st['input'] = batch.clone()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            m.on_train_batch_start(st['input'], batch_index)              # host scalars of THIS step (pinned buffers are read at replay)
            counts0 = [o.step_count for o in self.optimizers if hasattr(o, 'host_step_update')]
            ops.take_capture_refs()
            with torch.cuda.graph(g):
                st['loss'] = self._step_body(st['input'], batch_index)
            st['refs'] = ops.take_capture_refs()          # cached tensors the graph addresses (packed-weight table, sources, ...)
            # the capture ran the host-side bookkeeping of one step without executing it: undo, the replay below performs it
            for o, c in zip([o for o in self.optimizers if hasattr(o, 'host_step_update')], counts0):
                o.step_count = c
            st['graph'] = g
            st['logged'] = dict(m.logged)
        else:
            st['input'].copy_(batch, non_blocking=True)
            m.on_train_batch_start(st['input'], batch_index)
        for o in self._optimizers_stepped(key):
            o.host_step_update()
        st['graph'].replay()
        ops.bump_weights_epoch()                  # weights changed behind the Python-side caches (packed layouts, codebook copies)
        m.logged.update(st['logged'])
        return st['loss']

    def _optimizers_stepped(self, key):
        """optimizers whose step() is part of this variant's graph (the discriminator's is absent before its start epoch)"""
        variant = key[2]
        if len(self.optimizers) > 1 and len(variant) == 2 and not variant[0]:
            return [o for o in self.optimizers[:1] if hasattr(o, 'host_step_update')]
        return [o for o in self.optimizers if hasattr(o, 'host_step_update')]

    def _run_step_eager(self, batch, batch_index: int):
        m = self.model
        m.on_train_batch_start(batch, batch_index)
        return self._step_body(batch, batch_index)

    def _step_body(self, batch, batch_index: int):
        from . import ops
        dev = batch.device if torch.is_tensor(batch) and batch.is_cuda else None
        if dev is None:
            return self._step_body_inner(batch, batch_index)
        with ops.zero_arena.step(dev):                  # one memset instead of ~150 small zero fills per step
            return self._step_body_inner(batch, batch_index)

    def _step_body_inner(self, batch, batch_index: int):
        m = self.model
        if m.automatic_optimization:
            opt = self.optimizers[0]
            opt.zero_grad()
            loss = m.training_step(batch, batch_index)
            loss.backward()
            self.sync_gradients()
            opt.step()
        else:
            loss = m.training_step(batch, batch_index)
        # detached, as Lightning's loop hands losses on: a caller that keeps the returned loss (fit() below does, across the next
        # step) must not keep this step's autograd graph -- and its AccumulateGrad nodes, bound to the stream they were created
        # on -- alive into the next step, which may run under stream capture
        return loss.detach()

    @staticmethod
    def _device_batches(batches: Iterable, model: nn.Module):
        """Lightning's batch transfer (the reference feeds host batches from a DataLoader(pin_memory=True), vqvae/train.py:121-142):
        host batches reach the model's device through the DevicePrefetcher, one step ahead; device batches pass through."""
        import itertools
        p = next(model.parameters(), None)
        it = iter(batches)
        first = next(it, None)
        if first is None:
            return iter(())
        rest = itertools.chain([first], it)
        ts = DevicePrefetcher._tensors(first)
        if p is not None and p.is_cuda and ts and not ts[0].is_cuda:
            return DevicePrefetcher(rest, p.device)
        return rest

    def fit(self, model: LightningModule, batches: Iterable, steps_per_epoch: Optional[int] = None):
        if self.model is not model:
            self.attach(model)
        if steps_per_epoch is None:
            steps_per_epoch = len(batches)
        self.num_training_batches = steps_per_epoch
        model.train()
        model.on_train_start()
        loss = None
        for epoch in range(self.max_epochs):
            model.current_epoch = epoch
            for i, batch in enumerate(self._device_batches(batches, model)):
                if i >= steps_per_epoch:
                    break
                loss = self.run_step(batch, i)
            model.on_train_epoch_end()
        model.on_train_end()
        return loss

    # ---- evaluation loops (what pl.Trainer.validate / .test do around the module's hooks) -----------------------
    def _eval_loop(self, model: LightningModule, batches: Iterable, step, begin, end) -> Dict[str, Any]:
        if model.trainer is None:
            model.trainer = self                      # validation_step reads trainer.num_training_batches (model.py:341)
        was_training = model.training
        model.eval()
        try:
            if begin is not None:
                begin()
            with torch.no_grad():
                for i, batch in enumerate(self._device_batches(batches, model)):
                    step(batch, i)
            if end is not None:
                end()
        finally:
            model.train(was_training)
        return dict(model.logged)

    def validate(self, model: LightningModule, batches: Iterable) -> Dict[str, Any]:
        """validation_step over `batches` + on_validation_epoch_end (model.py:309-370); returns the logged validation/* scalars"""
        return self._eval_loop(model, batches, model.validation_step, None, getattr(model, 'on_validation_epoch_end', None))

    def test(self, model: LightningModule, batches: Iterable) -> Dict[str, Any]:
        """on_test_epoch_start / test_step / on_test_epoch_end (model.py:491-562, driven by vqvae/evaluate.py); returns the
        logged metrics (mse, psnr, ssim, used_codebook, perplexity [, rfid])"""
        return self._eval_loop(model, batches, model.test_step, getattr(model, 'on_test_epoch_start', None),
                               getattr(model, 'on_test_epoch_end', None))
