"""Host-side scalar schedules.

The reference imports LinearScheduler / CosineScheduler / LinearCosineScheduler from the external, un-vendored,
un-pinned `scheduling_utils.schedulers_cpp` package (vqvae/model.py:6; environment.yml:32).  PARITY UNPINNED: the
semantics below are inferred from the call sites vqvae/model.py:175-200,210-224 -- `.step(i) -> float`,
`.destroy()`.  Host scalars only; nothing here touches the GPU.
"""
from __future__ import annotations

import math


class _Base:
    def destroy(self) -> None:      # the C++ objects need explicit destruction (model.py:305-307); a no-op here
        pass


class LinearScheduler(_Base):
    def __init__(self, start_step: int, stop_step: int, start_value: float, stop_value: float):
        self.a, self.b, self.va, self.vb = int(start_step), int(stop_step), float(start_value), float(stop_value)

    def step(self, i: int) -> float:
        if i <= self.a:
            return self.va
        if i >= self.b:
            return self.vb
        t = (i - self.a) / float(self.b - self.a)
        return self.va + (self.vb - self.va) * t


class CosineScheduler(_Base):
    def __init__(self, start_step: int, stop_step: int, start_value: float, stop_value: float):
        self.a, self.b, self.va, self.vb = int(start_step), int(stop_step), float(start_value), float(stop_value)

    def step(self, i: int) -> float:
        if i <= self.a:
            return self.va
        if i >= self.b:
            return self.vb
        t = (i - self.a) / float(self.b - self.a)
        return self.vb + (self.va - self.vb) * 0.5 * (1.0 + math.cos(math.pi * t))


class LinearCosineScheduler(_Base):
    """linear warm-up from ~0 to start_value until th_step, then cosine to stop_value at stop_step (model.py:175)."""

    def __init__(self, start_step: int, stop_step: int, start_value: float, stop_value: float, th_step: int):
        self.th = int(th_step)
        self.lin = LinearScheduler(start_step, th_step, 1e-20, start_value)
        self.cos = CosineScheduler(th_step, stop_step, start_value, stop_value)

    def step(self, i: int) -> float:
        return self.lin.step(i) if i < self.th else self.cos.step(i)
