"""LTXScheduler -- host-side mirror of Scheduler/LTXScheduler.swift (sigma schedules are host float32 arithmetic in the
reference too; the Euler update itself runs on the GPU inside ltx_guided_euler_step / ltx_denoise_step)."""
from __future__ import annotations

from typing import List, Optional

import numpy as np

# Scheduler/LTXScheduler.swift:18-36
DISTILLED_SIGMA_VALUES = [1.0, 0.99375, 0.9875, 0.98125, 0.975, 0.909375, 0.725, 0.421875, 0.0]
STAGE_2_DISTILLED_SIGMA_VALUES = [0.909375, 0.725, 0.421875, 0.0]
BASE_SHIFT_ANCHOR = 1024
MAX_SHIFT_ANCHOR = 4096

_f = np.float32


class LTXScheduler:
    """Same surface as the Swift class: set_timesteps / set_custom_sigmas / sigmas / step_index bookkeeping."""

    def __init__(self, num_train_timesteps: int = 1000, is_distilled: bool = False):
        self.num_train_timesteps = num_train_timesteps
        self.is_distilled = is_distilled
        self.sigmas: List[float] = []
        self.step_index = 0

    @staticmethod
    def _mu(token_count: int, max_shift: float, base_shift: float) -> np.float32:
        x1, x2 = _f(BASE_SHIFT_ANCHOR), _f(MAX_SHIFT_ANCHOR)
        mm = (_f(max_shift) - _f(base_shift)) / (x2 - x1)
        b = _f(base_shift) - mm * x1
        return _f(token_count) * mm + b

    def set_timesteps(self, num_steps: int, distilled: bool = False, latent_token_count: Optional[int] = None,
                      max_shift: float = 2.05, base_shift: float = 0.95, stretch: bool = True, terminal: float = 0.1):
        """Scheduler/LTXScheduler.swift:74-182 (all arithmetic in float32 like Swift `Float`)."""
        self.is_distilled = distilled
        self.step_index = 0
        one = _f(1.0)
        if distilled:
            s = np.array([v for v in DISTILLED_SIGMA_VALUES if v > 0], dtype=np.float32)
            if latent_token_count is not None:
                e = np.exp(self._mu(min(latent_token_count, MAX_SHIFT_ANCHOR), max_shift, base_shift), dtype=np.float32)
                keep = (s == 0) | (s == 1.0)
                with np.errstate(divide="ignore"):
                    shifted = (e / (e + (one / s - one))).astype(np.float32)
                s = np.where(keep, s, shifted).astype(np.float32)
                if stretch:
                    last = one - s[-1]
                    if last > 0:
                        scale = last / (one - _f(terminal))
                        s = np.where(s == 0, _f(0), one - (one - s) / scale).astype(np.float32)
            self.sigmas = [float(v) for v in s] + [0.0]
            return self.sigmas
        tok = min(latent_token_count if latent_token_count is not None else MAX_SHIFT_ANCHOR, MAX_SHIFT_ANCHOR)
        s = (one - np.arange(num_steps + 1, dtype=np.float32) / _f(num_steps)).astype(np.float32)
        e = np.exp(self._mu(tok, max_shift, base_shift), dtype=np.float32)
        safe = np.where(s == 0, one, s)
        s = np.where(s == 0, _f(0), e / (e + (one / safe - one))).astype(np.float32)
        if stretch and num_steps > 0:
            om = one - s
            scale = om[num_steps - 1] / (one - _f(terminal))
            s = np.where(s == 0, _f(0), one - om / scale).astype(np.float32)
        self.sigmas = [float(v) for v in s]
        return self.sigmas

    def set_custom_sigmas(self, custom: List[float]):
        """:184-201"""
        if not custom:
            return
        s = list(custom)
        if s[-1] != 0.0:
            s.append(0.0)
        self.sigmas, self.step_index, self.is_distilled = s, 0, False

    @property
    def current_sigma(self) -> float:
        return self.sigmas[self.step_index] if self.step_index < len(self.sigmas) else 0.0

    @property
    def initial_sigma(self) -> float:
        return self.sigmas[0] if self.sigmas else 1.0

    @property
    def total_steps(self) -> int:
        return max(0, len(self.sigmas) - 1)

    @property
    def remaining_steps(self) -> int:
        return max(0, len(self.sigmas) - 1 - self.step_index)

    def reset(self):
        self.step_index = 0


def get_sigma_schedule(num_steps: int, distilled: bool = False, latent_token_count: Optional[int] = None) -> List[float]:
    """:372-385"""
    if distilled:
        return list(DISTILLED_SIGMA_VALUES)
    s = LTXScheduler()
    return s.set_timesteps(num_steps, False, latent_token_count)
