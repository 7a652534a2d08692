"""`SelectiveAdam` — drop-in for `gsplat.optimizers.SelectiveAdam`
(/root/reference/submodules/gsplat/gsplat/optimizers/selective_adam.py:6-89), the optimizer
splat_one selects with `visible_adam` (utils/gsplat_utils/gsplat_trainer.py:269-270, 719-730):
Adam whose moments and parameters move only for the Gaussians visible in the current step.
The update itself is one sm_100a kernel (csrc/optim.cu) through the C ABI.

What is kept from the reference because it is API contract, not implementation: the constructor
signature, `step(visibility)`, one tensor per param group, and the per-parameter state keys
(`step`, `exp_avg`, `exp_avg_sq`) that `torch.optim.Adam.state_dict()` / checkpoints of the
reference trainer carry.  `step` stays 0: the reference kernel applies no bias correction
(CS/adam.cu:33-40)."""
from typing import Dict

import torch
from torch import Tensor

from .wrapper import selective_adam_update

_STATE_KEYS = ("step", "exp_avg", "exp_avg_sq")


def _moments(state: Dict[str, Tensor], param: Tensor) -> Dict[str, Tensor]:
    """First/second moment buffers of `param`, created on first use with the reference's keys."""
    if not state:
        state[_STATE_KEYS[0]] = torch.tensor(0.0, dtype=torch.float32)
        for key in _STATE_KEYS[1:]:
            state[key] = torch.zeros_like(param, memory_format=torch.preserve_format)
    return state


class SelectiveAdam(torch.optim.Adam):
    """`visibility`: bool mask [N] over the leading dimension of every parameter; rows whose mask is
    False keep parameter AND moments untouched (selective_adam.py:9-13)."""

    def __init__(self, params, eps, betas):
        super().__init__(params=params, eps=eps, betas=betas)

    @torch.no_grad()
    def step(self, visibility: Tensor):
        n_rows = visibility.numel()
        for group in self.param_groups:
            if len(group["params"]) != 1:
                raise AssertionError("more than one tensor in group")  # same contract as the reference
            (param,) = group["params"]
            grad = param.grad
            if grad is None:
                continue
            st = _moments(self.state[param], param)
            b1, b2 = group["betas"]
            selective_adam_update(param, grad, st["exp_avg"], st["exp_avg_sq"], visibility, group["lr"], b1, b2,
                                  group["eps"], n_rows, param.numel() // n_rows)
