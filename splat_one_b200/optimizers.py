"""`SelectiveAdam` — drop-in for `gsplat.optimizers.SelectiveAdam`
(/root/reference/submodules/gsplat/gsplat/optimizers/selective_adam.py:6-89), the optimizer
splat_one selects with `visible_adam` (utils/gsplat_utils/gsplat_trainer.py:269-270, 719-730):
Adam whose moments and parameters move only for the Gaussians visible in the current step.
The update itself is one sm_100a kernel (csrc/optim.cu) through the C ABI."""
import torch

from .wrapper import selective_adam_update


class SelectiveAdam(torch.optim.Adam):
    """Same constructor and `step(visibility)` contract as the reference: one tensor per
    param group, `visibility` a bool mask [N] over the leading dimension, no bias correction
    (CS/adam.cu:33-40)."""

    def __init__(self, params, eps, betas):
        super().__init__(params=params, eps=eps, betas=betas)

    @torch.no_grad()
    def step(self, visibility):
        N = visibility.numel()
        for group in self.param_groups:
            lr = group["lr"]
            eps = group["eps"]
            beta1, beta2 = group["betas"]
            assert len(group["params"]) == 1, "more than one tensor in group"
            param = group["params"][0]
            if param.grad is None:
                continue
            state = self.state[param]
            if len(state) == 0:
                state["step"] = torch.tensor(0.0, dtype=torch.float32)
                state["exp_avg"] = torch.zeros_like(param, memory_format=torch.preserve_format)
                state["exp_avg_sq"] = torch.zeros_like(param, memory_format=torch.preserve_format)
            M = param.numel() // N
            selective_adam_update(param, param.grad, state["exp_avg"], state["exp_avg_sq"], visibility, lr, beta1,
                                  beta2, eps, N, M)
